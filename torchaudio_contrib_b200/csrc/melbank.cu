// K2: |.|^p + dense (num_bins x num_bands) contraction on tcgen05 tensor cores + dB epilogue.
// Replaces complex_norm (functional.py:126-128), apply_filterbank's torch.matmul (:183) and
// amplitude_to_db (:291-296) with one kernel.
//
// GEMM view:  D[frame, band] = sum_bin P[frame, bin] * FB[bin, band]
//   M = 128 frames per CTA  (TMEM lanes; lane == frame makes the epilogue stores coalesced along
//                            the contiguous time axis of the (n_seq, bands, frames) output)
//   N = bands               (TMEM columns, <= 128 per CTA; more bands -> blockIdx.y)
//   K = bins, consumed in slices of 32 (one 128-byte swizzle row of tf32)
// Precision: inputs are split v = hi + lo with hi = top 19 bits (exactly a tf32), lo = v - hi, and
//   D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi   ("3xTF32", fp32 accumulate in TMEM) -> ~2^-21 relative,
//   needed for the 1e-4 parity bar; single-pass tf32 (2^-11) is not enough.
// Sparsity: the host-built plan records, per K slice, the range of bands with a non-zero weight;
//   only that [band_lo, band_lo + n) block is stored, copied and multiplied (UMMA N = n).  A mel
//   matrix touches 16-32 of 128 bands per slice; a dense matrix degenerates to the full range.
//
// Warp roles (320 threads): warps 0-7 produce the A operand (power tile from smem, or |.|^p of values read
// from global memory -> tf32 hi/lo split -> tcgen05.st into the stage's TMEM columns) and later run the
// epilogue (TMEM -> registers -> dB -> global); warp 8 issues the MMAs (A from TMEM, B from smem) through one
// elected lane; warp 9 streams the plan blocks (and the power tiles) with 1-D bulk async copies.  Up to six
// stages, mbarrier full/ready/empty handshakes, tcgen05.commit releases a stage when its MMAs retired.
#include <stdlib.h>

#include "bandplan.cuh"
#include "tac_common.cuh"

namespace tac {

constexpr int kMbRows = 128;
constexpr int kMbBK = 32;
constexpr int kMbMaxStages = 8;
constexpr int kMbMaxChunks = 512;                    // slice table held in shared memory: up to 16384 bins
constexpr int kMbProducerWarps = 8;
constexpr int kMbProducerThreads = kMbProducerWarps * 32;
constexpr int kMbThreads = kMbProducerThreads + 64;
constexpr int kMbBandBlock = 128;
constexpr int kMbTileBytes = kMbRows * kMbBK * 4;          // 16 KB: one operand tile
// Stage = {A_hi, A_lo, B_hi, B_lo}.  Slots are sized at launch: A by the tile height, B by the widest
// band block of the plan, so a mel matrix (<= 48 bands per slice) gets 6-8 stages in flight instead of 3.
constexpr size_t kMbStageBudget = 192 * 1024;
constexpr size_t kMbSmemBytes = kMbStageBudget + kMbTileBytes + 1024;   // + over-read slack (UMMA M = 128) + alignment
constexpr uint32_t kPlanMagic = 0x7ac0fb01u;

struct FbPlanHeader {
  uint32_t magic;
  int32_t n_bins, n_bands, n_chunks, n_bblocks;
  int32_t max_n;      // widest block (bands) over all K slices
  int32_t band_off;   // byte offset of the band plan (bandplan.cuh) inside this blob, 0: the matrix has no such form
  int32_t band_cmax;  // its list length (entries per band)
  int32_t range_off;  // byte offset of the range plan (bandplan.cuh) inside this blob, 0: ranges too long (dense matrix)
  int32_t range_bytes;
  int32_t reserved[2];
};
struct FbPlanChunk {
  int32_t band_lo;    // first band of the block, relative to the band block, multiple of 16
  int32_t n;          // bands in the block, multiple of 16, 0 = nothing to do for this K slice
  int32_t blob_off;   // byte offset of the hi image from the start of the plan (lo image follows)
  int32_t reserved;
};

enum MelbankSource {
  SRC_TILES = 0,           // power tiles written by the STFT kernel (stft_params.cuh: power_tile_index)
  SRC_PUBLIC_REAL = 1,     // (n_seq, bins, frames)     reference layout of a magnitude / power spectrogram
  SRC_PUBLIC_COMPLEX = 2   // (n_seq, bins, frames, 2)  reference layout of a complex spectrogram
};

struct MelbankParams {
  const float* src;
  const unsigned char* plan;
  float* out;
  int64_t rows;          // frames handled by this launch
  int64_t g_base;        // flattened frame index (seq * frames + t) of row 0
  int64_t frames;        // frames per sequence
  int power_mode;        // complex input: 2 -> re^2+im^2, 1 -> sqrt, 0 -> pow(., power/2)
  float half_power;
  int n_bins, n_bands;
  int rows_per_tile;     // <= 128; SRC_TILES: the tile height the STFT kernel wrote (multiple of 8)
  int to_db;
  float amin, log10_ref;
  int debug;             // TAC_MB_DEBUG bit field for timing experiments (0 in production)
};

// offset of element (row r, 16-byte column c16) inside a 128B-swizzled K-major tile
__device__ __forceinline__ uint32_t swz_off(int r, int c16) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  // start address (>>4) | LBO = 1 (unused for swizzled K-major) | SBO = 1024 B between 8-row groups
  // | descriptor version 1 (sm_100) | layout type 2 = SWIZZLE_128B
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
  // c = f32 (1 << 4), a = b = tf32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M = 128 -> 8 at bit 24
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
}

__device__ __forceinline__ void split_tf32(const float4 v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
}

__device__ __forceinline__ float power_of(float re, float im, float half_power, int mode) {
  const float s = fmaf(re, re, im * im);
  if (mode == 2) return s;
  if (mode == 1) return sqrtf(s);
  return s > 0.0f ? exp2f(half_power * __log2f(s)) : (half_power == 0.0f ? 1.0f : 0.0f);
}

__device__ long long g_mb_trace[4][80];     // TAC_MB_DEBUG & 8: clock64 stamps of block 0 (loader, producer, mma, misc)
#define MB_TRACE(role, idx) do { if ((p.debug & 8) && blockIdx.x == 0 && blockIdx.y == 0 && (idx) < 80) g_mb_trace[role][idx] = clock64(); } while (0)

// TMEM map (512 columns allocated): [0, 128) accumulator D (lane = frame, column = band);
// [128 + 128 s, 128 + 128 s + 128): A operand of pipeline stage s = two 32-bin slices, each 32 columns of
// hi followed by 32 columns of lo (lane = frame, column = bin).  Feeding A from tensor memory matters
// here: with N = 16..48 bands an MMA does little math per operand byte, and an A operand read from shared
// memory (128 rows x 32 B per instruction, three times per k-step) made each UTCHMMA cost ~55 cycles of
// shared-memory bandwidth (measured); from TMEM only the small B block is read from shared memory.
// A stage carries TWO slices (64 bins) because the per-stage handshake chain (bulk-copy latency ~1100
// cycles, barrier wake-ups ~300) -- not the tensor pipe -- bounds the loop; fewer, larger stages halve it.
constexpr uint32_t kMbTmemCols = 512;
constexpr uint32_t kMbTmemA0 = 128;
constexpr int kMbTmemStages = 3;
constexpr int kMbGroup = 2;                  // slices per stage

template <int SRC>
__global__ void __launch_bounds__(kMbThreads, 1) melbank_kernel(const MelbankParams p) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  __shared__ uint64_t s_full[kMbMaxStages], s_a_ready[kMbMaxStages], s_empty[kMbMaxStages], s_accum, s_zeroed;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) FbPlanChunk s_chunks[kMbMaxChunks + kMbGroup];   // slice table (global reads cost ~500 cycles per step)

  // 1024-byte aligned stage buffers (swizzle atoms must not straddle 1 KB boundaries)
  unsigned char* stage0 = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const FbPlanHeader* hdr = reinterpret_cast<const FbPlanHeader*>(p.plan);
  const int n_chunks = hdr->n_chunks;
  const int n_groups = (n_chunks + kMbGroup - 1) / kMbGroup;
  {
    const int4* src = reinterpret_cast<const int4*>(p.plan + sizeof(FbPlanHeader)) + (size_t)blockIdx.y * n_chunks;
    for (int i = threadIdx.x; i < n_groups * kMbGroup; i += kMbThreads)
      reinterpret_cast<int4*>(s_chunks)[i] = (i < n_chunks) ? __ldg(src + i) : make_int4(0, 0, 0, 0);
  }
  const FbPlanChunk* chunks = s_chunks;

  // persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ... ; the stage barriers simply keep counting
  // across tiles, so the loader streams the next tile's first stages while this tile's epilogue stores drain
  const int tile_rows = p.rows_per_tile;
  const int64_t n_tiles = (p.rows + tile_rows - 1) / tile_rows;
  const uint32_t a_tile_bytes = (uint32_t)tile_rows * 128u;
  // stage = {raw A tiles of the two slices (SRC_TILES only), B0 = [hi | lo], B1 = [hi | lo]}
  const uint32_t a_slot = (SRC == SRC_TILES) ? ((kMbGroup * a_tile_bytes + 1023u) & ~1023u) : 0u;
  const uint32_t b_slot = ((uint32_t)max(hdr->max_n, 16) * 256u + 1023u) & ~1023u;      // hi + lo of one slice
  const uint32_t stage_bytes = a_slot + kMbGroup * b_slot;
  const int n_stages = max(2, min(kMbTmemStages, (int)(kMbStageBudget / stage_bytes)));

  if (warp == kMbProducerWarps && lane == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_a_ready[s], kMbProducerWarps);
      mbar_init(&s_empty[s], 1);
    }
    mbar_init(&s_accum, 1);
    mbar_init(&s_zeroed, kMbProducerWarps + 1);
    fence_mbar_init();
  }
  if (warp == kMbProducerWarps + 1) {
    tmem_alloc(&s_tmem, kMbTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (tid == 0) MB_TRACE(3, 0);

  // Per tile the accumulator is zeroed first (blocks of different K slices touch different column ranges, so
  // every MMA accumulates -- there is no single "first" MMA per column); producers and the MMA warp meet on an
  // mbarrier for that (one arrival per warp, phase = tile parity), the loader warp never waits for it.  (Round 1 used
  // named barrier 1 with a partial thread count here, which compute-sanitizer's synccheck reports as divergence.)
  uint32_t zeroed_parity = 0;
  auto accumulator_ready = [&]() {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_zeroed);
    mbar_wait(&s_zeroed, zeroed_parity);
    zeroed_parity ^= 1u;
    tc_fence_after();
  };

  // a group (stage) is active when at least one of its slices has a non-empty block
  auto group_active = [&](int g) { return (chunks[kMbGroup * g].n | chunks[kMbGroup * g + 1].n) != 0; };
  auto next_active = [&](int g) {
    while (g < n_groups && !group_active(g)) ++g;
    return g;
  };

  if (warp < kMbProducerWarps) {
    // =========================== A producers ====================================================
    // thread = (frame row r = 32 * (warp % 4) + lane, bin half kh = warp / 4): it owns 16 consecutive bins
    // of its row in every slice, splits them into tf32 hi / lo and stores both to the stage's TMEM columns
    // (a warp can only touch the TMEM lane quarter warp % 4, which is why rows map to lanes this way).
    const int q = warp & 3, kh = warp >> 2;
    const int r = 32 * q + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(32 * q) << 16);
    int it = 0;
    uint32_t tile_parity = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, tile_parity ^= 1u) {
    const int64_t row0 = tile * tile_rows;
    const int valid = (int)min((int64_t)tile_rows, p.rows - row0);
    const int64_t g_first = p.g_base + row0;
    {
      const uint32_t t0 = t_lane + (uint32_t)(64 * kh);       // the region this warp read in its last epilogue
#pragma unroll
      for (int j = 0; j < 4; ++j) tmem_zero16(t0 + 16 * j);
      tc_wait_st();
    }
    accumulator_ready();

    auto publish = [&](int s, const float (&v)[16 * kMbGroup]) {
#pragma unroll
      for (int j = 0; j < kMbGroup; ++j) {
        float hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          hi[i] = __uint_as_float(__float_as_uint(v[16 * j + i]) & 0xFFFFE000u);
          lo[i] = v[16 * j + i] - hi[i];
        }
        const uint32_t ta = t_lane + kMbTmemA0 + 128u * (uint32_t)s + 64u * (uint32_t)j + 16u * (uint32_t)kh;
        tmem_st16(ta, hi);
        tmem_st16(ta + 32u, lo);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_a_ready[s]);
    };

    if constexpr (SRC == SRC_TILES) {
      // the STFT kernel wrote the rows in the 128B-swizzled tile layout; the loader warp bulk-copies the two
      // (tile_rows x 32) blocks of this stage into shared memory; reading a row's 16-byte units through the
      // swizzle is bank-conflict free (8 consecutive rows hit 8 different 16-byte columns)
      for (int g = next_active(0); g < n_groups; g = next_active(g + 1), ++it) {
        const int s = it % n_stages;
        const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
        mbar_wait(&s_full[s], ph);
        if (tid == 0) MB_TRACE(1, 2 * it);
        const unsigned char* raw = stage0 + (size_t)s * stage_bytes;
        float v[16 * kMbGroup];
        if (!(p.debug & 1)) {
#pragma unroll
          for (int j = 0; j < kMbGroup; ++j) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 t4 = *reinterpret_cast<const float4*>(raw + j * a_tile_bytes + swz_off(r, 4 * kh + i));
              v[16 * j + 4 * i] = t4.x; v[16 * j + 4 * i + 1] = t4.y; v[16 * j + 4 * i + 2] = t4.z; v[16 * j + 4 * i + 3] = t4.w;
            }
          }
        }
        publish(s, v);       // TMEM columns of stage s are free: `full` implies the loader saw `empty`
        if (tid == 0) MB_TRACE(1, 2 * it + 1);
      }
    } else {
      // reference layouts: values come straight from global memory (lanes run along the contiguous time
      // axis); two register buffers alternate so the loads of group g+1 fly while group g is published
      const bool ok = r < valid;
      const int64_t gf = g_first + r;
      const int64_t seq = gf / p.frames, t = gf - seq * p.frames;
      const int64_t base = seq * p.n_bins * p.frames + t;
      auto load_group = [&](int g, float (&v)[16 * kMbGroup]) {
#pragma unroll
        for (int j = 0; j < kMbGroup; ++j) {
          const int k0 = (kMbGroup * g + j) * kMbBK + 16 * kh;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = k0 + i;
            float val = 0.0f;
            if (ok && k < p.n_bins) {
              const int64_t idx = base + (int64_t)k * p.frames;
              if constexpr (SRC == SRC_PUBLIC_COMPLEX) {
                const float2 z = __ldg(reinterpret_cast<const float2*>(p.src) + idx);
                val = power_of(z.x, z.y, p.half_power, p.power_mode);
              } else {
                val = ldg_stream_f1(p.src + idx);
              }
            }
            v[16 * j + i] = val;
          }
        }
      };
      auto publish_next = [&](const float (&v)[16 * kMbGroup]) {
        const int s = it % n_stages;
        const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
        mbar_wait(&s_empty[s], ph ^ 1u);          // MMAs of the stage's previous use have retired
        publish(s, v);
        ++it;
      };
      float buf_a[16 * kMbGroup], buf_b[16 * kMbGroup];
      int g0 = next_active(0);
      if (g0 < n_groups) load_group(g0, buf_a);
      while (g0 < n_groups) {
        const int g1 = next_active(g0 + 1);
        if (g1 < n_groups) load_group(g1, buf_b);
        publish_next(buf_a);
        if (g1 >= n_groups) break;
        const int g2 = next_active(g1 + 1);
        if (g2 < n_groups) load_group(g2, buf_a);
        publish_next(buf_b);
        g0 = g2;
      }
    }

    // =========================== epilogue =======================================================
    mbar_wait(&s_accum, tile_parity);
    tc_fence_after();
    if (tid == 0) MB_TRACE(3, 2);
    const bool ok = r < valid;
    const int64_t gf = g_first + r;
    const int64_t seq = gf / p.frames, t = gf - seq * p.frames;
    const int band0 = blockIdx.y * kMbBandBlock + 64 * kh;
    float* outp = p.out + (seq * p.n_bands + band0) * p.frames + t;
    const int n_here = min(64, p.n_bands - band0);           // bands this warp owns (<= 0: none)
    const bool to_db = p.to_db != 0;
    uint32_t acc[4][16];
#pragma unroll
    for (int j = 0; j < 4; ++j) tmem_ld16_nowait(t_lane + (uint32_t)(64 * kh + 16 * j), acc[j]);
    tc_wait_ld();
    if (tid == 0) MB_TRACE(3, 1);
    if (ok) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (16 * j + i < n_here) {
            float v = __uint_as_float(acc[j][i]);
            if (to_db) {
              float s2 = v * v;
              s2 = (s2 < p.amin) ? p.amin : s2;
              v = 10.0f * (log10f(s2) - p.log10_ref);
            }
            __stcs(outp, v);
          }
          outp += p.frames;
        }
      }
    }
    }   // tile loop
  } else if (warp == kMbProducerWarps) {
    // =========================== MMA issuer =====================================================
    // the whole warp walks the loop (converged), one elected lane issues
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    accumulator_ready();
    for (int g = next_active(0); g < n_groups; g = next_active(g + 1), ++it) {
      const int s = it % n_stages;
      const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
      if constexpr (SRC != SRC_TILES) mbar_wait(&s_full[s], ph);      // tiles: a_ready already implies full
      mbar_wait(&s_a_ready[s], ph);
      tc_fence_after();
      if (lane == 0) MB_TRACE(2, 2 * it);
      const uint32_t st = smem_u32(stage0 + (size_t)s * stage_bytes) + a_slot;
      if (elect_one()) {
        if (!(p.debug & 2)) {
#pragma unroll
          for (int j = 0; j < kMbGroup; ++j) {
            const int n = chunks[kMbGroup * g + j].n;
            if (n == 0) continue;
            const uint64_t b_hi = umma_desc_sw128(st + j * b_slot);
            const uint64_t b_lo = umma_desc_sw128(st + j * b_slot + (uint32_t)n * 128u);
            const uint32_t a_hi = tmem + kMbTmemA0 + 128u * (uint32_t)s + 64u * (uint32_t)j, a_lo = a_hi + 32u;
            const uint32_t idesc = umma_idesc_tf32(n);
            const uint32_t d = tmem + (uint32_t)chunks[kMbGroup * g + j].band_lo;
#pragma unroll
            for (int ks = 0; ks < kMbBK / 8; ++ks) {    // UMMA K = 8: 8 TMEM columns of A, 32 bytes of each B row
              tc_mma_tf32_ts(d, a_lo + 8 * ks, b_hi + 2 * ks, idesc, 1u);
              tc_mma_tf32_ts(d, a_hi + 8 * ks, b_lo + 2 * ks, idesc, 1u);
              tc_mma_tf32_ts(d, a_hi + 8 * ks, b_hi + 2 * ks, idesc, 1u);
            }
          }
        }
        tc_commit(&s_empty[s]);
      }
      __syncwarp();
      if (lane == 0) MB_TRACE(2, 2 * it + 1);
    }
    if (elect_one()) tc_commit(&s_accum);
    __syncwarp();
    }   // tile loop
  } else {
    // =========================== bulk-copy loader ===============================================
    // power tiles are read once -> evict-first after this read (measured: 112 MB of DRAM traffic per config-2 step,
    // 131 MB when they were kept evict-last, 173 MB without hints); the plan is read by every CTA -> evict-last
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int g = next_active(0); g < n_groups; g = next_active(g + 1), ++it) {
      const int s = it % n_stages;
      const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
      mbar_wait(&s_empty[s], ph ^ 1u);
      if (lane == 0) MB_TRACE(0, 2 * it);
      unsigned char* st = stage0 + (size_t)s * stage_bytes;
      const int c0 = kMbGroup * g;
      const int slices = min(kMbGroup, n_chunks - c0);
      const uint32_t b0 = (uint32_t)chunks[c0].n * 256u, b1 = (uint32_t)chunks[c0 + 1].n * 256u;   // hi + lo images
      if (elect_one()) {
        uint32_t a_bytes = 0;
        if constexpr (SRC == SRC_TILES) a_bytes = (p.debug & 4) ? 0u : (uint32_t)slices * a_tile_bytes;
        mbar_arrive_expect_tx(&s_full[s], a_bytes + b0 + b1);
        if constexpr (SRC == SRC_TILES) {
          // rows [r_lo, r_lo + tile_rows) of the STFT kernel's 128-row power tiles: one contiguous piece per
          // slice, two when the range straddles a 128-row tile (both 8-row aligned, so atoms stay whole)
          const int64_t r_lo = tile * tile_rows;
          const int64_t t0 = r_lo >> 7;
          const uint32_t in0 = (uint32_t)(r_lo & 127);
          const uint32_t n0 = min((uint32_t)tile_rows, 128u - in0);
          const unsigned char* base = reinterpret_cast<const unsigned char*>(p.src);
          for (int j = 0; j < slices && !(p.debug & 4); ++j) {
            const unsigned char* src0 = base + (((size_t)t0 * n_chunks + (c0 + j)) * 128 + in0) * 128;
            bulk_g2s_hint(st + j * a_tile_bytes, src0, n0 * 128u, &s_full[s], pol_stream);
            if (n0 < (uint32_t)tile_rows) {
              const unsigned char* src1 = base + (((size_t)(t0 + 1) * n_chunks + (c0 + j)) * 128) * 128;
              bulk_g2s_hint(st + j * a_tile_bytes + n0 * 128u, src1, ((uint32_t)tile_rows - n0) * 128u, &s_full[s], pol_stream);
            }
          }
        }
        if (b0) bulk_g2s_hint(st + a_slot, p.plan + chunks[c0].blob_off, b0, &s_full[s], pol_keep);
        if (b1) bulk_g2s_hint(st + a_slot + b_slot, p.plan + chunks[c0 + 1].blob_off, b1, &s_full[s], pol_keep);
      }
      __syncwarp();
      if (lane == 0) MB_TRACE(0, 2 * it + 1);
    }
    }   // tile loop
  }

  tc_fence_before();
  __syncthreads();
  if (tid == 0) MB_TRACE(3, 3);
  if (warp == kMbProducerWarps + 1) tmem_dealloc(tmem, kMbTmemCols);
}

int dump_melbank_trace() {
  long long h[4][80];
  TAC_CUDA_OK(cudaDeviceSynchronize());
  TAC_CUDA_OK(cudaMemcpyFromSymbol(h, g_mb_trace, sizeof(h)));
  const long long t0 = h[3][0];
  const char* names[4] = {"loader", "producer", "mma", "misc"};
  for (int r = 0; r < 4; ++r) {
    printf("%s:", names[r]);
    for (int i = 0; i < (r == 3 ? 4 : 72); ++i) printf(" %lld", h[r][i] ? h[r][i] - t0 : -1);
    printf("\n");
  }
  return TAC_OK;
}

template <int SRC>
static int launch_melbank(MelbankParams p, int64_t tiles, cudaStream_t stream) {
  TAC_CUDA_OK(cudaFuncSetAttribute(melbank_kernel<SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMbSmemBytes));
  const int bblocks = (p.n_bands + kMbBandBlock - 1) / kMbBandBlock;
  const int64_t ctas = tiles < sm_count() ? tiles : sm_count();     // persistent: one CTA per SM walks its tiles
  dim3 grid((unsigned)ctas, (unsigned)bblocks);
  static int debug_flags = -1;
  if (debug_flags < 0) {
    const char* e = getenv("TAC_MB_DEBUG");
    debug_flags = e ? atoi(e) : 0;
  }
  p.debug = debug_flags;
  LaunchProbe probe(KIND_MELBANK, stream);
  melbank_kernel<SRC><<<grid, kMbThreads, kMbSmemBytes, stream>>>(p);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

// power tiles written by the STFT kernel (pipeline path); tile_rows is the height it used
int launch_melbank_tiles(const float* tiles_ws, int64_t n_rows, int tile_rows, int64_t g_base, int64_t frames, int n_bins,
                         const void* plan_dev, int n_bands, int to_db, float ref, float amin, float* out,
                         cudaStream_t stream) {
  if (n_rows <= 0) return TAC_OK;
  TAC_REQUIRE((reinterpret_cast<uintptr_t>(tiles_ws) & 127) == 0 && tile_rows >= 8 && tile_rows <= kMbRows && (tile_rows & 7) == 0,
              TAC_ERR_INVALID, "melbank: power tiles must be 128-byte aligned, row range a multiple of 8 in [8, 128]");
  TAC_REQUIRE((reinterpret_cast<uintptr_t>(plan_dev) & 15) == 0, TAC_ERR_INVALID, "melbank: plan must be 16-byte aligned");
  MelbankParams p;
  memset(&p, 0, sizeof(p));
  p.src = tiles_ws;
  p.plan = static_cast<const unsigned char*>(plan_dev);
  p.out = out;
  p.rows = n_rows;
  p.g_base = g_base;
  p.frames = frames;
  p.n_bins = n_bins;
  p.n_bands = n_bands;
  p.rows_per_tile = tile_rows;
  p.to_db = to_db ? 1 : 0;
  p.amin = amin;
  p.log10_ref = log10f(ref);
  return launch_melbank<SRC_TILES>(p, (n_rows + tile_rows - 1) / tile_rows, stream);
}

// rows per tile that spreads `rows` frames evenly over the SMs (<= 128, multiple of 8)
int balanced_tile_rows(int64_t rows) {
  const int sms = sm_count();
  const int64_t waves = (rows + (int64_t)sms * kMbRows - 1) / ((int64_t)sms * kMbRows);
  int64_t tr = (rows + waves * sms - 1) / (waves * sms);
  tr = ((tr + 7) / 8) * 8;
  if (tr < 8) tr = 8;
  if (tr > kMbRows) tr = kMbRows;
  return (int)tr;
}

}  // namespace tac

// ---------------------------------------------------------------------------------------------
// plan construction (host)
// ---------------------------------------------------------------------------------------------
extern "C" int64_t tac_fbplan_bytes(int n_bins, int n_bands) {
  using namespace tac;
  if (n_bins <= 0 || n_bands <= 0) return 0;
  const int64_t n_chunks = (n_bins + kMbBK - 1) / kMbBK;
  const int64_t n_bblocks = (n_bands + kMbBandBlock - 1) / kMbBandBlock;
  return 128 + (int64_t)sizeof(FbPlanHeader) + n_chunks * n_bblocks * ((int64_t)sizeof(FbPlanChunk) + 2 * kMbTileBytes) +
         band_plan_capacity(n_bands) + range_plan_capacity(n_bins, n_bands);
}

extern "C" int tac_fbplan_build_host(const float* fb, int n_bins, int n_bands, void* plan_host, int64_t capacity,
                                     int64_t* used) {
  using namespace tac;
  TAC_REQUIRE(fb && plan_host && used, TAC_ERR_INVALID, "fbplan: null pointer");
  TAC_REQUIRE(n_bins > 0 && n_bands > 0, TAC_ERR_INVALID, "fbplan: bad shape (%d, %d)", n_bins, n_bands);
  TAC_REQUIRE(capacity >= tac_fbplan_bytes(n_bins, n_bands), TAC_ERR_WORKSPACE, "fbplan: buffer of %lld bytes is too small",
              (long long)capacity);
  const int n_chunks = (n_bins + kMbBK - 1) / kMbBK;
  const int n_bblocks = (n_bands + kMbBandBlock - 1) / kMbBandBlock;
  unsigned char* base = static_cast<unsigned char*>(plan_host);
  FbPlanHeader* hdr = reinterpret_cast<FbPlanHeader*>(base);
  memset(hdr, 0, sizeof(*hdr));
  hdr->magic = kPlanMagic;
  hdr->n_bins = n_bins;
  hdr->n_bands = n_bands;
  hdr->n_chunks = n_chunks;
  hdr->n_bblocks = n_bblocks;
  FbPlanChunk* table = reinterpret_cast<FbPlanChunk*>(base + sizeof(FbPlanHeader));
  int64_t off = (int64_t)sizeof(FbPlanHeader) + (int64_t)n_chunks * n_bblocks * sizeof(FbPlanChunk);
  off = (off + 127) & ~(int64_t)127;
  for (int bb = 0; bb < n_bblocks; ++bb) {
    const int band_begin = bb * kMbBandBlock;
    const int band_end = (band_begin + kMbBandBlock < n_bands) ? band_begin + kMbBandBlock : n_bands;
    for (int c = 0; c < n_chunks; ++c) {
      const int k_begin = c * kMbBK;
      const int k_end = (k_begin + kMbBK < n_bins) ? k_begin + kMbBK : n_bins;
      int lo = INT32_MAX, hi = -1;
      for (int k = k_begin; k < k_end; ++k)
        for (int b = band_begin; b < band_end; ++b)
          if (fb[(size_t)k * n_bands + b] != 0.0f) {
            if (b < lo) lo = b;
            if (b > hi) hi = b;
          }
      FbPlanChunk& e = table[(size_t)bb * n_chunks + c];
      memset(&e, 0, sizeof(e));
      if (hi < 0) continue;                                        // all-zero block: skipped by the kernel
      const int rel_lo = ((lo - band_begin) / 16) * 16;
      const int rel_hi = ((hi - band_begin) / 16 + 1) * 16;        // exclusive, <= 128
      const int n = rel_hi - rel_lo;
      e.band_lo = rel_lo;
      e.n = n;
      if (n > hdr->max_n) hdr->max_n = n;
      e.blob_off = (int32_t)off;
      float* img_hi = reinterpret_cast<float*>(base + off);
      float* img_lo = reinterpret_cast<float*>(base + off + (int64_t)n * 128);
      for (int j = 0; j < n; ++j) {
        const int b = band_begin + rel_lo + j;
        for (int kk = 0; kk < kMbBK; ++kk) {
          const int k = k_begin + kk;
          const float v = (b < n_bands && k < n_bins) ? fb[(size_t)k * n_bands + b] : 0.0f;
          uint32_t bits;
          memcpy(&bits, &v, 4);
          bits &= 0xFFFFE000u;
          float h;
          memcpy(&h, &bits, 4);
          const size_t idx = (size_t)(j >> 3) * 256 + (size_t)(j & 7) * 32 + (size_t)(((kk >> 2) ^ (j & 7)) << 2) + (kk & 3);
          img_hi[idx] = h;
          img_lo[idx] = v - h;
        }
      }
      off += 2 * (int64_t)n * 128;
    }
  }
  // the same matrix as a band plan for the fused STFT + filterbank kernel, when it has that form
  off = (off + 127) & ~(int64_t)127;
  const int64_t band_bytes = build_band_plan(fb, n_bins, n_bands, base + off, capacity - off);
  if (band_bytes > 0) {
    hdr->band_off = (int32_t)off;
    hdr->band_cmax = reinterpret_cast<const BandPlanHeader*>(base + off)->cmax;
    off += band_bytes;
  }
  // and as per-band bin ranges for the fused epilogue of the other fft lengths
  off = (off + 127) & ~(int64_t)127;
  const int64_t range_bytes = build_range_plan(fb, n_bins, n_bands, base + off, capacity - off);
  if (range_bytes > 0) {
    hdr->range_off = (int32_t)off;
    hdr->range_bytes = (int32_t)range_bytes;
    off += range_bytes;
  }
  *used = off;
  return TAC_OK;
}

// Handle of the one-kernel mel path for this plan and fft length, 0 when there is none: n_fft = 2048 -> the band plan
// (tac_fbplan_band_handle), 256 / 512 / 1024 / 4096 -> the range plan (offset | bytes << 32).
extern "C" int64_t tac_fbplan_fused_handle(const void* plan_host, int n_fft) {
  using namespace tac;
  if (!plan_host) return 0;
  const FbPlanHeader* hdr = static_cast<const FbPlanHeader*>(plan_host);
  if (hdr->magic != kPlanMagic || hdr->n_bins != n_fft / 2 + 1) return 0;
  if (n_fft == 2048) return tac_fbplan_band_handle(plan_host);
  if ((n_fft == 256 || n_fft == 512 || n_fft == 1024 || n_fft == 4096) && hdr->range_off > 0)
    return (int64_t)hdr->range_off | ((int64_t)hdr->range_bytes << 32);
  return 0;
}

extern "C" int64_t tac_fbplan_band_handle(const void* plan_host) {
  using namespace tac;
  if (!plan_host) return 0;
  const FbPlanHeader* hdr = static_cast<const FbPlanHeader*>(plan_host);
  if (hdr->magic != kPlanMagic || hdr->band_off <= 0) return 0;
  return (int64_t)hdr->band_off | ((int64_t)hdr->band_cmax << 48);
}

extern "C" int tac_debug_dump_melbank_trace(void) { return tac::dump_melbank_trace(); }

extern "C" int tac_power_mel_f32(const float* spec, int is_complex, float power, int64_t n_seq, int64_t frames, int n_bins,
                                 const void* plan_dev, int n_bands, int to_db, float ref, float amin, float* out,
                                 void* stream) {
  using namespace tac;
  TAC_REQUIRE(n_seq >= 0 && frames >= 0 && n_bins > 0 && n_bands > 0, TAC_ERR_INVALID, "power_mel: bad shape");
  TAC_REQUIRE(n_bins <= kMbMaxChunks * kMbBK, TAC_ERR_UNSUPPORTED, "power_mel: %d frequency bins exceed the %d supported",
              n_bins, kMbMaxChunks * kMbBK);
  if (n_seq * frames == 0) return TAC_OK;
  TAC_REQUIRE(spec && plan_dev && out, TAC_ERR_INVALID, "power_mel: null pointer");
  TAC_REQUIRE((reinterpret_cast<uintptr_t>(plan_dev) & 15) == 0, TAC_ERR_INVALID, "power_mel: plan must be 16-byte aligned");
  MelbankParams p;
  memset(&p, 0, sizeof(p));
  p.src = spec;
  p.plan = static_cast<const unsigned char*>(plan_dev);
  p.out = out;
  p.rows = n_seq * frames;
  p.g_base = 0;
  p.frames = frames;
  p.half_power = 0.5f * power;
  p.power_mode = power == 2.0f ? 2 : (power == 1.0f ? 1 : 0);
  p.n_bins = n_bins;
  p.n_bands = n_bands;
  p.rows_per_tile = balanced_tile_rows(p.rows);
  p.to_db = to_db ? 1 : 0;
  p.amin = amin;
  p.log10_ref = log10f(ref);
  const int64_t tiles = (p.rows + p.rows_per_tile - 1) / p.rows_per_tile;
  return is_complex ? launch_melbank<SRC_PUBLIC_COMPLEX>(p, tiles, as_stream(stream))
                    : launch_melbank<SRC_PUBLIC_REAL>(p, tiles, as_stream(stream));
}
