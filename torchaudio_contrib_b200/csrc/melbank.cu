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
// Warp roles (320 threads): warps 0-7 convert P tiles (global -> |.|^p -> hi/lo -> swizzled smem)
// and later run the epilogue (TMEM -> registers -> dB -> global); warp 8 lane 0 issues the MMAs;
// warp 9 lane 0 streams the plan blocks with 1-D bulk async copies.  Three 64 KB stages, mbarrier
// full/empty handshakes, tcgen05.commit releases a stage when its MMAs have read it.
#include "tac_common.cuh"

namespace tac {

constexpr int kMbRows = 128;
constexpr int kMbBK = 32;
constexpr int kMbStages = 3;
constexpr int kMbProducerWarps = 8;
constexpr int kMbProducerThreads = kMbProducerWarps * 32;
constexpr int kMbThreads = kMbProducerThreads + 64;
constexpr int kMbBandBlock = 128;
constexpr int kMbTileBytes = kMbRows * kMbBK * 4;          // 16 KB: one operand tile
constexpr int kMbStageBytes = 4 * kMbTileBytes;            // A_hi, A_lo, B_hi, B_lo
constexpr size_t kMbSmemBytes = (size_t)kMbStages * kMbStageBytes + 1024;   // + alignment slack
constexpr uint32_t kPlanMagic = 0x7ac0fb01u;

struct FbPlanHeader {
  uint32_t magic;
  int32_t n_bins, n_bands, n_chunks, n_bblocks;
  int32_t reserved[3];
};
struct FbPlanChunk {
  int32_t band_lo;    // first band of the block, relative to the band block, multiple of 16
  int32_t n;          // bands in the block, multiple of 16, 0 = nothing to do for this K slice
  int32_t blob_off;   // byte offset of the hi image from the start of the plan (lo image follows)
  int32_t reserved;
};

struct MelbankParams {
  const float* src;
  const unsigned char* plan;
  float* out;
  int64_t rows;          // frames handled by this launch
  int64_t g_base;        // flattened frame index (seq * frames + t) of row 0
  int64_t frames;        // frames per sequence
  int layout;            // 0: frame-major rows src[row * kpad + bin]; 1: public src[((seq*bins)+bin)*frames + t]
  int is_complex;        // public layout only: src holds (re, im) pairs
  int power_mode;        // complex input: 2 -> re^2+im^2, 1 -> sqrt, 0 -> pow(., power/2)
  float power;
  int n_bins, n_bands, kpad;
  int rows_per_tile;
  int to_db;
  float amin, log10_ref;
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

__device__ __forceinline__ float power_of(float re, float im, float power, int mode) {
  const float s = fmaf(re, re, im * im);
  if (mode == 2) return s;
  if (mode == 1) return sqrtf(s);
  return s > 0.0f ? powf(s, 0.5f * power) : (power == 0.0f ? 1.0f : 0.0f);
}

__global__ void __launch_bounds__(kMbThreads, 1) melbank_kernel(const MelbankParams p) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  __shared__ uint64_t s_full_b[kMbStages], s_a_ready[kMbStages], s_empty[kMbStages], s_accum;
  __shared__ uint32_t s_tmem;

  // 1024-byte aligned stage buffers (swizzle atoms must not straddle 1 KB boundaries)
  unsigned char* stage0 = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const FbPlanHeader* hdr = reinterpret_cast<const FbPlanHeader*>(p.plan);
  const int n_chunks = hdr->n_chunks;
  const FbPlanChunk* chunks = reinterpret_cast<const FbPlanChunk*>(p.plan + sizeof(FbPlanHeader)) + (size_t)blockIdx.y * n_chunks;

  const int64_t row0 = (int64_t)blockIdx.x * p.rows_per_tile;
  const int valid = (int)min((int64_t)p.rows_per_tile, p.rows - row0);

  if (warp == kMbProducerWarps && lane == 0) {
    for (int s = 0; s < kMbStages; ++s) {
      mbar_init(&s_full_b[s], 1);
      mbar_init(&s_a_ready[s], kMbProducerThreads);
      mbar_init(&s_empty[s], 1);
    }
    mbar_init(&s_accum, 1);
    fence_mbar_init();
  }
  if (warp == kMbProducerWarps + 1) {
    tmem_alloc(&s_tmem, kMbBandBlock);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  // zero the accumulator: blocks of different K slices touch different column ranges, so every
  // MMA accumulates (there is no single "first" MMA per column)
  if (warp < kMbProducerWarps) {
    const uint32_t t0 = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(64 * (warp >> 2));
#pragma unroll
    for (int j = 0; j < 4; ++j) tmem_zero16(t0 + 16 * j);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < kMbProducerWarps) {
    // =========================== P-tile producers ===============================================
    // layout 0: thread -> rows r = tid/8 + 32 i, 16-byte column c16 = tid % 8   (coalesced 128 B rows)
    // layout 1: thread -> row m = tid % 128, columns c16 = 4 * (tid / 128) + i  (coalesced along time)
    const int64_t g_first = p.g_base + row0;
    float4 cur[4], nxt[4];

    auto load_chunk = [&](int c, float4 (&v)[4]) {
      const int k0 = c * kMbBK;
      if (p.layout == 0) {
        const int c16 = tid & 7;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = (tid >> 3) + 32 * i;
          v[i] = (r < valid) ? ldg_stream_f4(reinterpret_cast<const float4*>(p.src + (row0 + r) * p.kpad + k0) + c16)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        const int m = tid & 127;
        const bool ok = m < valid;
        const int64_t g = g_first + m;
        const int64_t seq = g / p.frames, t = g - seq * p.frames;
        const int64_t base = seq * p.n_bins * p.frames + t;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kk = k0 + 4 * (4 * (tid >> 7) + i);
          float e[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = kk + j;
            float val = 0.0f;
            if (ok && k < p.n_bins) {
              const int64_t idx = base + (int64_t)k * p.frames;
              if (p.is_complex) {
                const float2 z = __ldg(reinterpret_cast<const float2*>(p.src) + idx);
                val = power_of(z.x, z.y, p.power, p.power_mode);
              } else {
                val = ldg_stream_f1(p.src + idx);
              }
            }
            e[j] = val;
          }
          v[i] = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
    };
    auto store_chunk = [&](unsigned char* a_hi, unsigned char* a_lo, const float4 (&v)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int r, c16;
        if (p.layout == 0) {
          r = (tid >> 3) + 32 * i;
          c16 = tid & 7;
        } else {
          r = tid & 127;
          c16 = 4 * (tid >> 7) + i;
        }
        float4 hi, lo;
        split_tf32(v[i], hi, lo);
        const uint32_t off = swz_off(r, c16);
        *reinterpret_cast<float4*>(a_hi + off) = hi;
        *reinterpret_cast<float4*>(a_lo + off) = lo;
      }
    };

    // first active slice
    int c = 0;
    while (c < n_chunks && chunks[c].n == 0) ++c;
    if (c < n_chunks) load_chunk(c, cur);
    int it = 0;
    while (c < n_chunks) {
      int c_next = c + 1;
      while (c_next < n_chunks && chunks[c_next].n == 0) ++c_next;
      if (c_next < n_chunks) load_chunk(c_next, nxt);
      const int s = it % kMbStages;
      const uint32_t ph = (uint32_t)(it / kMbStages) & 1u;
      mbar_wait(&s_empty[s], ph ^ 1u);
      unsigned char* st = stage0 + (size_t)s * kMbStageBytes;
      store_chunk(st, st + kMbTileBytes, cur);
      fence_proxy_async();
      mbar_arrive(&s_a_ready[s]);
#pragma unroll
      for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
      c = c_next;
      ++it;
    }

    // =========================== epilogue =======================================================
    mbar_wait(&s_accum, 0);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    const int m = 32 * q + lane;
    const bool ok = m < valid;
    const int64_t g = g_first + m;
    const int64_t seq = g / p.frames, t = g - seq * p.frames;
    const int band0 = blockIdx.y * kMbBandBlock + 64 * half;
    float* out_base = p.out + (seq * p.n_bands + band0) * p.frames + t;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc[16];
      tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(64 * half + 16 * j), acc);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int b = 16 * j + i;
        if (ok && band0 + b < p.n_bands) {
          float v = acc[i];
          if (p.to_db) {
            float s2 = v * v;
            s2 = (s2 < p.amin) ? p.amin : s2;
            v = 10.0f * (log10f(s2) - p.log10_ref);
          }
          out_base[(int64_t)b * p.frames] = v;
        }
      }
    }
  } else if (warp == kMbProducerWarps) {
    // =========================== MMA issuer =====================================================
    if (lane == 0) {
      int it = 0;
      for (int c = 0; c < n_chunks; ++c) {
        const int n = chunks[c].n;
        if (n == 0) continue;
        const int s = it % kMbStages;
        const uint32_t ph = (uint32_t)(it / kMbStages) & 1u;
        mbar_wait(&s_full_b[s], ph);
        mbar_wait(&s_a_ready[s], ph);
        tc_fence_after();
        const uint32_t st = smem_u32(stage0 + (size_t)s * kMbStageBytes);
        const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + kMbTileBytes);
        const uint64_t b_hi = umma_desc_sw128(st + 2 * kMbTileBytes), b_lo = umma_desc_sw128(st + 3 * kMbTileBytes);
        const uint32_t idesc = umma_idesc_tf32(n);
        const uint32_t d = tmem + (uint32_t)chunks[c].band_lo;
#pragma unroll
        for (int ks = 0; ks < kMbBK / 8; ++ks) {        // UMMA K = 8 tf32 = 32 bytes -> +2 in the address field
          tc_mma_tf32(d, a_lo + 2 * ks, b_hi + 2 * ks, idesc, 1u);
          tc_mma_tf32(d, a_hi + 2 * ks, b_lo + 2 * ks, idesc, 1u);
          tc_mma_tf32(d, a_hi + 2 * ks, b_hi + 2 * ks, idesc, 1u);
        }
        tc_commit(&s_empty[s]);
        ++it;
      }
      tc_commit(&s_accum);
    }
  } else {
    // =========================== plan block loader ==============================================
    if (lane == 0) {
      int it = 0;
      for (int c = 0; c < n_chunks; ++c) {
        const int n = chunks[c].n;
        if (n == 0) continue;
        const int s = it % kMbStages;
        const uint32_t ph = (uint32_t)(it / kMbStages) & 1u;
        mbar_wait(&s_empty[s], ph ^ 1u);
        unsigned char* st = stage0 + (size_t)s * kMbStageBytes;
        const uint32_t bytes = (uint32_t)n * 128u;
        mbar_arrive_expect_tx(&s_full_b[s], 2 * bytes);
        bulk_g2s(st + 2 * kMbTileBytes, p.plan + chunks[c].blob_off, bytes, &s_full_b[s]);
        bulk_g2s(st + 3 * kMbTileBytes, p.plan + chunks[c].blob_off + bytes, bytes, &s_full_b[s]);
        ++it;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMbProducerWarps + 1) tmem_dealloc(tmem, kMbBandBlock);
}

int launch_melbank(MelbankParams p, cudaStream_t stream) {
  if (p.rows <= 0) return TAC_OK;
  static bool configured[64] = {false};
  int dev = 0;
  TAC_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    TAC_CUDA_OK(cudaFuncSetAttribute(melbank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMbSmemBytes));
    configured[dev] = true;
  }
  const int sms = sm_count();
  int64_t tiles = (p.rows + kMbRows - 1) / kMbRows;
  if (tiles > sms) tiles = ((tiles + sms - 1) / sms) * sms;      // even out the last wave
  const int rpt = (int)((p.rows + tiles - 1) / tiles);
  tiles = (p.rows + rpt - 1) / rpt;
  p.rows_per_tile = rpt;
  const int bblocks = (p.n_bands + kMbBandBlock - 1) / kMbBandBlock;
  dim3 grid((unsigned)tiles, (unsigned)bblocks);
  LaunchProbe probe(KIND_MELBANK, stream);
  melbank_kernel<<<grid, kMbThreads, kMbSmemBytes, stream>>>(p);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

// frame-major rows written by the STFT kernel (pipeline path)
int launch_melbank_rows(const float* rows, int64_t n_rows, int64_t g_base, int64_t frames, int n_bins, int kpad,
                        const void* plan_dev, int n_bands, int to_db, float ref, float amin, float* out,
                        cudaStream_t stream) {
  TAC_REQUIRE((reinterpret_cast<uintptr_t>(rows) & 15) == 0 && (kpad & 31) == 0, TAC_ERR_INVALID,
              "melbank: power rows must be 16-byte aligned with a row length that is a multiple of 32");
  TAC_REQUIRE((reinterpret_cast<uintptr_t>(plan_dev) & 15) == 0, TAC_ERR_INVALID, "melbank: plan must be 16-byte aligned");
  MelbankParams p;
  memset(&p, 0, sizeof(p));
  p.src = rows;
  p.plan = static_cast<const unsigned char*>(plan_dev);
  p.out = out;
  p.rows = n_rows;
  p.g_base = g_base;
  p.frames = frames;
  p.layout = 0;
  p.n_bins = n_bins;
  p.n_bands = n_bands;
  p.kpad = kpad;
  p.to_db = to_db ? 1 : 0;
  p.amin = amin;
  p.log10_ref = log10f(ref);
  return launch_melbank(p, stream);
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
  return 128 + (int64_t)sizeof(FbPlanHeader) + n_chunks * n_bblocks * ((int64_t)sizeof(FbPlanChunk) + 2 * kMbTileBytes);
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
  *used = off;
  return TAC_OK;
}

extern "C" int tac_power_mel_f32(const float* spec, int is_complex, float power, int64_t n_seq, int64_t frames, int n_bins,
                                 const void* plan_dev, int n_bands, int to_db, float ref, float amin, float* out,
                                 void* stream) {
  using namespace tac;
  TAC_REQUIRE(n_seq >= 0 && frames >= 0 && n_bins > 0 && n_bands > 0, TAC_ERR_INVALID, "power_mel: bad shape");
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
  p.layout = 1;
  p.is_complex = is_complex ? 1 : 0;
  p.power = power;
  p.power_mode = power == 2.0f ? 2 : (power == 1.0f ? 1 : 0);
  p.n_bins = n_bins;
  p.n_bands = n_bands;
  p.kpad = kpad_for_bins(n_bins);
  p.to_db = to_db ? 1 : 0;
  p.amin = amin;
  p.log10_ref = log10f(ref);
  return launch_melbank(p, as_stream(stream));
}
