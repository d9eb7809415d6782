// K1: framing + padding + window + real FFT (+ optional |.|^p epilogue).
// Replaces the torch.stft call at reference functional.py:99-107 (and complex_norm :126-128 when
// the caller is the Spectrogram / Melspectrogram pipeline).
//
// Two kernels:
//
//  stft2048_kernel   n_fft = 2048, onesided.  One warp owns one frame at a time.  The frame's 2048
//      samples (8 KB, contiguous in HBM) arrive in the warp's private shared-memory slab through a
//      1-D bulk async copy (TMA engine, mbarrier completion); frames that touch the padding, or
//      unaligned inputs, are gathered with plain loads instead.  The real FFT is computed as a
//      1024-point complex FFT of z[n] = x[2n] + i x[2n+1], factored 32 x 32: each lane runs a
//      32-point FFT in registers, the 32x32 transpose goes through the same slab (row stride 33
//      -> conflict free), each lane runs the second 32-point FFT, and the real-FFT untangling
//      X[k] = E + W^k O pairs bin k with bin 1024-k by one shuffle exchange.  The slab is free
//      again after the transpose read, so the next frame's bulk copy overlaps the second half.
//  stft_dft_kernel       any n_fft in [2, 8192] that is not a power of two (400, 1200, 441 ...): direct DFT, correct-first
//  stft_generic_kernel   any power-of-two n_fft in [32, 8192], any padding mode, one- or two-sided:
//      one CTA per frame, radix-2 Stockham passes through shared memory.  Correct-first fallback.
//
// Output modes (OUT_*): the public layouts of the reference, and the frame-major power rows that
// feed the filterbank kernel (melbank.cu) inside the Melspectrogram pipeline.
#include <stdlib.h>

#include "bandplan.cuh"
#include "fft_regs.cuh"
#include "fft_stockham.cuh"
#include "stft_params.cuh"
#include "tac_common.cuh"

namespace tac {

// ---------------------------------------------------------------------------------------------
// shared pieces
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float spectral_power(float re, float im, float power, int mode) {
  const float s = fmaf(re, re, im * im);
  if (mode == 2) return s;
  if (mode == 1) return sqrtf(s);
  return s > 0.0f ? powf(s, 0.5f * power) : (power == 0.0f ? 1.0f : 0.0f);
}

__device__ __forceinline__ void emit_bin(const StftParams& p, int64_t seq, int64_t t, int64_t row, int k,
                                         float re, float im) {
  if (p.out_mode == OUT_POWER_ROWS) {
    p.out[power_tile_index(row, k, p.kpad)] = spectral_power(re, im, p.power, p.power_mode);
  } else if (p.out_mode == OUT_POWER_PUBLIC) {
    p.out[(seq * p.bins + k) * p.frames + t] = spectral_power(re, im, p.power, p.power_mode);
  } else {
    reinterpret_cast<float2*>(p.out)[(seq * p.bins + k) * p.frames + t] = make_float2(re, im);
  }
}

// ---------------------------------------------------------------------------------------------
// fast path: n_fft = 2048
// ---------------------------------------------------------------------------------------------
constexpr int kFastWarps = 16;
constexpr int kFastThreads = kFastWarps * 32;
constexpr int kSlabStride = 34;                              // complex per slab row: 272 B keeps 16-byte alignment, conflict free
constexpr int kSlabComplex = 32 * kSlabStride + 2;           // transposition slab (+2: 16 slabs hold a 1025 x 17 float2 output tile)
constexpr size_t kFastSmemBytes = 3 * 1024 * sizeof(float2)  // window pairs, tw1, tw2
                                  + kFastWarps * sizeof(uint64_t) + kFastWarps * kSlabComplex * sizeof(float2);
// OUT_MEL_FUSED: plus one power-spectrum stash per warp (bandplan.cuh) -- 232 320 of the 232 448 bytes a CTA may have
constexpr size_t kFusedSmemBytes = kFastSmemBytes + kFastWarps * kStashFloats * sizeof(float);
static_assert(kFusedSmemBytes <= 227 * 1024, "fused STFT + filterbank kernel exceeds the shared memory of one CTA");

// |X|^p with the exponent resolved at compile time (PMODE 2: p = 2, 1: p = 1, 0: runtime p through
// the hardware lg2 / ex2 units -- ~1e-6 relative, far inside the parity budget)
template <int PMODE>
__device__ __forceinline__ float fast_power(float re, float im, float half_power) {
  const float s = fmaf(re, re, im * im);
  if constexpr (PMODE == 2) return s;
  if constexpr (PMODE == 1) return sqrtf(s);
  return s > 0.0f ? exp2f(half_power * __log2f(s)) : (half_power == 0.0f ? 1.0f : 0.0f);
}

// Front half of a frame's FFT, shared by both n_fft = 2048 kernels.  In: the 2048 samples in `slab`.
// Out: v[n1] = pass-2 input of lane k2 = lane (pass-1 result transposed through the slab and multiplied by
// W_1024^(n1 k2)); the slab is free again on return.
__device__ __forceinline__ void fft2048_front(float2 (&v)[32], float2* slab, const float2* s_win, const float2* s_tw1, int lane) {
#pragma unroll
  for (int r = 0; r < 32; r += 2) {                // lane = n1, register r <-> z[n], n = n1 + 32 r, windowed
    const float2 x0 = slab[lane + 32 * r], x1 = slab[lane + 32 * r + 32];
    const float4 w = reinterpret_cast<const float4*>(s_win)[(r >> 1) * 32 + lane];
    v[r] = make_float2(x0.x * w.x, x0.y * w.y);
    v[r + 1] = make_float2(x1.x * w.z, x1.y * w.w);
  }
  __syncwarp();                                    // samples consumed; slab becomes the transpose buffer
  dit_fft_fma<32>(v);                              // pass 1: over r for fixed n1 = lane
#pragma unroll
  for (int k2 = 0; k2 < 32; ++k2) slab[k2 * kSlabStride + lane] = v[bit_reverse<32>(k2)];
  __syncwarp();
#pragma unroll
  for (int n1 = 0; n1 < 32; n1 += 2) {
    const float4 a = *reinterpret_cast<const float4*>(slab + lane * kSlabStride + n1);
    const float4 w = reinterpret_cast<const float4*>(s_tw1)[(n1 >> 1) * 32 + lane];   // W_1024^(n1 * k2), k2 = lane
    v[n1] = make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
    v[n1 + 1] = make_float2(fmaf(a.z, w.z, -a.w * w.w), fmaf(a.z, w.w, a.w * w.z));
  }
  __syncwarp();
}

// Real-FFT untangling of bin k = 32 K1 + lane from v[bit_reverse(k1)] = Z[32 k1 + lane] / 2: pairs with bin
// 1024 - k, held by lane (32 - lane) % 32 in register 31 - K1 (lane 0 pairs with itself, register (32 - K1) % 32).
template <int K1>
__device__ __forceinline__ float2 fft2048_untangle(const float2 (&v)[32], const float2 w, int lane, int partner) {
  const float2 z = v[bit_reverse<32>(K1)];
  float2 q;
  q.x = __shfl_sync(0xffffffffu, v[bit_reverse<32>(31 - K1)].x, partner);
  q.y = __shfl_sync(0xffffffffu, v[bit_reverse<32>(31 - K1)].y, partner);
  if (lane == 0) q = v[bit_reverse<32>((32 - K1) & 31)];
  const float a = z.x + q.x, b = z.y - q.y, gs = z.y + q.y, h = q.x - z.x;
  return make_float2(fmaf(w.x, gs, fmaf(-w.y, h, a)), fmaf(w.x, h, fmaf(w.y, gs, b)));
}

// W_2048^(32 K1 + lane) = (c, d): the table is pair-interleaved, so even K1 fetch (K1, K1 + 1) with one LDS.128 into
// `pair` and odd K1 take its second half (two LDS.64 16 bytes apart cost two wavefronts each)
template <int K1>
__device__ __forceinline__ float2 fft2048_tw2(const float2* s_tw2, int lane, float4& pair) {
  if constexpr ((K1 & 1) == 0) {
    pair = reinterpret_cast<const float4*>(s_tw2)[(K1 >> 1) * 32 + lane];
    return make_float2(pair.x, pair.y);
  } else {
    return make_float2(pair.z, pair.w);
  }
}

// The same twiddle without the table, for the mel kernel (which only needs K1 = 0 .. 16 and is bound by shared-memory
// bandwidth, not by the fp32 pipe): W_2048^(32 K1 + lane) = W_64^K1 * W_2048^lane, the first factor a compile-time
// constant, the second one float2 per lane (`base`, table row K1 = 0).  Four FP instructions instead of half an LDS.128.
__device__ constexpr float kW64[17][2] = {{1.000000000e+00f, -0.000000000e+00f}, {9.951847267e-01f, -9.801714033e-02f}, {9.807852804e-01f, -1.950903220e-01f}, {9.569403357e-01f, -2.902846773e-01f}, {9.238795325e-01f, -3.826834324e-01f}, {8.819212643e-01f, -4.713967368e-01f}, {8.314696123e-01f, -5.555702330e-01f}, {7.730104534e-01f, -6.343932842e-01f}, {7.071067812e-01f, -7.071067812e-01f}, {6.343932842e-01f, -7.730104534e-01f}, {5.555702330e-01f, -8.314696123e-01f}, {4.713967368e-01f, -8.819212643e-01f}, {3.826834324e-01f, -9.238795325e-01f}, {2.902846773e-01f, -9.569403357e-01f}, {1.950903220e-01f, -9.807852804e-01f}, {9.801714033e-02f, -9.951847267e-01f}, {6.123233996e-17f, -1.000000000e+00f}};
template <int K1>
__device__ __forceinline__ float2 fft2048_tw2_computed(const float2 base) {
  if constexpr (K1 == 0) return base;
  else if constexpr (K1 == 16) return make_float2(base.y, -base.x);           // W_64^16 = -i
  else {
    constexpr float c = kW64[K1][0], s = kW64[K1][1];
    return make_float2(fmaf(c, base.x, -s * base.y), fmaf(c, base.y, s * base.x));
  }
}

// table set-up shared by both kernels: pair-interleaved [j / 2][lane][j % 2] so one LDS.128 serves two registers
__device__ __forceinline__ void fft2048_tables(const StftParams& p, float2* s_win, float2* s_tw1, float2* s_tw2, int tid, int nthreads) {
  for (int i = tid; i < 1024; i += nthreads) {
    const int j = i >> 5, l = i & 31;                       // register index, lane
    const int slot = ((j >> 1) * 32 + l) * 2 + (j & 1);
    const float g = 0.5f * p.scale;
    s_win[slot] = make_float2(p.window[2 * i] * g, p.window[2 * i + 1] * g);
    float sn, cs;
    sincospif(-2.0f * (float)(j * l) / 1024.0f, &sn, &cs);
    s_tw1[slot] = make_float2(cs, sn);
    sincospif(-2.0f * (float)i / 2048.0f, &sn, &cs);
    s_tw2[slot] = make_float2(cs, sn);
  }
}

// OUT_MEL_FUSED: the frame's power spectrum (in this warp's stash) times a two-band filterbank (bandplan.cuh),
// optional dB epilogue (amplitude_to_db, functional.py:291-296), result to global memory.
// Replaces apply_filterbank's transpose + matmul + transpose (functional.py:183-184) for such matrices.
template <bool PEERS>
__device__ __forceinline__ void band_contract(const StftParams& p, float* stash, int lane, uint32_t seq, uint32_t t) {
  __syncwarp();                                    // the stash holds the whole frame
  // per-band lists of this lane's four bands (fast form: <= 128 bands, <= 4 entries each), fetched now so the
  // loads are long done when the sums are needed (a dependent index -> shared -> add chain per band cost 16 % of
  // the kernel's stall samples in the first version)
  uint4 ci[4];
  const bool fast = p.band_fast != 0;
  if (fast) {
#pragma unroll
    for (int j = 0; j < 4; ++j) ci[j] = __ldg(reinterpret_cast<const uint4*>(p.band_plan + p.band_off_fast) + lane * 4 + j);
  }
  float pw[32];
#pragma unroll
  for (int i = 1; i < 32; ++i) pw[i] = stash[lane * kStashStride + i];     // lane <- bins 32 lane + i
  pw[0] = stash[lane * kStashStride - (lane >= kStashShiftRow ? 1 : 0)];   // column 0 of rows >= 17: see bandplan.cuh
  const float p_last = (lane == 31) ? stash[kStashNyquist] : 0.0f;
  __syncwarp();                                    // stash consumed: it now takes the partial sums
  const uint4 meta = __ldg(reinterpret_cast<const uint4*>(p.band_plan + kBandOffMeta) + lane);
  const float4* wtab = reinterpret_cast<const float4*>(p.band_plan + kBandOffW) + lane;
  const uint32_t mask = meta.x;
  float* row = stash + lane * kStashStride;        // this lane's own (consumed) row takes what it stores
  // Running sums without a store inside the dependency chain: the sum after every bin gets its own register
  // (us[i]), the band step only selects what the next bin continues from, and the stores follow in a batch.
  // (Storing u and then overwriting the same register made every step wait for the store to read its operand.)
  float us[32];
  float u = 0.0f, v = 0.0f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 w = __ldg(wtab + j * 32);
    {
      const float un = fmaf(pw[2 * j], w.x, u), vn = fmaf(pw[2 * j], w.y, v);
      const bool step = (mask >> (2 * j)) & 1u;    // band b -> b + 1 after this bin
      us[2 * j] = un;
      u = step ? vn : un;
      v = step ? 0.0f : vn;
    }
    {
      const float un = fmaf(pw[2 * j + 1], w.z, u), vn = fmaf(pw[2 * j + 1], w.w, v);
      const bool step = (mask >> (2 * j + 1)) & 1u;
      us[2 * j + 1] = un;
      u = step ? vn : un;
      v = step ? 0.0f : vn;
    }
  }
  u = fmaf(p_last, __uint_as_float(meta.z), u);    // bin 1024 (weights are zero except on lane 31)
  v = fmaf(p_last, __uint_as_float(meta.w), v);
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if ((mask >> i) & 1u) row[i] = us[i];
  stash[meta.y] = u;
  stash[meta.y + 1] = v;
  __syncwarp();
  // lane = band: add up the band's slots in list order
  const uint16_t* comb = reinterpret_cast<const uint16_t*>(p.band_plan + kBandOffComb);
  // PEERS: the same offset inside every rank's full output; the stores below go out over NVLink to the peers'
  // memory as they are produced (no gather pass afterwards)
  const int64_t dst_off = (int64_t)(seq + (PEERS ? p.peer_seq0 : 0)) * p.out_seq_stride + (int64_t)t * p.out_t_stride;
  float* dst = (PEERS ? p.peer_out[0] : p.out) + dst_off;
  auto store = [&](int m, float r) {
    if constexpr (PEERS) {
      const int64_t o = dst_off + (int64_t)m * p.out_band_stride;
      if (p.peer_multicast) {
        multimem_st_f32(p.peer_out[0] + o, r);
      } else {
#pragma unroll 1
        for (int q = 0; q < p.n_peers; ++q) __stcs(p.peer_out[q] + o, r);
      }
    } else {
      __stcs(dst + (int64_t)m * p.out_band_stride, r);
    }
  };
  if (fast) {
    const unsigned char* sb = reinterpret_cast<const unsigned char*>(stash);
    float acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {                  // byte offsets into the stash; unused entries point at the zero
      const float a = *reinterpret_cast<const float*>(sb + ci[j].x), b = *reinterpret_cast<const float*>(sb + ci[j].y);
      const float c = *reinterpret_cast<const float*>(sb + ci[j].z);
      const float d = (p.band_cmax > 3) ? *reinterpret_cast<const float*>(sb + ci[j].w) : 0.0f;   // same bits: + 0
      acc[j] = ((a + b) + c) + d;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float r = acc[j];
      if (p.to_db) {
        float s2 = r * r;
        s2 = (s2 < p.amin) ? p.amin : s2;
        r = 10.0f * (log10f(s2) - p.log10_ref);
      }
      const int m = lane + 32 * j;
      if (m < p.n_bands) store(m, r);
    }
    __syncwarp();
    return;
  }
  for (int m = lane; m < p.n_bands; m += 32) {
    float acc = 0.0f;
    for (int c = 0; c < p.band_cmax; ++c) acc += stash[__ldg(comb + c * p.n_bands_pad + m)];
    if (p.to_db) {
      float s2 = acc * acc;
      s2 = (s2 < p.amin) ? p.amin : s2;
      acc = 10.0f * (log10f(s2) - p.log10_ref);
    }
    store(m, acc);
  }
  __syncwarp();                                    // slots consumed before the next frame's spectrum lands
}

// The kernel is specialised on the output mode and the exponent: the frame loop is ~2k fully
// unrolled instructions and has to stay resident in the instruction caches while 16 warps run
// through it at different phases (a first version that branched on these at run time was 12k
// instructions and spent most of its time stalled on instruction fetch).
// clock64 stamps of CTA 0 / warp 0 for timing experiments: compiled in only with -DTAC_K1_TRACE_BUILD (the counter and
// the run-time test cost two registers and ~10 instructions per frame in a kernel that sits at the 128-register cap)
__device__ long long g_k1_trace[64];
#ifdef TAC_K1_TRACE_BUILD
#define K1_TRACE(i) do { if (p.debug && blockIdx.x == 0 && threadIdx.x == 0 && (i) < 64) g_k1_trace[i] = clock64(); } while (0)
#define K1_TRACE_NEXT() do { K1_TRACE(trace_i); ++trace_i; } while (0)
// per-warp wall-clock stamps (globaltimer, ns) of every CTA: 0 entry, 1 tables ready, 2 first samples arrived,
// 3 first frame done, 4 second frame done, 5 last frame done, 6 frames processed
__device__ unsigned long long g_k1_stamps[160 * 16 * 8];
__device__ __forceinline__ unsigned long long k1_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define K1_STAMP(i) do { if (p.debug && (threadIdx.x & 31) == 0 && blockIdx.x < 160) g_k1_stamps[(blockIdx.x * 16 + (threadIdx.x >> 5)) * 8 + (i)] = k1_now(); } while (0)
#else
#define K1_TRACE(i) do { } while (0)
#define K1_TRACE_NEXT() do { } while (0)
#define K1_STAMP(i) do { } while (0)
#endif

template <int OUT_MODE_T, int PMODE>
__global__ void __launch_bounds__(kFastThreads, 1) stft2048_kernel(const StftParams p) {
  constexpr bool kPeers = OUT_MODE_T == OUT_MEL_FUSED_PEERS;
  constexpr int OUT_MODE = kPeers ? (int)OUT_MEL_FUSED : OUT_MODE_T;      // same body, only the band stores differ
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // tables are stored pair-interleaved, [j / 2][lane][j % 2], so that one LDS.128 fetches the entries of two
  // consecutive register indices of a lane (conflict free: consecutive lanes are 16 bytes apart)
  float2* s_win = reinterpret_cast<float2*>(smem_raw);      // (w[2n], w[2n+1]) * 0.5 * scale, n = lane + 32 r -> [r/2][lane][r%2]
  float2* s_tw1 = s_win + 1024;                             // W_1024^(n1 k2), k2 = lane                   -> [n1/2][lane][n1%2]
  float2* s_tw2 = s_tw1 + 1024;                             // W_2048^(32 k1 + lane)                      -> [k1/2][lane][k1%2]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_tw2 + 1024);
  float2* s_slab = reinterpret_cast<float2*>(s_bar + kFastWarps);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  K1_TRACE(0);
  K1_STAMP(0);

  // each warp owns its barrier and slab, so its first bulk copy can be issued before the tables are built: the
  // copy's latency (1.5 us from a cold HBM page, up to 7 us for the unluckiest of 2368 simultaneous copies) then
  // overlaps the table set-up instead of following it
  uint64_t* bar = s_bar + warp;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
#ifdef TAC_K1_TRACE_BUILD
  int trace_i = 2;
  int stamp_frames = 0;
#endif

  float2* slab = s_slab + warp * kSlabComplex;
  float* slab_f = reinterpret_cast<float*>(slab);
  // A CTA owns a contiguous chunk of the launch's frames (sizes differ by at most one) and deals it round-robin to
  // its warps.  (Dealing frame g to warp g % (16 gridDim) made 68 SMs run 9 rounds and 80 SMs 8 at config 2: the
  // slowest warp finished 15 us after the fastest, measured with the globaltimer stamps of the timing build.)
  constexpr uint32_t step = kFastWarps;                      // frame indices fit 31 bits (checked by the host)
  const float half_power = 0.5f * p.power;
  uint32_t parity = 0;
  // OUT_MEL_FUSED: this warp's power-spectrum stash, row k1 = bins 32 k1 .. 32 k1 + 31, stride 33
  float* stash = reinterpret_cast<float*>(s_slab + kFastWarps * kSlabComplex) + warp * kStashFloats;
  if constexpr (OUT_MODE == OUT_MEL_FUSED) {
    if (lane == 0) stash[kStashZero] = 0.0f;
    __syncwarp();
  }

  // (sequence, frame-in-sequence) of this warp's frames advance incrementally: a 64-bit division per frame
  // costs ~100 instructions in the hot loop (it did, twice per frame, in the first version)
  const uint32_t frames_u = (uint32_t)p.frames;
  const uint32_t step_seq = step / frames_u, step_t = step % frames_u;
  auto advance = [&](uint32_t& seq, uint32_t& t) {
    seq += step_seq;
    t += step_t;
    if (t >= frames_u) {
      t -= frames_u;
      ++seq;
    }
  };

  // a frame can use the bulk copy when it lies inside the sequence and everything is 16B aligned
  // Interior frames: the bulk copy is issued in the middle of the previous frame (slab just freed) and lands
  // while that frame finishes.  Frames that touch the padding are gathered instead, and that is deferred to the
  // end of the previous frame, when all 64 data registers are free and 32 loads per lane can be in flight.
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  FrameSpan span_cur, span_next;                                // which part of the slab the bulk copy fills
  span_cur.lo = 0; span_cur.hi = 2048; span_cur.bulk = false;
  span_next = span_cur;
  auto stage_bulk = [&](uint32_t seq, uint32_t t, FrameSpan& span) -> bool {   // true: bulk copy in flight; false: needs the gather
    const int start = (int)t * p.hop - p.pad;                   // 32-bit: n_samples + 2 n_fft < 2^31 (host check)
    span = frame_span<2048>(p, start);
    if (span.bulk && elect_one()) {
      fence_proxy_async();
      const uint32_t bytes = (uint32_t)(span.hi - span.lo) * sizeof(float);
      mbar_arrive_expect_tx(bar, bytes);
      bulk_g2s_hint(slab_f + span.lo, p.x + (int64_t)seq * p.seq_stride + (start + span.lo), bytes, bar, pol_stream);
    }
    return span.bulk;
  };
  auto stage_gather = [&](uint32_t seq, uint32_t t) {
    const int start = (int)t * p.hop - p.pad;
    gather_padded<64, 32>(slab_f, p.x + (int64_t)seq * p.seq_stride, start, (int)p.n_samples, p.pad_mode, lane);
  };

  const uint32_t n_all = (uint32_t)(p.g1 - p.g0);            // frames of this launch; gi indexes them
  const uint32_t per_cta = n_all / gridDim.x, extra = n_all % gridDim.x;
  const uint32_t chunk0 = blockIdx.x * per_cta + (blockIdx.x < extra ? blockIdx.x : extra);
  const uint32_t n_launch = chunk0 + per_cta + (blockIdx.x < extra ? 1u : 0u);          // end of this CTA's chunk
  uint32_t gi = chunk0 + warp;
  const int64_t g_first = p.g0 + gi;
  uint32_t seq = (uint32_t)(g_first / p.frames), t = (uint32_t)(g_first % p.frames);      // once per warp
  bool in_flight = false;
  if (gi < n_launch) {
    in_flight = stage_bulk(seq, t, span_cur);
    if (!in_flight) stage_gather(seq, t);
  }
  fft2048_tables(p, s_win, s_tw1, s_tw2, tid, kFastThreads);
  __syncthreads();
  K1_TRACE(1);
  K1_STAMP(1);

#pragma unroll 1
  for (; gi < n_launch; gi += step) {
    if (in_flight) {
      mbar_wait(bar, parity);
      parity ^= 1u;
      fill_padding<2048>(slab_f, span_cur, p.pad_mode, lane, p.x + (int64_t)seq * p.seq_stride, (int)t * p.hop - p.pad);
    } else {
      __syncwarp();
    }

    K1_TRACE_NEXT();                               // sample data arrived
#ifdef TAC_K1_TRACE_BUILD
    if (p.debug && blockIdx.x == 0 && lane == 0 && g_k1_trace[40 + warp] == 0) g_k1_trace[40 + warp] = clock64();
    if (stamp_frames == 0) K1_STAMP(2);
#endif
    float2 v[32];
    fft2048_front(v, slab, s_win, s_tw1, lane);    // slab free again afterwards: prefetch the next frame
    K1_TRACE_NEXT();                               // front half done
    uint32_t seq_next = seq, t_next = t;
    advance(seq_next, t_next);
    const bool has_next = gi + step < n_launch;
    in_flight = has_next ? stage_bulk(seq_next, t_next, span_next) : false;
    // ---- pass 2: 32-point FFT over n1 for fixed k2 = lane ---------------------------------------------
    dit_fft_fma<32>(v);
    // now v[bit_reverse(k1)] = Z[32 k1 + lane] / 2

    // ---- real-FFT untangling: bin k = 32 k1 + lane pairs with 1024 - k ------------------------------
    // destination of bin `lane` of this frame; bins 32 k1 + lane follow at a fixed stride
    float* dst;
    int64_t dst_stride = 0;                        // floats between consecutive k1 (public layouts only)
    if constexpr (OUT_MODE == OUT_POWER_ROWS) {
      dst = p.out + power_tile_index(gi, lane, p.kpad);            // + k1 * 4096 floats: immediate offsets below
    } else if constexpr (OUT_MODE == OUT_MEL_FUSED) {
      dst = stash + lane;                                          // + k1 * 33 floats
    } else if constexpr (OUT_MODE == OUT_POWER_PUBLIC) {
      dst = p.out + ((int64_t)seq * p.bins + lane) * p.frames + t;
      dst_stride = 32 * p.frames;
    } else {
      dst = p.out + 2 * (((int64_t)seq * p.bins + lane) * p.frames + t);
      dst_stride = 64 * p.frames;
    }
    const int partner = (32 - lane) & 31;
    float4 tw2_pair;
    if constexpr (OUT_MODE == OUT_MEL_FUSED) {
      // Only |X|^p is needed, so bins k and 1024 - k are produced together from the pair (Z[k], Z[1024 - k]) that
      // one shuffle exchange brings together: with E = (a, b), T = W^k O,  X[k] = E + T and X[1024 - k] = conj(E - T).
      // Steps k1 = 0 .. 15 cover every bin but 512 (its own mirror): half the shuffles, 10.5 instead of 15
      // instructions per bin.  The mirror of bin 32 k1 + lane sits in stash row 31 - k1, column 32 - lane
      // (lane 0: row 32 - k1, column 0, i.e. 33 floats after the row start; k1 = 0 gives the Nyquist slot).
      // (lane 0's mirror, column 0 of row 32 - k1 >= 17, goes one float lower: bank 31 - k1, the one bank the other
      // lanes' stores of this step leave free; at the natural place it shared a bank with lane 31's)
      float* dmir = stash + (lane == 0 ? kStashStride - 1 : 32 - lane);
      const float2 tw2_base = s_tw2[2 * lane];                     // W_2048^lane (table row k1 = 0)
      auto emit2 = [&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        const float2 w = fft2048_tw2_computed<k1>(tw2_base);
        const float2 z = v[bit_reverse<32>(k1)];
        float2 q;
        q.x = __shfl_sync(0xffffffffu, v[bit_reverse<32>(31 - k1)].x, partner);
        q.y = __shfl_sync(0xffffffffu, v[bit_reverse<32>(31 - k1)].y, partner);
        if (lane == 0) q = v[bit_reverse<32>((32 - k1) & 31)];
        const float a = z.x + q.x, b = z.y - q.y, gs = z.y + q.y, h = q.x - z.x;
        const float tr = fmaf(w.x, gs, -w.y * h), ti = fmaf(w.x, h, w.y * gs);
        dst[k1 * kStashStride] = fast_power<PMODE>(a + tr, b + ti, half_power);
        dmir[(31 - k1) * kStashStride] = fast_power<PMODE>(a - tr, b - ti, half_power);
      };
      static_for<16>(emit2);
      const float2 x512 = fft2048_untangle<16>(v, fft2048_tw2_computed<16>(tw2_base), lane, partner);
      if (lane == 0) dst[16 * kStashStride] = fast_power<PMODE>(x512.x, x512.y, half_power);
    } else {
      auto emit = [&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        const float2 x = fft2048_untangle<k1>(v, fft2048_tw2<k1>(s_tw2, lane, tw2_pair), lane, partner);
        if constexpr (OUT_MODE == OUT_POWER_ROWS) {
          st_global_hint(dst + k1 * 4096, fast_power<PMODE>(x.x, x.y, half_power), pol_keep);
        } else if constexpr (OUT_MODE == OUT_MEL_FUSED) {
          dst[k1 * kStashStride] = fast_power<PMODE>(x.x, x.y, half_power);
        } else if constexpr (OUT_MODE == OUT_COMPLEX_PUBLIC) {
          *reinterpret_cast<float2*>(dst) = x;
          dst += dst_stride;
        } else {
          *dst = fast_power<PMODE>(x.x, x.y, half_power);
          dst += dst_stride;
        }
      };
      static_for<32>(emit);
    }
    // Nyquist bin (and zero fill of the row padding in frame-major mode); dst now points at bin 1024 + lane
    {
      const float2 z0 = v[0];
      const float nyq = 2.0f * (z0.x - z0.y);
      if constexpr (OUT_MODE == OUT_POWER_ROWS) {
        st_global_hint(dst + 32 * 4096, (lane == 0) ? fast_power<PMODE>(nyq, 0.0f, half_power) : 0.0f, pol_keep);   // slice 32: Nyquist + zero fill
      } else if constexpr (OUT_MODE == OUT_MEL_FUSED) {
        (void)nyq;                                   // bin 1024 came out of step k1 = 0 as the mirror of bin 0
      } else if constexpr (OUT_MODE == OUT_POWER_PUBLIC) {
        if (lane == 0) *dst = fast_power<PMODE>(nyq, 0.0f, half_power);
      } else {
        if (lane == 0) *reinterpret_cast<float2*>(dst) = make_float2(nyq, 0.0f);
      }
    }
    if constexpr (OUT_MODE == OUT_MEL_FUSED) band_contract<kPeers>(p, stash, lane, seq, t);
    K1_TRACE_NEXT();                               // frame done
#ifdef TAC_K1_TRACE_BUILD
    if (stamp_frames == 0) K1_STAMP(3);
    if (stamp_frames == 1) K1_STAMP(4);
    K1_STAMP(5);
    ++stamp_frames;
    if (p.debug && lane == 0 && blockIdx.x < 160) g_k1_stamps[(blockIdx.x * 16 + warp) * 8 + 6] = stamp_frames;
#endif
    if (has_next && !in_flight) stage_gather(seq_next, t_next);
    seq = seq_next;
    t = t_next;
    span_cur = span_next;
  }
}

// ---------------------------------------------------------------------------------------------
// Backward of stft + |.|^p for n_fft = 2048: d loss / d (windowed frame) from d loss / d |X|^p, one warp
// per frame with the machinery of the forward kernel.  `gspec` is frame-major, kpad floats per frame (the filterbank
// adjoint writes it that way); the windowed frame gradient (2048 floats) goes to `frames_out`, which
// overlap_add_kernel (stft_backward.cu) folds into the waveform gradient.
//
// With P = Zh[k], Q = conj(Zh[1024 - k]) (Zh = the half-scaled complex FFT the forward pass builds X from:
// X_k = S + D, conj X_{1024-k} = S - D, S = P + Q, D = -i W^k (P - Q)), per-bin gains h_k (= g_k (p/2) |X_k|^(p-2)) and
// sigma = h_k + h_{1024-k}, delta = h_k - h_{1024-k}, the spectrum whose inverse complex FFT carries the frame gradient
// in (even, odd) sample pairs collapses to
//     Zt_k = 2 [ (sigma + delta Im W^k) P + i delta Re W^k Q ]                 (k = 0: twice that, H_0 = G_0 not G_0 / 2)
// i.e. untangling, gain and re-tangling are one two-term update per bin, in place.  The inverse FFT is
// conj . FFT . conj with the same two register passes and one transposition through the slab.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fft1024_mid(float2 (&u)[32], float2* slab, const float2* s_tw1, int lane) {
  dit_fft_fma<32>(u);
#pragma unroll
  for (int k2 = 0; k2 < 32; ++k2) slab[k2 * kSlabStride + lane] = u[bit_reverse<32>(k2)];
  __syncwarp();
#pragma unroll
  for (int n1 = 0; n1 < 32; n1 += 2) {
    const float4 a = *reinterpret_cast<const float4*>(slab + lane * kSlabStride + n1);
    const float4 w = reinterpret_cast<const float4*>(s_tw1)[(n1 >> 1) * 32 + lane];
    u[n1] = make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
    u[n1 + 1] = make_float2(fmaf(a.z, w.z, -a.w * w.w), fmaf(a.z, w.w, a.w * w.z));
  }
  __syncwarp();
}

// (p / 2) |X|^(p - 2) from |X|^2
template <int PMODE>
__device__ __forceinline__ float power_gain(float n2, float power) {
  if constexpr (PMODE == 2) return 1.0f;
  if constexpr (PMODE == 1) return n2 > 0.0f ? 0.5f * rsqrtf(n2) : 0.0f;
  return n2 > 0.0f ? 0.5f * power * exp2f((0.5f * power - 1.0f) * __log2f(n2)) : 0.0f;
}

template <int PMODE>
__global__ void __launch_bounds__(kFastThreads, 1)
stft2048_backward_kernel(const StftParams p, const float* __restrict__ gspec, float* __restrict__ frames_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* s_win = reinterpret_cast<float2*>(smem_raw);
  float2* s_tw1 = s_win + 1024;
  float2* s_tw2 = s_tw1 + 1024;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_tw2 + 1024);
  float2* s_slab = reinterpret_cast<float2*>(s_bar + kFastWarps);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t* bar = s_bar + warp;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  float2* slab = s_slab + warp * kSlabComplex;
  float* slab_f = reinterpret_cast<float*>(slab);
  float* grow = reinterpret_cast<float*>(s_slab + kFastWarps * kSlabComplex) + warp * kStashFloats;   // this frame's gspec row
  const uint64_t pol_stream = l2_policy_evict_first();
  constexpr uint32_t kRowBytes = 1056 * sizeof(float);

  const uint32_t n_all = (uint32_t)(p.g1 - p.g0);
  const uint32_t per_cta = n_all / gridDim.x, extra = n_all % gridDim.x;
  const uint32_t chunk0 = blockIdx.x * per_cta + (blockIdx.x < extra ? blockIdx.x : extra);
  const uint32_t chunk1 = chunk0 + per_cta + (blockIdx.x < extra ? 1u : 0u);
  const uint32_t frames_u = (uint32_t)p.frames;
  uint32_t gi = chunk0 + warp;
  const int64_t g_first = p.g0 + gi;
  uint32_t seq = (uint32_t)(g_first / p.frames), t = (uint32_t)(g_first % p.frames);
  FrameSpan span;
  // samples (bulk part) and the gradient row land on the same barrier; samples the bulk copy cannot take are gathered
  auto stage = [&](uint32_t g_idx, uint32_t sq, uint32_t tt) {
    const int start = (int)tt * p.hop - p.pad;
    span = frame_span<2048>(p, start);
    __syncwarp();
    if (elect_one()) {
      fence_proxy_async();
      const uint32_t bytes = span.bulk ? (uint32_t)(span.hi - span.lo) * sizeof(float) : 0u;
      mbar_arrive_expect_tx(bar, bytes + kRowBytes);
      if (span.bulk) bulk_g2s_hint(slab_f + span.lo, p.x + (int64_t)sq * p.seq_stride + (start + span.lo), bytes, bar, pol_stream);
      bulk_g2s_hint(grow, gspec + (int64_t)g_idx * 1056, kRowBytes, bar, pol_stream);
    }
    if (!span.bulk) gather_padded<64, 32>(slab_f, p.x + (int64_t)sq * p.seq_stride, start, (int)p.n_samples, p.pad_mode, lane);
  };
  if (gi < chunk1) stage(gi, seq, t);
  fft2048_tables(p, s_win, s_tw1, s_tw2, tid, kFastThreads);
  __syncthreads();

  uint32_t parity = 0;
  const int partner = (32 - lane) & 31;
#pragma unroll 1
  for (; gi < chunk1; gi += kFastWarps) {
    mbar_wait(bar, parity);
    parity ^= 1u;
    if (span.bulk) fill_padding<2048>(slab_f, span, p.pad_mode, lane, p.x + (int64_t)seq * p.seq_stride, (int)t * p.hop - p.pad);
    else __syncwarp();

    float2 v[32];
    fft2048_front(v, slab, s_win, s_tw1, lane);
    dit_fft_fma<32>(v);                            // v[bit_reverse(k1)] = Zh[32 k1 + lane]

    // ---- in place: Zh -> conj(Zt).  Step (k1, 31 - k1) reads registers k1 and 31 - k1 of this lane and of lane
    // 32 - lane; lane 0 pairs with its own registers 32 - k1 and k1 + 1 instead, the first of which the previous step
    // has already overwritten: it travels in `carry`.
    float2 carry = v[bit_reverse<32>(0)];          // lane 0, step 0: partner of bin 0 is bin 0 itself
    auto update = [&](float2 z, float2 q, float gk, float gm, float2 w, bool first) -> float2 {
      float hk = gk, hm = gm;
      if constexpr (PMODE != 2) {
        const float sx = z.x + q.x, sy = z.y - q.y, ax = z.x - q.x, ay = z.y + q.y;        // S, P - Q
        const float dx = fmaf(w.x, ay, w.y * ax), dy = -fmaf(w.x, ax, -w.y * ay);          // D = -i W (P - Q)
        const float xr = sx + dx, xi = sy + dy, mr = sx - dx, mi = sy - dy;
        hk *= power_gain<PMODE>(fmaf(xr, xr, xi * xi), p.power);
        hm *= power_gain<PMODE>(fmaf(mr, mr, mi * mi), p.power);
      }
      const float sigma = hk + hm, delta = hk - hm;
      const float a = fmaf(delta, w.y, sigma), b = delta * w.x;
      float zr = fmaf(a, z.x, b * q.y), zi = fmaf(a, z.y, b * q.x);
      if (first) {
        zr *= 2.0f;
        zi *= 2.0f;
      }
      return make_float2(zr, -zi);                 // conj(Zt) / 2: the factor 2 (and the window's) rides on the gains
    };
    auto pair_step = [&](auto k1c) {
      constexpr int k1 = decltype(k1c)::value;     // 0 .. 15, mirror register 31 - k1
      constexpr int ra = bit_reverse<32>(k1), rb = bit_reverse<32>(31 - k1);
      const float2 za = v[ra], zb = v[rb];
      float2 qa, qb;
      qa.x = __shfl_sync(0xffffffffu, zb.x, partner);
      qa.y = __shfl_sync(0xffffffffu, zb.y, partner);
      qb.x = __shfl_sync(0xffffffffu, za.x, partner);
      qb.y = __shfl_sync(0xffffffffu, za.y, partner);
      if (lane == 0) {
        qa = carry;                                // register (32 - k1) & 31
        qb = v[bit_reverse<32>(k1 + 1)];           // register 32 - (31 - k1)
        carry = zb;                                // register 31 - k1 = 32 - (k1 + 1): the next step's partner
      }
      const int ka = 32 * k1 + lane, kb = 32 * (31 - k1) + lane;
      const float ga = 4.0f * grow[ka], gam = 4.0f * grow[1024 - ka];
      const float gb = 4.0f * grow[kb], gbm = 4.0f * grow[1024 - kb];
      const float2 wa = s_tw2[((k1 >> 1) * 32 + lane) * 2 + (k1 & 1)];                      // W_2048^ka (pair-interleaved table)
      const float2 wb = s_tw2[(((31 - k1) >> 1) * 32 + lane) * 2 + ((31 - k1) & 1)];        // W_2048^kb
      v[ra] = update(za, qa, ga, gam, wa, k1 == 0 && lane == 0);
      v[rb] = update(zb, qb, gb, gbm, wb, false);
    };
    static_for<16>(pair_step);
    __syncwarp();                                  // gradient row consumed

    float2 u[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) u[r] = v[bit_reverse<32>(r)];      // natural order for the next pass (renaming only)
    fft1024_mid(u, slab, s_tw1, lane);             // slab and row buffer free: fetch the next frame
    uint32_t seq_next = seq, t_next = t + kFastWarps;
    while (t_next >= frames_u) {
      t_next -= frames_u;
      ++seq_next;
    }
    const bool has_next = gi + kFastWarps < chunk1;
    if (has_next) stage(gi + kFastWarps, seq_next, t_next);
    dit_fft_fma<32>(u);                            // u[bit_reverse(m1)] = Y[32 m1 + lane], frame samples = conj(Y) pairs

    float2* frow = reinterpret_cast<float2*>(frames_out + (int64_t)gi * 2048) + lane;
#pragma unroll
    for (int m1 = 0; m1 < 32; m1 += 2) {
      const float4 w = reinterpret_cast<const float4*>(s_win)[(m1 >> 1) * 32 + lane];       // window * scale / 2
      const float2 y0 = u[bit_reverse<32>(m1)], y1 = u[bit_reverse<32>(m1 + 1)];
      __stcs(frow + 32 * m1, make_float2(y0.x * w.x, -y0.y * w.y));
      __stcs(frow + 32 * (m1 + 1), make_float2(y1.x * w.z, -y1.y * w.w));
    }
    seq = seq_next;
    t = t_next;
  }
}

int launch_stft2048_backward(const StftParams& p, const float* gspec_fm, float* frames_out, cudaStream_t stream) {
  const int64_t n_frames = p.g1 - p.g0;
  if (n_frames <= 0) return TAC_OK;
  TAC_REQUIRE(p.n_fft == 2048 && p.onesided, TAC_ERR_UNSUPPORTED, "stft2048_backward: n_fft = 2048, onesided only");
  TAC_REQUIRE((reinterpret_cast<uintptr_t>(gspec_fm) & 15) == 0, TAC_ERR_INVALID, "stft2048_backward: gradient rows must be 16-byte aligned");
  int64_t want = (n_frames + kFastWarps - 1) / kFastWarps;
  const int grid = (int)(want < sm_count() ? want : sm_count());
  using Kernel = void (*)(const StftParams, const float*, float*);
  Kernel k = p.power_mode == 2 ? stft2048_backward_kernel<2> : (p.power_mode == 1 ? stft2048_backward_kernel<1> : stft2048_backward_kernel<0>);
  TAC_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemBytes));
  LaunchProbe probe(KIND_STFT, stream);
  k<<<grid, kFastThreads, kFusedSmemBytes, stream>>>(p, gspec_fm, frames_out);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

// n_fft = 2048 with the reference's public output layouts, (n_seq, 1025, frames[, 2]) with time innermost.
// A frame's bins are 4 * frames bytes apart there, so per-frame stores would scatter 4-byte writes over 1025
// rows.  Instead the CTA works on tiles of 16 consecutive frames of one sequence (warp w = frame t0 + w), the
// 16 warps drop their spectra into a shared [bin][16] tile that re-uses the 16 slabs, and all 512 threads write
// the tile out as 64-byte (power) / 128-byte (complex) row segments.  Costs three CTA barriers per tile and the
// load/compute overlap of the power-tile kernel; buys 4-8x fewer store transactions.
template <int OUT_MODE, int PMODE>
__global__ void __launch_bounds__(kFastThreads, 1) stft2048_public_kernel(const StftParams p) {
  static_assert(OUT_MODE == OUT_COMPLEX_PUBLIC || OUT_MODE == OUT_POWER_PUBLIC, "public layouts only");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* s_win = reinterpret_cast<float2*>(smem_raw);
  float2* s_tw1 = s_win + 1024;
  float2* s_tw2 = s_tw1 + 1024;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_tw2 + 1024);
  float2* s_slab = reinterpret_cast<float2*>(s_bar + kFastWarps);
  constexpr int kRow = 17;                                  // staged tile row stride (elements): odd -> conflict free

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  fft2048_tables(p, s_win, s_tw1, s_tw2, tid, kFastThreads);
  uint64_t* bar = s_bar + warp;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  float2* slab = s_slab + warp * kSlabComplex;
  float* slab_f = reinterpret_cast<float*>(slab);
  const float half_power = 0.5f * p.power;
  const uint32_t frames_u = (uint32_t)p.frames;
  const uint32_t tiles_per_seq = (frames_u + kFastWarps - 1) / kFastWarps;
  const int64_t n_tiles = p.n_seq * (int64_t)tiles_per_seq;
  const int partner = (32 - lane) & 31;
  uint32_t parity = 0;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t seq = (uint32_t)(tile / tiles_per_seq);
    const uint32_t t0 = (uint32_t)(tile - (int64_t)seq * tiles_per_seq) * kFastWarps;
    const uint32_t t = t0 + warp;
    const bool active = t < frames_u;                       // warp-uniform
    float2 v[32];
    if (active) {
      const int64_t start = (int64_t)t * p.hop - p.pad;
      const float* row = p.x + (int64_t)seq * p.seq_stride;
      const FrameSpan span = frame_span<2048>(p, start);
      if (span.bulk) {
        if (elect_one()) {
          fence_proxy_async();
          const uint32_t bytes = (uint32_t)(span.hi - span.lo) * sizeof(float);
          mbar_arrive_expect_tx(bar, bytes);
          bulk_g2s(slab_f + span.lo, row + (start + span.lo), bytes, bar);
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        fill_padding<2048>(slab_f, span, p.pad_mode, lane, row, start);
      } else {
        gather_padded<64, 32>(slab_f, row, (int)start, (int)p.n_samples, p.pad_mode, lane);
        __syncwarp();
      }
      fft2048_front(v, slab, s_win, s_tw1, lane);
      dit_fft_fma<32>(v);
    }
    __syncthreads();                                        // every slab is free: together they hold the output tile

    if (active) {
      float4 tw2_pair;
      auto stash = [&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        const float2 x = fft2048_untangle<k1>(v, fft2048_tw2<k1>(s_tw2, lane, tw2_pair), lane, partner);
        const int idx = (32 * k1 + lane) * kRow + warp;
        if constexpr (OUT_MODE == OUT_COMPLEX_PUBLIC) s_slab[idx] = x;
        else reinterpret_cast<float*>(s_slab)[idx] = fast_power<PMODE>(x.x, x.y, half_power);
      };
      static_for<32>(stash);
      if (lane == 0) {
        const float nyq = 2.0f * (v[0].x - v[0].y);
        if constexpr (OUT_MODE == OUT_COMPLEX_PUBLIC) s_slab[1024 * kRow + warp] = make_float2(nyq, 0.0f);
        else reinterpret_cast<float*>(s_slab)[1024 * kRow + warp] = fast_power<PMODE>(nyq, 0.0f, half_power);
      }
    }
    __syncthreads();

    // write-out: 16 threads per bin row, rows of this tile are `n_valid` consecutive frames
    const uint32_t n_valid = min((uint32_t)kFastWarps, frames_u - t0);
    const int f = tid & 15;
    const int64_t out_base = ((int64_t)seq * 1025) * p.frames + t0 + f;
    if (f < (int)n_valid) {
      for (int bin = tid >> 4; bin < 1025; bin += kFastThreads / 16) {
        const int64_t o = out_base + (int64_t)bin * p.frames;
        if constexpr (OUT_MODE == OUT_COMPLEX_PUBLIC) __stcs(reinterpret_cast<float2*>(p.out) + o, s_slab[bin * kRow + f]);
        else __stcs(p.out + o, reinterpret_cast<const float*>(s_slab)[bin * kRow + f]);
      }
    }
    __syncthreads();                                        // tile consumed: slabs may be refilled
  }
}

// ---------------------------------------------------------------------------------------------
// generic path: one CTA per frame, Stockham radix-2 in shared memory
// ---------------------------------------------------------------------------------------------
constexpr int kGenThreads = 256;

static size_t generic_smem_bytes(int n_fft) {
  const size_t c = (size_t)n_fft / 2;
  return sizeof(float2) * (2 * c + c / 2 + c + 1) + sizeof(float) * (size_t)n_fft + 16;
}

__global__ void __launch_bounds__(kGenThreads) stft_generic_kernel(const StftParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n_fft = p.n_fft, C = n_fft >> 1;
  float2* buf_a = reinterpret_cast<float2*>(smem_raw);
  float2* buf_b = buf_a + C;
  float2* tw_c = buf_b + C;               // W_C^m, m < C/2
  float2* tw_n = tw_c + (C >> 1);         // W_N^k, k <= C
  float* win = reinterpret_cast<float*>(tw_n + C + 1);
  const int tid = threadIdx.x;

  for (int i = tid; i < n_fft; i += kGenThreads) win[i] = p.window[i] * (0.5f * p.scale);
  for (int i = tid; i < (C >> 1); i += kGenThreads) {
    float sn, cs;
    sincospif(-2.0f * (float)i / (float)C, &sn, &cs);
    tw_c[i] = make_float2(cs, sn);
  }
  for (int i = tid; i <= C; i += kGenThreads) {
    float sn, cs;
    sincospif(-2.0f * (float)i / (float)n_fft, &sn, &cs);
    tw_n[i] = make_float2(cs, sn);
  }
  __syncthreads();

  for (int64_t g = p.g0 + blockIdx.x; g < p.g1; g += gridDim.x) {
    const int64_t seq = g / p.frames, t = g - seq * p.frames, start = t * p.hop - p.pad;
    const float* row = p.x + seq * p.seq_stride;
    for (int n = tid; n < C; n += kGenThreads) {
      const float x0 = fetch_padded(row, start + 2 * n, p.n_samples, p.pad_mode);
      const float x1 = fetch_padded(row, start + 2 * n + 1, p.n_samples, p.pad_mode);
      buf_a[n] = make_float2(x0 * win[2 * n], x1 * win[2 * n + 1]);
    }
    __syncthreads();
    const float2* src = stockham_forward<kGenThreads>(buf_a, buf_b, tw_c, C, tid);      // fft_stockham.cuh
    const int64_t out_row = g - p.g0;
    for (int k = tid; k <= C; k += kGenThreads) {
      const float2 z = src[k & (C - 1)];
      const float2 q = src[(C - k) & (C - 1)];
      const float a = z.x + q.x, b = z.y - q.y, gs = z.y + q.y, h = q.x - z.x;
      const float2 w = tw_n[k];
      const float xr = fmaf(w.x, gs, fmaf(-w.y, h, a));
      float xi = fmaf(w.x, h, fmaf(w.y, gs, b));
      if (k == 0 || k == C) xi = 0.0f;
      emit_bin(p, seq, t, out_row, k, xr, xi);
      if (!p.onesided && k > 0 && k < C) emit_bin(p, seq, t, out_row, n_fft - k, xr, -xi);
    }
    if (p.out_mode == OUT_POWER_ROWS)
      for (int k = C + 1 + tid; k < p.kpad; k += kGenThreads) p.out[power_tile_index(out_row, k, p.kpad)] = 0.0f;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// any other n_fft (not a power of two; e.g. 400 = 25 ms at 16 kHz, 1200, 441): direct DFT, one CTA per group of G frames.
// The reference passes any fft_length to torch.stft (functional.py:99); this is the path for the sizes the register /
// Stockham kernels do not cover.  Per frame and bin k:
//     X[k] = x[0] + (-1)^k x[N/2] + sum_{n=1}^{H} (x[n] + x[N-n]) cos(2 pi n k / N) - i (x[n] - x[N-n]) sin(2 pi n k / N),   H = (N-1)/2
// (the real input folded about N/2: half the multiply-adds).  A thread owns one bin for the G frames of the group: the
// folded samples (e, o) of the group sit in shared memory as [n][frame] so a 16-byte broadcast load serves two frames, the
// twiddle exp(-2 pi i n k / N) is read from a double-precision-built table every 16 steps (index n k mod N kept exactly)
// and rotated by exp(-2 pi i k / N) in between (16 steps: ~1e-6), which keeps the table gathers -- 32 different banks per
// warp -- off the inner loop.  fp32 accumulation over N/2 terms.
// ---------------------------------------------------------------------------------------------
constexpr int kDftThreads = 256;
constexpr int kDftResync = 16;
static size_t dft_smem_bytes(int n_fft, int g) {      // table (padded to 16 bytes), folded samples, unpaired samples
  return sizeof(float2) * (size_t)((n_fft + 1) & ~1) + sizeof(float2) * (size_t)g * (size_t)(n_fft / 2 + 1) + sizeof(float2) * (size_t)g + 64;
}

template <int G>
__global__ void __launch_bounds__(kDftThreads) stft_dft_kernel(const StftParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int N = p.n_fft, nb = N / 2 + 1, H = (N - 1) / 2;
  const bool even = (N & 1) == 0;
  float2* tab = reinterpret_cast<float2*>(smem_raw);                   // [N]  (cos, -sin)(2 pi j / N)
  float2* eo = tab + ((N + 1) & ~1);                                   // [H + 1][G]  (e, o) of n = 1..H (row 0 unused); 16-byte aligned
  float2* ends = eo + (size_t)(N / 2 + 1) * G;                         // [G]  (x[0], x[N/2] or 0)
  const int tid = threadIdx.x;
  for (int j = tid; j < N; j += kDftThreads) {
    double sn, cs;
    sincospi(-2.0 * (double)j / (double)N, &sn, &cs);
    tab[j] = make_float2((float)cs, (float)sn);
  }
  __syncthreads();
  const int64_t n_groups = (p.g1 - p.g0 + G - 1) / G;
  for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int64_t gbase = p.g0 + grp * G;
    for (int i = tid; i < G * (H + 1); i += kDftThreads) {            // stage: window, fold
      const int f = i % G, n = i / G;                                  // n = 0: the unpaired samples
      const int64_t g = gbase + f;
      float2 v = make_float2(0.0f, 0.0f);
      if (g < p.g1) {
        const int64_t seq = g / p.frames, t = g - seq * p.frames, start = t * p.hop - p.pad;
        const float* row = p.x + seq * p.seq_stride;
        if (n == 0) {
          v.x = fetch_padded(row, start, p.n_samples, p.pad_mode) * (p.window[0] * p.scale);
          v.y = even ? fetch_padded(row, start + N / 2, p.n_samples, p.pad_mode) * (p.window[N / 2] * p.scale) : 0.0f;
        } else {
          const float a = fetch_padded(row, start + n, p.n_samples, p.pad_mode) * (p.window[n] * p.scale);
          const float b = fetch_padded(row, start + N - n, p.n_samples, p.pad_mode) * (p.window[N - n] * p.scale);
          v = make_float2(a + b, a - b);
        }
      }
      if (n == 0) ends[f] = v; else eo[(size_t)n * G + f] = v;
    }
    __syncthreads();
    for (int k = tid; k < nb; k += kDftThreads) {
      float re[G], im[G];
#pragma unroll
      for (int f = 0; f < G; ++f) {
        const float2 e = ends[f];
        re[f] = e.x + ((k & 1) ? -e.y : e.y);
        im[f] = 0.0f;
      }
      const float2 rot = tab[k];                                       // exp(-2 pi i k / N)
      int idx = k;                                                     // (n k) mod N at n = 1
      const int step = (int)(((int64_t)kDftResync * k) % N);
      for (int n0 = 1; n0 <= H; n0 += kDftResync) {
        float2 w = tab[idx];
        const int n1 = min(n0 + kDftResync, H + 1);
        for (int n = n0; n < n1; ++n) {
          const float4* row4 = reinterpret_cast<const float4*>(eo + (size_t)n * G);
#pragma unroll
          for (int f2 = 0; f2 < G / 2; ++f2) {
            const float4 v = row4[f2];                                 // (e, o) of frames 2 f2, 2 f2 + 1
            re[2 * f2] = fmaf(v.x, w.x, re[2 * f2]);
            im[2 * f2] = fmaf(v.y, w.y, im[2 * f2]);
            re[2 * f2 + 1] = fmaf(v.z, w.x, re[2 * f2 + 1]);
            im[2 * f2 + 1] = fmaf(v.w, w.y, im[2 * f2 + 1]);
          }
          w = make_float2(fmaf(w.x, rot.x, -w.y * rot.y), fmaf(w.x, rot.y, w.y * rot.x));
        }
        idx += step;
        idx -= (idx >= N) ? N : 0;
      }
#pragma unroll
      for (int f = 0; f < G; ++f) {
        const int64_t g = gbase + f;
        if (g >= p.g1) continue;
        const int64_t seq = g / p.frames, t = g - seq * p.frames;
        const float xi = (k == 0 || 2 * k == N) ? 0.0f : im[f];
        emit_bin(p, seq, t, g - p.g0, k, re[f], xi);
        if (!p.onesided && k > 0 && 2 * k != N) emit_bin(p, seq, t, g - p.g0, N - k, re[f], -xi);
      }
    }
    if (p.out_mode == OUT_POWER_ROWS) {
      for (int f = 0; f < G; ++f) {
        const int64_t g = gbase + f;
        if (g < p.g1)
          for (int k = p.bins + tid; k < p.kpad; k += kDftThreads) p.out[power_tile_index(g - p.g0, k, p.kpad)] = 0.0f;
      }
    }
    __syncthreads();
  }
}

static int launch_stft_dft(const StftParams& p, cudaStream_t stream) {
  const int64_t n_frames = p.g1 - p.g0;
  const int g = p.n_fft <= 2048 ? 8 : 2;
  const size_t smem = dft_smem_bytes(p.n_fft, g);
  auto k = g == 8 ? stft_dft_kernel<8> : stft_dft_kernel<2>;
  if (smem > 48 * 1024) TAC_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_sm = (int)((200 * 1024) / (smem + 1024));
  const int64_t cap = (int64_t)sm_count() * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
  const int64_t groups = (n_frames + g - 1) / g;
  const int grid = (int)(groups < cap ? groups : cap);
  LaunchProbe probe(KIND_STFT, stream);
  k<<<grid, kDftThreads, smem, stream>>>(p);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

int dump_k1_trace() {
  long long h[64];
  TAC_CUDA_OK(cudaDeviceSynchronize());
#ifdef TAC_K1_TRACE_BUILD
  {
    static unsigned long long st[160 * 16 * 8];
    TAC_CUDA_OK(cudaMemcpyFromSymbol(st, g_k1_stamps, sizeof(st)));
    unsigned long long base = ~0ull;
    for (int w = 0; w < 160 * 16; ++w) if (st[w * 8] && st[w * 8] < base) base = st[w * 8];
    const char* names[6] = {"entry", "tables ready", "first samples", "frame 1 done", "frame 2 done", "last frame done"};
    for (int i = 0; i < 6; ++i) {
      double mn = 1e30, mx = 0, sum = 0;
      int n = 0;
      for (int w = 0; w < 160 * 16; ++w) {
        if (!st[w * 8] || !st[w * 8 + i]) continue;
        const double v = (double)(st[w * 8 + i] - base) * 1e-3;
        mn = v < mn ? v : mn; mx = v > mx ? v : mx; sum += v; ++n;
      }
      printf("stamp %-16s warps %4d  min %7.2f  mean %7.2f  max %7.2f us\n", names[i], n, mn, n ? sum / n : 0.0, mx);
    }
    int hist[16] = {0};
    double last_by_frames[16] = {0};
    for (int w = 0; w < 160 * 16; ++w) {
      if (!st[w * 8]) continue;
      const int f = (int)st[w * 8 + 6];
      if (f < 16) { ++hist[f]; const double v = (double)(st[w * 8 + 5] - base) * 1e-3; if (v > last_by_frames[f]) last_by_frames[f] = v; }
    }
    for (int f = 0; f < 16; ++f) if (hist[f]) printf("warps with %2d frames: %4d, latest finish %7.2f us\n", f, hist[f], last_by_frames[f]);
    // CTA 0: per-warp finish times
    printf("CTA 0 per-warp last-frame-done:");
    for (int w = 0; w < 16; ++w) printf(" %.1f", (double)(st[w * 8 + 5] - base) * 1e-3);
    printf("\n");
  }
#endif
  TAC_CUDA_OK(cudaMemcpyFromSymbol(h, g_k1_trace, sizeof(h)));
  printf("k1 trace (cycles since kernel entry of CTA 0):");
  for (int i = 0; i < 40; ++i) printf(" %lld", h[i] ? h[i] - h[0] : -1);
  printf("\nfirst data arrival per warp of CTA 0:");
  for (int i = 40; i < 56; ++i) printf(" %lld", h[i] ? h[i] - h[0] : -1);
  printf("\n");
  return TAC_OK;
}

static int g_mel_variant = -1;
static int mel_kernel_variant() {
  if (g_mel_variant < 0) {
    const char* e = getenv("TAC_MEL_VARIANT");                 // 0: pair kernel, 1: one frame per warp, 2: pair kernel + tcgen05 pass
    g_mel_variant = e ? atoi(e) : (getenv("TAC_MEL_SINGLE") ? 1 : 0);
    if (g_mel_variant < 0 || g_mel_variant > 2) g_mel_variant = 0;
  }
  return g_mel_variant;
}

int launch_stft(const StftParams& p, cudaStream_t stream) {
  const int64_t n_frames = p.g1 - p.g0;
  if (n_frames <= 0) return TAC_OK;
  if (!is_pow2(p.n_fft) || p.n_fft < 32) return launch_stft_dft(p, stream);
  if (p.onesided && (p.n_fft == 256 || p.n_fft == 512 || p.n_fft == 1024 || p.n_fft == 4096)) return launch_stft_warp(p, stream);
  if (p.n_fft == 2048 && p.onesided && (p.out_mode == OUT_MEL_FUSED || p.out_mode == OUT_MEL_FUSED_PEERS)) {
    // two frames per warp in packed fp32 pairs (stft_pair.cu); TAC_MEL_SINGLE=1 keeps the one-frame-per-warp kernel
    // below (A/B timing and the bit-equality test of the two)
    const int variant = mel_kernel_variant();
    if (variant == 2 && stft2048_pair_applies(p)) return launch_stft2048_pair_tc(p, stream);
    if (variant == 0 && stft2048_pair_applies(p)) return launch_stft2048_pair(p, stream);
  }
  if (p.n_fft == 2048 && p.onesided) {
    const bool whole = p.g0 == 0 && p.g1 == p.n_seq * p.frames;       // the tiled public kernel walks whole sequences
    int64_t want = (n_frames + kFastWarps - 1) / kFastWarps;
    if (whole && p.out_mode != OUT_POWER_ROWS) want = p.n_seq * ((p.frames + kFastWarps - 1) / kFastWarps);
    const int grid = (int)(want < sm_count() ? want : sm_count());
    using Kernel = void (*)(const StftParams);
    Kernel k = nullptr;
    if (p.out_mode == OUT_COMPLEX_PUBLIC) k = whole ? stft2048_public_kernel<OUT_COMPLEX_PUBLIC, 1> : stft2048_kernel<OUT_COMPLEX_PUBLIC, 1>;
    else if (p.out_mode == OUT_POWER_PUBLIC && whole)
      k = p.power_mode == 2 ? stft2048_public_kernel<OUT_POWER_PUBLIC, 2> : (p.power_mode == 1 ? stft2048_public_kernel<OUT_POWER_PUBLIC, 1> : stft2048_public_kernel<OUT_POWER_PUBLIC, 0>);
    else if (p.out_mode == OUT_POWER_PUBLIC)
      k = p.power_mode == 2 ? stft2048_kernel<OUT_POWER_PUBLIC, 2> : (p.power_mode == 1 ? stft2048_kernel<OUT_POWER_PUBLIC, 1> : stft2048_kernel<OUT_POWER_PUBLIC, 0>);
    else if (p.out_mode == OUT_MEL_FUSED)
      k = p.power_mode == 2 ? stft2048_kernel<OUT_MEL_FUSED, 2> : (p.power_mode == 1 ? stft2048_kernel<OUT_MEL_FUSED, 1> : stft2048_kernel<OUT_MEL_FUSED, 0>);
    else if (p.out_mode == OUT_MEL_FUSED_PEERS)
      k = p.power_mode == 2 ? stft2048_kernel<OUT_MEL_FUSED_PEERS, 2> : (p.power_mode == 1 ? stft2048_kernel<OUT_MEL_FUSED_PEERS, 1> : stft2048_kernel<OUT_MEL_FUSED_PEERS, 0>);
    else
      k = p.power_mode == 2 ? stft2048_kernel<OUT_POWER_ROWS, 2> : (p.power_mode == 1 ? stft2048_kernel<OUT_POWER_ROWS, 1> : stft2048_kernel<OUT_POWER_ROWS, 0>);
    const size_t smem = (p.out_mode == OUT_MEL_FUSED || p.out_mode == OUT_MEL_FUSED_PEERS) ? kFusedSmemBytes : kFastSmemBytes;
    TAC_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LaunchProbe probe(KIND_STFT, stream);
    k<<<grid, kFastThreads, smem, stream>>>(p);
  } else {
    const size_t smem = generic_smem_bytes(p.n_fft);
    if (smem > 48 * 1024)
      TAC_CUDA_OK(cudaFuncSetAttribute(stft_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = (int)((200 * 1024) / (smem + 1024));
    const int64_t cap = (int64_t)sm_count() * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
    const int grid = (int)(n_frames < cap ? n_frames : cap);
    LaunchProbe probe(KIND_STFT, stream);
    stft_generic_kernel<<<grid, kGenThreads, smem, stream>>>(p);
  }
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

int fill_stft_params(StftParams& p, const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                     const float* window, int n_fft, int hop, int center, int pad_mode, int normalized,
                     int onesided) {
  TAC_REQUIRE((x || n_seq == 0) && window, TAC_ERR_INVALID, "stft: null input or window pointer");   // an empty batch has no storage
  TAC_REQUIRE(n_seq >= 0 && n_samples >= 0 && seq_stride >= n_samples, TAC_ERR_INVALID,
              "stft: bad shape n_seq=%lld n_samples=%lld stride=%lld", (long long)n_seq, (long long)n_samples,
              (long long)seq_stride);
  TAC_REQUIRE(n_fft >= 2 && n_fft <= 8192, TAC_ERR_UNSUPPORTED, "stft: n_fft=%d outside [2, 8192] (the sizes the sm_100a kernels cover)", n_fft);
  TAC_REQUIRE(hop >= 1, TAC_ERR_INVALID, "stft: hop_length=%d must be positive", hop);
  TAC_REQUIRE(pad_mode >= TAC_PAD_REFLECT && pad_mode <= TAC_PAD_CIRCULAR, TAC_ERR_INVALID, "stft: unknown pad_mode %d", pad_mode);
  const int pad = center ? n_fft / 2 : 0;
  if (center && pad_mode == TAC_PAD_REFLECT)
    TAC_REQUIRE(pad < n_samples, TAC_ERR_INVALID,
                "stft: Padding size should be less than the corresponding input dimension (reflect pad %d, time %lld)", pad,
                (long long)n_samples);
  if (center && pad_mode == TAC_PAD_CIRCULAR)
    TAC_REQUIRE(pad <= n_samples, TAC_ERR_INVALID, "stft: circular padding %d wraps more than once (time %lld)", pad,
                (long long)n_samples);
  TAC_REQUIRE(n_samples + 2 * (int64_t)n_fft < ((int64_t)1 << 31), TAC_ERR_UNSUPPORTED,
              "stft: sequences of %lld samples exceed the 2^31 the kernels index", (long long)n_samples);
  TAC_REQUIRE(n_samples + 2 * pad >= n_fft, TAC_ERR_INVALID, "stft: input of %lld samples is shorter than n_fft=%d",
              (long long)n_samples, n_fft);
  p.x = x;
  p.window = window;
  p.n_seq = n_seq;
  p.n_samples = n_samples;
  p.seq_stride = seq_stride;
  p.n_fft = n_fft;
  p.hop = hop;
  p.pad = pad;
  p.pad_mode = pad_mode;
  p.onesided = onesided ? 1 : 0;
  p.bins = onesided ? n_fft / 2 + 1 : n_fft;
  p.frames = tac_stft_num_frames(n_samples, n_fft, hop, center);
  TAC_REQUIRE(n_seq * p.frames < ((int64_t)1 << 31) && p.frames < ((int64_t)1 << 31), TAC_ERR_UNSUPPORTED,
              "stft: %lld frames in one call exceed the 2^31 the kernels index; split the batch", (long long)(n_seq * p.frames));
  p.scale = normalized ? (float)(1.0 / sqrt((double)n_fft)) : 1.0f;
  p.g0 = 0;
  p.g1 = n_seq * p.frames;
  p.kpad = kpad_for_bins(p.bins);
  p.bulk_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (seq_stride & 3) == 0 && (hop & 3) == 0 && (pad & 3) == 0) ? 1 : 0;
  p.power = 1.0f;
  p.power_mode = 1;
  p.out_mode = OUT_COMPLEX_PUBLIC;
  p.out = nullptr;
  static int debug = -1;
  if (debug < 0) debug = getenv("TAC_K1_TRACE") ? 1 : 0;
  p.debug = debug;
  p.band_plan = nullptr;
  p.band_cmax = p.n_bands = p.n_bands_pad = p.to_db = 0;
  p.band_fast = p.band_off_fast = 0;
  p.amin = 0.0f;
  p.log10_ref = 0.0f;
  p.out_seq_stride = p.out_t_stride = p.out_band_stride = 0;
  for (int q = 0; q < kMaxPeers; ++q) p.peer_out[q] = nullptr;
  p.n_peers = 0;
  p.peer_multicast = 0;
  p.peer_seq0 = 0;
  return TAC_OK;
}

}  // namespace tac

extern "C" int tac_mel_kernel_variant(int variant) {
  const int prev = tac::mel_kernel_variant();
  if (variant >= 0) tac::g_mel_variant = variant <= 2 ? variant : 0;
  return prev;
}

extern "C" int tac_debug_dump_k1_trace(void) { return tac::dump_k1_trace(); }

extern "C" int64_t tac_stft_num_frames(int64_t n_samples, int n_fft, int hop, int center) {
  if (hop <= 0 || n_fft <= 0) return 0;
  const int64_t padded = n_samples + (center ? 2 * (int64_t)(n_fft / 2) : 0);
  return padded < n_fft ? 0 : 1 + (padded - n_fft) / hop;
}

extern "C" int tac_stft_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride, const float* window,
                            int n_fft, int hop, int center, int pad_mode, int normalized, int onesided, float* out,
                            void* stream) {
  using namespace tac;
  StftParams p;
  const int rc = fill_stft_params(p, x, n_seq, n_samples, seq_stride, window, n_fft, hop, center, pad_mode, normalized, onesided);
  if (rc != TAC_OK) return rc;
  TAC_REQUIRE(out || p.g1 == 0, TAC_ERR_INVALID, "stft: null output pointer");
  p.out = out;
  p.out_mode = OUT_COMPLEX_PUBLIC;
  return launch_stft(p, as_stream(stream));
}

extern "C" int tac_spectrogram_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                                   const float* window, int n_fft, int hop, int center, int pad_mode, int normalized,
                                   int onesided, float power, float* out, void* stream) {
  using namespace tac;
  StftParams p;
  const int rc = fill_stft_params(p, x, n_seq, n_samples, seq_stride, window, n_fft, hop, center, pad_mode, normalized, onesided);
  if (rc != TAC_OK) return rc;
  TAC_REQUIRE(out || p.g1 == 0, TAC_ERR_INVALID, "spectrogram: null output pointer");
  p.out = out;
  p.out_mode = OUT_POWER_PUBLIC;
  p.power = power;
  p.power_mode = power == 2.0f ? 2 : (power == 1.0f ? 1 : 0);
  return launch_stft(p, as_stream(stream));
}
