// K3 (round 2): the whole mel chain in one kernel, TWO FRAMES PER WARP in packed fp32 pairs.
// Replaces, like stft2048_kernel<OUT_MEL_FUSED> (stft.cu), the reference chain
//   torch.stft (functional.py:99-107) -> complex_norm (:126-128) -> apply_filterbank (:183-184) [-> amplitude_to_db :291-296]
// for fft_length = 2048, hop <= 512 and two-adjacent-band (triangular) filterbanks.
//
// Why: the one-frame-per-warp kernel is bound by instruction issue (2 075 warp-instructions per frame, 65 % of them
// FADD / FMUL / FFMA; issue-active 62 %, fp32 pipe 41 %, profiles/r01_prof_melfused.json).  sm_100 has two-wide fp32
// instructions (FADD2 / FMUL2 / FFMA2, f32x2.cuh) that do the same flops per clock in half the issue slots.  Here a
// warp carries the consecutive frames t = 2 j (A) and 2 j + 1 (B) of one sequence side by side: every real quantity
// of the FFT is a `pk` = (value of A, value of B), so the butterflies, twiddles, untangling, |X|^2 and the band walk
// are the instruction stream of the one-frame kernel with each fp32 instruction doing both frames (1 234 instead of
// 2 075 warp-instructions per frame).  Window, twiddles, band weights, masks and all index arithmetic are shared by
// the two frames (scalar broadcast operands).
//
// Memory per warp (18 752 B of shared memory, 12 warps = 24 frames in flight per SM):
//   * the two frames overlap by 2048 - hop samples, so they arrive as ONE bulk copy of 2048 + hop samples of the
//     (padded) row: frame A is floats [0, 2048) of the region, frame B floats [hop, hop + 2048);
//   * the same bytes then serve as the 32 x 32 transposition buffer of PAIR-complex values (float4 = re_A, re_B,
//     im_A, im_B; row stride 33 float4: STS.128 rows / LDS.128 columns, both conflict free);
//   * the stash of power PAIRS (float2; the band plan's indices unchanged) sits behind the sample region, so the next
//     pair's samples can land while the stash is in use.
// The per-lane constant tables -- window (64 floats per lane), inter-pass twiddles W_1024^(n1 k2) (64) and the band
// weights (64) -- live in TENSOR MEMORY: a lane of TMEM is exactly a per-lane table, `tcgen05.ld` delivers 16 of its
// columns to the owning thread without touching the shared-memory pipe (LDTM in the SASS), and the 24 KB they took in
// shared memory are what lets 12 warps of buffers fit.
//
// TC = true (stft_pair_tc.cu, tac_mel_kernel_variant(2)): the SECOND 32-point pass runs on the tcgen05 tensor cores.
// After the transposition lane = k2 holds Y'[n1][k2] (twiddled) for n1 = 0..31 of both frames; that is row k2 of a
// (32 x 64) real matrix per frame, and pass 2 is  D[k2, (k1, re|im)] = sum_(n1, re|im) A[k2, (n1, re|im)] B[(n1, ..), (k1, ..)]
// with the constant 64 x 64 matrix B = real form of the 32-point DFT.  Four warps (one per lane quadrant of tensor
// memory) form a group: their 4 x 32 rows are one M = 128 operand per frame slot, fed FROM TENSOR MEMORY
// (tcgen05.st of the tf32 hi / lo split of the registers; no shared-memory traffic for A), B (hi, lo images, 32 KB,
// 128B-swizzled K-major) sits in shared memory, 3xTF32: D = A_lo B_hi + A_hi B_lo + A_hi B_hi, fp32 accumulate, which
// keeps the 1e-4 parity bar (single-pass tf32 does not).  tcgen05.ld hands lane k2 its 32 outputs Z[32 k1 + k2] in
// natural order and the untangling carries on unchanged.  Tensor memory holds ONE group's operands at a time
// (2 frames x (64 hi + 64 lo + 64 D) columns next to the window / twiddle tables), so the two groups of a CTA take
// turns (mbarrier hand-off), which also puts them in anti-phase on the schedulers they share.
#pragma once
#include <stdlib.h>

#include "bandplan.cuh"
#include "f32x2.cuh"
#include "fft_regs.cuh"
#include "stft_params.cuh"
#include "tac_common.cuh"

namespace tac {

#ifndef PAIR_WARPS
#define PAIR_WARPS 8
#endif
constexpr int kPairWarps = PAIR_WARPS;                           // experiments: scripts/build_variant.sh -DPAIR_WARPS=8
constexpr int kPairThreads = kPairWarps * 32;
static_assert(kPairWarps % 4 == 0, "whole lane quadrants of tensor memory");
constexpr int kPairMaxHop = 512;
constexpr int kRegionFloats = 2048 + kPairMaxHop;               // samples of frames A and B
constexpr int kXStride = 33;                                    // transposition row stride, float4 elements
constexpr int kXBufBytes = 32 * kXStride * 16;                  // 16 896 B
constexpr int kStashOffBytes = kRegionFloats * 4;               // 10 240: behind the sample region
constexpr int kWarpBytes = kStashOffBytes + kStashFloats * 8;   // 18 752
static_assert(kWarpBytes >= kXBufBytes && kWarpBytes % 16 == 0, "transposition buffer inside the warp's block");
// shared memory: [band-weight table (TC only)] [W_2048^lane] [mbarriers] [per-warp blocks] [TC only: B images, 1 KB aligned]
constexpr size_t kPairBandTabBytes = 4 * 4 * 32 * sizeof(float4);              // 8 KB: [chunk][quarter][lane] float4
constexpr size_t kTcSlabBytes = 64 * 128;                                       // 64 rows (n) x 32 tf32 (k): one swizzled K slab
constexpr size_t kTcBBytes = 4 * kTcSlabBytes;                                  // hi k<32, hi k>=32, lo k<32, lo k>=32
template <bool TC>
constexpr size_t pair_smem_bytes() {
  return (TC ? kPairBandTabBytes : 0) + 32 * sizeof(float2) + 16 * sizeof(uint64_t) + (size_t)kPairWarps * kWarpBytes +
         (TC ? 1024 + kTcBBytes : 0);
}
static_assert(pair_smem_bytes<false>() <= 227 * 1024 && (kPairWarps != 8 || pair_smem_bytes<true>() <= 227 * 1024), "pair kernel exceeds the shared memory of one CTA");
// tensor-memory columns: per-lane tables, and (TC) one group's pass-2 operands
constexpr uint32_t kColWin = 0, kColTw1 = 64, kColBand = 128;
constexpr uint32_t kColPlanCi = 192, kColPlanMeta = 208;   // (!TC) the lane's band-plan constants: 4 x uint4 list offsets, uint4 meta
constexpr uint32_t kTcColA = 128, kTcColD = 384;          // frame slot f: A_hi at kTcColA + 128 f, A_lo 64 further, D at kTcColD + 64 f
template <bool TC>
constexpr uint32_t pair_tmem_cols() { return TC ? 512u : 256u; }
// mbarrier slots (s_bar): [0, 8) one per warp for its bulk copies; TC: a_ready[g], d_ready[g], r_free
constexpr int kBarAReady = 8, kBarDReady = 10, kBarRFree = 12;

// W_64^K1 for the untangling twiddle W_2048^(32 K1 + lane) = W_64^K1 * W_2048^lane (see stft.cu)
__device__ constexpr float kW64p[17][2] = {{1.000000000e+00f, -0.000000000e+00f}, {9.951847267e-01f, -9.801714033e-02f}, {9.807852804e-01f, -1.950903220e-01f}, {9.569403357e-01f, -2.902846773e-01f}, {9.238795325e-01f, -3.826834324e-01f}, {8.819212643e-01f, -4.713967368e-01f}, {8.314696123e-01f, -5.555702330e-01f}, {7.730104534e-01f, -6.343932842e-01f}, {7.071067812e-01f, -7.071067812e-01f}, {6.343932842e-01f, -7.730104534e-01f}, {5.555702330e-01f, -8.314696123e-01f}, {4.713967368e-01f, -8.819212643e-01f}, {3.826834324e-01f, -9.238795325e-01f}, {2.902846773e-01f, -9.569403357e-01f}, {1.950903220e-01f, -9.807852804e-01f}, {9.801714033e-02f, -9.951847267e-01f}, {6.123233996e-17f, -1.000000000e+00f}};
template <int K1>
__device__ __forceinline__ float2 pair_tw2(const float2 base) {
  if constexpr (K1 == 0) return base;
  else if constexpr (K1 == 16) return make_float2(base.y, -base.x);
  else {
    constexpr float c = kW64p[K1][0], s = kW64p[K1][1];
    return make_float2(fmaf(c, base.x, -s * base.y), fmaf(c, base.y, s * base.x));
  }
}

template <int PMODE>
__device__ __forceinline__ pk pair_power(pk re, pk im, float half_power) {
  const pk s = pfma(re, re, im * im);
  if constexpr (PMODE == 2) return s;
  else if constexpr (PMODE == 1) return mk2(sqrtf(lo(s)), sqrtf(hi(s)));
  else {
    const float a = lo(s), b = hi(s);
    const float z = (half_power == 0.0f) ? 1.0f : 0.0f;
    return mk2(a > 0.0f ? exp2f(half_power * __log2f(a)) : z, b > 0.0f ? exp2f(half_power * __log2f(b)) : z);
  }
}

__device__ __forceinline__ pk pair_shfl(pk x, int src) {
  return mk2(__shfl_sync(0xffffffffu, lo(x), src), __shfl_sync(0xffffffffu, hi(x), src));
}

// 4 / 16 raw columns of this thread's tensor-memory lane (bit patterns: the band plan's integer constants)
__device__ __forceinline__ uint4 tmem_ld4_u32(uint32_t taddr) {
  uint4 r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(taddr) : "memory");
  tc_wait_ld();
  return r;
}
__device__ __forceinline__ void tmem_st16_u32(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

// 16 columns of this thread's tensor-memory lane
__device__ __forceinline__ void tmem_table16(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }
// the band weights from shared memory ([chunk][quarter][lane] float4; TC, where tensor memory is taken by the operands)
__device__ __forceinline__ void smem_table16(const float4* tab, int chunk, int lane, float (&v)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = tab[(chunk * 4 + q) * 32 + lane];
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}
#define PAIR_TABLE16(kindcol, c, v) tmem_table16(t_lane + (kindcol) + 16 * (c), v)

// ---- tensor-core pass 2 (TC) ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pair_swz_off(int r, int c16) {          // 128B-swizzled K-major tile, as melbank.cu
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}
__device__ __forceinline__ uint64_t pair_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | (8u << 24);   // f32 += tf32 x tf32, M 128, N 64
// B[(n1, c), (k1, c')] of the real form of Z[k1] = sum_n1 Y[n1] exp(-2 pi i n1 k1 / 32); k = n1 + 32 c, n = k1 + 32 c'
__device__ __forceinline__ float pair_dft_entry(int k, int n) {
  const int n1 = k & 31, c = k >> 5, k1 = n & 31, cp = n >> 5;
  float sn, cs;
  sincospif((float)((n1 * k1) & 31) / 16.0f, &sn, &cs);
  return c == cp ? cs : (c == 1 ? sn : -sn);
}

// Everything about a pair's sample region that is not the plain interior bulk copy, out of line (inlined, the
// unrolled gather and padding loops made the kernel 10 k instructions and ptxas cloned half the frame loop).
// The region is samples [start, start + len) of the PADDED row.  Edge regions get the part inside the row by bulk copy
// and their padding rebuilt in place afterwards (reflection = mirror inside the region, replicate = edge value,
// constant = zeros); regions the bulk copy cannot take at all (circular padding, rows that are not 16-byte aligned,
// regions wider than the row) are gathered sample by sample, 16 loads in flight per lane.
struct RegionSpan {
  int lo, hi;        // region indices [lo, hi) filled by the bulk copy
  bool bulk;
};
__device__ __forceinline__ RegionSpan region_span(int start, int len, int n, int pad_mode, int bulk_ok) {
  RegionSpan f;
  f.lo = start < 0 ? -start : 0;
  f.hi = (start + len > n) ? n - start : len;
  const bool interior = f.lo == 0 && f.hi == len;
  f.bulk = bulk_ok && f.hi > f.lo && (interior || (pad_mode != 3 && (n & 3) == 0 && !(f.lo > 0 && f.hi < len)));
  return f;
}
static __device__ __noinline__ void pair_fixup_region(float* region, const float* __restrict__ row, int start, int len, int n,
                                               int pad_mode, int bulk_ok, int lane) {
  const RegionSpan f = region_span(start, len, n, pad_mode, bulk_ok);
  if (!f.bulk) {
    for (int base = 0; base < len; base += 512) {            // 16 loads per lane in flight
      float tmp[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int j = base + lane + 32 * u, s = start + j;
        const bool inside = (s >= 0) & (s < n);
        const float v = (j < len) ? __ldg(row + padded_index(s, n, pad_mode)) : 0.0f;
        tmp[u] = (pad_mode == 1 && !inside) ? 0.0f : v;
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int j = base + lane + 32 * u;
        if (j < len) region[j] = tmp[u];
      }
    }
  } else if (pad_mode == 0) {                                // reflect: x[-k] = x[k], x[n-1+k] = x[n-1-k]
    for (int j = lane; j < f.lo; j += 32) {
      const int src = 2 * f.lo - j;
      region[j] = (src < f.hi) ? region[src] : __ldg(row + (start + src));
    }
    for (int j = f.hi + lane; j < len; j += 32) {
      const int src = 2 * (f.hi - 1) - j;
      region[j] = (src >= f.lo) ? region[src] : __ldg(row + (start + src));
    }
  } else if (pad_mode == 2) {                                // replicate
    const float a = region[f.lo], b = region[f.hi - 1];
    for (int j = lane; j < f.lo; j += 32) region[j] = a;
    for (int j = f.hi + lane; j < len; j += 32) region[j] = b;
  } else {                                                   // constant
    for (int j = lane; j < f.lo; j += 32) region[j] = 0.0f;
    for (int j = f.hi + lane; j < len; j += 32) region[j] = 0.0f;
  }
  __syncwarp();
}

// Band contraction of both frames' power spectra (stash of float2 pairs) with the two-band plan, dB epilogue, stores.
// Same walk as band_contract (stft.cu); weights (from tensor memory), masks and list offsets are shared by the frames.
// quad: the lane's four sums are the ADJACENT bands 4 lane .. 4 lane + 3 (the plan's `quad` table, frame-major output):
// one 16-byte store per frame instead of four 4-byte stores.
template <bool PEERS, bool TC>
__device__ __forceinline__ void band_contract_pair(const StftParams& p, float2* stash, uint32_t t_lane, const float4* s_tab, int lane, int64_t off_a,
                                                   int64_t off_b, bool store_b, bool quad) {
  __syncwarp();
  // The lane's plan constants: from tensor memory (no global loads in the frame loop, so nothing here depends on L1,
  // which the warp blocks leave little of); the TC variant has its tensor memory full of operands and reads the plan.
  const bool fast = p.band_fast != 0;
  uint4 meta;
  if constexpr (TC) meta = __ldg(reinterpret_cast<const uint4*>(p.band_plan + kBandOffMeta) + lane);
  else meta = tmem_ld4_u32(t_lane + kColPlanMeta);
  pk pw[32];
#pragma unroll
  for (int i = 1; i < 32; ++i) {
    const float2 t = stash[lane * kStashStride + i];
    pw[i] = mk2(t.x, t.y);
  }
  {
    const float2 t = stash[lane * kStashStride - (lane >= kStashShiftRow ? 1 : 0)];
    pw[0] = mk2(t.x, t.y);
  }
  pk p_last = bc(0.0f);
  if (lane == 31) {
    const float2 t = stash[kStashNyquist];
    p_last = mk2(t.x, t.y);
  }
  __syncwarp();
  const uint32_t mask = meta.x;
  float2* row = stash + lane * kStashStride;
  pk u = bc(0.0f), v = bc(0.0f);
  const pk zero = bc(0.0f);
  // the partial sum of a finished band is parked in this lane's own (consumed) stash row as soon as it is complete
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float w[16];                                   // (w0, w1) of bins 8 q .. 8 q + 7 of this lane
    if constexpr (TC) smem_table16(s_tab, q, lane, w);
    else PAIR_TABLE16(kColBand, q, w);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int bin = 8 * q + i;
      const pk un = pfma(pw[bin], bc(w[2 * i]), u), vn = pfma(pw[bin], bc(w[2 * i + 1]), v);
      const bool step = (mask >> bin) & 1u;
      if (step) row[bin] = make_float2(lo(un), hi(un));
      u = psel(step, vn, un);
      v = psel(step, zero, vn);
    }
  }
  u = pfma(p_last, bc(__uint_as_float(meta.z)), u);
  v = pfma(p_last, bc(__uint_as_float(meta.w)), v);
  stash[meta.y] = make_float2(lo(u), hi(u));
  stash[meta.y + 1] = make_float2(lo(v), hi(v));
  __syncwarp();

  // 10 log10(s2) = (10 log10 2) log2(s2) with the hardware logarithm (MUFU.LG2: 2 ulp of the logarithm, < 3e-5 dB at
  // -100 dB against the 1e-3 dB bar); log10f was ~300 of the ~2 600 instructions per pair with the dB epilogue on.
  const float db_off = -10.0f * p.log10_ref;
  auto finish = [&](float r) {
    if (p.to_db) {
      float s2 = r * r;
      s2 = (s2 < p.amin) ? p.amin : s2;
      r = fmaf(__log2f(s2), 3.01029995663981195f, db_off);
    }
    return r;
  };
  auto store = [&](int64_t off, int m, float r) {
    const int64_t o = off + (int64_t)m * p.out_band_stride;
    if constexpr (PEERS) {
      if (p.peer_multicast) {
        multimem_st_f32(p.peer_out[0] + o, r);
      } else {
#pragma unroll 1
        for (int q = 0; q < p.n_peers; ++q) __stcs(p.peer_out[q] + o, r);
      }
    } else {
      __stcs(p.out + o, r);
    }
  };
  if (fast) {
    uint4 ci[4];
    if constexpr (TC) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        ci[j] = __ldg(reinterpret_cast<const uint4*>(p.band_plan + p.band_off_fast + (quad ? kBandFastBytes : 0)) + lane * 4 + j);
    } else {
      uint32_t raw[16];
      tmem_ld16_nowait(t_lane + kColPlanCi, raw);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 4; ++j) ci[j] = make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
    }
    const unsigned char* sb = reinterpret_cast<const unsigned char*>(stash);
    pk acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {                  // the plan's byte offsets address floats: pairs are twice as far
      const float2 a = *reinterpret_cast<const float2*>(sb + 2 * ci[j].x), b = *reinterpret_cast<const float2*>(sb + 2 * ci[j].y);
      const float2 c = *reinterpret_cast<const float2*>(sb + 2 * ci[j].z);
      const float2 d = (p.band_cmax > 3) ? *reinterpret_cast<const float2*>(sb + 2 * ci[j].w) : make_float2(0.0f, 0.0f);
      acc[j] = ((mk2(a.x, a.y) + mk2(b.x, b.y)) + mk2(c.x, c.y)) + mk2(d.x, d.y);
    }
    if (quad) {                                    // n_bands % 4 == 0 and 16-byte aligned rows (checked by the caller)
      if (4 * lane < p.n_bands) {
        const float4 ra = make_float4(finish(lo(acc[0])), finish(lo(acc[1])), finish(lo(acc[2])), finish(lo(acc[3])));
        const float4 rb = make_float4(finish(hi(acc[0])), finish(hi(acc[1])), finish(hi(acc[2])), finish(hi(acc[3])));
        auto store4 = [&](int64_t off, const float4 r) {
          const int64_t o = off + 4 * lane;
          if constexpr (PEERS) {
            if (p.peer_multicast) {
              multimem_st_f32x4(p.peer_out[0] + o, r);
            } else {
#pragma unroll 1
              for (int q = 0; q < p.n_peers; ++q) __stcs(reinterpret_cast<float4*>(p.peer_out[q] + o), r);
            }
          } else {
            __stcs(reinterpret_cast<float4*>(p.out + o), r);
          }
        };
        store4(off_a, ra);
        if (store_b) store4(off_b, rb);
      }
      __syncwarp();
      return;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = lane + 32 * j;
      if (m < p.n_bands) {
        store(off_a, m, finish(lo(acc[j])));
        if (store_b) store(off_b, m, finish(hi(acc[j])));
      }
    }
    __syncwarp();
    return;
  }
  const uint16_t* comb = reinterpret_cast<const uint16_t*>(p.band_plan + kBandOffComb);
  for (int m = lane; m < p.n_bands; m += 32) {
    pk acc = bc(0.0f);
    for (int c = 0; c < p.band_cmax; ++c) {
      const float2 t = stash[__ldg(comb + c * p.n_bands_pad + m)];
      acc = acc + mk2(t.x, t.y);
    }
    store(off_a, m, finish(lo(acc)));
    if (store_b) store(off_b, m, finish(hi(acc)));
  }
  __syncwarp();
}

// SHIFT: hop / 64 when hop is a multiple of 64 (frame B's register r is then frame A's register r + SHIFT: the two
// frames share their sample loads), 0 for any other even hop (separate loads).
// TC: pass 2 on the tensor cores (see the file header); pairs are then dealt statically, round by round, because the
// four warps of a group meet once per pair.
template <bool PEERS, int PMODE, int SHIFT, bool TC>
__global__ void __launch_bounds__(kPairThreads, 1) stft2048_pair_kernel(const StftParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const float4* s_tab = reinterpret_cast<const float4*>(smem_raw);   // band weights (TC only, 0 bytes otherwise)
  float2* s_twb = reinterpret_cast<float2*>(smem_raw + (TC ? kPairBandTabBytes : 0));        // W_2048^lane
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_twb + 32);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 15);
  uint32_t* s_next = s_tmem + 1;                              // next pair of this CTA's chunk nobody has taken yet
  unsigned char* s_blocks = reinterpret_cast<unsigned char*>(s_bar + 16);
  // TC: the B images behind the warp blocks, on a 1 KB boundary (swizzle atoms)
  const uint32_t b_base = (smem_u32(s_blocks + (size_t)kPairWarps * kWarpBytes) + 1023u) & ~1023u;

  // Programmatic dependent launch: the next launch of this kernel in the stream (launched with the stream-serialization
  // attribute, see launch_stft2048_pair_t) may be scheduled while this grid still runs -- its CTAs take an SM as soon as
  // this grid's CTA there has exited and do their set-up (tensor-memory allocation, barriers, twiddles) before
  // `griddepcontrol.wait`, which holds them until this grid has completed and its stores are visible.  Hides the launch
  // latency and the set-up behind the previous launch's tail; a no-op for any other neighbour in the stream.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int group = warp >> 2;                                 // TC: warps 0-3 / 4-7, one per lane quadrant each
  static_assert(!TC || kPairWarps == 8, "the tensor-core pass is laid out for two groups of four warps");
  uint64_t* bar = s_bar + warp;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (tid == 0) {
    *s_next = kPairWarps;                                      // the first kPairWarps pairs are dealt statically
    if constexpr (TC) {
      mbar_init(s_bar + kBarAReady, 4);
      mbar_init(s_bar + kBarAReady + 1, 4);
      mbar_init(s_bar + kBarDReady, 1);
      mbar_init(s_bar + kBarDReady + 1, 1);
      mbar_init(s_bar + kBarRFree, 4);
      fence_mbar_init();
    }
  }
  __syncwarp();

  unsigned char* block = s_blocks + (size_t)warp * kWarpBytes;
  float* region = reinterpret_cast<float*>(block);
  float4* xbuf = reinterpret_cast<float4*>(block);
  float2* stash = reinterpret_cast<float2*>(block + kStashOffBytes);
  if (lane == 0) stash[kStashZero] = make_float2(0.0f, 0.0f);
  __syncwarp();

  const float half_power = 0.5f * p.power;
  // frame-major output with whole float4 rows: the lane's four bands are adjacent (bandplan.cuh `quad`)
  bool quad = p.band_fast != 0 && p.out_band_stride == 1 && (p.n_bands & 3) == 0 && (p.out_t_stride & 3) == 0 && (p.out_seq_stride & 3) == 0;
  if constexpr (PEERS) {
    for (int q = 0; q < p.n_peers; ++q) quad = quad && (reinterpret_cast<uintptr_t>(p.peer_out[q]) & 15) == 0;
  } else {
    quad = quad && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
  }
  const uint32_t frames_u = (uint32_t)p.frames;
  const uint64_t pol_stream = l2_policy_evict_first();
  const int hop = p.hop, len = 2048 + hop, n_samples = (int)p.n_samples;

  // Pairs never straddle sequences: pair (seq, j) = frames 2 j and 2 j + 1 of sequence seq; an odd frame count leaves
  // the last pair of every sequence without a frame B (computed from the same region, result dropped).
  const uint32_t pairs_per_seq = (frames_u + 1) >> 1;
  const uint32_t n_pairs = (uint32_t)p.n_seq * pairs_per_seq;
  const uint32_t per_cta = n_pairs / gridDim.x, extra = n_pairs % gridDim.x;
  const uint32_t chunk0 = blockIdx.x * per_cta + (blockIdx.x < extra ? blockIdx.x : extra);
  const uint32_t chunk1 = chunk0 + per_cta + (blockIdx.x < extra ? 1u : 0u);
  // A warp takes pair chunk0 + warp first and then whatever pair of the chunk is next in line (shared counter), so the
  // warps of a CTA finish within one pair of each other whatever the chunk length (static dealing left the warps
  // with 5 rounds waiting for those with 6 at config 2: 14 % of all warp samples sat at the final barrier).
  // TC: round r gives warp w the pair chunk0 + 8 r + w; every warp walks all rounds (a warp without a pair only keeps
  // the group's barriers counting).
  uint32_t pj = chunk0 + warp;
  uint32_t seq = pj / pairs_per_seq, j = pj % pairs_per_seq;
  const uint32_t n_rounds = (chunk1 - chunk0 + kPairWarps - 1) / kPairWarps;

  // bulk copy of a pair's region: whatever part lies inside the row; the barrier is armed even when nothing can be
  // copied (0 bytes), so the frame loop waits unconditionally
  auto stage_bulk = [&](uint32_t sq, uint32_t jj) {
    const int start = (int)(2 * jj) * hop - p.pad;
    const RegionSpan f = region_span(start, len, n_samples, p.pad_mode, p.bulk_ok);
    if (elect_one()) {
      fence_proxy_async();
      const uint32_t bytes = f.bulk ? (uint32_t)(f.hi - f.lo) * 4u : 0u;
      mbar_arrive_expect_tx(bar, bytes);
      if (f.bulk) bulk_g2s_hint(region + f.lo, p.x + (int64_t)sq * p.seq_stride + (start + f.lo), bytes, bar, pol_stream);
    }
  };
  // ---- set-up that reads no global memory, then the dependency wait, then the first bulk copy ----
  if (warp == 0) {
    tmem_alloc(s_tmem, pair_tmem_cols<TC>());
    tmem_relinquish();
  }
  if (tid < 32) {
    float sn, cs;
    sincospif(-2.0f * (float)tid / 2048.0f, &sn, &cs);
    s_twb[tid] = make_float2(cs, sn);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");           // everything before us in the stream has completed
  if (pj < chunk1) stage_bulk(seq, j);

  // ---- per-lane tables into tensor memory (one warp per lane quadrant writes, all warps of the quadrant read) ----
  if constexpr (TC) {
    // B images: element (k, n) of the K slab k / 32 at row n, 16-byte column (k % 32) / 4 of the swizzled tile
    for (int idx = tid; idx < 64 * 64; idx += kPairThreads) {
      const int k = idx >> 6, n = idx & 63;
      const float v = pair_dft_entry(k, n);
      const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      const uint32_t off = (uint32_t)(k >> 5) * (uint32_t)kTcSlabBytes + pair_swz_off(n, (k & 31) >> 2) + (uint32_t)(k & 3) * 4u;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(b_base + off), "f"(h) : "memory");
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(b_base + 2u * (uint32_t)kTcSlabBytes + off), "f"(v - h) : "memory");
    }
    fence_proxy_async();                                       // generic-proxy writes -> the tensor core's reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  {
    // 12 chunks of 16 columns (4 window, 4 twiddle, 4 band weights) per lane quadrant, dealt to the warps of the
    // quadrant (warp, warp + 4, ...)
    const float g = 0.5f * p.scale;
    const float4* wtab = reinterpret_cast<const float4*>(p.band_plan + kBandOffW) + lane;
    const int third = warp >> 2;
#pragma unroll 1
    for (int chunk = third; chunk < 12; chunk += kPairWarps / 4) {
      const int kind = chunk >> 2, c = chunk & 3;
      float w[16];
      if (kind == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int n = lane + 32 * (8 * c + i);
          w[2 * i] = __ldg(p.window + 2 * n) * g;
          w[2 * i + 1] = __ldg(p.window + 2 * n + 1) * g;
        }
      } else if (kind == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float sn, cs;
          sincospif(-2.0f * (float)((8 * c + i) * lane) / 1024.0f, &sn, &cs);
          w[2 * i] = cs;
          w[2 * i + 1] = sn;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b = __ldg(wtab + (4 * c + i) * 32);    // bins 2 j, 2 j + 1 of this lane, j = 4 c + i
          w[4 * i] = b.x; w[4 * i + 1] = b.y; w[4 * i + 2] = b.z; w[4 * i + 3] = b.w;
        }
      }
      if constexpr (TC) {
        if (kind == 2) {                                      // band weights: shared memory (quadrant-independent)
          if ((warp & 3) == 0) {
            float4* tab = const_cast<float4*>(s_tab);
#pragma unroll
            for (int q = 0; q < 4; ++q) tab[(c * 4 + q) * 32 + lane] = make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
          }
          continue;
        }
      }
      tmem_st16(t_lane + (kind == 0 ? kColWin : (kind == 1 ? kColTw1 : kColBand)) + 16 * c, w);
    }
    if constexpr (!TC) {
      if (third == 0) {                                       // one warp per quadrant: this lane's plan constants
        uint32_t raw[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 c4 = p.band_fast ? __ldg(reinterpret_cast<const uint4*>(p.band_plan + p.band_off_fast + (quad ? kBandFastBytes : 0)) + lane * 4 + j)
                                       : make_uint4(0u, 0u, 0u, 0u);
          raw[4 * j] = c4.x; raw[4 * j + 1] = c4.y; raw[4 * j + 2] = c4.z; raw[4 * j + 3] = c4.w;
        }
        tmem_st16_u32(t_lane + kColPlanCi, raw);
        const uint4 m4 = __ldg(reinterpret_cast<const uint4*>(p.band_plan + kBandOffMeta) + lane);
#pragma unroll
        for (int j = 0; j < 16; ++j) raw[j] = 0u;
        raw[0] = m4.x; raw[1] = m4.y; raw[2] = m4.z; raw[3] = m4.w;
        tmem_st16_u32(t_lane + kColPlanMeta, raw);
      }
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  uint32_t parity = 0, round = 0;
  const int partner = (32 - lane) & 31;
  const float2 tw2_base = s_twb[lane];
#pragma unroll 1
  while (true) {
    bool valid = true;                                 // this warp has a pair in this round (always, without TC)
    if constexpr (TC) {
      if (round >= n_rounds) break;
      valid = pj < chunk1;
    } else {
      if (pj >= chunk1) break;
    }
    const uint32_t t_a = 2 * j;
    const bool has_b = t_a + 1 < frames_u;
    cx<pk> v[32];
    uint32_t pj_next = pj + kPairWarps, seq_next = seq, j_next = j;
    if (valid) {
    mbar_wait(bar, parity);
    parity ^= 1u;
    {
      const int start = (int)t_a * hop - p.pad;
      if (!(p.bulk_ok && start >= 0 && start + len <= n_samples))
        pair_fixup_region(region, p.x + (int64_t)seq * p.seq_stride, start, len, n_samples, p.pad_mode, p.bulk_ok, lane);
    }

    // ---- samples * window: lane = n1, register r <-> z[n1 + 32 r] of both frames -------------------------------
    {
      const float2* sa2 = reinterpret_cast<const float2*>(region);
      if constexpr (SHIFT > 0) {
        // z_B[n] = z_A[n + 32 SHIFT]: element r of frame B is element r + SHIFT of frame A, same lane -- 32 + SHIFT
        // loads serve both frames (40 instead of 64 LDS.64 per pair at hop 512)
        float2 e[32 + SHIFT];
#pragma unroll
        for (int r = 0; r < 32 + SHIFT; ++r) e[r] = sa2[lane + 32 * r];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float w[16];
          PAIR_TABLE16(kColWin, c, w);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 8 * c + i;
            v[r].x = mk2(e[r].x * w[2 * i], e[r + SHIFT].x * w[2 * i]);
            v[r].y = mk2(e[r].y * w[2 * i + 1], e[r + SHIFT].y * w[2 * i + 1]);
          }
        }
      } else {
        const float2* sb2 = reinterpret_cast<const float2*>(region + hop);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float w[16];
          PAIR_TABLE16(kColWin, c, w);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 8 * c + i;
            const float2 a = sa2[lane + 32 * r], b = sb2[lane + 32 * r];
            v[r].x = mk2(a.x * w[2 * i], b.x * w[2 * i]);
            v[r].y = mk2(a.y * w[2 * i + 1], b.y * w[2 * i + 1]);
          }
        }
      }
    }
    __syncwarp();                                    // samples consumed: the block becomes the transposition buffer
    dit_fft_fma_r<32, pk>(v);                        // pass 1 over r, both frames
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
      const cx<pk> e = v[bit_reverse<32>(k2)];
      xbuf[k2 * kXStride + lane] = make_float4(lo(e.x), hi(e.x), lo(e.y), hi(e.y));
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float w[16];                                   // W_1024^(n1 k2), k2 = lane, n1 = 8 c .. 8 c + 7
      PAIR_TABLE16(kColTw1, c, w);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n1 = 8 * c + i;
        const float4 a = xbuf[lane * kXStride + n1];
        const pk are = mk2(a.x, a.y), aim = mk2(a.z, a.w);
        v[n1].x = pfma(are, bc(w[2 * i]), aim * bc(-w[2 * i + 1]));
        v[n1].y = pfma(are, bc(w[2 * i + 1]), aim * bc(w[2 * i]));
      }
    }
    __syncwarp();                                    // block free again: fetch the next pair while this one finishes

    if constexpr (!TC) {
      pj_next = 0;
      if (lane == 0) pj_next = chunk0 + atomicAdd(s_next, 1u);
      pj_next = __shfl_sync(0xffffffffu, pj_next, 0);
    }
    seq_next = pj_next / pairs_per_seq;
    j_next = pj_next % pairs_per_seq;
    if (pj_next < chunk1) stage_bulk(seq_next, j_next);
    }   // valid

#ifdef PAIR_STOP_AFTER_TW1                           // cost-breakdown builds (scripts/build_variant.sh): stop here, keep the values live
    { pk acc = v[0].x; for (int i = 0; i < 32; ++i) acc = acc + v[i].x + v[i].y; if (lo(acc) + hi(acc) == 12345.678f) p.out[lane] = lo(acc); }
    pj = pj_next; seq = seq_next; j = j_next; continue;
#endif
    if constexpr (!TC) {
      dit_fft_fma_r<32, pk>(v);                      // pass 2 over n1: v[bit_reverse(k1)] = Z[32 k1 + lane] / 2
    } else {
      // ---- pass 2 on the tensor cores: rows = this warp's 32 lanes (k2), K = (n1, re | im), N = (k1, re | im) ---------
      uint64_t* a_ready = s_bar + kBarAReady + group;
      uint64_t* d_ready = s_bar + kBarDReady + group;
      uint64_t* r_free = s_bar + kBarRFree;
      const uint32_t use = 2 * round + (uint32_t)group;          // the groups take turns in the operand columns
      if (use > 0) mbar_wait(r_free, (use - 1) & 1u);
      tc_fence_after();
      if (valid) {
#pragma unroll
        for (int f = 0; f < 2; ++f) {
#pragma unroll
          for (int part = 0; part < 4; ++part) {               // k = 16 part + i: re of n1 = 0..31, then im
            float h[16], l[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n1 = (part & 1) * 16 + i;
              const pk val = part < 2 ? v[n1].x : v[n1].y;
              const float x = f ? hi(val) : lo(val);
              h[i] = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
              l[i] = x - h[i];
            }
            tmem_st16(t_lane + kTcColA + 128 * f + 16 * part, h);
            tmem_st16(t_lane + kTcColA + 128 * f + 64 + 16 * part, l);
          }
        }
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
      if ((warp & 3) == 0) {                                   // the group's first warp issues, through one elected lane
        mbar_wait(a_ready, round & 1u);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int f = 0; f < 2; ++f) {
            const uint32_t a_hi = tmem_base + kTcColA + 128 * f, a_lo = a_hi + 64, d = tmem_base + kTcColD + 64 * f;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {                   // UMMA K = 8: 8 columns of A, 32 bytes of each B row
              const uint32_t slab = b_base + (uint32_t)(ks >> 2) * (uint32_t)kTcSlabBytes;
              const uint64_t b_hi = pair_desc_sw128(slab) + 2 * (ks & 3);
              const uint64_t b_lo = pair_desc_sw128(slab + 2u * (uint32_t)kTcSlabBytes) + 2 * (ks & 3);
              tc_mma_tf32_ts(d, a_lo + 8 * ks, b_hi, kTcIdesc, ks > 0 ? 1u : 0u);
              tc_mma_tf32_ts(d, a_hi + 8 * ks, b_lo, kTcIdesc, 1u);
              tc_mma_tf32_ts(d, a_hi + 8 * ks, b_hi, kTcIdesc, 1u);
            }
          }
          tc_commit(d_ready);
        }
        __syncwarp();
      }
      mbar_wait(d_ready, round & 1u);
      tc_fence_after();
      if (valid) {
        uint32_t za[64], zb[64];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld16_nowait(t_lane + kTcColD + 16 * c, *reinterpret_cast<uint32_t(*)[16]>(&za[16 * c]));
          tmem_ld16_nowait(t_lane + kTcColD + 64 + 16 * c, *reinterpret_cast<uint32_t(*)[16]>(&zb[16 * c]));
        }
        tc_wait_ld();
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) {                      // the register labelling the untangling expects
          v[bit_reverse<32>(k1)].x = mk2(__uint_as_float(za[k1]), __uint_as_float(zb[k1]));
          v[bit_reverse<32>(k1)].y = mk2(__uint_as_float(za[32 + k1]), __uint_as_float(zb[32 + k1]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(r_free);
    }
#ifdef PAIR_STOP_AFTER_PASS2
    { pk acc = v[0].x; for (int i = 0; i < 32; ++i) acc = acc + v[i].x + v[i].y; if (lo(acc) + hi(acc) == 12345.678f) p.out[lane] = lo(acc); }
    pj = pj_next; seq = seq_next; j = j_next; continue;
#endif

    if (valid) {
    // ---- untangling + |X|^p of bins 32 k1 + lane and their mirrors 1024 - k, as in stft.cu (emit2) ------------
    {
      float2* dst = stash + lane;
      float2* dmir = stash + (lane == 0 ? kStashStride - 1 : 32 - lane);
      auto emit2 = [&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        const float2 w = pair_tw2<k1>(tw2_base);
        const cx<pk> z = v[bit_reverse<32>(k1)];
        cx<pk> q;
        q.x = pair_shfl(v[bit_reverse<32>(31 - k1)].x, partner);
        q.y = pair_shfl(v[bit_reverse<32>(31 - k1)].y, partner);
        if (lane == 0) q = v[bit_reverse<32>((32 - k1) & 31)];
        const pk a = z.x + q.x, b = z.y - q.y, gs = z.y + q.y, h = q.x - z.x;
        const pk tr = pfma(bc(w.x), gs, bc(-w.y) * h), ti = pfma(bc(w.x), h, bc(w.y) * gs);
        const pk pa = pair_power<PMODE>(a + tr, b + ti, half_power);
        const pk pm = pair_power<PMODE>(a - tr, b - ti, half_power);
        dst[k1 * kStashStride] = make_float2(lo(pa), hi(pa));
        dmir[(31 - k1) * kStashStride] = make_float2(lo(pm), hi(pm));
      };
      static_for<16>(emit2);
      {                                              // bin 512 + lane (lane 0 only: its own mirror)
        const float2 w = pair_tw2<16>(tw2_base);
        const cx<pk> z = v[bit_reverse<32>(16)];
        cx<pk> q;
        q.x = pair_shfl(v[bit_reverse<32>(15)].x, partner);
        q.y = pair_shfl(v[bit_reverse<32>(15)].y, partner);
        if (lane == 0) q = v[bit_reverse<32>(16)];
        const pk a = z.x + q.x, b = z.y - q.y, gs = z.y + q.y, h = q.x - z.x;
        const pk xr = pfma(bc(w.x), gs, pfma(bc(-w.y), h, a)), xi = pfma(bc(w.x), h, pfma(bc(w.y), gs, b));
        const pk pc = pair_power<PMODE>(xr, xi, half_power);
        if (lane == 0) dst[16 * kStashStride] = make_float2(lo(pc), hi(pc));
      }
    }
#ifdef PAIR_STOP_AFTER_UNTANGLE
    __syncwarp();
    { const float2 t = stash[lane * kStashStride + 3]; if (t.x + t.y == 12345.678f) p.out[lane] = t.x; }
    __syncwarp();
    pj = pj_next; seq = seq_next; j = j_next; continue;
#endif
    {
      const int64_t seq0 = PEERS ? p.peer_seq0 : 0;
      const int64_t off_a = ((int64_t)seq + seq0) * p.out_seq_stride + (int64_t)t_a * p.out_t_stride;
      band_contract_pair<PEERS, TC>(p, stash, t_lane, s_tab, lane, off_a, off_a + p.out_t_stride, has_b, quad);
    }
    }   // valid
    pj = pj_next;
    seq = seq_next;
    j = j_next;
    ++round;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, pair_tmem_cols<TC>());
}

bool stft2048_pair_applies(const StftParams& p);
static inline bool pair_pdl_enabled() {                       // TAC_PAIR_PDL=0 switches the programmatic dependent launch off (A/B timing)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TAC_PAIR_PDL");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}

// host side of one instantiation family (stft_pair.cu: TC = false, stft_pair_tc.cu: TC = true)
template <bool TC>
int launch_stft2048_pair_t(const StftParams& p, cudaStream_t stream) {
  const int64_t n_frames = p.g1 - p.g0;
  if (n_frames <= 0) return TAC_OK;
  const int64_t n_pairs = p.n_seq * ((p.frames + 1) / 2);
  TAC_REQUIRE(n_pairs < ((int64_t)1 << 31), TAC_ERR_UNSUPPORTED, "melspec: too many frames in one call");
  const int64_t want = (n_pairs + kPairWarps - 1) / kPairWarps;
  const int grid = (int)(want < sm_count() ? want : sm_count());
  using Kernel = void (*)(const StftParams);
  Kernel k;
  const bool peers = p.out_mode == OUT_MEL_FUSED_PEERS;
  if (p.hop == 512) {                                          // the headline hop: shared sample loads
    if (peers) k = p.power_mode == 2 ? stft2048_pair_kernel<true, 2, 8, TC> : (p.power_mode == 1 ? stft2048_pair_kernel<true, 1, 8, TC> : stft2048_pair_kernel<true, 0, 8, TC>);
    else k = p.power_mode == 2 ? stft2048_pair_kernel<false, 2, 8, TC> : (p.power_mode == 1 ? stft2048_pair_kernel<false, 1, 8, TC> : stft2048_pair_kernel<false, 0, 8, TC>);
  } else {
    if (peers) k = p.power_mode == 2 ? stft2048_pair_kernel<true, 2, 0, TC> : (p.power_mode == 1 ? stft2048_pair_kernel<true, 1, 0, TC> : stft2048_pair_kernel<true, 0, 0, TC>);
    else k = p.power_mode == 2 ? stft2048_pair_kernel<false, 2, 0, TC> : (p.power_mode == 1 ? stft2048_pair_kernel<false, 1, 0, TC> : stft2048_pair_kernel<false, 0, 0, TC>);
  }
  TAC_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair_smem_bytes<TC>()));
  LaunchProbe probe(KIND_STFT, stream);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(kPairThreads, 1, 1);
  cfg.dynamicSmemBytes = pair_smem_bytes<TC>();
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // see the kernel's first lines
  attr[0].val.programmaticStreamSerializationAllowed = pair_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TAC_CUDA_OK(cudaLaunchKernelEx(&cfg, k, p));
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

}  // namespace tac
