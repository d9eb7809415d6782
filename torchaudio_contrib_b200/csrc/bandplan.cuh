// Band plan: the form of a filterbank matrix consumed by the fused STFT + filterbank kernel (stft.cu,
// OUT_MEL_FUSED).  Applies to matrices whose every row (frequency bin) has at most two non-zeros, in
// adjacent columns -- every triangular filterbank (create_mel_filter, functional.py:131-169) is of that
// form: bin k feeds band b_k with weight w0[k] and band b_k + 1 with weight w1[k].
//
// The kernel has one warp per frame; after the FFT the warp transposes the 1025 power values through a
// shared-memory stash so that lane l owns the 32 CONSECUTIVE bins 32 l .. 32 l + 31 (lane 31 also bin 1024).
// The lane walks its bins keeping two running sums (u: band b, v: band b + 1).  Where b steps to b + 1 (bit i
// of the lane's segment mask: after local bin i) it stores u into its own, already consumed, stash row at
// column i and carries on with u = v, v = 0; at the end of its run it stores u and v at `end_pos`, `end_pos + 1`.
// A second step with lane = band sums the few stored floats that belong to a band (list `comb`, fixed order,
// so the result is deterministic).  The plan exists only when consecutive segments inside a lane's run step
// from band b to band b + 1 (or the value that would be orphaned is identically zero) -- true for chains of
// overlapping triangles.
//
// Blob layout (all offsets from the start of the band plan, which is 128-byte aligned inside the
// filterbank plan blob):
//   BandPlanHeader                                  32 B
//   w     [16][32] float4   (w0[i], w1[i], w0[i+1], w1[i+1]) of bins 32 l + i, i = 2 j     8192 B
//   meta  [32]     uint4    (segment mask, end_pos, w0 / w1 of bin 1024 as float bits [lane 31 only])   512 B
//   comb  [cmax][n_bands_pad] uint16   float index into the stash, padded with zero_idx
//   fast  [32][4][4] uint32  (header.fast_off != 0: <= 128 bands, cmax <= 4) byte offsets into the stash of the
//                            entries of bands l, l + 32, l + 64, l + 96 for lane l
//   quad  [32][4][4] uint32  (follows `fast`) the same for bands 4 l .. 4 l + 3: the frame-major layout stores them as one float4
#pragma once

#include <stdint.h>

namespace tac {

constexpr uint32_t kBandPlanMagic = 0x7ac0ba2du;
constexpr int kBandBins = 1025;                    // n_fft = 2048
constexpr int kStashStride = 33;                   // stash row stride (floats): odd -> conflict-free both ways
// Column 0 of rows 17 .. 32 (bins 544, 576 .. 1024) sits ONE FLOAT BELOW its natural place, in the padding slot of
// the previous row: those 16 values are all written by lane 0 as the mirrors of its bins 32 k1, in the same store
// instruction in which lane L >= 1 writes (row 31 - k1, column 32 - L); at the natural address lane 0 and lane 31 hit
// the same bank every step (16 extra shared-memory wavefronts per frame).
constexpr int kStashShiftRow = 17;
constexpr int kStashNyquist = 32 * kStashStride - 1;   // where bin 1024 sits (row 32, column 0, shifted)
constexpr int kStashEnd31 = 32 * kStashStride + 1;     // end-of-run pair of lane 31 (its row's last columns are taken)
constexpr int kStashZero = 1060;                   // a float that is always 0 (padding entries of `comb`)
constexpr int kStashFloats = 1064;                 // per warp
constexpr int kBandMaxComb = 16;
constexpr int kBandFastBytes = 32 * 16 * 4;        // size of the `fast` table; `quad` starts this far behind it

struct BandPlanHeader {
  uint32_t magic;
  int32_t n_bins, n_bands, n_stored, cmax, n_bands_pad, zero_idx, reserved;   // reserved = fast_off
};
constexpr int kBandOffW = 32;
constexpr int kBandOffMeta = kBandOffW + 16 * 32 * 16;
constexpr int kBandOffComb = kBandOffMeta + 32 * 16;

static inline int64_t band_plan_capacity(int n_bands) {
  const int64_t pad = ((int64_t)n_bands + 31) / 32 * 32;
  return kBandOffComb + (int64_t)kBandMaxComb * pad * 2 + 2 * 32 * 16 * 4 + 128;
}

// ---------------------------------------------------------------------------------------------------------------
// Range plan ("CSR by band"): the form consumed by the fused STFT + filterbank epilogue of the warp kernels for
// n_fft = 256 / 512 / 1024 / 4096 (stft_multi.cu, stft4096.cu: OUT_MEL_RANGE).  Band m is stored as the contiguous bin
// range [lo_m, lo_m + len_m) that covers its non-zeros, with the weights of that range; a lane owns bands m = lane + 32 j
// and sums w * |X|^p over each range from the frame's power spectrum in shared memory.  Any matrix qualifies whose
// ranges stay short (sum of lengths <= 4 n_bins): every triangular filterbank, not dense ones.
//   RangePlanHeader                              32 B
//   meta [n_bands_pad] int2   (lo | len << 16, offset of the band's first weight in `w`)
//   w    [nnz_pad]     float
constexpr uint32_t kRangePlanMagic = 0x7ac0c5a1u;
struct RangePlanHeader {
  uint32_t magic;
  int32_t n_bins, n_bands, n_bands_pad, nnz_pad, max_len, reserved[2];
};
static inline int64_t range_plan_capacity(int n_bins, int n_bands) {
  return 32 + ((int64_t)(n_bands + 31) / 32 * 32) * 8 + ((int64_t)4 * n_bins + 4) * 4 + 128;
}
// bytes written at `dst` (multiple of 16), or 0 when the ranges are too long (dense matrix) or too large for the kernels
int64_t build_range_plan(const float* fb, int n_bins, int n_bands, unsigned char* dst, int64_t capacity);

// Returns the bytes written at `dst` (multiple of 16), or 0 when the matrix is not of the two-adjacent-bands
// form (or needs more slots / list entries than the kernel holds).  Host code (bandplan.cu).
int64_t build_band_plan(const float* fb, int n_bins, int n_bands, unsigned char* dst, int64_t capacity);

}  // namespace tac
