// Harmonic-percussive source separation by median filtering (reference: torchaudio_contrib/beta_hpss.py:37-129, a beta
// module the reference does not export from its package; built for completeness of the drop-in surface).
//
//   perc[f, t] = median over the k frequency bins around f   (reflect padding, beta_hpss.py:88-90, :109-112)
//   harm[f, t] = median over the k frames around t           (:84-86, :113-114)
//   both raised to `power` (:94-95); soft masks (harm + eps) / (harm + perc + eps), (perc + eps) / (...) with eps = 1e-6
//   (:120-121) or hard masks harm > perc, harm < perc (:116-118); outputs mag * mask (:126).
//
// One thread per spectrogram element: the two windows (k <= 63 values each) are read into registers and the median is
// found by forgetful selection -- keep m + 2 candidates of the 2 m + 1 values, drop their minimum and maximum (neither
// can be the median), take in the next value, until three are left -- with every index resolved at compile time, so the
// candidates never leave the register file (~1.5 k compare-exchanges for k = 31 instead of a sort).  The window is padded
// to the template size with equally many -inf and +inf, which leaves the median where it was.  torch.median returns NaN
// when the window holds one; so does this.  A median is a selection: the result is bit-exact with the reference.
#include <math.h>

#include "tac_common.cuh"

namespace tac {

constexpr int kHpssThreads = 256;

__device__ __forceinline__ void cswap(float& a, float& b) {     // a <- min, b <- max
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  a = lo;
  b = hi;
}

// median of v[0..N), N odd, all finite or +-inf (NaN handled by the caller)
template <int N>
__device__ __forceinline__ float median_forgetful(float (&v)[N]) {
  static_assert(N % 2 == 1 && N >= 3, "odd window");
  constexpr int M = N / 2 + 2;                                   // candidates kept
  float a[M];
#pragma unroll
  for (int i = 0; i < M; ++i) a[i] = v[i];
#pragma unroll
  for (int s = M; s >= 3; --s) {                                 // s candidates in a[0..s)
#pragma unroll
    for (int i = 0; i < s / 2; ++i) cswap(a[i], a[s - 1 - i]);   // pairs ordered: minimum in the lower half, maximum in the upper
#pragma unroll
    for (int i = 1; i < (s + 1) / 2; ++i) cswap(a[0], a[i]);     // minimum -> a[0]
#pragma unroll
    for (int i = s / 2; i < s - 1; ++i) cswap(a[i], a[s - 1]);   // maximum -> a[s - 1]
    // drop both; the next unseen value (if any) takes the minimum's slot, the maximum's slot falls off the end
    if (M + (M - s) < N) a[0] = v[M + (M - s)];                  // true for every s > 3: exactly N - M values are taken in
  }
  return a[1];
}

__device__ __forceinline__ int reflect_index(int i, int n) {    // torch 'reflect' padding, |pad| < n
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

template <int KMAX>
__global__ void __launch_bounds__(kHpssThreads)
hpss_kernel(const float* __restrict__ mag, int64_t n_seq, int n_freq, int n_time, int k, float power, int power_mode, int hard, int mask_only,
            float* __restrict__ out_harm, float* __restrict__ out_perc, float* __restrict__ mask_harm, float* __restrict__ mask_perc) {
  const int64_t plane = (int64_t)n_freq * n_time, total = n_seq * plane;
  const int half = k / 2, fill = (KMAX - k) / 2;
  for (int64_t e = (int64_t)blockIdx.x * kHpssThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kHpssThreads) {
    const int64_t seq = e / plane;
    const int r = (int)(e - seq * plane), f = r / n_time, t = r - f * n_time;
    const float* base = mag + seq * plane;
    float wt[KMAX], wf[KMAX];
    bool nan_t = false, nan_f = false;
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      const int j = i - fill;                                   // position inside the real window
      float vt = (i < fill) ? -INFINITY : INFINITY, vf = vt;
      if (j >= 0 && j < k) {
        vt = __ldg(base + (int64_t)f * n_time + reflect_index(t + j - half, n_time));
        vf = __ldg(base + (int64_t)reflect_index(f + j - half, n_freq) * n_time + t);
        nan_t |= vt != vt;
        nan_f |= vf != vf;
        vt = (vt != vt) ? INFINITY : vt;
        vf = (vf != vf) ? INFINITY : vf;
      }
      wt[i] = vt;
      wf[i] = vf;
    }
    float harm = median_forgetful<KMAX>(wt), perc = median_forgetful<KMAX>(wf);
    if (nan_t) harm = NAN;
    if (nan_f) perc = NAN;
    if (power_mode == 2) {                                      // torch.pow(x, 2.) == x * x
      harm = harm * harm;
      perc = perc * perc;
    } else if (power_mode == 3) {                               // 0.5 -> sqrt
      harm = sqrtf(harm);
      perc = sqrtf(perc);
    } else if (power_mode == 0) {
      harm = powf(harm, power);
      perc = powf(perc, power);
    }
    float mh, mp;
    if (hard) {
      mh = harm > perc ? 1.0f : 0.0f;
      mp = harm < perc ? 1.0f : 0.0f;
    } else {
      const float den = harm + perc + 1e-6f;
      mh = (harm + 1e-6f) / den;
      mp = (perc + 1e-6f) / den;
    }
    mask_harm[e] = mh;
    mask_perc[e] = mp;
    if (!mask_only) {
      const float m = __ldg(base + r);
      out_harm[e] = m * mh;
      out_perc[e] = m * mp;
    }
  }
}

}  // namespace tac

extern "C" int tac_hpss_f32(const float* mag, int64_t n_seq, int n_freq, int n_time, int kernel_size, float power, int hard, int mask_only,
                            float* out_harm, float* out_perc, float* mask_harm, float* mask_perc, void* stream) {
  using namespace tac;
  TAC_REQUIRE(n_seq >= 0 && n_freq >= 1 && n_time >= 1, TAC_ERR_INVALID, "hpss: bad shape");
  TAC_REQUIRE(kernel_size >= 3 && (kernel_size & 1) == 1, TAC_ERR_INVALID, "hpss: kernel_size=%d must be odd and at least 3", kernel_size);
  TAC_REQUIRE(kernel_size <= 63, TAC_ERR_UNSUPPORTED, "hpss: kernel_size=%d, at most 63 is implemented", kernel_size);
  TAC_REQUIRE(kernel_size / 2 < n_freq && kernel_size / 2 < n_time, TAC_ERR_INVALID,
              "hpss: Padding size should be less than the corresponding input dimension (pad %d, freq %d, time %d)", kernel_size / 2, n_freq,
              n_time);
  const int64_t total = n_seq * (int64_t)n_freq * n_time;
  if (total == 0) return TAC_OK;
  TAC_REQUIRE(mag && mask_harm && mask_perc && (mask_only || (out_harm && out_perc)), TAC_ERR_INVALID, "hpss: null pointer");
  const int power_mode = power == 1.0f ? 1 : (power == 2.0f ? 2 : (power == 0.5f ? 3 : 0));
  const int64_t want = (total + kHpssThreads - 1) / kHpssThreads;
  const int64_t cap = (int64_t)sm_count() * 16;
  const int grid = (int)(want < cap ? want : cap);
  cudaStream_t st = as_stream(stream);
  LaunchProbe probe(KIND_POINTWISE, st);
  if (kernel_size <= 7) hpss_kernel<7><<<grid, kHpssThreads, 0, st>>>(mag, n_seq, n_freq, n_time, kernel_size, power, power_mode, hard, mask_only, out_harm, out_perc, mask_harm, mask_perc);
  else if (kernel_size <= 15) hpss_kernel<15><<<grid, kHpssThreads, 0, st>>>(mag, n_seq, n_freq, n_time, kernel_size, power, power_mode, hard, mask_only, out_harm, out_perc, mask_harm, mask_perc);
  else if (kernel_size <= 31) hpss_kernel<31><<<grid, kHpssThreads, 0, st>>>(mag, n_seq, n_freq, n_time, kernel_size, power, power_mode, hard, mask_only, out_harm, out_perc, mask_harm, mask_perc);
  else hpss_kernel<63><<<grid, kHpssThreads, 0, st>>>(mag, n_seq, n_freq, n_time, kernel_size, power, power_mode, hard, mask_only, out_harm, out_perc, mask_harm, mask_perc);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}
