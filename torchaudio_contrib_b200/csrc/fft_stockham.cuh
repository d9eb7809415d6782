// Stockham autosort FFT over C complex points held in shared memory, executed by all THREADS threads of a CTA.
// Forward transform (e^{-2 pi i j k / C}); radix 4 while four sub-transforms remain, one radix-2 pass if log2(C)
// is odd.  `tw_c[m]` = W_C^m for m < C/2 (the other half follows from W^(m + C/2) = -W^m).
// Returns the buffer holding the result (src or dst); every pass ends with __syncthreads().
// Used by stft_generic_kernel (stft.cu) and the backward kernel (stft_backward.cu).
#pragma once

#include <cuda_runtime.h>

namespace tac {

template <int THREADS>
__device__ __forceinline__ float2* stockham_forward(float2* src, float2* dst, const float2* __restrict__ tw_c, int C, int tid) {
  auto tw = [&](int m) -> float2 {
    const float2 w = tw_c[m & ((C >> 1) - 1)];
    return (m & (C >> 1)) ? make_float2(-w.x, -w.y) : w;
  };
  auto cmul = [](float2 a, float2 w) { return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x)); };
  int ns = 1;
  for (; ns * 4 <= C; ns <<= 2) {
    const int tw_stride = C / (4 * ns);
    for (int j = tid; j < (C >> 2); j += THREADS) {
      const int k = j & (ns - 1);
      const int m = k * tw_stride;                       // angle -2 pi k / (4 ns)
      const float2 v0 = src[j];
      const float2 v1 = cmul(src[j + (C >> 2)], tw(m));
      const float2 v2 = cmul(src[j + (C >> 1)], tw(2 * m));
      const float2 v3 = cmul(src[j + 3 * (C >> 2)], tw(3 * m));
      const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
      const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
      const int j0 = ((j - k) << 2) + k;
      dst[j0] = make_float2(s02.x + s13.x, s02.y + s13.y);
      dst[j0 + ns] = make_float2(d02.x + d13.y, d02.y - d13.x);          // d02 - i d13
      dst[j0 + 2 * ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
      dst[j0 + 3 * ns] = make_float2(d02.x - d13.y, d02.y + d13.x);      // d02 + i d13
    }
    __syncthreads();
    float2* tmp = src;
    src = dst;
    dst = tmp;
  }
  for (; ns < C; ns <<= 1) {
    const int tw_stride = C / (2 * ns);
    for (int j = tid; j < (C >> 1); j += THREADS) {
      const int k = j & (ns - 1);
      const float2 a = src[j];
      const float2 b = cmul(src[j + (C >> 1)], tw_c[k * tw_stride]);
      const int j0 = ((j - k) << 1) + k;
      dst[j0] = make_float2(a.x + b.x, a.y + b.y);
      dst[j0 + ns] = make_float2(a.x - b.x, a.y - b.y);
    }
    __syncthreads();
    float2* tmp = src;
    src = dst;
    dst = tmp;
  }
  return src;
}

}  // namespace tac
