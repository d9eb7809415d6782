// In-register radix-2 decimation-in-frequency FFTs of 2..64 complex points.
//
// Everything is resolved at compile time: the recursion is over template parameters, every
// array index is a constant and every twiddle is a literal, so the `float2 v[N]` array lives in
// registers and the trivial twiddles (1, -i, (1-i)/sqrt2, (-1-i)/sqrt2) cost no multiplies.
// Output is left in bit-reversed order: after `dif_fft<N>(v)`, X[k] sits in v[bit_reverse<N>(k)].
#pragma once

#include <cuda_runtime.h>

namespace tac {

// cos(2*pi*q/64), q = 0..16 (first quadrant); rounded once from the double value
__host__ __device__ constexpr float quarter_cos64(int q) {
  constexpr float t[17] = {1.0f,
                           0.99518472667219693f, 0.98078528040323043f, 0.95694033573220882f, 0.92387953251128674f,
                           0.88192126434835505f, 0.83146961230254524f, 0.77301045336273699f, 0.70710678118654757f,
                           0.63439328416364549f, 0.55557023301960229f, 0.47139673682599781f, 0.38268343236508984f,
                           0.29028467725446233f, 0.19509032201612833f, 0.09801714032956077f, 0.0f};
  return t[q];
}
// cos / sin of 2*pi*q/64 for 0 <= q < 32
__host__ __device__ constexpr float cos64(int q) { return q <= 16 ? quarter_cos64(q) : -quarter_cos64(32 - q); }
__host__ __device__ constexpr float sin64(int q) { return q <= 16 ? quarter_cos64(16 - q) : quarter_cos64(q - 16); }

template <int N>
__host__ __device__ constexpr int bit_reverse(int k) {
  int r = 0;
  for (int b = 1; b < N; b <<= 1) {
    r = (r << 1) | (k & 1);
    k >>= 1;
  }
  return r;
}

// d * W_N^J with W_N = exp(-2*pi*i/N), 0 <= J < N/2, N <= 64
template <int J, int N>
__device__ __forceinline__ float2 mul_twiddle(float2 d) {
  static_assert(N <= 64 && 64 % N == 0 && J >= 0 && 2 * J < N, "twiddle out of range");
  if constexpr (J == 0) {
    return d;
  } else if constexpr (4 * J == N) {             // -i
    return make_float2(d.y, -d.x);
  } else if constexpr (8 * J == N) {             // (1 - i) / sqrt(2)
    constexpr float h = 0.70710678118654757f;
    return make_float2((d.x + d.y) * h, (d.y - d.x) * h);
  } else if constexpr (8 * J == 3 * N) {         // (-1 - i) / sqrt(2)
    constexpr float h = 0.70710678118654757f;
    return make_float2((d.y - d.x) * h, -(d.x + d.y) * h);
  } else {
    constexpr float c = cos64(J * (64 / N));
    constexpr float s = sin64(J * (64 / N));     // W = c - i s
    return make_float2(fmaf(d.x, c, d.y * s), fmaf(d.y, c, -d.x * s));
  }
}

template <int N, int OFF, int TOTAL>
struct DifStage {
  template <int J>
  static __device__ __forceinline__ void butterflies(float2 (&v)[TOTAL]) {
    const float2 a = v[OFF + J];
    const float2 b = v[OFF + J + N / 2];
    v[OFF + J] = make_float2(a.x + b.x, a.y + b.y);
    v[OFF + J + N / 2] = mul_twiddle<J, N>(make_float2(a.x - b.x, a.y - b.y));
    if constexpr (J + 1 < N / 2) butterflies<J + 1>(v);
  }
  static __device__ __forceinline__ void run(float2 (&v)[TOTAL]) {
    butterflies<0>(v);
    if constexpr (N > 2) {
      DifStage<N / 2, OFF, TOTAL>::run(v);
      DifStage<N / 2, OFF + N / 2, TOTAL>::run(v);
    }
  }
};

// forward FFT of v[0..N), result in bit-reversed positions
template <int N>
__device__ __forceinline__ void dif_fft(float2 (&v)[N]) {
  static_assert(N >= 2 && N <= 64 && (N & (N - 1)) == 0, "N must be a power of two in [2, 64]");
  DifStage<N, 0, N>::run(v);
}

}  // namespace tac
