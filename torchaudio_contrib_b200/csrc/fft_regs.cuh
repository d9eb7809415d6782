// In-register radix-2 decimation-in-frequency FFTs of 2..64 complex points.
//
// Everything is resolved at compile time: the recursion is over template parameters, every
// array index is a constant and every twiddle is a literal, so the `float2 v[N]` array lives in
// registers and the trivial twiddles (1, -i, (1-i)/sqrt2, (-1-i)/sqrt2) cost no multiplies.
// Output is left in bit-reversed order: after `dif_fft<N>(v)`, X[k] sits in v[bit_reverse<N>(k)].
#pragma once

#include <cuda_runtime.h>

#include "f32x2.cuh"

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

// compile-time loop: f(std::integral_constant<int, 0>{}) ... f(std::integral_constant<int, N-1>{})
template <int I>
struct IntC {
  static constexpr int value = I;
};
template <int N, int I = 0, class Fn>
__device__ __forceinline__ void static_for(Fn&& f) {
  if constexpr (I < N) {
    f(IntC<I>{});
    static_for<N, I + 1>(f);
  }
}

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

// ---------------------------------------------------------------------------------------------
// FMA-form decimation-in-time variant (Linzer-Feig): same contract as dif_fft -- natural-order input in
// v[0..N), X[k] left in v[bit_reverse<N>(k)] -- but every butterfly with a non-trivial twiddle costs six
// FMAs instead of four add/sub plus a four-instruction complex multiply:
//      w b = c (b_r + t b_i) + i c (b_i - t b_r),  t = s / c        (|c| >= |s|, else the cotangent form)
//      X = u + c p,  Y = u - c p
// 32 points: 388 fp32 instructions instead of 456.  The DIT network runs on the virtually bit-reversed
// array A[i] = v[bit_reverse(i)] (a compile-time relabelling), which is what leaves X[k] in v[bit_reverse(k)].
// ---------------------------------------------------------------------------------------------
template <int J, int M>
__device__ __forceinline__ void dit_butterfly(float2& u, float2& b) {
  static_assert(M <= 64 && 64 % M == 0 && J >= 0 && 2 * J < M, "twiddle out of range");
  float pr, pi, g;                                 // w b = g * (pr + i pi)
  if constexpr (J == 0) {
    const float ur = u.x, ui = u.y;
    u = make_float2(ur + b.x, ui + b.y);
    b = make_float2(ur - b.x, ui - b.y);
    return;
  } else if constexpr (4 * J == M) {               // w = -i
    const float ur = u.x, ui = u.y, br = b.x, bi = b.y;
    u = make_float2(ur + bi, ui - br);
    b = make_float2(ur - bi, ui + br);
    return;
  } else {
    constexpr double c = (double)cos64(J * (64 / M)), sn = (double)sin64(J * (64 / M));
    if constexpr ((c >= 0 ? c : -c) >= (sn >= 0 ? sn : -sn)) {
      constexpr float t = (float)(sn / c);
      pr = fmaf(t, b.y, b.x);
      pi = fmaf(-t, b.x, b.y);
      g = (float)c;
    } else {
      constexpr float ct = (float)(c / sn);
      pr = fmaf(ct, b.x, b.y);
      pi = fmaf(ct, b.y, -b.x);
      g = (float)sn;
    }
    const float ur = u.x, ui = u.y;
    u = make_float2(fmaf(g, pr, ur), fmaf(g, pi, ui));
    b = make_float2(fmaf(-g, pr, ur), fmaf(-g, pi, ui));
  }
}

template <int N, int M, int K, int J>
__device__ __forceinline__ void dit_stage_walk(float2 (&v)[N]) {
  dit_butterfly<J, M>(v[bit_reverse<N>(K + J)], v[bit_reverse<N>(K + J + M / 2)]);
  if constexpr (J + 1 < M / 2) dit_stage_walk<N, M, K, J + 1>(v);
  else if constexpr (K + M < N) dit_stage_walk<N, M, K + M, 0>(v);
}

template <int N, int M>
__device__ __forceinline__ void dit_stages(float2 (&v)[N]) {
  dit_stage_walk<N, M, 0, 0>(v);
  if constexpr (2 * M <= N) dit_stages<N, 2 * M>(v);
}

template <int N>
__device__ __forceinline__ void dit_fft_fma(float2 (&v)[N]) {
  static_assert(N >= 2 && N <= 64 && (N & (N - 1)) == 0, "N must be a power of two in [2, 64]");
  dit_stages<N, 2>(v);
}

// ---------------------------------------------------------------------------------------------
// The same FMA-form DIT network over a generic real type R (float, or the packed pair `pk` of f32x2.cuh: two
// frames per warp, every butterfly one FADD2 / FFMA2 per component for both).  cx<R> v[N]; contract as above.
// The packed instructions have no negated addend, so the cotangent form keeps -pi and flips the sign of g.
// ---------------------------------------------------------------------------------------------
template <int J, int M, class R>
__device__ __forceinline__ void dit_butterfly_r(cx<R>& u, cx<R>& b) {
  static_assert(M <= 64 && 64 % M == 0 && J >= 0 && 2 * J < M, "twiddle out of range");
  if constexpr (J == 0) {
    const R ur = u.x, ui = u.y;
    u.x = ur + b.x; u.y = ui + b.y;
    b.x = ur - b.x; b.y = ui - b.y;
  } else if constexpr (4 * J == M) {               // w = -i
    const R ur = u.x, ui = u.y, br = b.x, bi = b.y;
    u.x = ur + bi; u.y = ui - br;
    b.x = ur - bi; b.y = ui + br;
  } else {
    constexpr double c = (double)cos64(J * (64 / M)), sn = (double)sin64(J * (64 / M));
    const R ur = u.x, ui = u.y;
    if constexpr ((c >= 0 ? c : -c) >= (sn >= 0 ? sn : -sn)) {
      constexpr float t = (float)(sn / c), g = (float)c;
      const R pr = pfma(t, b.y, b.x);              // w b = g (pr + i pi)
      const R pi = pfma(-t, b.x, b.y);
      u.x = pfma(g, pr, ur); u.y = pfma(g, pi, ui);
      b.x = pfma(-g, pr, ur); b.y = pfma(-g, pi, ui);
    } else {
      constexpr float ct = (float)(c / sn), g = (float)sn;
      const R pr = pfma(ct, b.x, b.y);             // w b = g (pr - i npi)
      const R npi = pfma(-ct, b.y, b.x);
      u.x = pfma(g, pr, ur); u.y = pfma(-g, npi, ui);
      b.x = pfma(-g, pr, ur); b.y = pfma(g, npi, ui);
    }
  }
}

template <int N, int M, int K, int J, class R>
__device__ __forceinline__ void dit_stage_walk_r(cx<R> (&v)[N]) {
  dit_butterfly_r<J, M, R>(v[bit_reverse<N>(K + J)], v[bit_reverse<N>(K + J + M / 2)]);
  if constexpr (J + 1 < M / 2) dit_stage_walk_r<N, M, K, J + 1, R>(v);
  else if constexpr (K + M < N) dit_stage_walk_r<N, M, K + M, 0, R>(v);
}

template <int N, int M, class R>
__device__ __forceinline__ void dit_stages_r(cx<R> (&v)[N]) {
  dit_stage_walk_r<N, M, 0, 0, R>(v);
  if constexpr (2 * M <= N) dit_stages_r<N, 2 * M, R>(v);
}

template <int N, class R>
__device__ __forceinline__ void dit_fft_fma_r(cx<R> (&v)[N]) {
  static_assert(N >= 2 && N <= 64 && (N & (N - 1)) == 0, "N must be a power of two in [2, 64]");
  dit_stages_r<N, 2, R>(v);
}

}  // namespace tac
