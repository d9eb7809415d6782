// Stand-alone elementwise stages, used when the modules are composed by hand (e.g. with a
// TimeStretch in between, reference tests/test_layers.py:98-101).  In the Spectrogram /
// Melspectrogram pipelines these are fused into the STFT epilogue / the filterbank epilogue.
//
//   complex_norm     functional.py:116-128   (n, 2) -> (n):  sqrt(re^2 + im^2), then .pow(power)
//   amplitude_to_db  functional.py:277-296   10 * (log10(max(x^2, amin)) - log10(ref))
//   db_to_amplitude  functional.py:299-314   sqrt(10^(x / 10 + log10(ref)))
//   angle            functional.py:187-191   (n, 2) -> (n):  atan2(im, re)
//   magphase         functional.py:194-201   (n, 2) -> (n), (n):  complex_norm and angle in one pass
//
// Pure streaming: 12 B / element (complex_norm, angle), 16 (magphase), 8 (amplitude_to_db, db_to_amplitude).
#include "tac_common.cuh"

namespace tac {

constexpr int kPwThreads = 256;

// same evaluation order as the reference: norm first, power second (quirk 2 in SURVEY section 0)
__device__ __forceinline__ float norm_then_pow(float re, float im, float power, int mode) {
  const float mag = sqrtf(fmaf(re, re, im * im));
  if (mode == 1) return mag;                 // power == 1
  if (mode == 2) return mag * mag;           // torch.pow(x, 2.) == x * x
  if (mode == 3) return sqrtf(mag);          // power == 0.5
  return powf(mag, power);
}

static inline int power_mode(float power) {
  return power == 1.0f ? 1 : (power == 2.0f ? 2 : (power == 0.5f ? 3 : 0));
}

__global__ void __launch_bounds__(kPwThreads)
complex_norm_kernel(const float2* __restrict__ z, int64_t n, float power, int mode, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * kPwThreads;
  const int64_t n2 = n >> 1;
  const bool aligned = ((reinterpret_cast<uintptr_t>(z) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
  if (aligned) {
    for (int64_t i = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < n2; i += stride) {
      const float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(z) + i);
      float2 r;
      r.x = norm_then_pow(v.x, v.y, power, mode);
      r.y = norm_then_pow(v.z, v.w, power, mode);
      __stcs(reinterpret_cast<float2*>(out) + i, r);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
      const float2 v = z[n - 1];
      out[n - 1] = norm_then_pow(v.x, v.y, power, mode);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < n; i += stride) {
      const float2 v = z[i];
      out[i] = norm_then_pow(v.x, v.y, power, mode);
    }
  }
}

__device__ __forceinline__ float to_db(float x, float amin, float log10_ref) {
  float s = x * x;
  s = (s < amin) ? amin : s;                 // torch.clamp(min=): NaN stays NaN
  return 10.0f * (log10f(s) - log10_ref);
}

__global__ void __launch_bounds__(kPwThreads)
amplitude_to_db_kernel(const float* __restrict__ x, int64_t n, float amin, float log10_ref, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * kPwThreads;
  const int64_t n4 = n >> 2;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (aligned) {
    for (int64_t i = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < n4; i += stride) {
      const float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(x) + i);
      float4 r;
      r.x = to_db(v.x, amin, log10_ref);
      r.y = to_db(v.y, amin, log10_ref);
      r.z = to_db(v.z, amin, log10_ref);
      r.w = to_db(v.w, amin, log10_ref);
      __stcs(reinterpret_cast<float4*>(out) + i, r);
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < n; i += stride)
      out[i] = to_db(x[i], amin, log10_ref);
  } else {
    for (int64_t i = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < n; i += stride)
      out[i] = to_db(x[i], amin, log10_ref);
  }
}

// functional.py:310-314: torch.pow(10.0, x / 10.0 + log10(ref)) then .pow(0.5)
__device__ __forceinline__ float from_db(float x, float log10_ref) {
  return sqrtf(powf(10.0f, x / 10.0f + log10_ref));
}

__global__ void __launch_bounds__(kPwThreads)
db_to_amplitude_kernel(const float* __restrict__ x, int64_t n, float log10_ref, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * kPwThreads;
  const int64_t n4 = n >> 2;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (aligned) {
    for (int64_t i = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < n4; i += stride) {
      const float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(x) + i);
      float4 r;
      r.x = from_db(v.x, log10_ref);
      r.y = from_db(v.y, log10_ref);
      r.z = from_db(v.z, log10_ref);
      r.w = from_db(v.w, log10_ref);
      __stcs(reinterpret_cast<float4*>(out) + i, r);
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < n; i += stride)
      out[i] = from_db(x[i], log10_ref);
  } else {
    for (int64_t i = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < n; i += stride) out[i] = from_db(x[i], log10_ref);
  }
}

// angle (mag == nullptr) or magphase: one read of z, one or two coalesced writes
__global__ void __launch_bounds__(kPwThreads)
magphase_kernel(const float2* __restrict__ z, int64_t n, float power, int mode, float* __restrict__ mag,
                float* __restrict__ phase) {
  const int64_t stride = (int64_t)gridDim.x * kPwThreads;
  const int64_t n2 = n >> 1;
  const bool aligned = ((reinterpret_cast<uintptr_t>(z) & 15) == 0) && ((reinterpret_cast<uintptr_t>(phase) & 7) == 0) &&
                       ((reinterpret_cast<uintptr_t>(mag) & 7) == 0);
  if (aligned) {
    for (int64_t i = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < n2; i += stride) {
      const float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(z) + i);
      __stcs(reinterpret_cast<float2*>(phase) + i, make_float2(atan2f(v.y, v.x), atan2f(v.w, v.z)));
      if (mag) __stcs(reinterpret_cast<float2*>(mag) + i,
                      make_float2(norm_then_pow(v.x, v.y, power, mode), norm_then_pow(v.z, v.w, power, mode)));
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
      const float2 v = z[n - 1];
      phase[n - 1] = atan2f(v.y, v.x);
      if (mag) mag[n - 1] = norm_then_pow(v.x, v.y, power, mode);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < n; i += stride) {
      const float2 v = z[i];
      phase[i] = atan2f(v.y, v.x);
      if (mag) mag[i] = norm_then_pow(v.x, v.y, power, mode);
    }
  }
}

// Adjoints of the remaining pointwise operators (the reference differentiates them through torch).
//   op 0  db_to_amplitude (functional.py:299-314): y = sqrt(ref 10^(x/10)) -> dy/dx = y ln(10) / 20;   a = y, g1 = dL/dy
//   op 1  magphase / angle (functional.py:187-201): a = z (n x 2), g1 = dL/d|z|^p or null, g2 = dL/dphase or null;
//         d|z|^p/dz = p |z|^(p-2) z (0 at z = 0, torch.norm's subgradient), dphase/dz = (-im, re) / |z|^2;  out = dL/dz (n x 2)
//   op 2  mu_law_decoding of float codes (functional.py:349-354): x = 2 c / mu - 1, y = sign(x) (exp(|x| L) - 1) / mu,
//         L = log1p(mu) -> dy/dc = 2 L exp(|x| L) / mu^2;   a = c, g1 = dL/dy, p0 = mu
__global__ void __launch_bounds__(kPwThreads)
pointwise_backward_kernel(int op, const float* __restrict__ a, const float* __restrict__ g1, const float* __restrict__ g2,
                          int64_t n, float p0, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * kPwThreads;
  for (int64_t i = (int64_t)blockIdx.x * kPwThreads + threadIdx.x; i < n; i += stride) {
    if (op == 0) {
      out[i] = g1[i] * a[i] * 0.11512925464970229f;                     // ln(10) / 20
    } else if (op == 1) {
      const float2 z = reinterpret_cast<const float2*>(a)[i];
      const float n2 = fmaf(z.x, z.x, z.y * z.y);
      float gr = 0.0f, gi = 0.0f;
      if (n2 > 0.0f) {
        if (g1) {
          const float k = g1[i] * (p0 == 1.0f ? rsqrtf(n2) : (p0 == 2.0f ? 2.0f : p0 * powf(n2, 0.5f * p0 - 1.0f)));
          gr = k * z.x;
          gi = k * z.y;
        }
        if (g2) {
          const float k = g2[i] / n2;
          gr = fmaf(-k, z.y, gr);
          gi = fmaf(k, z.x, gi);
        }
      }
      reinterpret_cast<float2*>(out)[i] = make_float2(gr, gi);
    } else {
      const float mu = p0, l1p = log1pf(mu);
      const float x = (a[i] / mu) * 2.0f - 1.0f;
      out[i] = g1[i] * (2.0f * l1p / (mu * mu)) * expf(fabsf(x) * l1p);
    }
  }
}

static int pw_grid(int64_t n_vec) {
  const int64_t want = (n_vec + kPwThreads - 1) / kPwThreads;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace tac

extern "C" int tac_complex_norm_f32(const float* z, int64_t n, float power, float* out, void* stream) {
  using namespace tac;
  TAC_REQUIRE(n >= 0, TAC_ERR_INVALID, "complex_norm: n=%lld", (long long)n);
  if (n == 0) return TAC_OK;
  TAC_REQUIRE(z && out, TAC_ERR_INVALID, "complex_norm: null pointer");
  LaunchProbe probe(KIND_POINTWISE, as_stream(stream));
  complex_norm_kernel<<<pw_grid((n + 1) / 2), kPwThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const float2*>(z), n, power, power_mode(power), out);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int tac_amplitude_to_db_f32(const float* x, int64_t n, float ref, float amin, float* out, void* stream) {
  using namespace tac;
  TAC_REQUIRE(n >= 0, TAC_ERR_INVALID, "amplitude_to_db: n=%lld", (long long)n);
  if (n == 0) return TAC_OK;
  TAC_REQUIRE(x && out, TAC_ERR_INVALID, "amplitude_to_db: null pointer");
  LaunchProbe probe(KIND_POINTWISE, as_stream(stream));
  amplitude_to_db_kernel<<<pw_grid((n + 3) / 4), kPwThreads, 0, as_stream(stream)>>>(x, n, amin, log10f(ref), out);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int tac_db_to_amplitude_f32(const float* x, int64_t n, float ref, float* out, void* stream) {
  using namespace tac;
  TAC_REQUIRE(n >= 0, TAC_ERR_INVALID, "db_to_amplitude: n=%lld", (long long)n);
  if (n == 0) return TAC_OK;
  TAC_REQUIRE(x && out, TAC_ERR_INVALID, "db_to_amplitude: null pointer");
  LaunchProbe probe(KIND_POINTWISE, as_stream(stream));
  db_to_amplitude_kernel<<<pw_grid((n + 3) / 4), kPwThreads, 0, as_stream(stream)>>>(x, n, log10f(ref), out);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int tac_magphase_f32(const float* z, int64_t n, float power, float* mag, float* phase, void* stream) {
  using namespace tac;
  TAC_REQUIRE(n >= 0, TAC_ERR_INVALID, "magphase: n=%lld", (long long)n);
  if (n == 0) return TAC_OK;
  TAC_REQUIRE(z && phase, TAC_ERR_INVALID, "magphase: null pointer");
  LaunchProbe probe(KIND_POINTWISE, as_stream(stream));
  magphase_kernel<<<pw_grid((n + 1) / 2), kPwThreads, 0, as_stream(stream)>>>(reinterpret_cast<const float2*>(z), n, power,
                                                                              power_mode(power), mag, phase);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int tac_pointwise_backward_f32(int op, const float* a, const float* g1, const float* g2, int64_t n, float p0,
                                          float* out, void* stream) {
  using namespace tac;
  TAC_REQUIRE(n >= 0 && op >= 0 && op <= 2, TAC_ERR_INVALID, "pointwise_backward: op=%d n=%lld", op, (long long)n);
  if (n == 0) return TAC_OK;
  TAC_REQUIRE(a && out && (g1 || (op == 1 && g2)), TAC_ERR_INVALID, "pointwise_backward: null pointer");
  LaunchProbe probe(KIND_POINTWISE, as_stream(stream));
  pointwise_backward_kernel<<<pw_grid(n), kPwThreads, 0, as_stream(stream)>>>(op, a, g1, g2, n, p0, out);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}
