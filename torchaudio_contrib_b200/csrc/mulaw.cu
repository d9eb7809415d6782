// K4 / K5: mu-law companding (reference: torchaudio_contrib/functional.py:317-354).
//
// Both kernels are pure streaming kernels: 12 algorithmic bytes per sample (fp32 in + int64 out,
// or int64 in + fp32 out), HBM-bound.  Bit-exactness with the reference's fp32 torch chain is
// obtained by table, not by re-deriving its transcendental functions:
//
//  encode  The reference quantiser  idx(x) = trunc(((s*log1p(mu|x|)/log1p(mu) + 1)/2)*mu + 0.5)
//          is monotone in x (checked exhaustively over all 2^32 floats), so it is fully
//          described by its decision levels thr[k] = min{x : idx(x) >= k}.  The kernel computes a
//          cheap estimate k0 = floor(v - 0.5) with the hardware lg2 and settles the last step with
//          ONE table compare:  idx = k0 + (x >= thr[k0 + 1]).  The estimate only has to be within
//          +-0.5 of the real-valued quantiser input, so the approximation error of lg2.approx
//          (1e-6) is irrelevant to the result.
//  decode  Codes 0..n_quantize-1 index a lookup table of the reference's decoded values; anything
//          else goes through the closed form.
//
// The tables are passed in as device pointers.  tac_mulaw_tables_host hands them out on the host: for n_quantize = 256
// (the reference's default) the levels found with the reference's own fp32 torch CPU chain are SHIPPED in the library
// (mulaw_table256.inc, scripts/gen_mulaw_table.py), so any caller of the C ABI is bit-exact without torch; other
// n_quantize are bisected here with the host libm (the Python package passes its torch-built table instead).
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "tac_common.cuh"

namespace tac {

#include "mulaw_table256.inc"

// ---- host-side decision levels (any n_quantize): the Python package's bisection, with libm ----------------------
static inline float key_to_float(int64_t key) {          // monotone key in [0, 2^32) -> the float it denotes
  const uint32_t bits = key >= ((int64_t)1 << 31) ? (uint32_t)(key - ((int64_t)1 << 31)) : (uint32_t)(((int64_t)1 << 32) - 1 - key);
  float f;
  memcpy(&f, &bits, 4);
  return f;
}
static inline int64_t float_to_key(float f) {
  uint32_t bits;
  memcpy(&bits, &f, 4);
  return bits >= 0x80000000u ? ((int64_t)1 << 32) - 1 - (int64_t)bits : (int64_t)bits + ((int64_t)1 << 31);
}
static inline int64_t quantise_host(float x, float mu) {   // functional.py:331-334 in fp32
  const float a = fabsf(x);
  const float sgn = x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f);
  const float comp = sgn * log1pf(mu * a) / log1pf(mu);
  const float v = (comp + 1.0f) / 2.0f * mu + 0.5f;
  if (!(fabsf(v) < 9.0e18f)) return INT64_MIN;              // what the float -> int64 conversion yields for inf / NaN
  return (int64_t)v;
}

constexpr int kMuLawThreads = 256;
constexpr int kMuLawSmemTableMax = 12288;   // floats (48 KB) -- n_quantize = 256 needs 4081

template <bool kSmemTable>
__global__ void __launch_bounds__(kMuLawThreads)
mulaw_encode_kernel(const float* __restrict__ x, int64_t n, long long* __restrict__ out,
                    const float* __restrict__ thr, int n_thr, int idx_min, float x_limit,
                    float mu, float half_mu_over_log2) {
  extern __shared__ float s_thr[];
  const float* table = thr;
  if (kSmemTable) {
    for (int i = threadIdx.x; i < n_thr; i += kMuLawThreads) s_thr[i] = thr[i];
    __syncthreads();
    table = s_thr;
  }
  const float half_mu = 0.5f * mu;
  const int j_max = n_thr - 1;

  auto encode_one = [&](float v) -> long long {
    const float a = fabsf(v);
    if (!(a <= x_limit)) return (long long)0x8000000000000000ull;     // overflow / NaN in the reference
    // real-valued quantiser input minus 0.5: mu/2 * (1 + sign * log(1+mu|x|)/log(1+mu))
    const float l = __log2f(fmaf(mu, a, 1.0f)) * half_mu_over_log2;
    const float est = half_mu + copysignf(l, v);
    // the reference truncates toward zero: floor for v >= 0, ceil for v < 0 (only reached for x < -1)
    int j = __float2int_rd(est) + (est < -0.5f ? 2 : 1) - idx_min;     // candidate index k0 + 1
    j = max(0, min(j, j_max));
    const float t = kSmemTable ? table[j] : __ldg(table + j);
    return (long long)(idx_min + j - 1 + (v >= t ? 1 : 0));
  };

  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * kMuLawThreads;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (aligned) {
    for (int64_t i = (int64_t)blockIdx.x * kMuLawThreads + threadIdx.x; i < n4; i += stride) {
      const float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(x) + i);
      longlong2 lo, hi;
      lo.x = encode_one(v.x);
      lo.y = encode_one(v.y);
      hi.x = encode_one(v.z);
      hi.y = encode_one(v.w);
      longlong2* o = reinterpret_cast<longlong2*>(out) + 2 * i;
      __stcs(o, lo);
      __stcs(o + 1, hi);
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * kMuLawThreads + threadIdx.x; i < n; i += stride)
      out[i] = encode_one(x[i]);
  } else {
    for (int64_t i = (int64_t)blockIdx.x * kMuLawThreads + threadIdx.x; i < n; i += stride)
      out[i] = encode_one(x[i]);
  }
}

__device__ __forceinline__ float mulaw_expand_closed_form(float code, float mu, float log1p_mu) {
  // functional.py:352-353, evaluated with the device exp (only reached for codes outside the table)
  const float y = (code / mu) * 2.0f - 1.0f;
  const float m = (expf(fabsf(y) * log1p_mu) - 1.0f) / mu;
  const float s = (y > 0.0f) ? 1.0f : ((y < 0.0f) ? -1.0f : 0.0f);
  return s * m;
}

template <typename CodeT>
__global__ void __launch_bounds__(kMuLawThreads)
mulaw_decode_kernel(const CodeT* __restrict__ codes, int64_t n, float* __restrict__ out,
                    const float* __restrict__ lut, int n_quantize, float mu, float log1p_mu) {
  extern __shared__ float s_lut[];
  for (int i = threadIdx.x; i < n_quantize; i += kMuLawThreads) s_lut[i] = lut[i];
  __syncthreads();

  auto decode_one = [&](CodeT c) -> float {
    if constexpr (sizeof(CodeT) == 8) {
      const long long k = (long long)c;
      if (k >= 0 && k < n_quantize) return s_lut[(int)k];
      return mulaw_expand_closed_form((float)k, mu, log1p_mu);
    } else {
      const float f = (float)c;
      const int k = __float2int_rz(f);
      if (f >= 0.0f && f < (float)n_quantize && (float)k == f) return s_lut[k];
      return mulaw_expand_closed_form(f, mu, log1p_mu);
    }
  };

  const int64_t stride = (int64_t)gridDim.x * kMuLawThreads;
  const int64_t n4 = n >> 2;
  const bool aligned = ((reinterpret_cast<uintptr_t>(codes) & 31) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (aligned) {
    for (int64_t i = (int64_t)blockIdx.x * kMuLawThreads + threadIdx.x; i < n4; i += stride) {
      CodeT c[4];
      if constexpr (sizeof(CodeT) == 8) {
        const longlong2 a = __ldcs(reinterpret_cast<const longlong2*>(codes) + 2 * i);
        const longlong2 b = __ldcs(reinterpret_cast<const longlong2*>(codes) + 2 * i + 1);
        c[0] = (CodeT)a.x; c[1] = (CodeT)a.y; c[2] = (CodeT)b.x; c[3] = (CodeT)b.y;
      } else {
        const float4 a = ldg_stream_f4(reinterpret_cast<const float4*>(codes) + i);
        c[0] = (CodeT)a.x; c[1] = (CodeT)a.y; c[2] = (CodeT)a.z; c[3] = (CodeT)a.w;
      }
      float4 r;
      r.x = decode_one(c[0]); r.y = decode_one(c[1]); r.z = decode_one(c[2]); r.w = decode_one(c[3]);
      __stcs(reinterpret_cast<float4*>(out) + i, r);
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * kMuLawThreads + threadIdx.x; i < n; i += stride)
      out[i] = decode_one(codes[i]);
  } else {
    for (int64_t i = (int64_t)blockIdx.x * kMuLawThreads + threadIdx.x; i < n; i += stride)
      out[i] = decode_one(codes[i]);
  }
}

static int streaming_grid(int64_t n_vec, int ctas_per_sm = 8) {
  const int64_t want = (n_vec + kMuLawThreads - 1) / kMuLawThreads;
  const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace tac

extern "C" int tac_mulaw_tables_host(int n_quantize, float* thresholds, int capacity, int* n_thresholds, int* idx_min,
                                     float* x_limit, float* decoded, int* exact) {
  using namespace tac;
  TAC_REQUIRE(n_quantize >= 2 && n_quantize <= 65536, TAC_ERR_INVALID, "mulaw_tables: n_quantize %d outside [2, 65536]", n_quantize);
  TAC_REQUIRE(n_thresholds && idx_min && x_limit, TAC_ERR_INVALID, "mulaw_tables: null output argument");
  if (n_quantize == 256) {
    *n_thresholds = kMuLaw256NThr;
    *idx_min = kMuLaw256IdxMin;
    memcpy(x_limit, &kMuLaw256XLimitBits, 4);
    if (exact) *exact = 1;
    if (thresholds) {
      TAC_REQUIRE(capacity >= kMuLaw256NThr, TAC_ERR_WORKSPACE, "mulaw_tables: %d thresholds, room for %d", kMuLaw256NThr, capacity);
      memcpy(thresholds, kMuLaw256Thr, sizeof(kMuLaw256Thr));
    }
    if (decoded) memcpy(decoded, kMuLaw256Dec, sizeof(kMuLaw256Dec));
    return TAC_OK;
  }
  const float mu = (float)(n_quantize - 1);
  // largest |x| whose code is still finite
  int64_t lo = (int64_t)1 << 31, hi = float_to_key(3.4028234663852886e38f);
  int64_t lim_key = hi;
  if (quantise_host(key_to_float(hi), mu) == INT64_MIN) {
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) / 2;
      if (quantise_host(key_to_float(mid), mu) != INT64_MIN) lo = mid; else hi = mid;
    }
    lim_key = lo;
  }
  const float lim = key_to_float(lim_key);
  const int64_t neg_key = float_to_key(-lim);
  const int64_t i_min = quantise_host(-lim, mu), i_max = quantise_host(lim, mu);
  const int64_t count = i_max - i_min + 1;
  TAC_REQUIRE(count >= 1 && count < ((int64_t)1 << 24), TAC_ERR_UNSUPPORTED, "mulaw_tables: %lld levels", (long long)count);
  *n_thresholds = (int)count;
  *idx_min = (int)i_min;
  *x_limit = lim;
  if (exact) *exact = 0;
  if (thresholds) {
    TAC_REQUIRE(capacity >= count, TAC_ERR_WORKSPACE, "mulaw_tables: %lld thresholds, room for %d", (long long)count, capacity);
    thresholds[0] = -INFINITY;
    for (int64_t t = i_min + 1; t <= i_max; ++t) {           // smallest float whose code is >= t
      int64_t a = neg_key, b = lim_key;                       // code(a) < t <= code(b)
      while (b - a > 1) {
        const int64_t mid = (a + b) / 2;
        if (quantise_host(key_to_float(mid), mu) >= t) b = mid; else a = mid;
      }
      thresholds[t - i_min] = key_to_float(b);
    }
  }
  if (decoded) {
    const float l1p = log1pf(mu);
    for (int c = 0; c < n_quantize; ++c) {                    // functional.py:351-353
      const float y = ((float)c / mu) * 2.0f - 1.0f;
      const float sgn = y > 0.0f ? 1.0f : (y < 0.0f ? -1.0f : 0.0f);
      decoded[c] = sgn * (expf(fabsf(y) * l1p) - 1.0f) / mu;
    }
  }
  return TAC_OK;
}

extern "C" int tac_mulaw_encode_f32_i64(const float* x, int64_t n, int n_quantize, const float* thresholds_dev,
                                        int n_thresholds, int idx_min, float x_limit, int64_t* out, void* stream) {
  using namespace tac;
  TAC_REQUIRE(n >= 0 && n_quantize >= 2, TAC_ERR_INVALID, "mulaw_encode: n=%lld n_quantize=%d", (long long)n, n_quantize);
  if (n == 0) return TAC_OK;
  TAC_REQUIRE(x && out && thresholds_dev && n_thresholds >= 1, TAC_ERR_INVALID, "mulaw_encode: null pointer / empty table");
  const float mu = (float)(n_quantize - 1);
  const float half_mu_over_log2 = (float)(0.5 * (double)mu / log2(1.0 + (double)mu));
  // 6 CTAs (1536 threads) per SM: measured 92.8 % of the HBM copy peak, against 74.6 % at full occupancy
  // (8 CTAs) and 78 % at 4 -- profiles/r01_notes.md
  const int grid = streaming_grid((n + 3) / 4, 6);
  LaunchProbe probe(KIND_MULAW, as_stream(stream));
  if (n_thresholds <= kMuLawSmemTableMax) {
    const size_t smem = (size_t)n_thresholds * sizeof(float);
    mulaw_encode_kernel<true><<<grid, kMuLawThreads, smem, as_stream(stream)>>>(
        x, n, reinterpret_cast<long long*>(out), thresholds_dev, n_thresholds, idx_min, x_limit, mu, half_mu_over_log2);
  } else {
    mulaw_encode_kernel<false><<<grid, kMuLawThreads, 0, as_stream(stream)>>>(
        x, n, reinterpret_cast<long long*>(out), thresholds_dev, n_thresholds, idx_min, x_limit, mu, half_mu_over_log2);
  }
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

template <typename CodeT>
static int launch_decode(const CodeT* codes, int64_t n, int n_quantize, const float* lut_dev, float* out, void* stream) {
  using namespace tac;
  TAC_REQUIRE(n >= 0 && n_quantize >= 2, TAC_ERR_INVALID, "mulaw_decode: n=%lld n_quantize=%d", (long long)n, n_quantize);
  if (n == 0) return TAC_OK;
  TAC_REQUIRE(codes && out && lut_dev, TAC_ERR_INVALID, "mulaw_decode: null pointer");
  TAC_REQUIRE(n_quantize <= kMuLawSmemTableMax, TAC_ERR_UNSUPPORTED, "mulaw_decode: n_quantize=%d exceeds the %d-entry table",
              n_quantize, kMuLawSmemTableMax);
  const float mu = (float)(n_quantize - 1);
  const float log1p_mu = (float)log1p((double)mu);
  const int grid = streaming_grid((n + 3) / 4, 6);
  LaunchProbe probe(KIND_MULAW, as_stream(stream));
  mulaw_decode_kernel<CodeT><<<grid, kMuLawThreads, (size_t)n_quantize * sizeof(float), as_stream(stream)>>>(
      codes, n, out, lut_dev, n_quantize, mu, log1p_mu);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int tac_mulaw_decode_i64_f32(const int64_t* codes, int64_t n, int n_quantize, const float* lut_dev,
                                        float* out, void* stream) {
  return launch_decode<long long>(reinterpret_cast<const long long*>(codes), n, n_quantize, lut_dev, out, stream);
}

extern "C" int tac_mulaw_decode_f32_f32(const float* codes, int64_t n, int n_quantize, const float* lut_dev,
                                        float* out, void* stream) {
  return launch_decode<float>(codes, n, n_quantize, lut_dev, out, stream);
}
