// float64 operator path.  The reference is dtype-generic: functional.py:99 hands a double waveform to torch.stft and every
// later stage (:126-128, :183, :291-296, :310-314, :349-354) computes in the dtype it is given.  These kernels keep that
// contract for double tensors -- plain, correct-first kernels, double arithmetic throughout; the tuned float32 kernels
// are the product path, this is the path a double-precision check of a pipeline (or a gradcheck harness) takes.
//
//   tac_stft_f64             framing + padding + window + DFT, direct N (N/2 + 1) sums from a double table of
//                            exp(-2 pi i j / N) (any n_fft in [2, 4096]); (n_seq, bins, frames, 2) double
//   tac_complex_norm_f64     sqrt(re^2 + im^2) then pow
//   tac_apply_filterbank_f64 out[s, m, t] = sum_k spec[s, k, t] fb[k, m]        (dense, any matrix)
//   tac_amplitude_to_db_f64 / tac_db_to_amplitude_f64 / tac_magphase_f64
//   tac_mulaw_decode_i64_f64 table of the n_quantize decoded doubles (built by the caller with the reference formula)
//   tac_mulaw_encode_f64_i64 the reference formula in double on the device (log1p), truncation toward zero
#include "stft_params.cuh"
#include "tac_common.cuh"

namespace tac {

constexpr int kF64Threads = 256;

__device__ __forceinline__ double fetch_padded_f64(const double* __restrict__ row, int64_t s, int64_t n, int pad_mode) {
  if (s >= 0 && s < n) return row[s];
  switch (pad_mode) {
    case 0: s = (s < 0) ? -s : 2 * (n - 1) - s; break;          // reflect
    case 2: s = (s < 0) ? 0 : n - 1; break;                     // replicate
    case 3: s = (s < 0) ? s + n : s - n; break;                 // circular
    default: return 0.0;                                        // constant
  }
  return (s >= 0 && s < n) ? row[s] : 0.0;
}

struct StftF64Params {
  const double* x;
  const double* window;
  double* out;
  int64_t n_seq, n_samples, seq_stride, frames;
  int n_fft, hop, pad, pad_mode, onesided, bins;
  double scale;
};

__global__ void __launch_bounds__(kF64Threads) stft_f64_kernel(const StftF64Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.n_fft, nb = N / 2 + 1;
  double2* tab = reinterpret_cast<double2*>(smem_raw);
  double* xw = reinterpret_cast<double*>(tab + N);
  const int tid = threadIdx.x;
  for (int j = tid; j < N; j += kF64Threads) {
    double sn, cs;
    sincospi(-2.0 * (double)j / (double)N, &sn, &cs);
    tab[j] = make_double2(cs, sn);
  }
  __syncthreads();
  const int64_t total = p.n_seq * p.frames;
  for (int64_t g = blockIdx.x; g < total; g += gridDim.x) {
    const int64_t seq = g / p.frames, t = g - seq * p.frames, start = t * p.hop - p.pad;
    const double* row = p.x + seq * p.seq_stride;
    for (int n = tid; n < N; n += kF64Threads) xw[n] = fetch_padded_f64(row, start + n, p.n_samples, p.pad_mode) * (p.window[n] * p.scale);
    __syncthreads();
    for (int k = tid; k < nb; k += kF64Threads) {
      double re = 0.0, im = 0.0;
      int idx = 0;
      for (int n = 0; n < N; ++n) {
        const double2 w = tab[idx];
        const double v = xw[n];
        re = fma(v, w.x, re);
        im = fma(v, w.y, im);
        idx += k;
        idx -= (idx >= N) ? N : 0;
      }
      if (k == 0 || 2 * k == N) im = 0.0;
      reinterpret_cast<double2*>(p.out)[(seq * p.bins + k) * p.frames + t] = make_double2(re, im);
      if (!p.onesided && k > 0 && 2 * k != N) reinterpret_cast<double2*>(p.out)[(seq * p.bins + (N - k)) * p.frames + t] = make_double2(re, -im);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ double norm_then_pow_f64(double re, double im, double power) {
  const double mag = sqrt(fma(re, re, im * im));
  if (power == 1.0) return mag;
  if (power == 2.0) return mag * mag;
  return pow(mag, power);
}

__global__ void __launch_bounds__(kF64Threads) complex_norm_f64_kernel(const double2* __restrict__ z, int64_t n, double power, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * kF64Threads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kF64Threads) {
    const double2 v = z[i];
    out[i] = norm_then_pow_f64(v.x, v.y, power);
  }
}

__global__ void __launch_bounds__(kF64Threads) magphase_f64_kernel(const double2* __restrict__ z, int64_t n, double power, double* __restrict__ mag,
                                                                  double* __restrict__ phase) {
  for (int64_t i = (int64_t)blockIdx.x * kF64Threads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kF64Threads) {
    const double2 v = z[i];
    if (mag) mag[i] = norm_then_pow_f64(v.x, v.y, power);
    phase[i] = atan2(v.y, v.x);
  }
}

__global__ void __launch_bounds__(kF64Threads) amplitude_to_db_f64_kernel(const double* __restrict__ x, int64_t n, double amin, double log10_ref,
                                                                         double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * kF64Threads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kF64Threads) {
    double s = x[i] * x[i];
    s = (s < amin) ? amin : s;
    out[i] = 10.0 * (log10(s) - log10_ref);
  }
}

__global__ void __launch_bounds__(kF64Threads) db_to_amplitude_f64_kernel(const double* __restrict__ x, int64_t n, double log10_ref, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * kF64Threads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kF64Threads)
    out[i] = sqrt(pow(10.0, x[i] / 10.0 + log10_ref));
}

// one thread per (band, frame) output, frames innermost (coalesced reads of spec rows and writes); the matrix column is
// read through the read-only cache
__global__ void __launch_bounds__(kF64Threads) apply_filterbank_f64_kernel(const double* __restrict__ spec, const double* __restrict__ fb, int64_t n_seq,
                                                                          int64_t frames, int n_bins, int n_bands, double* __restrict__ out) {
  const int64_t per_seq = (int64_t)n_bands * frames, total = n_seq * per_seq;
  for (int64_t i = (int64_t)blockIdx.x * kF64Threads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kF64Threads) {
    const int64_t s = i / per_seq, r = i - s * per_seq;
    const int m = (int)(r / frames);
    const int64_t t = r - (int64_t)m * frames;
    const double* col = spec + s * (int64_t)n_bins * frames + t;
    double acc = 0.0;
    for (int k = 0; k < n_bins; ++k) acc = fma(col[(int64_t)k * frames], __ldg(fb + (int64_t)k * n_bands + m), acc);
    out[i] = acc;
  }
}

__global__ void __launch_bounds__(kF64Threads) mulaw_decode_f64_kernel(const long long* __restrict__ codes, int64_t n, int n_quantize,
                                                                      const double* __restrict__ lut, double mu, double log1p_mu, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * kF64Threads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kF64Threads) {
    const long long c = codes[i];
    if (c >= 0 && c < n_quantize) {
      out[i] = __ldg(lut + c);
    } else {                                                    // functional.py:351-353 outside the table
      const double y = ((double)c / mu) * 2.0 - 1.0;
      const double sgn = y > 0.0 ? 1.0 : (y < 0.0 ? -1.0 : 0.0);
      out[i] = sgn * (exp(fabs(y) * log1p_mu) - 1.0) / mu;
    }
  }
}

__global__ void __launch_bounds__(kF64Threads) mulaw_encode_f64_kernel(const double* __restrict__ x, int64_t n, double mu, double log1p_mu,
                                                                      long long* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * kF64Threads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kF64Threads) {
    const double v = x[i];
    const double sgn = v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0);
    const double comp = sgn * log1p(mu * fabs(v)) / log1p_mu;   // functional.py:331-334
    const double q = (comp + 1.0) / 2.0 * mu + 0.5;
    out[i] = (fabs(q) < 9.0e18) ? (long long)q : (long long)0x8000000000000000ull;
  }
}

static int grid_for(int64_t n) {
  const int64_t want = (n + kF64Threads - 1) / kF64Threads;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace tac

extern "C" int tac_stft_f64(const double* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride, const double* window, int n_fft,
                            int hop, int center, int pad_mode, int normalized, int onesided, double* out, void* stream) {
  using namespace tac;
  TAC_REQUIRE((x || n_seq == 0) && window, TAC_ERR_INVALID, "stft_f64: null input or window pointer");
  TAC_REQUIRE(n_seq >= 0 && n_samples >= 0 && seq_stride >= n_samples, TAC_ERR_INVALID, "stft_f64: bad shape");
  TAC_REQUIRE(n_fft >= 2 && n_fft <= 4096, TAC_ERR_UNSUPPORTED, "stft_f64: n_fft=%d outside [2, 4096]", n_fft);
  TAC_REQUIRE(hop >= 1, TAC_ERR_INVALID, "stft_f64: hop_length=%d must be positive", hop);
  TAC_REQUIRE(pad_mode >= TAC_PAD_REFLECT && pad_mode <= TAC_PAD_CIRCULAR, TAC_ERR_INVALID, "stft_f64: unknown pad_mode %d", pad_mode);
  const int pad = center ? n_fft / 2 : 0;
  if (center && pad_mode == TAC_PAD_REFLECT)
    TAC_REQUIRE(pad < n_samples, TAC_ERR_INVALID,
                "stft: Padding size should be less than the corresponding input dimension (reflect pad %d, time %lld)", pad, (long long)n_samples);
  if (center && pad_mode == TAC_PAD_CIRCULAR)
    TAC_REQUIRE(pad <= n_samples, TAC_ERR_INVALID, "stft: circular padding %d wraps more than once (time %lld)", pad, (long long)n_samples);
  TAC_REQUIRE(n_samples + 2 * pad >= n_fft, TAC_ERR_INVALID, "stft: input of %lld samples is shorter than n_fft=%d", (long long)n_samples, n_fft);
  StftF64Params p;
  p.x = x; p.window = window; p.out = out;
  p.n_seq = n_seq; p.n_samples = n_samples; p.seq_stride = seq_stride;
  p.frames = tac_stft_num_frames(n_samples, n_fft, hop, center);
  p.n_fft = n_fft; p.hop = hop; p.pad = pad; p.pad_mode = pad_mode; p.onesided = onesided ? 1 : 0;
  p.bins = onesided ? n_fft / 2 + 1 : n_fft;
  p.scale = normalized ? 1.0 / sqrt((double)n_fft) : 1.0;
  const int64_t total = n_seq * p.frames;
  if (total <= 0) return TAC_OK;
  TAC_REQUIRE(out, TAC_ERR_INVALID, "stft_f64: null output pointer");
  const size_t smem = sizeof(double2) * (size_t)n_fft + sizeof(double) * (size_t)n_fft;
  if (smem > 48 * 1024) TAC_CUDA_OK(cudaFuncSetAttribute(stft_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t cap = (int64_t)sm_count() * 4;
  LaunchProbe probe(KIND_STFT, as_stream(stream));
  stft_f64_kernel<<<(int)(total < cap ? total : cap), kF64Threads, smem, as_stream(stream)>>>(p);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int tac_complex_norm_f64(const double* z, int64_t n, double power, double* out, void* stream) {
  using namespace tac;
  TAC_REQUIRE(n >= 0 && ((z && out) || n == 0), TAC_ERR_INVALID, "complex_norm_f64: bad arguments");
  if (n == 0) return TAC_OK;
  LaunchProbe probe(KIND_POINTWISE, as_stream(stream));
  complex_norm_f64_kernel<<<grid_for(n), kF64Threads, 0, as_stream(stream)>>>(reinterpret_cast<const double2*>(z), n, power, out);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int tac_magphase_f64(const double* z, int64_t n, double power, double* mag, double* phase, void* stream) {
  using namespace tac;
  TAC_REQUIRE(n >= 0 && ((z && phase) || n == 0), TAC_ERR_INVALID, "magphase_f64: bad arguments");
  if (n == 0) return TAC_OK;
  LaunchProbe probe(KIND_POINTWISE, as_stream(stream));
  magphase_f64_kernel<<<grid_for(n), kF64Threads, 0, as_stream(stream)>>>(reinterpret_cast<const double2*>(z), n, power, mag, phase);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int tac_amplitude_to_db_f64(const double* x, int64_t n, double ref, double amin, double* out, void* stream) {
  using namespace tac;
  TAC_REQUIRE(n >= 0 && ((x && out) || n == 0), TAC_ERR_INVALID, "amplitude_to_db_f64: bad arguments");
  if (n == 0) return TAC_OK;
  LaunchProbe probe(KIND_POINTWISE, as_stream(stream));
  amplitude_to_db_f64_kernel<<<grid_for(n), kF64Threads, 0, as_stream(stream)>>>(x, n, amin, log10(ref), out);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int tac_db_to_amplitude_f64(const double* x, int64_t n, double ref, double* out, void* stream) {
  using namespace tac;
  TAC_REQUIRE(n >= 0 && ((x && out) || n == 0), TAC_ERR_INVALID, "db_to_amplitude_f64: bad arguments");
  if (n == 0) return TAC_OK;
  LaunchProbe probe(KIND_POINTWISE, as_stream(stream));
  db_to_amplitude_f64_kernel<<<grid_for(n), kF64Threads, 0, as_stream(stream)>>>(x, n, log10(ref), out);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int tac_apply_filterbank_f64(const double* spec, const double* fb_dev, int64_t n_seq, int64_t frames, int n_bins, int n_bands,
                                        double* out, void* stream) {
  using namespace tac;
  TAC_REQUIRE(n_seq >= 0 && frames >= 0 && n_bins >= 1 && n_bands >= 1, TAC_ERR_INVALID, "apply_filterbank_f64: bad shape");
  const int64_t total = n_seq * frames * n_bands;
  if (total == 0) return TAC_OK;
  TAC_REQUIRE(spec && fb_dev && out, TAC_ERR_INVALID, "apply_filterbank_f64: null pointer");
  LaunchProbe probe(KIND_MELBANK, as_stream(stream));
  apply_filterbank_f64_kernel<<<grid_for(total), kF64Threads, 0, as_stream(stream)>>>(spec, fb_dev, n_seq, frames, n_bins, n_bands, out);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int tac_mulaw_decode_i64_f64(const int64_t* codes, int64_t n, int n_quantize, const double* lut_dev, double* out, void* stream) {
  using namespace tac;
  TAC_REQUIRE(n >= 0 && n_quantize >= 2 && ((codes && out && lut_dev) || n == 0), TAC_ERR_INVALID, "mulaw_decode_f64: bad arguments");
  if (n == 0) return TAC_OK;
  const double mu = (double)(n_quantize - 1);
  LaunchProbe probe(KIND_MULAW, as_stream(stream));
  mulaw_decode_f64_kernel<<<grid_for(n), kF64Threads, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(codes), n, n_quantize, lut_dev, mu,
                                                                           log1p(mu), out);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int tac_mulaw_encode_f64_i64(const double* x, int64_t n, int n_quantize, int64_t* out, void* stream) {
  using namespace tac;
  TAC_REQUIRE(n >= 0 && n_quantize >= 2 && ((x && out) || n == 0), TAC_ERR_INVALID, "mulaw_encode_f64: bad arguments");
  if (n == 0) return TAC_OK;
  const double mu = (double)(n_quantize - 1);
  LaunchProbe probe(KIND_MULAW, as_stream(stream));
  mulaw_encode_f64_kernel<<<grid_for(n), kF64Threads, 0, as_stream(stream)>>>(x, n, mu, log1p(mu), reinterpret_cast<long long*>(out));
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}
