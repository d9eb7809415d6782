// Backward passes of the hot path (the reference is differentiable w.r.t. the waveform through
// torch.stft, torch.norm/.pow, torch.matmul and the dB chain -- functional.py:99-107, :126-128, :183-184, :291-296).
//
//   stft_backward_kernel        d loss / d waveform from d loss / d stft (complex) or from d loss / d |stft|^p
//                               (the spectrum is recomputed from the waveform instead of being saved: 164 MB at
//                               config 2); one CTA per frame, any power-of-two n_fft, both through the Stockham
//                               FFT of fft_stockham.cuh.  The windowed frame gradients go to a workspace
//                               (frame-major, n_fft floats per frame);
//   overlap_add_kernel          sums, for every waveform sample, the (up to n_fft / hop) frame entries that cover its
//                               padded position plus those of the padded positions the forward pass filled from it
//                               (adjoint of the reflect / replicate / circular / constant centre padding).  No atomics:
//                               the gradient is deterministic, and a float atomic per frame sample (2048 per frame)
//                               had cost more than the two FFTs.
//   filterbank_backward_kernel  d loss / d spec[k, t] = sum_m d loss / d y[m, t] * fb[k, m]
//   pointwise                   amplitude_to_db and complex_norm
//
// Maths of the first: X_k = sum_n x_n w_n e^{-i theta}, theta = 2 pi k n / N, k = 0 .. N/2 (onesided).  With
// G_k = dL/dRe X_k + i dL/dIm X_k:  dL/d(x_n w_n) = Re sum_k G_k e^{+i theta}.  Put H_0 = Re G_0, H_{N/2} = Re G_{N/2},
// H_k = G_k / 2 otherwise, extend Hermitian: the sum is the inverse real FFT y = sum_{k < N} H_k e^{+i theta}, evaluated
// as one complex inverse FFT of size C = N/2:  Zt_k = (H_k + conj H_{C-k}) + i e^{+2 pi i k / N} (H_k - conj H_{C-k}),
// z = IFFT_C(Zt), y_{2n} = Re z_n, y_{2n+1} = Im z_n; and IFFT(v) = conj(FFT(conj v)).
// For |X|^p outputs G_k = g_k p |X_k|^(p-2) X_k (0 where X_k = 0); two-sided outputs fold bin N-k onto k first.
#include <math.h>

#include "fft_stockham.cuh"
#include "stft_params.cuh"
#include "tac_common.cuh"

namespace tac {

constexpr int kBwdThreads = 256;

struct StftBwdParams {
  StftParams f;            // geometry of the forward call (x may be null in complex mode)
  const float* grad_out;   // complex mode: (n_seq, bins, frames, 2); power mode: (n_seq, bins, frames); contiguous
  float* frames_out;       // workspace: (n_seq * frames, n_fft) windowed frame gradients
  int power_mode;          // -1: gradient of the complex spectrum; 2 / 1 / 0 as in the forward kernels
  float power;
};

static size_t bwd_smem_bytes(int n_fft) {
  const size_t c = (size_t)n_fft / 2;
  return sizeof(float2) * (2 * c + (c + 1) + c / 2 + (c + 1)) + sizeof(float) * (size_t)n_fft + 16;
}

__global__ void __launch_bounds__(kBwdThreads) stft_backward_kernel(const StftBwdParams bp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const StftParams& p = bp.f;
  const int n_fft = p.n_fft, C = n_fft >> 1;
  float2* buf_a = reinterpret_cast<float2*>(smem_raw);
  float2* buf_b = buf_a + C;
  float2* spec = buf_b + C;                // C + 1 bins: X, then H
  float2* tw_c = spec + C + 1;             // W_C^m, m < C/2
  float2* tw_n = tw_c + (C >> 1);          // W_N^k, k <= C
  float* win = reinterpret_cast<float*>(tw_n + C + 1);      // window * scale (no 1/2: see the untangling below)
  const int tid = threadIdx.x;

  for (int i = tid; i < n_fft; i += kBwdThreads) win[i] = p.window[i] * p.scale;
  for (int i = tid; i < (C >> 1); i += kBwdThreads) {
    float sn, cs;
    sincospif(-2.0f * (float)i / (float)C, &sn, &cs);
    tw_c[i] = make_float2(cs, sn);
  }
  for (int i = tid; i <= C; i += kBwdThreads) {
    float sn, cs;
    sincospif(-2.0f * (float)i / (float)n_fft, &sn, &cs);
    tw_n[i] = make_float2(cs, sn);
  }
  __syncthreads();

  const int64_t plane = p.frames;                                      // elements between consecutive bins of grad_out
  for (int64_t g = p.g0 + blockIdx.x; g < p.g1; g += gridDim.x) {
    const int64_t seq = g / p.frames, t = g - seq * p.frames, start = t * p.hop - p.pad;
    const int64_t gbase = seq * p.bins * plane + t;                   // grad_out[seq, 0, t]

    if (bp.power_mode >= 0) {
      // ---- recompute X_k of this frame (same arithmetic as stft_generic_kernel) ----
      const float* row = p.x + seq * p.seq_stride;
      for (int n = tid; n < C; n += kBwdThreads) {
        const float x0 = fetch_padded(row, start + 2 * n, p.n_samples, p.pad_mode);
        const float x1 = fetch_padded(row, start + 2 * n + 1, p.n_samples, p.pad_mode);
        buf_a[n] = make_float2(0.5f * x0 * win[2 * n], 0.5f * x1 * win[2 * n + 1]);
      }
      __syncthreads();
      const float2* z_fft = stockham_forward<kBwdThreads>(buf_a, buf_b, tw_c, C, tid);
      for (int k = tid; k <= C; k += kBwdThreads) {
        const float2 z = z_fft[k & (C - 1)];
        const float2 q = z_fft[(C - k) & (C - 1)];
        const float a = z.x + q.x, b = z.y - q.y, gs = z.y + q.y, h = q.x - z.x;
        const float2 w = tw_n[k];
        const float xr = fmaf(w.x, gs, fmaf(-w.y, h, a));
        float xi = fmaf(w.x, h, fmaf(w.y, gs, b));
        if (k == 0 || k == C) xi = 0.0f;
        // G_k = g_k p |X|^(p-2) X  (two-sided: bin N - k holds conj X_k, its gradient folds onto k)
        float gk = __ldg(bp.grad_out + gbase + (int64_t)k * plane);
        if (!p.onesided && k > 0 && k < C) gk += __ldg(bp.grad_out + gbase + (int64_t)(n_fft - k) * plane);
        float coef;
        if (bp.power_mode == 2) {
          coef = 2.0f * gk;
        } else {
          const float n2 = xr * xr + xi * xi;
          if (n2 > 0.0f) coef = (bp.power_mode == 1) ? gk * rsqrtf(n2) : gk * bp.power * powf(n2, 0.5f * bp.power - 1.0f);
          else coef = 0.0f;
        }
        spec[k] = make_float2(coef * xr, coef * xi);
      }
    } else {
      const float2* go = reinterpret_cast<const float2*>(bp.grad_out);
      for (int k = tid; k <= C; k += kBwdThreads) {
        float2 gk = __ldg(go + gbase + (int64_t)k * plane);
        if (!p.onesided && k > 0 && k < C) {                         // Re(G_k e^{i th} + G_{N-k} e^{-i th}) = Re((G_k + conj G_{N-k}) e^{i th})
          const float2 gm = __ldg(go + gbase + (int64_t)(n_fft - k) * plane);
          gk.x += gm.x;
          gk.y -= gm.y;
        }
        spec[k] = gk;
      }
    }
    __syncthreads();
    // ---- H_k, then conj(Zt_k) into buf_a ----
    for (int k = tid; k < C; k += kBwdThreads) {
      float2 hk = spec[k], hc = spec[C - k];
      if (k == 0) {
        hk = make_float2(hk.x, 0.0f);                                // H_0 = Re G_0, H_C = Re G_C
        hc = make_float2(hc.x, 0.0f);
      } else {
        hk = make_float2(0.5f * hk.x, 0.5f * hk.y);
        hc = make_float2(0.5f * hc.x, 0.5f * hc.y);
      }
      // s = H_k + conj H_{C-k},  d = H_k - conj H_{C-k},  Zt = s + i conj(W_N^k) d   (tw_n holds W_N^k = e^{-2 pi i k / N})
      const float sx = hk.x + hc.x, sy = hk.y - hc.y, dx = hk.x - hc.x, dy = hk.y + hc.y;
      const float2 w = tw_n[k];
      const float ex = w.x * dx + w.y * dy, ey = w.x * dy - w.y * dx;            // conj(w) * d
      buf_a[k] = make_float2(sx - ey, -(sy + ex));                              // conj(s + i e)
    }
    __syncthreads();
    const float2* r = stockham_forward<kBwdThreads>(buf_a, buf_b, tw_c, C, tid);   // z_n = conj(r_n)
    float2* frow = reinterpret_cast<float2*>(bp.frames_out + (g - p.g0) * n_fft);
    for (int n = tid; n < C; n += kBwdThreads) {
      const float2 v = r[n];
      frow[n] = make_float2(v.x * win[2 * n], -v.y * win[2 * n + 1]);
    }
    __syncthreads();
  }
}

// ---- any n_fft that is not a power of two: the adjoint of stft_dft_kernel (stft.cu), direct sums -------------------
// X_k = sum_n xw[n] (c_nk + i s_nk), (c, s) = (cos, -sin)(2 pi n k / N): with (Gr_k, Gi_k) the gradient w.r.t. (Re X_k, Im X_k)
// of the bins that exist (two-sided outputs folded onto k <= N/2 as in stft_backward_kernel),
//     dL/d xw[n] = sum_{k=0}^{N/2} Gr_k c_nk + Gi_k s_nk.
// |X|^p outputs: the spectrum is recomputed by the same direct sum, G_k = g_k p |X_k|^(p-2) X_k.  One CTA per frame,
// O(N^2) both ways; the table index (n k) mod N is kept incrementally (exact).  Correct-first like the forward kernel.
static size_t dft_bwd_smem_bytes(int n_fft) {
  return sizeof(float2) * (size_t)((n_fft + 1) & ~1) + sizeof(float2) * (size_t)(n_fft / 2 + 2) + sizeof(float) * 2 * (size_t)n_fft + 32;
}

__global__ void __launch_bounds__(kBwdThreads) stft_dft_backward_kernel(const StftBwdParams bp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const StftParams& p = bp.f;
  const int N = p.n_fft, nb = N / 2 + 1;
  float2* tab = reinterpret_cast<float2*>(smem_raw);                   // [N] (cos, -sin)(2 pi j / N)
  float2* spec = tab + ((N + 1) & ~1);                                 // [nb] G_k
  float* win = reinterpret_cast<float*>(spec + nb + 1);                // [N] window * scale
  float* xw = win + N;                                                 // [N] windowed samples (power modes)
  const int tid = threadIdx.x;
  for (int j = tid; j < N; j += kBwdThreads) {
    double sn, cs;
    sincospi(-2.0 * (double)j / (double)N, &sn, &cs);
    tab[j] = make_float2((float)cs, (float)sn);
    win[j] = p.window[j] * p.scale;
  }
  __syncthreads();
  const int64_t plane = p.frames;
  for (int64_t g = p.g0 + blockIdx.x; g < p.g1; g += gridDim.x) {
    const int64_t seq = g / p.frames, t = g - seq * p.frames, start = t * p.hop - p.pad;
    const int64_t gbase = seq * p.bins * plane + t;
    if (bp.power_mode >= 0) {
      const float* row = p.x + seq * p.seq_stride;
      for (int n = tid; n < N; n += kBwdThreads) xw[n] = fetch_padded(row, start + n, p.n_samples, p.pad_mode) * win[n];
      __syncthreads();
      for (int k = tid; k < nb; k += kBwdThreads) {
        float xr = 0.0f, xi = 0.0f;
        int idx = 0;
        for (int n = 0; n < N; ++n) {
          const float2 w = tab[idx];
          xr = fmaf(xw[n], w.x, xr);
          xi = fmaf(xw[n], w.y, xi);
          idx += k;
          idx -= (idx >= N) ? N : 0;
        }
        if (k == 0 || 2 * k == N) xi = 0.0f;
        float gk = __ldg(bp.grad_out + gbase + (int64_t)k * plane);
        if (!p.onesided && k > 0 && 2 * k != N) gk += __ldg(bp.grad_out + gbase + (int64_t)(N - k) * plane);
        float coef;
        if (bp.power_mode == 2) {
          coef = 2.0f * gk;
        } else {
          const float n2 = xr * xr + xi * xi;
          if (n2 > 0.0f) coef = (bp.power_mode == 1) ? gk * rsqrtf(n2) : gk * bp.power * powf(n2, 0.5f * bp.power - 1.0f);
          else coef = 0.0f;
        }
        spec[k] = make_float2(coef * xr, coef * xi);
      }
    } else {
      const float2* go = reinterpret_cast<const float2*>(bp.grad_out);
      for (int k = tid; k < nb; k += kBwdThreads) {
        float2 gk = __ldg(go + gbase + (int64_t)k * plane);
        if (!p.onesided && k > 0 && 2 * k != N) {
          const float2 gm = __ldg(go + gbase + (int64_t)(N - k) * plane);
          gk.x += gm.x;
          gk.y -= gm.y;
        }
        spec[k] = gk;
      }
    }
    __syncthreads();
    float* frow = bp.frames_out + (g - p.g0) * N;
    for (int n = tid; n < N; n += kBwdThreads) {
      float acc = 0.0f;
      int idx = 0;
      for (int k = 0; k < nb; ++k) {
        const float2 w = tab[idx];
        const float2 gk = spec[k];
        acc = fmaf(gk.x, w.x, acc);
        acc = fmaf(gk.y, w.y, acc);
        idx += n;
        idx -= (idx >= N) ? N : 0;
      }
      frow[n] = acc * win[n];
    }
    __syncthreads();
  }
}

// P(q) = sum over the frames t covering padded position q of frames_out[t][q - t hop]
__device__ __forceinline__ float ola_padded(const float* __restrict__ fr, int q, int frames, int n_fft, int hop) {
  int t_lo = q - n_fft + 1;
  t_lo = t_lo <= 0 ? 0 : (t_lo + hop - 1) / hop;
  int t_hi = q / hop;
  t_hi = t_hi > frames - 1 ? frames - 1 : t_hi;
  float acc = 0.0f;
  for (int t = t_lo; t <= t_hi; ++t) acc += __ldg(fr + (int64_t)t * n_fft + (q - t * hop));
  return acc;
}

// the padded positions that mirror / wrap onto waveform sample j (reflect / circular; replicate edges: see below)
__device__ __forceinline__ float ola_folded(const float* __restrict__ fr, int j, int n_samples, int frames, int n_fft, int hop,
                                            int pad, int pad_mode) {
  float acc = 0.0f;
  if (pad_mode == 0) {                                             // reflect: x[-k] = x[k], x[T-1+k] = x[T-1-k]
    if (j >= 1 && j <= pad) acc += ola_padded(fr, pad - j, frames, n_fft, hop);
    if (j >= n_samples - 1 - pad && j <= n_samples - 2) acc += ola_padded(fr, pad + 2 * (n_samples - 1) - j, frames, n_fft, hop);
  } else if (pad_mode == 3) {                                      // circular
    if (j >= n_samples - pad) acc += ola_padded(fr, j + pad - n_samples, frames, n_fft, hop);
    if (j < pad) acc += ola_padded(fr, j + pad + n_samples, frames, n_fft, hop);
  }
  return acc;
}

// grad_x[seq][j] = P(j + pad) + folded positions.  VEC: four consecutive samples per thread with 16-byte loads (hop, pad,
// n_fft and the row length multiples of 4: the four samples then sit in the same frames at the same offsets).
template <bool VEC>
__global__ void __launch_bounds__(256)
overlap_add_kernel(const float* __restrict__ frames_ws, int64_t n_seq, int n_samples, int frames, int n_fft, int hop, int pad,
                   int pad_mode, float* __restrict__ grad_x) {
  constexpr int kPer = VEC ? 4 : 1;
  const int j0 = (blockIdx.x * 256 + threadIdx.x) * kPer;
  if (j0 >= n_samples) return;
  for (int64_t seq = blockIdx.y; seq < n_seq; seq += gridDim.y) {
    const float* fr = frames_ws + seq * frames * (int64_t)n_fft;
    float* out = grad_x + seq * n_samples + j0;
    if constexpr (VEC) {
      const int q = j0 + pad;
      int t_lo = q + 3 - n_fft + 1;
      t_lo = t_lo <= 0 ? 0 : (t_lo + hop - 1) / hop;
      int t_hi = q / hop;
      t_hi = t_hi > frames - 1 ? frames - 1 : t_hi;
      float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      for (int t = t_lo; t <= t_hi; ++t) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(fr + (int64_t)t * n_fft + (q - t * hop)));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      if (pad > 0 && (j0 <= pad || j0 + 3 >= n_samples - 1 - pad)) {
        acc.x += ola_folded(fr, j0, n_samples, frames, n_fft, hop, pad, pad_mode);
        acc.y += ola_folded(fr, j0 + 1, n_samples, frames, n_fft, hop, pad, pad_mode);
        acc.z += ola_folded(fr, j0 + 2, n_samples, frames, n_fft, hop, pad, pad_mode);
        acc.w += ola_folded(fr, j0 + 3, n_samples, frames, n_fft, hop, pad, pad_mode);
      }
      *reinterpret_cast<float4*>(out) = acc;
    } else {
      float acc = ola_padded(fr, j0 + pad, frames, n_fft, hop);
      if (pad > 0) acc += ola_folded(fr, j0, n_samples, frames, n_fft, hop, pad, pad_mode);
      *out = acc;
    }
  }
}
// replicate padding: sample 0 also receives padded positions [0, pad), sample T - 1 positions (pad + T - 1, T + 2 pad)
__global__ void __launch_bounds__(256)
overlap_add_edges_kernel(const float* __restrict__ frames_ws, int n_samples, int frames, int n_fft, int hop, int pad,
                         float* __restrict__ grad_x) {
  const int64_t seq = blockIdx.x >> 1;
  const bool right = blockIdx.x & 1;
  const float* fr = frames_ws + seq * frames * (int64_t)n_fft;
  float acc = 0.0f;
  for (int q = threadIdx.x; q < pad; q += blockDim.x) acc += ola_padded(fr, right ? pad + n_samples + q : q, frames, n_fft, hop);
  __shared__ float part[8];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float sum = 0.0f;
    for (int w = 0; w < 8; ++w) sum += part[w];
    grad_x[seq * n_samples + (right ? n_samples - 1 : 0)] += sum;
  }
}

static int64_t backward_workspace_bytes(const StftParams& p) { return p.n_seq * p.frames * (int64_t)p.n_fft * 4; }

static int launch_overlap_add(const StftBwdParams& bp, float* grad_x, cudaStream_t stream) {
  const StftParams& p = bp.f;
  const bool vec = (p.hop & 3) == 0 && (p.pad & 3) == 0 && (p.n_fft & 3) == 0 && (p.n_samples & 3) == 0 &&
                   (reinterpret_cast<uintptr_t>(bp.frames_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad_x) & 15) == 0;
  const int per_block = 256 * (vec ? 4 : 1);
  dim3 grid((unsigned)((p.n_samples + per_block - 1) / per_block), (unsigned)(p.n_seq < 32768 ? p.n_seq : 32768), 1);
  {
    LaunchProbe probe(KIND_POINTWISE, stream);
    if (vec)
      overlap_add_kernel<true><<<grid, 256, 0, stream>>>(bp.frames_out, p.n_seq, (int)p.n_samples, (int)p.frames, p.n_fft, p.hop,
                                                          p.pad, p.pad_mode, grad_x);
    else
      overlap_add_kernel<false><<<grid, 256, 0, stream>>>(bp.frames_out, p.n_seq, (int)p.n_samples, (int)p.frames, p.n_fft, p.hop,
                                                           p.pad, p.pad_mode, grad_x);
  }
  TAC_CUDA_OK(cudaGetLastError());
  if (p.pad > 0 && p.pad_mode == TAC_PAD_REPLICATE) {
    LaunchProbe probe(KIND_POINTWISE, stream);
    overlap_add_edges_kernel<<<(int)(2 * p.n_seq), 256, 0, stream>>>(bp.frames_out, (int)p.n_samples, (int)p.frames, p.n_fft,
                                                                    p.hop, p.pad, grad_x);
    TAC_CUDA_OK(cudaGetLastError());
  }
  return TAC_OK;
}

static int launch_filterbank_backward(const float* grad_y, int64_t stride_seq, int64_t stride_band, int64_t stride_frame,
                                      const float* fb_dev, int64_t n_seq, int64_t frames, int n_bins, int n_bands, float* grad_spec,
                                      int frame_major_kpad, void* table_scratch, cudaStream_t stream);

static int launch_stft_backward(StftBwdParams& bp, float* grad_x, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  const StftParams& p = bp.f;
  if (p.n_seq == 0 || p.n_samples == 0) return TAC_OK;
  const int64_t n_frames = p.g1 - p.g0;
  if (n_frames <= 0) {                                   // no frame covers anything: the gradient is zero
    TAC_CUDA_OK(cudaMemsetAsync(grad_x, 0, (size_t)p.n_seq * p.n_samples * sizeof(float), stream));
    return TAC_OK;
  }
  TAC_REQUIRE(workspace && workspace_bytes >= backward_workspace_bytes(p), TAC_ERR_WORKSPACE,
              "stft backward: workspace of %lld bytes, %lld needed (tac_stft_backward_workspace_bytes)", (long long)workspace_bytes,
              (long long)backward_workspace_bytes(p));
  TAC_REQUIRE(p.n_seq * 2 < ((int64_t)1 << 31), TAC_ERR_UNSUPPORTED, "stft backward: too many sequences in one call");
  bp.frames_out = static_cast<float*>(workspace);
  const bool dft = !is_pow2(p.n_fft) || p.n_fft < 32;                 // sizes the Stockham kernel does not cover
  const size_t smem = dft ? dft_bwd_smem_bytes(p.n_fft) : bwd_smem_bytes(p.n_fft);
  if (dft) TAC_CUDA_OK(cudaFuncSetAttribute(stft_dft_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else TAC_CUDA_OK(cudaFuncSetAttribute(stft_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_sm = (int)((200 * 1024) / (smem + 1024));
  const int64_t cap = (int64_t)sm_count() * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
  const int grid = (int)(n_frames < cap ? n_frames : cap);
  {
    LaunchProbe probe(KIND_STFT, stream);
    if (dft) stft_dft_backward_kernel<<<grid, kBwdThreads, smem, stream>>>(bp);
    else stft_backward_kernel<<<grid, kBwdThreads, smem, stream>>>(bp);
  }
  TAC_CUDA_OK(cudaGetLastError());
  return launch_overlap_add(bp, grad_x, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// filterbank: grad_spec[s, k, t] = sum_m grad_y[s, m, t] * fb[k, m]        (adjoint of functional.py:183-184)
// One CTA per (sequence, 64 frames): the grad_y tile (n_bands x 64) sits in shared memory, thread (ty, tx) walks bins
// ty, ty + 4, ... for frame tx; fb[k, :] is read as a warp-wide broadcast and rows of zeros (outside the band range
// of bin k) are skipped through the per-bin [lo, hi) range computed once per CTA.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kFbTileT = 64;
constexpr int kFbThreads = 256;

__global__ void __launch_bounds__(kFbThreads)
filterbank_backward_kernel(const float* __restrict__ grad_y, int64_t sy_seq, int64_t sy_band, int64_t sy_frame,
                           const float* __restrict__ fb, int n_bins, int n_bands, int64_t frames, int tiles_per_seq,
                           int64_t n_jobs, float* __restrict__ grad_spec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);                       // [n_bands][kFbTileT + 1]
  int2* range = reinterpret_cast<int2*>(tile + (size_t)n_bands * (kFbTileT + 1));   // [n_bins]: non-zero columns [lo, hi)
  const int tid = threadIdx.x;
  // per-bin range of non-zero columns: a warp per bin, coalesced row reads, warp min / max
  // (a thread per bin walking its row made this prologue as expensive as the contraction itself)
  for (int k = tid >> 5; k < n_bins; k += kFbThreads / 32) {
    int lo = n_bands, hi = 0;
    for (int m = tid & 31; m < n_bands; m += 32)
      if (__ldg(fb + (int64_t)k * n_bands + m) != 0.0f) {
        lo = m < lo ? m : lo;
        hi = m + 1;
      }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((tid & 31) == 0) range[k] = make_int2(lo, hi);
  }
  for (int64_t job = blockIdx.x; job < n_jobs; job += gridDim.x) {
    const int64_t seq = job / tiles_per_seq;
    const int64_t t0 = (job - seq * tiles_per_seq) * kFbTileT;
    __syncthreads();                                                      // previous tile consumed / ranges written
    const float* gy = grad_y + seq * sy_seq;
    if (sy_band == 1) {                                                   // frame-major gradient: bands are contiguous
      for (int i = tid; i < n_bands * kFbTileT; i += kFbThreads) {
        const int tt = i / n_bands, m = i - tt * n_bands;
        tile[m * (kFbTileT + 1) + tt] = (t0 + tt < frames) ? __ldg(gy + (t0 + tt) * sy_frame + m) : 0.0f;
      }
    } else {
      for (int i = tid; i < n_bands * kFbTileT; i += kFbThreads) {
        const int m = i / kFbTileT, tt = i - m * kFbTileT;
        tile[m * (kFbTileT + 1) + tt] = (t0 + tt < frames) ? __ldg(gy + (int64_t)m * sy_band + (t0 + tt) * sy_frame) : 0.0f;
      }
    }
    __syncthreads();
    const int tx = tid & (kFbTileT - 1), ty = tid / kFbTileT;
    const bool live = t0 + tx < frames;
    float* out = grad_spec + seq * (int64_t)n_bins * frames + t0 + tx;
    for (int k = ty; k < n_bins; k += kFbThreads / kFbTileT) {
      const int2 rg = range[k];
      const float* w = fb + (int64_t)k * n_bands;
      float acc = 0.0f;
      for (int m = rg.x; m < rg.y; ++m) acc = fmaf(__ldg(w + m), tile[m * (kFbTileT + 1) + tx], acc);
      if (live) out[(int64_t)k * frames] = acc;
    }
  }
}

// The same contraction with a FRAME-MAJOR result, grad_spec[(s frames + t) kpad + k] (kpad = 1056 for 1025 bins,
// padding bins zero): the row layout stft2048_backward_kernel bulk-copies.  A warp takes a frame of the tile, lanes walk
// the bins, so the stores are coalesced; each bin's non-zero weights (at most four, true for every triangular
// filterbank) sit in a shared-memory table built once per CTA; wider rows fall back to reading the matrix.
// per-bin table of the matrix: non-zero column range and the first four weights of it.  One warp per bin, all bins in
// parallel (built by every CTA of the contraction kernel, one bin after the other, it cost 80 us of load latency).
__global__ void __launch_bounds__(kFbThreads)
filterbank_table_kernel(const float* __restrict__ fb, int n_bins, int n_bands, int kpad, float4* __restrict__ w_out,
                        int2* __restrict__ rg_out) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (kFbThreads / 32) + (threadIdx.x >> 5);
  if (k >= kpad) return;
  int lo = n_bands, hi = 0;
  if (k < n_bins)
    for (int m = lane; m < n_bands; m += 32)
      if (__ldg(fb + (int64_t)k * n_bands + m) != 0.0f) {
        lo = m < lo ? m : lo;
        hi = m + 1;
      }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if (hi <= lo) lo = hi = 0;
  float wv = 0.0f;
  if (lane < 4 && lo + lane < hi) wv = __ldg(fb + (int64_t)k * n_bands + lo + lane);
  const float w0 = __shfl_sync(0xffffffffu, wv, 0), w1 = __shfl_sync(0xffffffffu, wv, 1);
  const float w2 = __shfl_sync(0xffffffffu, wv, 2), w3 = __shfl_sync(0xffffffffu, wv, 3);
  if (lane == 0) {
    w_out[k] = make_float4(w0, w1, w2, w3);
    rg_out[k] = make_int2(lo, hi);
  }
}

__global__ void __launch_bounds__(kFbThreads)
filterbank_backward_fm_kernel(const float* __restrict__ grad_y, int64_t sy_seq, int64_t sy_band, int64_t sy_frame,
                              const float* __restrict__ fb, int n_bins, int n_bands, int64_t frames, int tiles_per_seq,
                              int64_t n_jobs, int kpad, const float4* __restrict__ g_w, const int2* __restrict__ g_rg,
                              float* __restrict__ grad_spec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_w = reinterpret_cast<float4*>(smem_raw);                      // [kpad] weights of columns lo .. lo + 3
  int2* s_rg = reinterpret_cast<int2*>(s_w + kpad);                       // [kpad] (lo, hi)
  float* tile = reinterpret_cast<float*>(s_rg + kpad);                    // [n_bands + 3][kFbTileT + 1], last 3 rows zero
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kRow = kFbTileT + 1;
  for (int k = tid; k < kpad; k += kFbThreads) {
    s_w[k] = __ldg(g_w + k);
    s_rg[k] = __ldg(g_rg + k);
  }
  for (int i = tid; i < 3 * kRow; i += kFbThreads) tile[n_bands * kRow + i] = 0.0f;
  for (int64_t job = blockIdx.x; job < n_jobs; job += gridDim.x) {
    const int64_t seq = job / tiles_per_seq;
    const int64_t t0 = (job - seq * tiles_per_seq) * kFbTileT;
    __syncthreads();
    const float* gy = grad_y + seq * sy_seq;
    if (sy_band == 1) {
      for (int i = tid; i < n_bands * kFbTileT; i += kFbThreads) {
        const int tt = i / n_bands, m = i - tt * n_bands;
        tile[m * kRow + tt] = (t0 + tt < frames) ? __ldg(gy + (t0 + tt) * sy_frame + m) : 0.0f;
      }
    } else {
      for (int i = tid; i < n_bands * kFbTileT; i += kFbThreads) {
        const int m = i / kFbTileT, tt = i - m * kFbTileT;
        tile[m * kRow + tt] = (t0 + tt < frames) ? __ldg(gy + (int64_t)m * sy_band + (t0 + tt) * sy_frame) : 0.0f;
      }
    }
    __syncthreads();
    for (int tt = warp; tt < kFbTileT && t0 + tt < frames; tt += kFbThreads / 32) {
      float* out = grad_spec + (seq * frames + t0 + tt) * (int64_t)kpad;
      for (int k = lane; k < kpad; k += 32) {
        const int2 rg = s_rg[k];
        float acc;
        if (rg.y - rg.x <= 4) {
          const float4 w = s_w[k];
          const float* col = tile + rg.x * kRow + tt;
          acc = fmaf(w.w, col[3 * kRow], fmaf(w.z, col[2 * kRow], fmaf(w.y, col[kRow], w.x * col[0])));
        } else {
          acc = 0.0f;
          for (int m = rg.x; m < rg.y; ++m) acc = fmaf(__ldg(fb + (int64_t)k * n_bands + m), tile[m * kRow + tt], acc);
        }
        out[k] = acc;
      }
    }
  }
}

// (n_seq, n_bins, frames) -> frame-major (n_seq * frames, kpad) with zero padding bins: what stft2048_backward_kernel
// bulk-copies, for a gradient that arrives in the reference's public Spectrogram layout.  32 x 32 tiles through shared memory.
__global__ void __launch_bounds__(256)
spec_to_frame_major_kernel(const float* __restrict__ src, int n_bins, int64_t frames, int kpad, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int64_t seq = blockIdx.z;
  const int k0 = blockIdx.y * 32;
  const int64_t t0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;                  // 32 x 8
  const float* s = src + seq * n_bins * frames;
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r;
    tile[r][tx] = (k < n_bins && t0 + tx < frames) ? __ldg(s + (int64_t)k * frames + t0 + tx) : 0.0f;
  }
  __syncthreads();
  float* d = dst + seq * frames * kpad;
  for (int r = ty; r < 32; r += 8) {
    const int64_t t = t0 + r;
    if (t < frames && k0 + tx < kpad) d[t * kpad + k0 + tx] = tile[tx][r];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// pointwise backward
// ---------------------------------------------------------------------------------------------------------------
// amplitude_to_db (functional.py:291-296): y = 10 log10(max(x^2, amin)) - c  ->  dy/dx = 20 / (ln 10 x) where x^2 >= amin
__global__ void amplitude_to_db_backward_kernel(const float* __restrict__ x, const float* __restrict__ g, int64_t n, float amin,
                                                float* __restrict__ out) {
  const float k = 8.685889638065035f;                                   // 20 / ln 10
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    out[i] = (v * v >= amin) ? g[i] * k / v : 0.0f;
  }
}
// complex_norm (functional.py:126-128): n = |z|, y = n^p  ->  dz = g p n^(p-2) z (0 at z = 0)
__global__ void complex_norm_backward_kernel(const float2* __restrict__ z, const float* __restrict__ g, int64_t n, float power,
                                             float2* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 v = z[i];
    const float n2 = v.x * v.x + v.y * v.y;
    float c = 0.0f;
    if (n2 > 0.0f) c = (power == 2.0f) ? 2.0f * g[i] : ((power == 1.0f) ? g[i] * rsqrtf(n2) : g[i] * power * powf(n2, 0.5f * power - 1.0f));
    out[i] = make_float2(c * v.x, c * v.y);
  }
}

static int pointwise_grid(int64_t n) {
  const int64_t want = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace tac

namespace tac {

// ---- gradient w.r.t. the window (round 2) ---------------------------------------------------------------------------
// X = FFT(x_frame * w * scale): dL/dw[n] = sum over frames of x_padded[start + n] * d[n], d = the frame gradient BEFORE the
// window multiply.  The caller runs tac_stft_backward_f32 with a window of ones, which leaves scale * d in the frame
// workspace; this reduces it against the padded waveform.  Two deterministic steps: partial sums over slices of the
// frames (one thread per window sample, coalesced along n), then the slices in fixed order.
constexpr int kWgSlices = 64;
__global__ void __launch_bounds__(256) window_grad_partial_kernel(const float* __restrict__ x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                                                                  const float* __restrict__ frames_ws, int64_t frames, int n_fft, int hop, int pad,
                                                                  int pad_mode, float* __restrict__ partial) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= n_fft) return;
  const int64_t total = n_seq * frames;
  const int64_t per = (total + kWgSlices - 1) / kWgSlices;
  const int64_t g0 = (int64_t)blockIdx.y * per, g1 = (g0 + per < total) ? g0 + per : total;
  float acc = 0.0f;
  for (int64_t g = g0; g < g1; ++g) {
    const int64_t seq = g / frames, t = g - seq * frames;
    const float xv = fetch_padded(x + seq * seq_stride, t * hop - pad + n, n_samples, pad_mode);
    acc = fmaf(xv, frames_ws[g * n_fft + n], acc);
  }
  partial[(int64_t)blockIdx.y * n_fft + n] = acc;
}
__global__ void __launch_bounds__(256) window_grad_final_kernel(const float* __restrict__ partial, int n_fft, float* __restrict__ grad_window) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= n_fft) return;
  float acc = 0.0f;
  for (int s = 0; s < kWgSlices; ++s) acc += partial[(int64_t)s * n_fft + n];
  grad_window[n] = acc;
}

}  // namespace tac

extern "C" int tac_window_grad_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride, const float* frames_ws, int n_fft,
                                   int hop, int center, int pad_mode, float* grad_window, void* scratch, int64_t scratch_bytes, void* stream) {
  using namespace tac;
  TAC_REQUIRE(grad_window && n_fft >= 2 && hop >= 1 && n_seq >= 0, TAC_ERR_INVALID, "window_grad: bad arguments");
  const int64_t frames = tac_stft_num_frames(n_samples, n_fft, hop, center);
  const int blocks = (n_fft + 255) / 256;
  if (n_seq * frames <= 0) {
    TAC_CUDA_OK(cudaMemsetAsync(grad_window, 0, sizeof(float) * (size_t)n_fft, as_stream(stream)));
    return TAC_OK;
  }
  TAC_REQUIRE(x && frames_ws, TAC_ERR_INVALID, "window_grad: null pointer");
  TAC_REQUIRE(scratch && scratch_bytes >= (int64_t)kWgSlices * n_fft * 4, TAC_ERR_WORKSPACE, "window_grad: scratch of %lld bytes needed",
              (long long)kWgSlices * n_fft * 4);
  LaunchProbe probe(KIND_STFT, as_stream(stream));
  window_grad_partial_kernel<<<dim3(blocks, kWgSlices), 256, 0, as_stream(stream)>>>(x, n_seq, n_samples, seq_stride, frames_ws, frames, n_fft, hop,
                                                                                   center ? n_fft / 2 : 0, pad_mode, static_cast<float*>(scratch));
  window_grad_final_kernel<<<blocks, 256, 0, as_stream(stream)>>>(static_cast<const float*>(scratch), n_fft, grad_window);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int64_t tac_stft_backward_workspace_bytes(int64_t n_seq, int64_t n_samples, int n_fft, int hop, int center) {
  if (n_fft <= 0 || hop <= 0 || n_seq <= 0) return 0;
  const int64_t rows = n_seq * tac_stft_num_frames(n_samples, n_fft, hop, center);
  // frame gradients; n_fft = 2048 adds the frame-major copy of a Spectrogram gradient (1056 floats per frame)
  return rows * (int64_t)n_fft * 4 + (n_fft == 2048 ? rows * 1056 * 4 : 0);
}

extern "C" int tac_stft_backward_f32(const float* grad_out, int64_t n_seq, int64_t n_samples, const float* window, int n_fft,
                                     int hop, int center, int pad_mode, int normalized, int onesided, float* grad_x,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace tac;
  StftBwdParams bp;
  TAC_REQUIRE((grad_out && grad_x) || n_seq == 0, TAC_ERR_INVALID, "stft_backward: null gradient pointer");
  const int rc = fill_stft_params(bp.f, grad_x /* placeholder, never read */, n_seq, n_samples, n_samples, window, n_fft, hop,
                                  center, pad_mode, normalized, onesided);
  if (rc != TAC_OK) return rc;
  bp.f.x = nullptr;
  bp.grad_out = grad_out;
  bp.power_mode = -1;
  bp.power = 1.0f;
  return launch_stft_backward(bp, grad_x, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int tac_spectrogram_backward_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                                            const float* window, int n_fft, int hop, int center, int pad_mode, int normalized,
                                            int onesided, float power, const float* grad_out, float* grad_x, void* workspace,
                                            int64_t workspace_bytes, void* stream) {
  using namespace tac;
  StftBwdParams bp;
  TAC_REQUIRE((grad_out && grad_x) || n_seq == 0, TAC_ERR_INVALID, "spectrogram_backward: null gradient pointer");
  const int rc = fill_stft_params(bp.f, x, n_seq, n_samples, seq_stride, window, n_fft, hop, center, pad_mode, normalized, onesided);
  if (rc != TAC_OK) return rc;
  bp.grad_out = grad_out;
  bp.power = power;
  bp.power_mode = power == 2.0f ? 2 : (power == 1.0f ? 1 : 0);
  StftParams& p = bp.f;
  const int64_t rows = p.n_seq * p.frames;
  if (n_fft == 2048 && onesided && rows > 0 && p.n_samples > 0 && p.n_seq < 65536) {
    // warp-per-frame kernel of stft.cu: it wants the gradient frame-major
    const int64_t frames_bytes = rows * 2048 * 4, need = frames_bytes + rows * 1056 * 4;
    TAC_REQUIRE(workspace && workspace_bytes >= need, TAC_ERR_WORKSPACE,
                "spectrogram_backward: workspace of %lld bytes, %lld needed (tac_stft_backward_workspace_bytes)",
                (long long)workspace_bytes, (long long)need);
    bp.frames_out = static_cast<float*>(workspace);
    float* gspec = bp.frames_out + rows * 2048;
    cudaStream_t st = as_stream(stream);
    {
      LaunchProbe probe(KIND_POINTWISE, st);
      dim3 grid((unsigned)((p.frames + 31) / 32), (unsigned)(1056 / 32), (unsigned)p.n_seq);
      spec_to_frame_major_kernel<<<grid, 256, 0, st>>>(grad_out, p.bins, p.frames, 1056, gspec);
    }
    TAC_CUDA_OK(cudaGetLastError());
    p.power = power;
    p.power_mode = bp.power_mode;
    const int rc2 = launch_stft2048_backward(p, gspec, bp.frames_out, st);
    if (rc2 != TAC_OK) return rc2;
    return launch_overlap_add(bp, grad_x, st);
  }
  return launch_stft_backward(bp, grad_x, workspace, workspace_bytes, as_stream(stream));
}

namespace tac {
static int launch_filterbank_backward(const float* grad_y, int64_t stride_seq, int64_t stride_band, int64_t stride_frame,
                                      const float* fb_dev, int64_t n_seq, int64_t frames, int n_bins, int n_bands, float* grad_spec,
                                      int frame_major_kpad, void* table_scratch, cudaStream_t stream) {
  TAC_REQUIRE(n_seq >= 0 && frames >= 0 && n_bins > 0 && n_bands > 0, TAC_ERR_INVALID, "filterbank_backward: bad shape");
  if (n_seq == 0 || frames == 0) return TAC_OK;
  TAC_REQUIRE(grad_y && fb_dev && grad_spec, TAC_ERR_INVALID, "filterbank_backward: null pointer");
  const int tiles_per_seq = (int)((frames + kFbTileT - 1) / kFbTileT);
  const int64_t jobs = n_seq * tiles_per_seq;
  const int64_t cap = (int64_t)sm_count() * 4;
  const int grid = (int)(jobs < cap ? jobs : cap);
  if (frame_major_kpad > 0) {
    const size_t smem = (sizeof(float4) + sizeof(int2)) * (size_t)frame_major_kpad + sizeof(float) * (size_t)(n_bands + 3) * (kFbTileT + 1);
    TAC_REQUIRE(smem <= 200 * 1024, TAC_ERR_UNSUPPORTED, "filterbank_backward: %d bands x %d bins exceed one CTA's shared memory",
                n_bands, n_bins);
    TAC_REQUIRE(table_scratch && (reinterpret_cast<uintptr_t>(table_scratch) & 15) == 0, TAC_ERR_INVALID,
                "filterbank_backward: missing scratch for the per-bin table");
    float4* g_w = static_cast<float4*>(table_scratch);                    // 24 bytes per bin, in the caller's workspace
    int2* g_rg = reinterpret_cast<int2*>(g_w + frame_major_kpad);
    TAC_CUDA_OK(cudaFuncSetAttribute(filterbank_backward_fm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
      LaunchProbe probe(KIND_MELBANK, stream);
      filterbank_table_kernel<<<(frame_major_kpad + 7) / 8, kFbThreads, 0, stream>>>(fb_dev, n_bins, n_bands, frame_major_kpad, g_w, g_rg);
    }
    LaunchProbe probe(KIND_MELBANK, stream);
    filterbank_backward_fm_kernel<<<grid, kFbThreads, smem, stream>>>(grad_y, stride_seq, stride_band, stride_frame, fb_dev, n_bins,
                                                                     n_bands, frames, tiles_per_seq, jobs, frame_major_kpad, g_w, g_rg,
                                                                     grad_spec);
  } else {
    LaunchProbe probe(KIND_MELBANK, stream);
    const size_t smem = sizeof(float) * (size_t)n_bands * (kFbTileT + 1) + sizeof(int2) * (size_t)n_bins;
    TAC_REQUIRE(smem <= 200 * 1024, TAC_ERR_UNSUPPORTED, "filterbank_backward: %d bands x %d bins exceed one CTA's shared memory",
                n_bands, n_bins);
    TAC_CUDA_OK(cudaFuncSetAttribute(filterbank_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    filterbank_backward_kernel<<<grid, kFbThreads, smem, stream>>>(grad_y, stride_seq, stride_band, stride_frame, fb_dev, n_bins,
                                                                  n_bands, frames, tiles_per_seq, jobs, grad_spec);
  }
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}
}  // namespace tac

extern "C" int tac_filterbank_backward_f32(const float* grad_y, int64_t stride_seq, int64_t stride_band, int64_t stride_frame,
                                           const float* fb_dev, int64_t n_seq, int64_t frames, int n_bins, int n_bands,
                                           float* grad_spec, void* stream) {
  return tac::launch_filterbank_backward(grad_y, stride_seq, stride_band, stride_frame, fb_dev, n_seq, frames, n_bins, n_bands,
                                         grad_spec, 0, nullptr, tac::as_stream(stream));
}

// Backward of the whole Melspectrogram chain (stft -> |.|^power -> filterbank), d loss / d waveform from d loss / d mel.
// Workspace = [filterbank-adjoint result | windowed frame gradients].  n_fft = 2048: frame-major filterbank adjoint ->
// stft2048_backward_kernel (one warp per frame, stft.cu) -> overlap-add; other sizes: the generic kernels above.
static int64_t melspec_backward_split(int64_t n_seq, int64_t frames, int n_fft, int n_bins, int64_t* spec_bytes) {
  const int64_t rows = n_seq * frames;
  int64_t a = (n_fft == 2048 ? rows * 1056 : rows * (int64_t)n_bins) * 4;
  a = (a + 255) & ~(int64_t)255;
  if (spec_bytes) *spec_bytes = a;
  return a + rows * (int64_t)n_fft * 4 + 32768;          // + the per-bin table of the frame-major filterbank adjoint
}

extern "C" int64_t tac_melspec_backward_workspace_bytes(int64_t n_seq, int64_t n_samples, int n_fft, int hop, int center) {
  if (n_fft <= 0 || hop <= 0 || n_seq <= 0) return 0;
  return melspec_backward_split(n_seq, tac_stft_num_frames(n_samples, n_fft, hop, center), n_fft, n_fft / 2 + 1, nullptr);
}

extern "C" int tac_melspec_backward_f32(const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride, const float* window,
                                        int n_fft, int hop, int center, int pad_mode, int normalized, float power,
                                        const float* fb_dev, int n_bands, const float* grad_y, int64_t stride_seq,
                                        int64_t stride_band, int64_t stride_frame, float* grad_x, void* workspace,
                                        int64_t workspace_bytes, void* stream) {
  using namespace tac;
  StftBwdParams bp;
  const int rc = fill_stft_params(bp.f, x, n_seq, n_samples, seq_stride, window, n_fft, hop, center, pad_mode, normalized, 1);
  if (rc != TAC_OK) return rc;
  StftParams& p = bp.f;
  if (p.n_seq == 0 || p.n_samples == 0) return TAC_OK;
  TAC_REQUIRE(grad_x, TAC_ERR_INVALID, "melspec_backward: null gradient pointer");
  cudaStream_t st = as_stream(stream);
  if (p.g1 <= 0) {
    TAC_CUDA_OK(cudaMemsetAsync(grad_x, 0, (size_t)p.n_seq * p.n_samples * sizeof(float), st));
    return TAC_OK;
  }
  int64_t spec_bytes = 0;
  const int64_t need = melspec_backward_split(p.n_seq, p.frames, n_fft, p.bins, &spec_bytes);
  TAC_REQUIRE(workspace && workspace_bytes >= need, TAC_ERR_WORKSPACE,
              "melspec_backward: workspace of %lld bytes, %lld needed (tac_melspec_backward_workspace_bytes)",
              (long long)workspace_bytes, (long long)need);
  float* gspec = static_cast<float*>(workspace);
  bp.frames_out = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + spec_bytes);
  bp.grad_out = gspec;
  bp.power = power;
  bp.power_mode = power == 2.0f ? 2 : (power == 1.0f ? 1 : 0);
  p.power = power;
  p.power_mode = bp.power_mode;
  const bool fast = n_fft == 2048;
  void* table = bp.frames_out + p.n_seq * p.frames * (int64_t)n_fft;     // the last 32 KB of the workspace
  int rc2 = launch_filterbank_backward(grad_y, stride_seq, stride_band, stride_frame, fb_dev, p.n_seq, p.frames, p.bins, n_bands,
                                       gspec, fast ? 1056 : 0, table, st);
  if (rc2 != TAC_OK) return rc2;
  if (fast) {
    rc2 = launch_stft2048_backward(p, gspec, bp.frames_out, st);
    if (rc2 != TAC_OK) return rc2;
    return launch_overlap_add(bp, grad_x, st);
  }
  return launch_stft_backward(bp, grad_x, bp.frames_out, workspace_bytes - spec_bytes - 32768, st);
}

extern "C" int tac_amplitude_to_db_backward_f32(const float* x, const float* grad_out, int64_t n, float amin, float* grad_x,
                                                void* stream) {
  using namespace tac;
  TAC_REQUIRE(n >= 0 && (n == 0 || (x && grad_out && grad_x)), TAC_ERR_INVALID, "amplitude_to_db_backward: bad arguments");
  if (n == 0) return TAC_OK;
  LaunchProbe probe(KIND_POINTWISE, as_stream(stream));
  amplitude_to_db_backward_kernel<<<pointwise_grid(n), 256, 0, as_stream(stream)>>>(x, grad_out, n, amin, grad_x);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

extern "C" int tac_complex_norm_backward_f32(const float* z, const float* grad_out, int64_t n, float power, float* grad_z,
                                             void* stream) {
  using namespace tac;
  TAC_REQUIRE(n >= 0 && (n == 0 || (z && grad_out && grad_z)), TAC_ERR_INVALID, "complex_norm_backward: bad arguments");
  if (n == 0) return TAC_OK;
  LaunchProbe probe(KIND_POINTWISE, as_stream(stream));
  complex_norm_backward_kernel<<<pointwise_grid(n), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2*>(z), grad_out, n,
                                                                                power, reinterpret_cast<float2*>(grad_z));
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}
