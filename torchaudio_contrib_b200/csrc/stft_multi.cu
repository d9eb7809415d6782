// K1b: the warp-level register FFT of stft.cu generalised to n_fft = 256 / 512 / 1024 / 4096.
//
// A frame of n_fft real samples is a complex FFT of C = n_fft / 2 points; C = R1 x R2.  For the small sizes both
// factors are at most 32, a warp handles F = 1024 / C frames at once and every lane owns 32 complex values; n_fft =
// 4096 is one frame per warp with 64 complex values per lane (8 warps of 255 registers instead of 16 of 128):
//
//     n_fft   C    R1 x R2   F (frames per warp batch)   pass-1 FFTs per lane   pass-2 FFTs per lane
//     4096  2048   32 x 64   1                           2 x 32-point           1 x 64-point
//     1024   512   16 x 32   2                           2 x 16-point           1 x 32-point
//      512   256   16 x 16   4                           2 x 16-point           2 x 16-point
//      256   128    8 x 16   8                           4 x  8-point           2 x 16-point
//
// Per frame: n = n1 + R2 n2, pass 1 transforms over n2 (R1 points) for each n1, the result goes through the
// warp's shared-memory slab laid out by pass-2 work item (frame, k2) with an odd row stride (conflict free),
// pass 2 applies W_C^(n1 k2) and transforms over n1 (R2 points): Z[R1 k1 + k2].  The real-FFT untangling
// pairs (k1, k2) with (R2-1-k1, R1-k2) inside the frame's group of R1 lanes by one shuffle.
// Frames are staged like in stft2048_kernel: one bulk async copy per interior frame into the slab, a gather
// for frames touching the padding.  n_fft = 2048 keeps its own kernels (stft.cu, stft_pair.cu); 8192 and the option
// surface that these do not cover (two-sided output) use stft_generic_kernel.
#include "bandplan.cuh"
#include "fft_regs.cuh"
#include "stft_params.cuh"
#include "tac_common.cuh"

namespace tac {

template <int LOG2N>
struct WarpFftShape {
  static constexpr int N = 1 << LOG2N, C = N / 2;
  static constexpr int R1 = (LOG2N == 12) ? 32 : ((LOG2N == 10) ? 16 : (LOG2N == 9 ? 16 : 8));
  static constexpr int R2 = C / R1;
  static constexpr int F = (C >= 1024) ? 1 : 1024 / C;
  static constexpr int V = F * C / 32;                   // complex values per lane: 32, or 64 at n_fft = 4096
  static constexpr int A1 = F * R2 / 32, A2 = F * R1 / 32;
  static constexpr int S = R2 + 1;                       // slab row stride (complex), odd
  static constexpr int WARPS = (V > 32) ? 8 : 16;
  static constexpr int THREADS = WARPS * 32;
  // floats per warp: >= F N samples, >= F R1 S complex (transposition) and >= F (C + 1) complex (the output tile of the public
  // layouts); for the small sizes SLAB = 4 or 8 (mod 32) so that the CTA-wide read-out of the tile spreads over the banks
  static constexpr int SLAB = (LOG2N == 12) ? 2 * 32 * 65 : (LOG2N == 8 ? 2216 : 2212);
  static_assert(A1 * R1 == V && A2 * R2 == V && F * R1 * S * 2 <= SLAB && F * N <= SLAB && 2 * F * (C + 1) <= SLAB && SLAB % 4 == 0, "shape");
};

template <int PMODE>
__device__ __forceinline__ float mw_power(float re, float im, float half_power) {
  const float s = fmaf(re, re, im * im);
  if constexpr (PMODE == 2) return s;
  if constexpr (PMODE == 1) return sqrtf(s);
  return s > 0.0f ? exp2f(half_power * __log2f(s)) : (half_power == 0.0f ? 1.0f : 0.0f);
}

template <int LOG2N, int OUT_MODE, int PMODE>
__global__ void __launch_bounds__(WarpFftShape<LOG2N>::THREADS, 1) stft_warp_kernel(const StftParams p) {
  using Sh = WarpFftShape<LOG2N>;
  constexpr int N = Sh::N, C = Sh::C, R1 = Sh::R1, R2 = Sh::R2, F = Sh::F, A1 = Sh::A1, A2 = Sh::A2, S = Sh::S, V = Sh::V;
  constexpr int kMwWarps = Sh::WARPS, kMwThreads = Sh::THREADS, kMwSlabFloats = Sh::SLAB;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* s_win = reinterpret_cast<float2*>(smem_raw);        // [C]        (w[2n], w[2n+1]) * 0.5 * scale
  float2* s_tw1 = s_win + C;                                  // [R2][R1]   W_C^(n1 k2)
  float2* s_tw2 = s_tw1 + C;                                  // [R2][R1]   W_N^(R1 k1 + k2)
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_tw2 + C);
  float* s_slab = reinterpret_cast<float*>(s_bar + kMwWarps);
  // OUT_MEL_RANGE: the range plan's meta and weights (bandplan.cuh), copied once per CTA
  const int2* s_meta = reinterpret_cast<const int2*>(s_slab + kMwWarps * kMwSlabFloats);
  const float* s_w = reinterpret_cast<const float*>(s_meta + (OUT_MODE == OUT_MEL_RANGE ? p.n_bands_pad : 0));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < C; i += kMwThreads) {
    const float g = 0.5f * p.scale;
    s_win[i] = make_float2(p.window[2 * i] * g, p.window[2 * i + 1] * g);
    const int a = i / R1, b = i % R1;                         // (n1, k2) resp. (k1, k2)
    float sn, cs;
    sincospif(-2.0f * (float)(a * b) / (float)C, &sn, &cs);
    s_tw1[i] = make_float2(cs, sn);
    sincospif(-2.0f * (float)(R1 * a + b) / (float)N, &sn, &cs);
    s_tw2[i] = make_float2(cs, sn);
  }
  if constexpr (OUT_MODE == OUT_MEL_RANGE) {
    const uint4* src = reinterpret_cast<const uint4*>(p.band_plan + 32);
    uint4* dst = reinterpret_cast<uint4*>(s_slab + kMwWarps * kMwSlabFloats);
    for (int i = tid; i < p.band_cmax; i += kMwThreads) dst[i] = __ldg(src + i);      // band_cmax: uint4 count of meta + weights
  }
  uint64_t* bar = s_bar + warp;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  float* slab_f = s_slab + warp * kMwSlabFloats;
  float2* slab = reinterpret_cast<float2*>(slab_f);
  const float half_power = 0.5f * p.power;
  const uint32_t frames_u = (uint32_t)p.frames;
  const int64_t n_batches = (p.g1 - p.g0 + F - 1) / F;
  uint32_t parity = 0;

  // Public layouts (complex / power, (n_seq, bins, frames[, 2])): the 16 warps of a CTA hold WARPS * F CONSECUTIVE frames per
  // round; each warp leaves its frames' spectra in its slab and the whole CTA writes the tile out, a bin's consecutive frames
  // as one run (a lane-per-bin store scatters 4- or 8-byte writes over `bins` rows: the complex STFT at 512 took 1.9x the
  // power spectrogram).  The round loop is therefore uniform over the CTA; a warp without a batch only joins the barriers.
  // (n_fft = 256 keeps the per-lane stores: its 8 frames per warp already give 32- / 64-byte runs and the two CTA barriers per
  // round cost more than they save there: 0.101 against 0.094 ms.)
  constexpr bool kTile = (OUT_MODE == OUT_COMPLEX_PUBLIC || OUT_MODE == OUT_POWER_PUBLIC) && LOG2N != 8;
  constexpr int kTileFrames = kMwWarps * F;
  int64_t* s_fr = reinterpret_cast<int64_t*>(s_slab + kMwWarps * kMwSlabFloats);      // kTile: output offset of (frame, bin 0), -1: no frame
  for (int64_t base = (int64_t)blockIdx.x * kMwWarps; base < n_batches; base += (int64_t)gridDim.x * kMwWarps) {
    const int64_t batch = base + warp;
    const bool have = batch < n_batches;
    if (!kTile && !have) continue;
    if (have) {
    const uint32_t gb = (uint32_t)(p.g0 + batch * F);         // first frame of the batch (flattened index)
    // ---- stage the F frames: bulk copy where possible, gather otherwise ---------------------------------
    __syncwarp();                                             // previous batch is done with the slab
    uint32_t bulk_bytes = 0;
    bool any_edge = false;                                    // some bulk-copied frame sticks out of its row
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const uint32_t g = gb + f;
      const uint32_t seq = g / frames_u, t = g - seq * frames_u;
      const int64_t start = (int64_t)t * p.hop - p.pad;
      const bool live = (int64_t)g < p.g1;
      const FrameSpan span = frame_span<N>(p, start);
      if (live && span.bulk) {
        bulk_bytes += (uint32_t)(span.hi - span.lo) * 4;
        any_edge |= (span.hi - span.lo) != N;
      }
    }
    if (elect_one()) {
      fence_proxy_async();
      if (bulk_bytes) mbar_arrive_expect_tx(bar, bulk_bytes);
    }
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const uint32_t g = gb + f;
      const uint32_t seq = g / frames_u, t = g - seq * frames_u;
      const int64_t start = (int64_t)t * p.hop - p.pad;
      const bool live = (int64_t)g < p.g1;
      const float* row = p.x + (int64_t)seq * p.seq_stride;
      const FrameSpan span = frame_span<N>(p, start);
      if (live && span.bulk) {
        if (elect_one()) bulk_g2s(slab_f + f * N + span.lo, row + (start + span.lo), (uint32_t)(span.hi - span.lo) * 4, bar);
      } else {
        gather_padded<N / 32>(slab_f + f * N, row, (int)start, (int)p.n_samples, p.pad_mode, lane, live);
      }
    }
    if (bulk_bytes) {
      mbar_wait(bar, parity);
      parity ^= 1u;
#pragma unroll
      for (int f = 0; f < F && any_edge; ++f) {                // frames that stick out of their row: fill from the slab
        const uint32_t g = gb + f;
        const uint32_t seq = g / frames_u, t = g - seq * frames_u;
        const int64_t start = (int64_t)t * p.hop - p.pad;
        const FrameSpan span = frame_span<N>(p, start);
        if ((int64_t)g < p.g1 && span.bulk)
          fill_padding<N>(slab_f + f * N, span, p.pad_mode, lane, p.x + (int64_t)seq * p.seq_stride, start);
      }
    }
    __syncwarp();

    // ---- pass 1: item i1 = a * 32 + lane = (frame f1, n1); R1-point FFT over n2 -------------------------
    float2 v[V];
#pragma unroll
    for (int a = 0; a < A1; ++a) {
      const int i1 = a * 32 + lane, f1 = i1 / R2, n1 = i1 % R2;
#pragma unroll
      for (int n2 = 0; n2 < R1; ++n2) {
        const float2 xs = slab[f1 * C + n1 + R2 * n2];
        const float2 w = s_win[n1 + R2 * n2];
        v[a * R1 + n2] = make_float2(xs.x * w.x, xs.y * w.y);
      }
    }
    __syncwarp();                                             // samples consumed: the slab becomes the transpose buffer
#pragma unroll
    for (int a = 0; a < A1; ++a) {
      float2 w[R1];
#pragma unroll
      for (int i = 0; i < R1; ++i) w[i] = v[a * R1 + i];
      dit_fft_fma<R1>(w);
      const int i1 = a * 32 + lane, f1 = i1 / R2, n1 = i1 % R2;
#pragma unroll
      for (int k2 = 0; k2 < R1; ++k2) slab[(f1 * R1 + k2) * S + n1] = w[bit_reverse<R1>(k2)];
    }
    __syncwarp();

    // ---- pass 2: item i2 = b * 32 + lane = (frame f2, k2); twiddle, R2-point FFT over n1, untangle -------
    const int kk2 = lane % R1;                                // k2 of this lane's items (R1 divides 32)
    const int partner = (lane & ~(R1 - 1)) | ((R1 - kk2) & (R1 - 1));
    if constexpr (OUT_MODE == OUT_MEL_RANGE) {
      // Fused filterbank (replaces apply_filterbank's transpose + matmul + transpose, functional.py:183-184, and the dB
      // pass :291-296): every pass-2 input is taken out of the slab first, so that the slab can take the F frames' power
      // spectra ((C + 1) floats each); then lane <-> bands lane + 32 j sum their bin ranges for all F frames at once.
      constexpr int SP = C + 1;
      float2 uu[V];
#pragma unroll
      for (int b = 0; b < A2; ++b) {
        const int i2 = b * 32 + lane;
#pragma unroll
        for (int n1 = 0; n1 < R2; ++n1) {
          const float2 a = slab[i2 * S + n1];
          const float2 w = s_tw1[n1 * R1 + kk2];
          uu[b * R2 + n1] = make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
        }
      }
      __syncwarp();
#pragma unroll
      for (int b = 0; b < A2; ++b) {
        const int f2 = (b * 32 + lane) / R1;
        float2 u[R2];
#pragma unroll
        for (int i = 0; i < R2; ++i) u[i] = uu[b * R2 + i];
        dit_fft_fma<R2>(u);
        float* ps = slab_f + f2 * SP;
#pragma unroll
        for (int k1 = 0; k1 < R2; ++k1) {
          const float2 z = u[bit_reverse<R2>(k1)];
          float2 q;
          q.x = __shfl_sync(0xffffffffu, u[bit_reverse<R2>(R2 - 1 - k1)].x, partner);
          q.y = __shfl_sync(0xffffffffu, u[bit_reverse<R2>(R2 - 1 - k1)].y, partner);
          if (kk2 == 0) q = u[bit_reverse<R2>((R2 - k1) % R2)];
          const float a = z.x + q.x, bb = z.y - q.y, gs = z.y + q.y, h = q.x - z.x;
          const float2 w = s_tw2[k1 * R1 + kk2];
          ps[R1 * k1 + kk2] = mw_power<PMODE>(fmaf(w.x, gs, fmaf(-w.y, h, a)), fmaf(w.x, h, fmaf(w.y, gs, bb)), half_power);
        }
        if (kk2 == 0) ps[C] = mw_power<PMODE>(2.0f * (u[0].x - u[0].y), 0.0f, half_power);
      }
      __syncwarp();
      uint32_t seq_f[F], t_f[F];
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const uint32_t g = gb + f;
        seq_f[f] = g / frames_u;
        t_f[f] = g - seq_f[f] * frames_u;
      }
      const int nbj = p.n_bands_pad >> 5;
      for (int j = 0; j < nbj; ++j) {
        const int m = lane + 32 * j;
        const int2 me = s_meta[m];
        const int lo = me.x & 0xffff, len = me.x >> 16;
        const float* w = s_w + me.y;
        const float* pp = slab_f + lo;
        float acc[F];
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] = 0.0f;
        for (int i = 0; i < len; ++i) {
          const float wv = w[i];
#pragma unroll
          for (int f = 0; f < F; ++f) acc[f] = fmaf(pp[f * SP + i], wv, acc[f]);
        }
        if (m < p.n_bands) {
#pragma unroll
          for (int f = 0; f < F; ++f) {
            if ((int64_t)(gb + f) >= p.g1) continue;
            float r = acc[f];
            if (p.to_db) {
              float s2 = r * r;
              s2 = (s2 < p.amin) ? p.amin : s2;
              r = 10.0f * (log10f(s2) - p.log10_ref);
            }
            __stcs(p.out + (int64_t)seq_f[f] * p.out_seq_stride + (int64_t)t_f[f] * p.out_t_stride + (int64_t)m * p.out_band_stride, r);
          }
        }
      }
      continue;
    }
    if constexpr (kTile) {
      // every pass-2 input leaves the slab first (as in the fused-filterbank branch), then the slab takes the tile
      constexpr int SP = C + 1;
      float2 uu[V];
#pragma unroll
      for (int b = 0; b < A2; ++b) {
        const int i2 = b * 32 + lane;
#pragma unroll
        for (int n1 = 0; n1 < R2; ++n1) {
          const float2 a = slab[i2 * S + n1];
          const float2 w = s_tw1[n1 * R1 + kk2];
          uu[b * R2 + n1] = make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
        }
      }
      __syncwarp();
#pragma unroll
      for (int b = 0; b < A2; ++b) {
        const int f2 = (b * 32 + lane) / R1;
        float2 u[R2];
#pragma unroll
        for (int i = 0; i < R2; ++i) u[i] = uu[b * R2 + i];
        dit_fft_fma<R2>(u);
        auto put = [&](int k, float re, float im) {
          if constexpr (OUT_MODE == OUT_COMPLEX_PUBLIC) slab[f2 * SP + k] = make_float2(re, im);
          else slab_f[f2 * SP + k] = mw_power<PMODE>(re, im, half_power);
        };
#pragma unroll
        for (int k1 = 0; k1 < R2; ++k1) {
          const float2 z = u[bit_reverse<R2>(k1)];
          float2 q;
          q.x = __shfl_sync(0xffffffffu, u[bit_reverse<R2>(R2 - 1 - k1)].x, partner);
          q.y = __shfl_sync(0xffffffffu, u[bit_reverse<R2>(R2 - 1 - k1)].y, partner);
          if (kk2 == 0) q = u[bit_reverse<R2>((R2 - k1) % R2)];
          const float a = z.x + q.x, bb = z.y - q.y, gs = z.y + q.y, h = q.x - z.x;
          const float2 w = s_tw2[k1 * R1 + kk2];
          put(R1 * k1 + kk2, fmaf(w.x, gs, fmaf(-w.y, h, a)), fmaf(w.x, h, fmaf(w.y, gs, bb)));
        }
        if (kk2 == 0) put(C, 2.0f * (u[0].x - u[0].y), 0.0f);
      }
    } else {
#pragma unroll
    for (int b = 0; b < A2; ++b) {
      const int i2 = b * 32 + lane, f2 = i2 / R1;
      float2 u[R2];
#pragma unroll
      for (int n1 = 0; n1 < R2; ++n1) {
        const float2 a = slab[i2 * S + n1];
        const float2 w = s_tw1[n1 * R1 + kk2];
        u[n1] = make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
      }
      dit_fft_fma<R2>(u);                                     // u[bit_reverse(k1)] = Z[R1 k1 + k2] / 2

      const uint32_t g = gb + f2;
      const bool live = (int64_t)g < p.g1;
      const uint32_t seq = g / frames_u, t = g - seq * frames_u;
      const int64_t row = (int64_t)g - p.g0;
      auto emit = [&](int k, float re, float im) {
        if (!live) return;
        if constexpr (OUT_MODE == OUT_POWER_ROWS) {
          p.out[power_tile_index(row, k, p.kpad)] = mw_power<PMODE>(re, im, half_power);
        } else if constexpr (OUT_MODE == OUT_POWER_PUBLIC) {
          p.out[((int64_t)seq * p.bins + k) * p.frames + t] = mw_power<PMODE>(re, im, half_power);
        } else {
          reinterpret_cast<float2*>(p.out)[((int64_t)seq * p.bins + k) * p.frames + t] = make_float2(re, im);
        }
      };
#pragma unroll
      for (int k1 = 0; k1 < R2; ++k1) {
        const float2 z = u[bit_reverse<R2>(k1)];
        float2 q;
        q.x = __shfl_sync(0xffffffffu, u[bit_reverse<R2>(R2 - 1 - k1)].x, partner);
        q.y = __shfl_sync(0xffffffffu, u[bit_reverse<R2>(R2 - 1 - k1)].y, partner);
        if (kk2 == 0) q = u[bit_reverse<R2>((R2 - k1) % R2)];
        const float a = z.x + q.x, bb = z.y - q.y, gs = z.y + q.y, h = q.x - z.x;
        const float2 w = s_tw2[k1 * R1 + kk2];
        emit(R1 * k1 + kk2, fmaf(w.x, gs, fmaf(-w.y, h, a)), fmaf(w.x, h, fmaf(w.y, gs, bb)));
      }
      // Nyquist bin C (owned by the k2 = 0 lane) and, for power tiles, the zero fill of its 32-bin slice
      const float2 z0 = u[0];
      const float nyq = 2.0f * (z0.x - z0.y);
      if constexpr (OUT_MODE == OUT_POWER_ROWS) {
        if (live) {
#pragma unroll
          for (int j = 0; j < 32 / R1; ++j) {
            const int kq = kk2 + R1 * j;
            p.out[power_tile_index(row, C + kq, p.kpad)] = (kq == 0) ? mw_power<PMODE>(nyq, 0.0f, half_power) : 0.0f;
          }
        }
      } else {
        if (kk2 == 0) emit(C, nyq, 0.0f);
      }
    }
    }   // !kTile
    }   // have
    if constexpr (kTile) {
      constexpr int SP = C + 1;
      __syncthreads();                                        // every warp's part of the tile is in its slab
      for (int i = tid; i < kTileFrames; i += kMwThreads) {
        const int64_t g = p.g0 + base * F + i;
        int64_t off = -1;
        if (g < p.g1) {
          const int64_t seq = g / p.frames, t = g - seq * p.frames;
          off = seq * p.bins * p.frames + t;
        }
        s_fr[i] = off;
      }
      __syncthreads();
      constexpr int FPW = kTileFrames < 32 ? kTileFrames : 32;    // frames per warp instruction
      constexpr int BPW = 32 / FPW;                               // bins per warp instruction
      const int fl = lane % FPW, bl = lane / FPW;
      for (int bin0 = warp * BPW; bin0 <= C; bin0 += kMwWarps * BPW) {
        const int bin = bin0 + bl;
#pragma unroll
        for (int c = 0; c < kTileFrames / FPW; ++c) {
          const int fr = c * FPW + fl;
          const int64_t off = s_fr[fr];
          if (bin <= C && off >= 0) {
            const float* src = s_slab + (fr / F) * kMwSlabFloats;
            if constexpr (OUT_MODE == OUT_COMPLEX_PUBLIC)
              __stcs(reinterpret_cast<float2*>(p.out) + off + (int64_t)bin * p.frames, reinterpret_cast<const float2*>(src)[(fr % F) * SP + bin]);
            else
              __stcs(p.out + off + (int64_t)bin * p.frames, src[(fr % F) * SP + bin]);
          }
        }
      }
      __syncthreads();                                        // tile written: the slabs may take the next round's samples
    }
  }
}

template <int LOG2N>
static size_t warp_kernel_smem() {
  using Sh = WarpFftShape<LOG2N>;
  return 3 * (size_t)Sh::C * sizeof(float2) + Sh::WARPS * sizeof(uint64_t) + (size_t)Sh::WARPS * Sh::SLAB * sizeof(float) +
         (size_t)Sh::WARPS * Sh::F * sizeof(int64_t);        // + the frame table of the public-layout tile
}

template <int LOG2N>
static int launch_warp_kernel(const StftParams& p, cudaStream_t stream) {
  using Kernel = void (*)(const StftParams);
  Kernel k = nullptr;
  if (p.out_mode == OUT_COMPLEX_PUBLIC) k = stft_warp_kernel<LOG2N, OUT_COMPLEX_PUBLIC, 1>;
  else if (p.out_mode == OUT_POWER_PUBLIC)
    k = p.power_mode == 2 ? stft_warp_kernel<LOG2N, OUT_POWER_PUBLIC, 2>
                          : (p.power_mode == 1 ? stft_warp_kernel<LOG2N, OUT_POWER_PUBLIC, 1> : stft_warp_kernel<LOG2N, OUT_POWER_PUBLIC, 0>);
  else if (p.out_mode == OUT_MEL_RANGE)
    k = p.power_mode == 2 ? stft_warp_kernel<LOG2N, OUT_MEL_RANGE, 2>
                          : (p.power_mode == 1 ? stft_warp_kernel<LOG2N, OUT_MEL_RANGE, 1> : stft_warp_kernel<LOG2N, OUT_MEL_RANGE, 0>);
  else
    k = p.power_mode == 2 ? stft_warp_kernel<LOG2N, OUT_POWER_ROWS, 2>
                          : (p.power_mode == 1 ? stft_warp_kernel<LOG2N, OUT_POWER_ROWS, 1> : stft_warp_kernel<LOG2N, OUT_POWER_ROWS, 0>);
  const size_t smem = warp_kernel_smem<LOG2N>() + (p.out_mode == OUT_MEL_RANGE ? (size_t)p.band_cmax * 16 : 0);
  TAC_REQUIRE(smem <= 227 * 1024, TAC_ERR_UNSUPPORTED, "melspec: range plan of %d bytes does not fit beside the FFT buffers", p.band_cmax * 16);
  TAC_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t batches = (p.g1 - p.g0 + WarpFftShape<LOG2N>::F - 1) / WarpFftShape<LOG2N>::F;
  const int64_t want = (batches + WarpFftShape<LOG2N>::WARPS - 1) / WarpFftShape<LOG2N>::WARPS;
  const int grid = (int)(want < sm_count() ? want : sm_count());
  LaunchProbe probe(KIND_STFT, stream);
  k<<<grid, WarpFftShape<LOG2N>::THREADS, smem, stream>>>(p);
  TAC_CUDA_OK(cudaGetLastError());
  return TAC_OK;
}

// n_fft = 256 / 512 / 1024 / 4096, one-sided output: returns TAC_ERR_UNSUPPORTED for anything else (caller falls
// through to the generic kernel)
int launch_stft_warp(const StftParams& p, cudaStream_t stream) {
  if (!p.onesided) return TAC_ERR_UNSUPPORTED;
  switch (p.n_fft) {
    case 256: return launch_warp_kernel<8>(p, stream);
    case 512: return launch_warp_kernel<9>(p, stream);
    case 1024: return launch_warp_kernel<10>(p, stream);
    case 4096: return launch_warp_kernel<12>(p, stream);
    default: return TAC_ERR_UNSUPPORTED;
  }
}

}  // namespace tac
