// Launch description shared by the STFT kernels (stft.cu) and the pipeline driver (abi.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tac {

enum StftOutMode {
  OUT_COMPLEX_PUBLIC = 0,   // (n_seq, bins, frames, 2)   -- reference layout of `stft`
  OUT_POWER_PUBLIC = 1,     // (n_seq, bins, frames)      -- reference layout of `Spectrogram`
  OUT_POWER_ROWS = 2,       // |X|^p in "power tiles": the swizzled tensor-core operand layout (below)
  OUT_MEL_FUSED = 3,        // |X|^p contracted with a two-band filterbank inside the kernel (bandplan.cuh) [+ dB]
  OUT_MEL_FUSED_PEERS = 4,  // the same, every frame's bands stored into the output buffers of all ranks of the box (peers.cu)
  OUT_MEL_RANGE = 5         // |X|^p contracted with a range plan (bandplan.cuh) inside the warp kernels for n_fft != 2048 [+ dB]
};
constexpr int kMaxPeers = 8;           // GPUs of one NVSwitch box

struct StftParams {
  const float* x;          // (n_seq, n_samples), rows seq_stride apart
  const float* window;     // n_fft floats, centre-padded
  float* out;
  int64_t n_seq, n_samples, seq_stride;
  int64_t frames;          // frames per sequence
  int64_t g0, g1;          // flattened frame range [g0, g1) handled by this launch (g = seq * frames + t)
  int n_fft, hop, pad, pad_mode;
  int onesided, bins, kpad;
  int bulk_ok;             // interior frames may use the 1-D bulk copy (16 B alignment holds)
  int out_mode;
  int power_mode;          // 2: |X|^2, 1: |X|, 0: |X|^power
  float power;
  float scale;             // n_fft^-0.5 when normalized, else 1
  int debug;               // TAC_K1_TRACE: clock stamps of CTA 0 / warp 0 (timing experiments only)
  // OUT_MEL_FUSED only: the band plan and where band m of frame (seq, t) goes:
  //   out[seq * out_seq_stride + t * out_t_stride + m * out_band_stride]
  const unsigned char* band_plan;
  int band_cmax, n_bands, n_bands_pad, to_db;
  int band_fast, band_off_fast;   // per-lane 4 x 4 list form present (<= 128 bands, <= 4 entries), its byte offset
  float amin, log10_ref;
  int64_t out_seq_stride, out_t_stride, out_band_stride;
  // OUT_MEL_FUSED_PEERS only: base of every rank's full output (its own included, peer-mapped pointers for the
  // others); `out` is unused and the strides above address the full (all ranks) tensor from peer_seq0 on
  float* peer_out[kMaxPeers];
  int n_peers;
  int peer_multicast;      // peer_out[0] is a MULTICAST address (multicast.cu): one multimem store reaches every replica
  int64_t peer_seq0;       // first sequence of this rank inside the full output
};

// Power tiles (OUT_POWER_ROWS): frames are grouped in tiles of 128; for each tile and each 32-bin slice the
// (128 x 32) fp32 block is stored contiguously (16 KB) in the 128B-swizzled K-major layout the tcgen05 A
// operand wants (8-row x 128-byte atoms, 16-byte columns XOR-ed with the row index), so the filterbank kernel
// fetches any 8-row-aligned range of a block with one bulk copy, and the STFT kernel addresses the 33 slices of
// a frame as base + slice * 16 KB (an immediate offset: no address arithmetic in its store loop):
//   float index = ((row / 128) * (kpad / 32) + bin / 32) * 4096 + (ri / 8) * 256 + (ri % 8) * 32
//                 + (((kk / 4) ^ (ri % 8)) * 4) + kk % 4,        ri = row % 128, kk = bin % 32
constexpr int kPowerTileRows = 128;
__host__ __device__ inline int64_t power_tile_index(int64_t row, int bin, int kpad) {
  const int64_t tile = row >> 7;
  const int ri = (int)(row & 127), kk = bin & 31;
  return (tile * (kpad >> 5) + (bin >> 5)) * 4096 + (ri >> 3) * 256 + (ri & 7) * 32 + ((((kk >> 2) ^ (ri & 7)) << 2) | (kk & 3));
}
// bytes of workspace for `rows` frames (whole tiles, plus one tile of slack for 8-row-aligned over-reads)
__host__ __device__ inline int64_t power_tile_bytes(int64_t rows, int kpad) {
  return ((rows + kPowerTileRows - 1) / kPowerTileRows + 1) * (int64_t)kPowerTileRows * kpad * 4;
}

#ifdef __CUDACC__
// x[s] under torch.stft's centre padding (torch.nn.functional.pad modes), 0 <= pad < n guaranteed by the host
__device__ __forceinline__ float fetch_padded(const float* __restrict__ row, int64_t s, int64_t n, int pad_mode) {
  if (s >= 0 && s < n) return __ldg(row + s);
  switch (pad_mode) {
    case 0: s = (s < 0) ? -s : 2 * (n - 1) - s; break;          // TAC_PAD_REFLECT
    case 2: s = (s < 0) ? 0 : n - 1; break;                     // TAC_PAD_REPLICATE
    case 3: s = (s < 0) ? s + n : s - n; break;                 // TAC_PAD_CIRCULAR
    default: return 0.0f;                                       // TAC_PAD_CONSTANT
  }
  return (s >= 0 && s < n) ? __ldg(row + s) : 0.0f;
}

// Gather of 32 * COUNT consecutive padded samples (lane-strided) into shared memory, in explicit batches of
// BATCH loads so that BATCH global loads are in flight per lane (a
// load-store-load-store sequence costs a full memory latency per sample: measured 16 us per edge frame).
// 32-bit index arithmetic: the host guarantees n_samples + n_fft < 2^31.
__device__ __forceinline__ int padded_index(int s, int n, int pad_mode) {
  int r = s;
  if (pad_mode == 0) r = (s < 0) ? -s : ((s >= n) ? 2 * (n - 1) - s : s);          // reflect
  else if (pad_mode == 3) r = (s < 0) ? s + n : ((s >= n) ? s - n : s);              // circular
  return r < 0 ? 0 : (r >= n ? n - 1 : r);                                           // replicate / safety clamp
}
template <int COUNT, int BATCH = 8>
__device__ __forceinline__ void gather_padded(float* dst, const float* __restrict__ row, int start, int n, int pad_mode, int lane,
                                              bool live = true) {
  static_assert(COUNT % BATCH == 0, "whole batches");
#pragma unroll 1
  for (int base = 0; base < COUNT; base += BATCH) {
    float tmp[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int s = start + lane + 32 * (base + u);
      const bool inside = (s >= 0) & (s < n);
      const float v = live ? __ldg(row + padded_index(s, n, pad_mode)) : 0.0f;
      tmp[u] = (pad_mode == 1 && !inside) ? 0.0f : v;
    }
#pragma unroll
    for (int u = 0; u < BATCH; ++u) dst[lane + 32 * (base + u)] = tmp[u];
  }
}

// ---- frame staging shared by the warp-level kernels ------------------------------------------------------
// A frame is samples [start, start + N) of its row, start = t * hop - pad.  With 16-byte alignment
// (p.bulk_ok, and n_samples % 4 == 0 for frames that cross the end) the part that lies inside the row is
// fetched with ONE bulk async copy into its place in the slab; whatever sticks out on the left / right is
// then produced from the slab itself: reflection = mirror inside the slab, replicate = edge value, constant =
// zeros.  (Per-sample gathers with index arithmetic made an edge frame ~8x as expensive as an interior one and
// the warps owning the first two frames of a sequence set the kernel's makespan.)  Circular padding and frames
// wider than the row use the gather.
struct FrameSpan {
  int lo, hi;        // slab indices [lo, hi) filled by the bulk copy (0 / N for interior frames)
  bool bulk;
};
template <int N>
__device__ __forceinline__ FrameSpan frame_span(const StftParams& p, int start) {   // 32-bit: n_samples + 2 n_fft < 2^31
  FrameSpan f;
  const int n = (int)p.n_samples;
  const int lo = start < 0 ? -start : 0;
  const int hi = (start + N > n) ? n - start : N;
  f.lo = lo;
  f.hi = hi;
  const bool interior = lo == 0 && hi == N;
  f.bulk = p.bulk_ok && hi > lo && (interior || (p.pad_mode != 3 && (p.n_samples & 3) == 0 && !(lo > 0 && hi < N)));
  return f;
}
// after the bulk copy has landed: fill slab[0, lo) and slab[hi, N) (warp-cooperative, then __syncwarp).
// `row` / `start` locate the frame in its row: a reflected sample whose source lies outside the part of the
// frame that was copied (only when the padding is as long as the valid part, e.g. sample 0 of frame 0) is
// read from global memory.
template <int N>
__device__ __forceinline__ void fill_padding(float* slab_f, const FrameSpan f, int pad_mode, int lane,
                                             const float* __restrict__ row, int start) {
  if (f.lo == 0 && f.hi == N) return;
  if (pad_mode == 0) {                                   // reflect: x[-k] = x[k], x[n-1+k] = x[n-1-k]
    for (int j = lane; j < f.lo; j += 32) {
      const int src = 2 * f.lo - j;
      slab_f[j] = (src < f.hi) ? slab_f[src] : __ldg(row + (start + src));
    }
    for (int j = f.hi + lane; j < N; j += 32) {
      const int src = 2 * (f.hi - 1) - j;
      slab_f[j] = (src >= f.lo) ? slab_f[src] : __ldg(row + (start + src));
    }
  } else if (pad_mode == 2) {                            // replicate
    const float a = slab_f[f.lo], b = slab_f[f.hi - 1];
    for (int j = lane; j < f.lo; j += 32) slab_f[j] = a;
    for (int j = f.hi + lane; j < N; j += 32) slab_f[j] = b;
  } else {                                               // constant
    for (int j = lane; j < f.lo; j += 32) slab_f[j] = 0.0f;
    for (int j = f.hi + lane; j < N; j += 32) slab_f[j] = 0.0f;
  }
  __syncwarp();
}
#endif

int launch_stft_warp(const StftParams& p, cudaStream_t stream);   // stft_multi.cu: n_fft = 256 / 512 / 1024
int fill_stft_params(StftParams& p, const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                     const float* window, int n_fft, int hop, int center, int pad_mode, int normalized, int onesided);
int launch_stft(const StftParams& p, cudaStream_t stream);
// stft_pair.cu: the one-kernel mel path (OUT_MEL_FUSED / OUT_MEL_FUSED_PEERS, n_fft = 2048) with two frames per warp
bool stft2048_pair_applies(const StftParams& p);    // hop <= 512 and even, whole sequences
int launch_stft2048_pair(const StftParams& p, cudaStream_t stream);
int launch_stft2048_pair_tc(const StftParams& p, cudaStream_t stream);   // stft_pair_tc.cu: second FFT pass on tcgen05
// stft.cu: backward of stft + |.|^p for n_fft = 2048; gspec_fm: (frames of the launch, 1056) frame-major gradient of
// |X|^p, frames_out: (frames of the launch, 2048) windowed frame gradients (p.power / p.power_mode as in the forward pass)
int launch_stft2048_backward(const StftParams& p, const float* gspec_fm, float* frames_out, cudaStream_t stream);

}  // namespace tac
