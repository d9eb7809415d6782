// Launch description shared by the STFT kernels (stft.cu) and the pipeline driver (abi.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tac {

enum StftOutMode {
  OUT_COMPLEX_PUBLIC = 0,   // (n_seq, bins, frames, 2)   -- reference layout of `stft`
  OUT_POWER_PUBLIC = 1,     // (n_seq, bins, frames)      -- reference layout of `Spectrogram`
  OUT_POWER_ROWS = 2        // |X|^p in "power tiles": the swizzled tensor-core operand layout (below)
};

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
#endif

int launch_stft_warp(const StftParams& p, cudaStream_t stream);   // stft_multi.cu: n_fft = 256 / 512 / 1024
int fill_stft_params(StftParams& p, const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                     const float* window, int n_fft, int hop, int center, int pad_mode, int normalized, int onesided);
int launch_stft(const StftParams& p, cudaStream_t stream);

}  // namespace tac
