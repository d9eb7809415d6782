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
  int tile_rows;           // OUT_POWER_ROWS: frames per tile (multiple of 8, <= 128)
  int bulk_ok;             // interior frames may use the 1-D bulk copy (16 B alignment holds)
  int out_mode;
  int power_mode;          // 2: |X|^2, 1: |X|, 0: |X|^power
  float power;
  float scale;             // n_fft^-0.5 when normalized, else 1
};

// Power tiles (OUT_POWER_ROWS): frames are grouped in tiles of `tile_rows`; for each tile and each
// 32-bin slice the (tile_rows x 32) block is stored contiguously in the 128B-swizzled K-major layout
// the tcgen05 A operand wants, so the filterbank kernel fetches it with one bulk copy:
//   float index = ((row / tile_rows) * (kpad / 32) + bin / 32) * tile_rows * 32
//                 + (ri / 8) * 256 + (ri % 8) * 32 + (((kk / 4) ^ (ri % 8)) * 4) + kk % 4,   ri = row % tile_rows, kk = bin % 32
__host__ __device__ inline int64_t power_tile_index(int64_t row, int bin, int tile_rows, int kpad) {
  const int64_t tile = row / tile_rows;
  const int ri = (int)(row - tile * tile_rows), kk = bin & 31;
  return (tile * (kpad >> 5) + (bin >> 5)) * (int64_t)tile_rows * 32 + (ri >> 3) * 256 + (ri & 7) * 32 +
         ((((kk >> 2) ^ (ri & 7)) << 2) | (kk & 3));
}

int fill_stft_params(StftParams& p, const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                     const float* window, int n_fft, int hop, int center, int pad_mode, int normalized, int onesided);
int launch_stft(const StftParams& p, cudaStream_t stream);

}  // namespace tac
