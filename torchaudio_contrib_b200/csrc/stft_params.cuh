// Launch description shared by the STFT kernels (stft.cu) and the pipeline driver (abi.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tac {

enum StftOutMode {
  OUT_COMPLEX_PUBLIC = 0,   // (n_seq, bins, frames, 2)   -- reference layout of `stft`
  OUT_POWER_PUBLIC = 1,     // (n_seq, bins, frames)      -- reference layout of `Spectrogram`
  OUT_POWER_ROWS = 2        // (g1 - g0, kpad) frame-major |X|^p rows for the filterbank kernel
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

int fill_stft_params(StftParams& p, const float* x, int64_t n_seq, int64_t n_samples, int64_t seq_stride,
                     const float* window, int n_fft, int hop, int center, int pad_mode, int normalized, int onesided);
int launch_stft(const StftParams& p, cudaStream_t stream);

}  // namespace tac
