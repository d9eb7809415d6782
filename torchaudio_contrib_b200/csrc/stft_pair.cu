// K3: the one-kernel mel chain with two frames per warp (stft_pair_kernel.cuh), CUDA-core passes; host dispatch.
#include "stft_pair_kernel.cuh"

namespace tac {

bool stft2048_pair_applies(const StftParams& p) {
  return p.n_fft == 2048 && p.onesided && p.hop <= kPairMaxHop && (p.hop & 1) == 0 && p.g0 == 0 && p.g1 == p.n_seq * p.frames &&
         (p.out_mode == OUT_MEL_FUSED || p.out_mode == OUT_MEL_FUSED_PEERS);
}

int launch_stft2048_pair(const StftParams& p, cudaStream_t stream) { return launch_stft2048_pair_t<false>(p, stream); }

}  // namespace tac
