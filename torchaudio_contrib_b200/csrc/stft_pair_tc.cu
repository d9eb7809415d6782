// K3 with the second 32-point FFT pass on the tcgen05 tensor cores (stft_pair_kernel.cuh, TC = true);
// selected with tac_mel_kernel_variant(2) / TAC_MEL_VARIANT=2.
#include "stft_pair_kernel.cuh"

namespace tac {

int launch_stft2048_pair_tc(const StftParams& p, cudaStream_t stream) { return launch_stft2048_pair_t<true>(p, stream); }

}  // namespace tac
