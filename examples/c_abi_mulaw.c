/* A caller of the C ABI with no Python and no torch: mu-law encode / decode of a file of float32 samples.
 *
 *   cc -O2 -I include -I /usr/local/cuda/include examples/c_abi_mulaw.c -o c_abi_mulaw \
 *      -L torchaudio_contrib_b200/lib -ltac_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/torchaudio_contrib_b200/lib
 *   ./c_abi_mulaw samples.f32 codes.i64 decoded.f32
 *
 * The decision levels come from tac_mulaw_tables_host (shipped in the library for n_quantize = 256), so the codes are
 * bit-identical to the reference's torch CPU chain (functional.py:317-354) without torch anywhere in the process.
 * tests/test_gpu_parity.py::test_c_abi_without_python builds and runs this and compares with the oracle. */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "tac_b200.h"

#define CHECK_TAC(call)                                                        \
  do {                                                                         \
    int rc_ = (call);                                                          \
    if (rc_ != 0) {                                                            \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, tac_last_error());   \
      return 2;                                                                \
    }                                                                          \
  } while (0)
#define CHECK_CUDA(call)                                                       \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess) {                                                   \
      fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e_));       \
      return 3;                                                                \
    }                                                                          \
  } while (0)

int main(int argc, char** argv) {
  if (argc != 4) {
    fprintf(stderr, "usage: %s samples.f32 codes.i64 decoded.f32\n", argv[0]);
    return 1;
  }
  FILE* in = fopen(argv[1], "rb");
  if (!in) return 1;
  fseek(in, 0, SEEK_END);
  const long bytes = ftell(in);
  fseek(in, 0, SEEK_SET);
  const int64_t n = bytes / 4;
  float* x = (float*)malloc((size_t)n * 4);
  if (fread(x, 4, (size_t)n, in) != (size_t)n) return 1;
  fclose(in);

  const int n_quantize = 256;
  int n_thr = 0, idx_min = 0, exact = 0;
  float x_limit = 0.0f;
  CHECK_TAC(tac_mulaw_tables_host(n_quantize, NULL, 0, &n_thr, &idx_min, &x_limit, NULL, &exact));
  float* thr = (float*)malloc((size_t)n_thr * 4);
  float dec_tab[256];
  CHECK_TAC(tac_mulaw_tables_host(n_quantize, thr, n_thr, &n_thr, &idx_min, &x_limit, dec_tab, &exact));
  if (!exact) {
    fprintf(stderr, "the shipped table should be exact for 256 levels\n");
    return 4;
  }

  float *d_x, *d_thr, *d_lut, *d_dec;
  int64_t* d_codes;
  CHECK_CUDA(cudaMalloc((void**)&d_x, (size_t)n * 4));
  CHECK_CUDA(cudaMalloc((void**)&d_codes, (size_t)n * 8));
  CHECK_CUDA(cudaMalloc((void**)&d_dec, (size_t)n * 4));
  CHECK_CUDA(cudaMalloc((void**)&d_thr, (size_t)n_thr * 4));
  CHECK_CUDA(cudaMalloc((void**)&d_lut, 256 * 4));
  CHECK_CUDA(cudaMemcpy(d_x, x, (size_t)n * 4, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_thr, thr, (size_t)n_thr * 4, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_lut, dec_tab, 256 * 4, cudaMemcpyHostToDevice));
  CHECK_TAC(tac_mulaw_encode_f32_i64(d_x, n, n_quantize, d_thr, n_thr, idx_min, x_limit, d_codes, NULL));
  CHECK_TAC(tac_mulaw_decode_i64_f32(d_codes, n, n_quantize, d_lut, d_dec, NULL));
  CHECK_CUDA(cudaDeviceSynchronize());

  int64_t* codes = (int64_t*)malloc((size_t)n * 8);
  float* dec = (float*)malloc((size_t)n * 4);
  CHECK_CUDA(cudaMemcpy(codes, d_codes, (size_t)n * 8, cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(dec, d_dec, (size_t)n * 4, cudaMemcpyDeviceToHost));
  FILE* o1 = fopen(argv[2], "wb");
  FILE* o2 = fopen(argv[3], "wb");
  if (!o1 || !o2) return 1;
  fwrite(codes, 8, (size_t)n, o1);
  fwrite(dec, 4, (size_t)n, o2);
  fclose(o1);
  fclose(o2);
  printf("encoded and decoded %lld samples through the C ABI (library version %d, %lld kernel launches)\n", (long long)n, tac_version(),
         (long long)tac_launch_count());
  return 0;
}
