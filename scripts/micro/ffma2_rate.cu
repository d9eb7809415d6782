// Microbenchmark: issue rate of FFMA (scalar) vs FFMA2 / FADD2 / FMUL2 (packed f32x2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu ; run on the GPU box.
// Prints cycles per warp-instruction per SM sub-partition for 1, 2, 4 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float s) {
  constexpr int NACC = 8;
  float a1[NACC]; u64 a2[NACC];
  for (int i = 0; i < NACC; ++i) { a1[i] = threadIdx.x + i; a2[i] = mk(threadIdx.x + i, i); }
  const u64 s2 = mk(s, s + 1.0f);
  const float t = s * 0.5f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        if (MODE == 0) a1[i] = fma1(a1[i], s, t);                 // FFMA, 3 register operands
        else if (MODE == 1) a2[i] = fma2(a2[i], s2, a2[(i + 1) % NACC]);   // FFMA2, packed operands
        else if (MODE == 2) a2[i] = fma2(a2[i], mk(s, s), a2[i]);  // FFMA2 with a broadcast scalar operand
        else if (MODE == 3) a2[i] = add2(a2[i], s2);              // FADD2
        else if (MODE == 4) { a1[i] = fma1(a1[i], s, t); a2[i] = add2(a2[i], s2); }   // mixed: FFMA + FADD2
      }
    }
  }
  long long t1 = clock64();
  float r = 0;
  for (int i = 0; i < NACC; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a2[i])); r += a1[i] + x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  k<MODE><<<148, warps_per_sm * 32>>>(out, cyc, iters, 1.0001f);
  k<MODE><<<148, warps_per_sm * 32>>>(out, cyc, iters, 1.0001f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < 148; ++i) mean += h[i]; mean /= 148;
  const double n_inst = (double)iters * 64 * (MODE == 4 ? 2 : 1);         // per warp
  const double per_smsp = n_inst * warps_per_sm / 4.0;
  printf("%-28s warps/SM %2d: %.3f cycles per warp-instruction per SMSP  (%s)\n", name, warps_per_sm, mean / per_smsp, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 16}) {
    run<0>("FFMA", w); run<1>("FFMA2 packed", w); run<2>("FFMA2 broadcast operand", w); run<3>("FADD2", w); run<4>("FFMA + FADD2 interleaved", w);
  }
  return 0;
}
