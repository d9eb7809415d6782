// Microbenchmark: dependent-issue latency of FFMA vs FFMA2 / FADD2 on sm_100a: one warp per SM sub-partition runs ILP
// independent chains; cycles per instruction * ILP at ILP = 1 is the latency, the ILP at which it stops falling is what a
// warp needs to keep the pipe busy on its own.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_latency ffma2_latency.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int MODE, int ILP>
__global__ void k(float* out, long long* cyc, int iters, float s) {
  float a1[ILP]; u64 a2[ILP];
  for (int i = 0; i < ILP; ++i) { a1[i] = threadIdx.x + i; a2[i] = mk(threadIdx.x + i, i); }
  const u64 s2 = mk(s, s + 1.0f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 64 / ILP; ++rep) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (MODE == 0) a1[i] = fma1(a1[i], s, a1[i]);
        else if (MODE == 1) a2[i] = fma2(a2[i], s2, a2[i]);
        else a2[i] = add2(a2[i], s2);
      }
    }
  }
  long long t1 = clock64();
  float r = 0;
  for (int i = 0; i < ILP; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a2[i])); r += a1[i] + x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int ILP>
void run(const char* name, int warps) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  k<MODE, ILP><<<148, warps * 32>>>(out, cyc, iters, 1.0001f);
  k<MODE, ILP><<<148, warps * 32>>>(out, cyc, iters, 1.0001f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < 148; ++i) mean += h[i]; mean /= 148;
  printf("%-6s ILP %d, %d warp(s) per SMSP: %.2f cycles per instruction of one warp (%s)\n", name, ILP, warps / 4, mean / (iters * 64.0), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 12}) {
    run<0, 1>("FFMA", w); run<0, 2>("FFMA", w); run<0, 4>("FFMA", w);
    run<1, 1>("FFMA2", w); run<1, 2>("FFMA2", w); run<1, 4>("FFMA2", w); run<1, 8>("FFMA2", w);
    run<2, 1>("FADD2", w); run<2, 2>("FADD2", w); run<2, 4>("FADD2", w);
  }
  return 0;
}
