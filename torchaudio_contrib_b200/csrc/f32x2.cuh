// Packed pairs of fp32 values and the sm_100 two-wide fp32 instructions (SASS FADD2 / FMUL2 / FFMA2).
//
// A `pk` holds the same quantity of TWO independent problems (here: two STFT frames handled by one warp) in an
// aligned 64-bit register pair.  One FFMA2 does both frames' multiply-add in ONE issue slot; measured on B200
// (scripts/micro/ffma2_rate.cu, profiles/r02_ffma2_rate.txt): FFMA 1.03 cycles per warp-instruction per SM
// sub-partition, FFMA2 / FADD2 / FMUL2 2.01 -- the same flops per clock in half the issue slots, which is what a
// kernel bound by instruction issue (the one-frame-per-warp mel kernel: issue-active 62 %, fp32 pipe 41 %) wants.
// A scalar operand shared by both halves costs nothing: ptxas turns `bc(s)` into a broadcast operand (`R4.F32`),
// with negation folded in, and a compile-time constant into an immediate.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tac {

struct pk {
  unsigned long long v;
};

__device__ __forceinline__ pk mk2(float a, float b) {
  pk r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ pk bc(float s) { return mk2(s, s); }
__device__ __forceinline__ float lo(pk a) { return __uint_as_float((uint32_t)a.v); }
__device__ __forceinline__ float hi(pk a) { return __uint_as_float((uint32_t)(a.v >> 32)); }

__device__ __forceinline__ pk operator+(pk a, pk b) {
  pk d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ pk operator-(pk a, pk b) {
  pk d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ pk operator*(pk a, pk b) {
  pk d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ pk pfma(pk a, pk b, pk c) {
  pk d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return d;
}
// scalar * pair (+ pair): the scalar is shared by both halves
__device__ __forceinline__ pk operator*(float s, pk a) { return a * bc(s); }
__device__ __forceinline__ pk pfma(float s, pk b, pk c) { return pfma(bc(s), b, c); }
__device__ __forceinline__ pk psel(bool c, pk a, pk b) {
  pk r;
  r.v = c ? a.v : b.v;
  return r;
}

// the scalar spellings of the same operations, so that code templated on the real type reads the same for both
__device__ __forceinline__ float pfma(float a, float b, float c) { return fmaf(a, b, c); }

// complex value over either real type
template <class R>
struct cx {
  R x, y;
};

}  // namespace tac
