// f32x2.cuh -- Blackwell packed-pair FP32 arithmetic: fma.rn.f32x2 does two FMAs per lane per instruction (same FMA-pipe
// rate as FFMA, half the issue slots; measured, tools/microbench/ffma2.cu).  ptxas folds a broadcast operand {a, a} into the
// .F32 form and a swapped / negated pair {hi, -lo} into the .LO_HI.NP operand modifier, so neither costs an instruction.
#pragma once

namespace doa {
namespace {

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

}  // namespace
}  // namespace doa
