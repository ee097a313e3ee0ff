// Packed FP32 pairs for sm_100a (PTX ISA 8.6 add/sub/mul/fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2).
// A value of type f2 is a 64-bit register pair (lo, hi); complex numbers are held as (re, im).
// ptxas folds half swaps and per-half negations that are written with mov.b64 pack/unpack into the operand
// modifiers of the packed instruction (.LO_HI / .NP ...), so they cost no instructions.
// One packed instruction does the work of two scalar ones in ONE issue slot -- the arithmetic (and rounding)
// is identical to the scalar form.
#pragma once

typedef unsigned long long f2;

__device__ __forceinline__ f2 pk(float lo, float hi)
{
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo(f2 v)
{
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    (void)b;
    return a;
}
__device__ __forceinline__ float hi(f2 v)
{
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    (void)a;
    return b;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b)
{
    f2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b)
{
    f2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b)
{
    f2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c)
{
    f2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// (lo, hi) -> (hi, -lo): multiplication of a complex number by -i
__device__ __forceinline__ f2 mul_mi(f2 v) { return pk(hi(v), -lo(v)); }
// (lo, hi) -> (-hi, lo): multiplication of a complex number by +i
__device__ __forceinline__ f2 mul_i(f2 v) { return pk(-hi(v), lo(v)); }
// scalar broadcast (s, s): ptxas encodes it as a .F32 operand, no instruction
__device__ __forceinline__ f2 bc(float s) { return pk(s, s); }
// (lo, hi) -> (hi, lo)
__device__ __forceinline__ f2 swp(f2 v) { return pk(hi(v), lo(v)); }
// (lo, hi) -> (-lo, -hi)
__device__ __forceinline__ f2 neg2(f2 v) { return pk(-lo(v), -hi(v)); }
