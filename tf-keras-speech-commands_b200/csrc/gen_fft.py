#!/usr/bin/env python3
"""Generates fft_gen.cuh: straight-line in-register complex FFTs (N = 8, 16, 32) for sm_100a.

The header holds the PACKED forms (emit_fft_packed: f32x2 instructions on (re, im) register pairs).  The scalar
emitters (emit_fft, emit_fft_tw) are kept as the readable derivation of the same butterflies and are what the packed
forms were checked against; they are not emitted any more.


Radix-2 decimation-in-time, fully unrolled, twiddles as float literals.  Input is expected in
bit-reversed order (the caller permutes for free: every index is a compile-time constant, so the
arrays live in registers); output is in natural order.

Butterfly forms (a' = a + w*b, b' = a - w*b):
  w = 1         4 FADD
  w = -i        4 FADD
  general       6 FFMA:   a'r = fma(wr, br, fma(-wi, bi, ar));  a'i = fma(wr, bi, fma(wi, br, ai))
                          b'  = 2a - a'   (one FFMA per component)
N = 32: 46 trivial + 34 general butterflies = 388 FP32 instructions (nominal 5 N log2 N = 800 flop).

Run:  python gen_fft.py > fft_gen.cuh
"""
import math


def bitrev(i, bits):
    r = 0
    for b in range(bits):
        r = (r << 1) | ((i >> b) & 1)
    return r


def lit(x):
    s = repr(float(x))
    if 'e' not in s and '.' not in s:
        s += '.0'
    return s + 'f'


def emit_fft(n):
    bits = n.bit_length() - 1
    out = []
    out.append('// %d-point complex FFT, forward (e^{-2*pi*i*k*n/N}), in place.' % n)
    out.append('// In: xr/xi[i] = z[bitrev%d(i)].  Out: xr/xi[k] = Z[k].' % bits)
    out.append('__device__ __forceinline__ void fft%d_dit(float (&xr)[%d], float (&xi)[%d])' % (n, n, n))
    out.append('{')
    out.append('    float ar, ai, br, bi;')
    span = 1
    n_triv = n_gen = 0
    while span < n:
        out.append('    // ---- stage span=%d' % span)
        for g in range(0, n, 2 * span):
            for k in range(span):
                i, j = g + k, g + k + span
                if k == 0:
                    n_triv += 1
                    out.append('    ar = xr[%d]; ai = xi[%d]; br = xr[%d]; bi = xi[%d];' % (i, i, j, j))
                    out.append('    xr[%d] = ar + br; xi[%d] = ai + bi; xr[%d] = ar - br; xi[%d] = ai - bi;' % (i, i, j, j))
                elif 2 * k == span:
                    n_triv += 1    # w = -i : w*b = (bi, -br)
                    out.append('    ar = xr[%d]; ai = xi[%d]; br = xr[%d]; bi = xi[%d];' % (i, i, j, j))
                    out.append('    xr[%d] = ar + bi; xi[%d] = ai - br; xr[%d] = ar - bi; xi[%d] = ai + br;' % (i, i, j, j))
                else:
                    n_gen += 1
                    ang = -2.0 * math.pi * k / (2 * span)
                    wr, wi = math.cos(ang), math.sin(ang)
                    out.append('    ar = xr[%d]; ai = xi[%d]; br = xr[%d]; bi = xi[%d];' % (i, i, j, j))
                    out.append('    xr[%d] = __fmaf_rn(%s, br, __fmaf_rn(%s, bi, ar)); xi[%d] = __fmaf_rn(%s, bi, __fmaf_rn(%s, br, ai));'
                               % (i, lit(wr), lit(-wi), i, lit(wr), lit(wi)))
                    out.append('    xr[%d] = __fmaf_rn(2.0f, ar, -xr[%d]); xi[%d] = __fmaf_rn(2.0f, ai, -xi[%d]);' % (j, i, j, i))
        span *= 2
    out.append('}')
    out.insert(3, '// %d trivial (4 FADD) + %d general (6 FFMA) butterflies = %d FP32 instructions'
               % (n_triv, n_gen, 4 * n_triv + 6 * n_gen))
    return '\n'.join(out)


def emit_fft_tw(n):
    """Variant whose first stage also applies per-element input twiddles (the inter-pass twiddles of a two-pass
    FFT): x = z * c before the butterfly.  Fused form per first-stage butterfly (a = z[n]*c[n], b = z[m]*c[m]):
        p   = z[n] * c[n]                      2 FMUL + 2 FFMA
        a'  = p + z[m] * c[m]                  4 FFMA
        b'  = 2p - a'                          2 FFMA         -> 10 instead of 8 + 4."""
    bits = n.bit_length() - 1
    body = emit_fft(n).split('\n')
    start = next(i for i, l in enumerate(body) if 'stage span=1' in l)
    end = next(i for i, l in enumerate(body) if 'stage span=2' in l)
    out = []
    out.append('// %d-point complex FFT with fused input twiddles: computes FFT of z[n] * c[n].' % n)
    out.append('// In: zr/zi/cr/ci in NATURAL order.  Out: xr/xi[k] = Z[k] in natural order.')
    out.append('__device__ __forceinline__ void fft%d_dit_tw(const float (&zr)[%d], const float (&zi)[%d], const float (&cr)[%d],' % (n, n, n, n))
    out.append('                                             const float (&ci)[%d], float (&xr)[%d], float (&xi)[%d])' % (n, n, n))
    out.append('{')
    out.append('    float ar, ai, br, bi;')
    out.append('    float pr, pi;')
    out.append('    // ---- stage span=1 fused with the input twiddles')
    for i in range(0, n, 2):
        na, nb = bitrev(i, bits), bitrev(i + 1, bits)
        out.append('    pr = __fmaf_rn(zr[%d], cr[%d], -zi[%d] * ci[%d]); pi = __fmaf_rn(zr[%d], ci[%d], zi[%d] * cr[%d]);'
                   % (na, na, na, na, na, na, na, na))
        out.append('    xr[%d] = __fmaf_rn(zr[%d], cr[%d], __fmaf_rn(-zi[%d], ci[%d], pr)); xi[%d] = __fmaf_rn(zr[%d], ci[%d], __fmaf_rn(zi[%d], cr[%d], pi));'
                   % (i, nb, nb, nb, nb, i, nb, nb, nb, nb))
        out.append('    xr[%d] = __fmaf_rn(2.0f, pr, -xr[%d]); xi[%d] = __fmaf_rn(2.0f, pi, -xi[%d]);' % (i + 1, i, i + 1, i))
    out += body[end:]
    return '\n'.join(out)


def emit_fft_packed(n, with_tw=False):
    """Packed (f32x2) form: x[i] is a 64-bit (re, im) pair.  Butterflies (a' = a + w b, b' = a - w b):
        w = 1      a' = add2(a, b); b' = sub2(a, b)                                     2 instructions
        w = -i     t = (b.im, -b.re) [operand modifiers]; a' = add2(a, t); b' = sub2(a, t)   2
        general    a' = fma2(b, wr, a); a' = fma2(i*b, wi, a'); b' = fma2(2, a, -a')   (i*b = (-b.im, b.re) is an operand modifier)   3
    i.e. half the issue slots of the scalar form with identical rounding.
    with_tw: the first stage also multiplies its inputs by per-element twiddles c (natural order arrays z, c):
        p = mul2(z_a, cr); p = fma2(i*z_a, ci, p); a' = fma2(z_b, dr, p); a' = fma2(i*z_b, di, a'); b' = fma2(2, p, -a')
                                                                                        5 instead of 10."""
    bits = n.bit_length() - 1
    out = []
    name = 'fft%d_p2%s' % (n, '_tw' if with_tw else '')
    if with_tw:
        out.append('// %d-point complex FFT of z[n] * c[n], packed f32x2.  In: z, c natural order.  Out: x natural order.' % n)
        out.append('__device__ __forceinline__ void %s(const f2 (&z)[%d], const f2 (&c)[%d], f2 (&x)[%d])' % (name, n, n, n))
    else:
        out.append('// %d-point complex FFT, packed f32x2, in place.  In: x[i] = z[bitrev(i)].  Out: x[k] = Z[k].' % n)
        out.append('__device__ __forceinline__ void %s(f2 (&x)[%d])' % (name, n))
    out.append('{')
    out.append('    f2 a, b, t;')
    count = 0
    span = 1
    first = True
    while span < n:
        out.append('    // ---- stage span=%d' % span)
        for g in range(0, n, 2 * span):
            for k in range(span):
                i, j = g + k, g + k + span
                if first and with_tw:
                    na, nb = bitrev(i, bits), bitrev(j, bits)
                    out.append('    t = mul2(z[%d], bc(lo(c[%d]))); t = fma2(mul_i(z[%d]), bc(hi(c[%d])), t);' % (na, na, na, na))
                    out.append('    a = fma2(z[%d], bc(lo(c[%d])), t); a = fma2(mul_i(z[%d]), bc(hi(c[%d])), a);' % (nb, nb, nb, nb))
                    out.append('    x[%d] = a; x[%d] = fma2(pk(2.0f, 2.0f), t, neg2(a));' % (i, j))
                    count += 5
                elif k == 0:
                    out.append('    a = x[%d]; b = x[%d]; x[%d] = add2(a, b); x[%d] = sub2(a, b);' % (i, j, i, j))
                    count += 2
                elif 2 * k == span:
                    out.append('    a = x[%d]; t = mul_mi(x[%d]); x[%d] = add2(a, t); x[%d] = sub2(a, t);' % (i, j, i, j))
                    count += 2
                else:
                    ang = -2.0 * math.pi * k / (2 * span)
                    wr, wi = math.cos(ang), math.sin(ang)
                    out.append('    a = x[%d]; b = x[%d]; t = fma2(b, bc(%s), a); t = fma2(mul_i(b), bc(%s), t);'
                               % (i, j, lit(wr), lit(wi)))
                    out.append('    x[%d] = t; x[%d] = fma2(pk(2.0f, 2.0f), a, neg2(t));' % (i, j))
                    count += 3
        span *= 2
        first = False
    out.append('}')
    out.insert(2 if not with_tw else 2, '// %d packed instructions' % count)
    return '\n'.join(out)


def main():
    print('// GENERATED by gen_fft.py -- do not edit.  See that file for the derivation.')
    print('#pragma once')
    print()
    print('// bit reversal of i over `bits` bits; folds to a constant in unrolled code')
    print('__host__ __device__ constexpr int scf_bitrev(int i, int bits)')
    print('{')
    print('    int r = 0;')
    print('    for (int b = 0; b < bits; ++b) r = (r << 1) | ((i >> b) & 1);')
    print('    return r;')
    print('}')
    print()
    print('#include "f32x2.cuh"')
    print()
    for n in (8, 16, 32):
        print(emit_fft_packed(n))
        print()
    print(emit_fft_packed(32, with_tw=True))
    print()


if __name__ == '__main__':
    main()
