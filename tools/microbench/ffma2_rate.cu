// Issue rate of packed / scalar FP32 instructions on sm_100a by operand form: is a 3-register-operand FFMA2 limited by
// register-file read ports (even / odd banks) rather than by the FMA pipe?   nvcc -arch=sm_100a -O3 ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fmas(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int CH = 8, IT = 512;
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long* cyc, float seed, int* smid)
{
    f2 d[CH], a[CH], b[CH];
    float s[CH];
    for (int i = 0; i < CH; ++i) {
        d[i] = pk(threadIdx.x * 1e-3f + i, seed);
        a[i] = pk(0.999f + i * 1e-5f * seed, 0.998f);
        b[i] = pk(1e-4f * seed + i, 2e-4f);
        s[i] = 0.5f + seed * i;
    }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < IT; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (MODE == 0) d[i] = fma2(a[i], b[i], d[i]);                              // 3 distinct register pairs
                if (MODE == 1) d[i] = fma2(a[i], pk(0.70710678f, 0.70710678f), d[i]);      // 2 pairs + immediate
                if (MODE == 2) d[i] = fma2(a[i], pk(s[i], s[i]), d[i]);                    // 2 pairs + scalar register
                if (MODE == 3) d[i] = add2(a[i], d[i]);                                    // FADD2: 2 pairs
                if (MODE == 4) d[i] = mul2(d[i], pk(s[i], s[i]));                          // FMUL2: pair x scalar
                if (MODE == 5) d[i] = fma2(a[0], b[0], d[i]);                              // same a, b in every instruction (reuse)
                if (MODE == 6) d[i] = fma2(a[i], pk(s[0], s[0]), d[i]);                    // scalar shared by all (reuse)
                if (MODE == 7) { float x = fmas(lo(a[i]), lo(b[i]), lo(d[i])); d[i] = pk(x, x); }   // scalar FFMA, 3 registers
                if (MODE == 8) { float x = fmas(lo(d[i]), 0.999f, lo(b[i])); d[i] = pk(x, x); }     // scalar FFMA, 2 registers + imm
                if (MODE == 9) d[i] = fma2(d[i], d[i], a[i]);                              // (pair, same pair, pair): 2 distinct
            }
        }
    }
    long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < CH; ++i) acc += lo(d[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) {
        unsigned id;
        asm("mov.u32 %0, %%smid;" : "=r"(id));
        cyc[2 * blockIdx.x] = t0;
        cyc[2 * blockIdx.x + 1] = t1;
        smid[blockIdx.x] = (int)id;
    }
}

template <int MODE>
void run(const char* name, int ctas_per_sm)
{
    const int sms = 148, n = sms * ctas_per_sm;
    float* out; long long* cyc; int* smid;
    cudaMalloc(&out, sizeof(float) * n * 256);
    cudaMalloc(&cyc, sizeof(long long) * 2 * n);
    cudaMalloc(&smid, sizeof(int) * n);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<n, 256>>>(out, cyc, 1.0f, smid);
    cudaEventRecord(e0);
    k<MODE><<<n, 256>>>(out, cyc, 1.0f, smid);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    static long long h[2 * 148 * 8]; static int hs[148 * 8];
    cudaMemcpy(h, cyc, sizeof(long long) * 2 * n, cudaMemcpyDeviceToHost);
    cudaMemcpy(hs, smid, sizeof(int) * n, cudaMemcpyDeviceToHost);
    // per SM: span of its blocks and how many it ran
    static long long lo_[256], hi_[256]; static int cnt[256];
    for (int i = 0; i < 256; ++i) { lo_[i] = 1LL << 62; hi_[i] = 0; cnt[i] = 0; }
    for (int i = 0; i < n; ++i) { int s = hs[i] & 255; if (h[2*i] < lo_[s]) lo_[s] = h[2*i]; if (h[2*i+1] > hi_[s]) hi_[s] = h[2*i+1]; cnt[s]++; }
    double worst = 0; int used = 0, maxc = 0;
    for (int s = 0; s < 256; ++s) if (cnt[s]) {
        const double inst = cnt[s] * 8 / 4.0 * IT * 4.0 * CH;
        const double c = (hi_[s] - lo_[s]) / inst;
        if (c > worst) worst = c;
        used++; if (cnt[s] > maxc) maxc = cnt[s];
    }
    const double inst_total = (double)n * 8 * IT * 4.0 * CH;            // warp-instructions
    printf("%-44s %d CTAs/SM (%d SMs used, max %d per SM): %.2f clk64 / warp-inst / SMSP;  wall: %.3f warp-inst per ns per SMSP\n",
           name, ctas_per_sm, used, maxc, worst, inst_total / (ms * 1e-3) / 1e9 / (148 * 4));
    cudaFree(out); cudaFree(cyc); cudaFree(smid);
}

int main()
{
    for (int c = 2; c <= 4; c += 2) {
        run<0>("FFMA2  pair, pair, pair (3 distinct)", c);
        run<1>("FFMA2  pair, imm, pair", c);
        run<2>("FFMA2  pair, scalar reg, pair", c);
        run<3>("FADD2  pair, pair", c);
        run<4>("FMUL2  pair, scalar reg", c);
        run<5>("FFMA2  same a, b every time (reuse)", c);
        run<6>("FFMA2  pair, shared scalar, pair (reuse)", c);
        run<7>("FFMA   3 registers", c);
        run<8>("FFMA   2 registers + imm", c);
        run<9>("FFMA2  d*d + a (2 distinct pairs)", c);
    }
    return 0;
}
