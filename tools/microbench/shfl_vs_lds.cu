// Micro-benchmark: throughput of SHFL.IDX against STS.64+LDS.64 exchanges on sm_100a, alone and mixed.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o shfl_vs_lds shfl_vs_lds.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(768, 1) k(float* out, int iters)
{
    extern __shared__ float2 sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* w = sm + warp * 32 * 17;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
    const int src = (32 - lane) & 31;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __shfl_sync(0xffffffffu, v[i], src) + 1.f;
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i * 32 + lane] = make_float2(v[2 * i], v[2 * i + 1]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 t = w[i * 32 + src];
                v[2 * i] = t.x + 1.f;
                v[2 * i + 1] = t.y + 1.f;
            }
            __syncwarp();
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* d, int sms)
{
    const int iters = 2000;
    const size_t smem = 24 * 32 * 17 * sizeof(float2);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<MODE><<<sms, 768, smem>>>(d, 10);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms, 768, smem>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * khz * 1e3;
    // per SM: 24 warps x iters x (32 floats exchanged per lane)
    printf("%-28s %.3f ms  -> %.2f SM-cycles per warp-exchange of 32 floats/lane (err %s)\n", name, ms, cyc / (24.0 * iters),
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* d; cudaMalloc(&d, p.multiProcessorCount * 768 * sizeof(float));
    run<0>("32 x SHFL.IDX", d, p.multiProcessorCount);
    run<1>("16 x STS.64 + 16 x LDS.64", d, p.multiProcessorCount);
    run<2>("both", d, p.multiProcessorCount);
    return 0;
}
