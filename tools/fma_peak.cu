// Measures the fp32 FFMA ceiling of the device (register-only dependent chains, 8 independent accumulators
// per thread) -- the compute roofline the fused RGL kernels are bound by.  nvcc -arch=sm_100a -O3 fma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) fma_kernel(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wpsm : {4, 8, 16, 32}) {
        const int blocks = sms * wpsm / 8, iters = 20000;
        fma_kernel<<<blocks, 256>>>(out, 1000, 1.0001f, 0.5f);
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            fma_kernel<<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double fma = (double)blocks * 256 * iters * 128.0;
        printf("warps/SM %2d: %.3f ms  %.2f TFMA/s  = %.2f TFLOP/s fp32  (%.1f FMA/clk/SM at 1.965 GHz)\n", wpsm, best,
               fma / best / 1e9, 2 * fma / best / 1e9, fma / (best * 1e-3) / sms / 1.965e9);
    }
    return 0;
}
