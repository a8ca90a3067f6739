// Microbenchmarks: (1) throughput of legacy warp-level mma.sync.m16n8k8 TF32 on this GPU, (2) accuracy of the
// 3xTF32 split (hi/lo) product against fp32 FFMA and fp64 on a 16x32x32 tile.   nvcc -arch=sm_100a -O3 mma_peak.cu
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }

__global__ void __launch_bounds__(256) mma_kernel(float* out, int iters) {
    float d[8][4];
    uint32_t a[4], b[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    for (int j = 0; j < 4; ++j) a[j] = to_tf32(1.0f + threadIdx.x * 1e-3f + j);
    b[0] = to_tf32(0.5f); b[1] = to_tf32(0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) mma_tf32(d[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    if (s == 123.456f) out[0] = s;
}

// accuracy: D[16x32] = X[16x32] * W[32x32]; one warp; compare fp32 FMA, 1xTF32, 3xTF32 against fp64
__global__ void acc_kernel(const float* X, const float* W, float* D1, float* D3) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    for (int nt = 0; nt < 4; ++nt) {
        float d1[4] = {0, 0, 0, 0}, d3[4] = {0, 0, 0, 0};
        for (int ks = 0; ks < 4; ++ks) {
            float af[4] = {X[g * 32 + ks * 8 + t], X[(g + 8) * 32 + ks * 8 + t], X[g * 32 + ks * 8 + t + 4], X[(g + 8) * 32 + ks * 8 + t + 4]};
            float bf[2] = {W[(ks * 8 + t) * 32 + nt * 8 + g], W[(ks * 8 + t + 4) * 32 + nt * 8 + g]};
            uint32_t ah[4], al[4], bh[2], bl[2];
            for (int j = 0; j < 4; ++j) { ah[j] = to_tf32(af[j]); al[j] = to_tf32(af[j] - __uint_as_float(ah[j])); }
            for (int j = 0; j < 2; ++j) { bh[j] = to_tf32(bf[j]); bl[j] = to_tf32(bf[j] - __uint_as_float(bh[j])); }
            mma_tf32(d1, ah, bh);
            mma_tf32(d3, al, bh);      // small terms first
            mma_tf32(d3, ah, bl);
            mma_tf32(d3, ah, bh);
        }
        const int c = nt * 8 + 2 * t;
        D1[g * 32 + c] = d1[0]; D1[g * 32 + c + 1] = d1[1]; D1[(g + 8) * 32 + c] = d1[2]; D1[(g + 8) * 32 + c + 1] = d1[3];
        D3[g * 32 + c] = d3[0]; D3[g * 32 + c + 1] = d3[1]; D3[(g + 8) * 32 + c] = d3[2]; D3[(g + 8) * 32 + c + 1] = d3[3];
    }
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wpsm : {4, 8, 16, 32}) {
        const int blocks = sms * wpsm / 8, iters = 4000;
        mma_kernel<<<blocks, 256>>>(out, 100);
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            mma_kernel<<<blocks, 256>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double macs = (double)blocks * 8 * iters * 8 * 1024.0;     // warps * iters * 8 mma * (16*8*8)
        printf("mma.sync m16n8k8 tf32, warps/SM %2d: %.3f ms  %.1f TMAC/s = %.1f TFLOP/s  (%.0f MAC/clk/SM at 1.965 GHz)\n", wpsm, best,
               macs / best / 1e9, 2 * macs / best / 1e9, macs / (best * 1e-3) / sms / 1.965e9);
    }
    // accuracy
    float hX[512], hW[1024], *dX, *dW, *dD1, *dD3, hD1[512], hD3[512];
    srand(1);
    for (int i = 0; i < 512; ++i) hX[i] = 20.0f * ((float)rand() / RAND_MAX) * ((rand() & 1) ? 1.f : 0.3f);   // relu-like activations up to 20
    for (int i = 0; i < 1024; ++i) { float u = (float)rand() / RAND_MAX, v = (float)rand() / RAND_MAX; hW[i] = sqrtf(-2 * logf(u + 1e-9f)) * cosf(6.2831853f * v); }
    cudaMalloc(&dX, 2048); cudaMalloc(&dW, 4096); cudaMalloc(&dD1, 2048); cudaMalloc(&dD3, 2048);
    cudaMemcpy(dX, hX, 2048, cudaMemcpyHostToDevice); cudaMemcpy(dW, hW, 4096, cudaMemcpyHostToDevice);
    acc_kernel<<<1, 32>>>(dX, dW, dD1, dD3);
    cudaMemcpy(hD1, dD1, 2048, cudaMemcpyDeviceToHost); cudaMemcpy(hD3, dD3, 2048, cudaMemcpyDeviceToHost);
    double e1x = 0, e3x = 0, ef = 0, scale = 0;
    for (int r = 0; r < 16; ++r) for (int c = 0; c < 32; ++c) {
        double ref = 0; float f = 0.f;
        for (int k = 0; k < 32; ++k) { ref += (double)hX[r * 32 + k] * hW[k * 32 + c]; f = fmaf(hX[r * 32 + k], hW[k * 32 + c], f); }
        e1x = fmax(e1x, fabs(hD1[r * 32 + c] - ref)); e3x = fmax(e3x, fabs(hD3[r * 32 + c] - ref)); ef = fmax(ef, fabs(f - ref));
        scale = fmax(scale, fabs(ref));
    }
    printf("accuracy vs fp64 at scale %.1f: fp32 FMA %.3e (rel %.2e) | 1xTF32 %.3e (rel %.2e) | 3xTF32 %.3e (rel %.2e)\n", scale, ef, ef / scale,
           e1x, e1x / scale, e3x, e3x / scale);
    return 0;
}
