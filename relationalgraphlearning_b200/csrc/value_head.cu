// Value head for sm_100a: V[B] = mlp(32,[32,100,100,1])(E[B,32])  (crowd_nav/policy/value_estimator.py:9,19).
// Persistent CTAs, the 73 KB packed weight blob resident in shared memory (one TMA bulk copy), 32-row tiles,
// every layer a set of (16-row block) x (32-column block) register tiles spread over the 8 warps.
#include "kernels.h"

namespace rgl {

constexpr int VT = 32;            // rows (states) per tile
constexpr int LDV = VHP + 4;      // 132: padded stride of the 128-wide hidden rows

__global__ void __launch_bounds__(256, 1) value_head_kernel(const float* __restrict__ E, int B, const float* __restrict__ vwg,
                                                            float* __restrict__ V, float* __restrict__ sv0, float* __restrict__ sv1,
                                                            float* __restrict__ sv2, int ntiles, int use_tma) {
    constexpr int RT = 2, RB = 16;
    extern __shared__ __align__(128) float smem[];
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem);
    uint64_t* bar_in = bar_w + 1;
    float* vw = smem + 4;
    float* X0 = vw + VALUE_FLOATS;     // [32][36] input rows
    float* X1 = X0 + VT * LDX;         // [32][36]
    float* H1 = X1 + VT * LDX;         // [32][132]
    float* H2 = H1 + VT * LDV;         // [32][132]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int rg = lane & 7, cg = lane >> 3;

    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_in, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(bar_w, VALUE_FLOATS * 4u);
        bulk_g2s(vw, vwg, VALUE_FLOATS * 4u, bar_w);
    }
    auto load_tile = [&](int t) {
        const int s0 = t * VT, cnt = min(VT, B - s0);
        if (use_tma && cnt == VT) {
            if (warp == 0) {                       // one 128-byte bulk copy per row into the padded layout
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_arrive_expect_tx(bar_in, VT * XD * 4u);
                }
                __syncwarp();
                bulk_g2s(X0 + lane * LDX, E + (size_t)(s0 + lane) * XD, XD * 4u, bar_in);
            }
        } else {
            for (int idx = tid; idx < VT * XD; idx += blockDim.x) {
                const int s = idx / XD, c = idx - s * XD;
                X0[s * LDX + c] = s < cnt ? __ldg(E + (size_t)s0 * XD + idx) : 0.f;
            }
        }
    };
    uint32_t in_parity = 0;
    if ((int)blockIdx.x < ntiles) load_tile(blockIdx.x);
    mbar_wait(bar_w, 0);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int s0 = tile * VT, cnt = min(VT, B - s0);
        if (use_tma && cnt == VT) {
            mbar_wait(bar_in, in_parity);
            in_parity ^= 1;
        }
        __syncthreads();

        // layer 0: 32 -> 32, relu
        for (int rb = warp; rb < VT / RB; rb += nwarps) {
            float acc[RT][8];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const float4 b = lds128(vw + V_B0 + cg * 4 + 16 * m);
#pragma unroll
                for (int q = 0; q < RT; ++q) { acc[q][4 * m] = b.x; acc[q][4 * m + 1] = b.y; acc[q][4 * m + 2] = b.z; acc[q][4 * m + 3] = b.w; }
            }
            tile_gemm<RT, 2>(acc, X0 + (rb * RB + rg) * LDX, LDX, vw + V_W0 + cg * 4, XD, XD);
#pragma unroll
            for (int q = 0; q < RT; ++q)
#pragma unroll
                for (int m = 0; m < 2; ++m)
                    sts128(X1 + (rb * RB + rg + 8 * q) * LDX + cg * 4 + 16 * m,
                           make_float4(fmaxf(acc[q][4 * m], 0.f), fmaxf(acc[q][4 * m + 1], 0.f),
                                       fmaxf(acc[q][4 * m + 2], 0.f), fmaxf(acc[q][4 * m + 3], 0.f)));
        }
        __syncthreads();
        if (sv0) {                                   // training forward: save relu(layer 0)
            for (int idx = tid; idx < VT * XD; idx += blockDim.x) {
                const int s = idx / XD, c = idx - s * XD;
                if (s < cnt) sv0[(size_t)(s0 + s) * XD + c] = X1[s * LDX + c];
            }
        }
        // X0 is dead: prefetch the next tile's rows
        if (tile + (int)gridDim.x < ntiles) load_tile(tile + gridDim.x);

        // layer 1: 32 -> 100 (padded 128), relu.  items = (row block, 32-column block)
        for (int it = warp; it < (VT / RB) * 4; it += nwarps) {
            const int rb = it >> 2, cb = it & 3;
            float acc[RT][8];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const float4 b = lds128(vw + V_B1 + cb * 32 + cg * 4 + 16 * m);
#pragma unroll
                for (int q = 0; q < RT; ++q) { acc[q][4 * m] = b.x; acc[q][4 * m + 1] = b.y; acc[q][4 * m + 2] = b.z; acc[q][4 * m + 3] = b.w; }
            }
            tile_gemm<RT, 2>(acc, X1 + (rb * RB + rg) * LDX, LDX, vw + V_W1 + cb * 32 + cg * 4, VHP, XD);
#pragma unroll
            for (int q = 0; q < RT; ++q)
#pragma unroll
                for (int m = 0; m < 2; ++m)
                    sts128(H1 + (rb * RB + rg + 8 * q) * LDV + cb * 32 + cg * 4 + 16 * m,
                           make_float4(fmaxf(acc[q][4 * m], 0.f), fmaxf(acc[q][4 * m + 1], 0.f),
                                       fmaxf(acc[q][4 * m + 2], 0.f), fmaxf(acc[q][4 * m + 3], 0.f)));
        }
        __syncthreads();
        if (sv1) {
            for (int idx = tid; idx < VT * VHP; idx += blockDim.x) {
                const int s = idx / VHP, c = idx - s * VHP;
                if (s < cnt) sv1[(size_t)(s0 + s) * VHP + c] = H1[s * LDV + c];
            }
        }
        // layer 2: 100 -> 100 (padded 128), relu
        for (int it = warp; it < (VT / RB) * 4; it += nwarps) {
            const int rb = it >> 2, cb = it & 3;
            float acc[RT][8];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const float4 b = lds128(vw + V_B2 + cb * 32 + cg * 4 + 16 * m);
#pragma unroll
                for (int q = 0; q < RT; ++q) { acc[q][4 * m] = b.x; acc[q][4 * m + 1] = b.y; acc[q][4 * m + 2] = b.z; acc[q][4 * m + 3] = b.w; }
            }
            tile_gemm<RT, 2>(acc, H1 + (rb * RB + rg) * LDV, LDV, vw + V_W2 + cb * 32 + cg * 4, VHP, VH);
#pragma unroll
            for (int q = 0; q < RT; ++q)
#pragma unroll
                for (int m = 0; m < 2; ++m)
                    sts128(H2 + (rb * RB + rg + 8 * q) * LDV + cb * 32 + cg * 4 + 16 * m,
                           make_float4(fmaxf(acc[q][4 * m], 0.f), fmaxf(acc[q][4 * m + 1], 0.f),
                                       fmaxf(acc[q][4 * m + 2], 0.f), fmaxf(acc[q][4 * m + 3], 0.f)));
        }
        __syncthreads();
        if (sv2) {
            for (int idx = tid; idx < VT * VHP; idx += blockDim.x) {
                const int s = idx / VHP, c = idx - s * VHP;
                if (s < cnt) sv2[(size_t)(s0 + s) * VHP + c] = H2[s * LDV + c];
            }
        }
        // layer 3: 100 -> 1.  8 lanes per row, 16 (padded) k each, butterfly reduce.
        {
            const int row = tid >> 3, part = tid & 7;          // 256 threads = 32 rows x 8
            const float* h = H2 + row * LDV + part * 16;
            const float* w = vw + V_W3 + part * 16;
            float d = 0.f;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                const float4 hv = lds128(h + 4 * k4), wv = lds128(w + 4 * k4);
                d = fmaf(hv.x, wv.x, d); d = fmaf(hv.y, wv.y, d); d = fmaf(hv.z, wv.z, d); d = fmaf(hv.w, wv.w, d);
            }
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            d += __shfl_xor_sync(0xffffffffu, d, 4);
            if (part == 0 && row < cnt) V[s0 + row] = d + vw[V_B3];
        }
        // no trailing barrier needed: the next iteration's first __syncthreads orders H2 reads before
        // layer-2 rewrites (two barriers later) and X0 was refilled after its last reader.
    }
}

size_t value_smem_bytes() { return (4 + VALUE_FLOATS + 2 * VT * LDX + 2 * VT * LDV) * sizeof(float); }

cudaError_t run_value_head(const float* E, int B, const float* vw, float* V, float* v0, float* v1, float* v2, int use_tma, int num_sms,
                           cudaStream_t st) {
    const size_t smem = value_smem_bytes();
    if (cudaError_t e = ensure_dyn_smem(value_head_kernel, (int)smem)) return e;
    const int ntiles = (B + VT - 1) / VT;
    const int grid = ntiles < num_sms ? ntiles : num_sms;
    value_head_kernel<<<grid, 256, smem, st>>>(E, B, vw, V, v0, v1, v2, ntiles, use_tma);
    return cudaGetLastError();
}

}  // namespace rgl
