// Stand-alone GCN layer for sm_100a: Hout = relu((A X) W) (+ X), A given or A = softmax(X w_a X^T)
// (crowd_nav/policy/graph_model.py:119-128 and :65-66) on node features that already live in HBM.
// This is the unit whose roofline is HBM (1 536 B of compulsory traffic per 6-node state for 14.6 kFLOP,
// SURVEY.md 8(d)); the fused graph_forward kernel keeps the same math on-chip.
//
// Rows are state-major (row = state*n + node), i.e. the HBM order: a tile of TS states is one contiguous
// block of TS*n 128-byte rows, landed in padded (stride-36) shared memory by one TMA bulk copy per row.
#include "kernels.h"

namespace rgl {

template <int TS>
__global__ void __launch_bounds__(TS == 32 ? 192 : 384, TS == 32 ? 2 : 1) gcn_layer_kernel(const float* __restrict__ X, const float* __restrict__ Ag,
                                                           const float* __restrict__ Wg, const float* __restrict__ wag,
                                                           int B, int n, int flags, float* __restrict__ Hout,
                                                           float* __restrict__ Aout, int ntiles) {
    // TS = 32: 32-row register tiles (RT = 4, least shared-memory traffic per FFMA), one warp per 32 rows, two CTAs
    // per SM so that the load / compute / store phases of different tiles overlap; TS = 16: 16-row tiles.
    constexpr int RT = TS == 32 ? 4 : 2, RB = 8 * RT;
    extern __shared__ __align__(128) float smem[];
    const int R = n * TS;
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem);
    uint64_t* bar_in = bar_w + 1;
    float* W = smem + 4;                 // [32][32]
    float* WA = W + XD * XD;             // [32][32]
    float* XB = WA + XD * XD;            // [R][36]
    float* YB = XB + R * LDX;            // [R][36]
    float* AB = YB + R * LDX;            // [TS][n][n]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int rg = lane & 7, cg = lane >> 3;
    const bool skip = flags & RGL_FLAG_SKIP;

    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_in, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(bar_w, (wag ? 2u : 1u) * XD * XD * 4u);
        bulk_g2s(W, Wg, XD * XD * 4u, bar_w);
        if (wag) bulk_g2s(WA, wag, XD * XD * 4u, bar_w);
    }
    mbar_wait(bar_w, 0);

    uint32_t in_parity = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int s0 = tile * TS, cnt = min(TS, B - s0), rows = cnt * n;
        __syncthreads();                 // previous tile fully consumed (XB/YB/AB reuse)
        // ---- stage X rows (and A) ----
        if (tid == 0) {
            fence_proxy_async();
            mbar_arrive_expect_tx(bar_in, (uint32_t)rows * XD * 4u);
        }
        __syncthreads();
        for (int r = tid; r < rows; r += blockDim.x)
            bulk_g2s(XB + r * LDX, X + ((size_t)s0 * n + r) * XD, XD * 4u, bar_in);
        for (int r = rows + tid; r < R; r += blockDim.x)         // zero rows of a partial tile
            for (int c = 0; c < XD; ++c) XB[r * LDX + c] = 0.f;
        if (Ag) {
            const int na = cnt * n * n;
            for (int idx = tid; idx < TS * n * n; idx += blockDim.x)
                AB[idx] = idx < na ? __ldg(Ag + (size_t)s0 * n * n + idx) : 0.f;
        }
        mbar_wait(bar_in, in_parity);
        in_parity ^= 1;
        __syncthreads();

        if (!Ag) {
            // Y = X w_a
            for (int rb = warp; rb * RB < R; rb += nwarps) {
                float acc[RT][8];
#pragma unroll
                for (int q = 0; q < RT; ++q)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[q][c] = 0.f;
                tile_gemm_pf<RT, 2, XD>(acc, XB + (rb * RB + rg) * LDX, LDX, WA + cg * 4, XD);
#pragma unroll
                for (int q = 0; q < RT; ++q)
#pragma unroll
                    for (int m = 0; m < 2; ++m)
                        sts128(YB + (rb * RB + rg + 8 * q) * LDX + cg * 4 + 16 * m,
                               make_float4(acc[q][4 * m], acc[q][4 * m + 1], acc[q][4 * m + 2], acc[q][4 * m + 3]));
            }
            __syncthreads();
            // similarity row + softmax
            for (int r = tid; r < R; r += blockDim.x) {
                const int s = r / n;
                float4 y[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) y[c] = lds128(YB + r * LDX + 4 * c);
                float* arow = AB + r * n;
                float mx = -INFINITY;
                for (int j = 0; j < n; ++j) {
                    const float* x = XB + (s * n + j) * LDX;
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 xv = lds128(x + 4 * c);
                        d0 = fmaf(y[c].x, xv.x, d0); d1 = fmaf(y[c].y, xv.y, d1);
                        d2 = fmaf(y[c].z, xv.z, d2); d3 = fmaf(y[c].w, xv.w, d3);
                    }
                    const float d = (d0 + d1) + (d2 + d3);
                    arow[j] = d;
                    mx = fmaxf(mx, d);
                }
                float sum = 0.f;
                for (int j = 0; j < n; ++j) {
                    const float e = expf(arow[j] - mx);
                    arow[j] = e;
                    sum += e;
                }
                for (int j = 0; j < n; ++j) arow[j] = arow[j] / sum;
            }
            __syncthreads();
        }
        if (Aout) {
            const int na = cnt * n * n;
            for (int idx = tid; idx < na; idx += blockDim.x) Aout[(size_t)s0 * n * n + idx] = AB[idx];
        }
        // AH = A . X (per state)
        for (int r = tid; r < R; r += blockDim.x) {
            const int s = r / n;
            float acc[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = 0.f;
            const float* arow = AB + r * n;
            for (int j = 0; j < n; ++j) {
                const float aij = arow[j];
                const float* h = XB + (s * n + j) * LDX;
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 hv = lds128(h + 4 * c4);
                    acc[4 * c4 + 0] = fmaf(aij, hv.x, acc[4 * c4 + 0]);
                    acc[4 * c4 + 1] = fmaf(aij, hv.y, acc[4 * c4 + 1]);
                    acc[4 * c4 + 2] = fmaf(aij, hv.z, acc[4 * c4 + 2]);
                    acc[4 * c4 + 3] = fmaf(aij, hv.w, acc[4 * c4 + 3]);
                }
            }
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4)
                sts128(YB + r * LDX + 4 * c4, make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]));
        }
        __syncthreads();
        // H' = relu(AH W) (+ X) -> HBM
        for (int rb = warp; rb * RB < R; rb += nwarps) {
            float acc[RT][8];
#pragma unroll
            for (int q = 0; q < RT; ++q)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[q][c] = 0.f;
            tile_gemm_pf<RT, 2, XD>(acc, YB + (rb * RB + rg) * LDX, LDX, W + cg * 4, XD);
#pragma unroll
            for (int q = 0; q < RT; ++q) {
                const int r = rb * RB + rg + 8 * q;
                if (r < rows) {
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        float4 v = make_float4(fmaxf(acc[q][4 * m], 0.f), fmaxf(acc[q][4 * m + 1], 0.f),
                                               fmaxf(acc[q][4 * m + 2], 0.f), fmaxf(acc[q][4 * m + 3], 0.f));
                        if (skip) {
                            const float4 h = lds128(XB + r * LDX + cg * 4 + 16 * m);
                            v.x += h.x; v.y += h.y; v.z += h.z; v.w += h.w;
                        }
                        *reinterpret_cast<float4*>(Hout + ((size_t)s0 * n + r) * XD + cg * 4 + 16 * m) = v;
                    }
                }
            }
        }
    }
}

static size_t gcn_smem_bytes(int TS, int n) {
    return (4 + 2 * XD * XD + 2 * (size_t)n * TS * LDX + (size_t)TS * n * n) * sizeof(float);
}

template <int TS>
static cudaError_t launch_gcn(const float* X, const float* A, const float* W, const float* wa, int B, int n, int flags,
                              float* Hout, float* Aout, int num_sms, size_t max_smem, cudaStream_t st) {
    const size_t smem = gcn_smem_bytes(TS, n);
    if (smem > max_smem) return cudaErrorInvalidConfiguration;
    if (cudaError_t e = ensure_dyn_smem(gcn_layer_kernel<TS>, (int)max_smem)) return e;
    const int ntiles = (B + TS - 1) / TS;
    int nwarps = TS == 32 ? n : n * TS / 16;            // one warp per 32-row (TS = 32) / 16-row block
    const int cap = TS == 32 ? 6 : 12;
    if (nwarps > cap) nwarps = cap;
    int per_sm = (int)((228 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    const int max_cta = TS == 32 ? 2 : 1;               // register budget of the kernel's __launch_bounds__
    if (per_sm > max_cta) per_sm = max_cta;
    const int grid = ntiles < num_sms * per_sm ? ntiles : num_sms * per_sm;
    gcn_layer_kernel<TS><<<grid, nwarps * 32, smem, st>>>(X, A, W, wa, B, n, flags, Hout, Aout, ntiles);
    return cudaGetLastError();
}

cudaError_t run_gcn_layer(const float* X, const float* A, const float* W, const float* wa, int B, int n, int flags,
                          float* Hout, float* Aout, int num_sms, size_t max_smem, cudaStream_t st) {
    if (gcn_smem_bytes(32, n) <= max_smem / 2 || gcn_smem_bytes(16, n) > max_smem)
        return launch_gcn<32>(X, A, W, wa, B, n, flags, Hout, Aout, num_sms, max_smem, st);
    return launch_gcn<16>(X, A, W, wa, B, n, flags, Hout, Aout, num_sms, max_smem, st);
}

}  // namespace rgl
