// Backward kernels of the value-estimator training step (crowd_nav/utils/trainer.py:122-131 calls loss.backward()
// through ValueEstimator -> RGL; this file is that backward, hand-written).
//
// The forward of a training step is the fused graph_forward / value_head kernels run with activation saves
// (kernels.h GraphSave).  The backward is a short sequence of three generic kernels over those saved rows:
//   rows_linear_bwd   y = x W (+b) layers: data gradient G W^T, weight gradient x^T G (register-tiled, accumulated
//                     per CTA and flushed with one atomicAdd per element), bias gradient; optional relu mask
//   attn_layer_bwd    H' = relu(A H W) + H :  gH += A^T gM,  gA += gM H^T           (per state)
//   sim_bwd           A = softmax(Y X^T):  gS, gY = gS X, gX += gS^T Y            (per state)
#include "kernels.h"

namespace rgl {

// row r of a logical [R, width] matrix: ptr + (r / rpg) * gstride + (r % rpg) * ld   (grouped rows: e.g. the robot
// row of every state inside a [B, n, 32] tensor is rpg = 1, gstride = n*32)
struct Rows {
    float* ptr;
    int ld;
    int rpg;
    long long gstride;
    __device__ __forceinline__ float* row(int r) const { return ptr + (long long)(r / rpg) * gstride + (long long)(r % rpg) * ld; }
};

struct LinBwdArgs {
    Rows G, mask, Xin, Gin;
    int N, K, R;
    const float* W;       // optional (data gradient)
    int w_layout;         // 0: W is [N][K] (nn.Linear.weight), 1: W is [K][N] (w_a / Ws used as x @ W)
    int accumulate;       // Gin += instead of =
    float* dW;            // optional, same layout as W
    float* db;            // optional [N]
    int ntiles;
};

constexpr int LT = 64;    // rows per tile

__global__ void __launch_bounds__(256, 1) rows_linear_bwd_kernel(const LinBwdArgs a) {
    extern __shared__ __align__(128) float smem[];
    const int N = a.N, K = a.K;
    const int NP4 = (N + 3) & ~3, NP32 = (N + 31) & ~31, KP32 = (K + 31) & ~31;
    const int ldw = KP32 + 4, ldg = NP32 + 4, ldx = KP32 + 4;
    float* Ws = smem;                       // [NP4][ldw]   Ws[n][k] = dY_n/dX_k weight
    float* Gs = Ws + NP4 * ldw;             // [LT][ldg]
    float* Xs = Gs + LT * ldg;              // [LT][ldx]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rg = lane & 7, cg = lane >> 3;

    if (a.W) {
        for (int idx = tid; idx < NP4 * ldw; idx += blockDim.x) {
            const int nn = idx / ldw, k = idx - nn * ldw;
            float v = 0.f;
            if (nn < N && k < K) v = a.w_layout == 0 ? a.W[(size_t)nn * K + k] : a.W[(size_t)k * N + nn];
            Ws[idx] = v;
        }
    }
    const int KB = KP32 / 32, NB = NP32 / 32;            // weight-gradient blocks (<= 16 items, 2 per warp)
    float wacc[2][32];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int c = 0; c < 32; ++c) wacc[t][c] = 0.f;
    float bacc = 0.f;

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int r0 = tile * LT;
        __syncthreads();
        // ---- stage G (masked) and Xin, zero padded ----
        for (int idx = tid; idx < LT * ldg; idx += blockDim.x) {
            const int r = idx / ldg, c = idx - r * ldg;
            float v = 0.f;
            if (r0 + r < a.R && c < N) {
                v = a.G.row(r0 + r)[c];
                if (a.mask.ptr && !(a.mask.row(r0 + r)[c] > 0.f)) v = 0.f;
            }
            Gs[idx] = v;
        }
        if (a.Xin.ptr) {
            for (int idx = tid; idx < LT * ldx; idx += blockDim.x) {
                const int r = idx / ldx, c = idx - r * ldx;
                Xs[idx] = (r0 + r < a.R && c < K) ? a.Xin.row(r0 + r)[c] : 0.f;
            }
        }
        __syncthreads();
        // ---- data gradient: Gin[r][k] = sum_n G[r][n] W[n][k] ----
        if (a.W && a.Gin.ptr) {
            for (int it = warp; it < (LT / 16) * KB; it += 8) {
                const int rb = it / KB, cb = it - rb * KB;
                float acc[2][8];
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[q][c] = 0.f;
                tile_gemm<2, 2>(acc, Gs + (rb * 16 + rg) * ldg, ldg, Ws + cb * 32 + cg * 4, ldw, NP4);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int r = r0 + rb * 16 + rg + 8 * q;
                    if (r < a.R) {
                        float* o = a.Gin.row(r);
#pragma unroll
                        for (int m = 0; m < 2; ++m)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int k = cb * 32 + cg * 4 + 16 * m + j;
                                if (k < K) o[k] = a.accumulate ? o[k] + acc[q][4 * m + j] : acc[q][4 * m + j];
                            }
                    }
                }
            }
        }
        // ---- weight gradient: dW[k][n] += sum_r x[r][k] g[r][n]; thread tile 4 k x 8 n, rows streamed ----
        if (a.dW) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int it = warp + 8 * t;
                if (it < KB * NB) {
                    const int kb = it / NB, nb = it - kb * NB;
                    const float* xp = Xs + kb * 32 + rg * 4;
                    const float* gp = Gs + nb * 32 + cg * 8;
#pragma unroll 4
                    for (int r = 0; r < LT; ++r) {
                        const float4 xv = lds128(xp + r * ldx);
                        const float4 g0 = lds128(gp + r * ldg), g1 = lds128(gp + r * ldg + 4);
                        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            wacc[t][8 * i + 0] = fmaf(xs[i], g0.x, wacc[t][8 * i + 0]); wacc[t][8 * i + 1] = fmaf(xs[i], g0.y, wacc[t][8 * i + 1]);
                            wacc[t][8 * i + 2] = fmaf(xs[i], g0.z, wacc[t][8 * i + 2]); wacc[t][8 * i + 3] = fmaf(xs[i], g0.w, wacc[t][8 * i + 3]);
                            wacc[t][8 * i + 4] = fmaf(xs[i], g1.x, wacc[t][8 * i + 4]); wacc[t][8 * i + 5] = fmaf(xs[i], g1.y, wacc[t][8 * i + 5]);
                            wacc[t][8 * i + 6] = fmaf(xs[i], g1.z, wacc[t][8 * i + 6]); wacc[t][8 * i + 7] = fmaf(xs[i], g1.w, wacc[t][8 * i + 7]);
                        }
                    }
                }
            }
        }
        if (a.db && tid < N) {
            float sacc = 0.f;
            for (int r = 0; r < LT; ++r) sacc += Gs[r * ldg + tid];
            bacc += sacc;
        }
    }
    // ---- flush the per-CTA partial sums ----
    if (a.dW) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int it = warp + 8 * t;
            if (it < KB * NB) {
                const int kb = it / NB, nb = it - kb * NB;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k = kb * 32 + rg * 4 + i, nn = nb * 32 + cg * 8 + j;
                        if (k < K && nn < N) atomicAdd(a.dW + (a.w_layout == 0 ? (size_t)nn * K + k : (size_t)k * N + nn), wacc[t][8 * i + j]);
                    }
            }
        }
    }
    if (a.db && tid < N) atomicAdd(a.db + tid, bacc);
}

// ---- layer backward through A: one thread per (state, node j) ------------------------------------------------------
// gHprev[b,j,:] = (skip ? gH[b,j,:] : 0) + sum_i A[b,i,j] gM[b,i,:]        gA[b,i,j] (+)= gM[b,i,:] . Hprev[b,j,:]
__global__ void attn_layer_bwd_kernel(const float* __restrict__ A, const float* __restrict__ Hprev, const float* __restrict__ gM,
                                      const float* __restrict__ gH, int skip, float* __restrict__ gHprev, float* __restrict__ gA,
                                      int accumulate_gA, int B, int n) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * n) return;
    const int b = idx / n, j = idx - b * n;
    float hp[32], acc[32];
    const float4* hp4 = reinterpret_cast<const float4*>(Hprev + (size_t)idx * 32);
#pragma unroll
    for (int c = 0; c < 8; ++c) { const float4 v = hp4[c]; hp[4 * c] = v.x; hp[4 * c + 1] = v.y; hp[4 * c + 2] = v.z; hp[4 * c + 3] = v.w; }
    if (skip) {
        const float4* g4 = reinterpret_cast<const float4*>(gH + (size_t)idx * 32);
#pragma unroll
        for (int c = 0; c < 8; ++c) { const float4 v = g4[c]; acc[4 * c] = v.x; acc[4 * c + 1] = v.y; acc[4 * c + 2] = v.z; acc[4 * c + 3] = v.w; }
    } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    }
    for (int i = 0; i < n; ++i) {
        const float aij = A[((size_t)b * n + i) * n + j];
        const float4* gm4 = reinterpret_cast<const float4*>(gM + ((size_t)b * n + i) * 32);
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 v = gm4[c];
            acc[4 * c] = fmaf(aij, v.x, acc[4 * c]); acc[4 * c + 1] = fmaf(aij, v.y, acc[4 * c + 1]);
            acc[4 * c + 2] = fmaf(aij, v.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(aij, v.w, acc[4 * c + 3]);
            d = fmaf(v.x, hp[4 * c], d); d = fmaf(v.y, hp[4 * c + 1], d); d = fmaf(v.z, hp[4 * c + 2], d); d = fmaf(v.w, hp[4 * c + 3], d);
        }
        float* ga = gA + ((size_t)b * n + i) * n + j;
        *ga = accumulate_gA ? *ga + d : d;
    }
    float4* o4 = reinterpret_cast<float4*>(gHprev + (size_t)idx * 32);
#pragma unroll
    for (int c = 0; c < 8; ++c) o4[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
}

// ---- similarity backward: A = softmax_j(Y_i . X_j) -----------------------------------------------------------------
// gS_ij = A_ij (gA_ij - sum_k gA_ik A_ik);  gY_i = sum_j gS_ij X_j;  gX_j += sum_i gS_ij Y_i
// one thread per (state, node); a CTA owns whole states so gS can be exchanged through shared memory.
__global__ void sim_bwd_kernel(const float* __restrict__ A, const float* __restrict__ gA, const float* __restrict__ X,
                               const float* __restrict__ Y, float* __restrict__ gY, float* __restrict__ gX, int B, int n, int spb) {
    extern __shared__ float gS[];                 // [spb][n][n]
    const int sl = threadIdx.x / n, i = threadIdx.x - sl * n;
    const int b = blockIdx.x * spb + sl;
    const bool active = sl < spb && b < B;
    if (active) {
        const float* arow = A + ((size_t)b * n + i) * n;
        const float* grow = gA + ((size_t)b * n + i) * n;
        float dot = 0.f;
        for (int j = 0; j < n; ++j) dot = fmaf(grow[j], arow[j], dot);
        float acc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = 0.f;
        for (int j = 0; j < n; ++j) {
            const float gs = arow[j] * (grow[j] - dot);
            gS[(sl * n + i) * n + j] = gs;
            const float4* x4 = reinterpret_cast<const float4*>(X + ((size_t)b * n + j) * 32);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 v = x4[c];
                acc[4 * c] = fmaf(gs, v.x, acc[4 * c]); acc[4 * c + 1] = fmaf(gs, v.y, acc[4 * c + 1]);
                acc[4 * c + 2] = fmaf(gs, v.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(gs, v.w, acc[4 * c + 3]);
            }
        }
        float4* o4 = reinterpret_cast<float4*>(gY + ((size_t)b * n + i) * 32);
#pragma unroll
        for (int c = 0; c < 8; ++c) o4[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
    }
    __syncthreads();
    if (active) {
        const int j = i;
        float4* o4 = reinterpret_cast<float4*>(gX + ((size_t)b * n + j) * 32);
        float acc[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) { const float4 v = o4[c]; acc[4 * c] = v.x; acc[4 * c + 1] = v.y; acc[4 * c + 2] = v.z; acc[4 * c + 3] = v.w; }
        for (int ii = 0; ii < n; ++ii) {
            const float gs = gS[(sl * n + ii) * n + j];
            const float4* y4 = reinterpret_cast<const float4*>(Y + ((size_t)b * n + ii) * 32);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 v = y4[c];
                acc[4 * c] = fmaf(gs, v.x, acc[4 * c]); acc[4 * c + 1] = fmaf(gs, v.y, acc[4 * c + 1]);
                acc[4 * c + 2] = fmaf(gs, v.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(gs, v.w, acc[4 * c + 3]);
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) o4[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
static Rows to_rows(const RglRows* r) {
    Rows o;
    if (r) { o.ptr = r->ptr; o.ld = r->ld; o.rpg = r->rows_per_group > 0 ? r->rows_per_group : 1; o.gstride = r->group_stride; }
    else { o.ptr = nullptr; o.ld = 0; o.rpg = 1; o.gstride = 0; }
    if (o.rpg == 1 && o.gstride == 0) o.gstride = o.ld;       // plain matrix: one row per group
    return o;
}

cudaError_t run_linear_bwd(const RglRows* G, int N, const RglRows* mask, const RglRows* Xin, int K, const float* W, int w_layout,
                           const RglRows* Gin, int accumulate, float* dW, float* db, int R, int num_sms, size_t max_smem,
                           cudaStream_t st) {
    LinBwdArgs a;
    a.G = to_rows(G); a.mask = to_rows(mask); a.Xin = to_rows(Xin); a.Gin = to_rows(Gin);
    a.N = N; a.K = K; a.R = R; a.W = W; a.w_layout = w_layout; a.accumulate = accumulate; a.dW = dW; a.db = db;
    a.ntiles = (R + LT - 1) / LT;
    const int NP4 = (N + 3) & ~3, NP32 = (N + 31) & ~31, KP32 = (K + 31) & ~31;
    const size_t smem = ((size_t)NP4 * (KP32 + 4) + (size_t)LT * (NP32 + 4) + (size_t)LT * (KP32 + 4)) * sizeof(float);
    if (smem > max_smem) return cudaErrorInvalidConfiguration;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(rows_linear_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int grid = a.ntiles < num_sms ? a.ntiles : num_sms;
    rows_linear_bwd_kernel<<<grid, 256, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t run_attn_layer_bwd(const float* A, const float* Hprev, const float* gM, const float* gH, int skip, float* gHprev,
                               float* gA, int accumulate_gA, int B, int n, cudaStream_t st) {
    const int total = B * n;
    attn_layer_bwd_kernel<<<(total + 127) / 128, 128, 0, st>>>(A, Hprev, gM, gH, skip, gHprev, gA, accumulate_gA, B, n);
    return cudaGetLastError();
}

cudaError_t run_sim_bwd(const float* A, const float* gA, const float* X, const float* Y, float* gY, float* gX, int B, int n,
                        cudaStream_t st) {
    const int spb = 256 / n;
    const int grid = (B + spb - 1) / spb;
    sim_bwd_kernel<<<grid, spb * n, (size_t)spb * n * n * sizeof(float), st>>>(A, gA, X, Y, gY, gX, B, n, spb);
    return cudaGetLastError();
}

}  // namespace rgl
