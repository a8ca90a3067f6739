// Backward kernels of the value-estimator training step (crowd_nav/utils/trainer.py:122-131 calls loss.backward()
// through ValueEstimator -> RGL; this file is that backward, hand-written).
//
// The forward of a training step is the fused graph_forward / value_head kernels run with activation saves
// (kernels.h GraphSave).  The backward is a short sequence of three generic kernels over those saved rows:
//   rows_linear_bwd   y = x W (+b) layers: data gradient G W^T, weight gradient x^T G (register-tiled, accumulated
//                     per CTA and flushed with one atomicAdd per element), bias gradient; optional relu mask
//   attn_layer_bwd    H' = relu(A H W) + H :  gH += A^T gM,  gA += gM H^T           (per state)
//   sim_bwd           A = softmax(Y X^T):  gS, gY = gS X, gX += gS^T Y            (per state)
#include <stdlib.h>
#include "kernels.h"
#include "train_common.cuh"

namespace rgl {

constexpr int LT = 64;    // rows per tile

__global__ void __launch_bounds__(256, 2) rows_linear_bwd_kernel(const LinBwdArgs a) {
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
    const int items = KB * NB;
    const int nsplit = items >= 8 ? 1 : 8 / items;
    float wacc[2][32];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int c = 0; c < 32; ++c) wacc[t][c] = 0.f;
    float bacc = 0.f;

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int r0 = tile * LT;
        __syncthreads();
        // ---- stage G (masked) and Xin, zero padded (float4 path when every row is 16-byte aligned) ----
        if (a.vecG) {
            const int n4 = ldg >> 2;
            for (int idx = tid; idx < LT * n4; idx += blockDim.x) {
                const int r = idx / n4, c = (idx - r * n4) << 2;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r0 + r < a.R && c < N) {
                    v = *reinterpret_cast<const float4*>(a.G.row(r0 + r) + c);
                    if (a.mask.ptr) {
                        const float4 m = *reinterpret_cast<const float4*>(a.mask.row(r0 + r) + c);
                        v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
                    }
                }
                sts128(Gs + r * ldg + c, v);
            }
        } else {
            for (int idx = tid; idx < LT * ldg; idx += blockDim.x) {
                const int r = idx / ldg, c = idx - r * ldg;
                float v = 0.f;
                if (r0 + r < a.R && c < N) {
                    v = a.G.row(r0 + r)[c];
                    if (a.mask.ptr && !(a.mask.row(r0 + r)[c] > 0.f)) v = 0.f;
                }
                Gs[idx] = v;
            }
        }
        if (a.Xin.ptr) {
            if (a.vecX) {
                const int k4 = ldx >> 2;
                for (int idx = tid; idx < LT * k4; idx += blockDim.x) {
                    const int r = idx / k4, c = (idx - r * k4) << 2;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r0 + r < a.R && c < K) v = *reinterpret_cast<const float4*>(a.Xin.row(r0 + r) + c);
                    sts128(Xs + r * ldx + c, v);
                }
            } else {
                for (int idx = tid; idx < LT * ldx; idx += blockDim.x) {
                    const int r = idx / ldx, c = idx - r * ldx;
                    Xs[idx] = (r0 + r < a.R && c < K) ? a.Xin.row(r0 + r)[c] : 0.f;
                }
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
                // >= 8 (k-block, n-block) items: item = warp + 8t over all rows; fewer: the rows of an item are split over
                // nsplit warps (partials are combined through shared memory at the end)
                int it, rbeg, rend;
                if (nsplit == 1) { it = warp + 8 * t; rbeg = 0; rend = LT; }
                else { it = warp % items; const int part = warp / items; rbeg = part * (LT / nsplit); rend = (t == 0 && part < nsplit) ? rbeg + LT / nsplit : rbeg; }
                if (it < items && rbeg < rend) {
                    const int kb = it / NB, nb = it - kb * NB;
                    const float* xp = Xs + kb * 32 + rg * 4;
                    const float* gp = Gs + nb * 32 + cg * 8;
#pragma unroll 4
                    for (int r = rbeg; r < rend; ++r) {
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
        if (nsplit == 1) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int it = warp + 8 * t;
                if (it < items) {
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
        } else {
            // combine the row-split partials of each item in shared memory (Gs/Xs are free now), then one atomic per element
            __syncthreads();
            float* red = smem;                                 // [8 warps][1024]: the launcher allocates >= 32 KB
            if (warp < items * nsplit) {
#pragma unroll
                for (int c = 0; c < 32; ++c) red[warp * 1024 + c * 32 + lane] = wacc[0][c];
            }
            __syncthreads();
            for (int e = tid; e < items * 1024; e += blockDim.x) {
                const int it = e >> 10, w = e & 1023;
                const int c = w >> 5, ln = w & 31;             // accumulator c of lane ln
                float sum = 0.f;
                for (int part = 0; part < nsplit; ++part) sum += red[(part * items + it) * 1024 + w];
                const int kb = it / NB, nb = it - kb * NB;
                const int k = kb * 32 + (ln & 7) * 4 + (c >> 3), nn = nb * 32 + (ln >> 3) * 8 + (c & 7);
                if (k < K && nn < N) atomicAdd(a.dW + (a.w_layout == 0 ? (size_t)nn * K + k : (size_t)k * N + nn), sum);
            }
        }
    }
    if (a.db && tid < N) atomicAdd(a.db + tid, bacc);
}

// ---- the same layer backward on the tensor cores (warp-level mma.sync m16n8k8, 3xTF32: fp32-grade products) -------
// The FMA kernel above spends 65 us per launch on ~10 us of HBM traffic (68 % of the C4 training step,
// profiles/r1_launches_train_b8192_nh10.md): every tile is load -> barrier -> FMA -> barrier with nothing in flight.
// Per tile of MROWS rows, 8 warps:
//   data gradient    Gin[r][k] = sum_n G[r][n] W[n][k]     A = G rows, B = W; a warp owns 16 rows (and, for 64-row tiles,
//                                                          half of the k columns)
//   weight gradient  dW[k][n] += sum_r X[r][k] G[r][n]     (16 x 8) tiles of dW dealt round-robin to the warps, the row
//                                                          index is the contraction: A = X^T, B = G, both read straight
//                                                          from the row-major staging tiles (strides = 8 mod 32 words:
//                                                          conflict-free fragment loads)
// Staging is cp.async (16-byte, zero-filled out of range) into a DOUBLE buffer: the next tile's G / mask / X rows are in
// flight while the current tile is multiplied; the relu mask is applied in place by the thread that copied the chunk.
// dW / db accumulate in registers over the tiles of a persistent CTA and are flushed with one atomicAdd per element.
//   <128, 4, true>   N, K <= 64: every layer whose row count is B*n (GCN layers, w_a, embeddings, motion head)
//   <64, 16, false>  N, K <= 128: the 100-wide value-head layers (B rows; single buffer: the tiles fill shared memory)
__device__ __forceinline__ void cp_async16(uint32_t dst_s, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst_s), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NLEFT>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(NLEFT) : "memory"); }

// Every shape parameter is a template argument (N8 = ceil(N/8), K16 = ceil(K/16)): the first version took them at run time
// and spent 95 % of its instructions on index arithmetic (ncu: 2 550 warp instructions per warp and tile for 96 HMMA,
// instruction-cache hit rate 80 %).
// F0 (fused first layer of a two-layer MLP, K = 64): the data gradient G W is not written to memory; it is multiplied by
// (Xin > 0) -- Xin is the hidden activation relu(x0 W0^T + b0) -- in place in the staged Xin tile and contracted with the
// MLP input x0 (a.X0, K0 <= 16 columns) right here: dW0 += (G W . (Xin > 0))^T x0, db0 += its column sums.  One launch
// and one pass over the rows for both Linear layers of w_r / w_h (graph_model.py:41-42).
template <int N8, int K16, int MROWS, bool DB, bool F0 = false>
__global__ void __launch_bounds__(256, (DB || MROWS <= 32) ? 2 : 1) rows_linear_bwd_mma_kernel(const LinBwdArgs a) {
    extern __shared__ __align__(128) float smem[];
    constexpr int NP = N8 * 8, KP = K16 * 16, K8 = K16 * 2;
    constexpr int ldg = ((NP + 31) / 32) * 32 + 8, ldx = ((KP + 31) / 32) * 32 + 8;      // = 8 mod 32 words
    constexpr int WPR = 8 / (MROWS / 16);   // warps sharing a 16-row block in the data gradient (1 or 2): they split the k columns
    constexpr int KPW = (K8 + WPR - 1) / WPR;                // k column tiles per warp in the data gradient
    // (16 x 8) tiles of dW, dealt to the warps as a WM x WN grid: warp (wm, wn) owns the tiles (mt, nt) = (wm + WM j, wn + WN q).
    // Every A fragment (X^T, 16 input columns x 8 rows) is split once and reused for the warp's NTW output-column tiles, every
    // B fragment (G) for its MTW tiles -- the first version dealt the tiles round-robin and re-split both operands per tile
    // (100 x 100 layer: 72 splits and 72 fragment loads per k-step and warp, now 22).
    constexpr int WM = K16 >= 8 ? 8 : (K16 >= 4 ? 4 : (K16 >= 2 ? 2 : 1)), WN = 8 / WM;
    constexpr int MTW = (K16 + WM - 1) / WM, NTW = (N8 + WN - 1) / WN, TW = MTW * NTW;
    constexpr int ld0 = 24;                                  // x0 tile: 16 columns (K0 <= 16, zero padded), stride = 24 mod 32 words
    constexpr int tileG = MROWS * ldg, tileX = MROWS * ldx, tile0 = F0 ? MROWS * ld0 : 0, stage_floats = 2 * tileG + tileX + tile0;
    constexpr int n4 = NP / 4, k4 = KP / 4;
    const int N = a.N, K = a.K;
    float* Ws = smem;                       // [NP][ldx]   Ws[n][k]
    float* Gs0 = Ws + NP * ldx;             // per stage: [MROWS][ldg] masked gradient | [MROWS][ldg] mask | [MROWS][ldx] layer input
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    if (a.W && a.vecW) {
        // nn.Linear layout with 16-byte aligned rows: the whole matrix in ONE cp.async group (a single global round trip;
        // the group is waited for together with the first tile's).  Padding rows / columns are zero-filled.
        for (int idx = tid; idx < NP * (ldx / 4); idx += 256) {
            const int nn = idx / (ldx / 4), k = (idx - nn * (ldx / 4)) << 2;
            const bool in = nn < N && k < K;
            cp_async16(smem_u32(Ws + nn * ldx + k), in ? a.W + (size_t)nn * K + k : a.W, in ? 16u : 0u);
        }
    } else if (a.W) {
        // 8 independent loads in flight per thread (a rolled loop would serialise one global round trip per element)
        for (int base = tid; base < NP * ldx; base += 8 * 256) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + 256 * u;
                const int nn = idx / ldx, k = idx - nn * ldx;
                v[u] = (idx < NP * ldx && nn < N && k < K) ? (a.w_layout == 0 ? a.W[(size_t)nn * K + k] : a.W[(size_t)k * N + nn]) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (base + 256 * u < NP * ldx) Ws[base + 256 * u] = v[u];
        }
    }
    float wacc[TW][4];
#pragma unroll
    for (int i = 0; i < TW; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) wacc[i][c] = 0.f;
    float bacc = 0.f;
    float w0acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};      // F0: dW0 fragment (16 hidden units x 16 input columns)
    float bacc0 = 0.f;
    if (F0) {                                                // columns K0..15 of the x0 tiles stay zero for the whole kernel
        for (int st = 0; st < (DB ? 2 : 1); ++st)
            for (int i = tid; i < tile0; i += 256) Gs0[st * stage_floats + 2 * tileG + tileX + i] = 0.f;
        __syncthreads();
    }

    // issue the copies of one tile into stage `st` (vector path: cp.async; otherwise plain loads / stores by the same thread)
    auto stage = [&](int tile, int st) {
        float* Gs = Gs0 + st * stage_floats;
        float* Ms = Gs + tileG;
        float* Xs = Ms + tileG;
        const int r0 = tile * MROWS;
        if (a.vecG) {
#pragma unroll
            for (int i = 0; i < (MROWS * n4 + 255) / 256; ++i) {
                const int idx = tid + 256 * i;
                const int r = idx / n4, c = (idx - r * n4) << 2;
                if (idx < MROWS * n4) {
                    const bool in = r0 + r < a.R && c < N;
                    cp_async16(smem_u32(Gs + r * ldg + c), in ? a.G.row(r0 + r) + c : a.G.ptr, in ? 16u : 0u);
                    if (a.mask.ptr) cp_async16(smem_u32(Ms + r * ldg + c), in ? a.mask.row(r0 + r) + c : a.mask.ptr, in ? 16u : 0u);
                }
            }
        } else {
            for (int idx = tid; idx < MROWS * NP; idx += 256) {
                const int r = idx / NP, c = idx - r * NP;
                float v = 0.f;
                if (r0 + r < a.R && c < N) {
                    v = a.G.row(r0 + r)[c];
                    if (a.mask.ptr && !(a.mask.row(r0 + r)[c] > 0.f)) v = 0.f;
                }
                Gs[r * ldg + c] = v;
            }
        }
        if (a.Xin.ptr) {
            if (a.vecX) {
#pragma unroll
                for (int i = 0; i < (MROWS * k4 + 255) / 256; ++i) {
                    const int idx = tid + 256 * i;
                    const int r = idx / k4, c = (idx - r * k4) << 2;
                    if (idx < MROWS * k4) {
                        const bool in = r0 + r < a.R && c < K;
                        cp_async16(smem_u32(Xs + r * ldx + c), in ? a.Xin.row(r0 + r) + c : a.Xin.ptr, in ? 16u : 0u);
                    }
                }
            } else {
                for (int idx = tid; idx < MROWS * KP; idx += 256) {
                    const int r = idx / KP, c = idx - r * KP;
                    Xs[r * ldx + c] = (r0 + r < a.R && c < K) ? a.Xin.row(r0 + r)[c] : 0.f;
                }
            }
        }
        if (F0) {
            float* X0s = Xs + tileX;
            for (int idx = tid; idx < MROWS * a.K0; idx += 256) {
                const int r = idx / a.K0, c = idx - r * a.K0;
                const bool in = r0 + r < a.R;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(smem_u32(X0s + r * ld0 + c)),
                             "l"(in ? a.X0.row(r0 + r) + c : a.X0.ptr), "r"(in ? 4u : 0u) : "memory");
            }
        }
        cp_async_commit();
    };

    int it = 0;
    if (blockIdx.x < a.ntiles) stage(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
        const int st = DB ? (it & 1) : 0;
        float* Gs = Gs0 + st * stage_floats;
        float* Ms = Gs + tileG;
        float* Xs = Ms + tileG;
        const int r0 = tile * MROWS;
        const int next = tile + gridDim.x;
        if (DB && next < a.ntiles) { stage(next, st ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        // relu mask, in place, by the thread that copied the chunk (its own cp.async results are visible to it)
        if (a.vecG && a.mask.ptr) {
#pragma unroll
            for (int i = 0; i < (MROWS * n4 + 255) / 256; ++i) {
                const int idx = tid + 256 * i;
                const int r = idx / n4, c = (idx - r * n4) << 2;
                if (idx < MROWS * n4) {
                    float4 v = lds128(Gs + r * ldg + c);
                    const float4 m = lds128(Ms + r * ldg + c);
                    v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
                    sts128(Gs + r * ldg + c, v);
                }
            }
        }
        __syncthreads();

        // ---- data gradient: a warp owns 16 rows (x a share of the k columns when two warps split a row block) ----
        auto data_grad = [&]() {
        if (a.W && (F0 || a.Gin.ptr)) {
            const int rb = warp / WPR, part = warp - rb * WPR;
            const int nt0 = part * KPW;
            float acc[KPW][4];
#pragma unroll
            for (int nt = 0; nt < KPW; ++nt)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[nt][c] = 0.f;
            const float* gr = Gs + (rb * 16 + g) * ldg + t;
            const float* wr0 = Ws + t * ldx + nt0 * 8 + g;
#pragma unroll 2
            for (int ks = 0; ks < N8; ++ks) {
                uint32_t ahi[4], alo[4];
                split_tf32(gr[ks * 8], ahi[0], alo[0]);
                split_tf32(gr[ks * 8 + 8 * ldg], ahi[1], alo[1]);
                split_tf32(gr[ks * 8 + 4], ahi[2], alo[2]);
                split_tf32(gr[ks * 8 + 8 * ldg + 4], ahi[3], alo[3]);
                const float* wr = wr0 + ks * 8 * ldx;
#pragma unroll
                for (int nt = 0; nt < KPW; ++nt) {
                    if (WPR == 1 || nt0 + nt < K8) {
                        uint32_t bh0, bl0, bh1, bl1;
                        split_tf32(wr[nt * 8], bh0, bl0);
                        split_tf32(wr[nt * 8 + 4 * ldx], bh1, bl1);
                        mma_tf32(acc[nt], alo, bh0, bh1);
                        mma_tf32(acc[nt], ahi, bl0, bl1);
                        mma_tf32(acc[nt], ahi, bh0, bh1);
                    }
                }
            }
            if (F0) {
                // the hidden layer's gradient stays on chip: masked with (hidden activation > 0) it replaces that activation
                // in the staged tile (each element is read and overwritten by the one thread that owns it; the weight
                // gradient that needs the activation has already run)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float* o = Xs + (rb * 16 + g + 8 * q) * ldx + nt0 * 8 + 2 * t;
#pragma unroll
                    for (int nt = 0; nt < KPW; ++nt) {
                        if (WPR == 1 || nt0 + nt < K8) {
                            const float2 h = *reinterpret_cast<const float2*>(o + nt * 8);
                            *reinterpret_cast<float2*>(o + nt * 8) = make_float2(h.x > 0.f ? acc[nt][2 * q] : 0.f, h.y > 0.f ? acc[nt][2 * q + 1] : 0.f);
                        }
                    }
                }
                return;
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int r = r0 + rb * 16 + g + 8 * q;
                if (r < a.R) {
                    float* o = a.Gin.row(r) + nt0 * 8 + 2 * t;
#pragma unroll
                    for (int nt = 0; nt < KPW; ++nt) {
                        const int k = (nt0 + nt) * 8 + 2 * t;
                        if (k + 1 < K && ((reinterpret_cast<uintptr_t>(o + nt * 8) & 7u) == 0)) {
                            float2 v = make_float2(acc[nt][2 * q], acc[nt][2 * q + 1]);
                            if (a.accumulate) { const float2 old = *reinterpret_cast<const float2*>(o + nt * 8); v.x += old.x; v.y += old.y; }
                            *reinterpret_cast<float2*>(o + nt * 8) = v;
                        } else {
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                if (k + j < K) o[nt * 8 + j] = a.accumulate ? o[nt * 8 + j] + acc[nt][2 * q + j] : acc[nt][2 * q + j];
                        }
                    }
                }
            }
        }
        };
        // ---- weight gradient: the rows of the tile are the contraction ----
        auto weight_grad = [&]() {
        if (a.dW) {
            const int wm = warp % WM, wn = warp / WM;
#pragma unroll 2
            for (int ks = 0; ks < MROWS / 8; ++ks) {
                uint32_t ahi[MTW][4], alo[MTW][4];
#pragma unroll
                for (int jm = 0; jm < MTW; ++jm) {
                    const int mt = wm + WM * jm;
                    if (K16 % WM == 0 || mt < K16) {
                        const float* xp = Xs + (ks * 8 + t) * ldx + mt * 16 + g;
                        split_tf32(xp[0], ahi[jm][0], alo[jm][0]);
                        split_tf32(xp[8], ahi[jm][1], alo[jm][1]);
                        split_tf32(xp[4 * ldx], ahi[jm][2], alo[jm][2]);
                        split_tf32(xp[4 * ldx + 8], ahi[jm][3], alo[jm][3]);
                    }
                }
#pragma unroll
                for (int q = 0; q < NTW; ++q) {
                    const int nt = wn + WN * q;
                    if (N8 % WN == 0 || nt < N8) {
                        const float* gp = Gs + (ks * 8 + t) * ldg + nt * 8 + g;
                        uint32_t bh0, bl0, bh1, bl1;
                        split_tf32(gp[0], bh0, bl0);
                        split_tf32(gp[4 * ldg], bh1, bl1);
#pragma unroll
                        for (int jm = 0; jm < MTW; ++jm) {
                            if (K16 % WM == 0 || wm + WM * jm < K16) {
                                mma_tf32(wacc[jm * NTW + q], alo[jm], bh0, bh1);
                                mma_tf32(wacc[jm * NTW + q], ahi[jm], bl0, bl1);
                                mma_tf32(wacc[jm * NTW + q], ahi[jm], bh0, bh1);
                            }
                        }
                    }
                }
            }
        }
        if (a.db && (tid & 127) < N) {
            // column sums: both halves of the CTA take half of the rows, four independent partial sums each
            const float* gc = Gs + (tid >> 7) * (MROWS / 2) * ldg + (tid & 127);
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
            for (int r = 0; r < MROWS / 2; r += 4) {
                s0 += gc[r * ldg]; s1 += gc[(r + 1) * ldg]; s2 += gc[(r + 2) * ldg]; s3 += gc[(r + 3) * ldg];
            }
            bacc += (s0 + s1) + (s2 + s3);
        }
        };
        if (!F0) {
            data_grad();
            weight_grad();
        } else {
            weight_grad();                      // needs the hidden activation tile ...
            __syncthreads();
            data_grad();                        // ... which the masked hidden gradient then replaces
            __syncthreads();
            // first layer: dW0[m][c] += sum_rows Hg[row][m] x0[row][c]; warp = (16 hidden units, half of the rows)
            {
                const float* X0s = Xs + tileX;
                const int mt = warp & 3, kh = warp >> 2;
                const float* hp = Xs + t * ldx + mt * 16 + g;
                const float* xq = X0s + t * ld0 + g;
#pragma unroll 2
                for (int ks = kh * (MROWS / 16); ks < (kh + 1) * (MROWS / 16); ++ks) {
                    uint32_t ahi[4], alo[4];
                    split_tf32(hp[(ks * 8) * ldx], ahi[0], alo[0]);
                    split_tf32(hp[(ks * 8) * ldx + 8], ahi[1], alo[1]);
                    split_tf32(hp[(ks * 8 + 4) * ldx], ahi[2], alo[2]);
                    split_tf32(hp[(ks * 8 + 4) * ldx + 8], ahi[3], alo[3]);
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        if (nt * 8 < a.K0) {
                            uint32_t bh0, bl0, bh1, bl1;
                            split_tf32(xq[(ks * 8) * ld0 + nt * 8], bh0, bl0);
                            split_tf32(xq[(ks * 8 + 4) * ld0 + nt * 8], bh1, bl1);
                            mma_tf32(w0acc[nt], alo, bh0, bh1);
                            mma_tf32(w0acc[nt], ahi, bl0, bl1);
                            mma_tf32(w0acc[nt], ahi, bh0, bh1);
                        }
                    }
                }
                if (tid < 128) {                // db0: column sums of the masked hidden gradient, half of the rows each
                    const float* hc = Xs + (tid >> 6) * (MROWS / 2) * ldx + (tid & 63);
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
                    for (int r = 0; r < MROWS / 2; r += 2) { s0 += hc[r * ldx]; s1 += hc[(r + 1) * ldx]; }
                    bacc0 += s0 + s1;
                }
            }
        }
        __syncthreads();                    // every warp is done with this stage before it is refilled
        if (!DB && next < a.ntiles) stage(next, 0);
    }
    // ---- flush the per-CTA partial sums: accumulator fragment (m16n8): [0],[1] = row g, cols 2t, 2t+1; [2],[3] = row g+8 ----
    // Each warp transposes its (16 x 8) fragment through 1 KB of the now idle staging memory so that a lane owns four consecutive
    // elements of dW and issues ONE 16-byte vector reduction (REDG.E.ADD.F32x4) for them: a quarter of the reduction
    // instructions (value layer 2: 256 CTAs x 10 000 elements; 19.6 -> 18.2 us per launch).
    if (a.dW) {
        float* scr = Gs0 + warp * 256;
#pragma unroll
        for (int i = 0; i < TW; ++i) {
            const int mt = warp % WM + WM * (i / NTW), nt = warp / WM + WN * (i % NTW);
            if (mt < K16 && nt < N8) {
                if (a.w_layout == 0 && (K & 3)) {       // rows of dW not 16-byte aligned (K = 5, 9): scalar reductions from the fragment
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int k = mt * 16 + g + 8 * (c >> 1), nn = nt * 8 + 2 * t + (c & 1);
                        if (k < K && nn < N) atomicAdd(a.dW + (size_t)nn * K + k, wacc[i][c]);
                    }
                    continue;
                }
                float4 v;
                size_t off;
                int left;                   // valid elements among the lane's four (<= 0: none)
                if (a.w_layout == 0) {      // dW[nn][k]: k contiguous.  scratch [8 n][20]
#pragma unroll
                    for (int c = 0; c < 4; ++c) scr[(2 * t + (c & 1)) * 20 + g + 8 * (c >> 1)] = wacc[i][c];
                    __syncwarp();
                    const int nl = lane >> 2, k4 = (lane & 3) * 4;
                    v = lds128(scr + nl * 20 + k4);
                    const int nn = nt * 8 + nl, k = mt * 16 + k4;
                    off = (size_t)nn * K + k;
                    left = nn < N ? K - k : 0;
                } else {                    // dW[k][nn]: nn contiguous.  scratch [16 k][12]
#pragma unroll
                    for (int c = 0; c < 4; ++c) scr[(g + 8 * (c >> 1)) * 12 + 2 * t + (c & 1)] = wacc[i][c];
                    __syncwarp();
                    const int kl = lane >> 1, n4 = (lane & 1) * 4;
                    v = lds128(scr + kl * 12 + n4);
                    const int k = mt * 16 + kl, nn = nt * 8 + n4;
                    off = (size_t)k * N + nn;
                    left = k < K ? N - nn : 0;
                }
                __syncwarp();
                float* dst = a.dW + off;
                if (left >= 4 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                } else {
                    if (left > 0) atomicAdd(dst, v.x);
                    if (left > 1) atomicAdd(dst + 1, v.y);
                    if (left > 2) atomicAdd(dst + 2, v.z);
                    if (left > 3) atomicAdd(dst + 3, v.w);
                }
            }
        }
    }
    if (a.db && (tid & 127) < N) atomicAdd(a.db + (tid & 127), bacc);
    if (F0) {
        const int mt = warp & 3;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int m = mt * 16 + g + 8 * (c >> 1), col = nt * 8 + 2 * t + (c & 1);
                if (col < a.K0) atomicAdd(a.dW0 + (size_t)m * a.K0 + col, w0acc[nt][c]);
            }
        if (a.db0 && tid < 128) atomicAdd(a.db0 + (tid & 63), bacc0);
    }
}

template <int N8, int K16, int MROWS, bool DB, bool F0 = false>
static cudaError_t launch_bwd_mma(LinBwdArgs& a, int num_sms, size_t max_smem, cudaStream_t st) {
    constexpr int NP = N8 * 8, KP = K16 * 16;
    constexpr int ldg = ((NP + 31) / 32) * 32 + 8, ldx = ((KP + 31) / 32) * 32 + 8;
    constexpr size_t stage = (size_t)2 * MROWS * ldg + (size_t)MROWS * ldx + (F0 ? (size_t)MROWS * 24 : 0);
    constexpr size_t sm = ((size_t)NP * ldx + stage * (DB ? 2 : 1)) * sizeof(float);
    static_assert(sm <= 227 * 1024, "staging tiles exceed the shared memory of an SM");
    if (sm > max_smem) return cudaErrorInvalidConfiguration;
    if (cudaError_t e = ensure_dyn_smem(rows_linear_bwd_mma_kernel<N8, K16, MROWS, DB, F0>, (int)max_smem)) return e;
    a.ntiles = (a.R + MROWS - 1) / MROWS;
    int per_sm = (int)((227 * 1024) / (sm + 1024));
    // registers: __launch_bounds__(256, 2) guarantees two CTAs per SM; a third fits when the kernel needs <= 85 registers
    cudaFuncAttributes fa;
    int reg_cap = (DB || MROWS <= 32) ? 2 : 1;
    if (DB && cudaFuncGetAttributes(&fa, rows_linear_bwd_mma_kernel<N8, K16, MROWS, DB, F0>) == cudaSuccess && fa.numRegs * 256 * 3 <= 65536) reg_cap = 3;
    per_sm = per_sm < 1 ? 1 : (per_sm > reg_cap ? reg_cap : per_sm);
    const int cap = num_sms * per_sm;
    rows_linear_bwd_mma_kernel<N8, K16, MROWS, DB, F0><<<a.ntiles < cap ? a.ntiles : cap, 256, sm, st>>>(a);
    return cudaGetLastError();
}

// the layer shapes of the path (N8 = ceil(N/8), K16 = ceil(K/16)); cudaErrorNotSupported: no instantiation
static cudaError_t dispatch_bwd_mma(LinBwdArgs& a, int num_sms, size_t max_smem, cudaStream_t st) {
    const int N8 = (a.N + 7) / 8, K16 = (a.K + 15) / 16;
#define RGL_BWD_CASE(n8, k16, rows, db) if (N8 == n8 && K16 == k16) return launch_bwd_mma<n8, k16, rows, db>(a, num_sms, max_smem, st)
    static const char* rows_env = getenv("RGL_BWD_ROWS");          // experiments only: 128-row tiles (one CTA per SM)
    if (rows_env && rows_env[0] == '1') {
        RGL_BWD_CASE(4, 2, 128, true);
        RGL_BWD_CASE(4, 4, 128, true);
        RGL_BWD_CASE(8, 1, 128, true);
    }
    if (a.dW0) {                        // fused two-layer MLP backward (w_r / w_h): N = 32, K = 64
        if (N8 != 4 || K16 != 4 || a.K != 64 || a.K0 < 1 || a.K0 > 16 || !a.W || !a.X0.ptr) return cudaErrorNotSupported;
        if ((a.R + 63) / 64 <= num_sms) return launch_bwd_mma<4, 4, 32, true, true>(a, num_sms, max_smem, st);
        return launch_bwd_mma<4, 4, 64, true, true>(a, num_sms, max_smem, st);
    }
    if ((a.R + 63) / 64 <= num_sms) {
        // fewer 64-row tiles than SMs (the B-row layers: value head, robot embedding): 32-row tiles double the CTAs and the
        // warps per SM -- these launches are bound by the latency of ONE tile, not by throughput
        RGL_BWD_CASE(13, 7, 32, false);
        RGL_BWD_CASE(13, 2, 32, false);
        RGL_BWD_CASE(1, 7, 32, false);
        RGL_BWD_CASE(4, 2, 32, true);
        RGL_BWD_CASE(4, 4, 32, true);
    }
    RGL_BWD_CASE(4, 2, 64, true);       // 32 x 32: GCN layers, w_a, value layer 0          (64-row tiles: two CTAs per SM)
    RGL_BWD_CASE(4, 4, 64, true);       // 32 x 64: embedding layer 2 (w_r.2, w_h.2)
    RGL_BWD_CASE(8, 1, 64, true);       // 64 x 5 / 64 x 9: embedding layer 1 (w_h.0, w_r.0)
    RGL_BWD_CASE(8, 2, 64, true);       // 64 x 32: motion head layer 0
    RGL_BWD_CASE(1, 4, 64, true);       // 5 x 64: motion head layer 1
    RGL_BWD_CASE(1, 7, 64, false);      // 1 x 100: value layer 3
    RGL_BWD_CASE(13, 7, 64, false);     // 100 x 100: value layer 2
    RGL_BWD_CASE(13, 2, 64, false);     // 100 x 32: value layer 1
#undef RGL_BWD_CASE
    return cudaErrorNotSupported;
}

// ---- layer backward through A: FOUR threads per (state, node j), 8 feature columns each ---------------------------
// gHprev[b,j,:] = (skip ? gH[b,j,:] : 0) + sum_i A[b,i,j] gM[b,i,:]        gA[b,i,j] (+)= gM[b,i,:] . Hprev[b,j,:]
// (one thread per row ran n dependent 128-byte row loads with 19 warps per SM: latency-bound, 22 us for 9 us of traffic;
//  a quad of lanes splits the row, the dot product is closed with two shuffles)
// mask (optional, [B,n,32]): gM is multiplied by (mask > 0) on load -- the relu mask of the reassociated layer
// H' = relu(A (H W)), where this kernel runs FIRST (gM = gH', Hprev = H W) and the linear backward second.
// up_rows: the first up_rows node rows of gM carry gradient, the others are zero and skipped (1 for the top layer of the
// value step, whose head reads the robot row only; n otherwise).
__global__ void attn_layer_bwd_kernel(const float* __restrict__ A, const float* __restrict__ Hprev, const float* __restrict__ gM,
                                      const float* __restrict__ gH, int skip, float* __restrict__ gHprev, float* __restrict__ gA,
                                      int accumulate_gA, int B, int n, const float* __restrict__ mask, int up_rows) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int q = (int)(t & 3);
    const long long total = (long long)B * n;
    const bool live = (t >> 2) < total;
    const long long row = live ? (t >> 2) : total - 1;           // idle quads shadow the last row (the shuffles stay warp-wide)
    const int b = (int)(row / n), j = (int)(row - (long long)b * n);
    float hp[8], acc[8];
    {
        const float4* hp4 = reinterpret_cast<const float4*>(Hprev + row * 32 + 8 * q);
        const float4 u = hp4[0], v = hp4[1];
        hp[0] = u.x; hp[1] = u.y; hp[2] = u.z; hp[3] = u.w; hp[4] = v.x; hp[5] = v.y; hp[6] = v.z; hp[7] = v.w;
    }
    if (skip) {
        const float4* g4 = reinterpret_cast<const float4*>(gH + row * 32 + 8 * q);
        const float4 u = g4[0], v = g4[1];
        acc[0] = u.x; acc[1] = u.y; acc[2] = u.z; acc[3] = u.w; acc[4] = v.x; acc[5] = v.y; acc[6] = v.z; acc[7] = v.w;
    } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    }
    const float* Ab = A + (size_t)b * n * n + j;
    const float* gMb = gM + (size_t)b * n * 32 + 8 * q;
    const float* mkb = mask ? mask + (size_t)b * n * 32 + 8 * q : nullptr;
    float* gAb = gA + (size_t)b * n * n + j;
    // rows >= up_rows of gM are zero by contract and never read (value head: only the robot row carries gradient)
    if (!accumulate_gA && q == 0 && live)
        for (int i = up_rows; i < n; ++i) gAb[i * n] = 0.f;
#pragma unroll 4
    for (int i = 0; i < up_rows; ++i) {
        const float aij = Ab[i * n];
        const float4 u = *reinterpret_cast<const float4*>(gMb + i * 32), v = *reinterpret_cast<const float4*>(gMb + i * 32 + 4);
        float gm[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
        if (mkb) {
            const float4 mu = *reinterpret_cast<const float4*>(mkb + i * 32), mv = *reinterpret_cast<const float4*>(mkb + i * 32 + 4);
            const float mk[8] = {mu.x, mu.y, mu.z, mu.w, mv.x, mv.y, mv.z, mv.w};
#pragma unroll
            for (int c = 0; c < 8; ++c) gm[c] = mk[c] > 0.f ? gm[c] : 0.f;
        }
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) { acc[c] = fmaf(aij, gm[c], acc[c]); d = fmaf(gm[c], hp[c], d); }
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        if (q == 0 && live) gAb[i * n] = accumulate_gA ? gAb[i * n] + d : d;
    }
    if (live) {
        float4* o4 = reinterpret_cast<float4*>(gHprev + row * 32 + 8 * q);
        o4[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o4[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

// ---- similarity backward: A = softmax_j(Y_i . X_j), four threads per (state, node), 8 columns each --------------------
// gS_ij = A_ij (gA_ij - sum_k gA_ik A_ik);  gY_i = sum_j gS_ij X_j;  gX_j += sum_i gS_ij Y_i
// a CTA owns whole states so gS can be exchanged through shared memory.
__global__ void sim_bwd_kernel(const float* __restrict__ A, const float* __restrict__ gA, const float* __restrict__ X,
                               const float* __restrict__ Y, float* __restrict__ gY, float* __restrict__ gX, int B, int n, int spb) {
    extern __shared__ float gS[];                 // [spb][n][n]
    const int q = threadIdx.x & 3, r = threadIdx.x >> 2;
    const int sl = r / n, i = r - sl * n;
    const int b = blockIdx.x * spb + sl;
    const bool active = sl < spb && b < B;
    const size_t row = (size_t)b * n + i;
    if (active) {
        const float* arow = A + row * n;
        const float* grow = gA + row * n;
        float dot = 0.f;
        for (int j = 0; j < n; ++j) dot = fmaf(grow[j], arow[j], dot);
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.f;
        const float* Xb = X + (size_t)b * n * 32 + 8 * q;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const float gs = arow[j] * (grow[j] - dot);
            if (q == 0) gS[(sl * n + i) * n + j] = gs;
            const float4 u = *reinterpret_cast<const float4*>(Xb + j * 32), v = *reinterpret_cast<const float4*>(Xb + j * 32 + 4);
            acc[0] = fmaf(gs, u.x, acc[0]); acc[1] = fmaf(gs, u.y, acc[1]); acc[2] = fmaf(gs, u.z, acc[2]); acc[3] = fmaf(gs, u.w, acc[3]);
            acc[4] = fmaf(gs, v.x, acc[4]); acc[5] = fmaf(gs, v.y, acc[5]); acc[6] = fmaf(gs, v.z, acc[6]); acc[7] = fmaf(gs, v.w, acc[7]);
        }
        float4* o4 = reinterpret_cast<float4*>(gY + row * 32 + 8 * q);
        o4[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o4[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    __syncthreads();
    if (active) {
        const int j = i;
        float4* o4 = reinterpret_cast<float4*>(gX + row * 32 + 8 * q);
        const float4 u0 = o4[0], v0 = o4[1];
        float acc[8] = {u0.x, u0.y, u0.z, u0.w, v0.x, v0.y, v0.z, v0.w};
        const float* Yb = Y + (size_t)b * n * 32 + 8 * q;
#pragma unroll 4
        for (int ii = 0; ii < n; ++ii) {
            const float gs = gS[(sl * n + ii) * n + j];
            const float4 u = *reinterpret_cast<const float4*>(Yb + ii * 32), v = *reinterpret_cast<const float4*>(Yb + ii * 32 + 4);
            acc[0] = fmaf(gs, u.x, acc[0]); acc[1] = fmaf(gs, u.y, acc[1]); acc[2] = fmaf(gs, u.z, acc[2]); acc[3] = fmaf(gs, u.w, acc[3]);
            acc[4] = fmaf(gs, v.x, acc[4]); acc[5] = fmaf(gs, v.y, acc[5]); acc[6] = fmaf(gs, v.z, acc[6]); acc[7] = fmaf(gs, v.w, acc[7]);
        }
        o4[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o4[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

// ---- staged layer backward through A, optionally fused with the similarity backward (tcgen05 training forward) ------
// The two kernels above walk n rows per thread straight from global memory: ncu shows them latency-bound (long-scoreboard
// stalls 24 of 26 stall cycles per issue, 22 % issue utilisation, 33 / 29 us for 9 / 11 us of traffic at C4).  Here a CTA
// owns spb whole states, fetches every operand of those states with ONE group of cp.async copies (a single global round
// trip), and runs the row loops out of shared memory:
//   part 1   gZ[j] = sum_{i < up_rows} A[i][j] (gM[i] . mask[i]);   gA[i][j] (+)= (gM[i] . mask[i]) . Z[j]
//   SIM      gS = A (gA - rowsum(gA A));   gY[i] = sum_j gS[i][j] X[j];   gX[j] (+)= sum_i gS[i][j] Y[i]
// With SIM the attention gradient gA of layer 0 never leaves shared memory.  Thread = (state slot, node, column octet).
struct AttnSimArgs {
    const float *A, *Z, *gM, *mask, *gA_in, *X, *Y;
    float *gZ, *gA_out, *gY, *gX;
    int up_rows, gx_accumulate, B, n, spb;
};

__device__ __forceinline__ void cp_async4(uint32_t dst_s, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst_s), "l"(src) : "memory");
}

// NT: compile-time node count (6, 11, 21: the counts of the tcgen05 forward; 0 = run-time n): the row loops unroll fully and
// every shared-memory address becomes base + immediate.
template <bool SIM, int NT>
__global__ void __launch_bounds__(256) attn_sim_bwd_kernel(const AttnSimArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int n = NT ? NT : a.n, nn = n * n, spb = NT ? (64 / (NT ? NT : 1)) : a.spb;
    const int b0 = blockIdx.x * spb;
    const int cnt = a.B - b0 < spb ? a.B - b0 : spb;
    const int tile = spb * n * 32;
    float* gMs = sm;                              // [spb][n][32] upstream gradient (masked in place)
    float* Mks = gMs + tile;                      // relu mask
    float* Zs = Mks + tile;                       // Z_l = H_{l-1} W_l
    float* Xs = Zs + tile;                        // SIM only
    float* Ys = Xs + tile;
    float* As = SIM ? Ys + tile : Xs;             // [spb][n*n]
    float* gAs = As + spb * nn;
    float* gSs = gAs + spb * nn;                  // SIM only
    const int tid = threadIdx.x, nt = blockDim.x;
    const int rows = cnt * n;
    const size_t g0 = (size_t)b0 * n * 32;

    for (int c = tid; c < rows * 8; c += nt) {
        const int r = c >> 3, o = r * 32 + ((c & 7) << 2);
        if (r % n < a.up_rows) {
            cp_async16(smem_u32(gMs + o), a.gM + g0 + o, 16u);
            if (a.mask) cp_async16(smem_u32(Mks + o), a.mask + g0 + o, 16u);
        }
        cp_async16(smem_u32(Zs + o), a.Z + g0 + o, 16u);
        if (SIM) {
            cp_async16(smem_u32(Xs + o), a.X + g0 + o, 16u);
            cp_async16(smem_u32(Ys + o), a.Y + g0 + o, 16u);
        }
    }
    for (int c = tid; c < cnt * nn; c += nt) {
        cp_async4(smem_u32(As + c), a.A + (size_t)b0 * nn + c);
        if (a.gA_in) cp_async4(smem_u32(gAs + c), a.gA_in + (size_t)b0 * nn + c);
        else gAs[c] = 0.f;
    }
    cp_async_commit();

    const int q = tid & 3, r = tid >> 2;
    const int sl = r / n, j = r - sl * n;
    const bool live = r < rows;
    float old[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) old[c] = 0.f;
    if (SIM && a.gx_accumulate && live) {          // this thread's row of gX, fetched next to the staging copies
        const float4* p = reinterpret_cast<const float4*>(a.gX + g0 + r * 32 + 8 * q);
        const float4 u = p[0], v = p[1];
        old[0] = u.x; old[1] = u.y; old[2] = u.z; old[3] = u.w; old[4] = v.x; old[5] = v.y; old[6] = v.z; old[7] = v.w;
    }
    cp_async_wait<0>();
    if (a.mask) {                                  // in place, by the thread that copied the chunk
        for (int c = tid; c < rows * 8; c += nt) {
            const int rr = c >> 3, o = rr * 32 + ((c & 7) << 2);
            if (rr % n < a.up_rows) {
                float4 v = lds128(gMs + o);
                const float4 m = lds128(Mks + o);
                v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
                sts128(gMs + o, v);
            }
        }
    }
    __syncthreads();

    // ---- part 1: thread (sl, j, q) ----
    {
        float hp[8], acc[8];
        const int rs = live ? r : 0;               // idle threads shadow row 0 (the shuffles stay warp-wide)
        const int sls = live ? sl : 0, js = live ? j : 0;
        const float4 u = lds128(Zs + rs * 32 + 8 * q), v = lds128(Zs + rs * 32 + 8 * q + 4);
        hp[0] = u.x; hp[1] = u.y; hp[2] = u.z; hp[3] = u.w; hp[4] = v.x; hp[5] = v.y; hp[6] = v.z; hp[7] = v.w;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.f;
        const float* Ab = As + sls * nn + js;
        float* gAb = gAs + sls * nn + js;
        const float* gMb = gMs + sls * n * 32 + 8 * q;
        auto row = [&](int i) {
            const float aij = Ab[i * n];
            const float4 gu = lds128(gMb + i * 32), gv = lds128(gMb + i * 32 + 4);
            const float gm[8] = {gu.x, gu.y, gu.z, gu.w, gv.x, gv.y, gv.z, gv.w};
            float d = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) { acc[c] = fmaf(aij, gm[c], acc[c]); d = fmaf(gm[c], hp[c], d); }
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            if (q == 0 && live) gAb[i * n] += d;
        };
        if (NT && a.up_rows == NT) {
#pragma unroll
            for (int i = 0; i < (NT ? NT : 1); ++i) row(i);
        } else {
#pragma unroll 2
            for (int i = 0; i < a.up_rows; ++i) row(i);
        }
        if (live) {
            float4* o4 = reinterpret_cast<float4*>(a.gZ + g0 + r * 32 + 8 * q);
            o4[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            o4[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
    }
    __syncthreads();
    if (!SIM) {
        if (a.gA_out)
            for (int c = tid; c < cnt * nn; c += nt) a.gA_out[(size_t)b0 * nn + c] = gAs[c];
        return;
    }
    // ---- similarity backward: thread (sl, i = j, q) ----
    if (live) {
        const int i = j;
        const float* arow = As + sl * nn + i * n;
        const float* grow = gAs + sl * nn + i * n;
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < n; ++k) dot = fmaf(grow[k], arow[k], dot);
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.f;
        const float* Xb = Xs + sl * n * 32 + 8 * q;
#pragma unroll
        for (int k = 0; k < n; ++k) {
            const float gs = arow[k] * (grow[k] - dot);
            if (q == 0) gSs[sl * nn + i * n + k] = gs;
            const float4 u = lds128(Xb + k * 32), v = lds128(Xb + k * 32 + 4);
            acc[0] = fmaf(gs, u.x, acc[0]); acc[1] = fmaf(gs, u.y, acc[1]); acc[2] = fmaf(gs, u.z, acc[2]); acc[3] = fmaf(gs, u.w, acc[3]);
            acc[4] = fmaf(gs, v.x, acc[4]); acc[5] = fmaf(gs, v.y, acc[5]); acc[6] = fmaf(gs, v.z, acc[6]); acc[7] = fmaf(gs, v.w, acc[7]);
        }
        float4* o4 = reinterpret_cast<float4*>(a.gY + g0 + r * 32 + 8 * q);
        o4[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o4[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    __syncthreads();
    if (live) {
        const float* Yb = Ys + sl * n * 32 + 8 * q;
        const float* gsc = gSs + sl * nn + j;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float gs = gsc[i * n];
            const float4 u = lds128(Yb + i * 32), v = lds128(Yb + i * 32 + 4);
            old[0] = fmaf(gs, u.x, old[0]); old[1] = fmaf(gs, u.y, old[1]); old[2] = fmaf(gs, u.z, old[2]); old[3] = fmaf(gs, u.w, old[3]);
            old[4] = fmaf(gs, v.x, old[4]); old[5] = fmaf(gs, v.y, old[5]); old[6] = fmaf(gs, v.z, old[6]); old[7] = fmaf(gs, v.w, old[7]);
        }
        float4* o4 = reinterpret_cast<float4*>(a.gX + g0 + r * 32 + 8 * q);
        o4[0] = make_float4(old[0], old[1], old[2], old[3]);
        o4[1] = make_float4(old[4], old[5], old[6], old[7]);
    }
}

// ---- temporal-difference loss of the value step (crowd_nav/utils/trainer.py:125-129) in one launch ---------------------
//   target = reward + gamma_bar * V_next        (two rounded fp32 operations, like the tensor expression :126)
//   loss  += sum_b (V - target)^2 * inv_count   (MSELoss(mean) over the GLOBAL batch: inv_count = 1 / global batch)
//   gV     = 2 (V - target) * inv_count         (what loss.backward() hands to the value head)
// `loss` is accumulated with one atomicAdd per block (zero it first).
__global__ void td_loss_kernel(const float* __restrict__ V, const float* __restrict__ reward, const float* __restrict__ Vnext, int B,
                               float gamma_bar, float inv_count, float* __restrict__ loss, float* __restrict__ gV) {
    __shared__ float part[8];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    float sq = 0.f;
    if (b < B) {
        const float target = __fadd_rn(reward[b], __fmul_rn(gamma_bar, Vnext[b]));
        const float d = V[b] - target;
        sq = d * d * inv_count;
        if (gV) gV[b] = 2.f * d * inv_count;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
        atomicAdd(loss, s);
    }
}

cudaError_t run_td_loss(const float* V, const float* reward, const float* Vnext, int B, float gamma_bar, float inv_count, float* loss,
                        float* gV, cudaStream_t st) {
    td_loss_kernel<<<(B + 255) / 256, 256, 0, st>>>(V, reward, Vnext, B, gamma_bar, inv_count, loss, gV);
    return cudaGetLastError();
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
    a.X0 = to_rows(nullptr); a.K0 = 0; a.dW0 = nullptr; a.db0 = nullptr;
    auto vec_ok = [](const Rows& r, int width) {
        return r.ptr != nullptr && (reinterpret_cast<uintptr_t>(r.ptr) & 15u) == 0 && (r.ld & 3) == 0 && (r.gstride & 3) == 0 && (width & 3) == 0;
    };
    a.vecG = vec_ok(a.G, N) && (a.mask.ptr == nullptr || vec_ok(a.mask, N));
    a.vecX = vec_ok(a.Xin, K);
    a.vecW = W != nullptr && w_layout == 0 && (K & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15u) == 0;
    // N = 32, K = 32 / 64: tcgen05 kernel (linear_bwd_tc.cu); other shapes: mma.sync 3xTF32 kernel below.
    // RGL_BWD_VARIANT (experiments only): t = tcgen05 wherever it applies, m = mma.sync for every shape, f = fp32-FMA kernel.
    static const char* variant = getenv("RGL_BWD_VARIANT");
    if (!variant || variant[0] == 't') {
        const cudaError_t e = run_linear_bwd_tc(a, num_sms, max_smem, st);
        if (e != cudaErrorNotSupported && e != cudaErrorInvalidConfiguration) return e;
    }
    if (!(variant && variant[0] == 'f')) {
        const cudaError_t e = dispatch_bwd_mma(a, num_sms, max_smem, st);
        if (e != cudaErrorNotSupported && e != cudaErrorInvalidConfiguration) return e;
    }
    a.ntiles = (R + LT - 1) / LT;
    const int NP4 = (N + 3) & ~3, NP32 = (N + 31) & ~31, KP32 = (K + 31) & ~31;
    size_t smem = ((size_t)NP4 * (KP32 + 4) + (size_t)LT * (NP32 + 4) + (size_t)LT * (KP32 + 4)) * sizeof(float);
    if (dW && smem < 8 * 1024 * sizeof(float)) smem = 8 * 1024 * sizeof(float);      // row-split reduction scratch
    if (smem > max_smem) return cudaErrorInvalidConfiguration;
    if (cudaError_t e = ensure_dyn_smem(rows_linear_bwd_kernel, (int)max_smem)) return e;
    // several CTAs per SM hide the global-load latency of the staging phase (the tiles are small)
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);      // register budget: __launch_bounds__(256, 2)
    const int cap = num_sms * per_sm;
    const int grid = a.ntiles < cap ? a.ntiles : cap;
    rows_linear_bwd_kernel<<<grid, 256, smem, st>>>(a);
    return cudaGetLastError();
}

// Both Linear layers of a two-layer embedding MLP (w_r / w_h: K0 -> 64 -> 32, relu after each) in one launch:
//   G = gX . (X > 0) [R,32];  dW1 += G^T hidden, db1 += colsum G;  Hg = (G W1) . (hidden > 0) (never written);
//   dW0 += Hg^T x0, db0 += colsum Hg.
cudaError_t run_mlp2_bwd(const RglRows* G, const RglRows* mask, const RglRows* hidden, const float* W1, const RglRows* X0, int K0,
                         float* dW1, float* db1, float* dW0, float* db0, int R, int num_sms, size_t max_smem, cudaStream_t st) {
    LinBwdArgs a;
    a.G = to_rows(G); a.mask = to_rows(mask); a.Xin = to_rows(hidden); a.Gin = to_rows(nullptr);
    a.N = 32; a.K = 64; a.R = R; a.W = W1; a.w_layout = 0; a.accumulate = 0; a.dW = dW1; a.db = db1;
    a.X0 = to_rows(X0); a.K0 = K0; a.dW0 = dW0; a.db0 = db0;
    auto vec_ok = [](const Rows& r, int width) {
        return r.ptr != nullptr && (reinterpret_cast<uintptr_t>(r.ptr) & 15u) == 0 && (r.ld & 3) == 0 && (r.gstride & 3) == 0 && (width & 3) == 0;
    };
    a.vecG = vec_ok(a.G, 32) && (a.mask.ptr == nullptr || vec_ok(a.mask, 32));
    a.vecX = vec_ok(a.Xin, 64);
    a.vecW = (reinterpret_cast<uintptr_t>(W1) & 15u) == 0;
    return dispatch_bwd_mma(a, num_sms, max_smem, st);
}

cudaError_t run_attn_layer_bwd(const float* A, const float* Hprev, const float* gM, const float* gH, int skip, float* gHprev,
                               float* gA, int accumulate_gA, int B, int n, const float* mask, int up_rows, cudaStream_t st) {
    const long long threads = (long long)B * n * 4;
    attn_layer_bwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(A, Hprev, gM, gH, skip, gHprev, gA, accumulate_gA, B, n, mask, up_rows);
    return cudaGetLastError();
}

cudaError_t run_attn_sim_bwd(const float* A, const float* Z, const float* gM, const float* mask, int up_rows, const float* gA_in,
                             float* gZ, float* gA_out, const float* X, const float* Y, float* gY, float* gX, int gx_accumulate,
                             int B, int n, size_t max_smem, cudaStream_t st) {
    AttnSimArgs a;
    a.A = A; a.Z = Z; a.gM = gM; a.mask = mask; a.gA_in = gA_in; a.X = X; a.Y = Y;
    a.gZ = gZ; a.gA_out = gA_out; a.gY = gY; a.gX = gX;
    a.up_rows = up_rows; a.gx_accumulate = gx_accumulate; a.B = B; a.n = n;
    const bool sim = X != nullptr;
    a.spb = 64 / n > 0 ? 64 / n : 1;              // spb * n rows x 4 threads <= 256
    const int threads = ((a.spb * n * 4 + 31) / 32) * 32;
    const size_t smem = ((size_t)a.spb * n * 32 * (sim ? 5 : 3) + (size_t)a.spb * n * n * (sim ? 3 : 2)) * sizeof(float);
    if (smem > max_smem) return cudaErrorInvalidConfiguration;
    const int grid = (B + a.spb - 1) / a.spb;
#define RGL_AS_LAUNCH(SIMV, NTV) do { \
        if (cudaError_t e = ensure_dyn_smem(attn_sim_bwd_kernel<SIMV, NTV>, (int)max_smem)) return e; \
        attn_sim_bwd_kernel<SIMV, NTV><<<grid, threads, smem, st>>>(a); } while (0)
    if (sim) {
        if (n == 6) RGL_AS_LAUNCH(true, 6); else if (n == 11) RGL_AS_LAUNCH(true, 11); else if (n == 21) RGL_AS_LAUNCH(true, 21);
        else RGL_AS_LAUNCH(true, 0);
    } else {
        if (n == 6) RGL_AS_LAUNCH(false, 6); else if (n == 11) RGL_AS_LAUNCH(false, 11); else if (n == 21) RGL_AS_LAUNCH(false, 21);
        else RGL_AS_LAUNCH(false, 0);
    }
#undef RGL_AS_LAUNCH
    return cudaGetLastError();
}

cudaError_t run_sim_bwd(const float* A, const float* gA, const float* X, const float* Y, float* gY, float* gX, int B, int n,
                        cudaStream_t st) {
    const int spb = 64 / n > 0 ? 64 / n : 1;                 // states per CTA: spb * n rows x 4 threads <= 256 (n <= 32: <= 128 threads for spb = 1)
    const int grid = (B + spb - 1) / spb;
    sim_bwd_kernel<<<grid, spb * n * 4, (size_t)spb * n * n * sizeof(float), st>>>(A, gA, X, Y, gY, gX, B, n, spb);
    return cudaGetLastError();
}

}  // namespace rgl
