// Fused RGL graph forward for sm_100a: embedding MLPs -> similarity softmax -> num_layer GCN layers
// (-> optional state-predictor head), one launch, everything between the 136-byte input state and the
// requested outputs stays in shared memory / registers.
//
// Replaces crowd_nav/policy/graph_model.py:99-130 (RGL.forward), :63-66 (embedded_gaussian similarity)
// and the head of crowd_nav/policy/state_predictor.py:28,36.
//
// Work decomposition
//   CTA      = a tile of TS states (persistent loop over tiles), weights resident in shared memory
//              (one TMA bulk copy of the packed blob per CTA), next tile's raw states prefetched by TMA
//              while the current tile computes.
//   rows     = graph nodes of the tile, agent-major: row = agent*TS + state, so a 16-row block is one
//              agent of 16 consecutive states (uniform weights, robot vs human) and the robot rows are
//              rows [0,TS).
//   GEMMs    = per-warp register tiles (common.cuh tile_gemm_pf: fully unrolled, operands of the next
//              k-step prefetched into registers) over row-major, stride-36 smem rows.
//   per-state= similarity row + softmax and A.H.  With the node count known at compile time (N = 6, 11, 21:
//              Nh = 5, 10, 20) a node row is handled by a lane PAIR (each lane half of the logits / half of
//              the columns, partner exchange by shuffle); N = 0 is the generic run-time-n path.
#include <stdlib.h>
#include "kernels.h"

namespace rgl {

template <int NCOL>
__device__ __forceinline__ void ah_row(const float* __restrict__ AB, const float* __restrict__ XB, float* __restrict__ YB,
                                       int r, int i, int s, int n, int TS, int c0) {
    float acc[NCOL];
#pragma unroll
    for (int c = 0; c < NCOL; ++c) acc[c] = 0.f;
    const float* arow = AB + (i * n) * TS + s;
    for (int j = 0; j < n; ++j) {
        const float aij = arow[j * TS];
        const float* h = XB + (j * TS + s) * LDX + c0;
#pragma unroll
        for (int c4 = 0; c4 < NCOL / 4; ++c4) {
            const float4 hv = lds128(h + 4 * c4);
            acc[4 * c4 + 0] = fmaf(aij, hv.x, acc[4 * c4 + 0]);
            acc[4 * c4 + 1] = fmaf(aij, hv.y, acc[4 * c4 + 1]);
            acc[4 * c4 + 2] = fmaf(aij, hv.z, acc[4 * c4 + 2]);
            acc[4 * c4 + 3] = fmaf(aij, hv.w, acc[4 * c4 + 3]);
        }
    }
    float* y = YB + r * LDX + c0;
#pragma unroll
    for (int c4 = 0; c4 < NCOL / 4; ++c4)
        sts128(y + 4 * c4, make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]));
}

template <int RT>
__device__ __forceinline__ void store_tile(float* base, int cg, const float (&acc)[RT][8]) {
#pragma unroll
    for (int q = 0; q < RT; ++q)
#pragma unroll
        for (int m = 0; m < 2; ++m)
            sts128(base + q * 8 * LDX + cg * 4 + 16 * m,
                   make_float4(acc[q][4 * m], acc[q][4 * m + 1], acc[q][4 * m + 2], acc[q][4 * m + 3]));
}
template <int RT>
__device__ __forceinline__ void store_tile_relu(float* base, int cg, const float (&acc)[RT][8]) {
#pragma unroll
    for (int q = 0; q < RT; ++q)
#pragma unroll
        for (int m = 0; m < 2; ++m)
            sts128(base + q * 8 * LDX + cg * 4 + 16 * m,
                   make_float4(fmaxf(acc[q][4 * m], 0.f), fmaxf(acc[q][4 * m + 1], 0.f), fmaxf(acc[q][4 * m + 2], 0.f),
                               fmaxf(acc[q][4 * m + 3], 0.f)));
}
template <int RT>
__device__ __forceinline__ void zero_tile(float (&acc)[RT][8]) {
#pragma unroll
    for (int q = 0; q < RT; ++q)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[q][c] = 0.f;
}
template <int RT>
__device__ __forceinline__ void bias_tile(float (&acc)[RT][8], const float* b, int cg) {
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const float4 v = lds128(b + cg * 4 + 16 * m);
#pragma unroll
        for (int q = 0; q < RT; ++q) {
            acc[q][4 * m + 0] = v.x; acc[q][4 * m + 1] = v.y; acc[q][4 * m + 2] = v.z; acc[q][4 * m + 3] = v.w;
        }
    }
}

// launch bounds: RPT > 0 (row-per-thread per-state phases, RT = 4) runs n warps per CTA and two CTAs per SM when
// shared memory allows (N = 6); the other variants run one CTA per SM.
template <int TS, int RT, int N, int RPT, bool MMA = false>
__global__ void __launch_bounds__(RPT > 0 ? N * 32 : (N == 6 ? 384 : 512), (RPT > 0 && N <= 8) ? 2 : 1)
graph_forward_kernel(const GraphArgs a) {
    constexpr int MT = RT / 2;                // 16-row m-tiles per warp row block on the tensor-core path
    constexpr int RB = 8 * RT;
    static_assert(TS % RB == 0, "a row block must not straddle two agents");
    extern __shared__ __align__(128) float smem[];

    const int n = N > 0 ? N : a.Nh + 1;
    const int Nh = n - 1;
    const int R = n * TS;
    const int nrb = R / RB;
    const int gwf = graph_floats(a.L);
    constexpr int NP = (N + 3) & ~3;          // attention-row length rounded to float4
    constexpr int NPS = NP + 4;               // its smem stride (bank-conflict-free for 8 consecutive rows)

    // ---- shared memory carve-up (all offsets multiples of 4 floats) ----
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem);
    uint64_t* bar_in = bar_w + 1;
    float* gw = smem + 4;
    float* mw = gw + gwf;
    float* XB = mw + (a.mw ? MOTION_FLOATS : 0);
    float* YB = XB + R * LDX;
    float* AB = YB + R * LDX;                 // generic: [n*n][TS]; compile-time N: [R][NP]
    float* rawR = AB + (N > 0 ? R * NPS : n * n * TS);
    float* rawH = rawR + TS * RD;             // TS*9 is a multiple of 4 for TS in {16,32}

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int rg = lane & 7, cg = lane >> 3;
    const bool skip = a.flags & RGL_FLAG_SKIP, layerwise = a.flags & RGL_FLAG_LAYERWISE;

    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_in, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = gwf * 4u + (a.mw ? MOTION_FLOATS * 4u : 0u);
        mbar_arrive_expect_tx(bar_w, bytes);
        bulk_g2s(gw, a.gw, gwf * 4u, bar_w);
        if (a.mw) bulk_g2s(mw, a.mw, MOTION_FLOATS * 4u, bar_w);
    }

    // stage the raw states of tile t (TMA bulk copy when the tile is full and aligned, else plain loads)
    auto load_tile = [&](int t) {
        const int s0 = t * TS;
        const int cnt = min(TS, a.B - s0);
        if (a.use_tma && cnt == TS) {
            if (tid == 0) {
                fence_proxy_async();
                mbar_arrive_expect_tx(bar_in, (uint32_t)(TS * RD + TS * Nh * HD) * 4u);
                bulk_g2s(rawR, a.robot + (size_t)s0 * RD, TS * RD * 4u, bar_in);
                bulk_g2s(rawH, a.humans + (size_t)s0 * Nh * HD, (uint32_t)(TS * Nh * HD) * 4u, bar_in);
            }
        } else {
            for (int idx = tid; idx < TS * RD; idx += blockDim.x) {
                const int s = idx / RD;
                rawR[idx] = s < cnt ? __ldg(a.robot + (size_t)s0 * RD + idx) : 0.f;
            }
            const int hw = Nh * HD;
            for (int idx = tid; idx < TS * hw; idx += blockDim.x) {
                const int s = idx / hw, rem = idx - s * hw;
                rawH[idx] = s < cnt ? __ldg(a.humans + (size_t)((s0 + s) / a.hb) * hw + rem) : 0.f;
            }
        }
    };

    uint32_t in_parity = 0;
    if ((int)blockIdx.x < a.ntiles) load_tile(blockIdx.x);
    mbar_wait(bar_w, 0);

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int s0 = tile * TS;
        const int cnt = min(TS, a.B - s0);
        if (a.use_tma && cnt == TS) {
            mbar_wait(bar_in, in_parity);
            in_parity ^= 1;
        }
        __syncthreads();

        // ================= embedding: X = relu(W1 relu(W0 x + b0) + b1); Y = X w_a =================
        for (int rb = warp; rb < nrb; rb += nwarps) {
            const int r0 = rb * RB;
            const int agent = r0 / TS;
            const int sb = r0 - agent * TS;
            const float *W0, *B0, *W1, *B1;
            int K0;
            const float* xrow[RT];
            if (agent == 0) {
                W0 = gw + G_WR0; B0 = gw + G_BR0; W1 = gw + G_WR1; B1 = gw + G_BR1; K0 = RD;
#pragma unroll
                for (int q = 0; q < RT; ++q) xrow[q] = rawR + (sb + rg + 8 * q) * RD;
            } else {
                W0 = gw + G_WH0; B0 = gw + G_BH0; W1 = gw + G_WH1; B1 = gw + G_BH1; K0 = HD;
#pragma unroll
                for (int q = 0; q < RT; ++q) xrow[q] = rawH + ((sb + rg + 8 * q) * Nh + (agent - 1)) * HD;
            }
            float* scr = YB + (r0 + rg) * LDX;       // this warp's own rows of YB double as the hidden scratch
            float* xo = XB + (r0 + rg) * LDX;
            if constexpr (MMA) {
                // layer 2 (64 -> 32) and Y = X w_a on the tensor cores (3xTF32); layer 1 (K = 5 / 9) stays on the FMA pipe
                float macc[MT][4][4];
                cfrag_fill<MT>(macc, B1, lane);
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    float acc1[RT][8];
                    bias_tile<RT>(acc1, B0 + half * 32, cg);
                    tile_gemm_smallk<RT, 2>(acc1, xrow, W0 + half * 32 + cg * 4, HID, K0);
                    store_tile_relu<RT>(scr, cg, acc1);
                    __syncwarp();
                    mma_gemm_3xtf32<4, MT>(macc, YB + r0 * LDX, W1 + half * 32 * LDW, lane);
                    __syncwarp();
                }
                cfrag_store<true, MT>(XB + r0 * LDX, macc, lane);
                __syncwarp();
                cfrag_fill<MT>(macc, nullptr, lane);
                mma_gemm_3xtf32<4, MT>(macc, XB + r0 * LDX, gw + G_WA, lane);
                cfrag_store<false, MT>(YB + r0 * LDX, macc, lane);
            } else {
            float acc2[RT][8];
            bias_tile<RT>(acc2, B1, cg);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                float acc1[RT][8];
                bias_tile<RT>(acc1, B0 + half * 32, cg);
                tile_gemm_smallk<RT, 2>(acc1, xrow, W0 + half * 32 + cg * 4, HID, K0);
                store_tile_relu<RT>(scr, cg, acc1);
                if (a.save) {
#pragma unroll
                    for (int q = 0; q < RT; ++q) {
                        const int s = sb + rg + 8 * q;
                        if (s < cnt) {
                            float* dst = agent == 0 ? a.sv.a1r + (size_t)(s0 + s) * HID
                                                    : a.sv.a1h + ((size_t)(s0 + s) * Nh + (agent - 1)) * HID;
#pragma unroll
                            for (int m = 0; m < 2; ++m)
                                *reinterpret_cast<float4*>(dst + half * 32 + cg * 4 + 16 * m) =
                                    make_float4(fmaxf(acc1[q][4 * m], 0.f), fmaxf(acc1[q][4 * m + 1], 0.f),
                                                fmaxf(acc1[q][4 * m + 2], 0.f), fmaxf(acc1[q][4 * m + 3], 0.f));
                        }
                    }
                }
                __syncwarp();
                tile_gemm_pf<RT, 2, 32>(acc2, scr, LDX, W1 + half * 32 * LDW + cg * 4, LDW);
                __syncwarp();
            }
            store_tile_relu<RT>(xo, cg, acc2);
            __syncwarp();
            float accy[RT][8];
            zero_tile<RT>(accy);
            tile_gemm_pf<RT, 2, XD>(accy, xo, LDX, gw + G_WA + cg * 4, LDW);
            store_tile<RT>(scr, cg, accy);
            }
        }
        __syncthreads();

        // raw inputs are dead: prefetch the next tile's states under this tile's GCN layers
        if (tile + (int)gridDim.x < a.ntiles) load_tile(tile + gridDim.x);

        // training forward: copy a [R][36] smem buffer (agent-major rows) to a state-major [B,n,32] HBM tensor
        auto dump_rows = [&](const float* buf, float* dst) {
            for (int idx = tid; idx < R * 8; idx += blockDim.x) {
                const int r = idx >> 3, c4 = idx & 7;
                const int i = r / TS, s = r - i * TS;
                if (s < cnt) *reinterpret_cast<float4*>(dst + ((size_t)(s0 + s) * n + i) * XD + 4 * c4) = lds128(buf + r * LDX + 4 * c4);
            }
        };
        if (a.save) {
            dump_rows(XB, a.sv.X);
            dump_rows(YB, a.sv.Y);
        }

        // ================= GCN layers =================
        for (int l = 0; l < a.L; ++l) {
            const bool last = (l == a.L - 1);
            const bool robot_only = last && a.H == nullptr && a.S == nullptr && !a.save;
            const int rows = robot_only ? TS : R;        // node rows that must be produced by this layer

            if (l == 0 || layerwise) {
                if (l > 0) {                              // Y = H w_a for the layerwise graph
                    for (int rb = warp; rb * RB < rows; rb += nwarps) {
                        if constexpr (MMA) {
                            float macc[MT][4][4];
                            cfrag_fill<MT>(macc, nullptr, lane);
                            mma_gemm_3xtf32<4, MT>(macc, XB + rb * RB * LDX, gw + G_WA, lane);
                            cfrag_store<false, MT>(YB + rb * RB * LDX, macc, lane);
                        } else {
                            float accy[RT][8];
                            zero_tile<RT>(accy);
                            tile_gemm_pf<RT, 2, XD>(accy, XB + (rb * RB + rg) * LDX, LDX, gw + G_WA + cg * 4, LDW);
                            store_tile<RT>(YB + (rb * RB + rg) * LDX, cg, accy);
                        }
                    }
                    __syncthreads();
                }
                // ---- similarity row + softmax: A[i][:] = softmax_j( Y[i] . X[j] ) ----
                if constexpr (RPT > 0) {
                    // lane pair = RPT node rows of one state; each lane computes half of the N logits of those rows
                    // (every X_j row load is reused RPT times), partners swap halves by shuffle
                    constexpr int NG = N / RPT, JH = (N + 1) / 2;
                    static_assert(N % RPT == 0, "rows-per-thread must divide the node count");
                    for (int t = tid; t < NG * TS * 2; t += blockDim.x) {
                        const int half = t & 1, u = t >> 1;
                        const int ip = u / TS, s = u - ip * TS;
                        if (robot_only && ip > 0) continue;        // warp-uniform: TS*2 is a multiple of 32
                        const int rot = half * 4;
                        float4 y[RPT][8];
#pragma unroll
                        for (int q = 0; q < RPT; ++q)
#pragma unroll
                            for (int c = 0; c < 8; ++c) y[q][c] = lds128(YB + ((ip * RPT + q) * TS + s) * LDX + 4 * ((c + rot) & 7));
                        float lg[RPT][JH], ot[RPT][JH];
#pragma unroll
                        for (int jj = 0; jj < JH; ++jj) {
                            const int j = half * JH + jj;
                            float d[RPT][4];
#pragma unroll
                            for (int q = 0; q < RPT; ++q) d[q][0] = d[q][1] = d[q][2] = d[q][3] = 0.f;
                            if (j < N) {
                                const float* x = XB + (j * TS + s) * LDX;
#pragma unroll
                                for (int c = 0; c < 8; ++c) {
                                    const float4 xv = lds128(x + 4 * ((c + rot) & 7));
#pragma unroll
                                    for (int q = 0; q < RPT; ++q) {
                                        d[q][0] = fmaf(y[q][c].x, xv.x, d[q][0]); d[q][1] = fmaf(y[q][c].y, xv.y, d[q][1]);
                                        d[q][2] = fmaf(y[q][c].z, xv.z, d[q][2]); d[q][3] = fmaf(y[q][c].w, xv.w, d[q][3]);
                                    }
                                }
                            }
#pragma unroll
                            for (int q = 0; q < RPT; ++q) lg[q][jj] = (d[q][0] + d[q][1]) + (d[q][2] + d[q][3]);
                        }
#pragma unroll
                        for (int q = 0; q < RPT; ++q)
#pragma unroll
                            for (int jj = 0; jj < JH; ++jj) ot[q][jj] = __shfl_xor_sync(0xffffffffu, lg[q][jj], 1);
#pragma unroll
                        for (int q = 0; q < RPT; ++q) {
                            const int i = ip * RPT + q;
                            float p[NP];
#pragma unroll
                            for (int j = 0; j < NP; ++j) {
                                const int jj = j < JH ? j : j - JH;
                                const bool mine = (j < JH) == (half == 0);
                                p[j] = j < N ? (mine ? lg[q][jj] : ot[q][jj]) : -INFINITY;
                            }
                            float mx = p[0];
#pragma unroll
                            for (int j = 1; j < N; ++j) mx = fmaxf(mx, p[j]);
                            float sum = 0.f;
#pragma unroll
                            for (int j = 0; j < N; ++j) { p[j] = expf(p[j] - mx); sum += p[j]; }
#pragma unroll
                            for (int j = 0; j < NP; ++j) p[j] = j < N ? p[j] / sum : 0.f;
                            if ((q & 1) == half) {                 // the pair shares the write-back work
#pragma unroll
                                for (int j4 = 0; j4 < NP / 4; ++j4)
                                    sts128(AB + (i * TS + s) * NPS + 4 * j4, make_float4(p[4 * j4], p[4 * j4 + 1], p[4 * j4 + 2], p[4 * j4 + 3]));
                                if (a.A0 != nullptr && l == 0 && (s0 + s) == 0) {
#pragma unroll
                                    for (int j = 0; j < N; ++j) a.A0[i * N + j] = p[j];
                                }
                            }
                        }
                    }
                } else if constexpr (N > 0) {
                    // lane pair per node row: lane&1 selects which half of the N logits this lane computes
                    constexpr int JH = (N + 1) / 2;
                    for (int task = warp; task * 16 < rows; task += nwarps) {
                        const int r = task * 16 + (lane >> 1), half = lane & 1;
                        const int i = r / TS, s = r - i * TS;
                        // the odd lane walks the 16-byte chunks rotated by 4 so that the pair never hits the same banks
                        const int rot = half * 4;
                        float4 y[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) y[c] = lds128(YB + r * LDX + 4 * ((c + rot) & 7));
                        float lg[JH], ot[JH];
#pragma unroll
                        for (int jj = 0; jj < JH; ++jj) {
                            const int j = half * JH + jj;
                            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
                            if (j < N) {
                                const float* x = XB + (j * TS + s) * LDX;
#pragma unroll
                                for (int c = 0; c < 8; ++c) {
                                    const float4 xv = lds128(x + 4 * ((c + rot) & 7));
                                    d0 = fmaf(y[c].x, xv.x, d0); d1 = fmaf(y[c].y, xv.y, d1);
                                    d2 = fmaf(y[c].z, xv.z, d2); d3 = fmaf(y[c].w, xv.w, d3);
                                }
                            }
                            lg[jj] = (d0 + d1) + (d2 + d3);
                        }
#pragma unroll
                        for (int jj = 0; jj < JH; ++jj) ot[jj] = __shfl_xor_sync(0xffffffffu, lg[jj], 1);
                        float p[NP];
#pragma unroll
                        for (int j = 0; j < NP; ++j) {
                            const int jj = j < JH ? j : j - JH;
                            const bool mine = (j < JH) == (half == 0);
                            p[j] = j < N ? (mine ? lg[jj] : ot[jj]) : -INFINITY;
                        }
                        float mx = p[0];
#pragma unroll
                        for (int j = 1; j < N; ++j) mx = fmaxf(mx, p[j]);
                        float sum = 0.f;
#pragma unroll
                        for (int j = 0; j < N; ++j) { p[j] = expf(p[j] - mx); sum += p[j]; }
#pragma unroll
                        for (int j = 0; j < NP; ++j) p[j] = j < N ? p[j] / sum : 0.f;
                        if (half == 0) {
#pragma unroll
                            for (int j4 = 0; j4 < NP / 4; ++j4)
                                sts128(AB + r * NPS + 4 * j4, make_float4(p[4 * j4], p[4 * j4 + 1], p[4 * j4 + 2], p[4 * j4 + 3]));
                            if (a.A0 != nullptr && l == 0 && (s0 + s) == 0) {
#pragma unroll
                                for (int j = 0; j < N; ++j) a.A0[i * N + j] = p[j];
                            }
                        }
                    }
                } else {
                    for (int r = warp * 32 + lane; r < rows; r += nwarps * 32) {
                        const int i = r / TS, s = r - i * TS;
                        float4 y[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) y[c] = lds128(YB + r * LDX + 4 * c);
                        float* arow = AB + (i * n) * TS + s;
                        float mx = -INFINITY;
                        for (int j = 0; j < n; ++j) {
                            const float* x = XB + (j * TS + s) * LDX;
                            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const float4 xv = lds128(x + 4 * c);
                                d0 = fmaf(y[c].x, xv.x, d0); d1 = fmaf(y[c].y, xv.y, d1);
                                d2 = fmaf(y[c].z, xv.z, d2); d3 = fmaf(y[c].w, xv.w, d3);
                            }
                            const float d = (d0 + d1) + (d2 + d3);
                            arow[j * TS] = d;
                            mx = fmaxf(mx, d);
                        }
                        float sum = 0.f;
                        for (int j = 0; j < n; ++j) {
                            const float e = expf(arow[j * TS] - mx);
                            arow[j * TS] = e;
                            sum += e;
                        }
                        const bool writeA0 = a.A0 != nullptr && l == 0 && (s0 + s) == 0;
                        for (int j = 0; j < n; ++j) {
                            const float p = arow[j * TS] / sum;
                            arow[j * TS] = p;
                            if (writeA0) a.A0[i * n + j] = p;
                        }
                    }
                }
                __syncthreads();
                if (a.save && l == 0) {
                    for (int idx = tid; idx < R * n; idx += blockDim.x) {
                        const int r = idx / n, j = idx - r * n;
                        const int i = r / TS, s = r - i * TS;
                        const float v = N > 0 ? AB[r * NPS + j] : AB[(i * n + j) * TS + s];
                        if (s < cnt) a.sv.A[((size_t)(s0 + s) * n + i) * n + j] = v;
                    }
                }
            }

            // ---- AH = A . H  (per state) ----
            if constexpr (RPT > 0) {
                constexpr int NG = N / RPT;
                for (int t = tid; t < NG * TS * 2; t += blockDim.x) {
                    const int half = t & 1, u = t >> 1;
                    const int ip = u / TS, s = u - ip * TS;
                    if (robot_only && ip > 0) continue;
                    const int c0 = half * 16;
                    float p[RPT][NP];
#pragma unroll
                    for (int q = 0; q < RPT; ++q)
#pragma unroll
                        for (int j4 = 0; j4 < NP / 4; ++j4) {
                            const float4 v = lds128(AB + ((ip * RPT + q) * TS + s) * NPS + 4 * j4);
                            p[q][4 * j4] = v.x; p[q][4 * j4 + 1] = v.y; p[q][4 * j4 + 2] = v.z; p[q][4 * j4 + 3] = v.w;
                        }
                    float acc[RPT][16];
#pragma unroll
                    for (int q = 0; q < RPT; ++q)
#pragma unroll
                        for (int c = 0; c < 16; ++c) acc[q][c] = 0.f;
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        const float* h = XB + (j * TS + s) * LDX + c0;
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            const float4 hv = lds128(h + 4 * c4);
#pragma unroll
                            for (int q = 0; q < RPT; ++q) {
                                acc[q][4 * c4 + 0] = fmaf(p[q][j], hv.x, acc[q][4 * c4 + 0]);
                                acc[q][4 * c4 + 1] = fmaf(p[q][j], hv.y, acc[q][4 * c4 + 1]);
                                acc[q][4 * c4 + 2] = fmaf(p[q][j], hv.z, acc[q][4 * c4 + 2]);
                                acc[q][4 * c4 + 3] = fmaf(p[q][j], hv.w, acc[q][4 * c4 + 3]);
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < RPT; ++q)
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4)
                            sts128(YB + ((ip * RPT + q) * TS + s) * LDX + c0 + 4 * c4,
                                   make_float4(acc[q][4 * c4], acc[q][4 * c4 + 1], acc[q][4 * c4 + 2], acc[q][4 * c4 + 3]));
                }
            } else if constexpr (N > 0) {
                for (int task = warp; task * 16 < rows; task += nwarps) {
                    const int r = task * 16 + (lane >> 1), c0 = (lane & 1) * 16;
                    const int i = r / TS, s = r - i * TS;
                    float p[NP];
#pragma unroll
                    for (int j4 = 0; j4 < NP / 4; ++j4) {
                        const float4 v = lds128(AB + r * NPS + 4 * j4);
                        p[4 * j4] = v.x; p[4 * j4 + 1] = v.y; p[4 * j4 + 2] = v.z; p[4 * j4 + 3] = v.w;
                    }
                    float acc[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        const float* h = XB + (j * TS + s) * LDX + c0;
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            const float4 hv = lds128(h + 4 * c4);
                            acc[4 * c4 + 0] = fmaf(p[j], hv.x, acc[4 * c4 + 0]);
                            acc[4 * c4 + 1] = fmaf(p[j], hv.y, acc[4 * c4 + 1]);
                            acc[4 * c4 + 2] = fmaf(p[j], hv.z, acc[4 * c4 + 2]);
                            acc[4 * c4 + 3] = fmaf(p[j], hv.w, acc[4 * c4 + 3]);
                        }
                    }
                    (void)i;
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4)
                        sts128(YB + r * LDX + c0 + 4 * c4, make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]));
                }
            } else {
                const int chunks = (rows + 31) / 32;
                if (chunks >= nwarps) {
                    for (int r = warp * 32 + lane; r < rows; r += nwarps * 32) {
                        const int i = r / TS;
                        ah_row<32>(AB, XB, YB, r, i, r - i * TS, n, TS, 0);
                    }
                } else {
                    for (int it = warp; it < chunks * 2; it += nwarps) {
                        const int r = (it >> 1) * 32 + lane;
                        if (r < rows) {
                            const int i = r / TS;
                            ah_row<16>(AB, XB, YB, r, i, r - i * TS, n, TS, (it & 1) * 16);
                        }
                    }
                }
            }
            __syncthreads();
            if (a.save) dump_rows(YB, a.sv.M[l]);

            // ---- H' = relu(AH . W_l) (+ H), in place over XB; last layer streams the outputs to HBM ----
            for (int rb = warp; rb * RB < rows; rb += nwarps) {
                const int r0 = rb * RB;
                const int agent = r0 / TS;
                float* xo = XB + (r0 + rg) * LDX;
                if constexpr (MMA) {
                    float macc[MT][4][4];
                    cfrag_fill<MT>(macc, nullptr, lane);
                    mma_gemm_3xtf32<4, MT>(macc, YB + r0 * LDX, gw + G_WS + l * XD * LDW, lane);
                    const int g = lane >> 2, t = lane & 3;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const int rr = r0 + mt * 16 + g + 8 * hh;        // node row; s = its state inside the tile
                            const int s = rr - agent * TS;
#pragma unroll
                            for (int nt = 0; nt < 4; ++nt) {
                                float2 v = make_float2(fmaxf(macc[mt][nt][2 * hh], 0.f), fmaxf(macc[mt][nt][2 * hh + 1], 0.f));
                                float* p = XB + rr * LDX + nt * 8 + 2 * t;
                                if (skip) {
                                    const float2 h = *reinterpret_cast<const float2*>(p);
                                    v.x += h.x; v.y += h.y;
                                }
                                *reinterpret_cast<float2*>(p) = v;
                                if (last && s < cnt) {
                                    const size_t gs = (size_t)(s0 + s);
                                    if (a.H) *reinterpret_cast<float2*>(a.H + (gs * n + agent) * XD + nt * 8 + 2 * t) = v;
                                    if (a.E && agent == 0) *reinterpret_cast<float2*>(a.E + gs * XD + nt * 8 + 2 * t) = v;
                                }
                            }
                        }
                } else {
                float acc[RT][8];
                zero_tile<RT>(acc);
                tile_gemm_pf<RT, 2, XD>(acc, YB + (r0 + rg) * LDX, LDX, gw + G_WS + l * XD * LDW + cg * 4, LDW);
#pragma unroll
                for (int q = 0; q < RT; ++q) {
                    const int s = r0 + rg + 8 * q - agent * TS;
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        float4 v = make_float4(fmaxf(acc[q][4 * m], 0.f), fmaxf(acc[q][4 * m + 1], 0.f),
                                               fmaxf(acc[q][4 * m + 2], 0.f), fmaxf(acc[q][4 * m + 3], 0.f));
                        float* p = xo + q * 8 * LDX + cg * 4 + 16 * m;
                        if (a.save && s < cnt)
                            *reinterpret_cast<float4*>(a.sv.Rl[l] + ((size_t)(s0 + s) * n + agent) * XD + cg * 4 + 16 * m) = v;
                        if (skip) {
                            const float4 h = lds128(p);
                            v.x += h.x; v.y += h.y; v.z += h.z; v.w += h.w;
                        }
                        sts128(p, v);
                        if (a.save && s < cnt)
                            *reinterpret_cast<float4*>(a.sv.Hl[l] + ((size_t)(s0 + s) * n + agent) * XD + cg * 4 + 16 * m) = v;
                        if (last && s < cnt) {
                            const size_t gs = (size_t)(s0 + s);
                            if (a.H) *reinterpret_cast<float4*>(a.H + (gs * n + agent) * XD + cg * 4 + 16 * m) = v;
                            if (a.E && agent == 0) *reinterpret_cast<float4*>(a.E + gs * XD + cg * 4 + 16 * m) = v;
                        }
                    }
                }
                }
                // ---- state-predictor head on human rows: S = W1 relu(W0 h + b0) + b1 (32 -> 64 -> 5) ----
                if (last && a.S != nullptr && agent >= 1) {
                    __syncwarp();
                    constexpr int LPR = 32 / RB;          // lanes per row in the 64->5 dot (2 for RB=16, 1 for RB=32)
                    constexpr int KPL = 32 / LPR;         // k per lane per half
                    const int rl = lane % RB, kh = lane / RB;
                    float part[HD];
#pragma unroll
                    for (int c = 0; c < HD; ++c) part[c] = 0.f;
                    float* scr = YB + (r0 + rg) * LDX;
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        float acc1[RT][8];
                        bias_tile<RT>(acc1, mw + M_B0 + half * 32, cg);
                        tile_gemm_pf<RT, 2, XD>(acc1, xo, LDX, mw + M_W0 + half * 32 + cg * 4, MH);
                        store_tile_relu<RT>(scr, cg, acc1);
                        if (a.save && a.sv.mh) {
#pragma unroll
                            for (int q = 0; q < RT; ++q) {
                                const int s = r0 + rg + 8 * q - agent * TS;
                                if (s < cnt) {
                                    float* dst = a.sv.mh + ((size_t)(s0 + s) * Nh + (agent - 1)) * MH + half * 32 + cg * 4;
#pragma unroll
                                    for (int m = 0; m < 2; ++m)
                                        *reinterpret_cast<float4*>(dst + 16 * m) =
                                            make_float4(fmaxf(acc1[q][4 * m], 0.f), fmaxf(acc1[q][4 * m + 1], 0.f),
                                                        fmaxf(acc1[q][4 * m + 2], 0.f), fmaxf(acc1[q][4 * m + 3], 0.f));
                                }
                            }
                        }
                        __syncwarp();
                        const float* hrow = YB + (r0 + rl) * LDX + kh * KPL;
#pragma unroll
                        for (int k4 = 0; k4 < KPL / 4; ++k4) {
                            const float4 hv = lds128(hrow + 4 * k4);
#pragma unroll
                            for (int c = 0; c < HD; ++c) {
                                const float4 wv = lds128(mw + M_W1 + c * MH + half * 32 + kh * KPL + 4 * k4);
                                part[c] = fmaf(hv.x, wv.x, part[c]); part[c] = fmaf(hv.y, wv.y, part[c]);
                                part[c] = fmaf(hv.z, wv.z, part[c]); part[c] = fmaf(hv.w, wv.w, part[c]);
                            }
                        }
                        __syncwarp();
                    }
                    if (LPR == 2) {
#pragma unroll
                        for (int c = 0; c < HD; ++c) part[c] += __shfl_xor_sync(0xffffffffu, part[c], 16);
                    }
                    const int s = r0 + rl - agent * TS;
                    if (kh == 0 && s < cnt) {
                        float* so = a.S + ((size_t)(s0 + s) * Nh + (agent - 1)) * HD;
#pragma unroll
                        for (int c = 0; c < HD; ++c) so[c] = part[c] + mw[M_B1 + c];
                    }
                }
            }
            if (!last) __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------
static size_t graph_smem_bytes(int TS, int Nh, int L, bool motion, bool ctn) {
    const int n = Nh + 1;
    const size_t ab = ctn ? (size_t)n * TS * (((n + 3) & ~3) + 4) : (size_t)n * n * TS;
    size_t fl = 4 + graph_floats(L) + (motion ? MOTION_FLOATS : 0) + 2 * (size_t)n * TS * LDX + ab + TS * RD +
                ((TS * Nh * HD + 3) & ~3);
    return fl * sizeof(float);
}

template <int TS, int RT, int N, int RPT, bool MMA = false>
static cudaError_t launch_graph(const GraphArgs& a, int num_sms, size_t max_smem, cudaStream_t st) {
    const int n = a.Nh + 1;
    const size_t smem = graph_smem_bytes(TS, a.Nh, a.L, a.mw != nullptr, N > 0);
    if (smem > max_smem) return cudaErrorInvalidConfiguration;
    GraphArgs b = a;
    b.ntiles = (a.B + TS - 1) / TS;
    const int nrb = n * TS / (8 * RT);
    const int cap = RPT > 0 ? n : (N == 6 ? 12 : 16);
    const int nwarps = nrb < cap ? nrb : cap;
    if (cudaError_t e = ensure_dyn_smem(graph_forward_kernel<TS, RT, N, RPT, MMA>, (int)max_smem)) return e;
    int per_sm = (int)((228 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    const int max_cta = (RPT > 0 && N <= 8) ? 2 : 1;          // matches the kernel's __launch_bounds__ (register budget)
    if (per_sm > max_cta) per_sm = max_cta;
    const int grid = b.ntiles < num_sms * per_sm ? b.ntiles : num_sms * per_sm;
    graph_forward_kernel<TS, RT, N, RPT, MMA><<<grid, nwarps * 32, smem, st>>>(b);
    return cudaGetLastError();
}

template <int TS, int RT, int N, int RPT>
static cudaError_t launch_pick(const GraphArgs& a, bool mma, int num_sms, size_t max_smem, cudaStream_t st) {
    return mma ? launch_graph<TS, RT, N, RPT, true>(a, num_sms, max_smem, st) : launch_graph<TS, RT, N, RPT, false>(a, num_sms, max_smem, st);
}

template <int N>
static cudaError_t dispatch_tile(const GraphArgs& a, int num_sms, size_t max_smem, cudaStream_t st) {
    // Numerics: inference runs the shared-weight GEMMs on the tensor cores (mma.sync, 3xTF32 split, ~3e-6 relative to
    // fp32) unless the caller sets RGL_FLAG_FP32_FMA; the training forward (activation saves) always uses fp32 FFMA.
    // RGL_GRAPH_VARIANT (experiments only): '2'/'4' = 16-/32-row FFMA tiles, 'n'/'m' = the same tiles with mma.
    static const char* force = getenv("RGL_GRAPH_VARIANT");
    const bool mma = (force ? (force[0] == 'm' || force[0] == 'n') : true) && !a.save && !(a.flags & RGL_FLAG_FP32_FMA);
    // pick the largest state tile that fits in shared memory; small batches prefer more CTAs
    const bool fits32 = graph_smem_bytes(32, a.Nh, a.L, a.mw != nullptr, N > 0) <= max_smem;
    const int tiles32 = (a.B + 31) / 32;
    if (fits32 && tiles32 >= num_sms / 2) {
        if constexpr (N == 6) {
            // beyond one wave of tiles (or multi-stream serving): 32-row tiles, 2 node rows per thread in the per-state
            // phases, two 6-warp CTAs per SM; up to one wave: 16-row tiles, 12 warps per CTA (lowest latency)
            const bool big = force ? (force[0] == '4' || force[0] == 'm') : (tiles32 > num_sms || (a.flags & RGL_FLAG_THROUGHPUT));
            if (big) return launch_pick<32, 4, 6, 2>(a, mma, num_sms, max_smem, st);
        }
        return launch_pick<32, 2, N, 0>(a, mma, num_sms, max_smem, st);
    }
    if (graph_smem_bytes(16, a.Nh, a.L, a.mw != nullptr, N > 0) <= max_smem) return launch_pick<16, 2, N, 0>(a, mma, num_sms, max_smem, st);
    return cudaErrorInvalidConfiguration;
}

cudaError_t run_graph_forward(const GraphArgs& a, int num_sms, size_t max_smem, cudaStream_t st) {
    switch (a.Nh + 1) {
        case 6: return dispatch_tile<6>(a, num_sms, max_smem, st);
        case 11: return dispatch_tile<11>(a, num_sms, max_smem, st);
        case 21: return dispatch_tile<21>(a, num_sms, max_smem, st);
        default: return dispatch_tile<0>(a, num_sms, max_smem, st);
    }
}

}  // namespace rgl
