// Fused RGL graph forward on the 5th-generation tensor cores (tcgen05 / TMEM) of sm_100a.
//
// Same contract as graph_forward.cu (crowd_nav/policy/graph_model.py:99-130 RGL.forward, :63-66 embedded_gaussian
// similarity, crowd_nav/policy/state_predictor.py:28,36 motion head), different machine mapping:
//
//   group   = 128 threads = one UMMA M-tile of 128 graph-node rows = SPT = 128 / n whole states.  A thread owns ONE node
//             row end to end (TMEM lane = row, tcgen05.ld/st 32x32b: thread t <-> lane t).  Groups are independent
//             pipelines (own mbarrier, own named barrier, own 128 TMEM columns, own 18 KB feature buffer) that only
//             share the weight tiles of their CTA: while one group waits for its MMAs the others compute.
//   rows    = robot-first inside a tile: rows [0,SPT) are the robots of the SPT states, rows SPT + s*Nh + j the humans,
//             so only the first warp holds robot rows (value path: the last layer runs for that warp only).
//   GEMMs   = every shared-weight product (90% of the MACs) is tcgen05.mma kind::tf32, M=128, A operand in TMEM
//             (written by the row-owning threads with tcgen05.st), B = weights resident in shared memory as UMMA
//             SWIZZLE_128B K-major tiles, fp32 accumulation in TMEM.  fp32 accuracy comes from the 3xTF32 split
//             x = hi + lo:  D += lo*Whi + hi*Wlo + hi*Whi  (measured 2.5e-7 relative on a K=32 dot, tools/umma_probe.cu;
//             plain fp32 FMA chains give 2e-7).  A-in-TMEM matters: with A in shared memory an N=32 MMA is bound by
//             the 4 KB operand fetch (85 cycles measured instead of 16).
//             The robot / human embedding MLPs differ, so layer 1 concatenates them along k (robot features, human
//             features and two indicator columns that carry the biases; a row has zeros in the other agent type's
//             slots) and layer 2 stacks them along n (N=64: columns 0-31 human weights, 32-63 robot weights; each row
//             keeps the half that belongs to it).
//             Y = H w_a and H W_0 share their A operand (one N=64 chain over the stacked tile [w_a^T ; Ws[0]^T]) and the
//             GCN layer is evaluated as relu(A (H W)) -- the reference's (A H) W reassociated.  The MMAs are issued by the
//             elect.sync lane of the group's first warp (a plain `lane == 0` guard makes the compiler wrap every
//             UTCHMMA in a divergence loop: 45 instead of 16-32 cycles per MMA).
//   per-state work (similarity row, softmax, A.(HW)) stays on the FMA pipe (packed FFMA2), one node row per thread,
//             neighbours' rows read from the group's padded shared-memory buffer; attention weights never leave registers.
#include <stdlib.h>
#include <string.h>
#include "kernels.h"
#include "tc_common.cuh"
#include "tma_maps.cuh"

namespace rgl {

// [128][32] fp32 row buffer of a group, rows padded to 36 floats (144 B = 4 banks past a multiple of 32): row-per-thread
// LDS/STS.128 are conflict-free (8 consecutive rows cover all 32 banks) and every chunk address is row base + immediate
// (the buffer is no MMA operand -- the A operand lives in TMEM -- so it does not need the UMMA swizzle).
constexpr int XF_ROW = 144;                 // bytes
constexpr int XF_GROUP = 128 * XF_ROW;      // 18 KB per group
__device__ __forceinline__ uint32_t row_ptr(uint32_t xf_s, int row) { return xf_s + row * XF_ROW; }
__device__ __forceinline__ void xf_store_row(uint32_t rp, const float (&v)[32]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) sts128s(rp + c * 16, make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
}

// ---- optional phase trace (tools/trace_tc.cu builds this file with -DRGL_TC_TRACE): cycles between the marks below, summed
// over every tile of group 0 of every CTA by a non-issuing thread (lane 0 of the group's second warp) ----
#ifdef RGL_TC_TRACE
__device__ unsigned long long g_tc_trace[32];
#define TC_MARK(k) do { if (grp == 0 && gt == 32) { const long long t_ = clock64(); atomicAdd(&g_tc_trace[k], (unsigned long long)(t_ - tprev)); tprev = t_; } } while (0)
void tc_trace_read(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(g_tc_trace)); }
void tc_trace_reset() { unsigned long long z[32] = {}; cudaMemcpyToSymbol(g_tc_trace, z, sizeof(z)); }
#else
#define TC_MARK(k) do { } while (0)
#endif

// state -> row of the humans tensor (humans_bcast consecutive states share one human set); the common hb == 1 skips a 64-bit division
__device__ __forceinline__ long hgroup(long gs, int hb) { return hb == 1 ? gs : (long)((unsigned long long)gs / (unsigned)hb); }

// --------------------------------------------------------------------------------------------------- kernel
constexpr int TC_COLS = 128;          // TMEM columns per group: [0,64) accumulators, [64,96) A hi, [96,128) A lo
constexpr int C_D = 0, C_AHI = 64, C_ALO = 96;

template <int N, int G>
__global__ void __launch_bounds__(128 * G, G <= 2 ? 2 : 1) graph_forward_tc_kernel(const GraphArgs a, const __grid_constant__ CUtensorMap mapHr,
                                                                                     const __grid_constant__ CUtensorMap mapHh, const int tma_out) {
    constexpr int NMAX = N > 0 ? N : RGL_MAX_HUMANS + 1;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* smem = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));

    const int n = N > 0 ? N : a.Nh + 1;
    const int Nh = n - 1;
    const int SPT = 128 / n;                        // whole states per 128-row tile
    const int twf = tc_graph_floats(a.L);
    float* tw = smem;                               // graph operand tiles (1024 B aligned)
    float* tm = tw + twf;                           // motion operand tiles (only when S is requested)
    float* xf_all = tm + (a.mw ? TMOTION_FLOATS : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(xf_all + G * (XF_GROUP / 4));      // [0],[1] weights; [2+g] group g
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2 + G);

    const int tid = threadIdx.x, gt = tid & 127;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler
    const int grp = warp >> 2, wq = warp & 3;
    const bool skip = a.flags & RGL_FLAG_SKIP, layerwise = a.flags & RGL_FLAG_LAYERWISE;

    // row identity inside a tile (the same for every tile)
    const bool is_robot = gt < SPT;
    const int hrow = gt - SPT;
    const int s_loc = is_robot ? gt : hrow / Nh;                      // state inside the tile
    const int hum = is_robot ? 0 : hrow - s_loc * Nh;                 // human index
    const bool row_used = gt < SPT * n;
    const int node = is_robot ? 0 : hum + 1;
    // rows of this thread's state in the group's row buffer.  The tile's tail rows (gt >= SPT * n) run the per-state loops and drop
    // the result; they read state 0's rows so that every read stays inside the buffer
    const int s_rd = row_used ? s_loc : 0;
    const int hbase = SPT + s_rd * Nh;                                // first human row of this thread's state

    const int ntiles = a.ntiles;
    const int tstride = gridDim.x * G;
    float xr[RD];
    auto load_raw = [&](int tile) {
        const long gs = (long)tile * SPT + s_loc;
#pragma unroll
        for (int k = 0; k < RD; ++k) xr[k] = 0.f;
        if (tile < ntiles && row_used && gs < a.B) {
            if (is_robot) {
                const float* p = a.robot + gs * RD;
#pragma unroll
                for (int k = 0; k < RD; ++k) xr[k] = __ldg(p + k);
            } else {
                const float* p = a.humans + (hgroup(gs, a.hb) * Nh + hum) * HD;
#pragma unroll
                for (int k = 0; k < HD; ++k) xr[k] = __ldg(p + k);
            }
        }
    };
    // next tile's raw row: prefetched into L2 while this tile computes (a register prefetch would be spilled: the row
    // would have to stay live across the whole tile)
    auto prefetch_raw = [&](int tile) {
        const long gs = (long)tile * SPT + s_loc;
        if (tile < ntiles && row_used && gs < a.B) {
            const float* p = is_robot ? a.robot + gs * RD : a.humans + (hgroup(gs, a.hb) * Nh + hum) * HD;
            asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(p + (is_robot ? RD - 1 : HD - 1)));
        }
    };
    int tile = blockIdx.x * G + grp;

    if (warp == 0) tmem_alloc(tslot, TC_COLS * G);
    if (tid == 0) {
        for (int i = 0; i < 2 + G; ++i) mbar_init(bars + i, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // PDL: everything above (TMEM allocation, barrier set-up) may overlap the tail of the previous kernel on the stream;
    // the states, the packed weights and every output buffer are only touched after this point
    pdl_wait();
    if (a.flags & RGL_INTERNAL_PDL_EARLY) pdl_trigger();      // small grids: let the next kernel's CTAs launch right away (kernels.h)
    load_raw(tile);                          // the first tile's raw rows: in flight under the weight TMA
    if (tid == 0) {
        // stage 0: layer-1 embedding tiles (the first MMA needs only these); stage 1: everything else
        const float* src = a.gw + graph_tc_off(a.L);
        mbar_arrive_expect_tx(bars + 0, T_W1 * 4u);
        bulk_g2s(tw, src, T_W1 * 4u, bars + 0);
        const uint32_t rest = (uint32_t)(twf - T_W1) * 4u;
        mbar_arrive_expect_tx(bars + 1, rest + (a.mw ? TMOTION_FLOATS * 4u : 0u));
        bulk_g2s(tw + T_W1, src + T_W1, rest, bars + 1);
        if (a.mw) bulk_g2s(tm, a.mw + MOTION_TC_OFF, TMOTION_FLOATS * 4u, bars + 1);
    }

    const uint32_t tbase = __shfl_sync(0xffffffffu, *tslot, 0);
    const uint32_t tg = tbase + grp * TC_COLS;                        // this group's TMEM columns (lane field 0: MMA view)
    const uint32_t tl = tg + ((uint32_t)(wq * 32) << 16);             // this warp's lane quadrant (ld / st view)
    const uint32_t xf_s = __shfl_sync(0xffffffffu, smem_u32(xf_all), 0) + grp * XF_GROUP;
    uint64_t* gbar = bars + 2 + grp;
    uint32_t par = 0;
    const uint32_t tw_s = __shfl_sync(0xffffffffu, smem_u32(tw), 0), tm_s = __shfl_sync(0xffffffffu, smem_u32(tm), 0);   // warp-uniform for the compiler
    const bool issuer = wq == 0;                                      // warp-uniform; lane 0 of that warp issues the MMAs
#ifdef RGL_TC_TRACE
    long long tprev = clock64();
#endif
    // (trace slots 20-23: time spent BEFORE the waits = the thread work; slots 0-14: the waits themselves)
    auto group_sync = [&]() { TC_MARK(22); asm volatile("bar.sync %0, 128;" :: "r"(grp + 1) : "memory"); };
    auto mma_wait = [&]() { TC_MARK(21); mbar_wait_sleepy(gbar, par); par ^= 1; tc_fence_after(); };
    // A operand written -> visible to the MMAs issued after the barrier
    auto publish = [&]() { TC_MARK(20); tmem_st_wait(); TC_MARK(23); tc_fence_before(); group_sync(); };

    const uint32_t my_row = row_ptr(xf_s, gt);
    const bool robot_warps = wq * 32 < SPT;                           // warp-uniform: this warp holds at least one robot row
    bool first = true;

    for (; tile < ntiles; tile += tstride) {
        const long s0 = (long)tile * SPT;
        const int cnt = (int)min((long)SPT, (long)a.B - s0);
        const bool valid = row_used && s_loc < cnt;
        if (tma_out && gt == 0) tma_store_wait_read();      // the previous tile's tensor stores have read the staging rows
#ifdef RGL_TC_TRACE
        tprev = clock64();
        if (grp == 0 && gt == 32) atomicAdd(&g_tc_trace[31], 1ull);
#endif

        // ================= embedding layer 1: hidden = relu([x_r | x_h | 1_r | 1_h] . W0cat^T), K = 16, N = 64 =================
        {
            float a0[16];
#pragma unroll
            for (int k = 0; k < RD; ++k) a0[k] = is_robot ? xr[k] : 0.f;
#pragma unroll
            for (int k = 0; k < HD; ++k) a0[RD + k] = is_robot ? 0.f : xr[k];
            a0[14] = (valid && is_robot) ? 1.f : 0.f;
            a0[15] = (valid && !is_robot) ? 1.f : 0.f;
            st_split<16>(tl + C_AHI, tl + C_ALO, a0);
        }
        publish();
        TC_MARK(0);   // after: publish()
        if (issuer) {
            if (elect_one()) {
                if (first) mbar_wait(bars + 0, 0);
                tc_fence_after();
                issue_gemm<2>(tg + C_D, tg + C_AHI, tg + C_ALO, tw_s + T_W0 * 4, tw_s + T_W0 * 4 + 64, umma_idesc(128, 64), 0);
                umma_commit(gbar);
            }
            __syncwarp();
        }
        prefetch_raw(tile + tstride);
        mma_wait();
        TC_MARK(1);   // after: mma_wait()

        // ================= embedding layer 2: [X_h | X_r] = hidden . [W1_h ; W1_r]^T, K = 64 in two halves, N = 64 =================
        float x[32];
        {
            uint32_t h0[32], h1[32];
            tmem_ld64(tl + C_D, h0, h1);     // the whole hidden row leaves TMEM before the accumulator columns are reused
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(__uint_as_float(h0[j]), 0.f);
            st_split<32>(tl + C_AHI, tl + C_ALO, v);
            publish();
            TC_MARK(2);   // after: publish()
            if (issuer) {
                if (elect_one()) {
                    if (first) mbar_wait(bars + 1, 0);
                    tc_fence_after();
                    issue_gemm<4>(tg + C_D, tg + C_AHI, tg + C_ALO, tw_s + T_W1 * 4, tw_s + (T_W1 + 4096) * 4, umma_idesc(128, 64), 0);
                    umma_commit(gbar);
                }
                __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(__uint_as_float(h1[j]), 0.f);
            mma_wait();                      // first half consumed: its A columns may be overwritten
            TC_MARK(3);   // after: mma_wait()
            st_split<32>(tl + C_AHI, tl + C_ALO, v);
            publish();
            TC_MARK(4);   // after: publish()
            if (issuer) {
                if (elect_one()) {
                    tc_fence_after();
                    issue_gemm<4>(tg + C_D, tg + C_AHI, tg + C_ALO, tw_s + (T_W1 + 2048) * 4, tw_s + (T_W1 + 4096 + 2048) * 4, umma_idesc(128, 64), 1);
                    umma_commit(gbar);
                }
                __syncwarp();
            }
            if (first) { mbar_wait(bars + 1, 0); first = false; }      // the biases arrive with stage 1
            mma_wait();
            TC_MARK(5);   // after: mma_wait()
            tmem_ld64(tl + C_D, h0, h1);     // h0 = human-weight version, h1 = robot-weight version
            const float* bias = tw + tc_bias_off(a.L) + (is_robot ? TC_RBIAS : 0);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 b = lds128(bias + 4 * c);
                x[4 * c + 0] = fmaxf(__uint_as_float(is_robot ? h1[4 * c + 0] : h0[4 * c + 0]) + b.x, 0.f);
                x[4 * c + 1] = fmaxf(__uint_as_float(is_robot ? h1[4 * c + 1] : h0[4 * c + 1]) + b.y, 0.f);
                x[4 * c + 2] = fmaxf(__uint_as_float(is_robot ? h1[4 * c + 2] : h0[4 * c + 2]) + b.z, 0.f);
                x[4 * c + 3] = fmaxf(__uint_as_float(is_robot ? h1[4 * c + 3] : h0[4 * c + 3]) + b.w, 0.f);
            }
        }
        xf_store_row(my_row, x);             // feature rows for the neighbours' similarity logits

        // ================= GCN layers: H' = relu(A (H W_l)) (+ H) =================
        // (the reference evaluates (A H) W_l; the products are reassociated so that H W_l shares its A operand with Y = H w_a)
        float p[NMAX];
        for (int l = 0; l < a.L; ++l) {
            const bool last = (l == a.L - 1);
            const bool robot_only = last && a.H == nullptr && a.S == nullptr;
            const bool active = !robot_only || robot_warps;          // warp-uniform
            const bool sim = (l == 0) || layerwise;

            st_split<32>(tl + C_AHI, tl + C_ALO, x);
            publish();       // also: feature rows in xf visible; every read of the previous layer's H W rows is done
            TC_MARK(6);   // after: publish()
            if (issuer) {
                if (elect_one()) {
                    tc_fence_after();
                    if (l == 0) {
                        // one N = 64 MMA chain: columns [0,32) = Y = X w_a, columns [32,64) = X Ws[0]
                        issue_gemm<4>(tg + C_D, tg + C_AHI, tg + C_ALO, tw_s + T_WA * 4, tw_s + (T_WA + 2048) * 4, umma_idesc(128, 64), 0);
                    } else {
                        if (layerwise)
                            issue_gemm<4>(tg + C_D, tg + C_AHI, tg + C_ALO, tw_s + T_WA * 4, tw_s + (T_WA + 2048) * 4, umma_idesc(128, 32), 0);
                        const uint32_t ws = tw_s + (T_WS1 + (l - 1) * 2048) * 4;
                        issue_gemm<4>(tg + C_D + 32, tg + C_AHI, tg + C_ALO, ws, ws + 4096, umma_idesc(128, 32), 0);
                    }
                    umma_commit(gbar);
                }
                __syncwarp();
            }
            mma_wait();
            TC_MARK(7);   // after: mma_wait()

            if (sim) {
                // ---- this row's similarity logits against the state's feature rows + softmax (FMA pipe) ----
                if (active) {                                // warp-uniform: tcgen05.ld is warp-collective
                    uint32_t yr[32];
                    tmem_ld32(tl + C_D, yr);
                    if (row_used) {
                    float mx = -INFINITY;
#pragma unroll
                    for (int j = 0; j < NMAX; ++j) {
                        if (N > 0 || j < n) {
                            const uint32_t rp = row_ptr(xf_s, j == 0 ? s_rd : hbase + j - 1);
                            f32x2 d01 = 0ull, d23 = 0ull;           // four partial sums, two per packed FFMA2
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                f32x2 xlo, xhi;
                                lds128s2(rp + c * 16, xlo, xhi);
                                fma2(d01, pack2u(yr[4 * c + 0], yr[4 * c + 1]), xlo);
                                fma2(d23, pack2u(yr[4 * c + 2], yr[4 * c + 3]), xhi);
                            }
                            float d0, d1, d2, d3;
                            unpack2(d01, d0, d1);
                            unpack2(d23, d2, d3);
                            p[j] = (d0 + d1) + (d2 + d3);
                            mx = fmaxf(mx, p[j]);
                        }
                    }
                    float sum = 0.f;
#pragma unroll
                    for (int j = 0; j < NMAX; ++j)
                        if (N > 0 || j < n) { p[j] = expf(p[j] - mx); sum += p[j]; }
                    const float rs = 1.f / sum;             // one division; p * (1/sum) differs from p / sum by <= 1 ulp
#pragma unroll
                    for (int j = 0; j < NMAX; ++j)
                        if (N > 0 || j < n) p[j] *= rs;
                    if (a.A0 != nullptr && l == 0 && tile == 0 && s_loc == 0) {
#pragma unroll
                        for (int j = 0; j < NMAX; ++j)
                            if (N > 0 || j < n) a.A0[node * n + j] = p[j];
                    }
                    }
                }
                group_sync();                                // every read of the feature rows is done
                TC_MARK(8);   // after: group_sync()
            }

            // ---- H W rows -> xf; H' = relu(sum_j A[i][j] (H W)[j]) (+ H) ----
            {
                uint32_t hw[32];
                tmem_ld32(tl + C_D + 32, hw);
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    sts128s(my_row + c * 16, make_float4(__uint_as_float(hw[4 * c]), __uint_as_float(hw[4 * c + 1]),
                                                           __uint_as_float(hw[4 * c + 2]), __uint_as_float(hw[4 * c + 3])));
            }
            tc_fence_before();
            group_sync();
            TC_MARK(9);   // after: group_sync()
            if (active && row_used) {
                f32x2 acc2[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) acc2[c] = 0ull;
#pragma unroll
                for (int j = 0; j < NMAX; ++j) {
                    if (N > 0 || j < n) {
                        const uint32_t rp = row_ptr(xf_s, j == 0 ? s_rd : hbase + j - 1);
                        const f32x2 pj = pack2(p[j], p[j]);
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            f32x2 hlo, hhi;
                            lds128s2(rp + c * 16, hlo, hhi);
                            fma2(acc2[2 * c], pj, hlo);
                            fma2(acc2[2 * c + 1], pj, hhi);
                        }
                    }
                }
                float acc[32];
#pragma unroll
                for (int c = 0; c < 16; ++c) unpack2(acc2[c], acc[2 * c], acc[2 * c + 1]);
                if (skip) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) x[c] += fmaxf(acc[c], 0.f);
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) x[c] = fmaxf(acc[c], 0.f);
                }
            }
            if (!last && layerwise) {                        // the next layer's similarity needs the new feature rows
                group_sync();
                TC_MARK(10);   // after: group_sync()
                xf_store_row(my_row, x);
            }
        }

        // ================= outputs =================
        // PDL: this CTA's last tile has left the tensor pipe -- let the next kernel on the stream launch and run its
        // prologue under the output phase (it waits for this grid to complete before it reads or writes memory)
        if (tile + tstride >= ntiles) pdl_trigger();
        if (a.E != nullptr && is_robot && valid) {
            float* e = a.E + (s0 + s_loc) * XD;
#pragma unroll
            for (int c = 0; c < 8; ++c) *reinterpret_cast<float4*>(e + 4 * c) = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
        }
        if (a.S != nullptr) {
            // state-predictor head on every row (robot rows are computed and dropped): relu(H W0^T + b0) on the tensor
            // cores (N = 64), the 64 -> 5 output layer per thread on the FMA pipe
            st_split<32>(tl + C_AHI, tl + C_ALO, x);
            publish();
            TC_MARK(11);   // after: publish()
            if (issuer) {
                if (elect_one()) {
                    tc_fence_after();
                    issue_gemm<4>(tg + C_D, tg + C_AHI, tg + C_ALO, tm_s + TM_W0 * 4, tm_s + (TM_W0 + 2048) * 4, umma_idesc(128, 64), 0);
                    umma_commit(gbar);
                }
                __syncwarp();
            }
            mma_wait();
            TC_MARK(12);   // after: mma_wait()
            uint32_t h0[32], h1[32];
            tmem_ld64(tl + C_D, h0, h1);
            f32x2 part2[HD];                         // two partial sums per output (packed FFMA2)
#pragma unroll
            for (int c = 0; c < HD; ++c) part2[c] = 0ull;
#pragma unroll
            for (int k4 = 0; k4 < 16; ++k4) {
                const float4 b = lds128(tm + TM_B0 + 4 * k4);
                const uint32_t* hh = k4 < 8 ? h0 : h1;
                const int o = (k4 & 7) * 4;
                const f32x2 v01 = pack2(fmaxf(__uint_as_float(hh[o + 0]) + b.x, 0.f), fmaxf(__uint_as_float(hh[o + 1]) + b.y, 0.f));
                const f32x2 v23 = pack2(fmaxf(__uint_as_float(hh[o + 2]) + b.z, 0.f), fmaxf(__uint_as_float(hh[o + 3]) + b.w, 0.f));
#pragma unroll
                for (int c = 0; c < HD; ++c) {
                    f32x2 wlo, whi;
                    lds128s2(tm_s + (TM_W1 + c * MH + 4 * k4) * 4, wlo, whi);
                    fma2(part2[c], v01, wlo);
                    fma2(part2[c], v23, whi);
                }
            }
            float part[HD];
#pragma unroll
            for (int c = 0; c < HD; ++c) {
                float lo_, hi_;
                unpack2(part2[c], lo_, hi_);
                part[c] = (lo_ + hi_) + tm[TM_B1 + c];
            }
            if (!is_robot && valid) {
                float* so = a.S + ((s0 + s_loc) * Nh + hum) * HD;
#pragma unroll
                for (int c = 0; c < HD; ++c) so[c] = part[c];
            }
        }
        if (a.H != nullptr) {
            group_sync();                                    // every read of the last H W rows is done
            if (tma_out) {
                // dense SWIZZLE_128B staging (robot rows at [0,SPT), human rows from the next multiple of 8 rows, so both boxes
                // start 1 KB aligned); two tensor stores (node 0 of SPT states; nodes 1..Nh of SPT states) write the tile in
                // HBM order, clipped to the batch by the tensor map
                const int hb8 = (SPT + 7) & ~7;
                const int srow = is_robot ? gt : hb8 + hrow;
                const uint32_t rp = xf_s + srow * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    sts128s(rp + (((c ^ srow) & 7) << 4), make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]));
                fence_proxy_async();
                group_sync();
                if (gt == 0) {
                    tma_store_3d(&mapHr, 0, 0, (int)s0, xf_s);
                    tma_store_3d(&mapHh, 0, 1, (int)s0, xf_s + hb8 * 128);
                    tma_commit();
                }
            } else {
                // stage the final rows in xf, then copy out in HBM order: 512 contiguous bytes per warp instruction
                xf_store_row(my_row, x);
                group_sync();
                float* dst = a.H + s0 * n * XD;
                const int chunks = cnt * n * 8;
                for (int idx = gt; idx < chunks; idx += 128) {
                    const int orow = idx >> 3, c = idx & 7;
                    const int s = orow / n, i = orow - s * n;
                    const int srow = i == 0 ? s : SPT + s * Nh + i - 1;
                    *reinterpret_cast<float4*>(dst + (size_t)idx * 4) = lds128s(row_ptr(xf_s, srow) + c * 16);
                }
            }
            // the next tile writes xf only after further group barriers (and after thread 0 has seen the stores read it)
        }
        load_raw(tile + tstride);            // next tile's raw rows (prefetched into L2 above): consumed at the top of the loop
    }

    // teardown: the bulk copies must have landed before the CTA's shared memory is released
    if (tma_out && gt == 0) tma_store_wait_all();
    if (tid == 0) { mbar_wait(bars + 0, 0); mbar_wait(bars + 1, 0); }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TC_COLS * G);
}

// ---------------------------------------------------------------------------------------------------
static size_t tc_smem_bytes(int L, bool motion, int G) {
    return 1024 + ((size_t)tc_graph_floats(L) + (motion ? TMOTION_FLOATS : 0) + (size_t)G * (XF_GROUP / 4)) * 4 + (2 + G) * 8 + 16;
}

template <int N, int G>
static cudaError_t launch_tc(const GraphArgs& a, int num_sms, size_t max_smem, cudaStream_t st) {
    const int n = a.Nh + 1;
    const size_t smem = tc_smem_bytes(a.L, a.mw != nullptr, G);
    if (smem > max_smem) return cudaErrorInvalidConfiguration;
    GraphArgs b = a;
    if (pdl_early(a.B)) b.flags |= RGL_INTERNAL_PDL_EARLY;
    const int spt = 128 / n;
    b.ntiles = (a.B + spt - 1) / spt;
    if (cudaError_t e = ensure_dyn_smem(graph_forward_tc_kernel<N, G>, (int)max_smem)) return e;
    int per_sm = (int)((228 * 1024) / (smem + 1024));
    const int max_cta = G <= 2 ? 2 : 1;                       // matches __launch_bounds__ and the 512 TMEM columns of an SM
    if (per_sm > max_cta) per_sm = max_cta;
    if (per_sm < 1) per_sm = 1;
    const int want = (b.ntiles + G - 1) / G;
    const int grid = want < num_sms * per_sm ? want : num_sms * per_sm;
    // H leaves through TMA tensor stores when the driver offers cuTensorMapEncodeTiled (RGL_TC_TMA_OUT=0 disables: experiments)
    static const char* tma_env = getenv("RGL_TC_TMA_OUT");
    CUtensorMap mr, mh;
    memset(&mr, 0, sizeof(mr));
    memset(&mh, 0, sizeof(mh));
    int tma_out = 0;
    if (a.H != nullptr && !(tma_env && tma_env[0] == '0'))
        tma_out = make_state_map(&mr, a.H, a.B, n, 1, spt) && make_state_map(&mh, a.H, a.B, n, n - 1, spt) ? 1 : 0;
    return launch_pdl(graph_forward_tc_kernel<N, G>, dim3(grid), dim3(128 * G), smem, st, b, mr, mh, tma_out);
}

template <int N>
static cudaError_t dispatch_tc(const GraphArgs& a, int num_sms, size_t max_smem, cudaStream_t st) {
    // RGL_TC_GROUPS (experiments only): force the number of 128-row groups per CTA
    static const char* force = getenv("RGL_TC_GROUPS");
    const int n = a.Nh + 1, spt = 128 / n;
    const int ntiles = (a.B + spt - 1) / spt;
    int g = force ? atoi(force) : 0;
    if (g != 1 && g != 2 && g != 4) {
        // the state-predictor tiles do not leave room for two CTAs per SM: one CTA of four groups shares one weight copy.
        // Small batches (at most two tiles per SM): one group per CTA so that the tiles spread over all SMs.
        // RGL_FLAG_THROUGHPUT (the caller keeps several launches in flight): two groups per CTA even for small batches, so
        // that concurrent launches share SMs at the steady-state occupancy (measured at B = 4096: 850 vs 677 M states/s).
        // (with the motion tiles a one-group CTA still fits twice per SM: small state-predictor batches spread over all SMs)
        if (a.mw != nullptr) g = ntiles <= 2 * num_sms ? 1 : 4;
        else g = (ntiles <= 2 * num_sms && !(a.flags & RGL_FLAG_THROUGHPUT)) ? 1 : 2;
    }
    if (g == 1) return launch_tc<N, 1>(a, num_sms, max_smem, st);
    if (g == 2) return launch_tc<N, 2>(a, num_sms, max_smem, st);
    return launch_tc<N, 4>(a, num_sms, max_smem, st);
}

cudaError_t run_graph_forward_tc(const GraphArgs& a, int num_sms, size_t max_smem, cudaStream_t st) {
    switch (a.Nh + 1) {
        case 6: return dispatch_tc<6>(a, num_sms, max_smem, st);
        case 11: return dispatch_tc<11>(a, num_sms, max_smem, st);
        case 21: return dispatch_tc<21>(a, num_sms, max_smem, st);
        default: return dispatch_tc<0>(a, num_sms, max_smem, st);
    }
}

}  // namespace rgl
