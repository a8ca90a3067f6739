// Fused RGL graph forward on tcgen05 / TMEM, ROW-PAIRED per-state phases (the throughput variant of graph_forward_tc.cu).
//
// Same contract and the same GEMM mapping as graph_forward_tc.cu (crowd_nav/policy/graph_model.py:99-130; 128-row tiles, one
// node row per thread / TMEM lane, every shared-weight product a 3xTF32 tcgen05.mma with the A operand in TMEM).  What
// changes is the part that bounded that kernel: the per-state phases (similarity logits, A.(HW)) read every neighbour row
// of the state once per node row -- n * 128 B per thread and phase, 82 % of all shared-memory wavefronts, LSU data pipe 72 %
// busy at steady state (profiles/r1_tc_graph_forward_b1m_full.md).  Here two adjacent lanes own two node rows of the SAME
// state and split the 32 feature columns between them:
//
//   lane 2m   (row r)     columns  0-15 of rows r and r+1
//   lane 2m+1 (row r+1)   columns 16-31 of rows r and r+1
//
// so a neighbour row's 64-byte half is loaded once and used for two rows (half the LDS.128 per FFMA), and the halves are
// exchanged with warp shuffles: 16 for the partner's Y half and 6 + 6 for the partial logits / attention weights in the
// similarity, 16 per GCN layer for the partner row's half of A.(HW).  Per row and tile: 72 LDS.128 + 60 SHFL instead of
// 144 LDS.128.  Row-local data (TMEM loads / stores, operand splits, epilogues) stays in natural column order; lane parity
// only selects which half is kept and which is sent (SEL).
//
// Rows are STATE-MAJOR inside a tile (row = state * NP + node, the HBM order) with NP = n rounded up to even, so that the
// lanes of a pair always belong to one state; for odd n the last row of a state is a padding row (zero input, never read,
// never stored).  H leaves through ONE 3-D tensor-map TMA store per tile.
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

// training forward (a.save): this thread's node row of an activation, 32 floats -> dst (row-major [B*n][32])
__device__ __forceinline__ void save_row32(float* dst, const float (&v)[32]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

// ---- optional phase trace (tools/trace_tc.cu builds this file with -DRGL_TC_TRACE): cycles between the marks below, summed
// over every tile of group 0 of every CTA by a non-issuing thread (lane 0 of the group's second warp) ----
#ifdef RGL_TC_TRACE
__device__ unsigned long long g_tp_trace[32];
#define TC_MARK(k) do { if (grp == 0 && gt == 32) { const long long t_ = clock64(); atomicAdd(&g_tp_trace[k], (unsigned long long)(t_ - tprev)); tprev = t_; } } while (0)
void tp_trace_read(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_tp_trace, sizeof(g_tp_trace)); }
void tp_trace_reset() { unsigned long long z[32] = {}; cudaMemcpyToSymbol(g_tp_trace, z, sizeof(z)); }
#else
#define TC_MARK(k) do { } while (0)
#endif

// state -> row of the humans tensor (humans_bcast consecutive states share one human set); the common hb == 1 skips a 64-bit division
__device__ __forceinline__ long hgroup(long gs, int hb) { return hb == 1 ? gs : (long)((unsigned long long)gs / (unsigned)hb); }

// --------------------------------------------------------------------------------------------------- kernel
constexpr int TC_COLS = 128;          // TMEM columns per group: [0,64) accumulators, [64,96) A hi, [96,128) A lo
constexpr int C_D = 0, C_AHI = 64, C_ALO = 96;

template <int N, int G, bool SAVE>
__global__ void __launch_bounds__(128 * G, G <= 2 ? 2 : 1) graph_forward_tp_kernel(const GraphArgs a, const __grid_constant__ CUtensorMap mapH,
                                                                                     const int tma_out) {
    static_assert(N >= 2, "the paired kernel is instantiated for compile-time node counts");
    constexpr int NMAX = N;
    constexpr int NP = (N + 1) & ~1;                // rows per state inside a tile (even: the lanes of a pair share a state)
    constexpr int SPT = 128 / NP;                   // whole states per 128-row tile
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* smem = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));

    constexpr int n = N;
    constexpr int Nh = N - 1;
    const int twf = tc_graph_floats(a.L);
    float* tw = smem;                               // graph operand tiles (1024 B aligned)
    float* tm = tw + twf;                           // motion operand tiles (only when S is requested)
    float* xf_all = tm + (a.mw ? TMOTION_FLOATS : 0);
    float* stage_all = xf_all + G * (XF_GROUP / 4);                                 // SAVE only: 4 KB per warp (coalesced activation saves)
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_all + (SAVE ? G * 4096 : 0));  // [0],[1] weights; [2+g] group g
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2 + G);

    const int tid = threadIdx.x, gt = tid & 127;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler
    const int grp = warp >> 2, wq = warp & 3;
    const bool skip = a.flags & RGL_FLAG_SKIP, layerwise = a.flags & RGL_FLAG_LAYERWISE;

    // row identity inside a tile (the same for every tile): state-major, NP rows per state
    const int s_loc = gt / NP;                                        // state inside the tile
    const int node = gt - s_loc * NP;
    const bool row_used = s_loc < SPT && node < N;                    // padding rows (odd n) and the tile's tail rows idle
    const bool is_robot = node == 0;
    const int hum = node - 1;                                         // human index (node >= 1)
    // first row of this thread's state.  The tile's tail rows (s_loc == SPT: 128 is no multiple of NP) run the per-state loops
    // like everyone else and drop the result; they read state 0's rows so that every read stays inside the group's row buffer
    // (compute-sanitizer racecheck: with sbase = s_loc * NP they read up to N - 1 rows past it, into the save staging area)
    const int sbase = (s_loc < SPT ? s_loc : 0) * NP;
    const bool odd = gt & 1;                                          // lane parity: which 16-column half this thread computes

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
    // ---- training saves (SAVE): a warp's valid rows are CONSECUTIVE rows of the [B*n, .] activation tensors (state-major tiles),
    // so each warp transposes its rows through 4 KB of shared memory and writes them with fully coalesced 512-byte stores
    // (a thread writing its own 128-byte row costs 32 partial-sector requests per instruction: 2.4 TB/s instead of HBM speed)
    const uint32_t stage_s = SAVE ? __shfl_sync(0xffffffffu, smem_u32(stage_all), 0) + (uint32_t)(grp * 4 + wq) * 4096u : 0u;
    const int lane = tid & 31;
    auto warp_save = [&](float* base, int pitch, long grow_, bool ok, const float (&v)[32]) {
        // base: tensor base (+ column offset); pitch: floats per tensor row; grow_: this lane's tensor row; ok: the lane holds a real row
        const unsigned okm = __ballot_sync(0xffffffffu, ok);
        if (okm == 0u) return;
        const int di = __popc(okm & ((1u << lane) - 1u));                 // dense index of this lane's row inside the warp's block
        const int nrow = __popc(okm);
        const long first = __shfl_sync(0xffffffffu, grow_, __ffs(okm) - 1);  // tensor row of the block's first row
        __syncwarp();                                                      // the previous save has been copied out
        if (ok) {
            const uint32_t rp = stage_s + di * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                sts128s(rp + (((c ^ di) & 7) << 4), make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int q = k * 32 + lane, r = q >> 3, c = q & 7;
            if (r < nrow)
                *reinterpret_cast<float4*>(base + (first + r) * pitch + 4 * c) = lds128s(stage_s + r * 128 + (((c ^ r) & 7) << 4));
        }
    };
    bool first = true;

    for (; tile < ntiles; tile += tstride) {
        const long s0 = (long)tile * SPT;
        const int cnt = (int)min((long)SPT, (long)a.B - s0);
        const bool valid = row_used && s_loc < cnt;
        const long grow = (s0 + s_loc) * n + node;           // this thread's row of the [B, n, .] activation tensors (training saves)
        const bool saving = SAVE && valid;                   // training forward: activation saves (compile-time: the inference kernels carry none of it)
        if (tma_out && gt == 0) tma_store_wait_read();      // the previous tile's tensor stores have read the staging rows
#ifdef RGL_TC_TRACE
        tprev = clock64();
        if (grp == 0 && gt == 32) atomicAdd(&g_tp_trace[31], 1ull);
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
            if (SAVE) warp_save(a.sv.a1r, 64, grow, saving, v);            // relu(hidden) [B, n, 64], columns 0-31
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
            if (SAVE) warp_save(a.sv.a1r + 32, 64, grow, saving, v);       // columns 32-63
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
        if (SAVE) warp_save(a.sv.X, 32, grow, saving, x);

        // ================= GCN layers: H' = relu(A (H W_l)) (+ H) =================
        // (the reference evaluates (A H) W_l; the products are reassociated so that H W_l shares its A operand with Y = H w_a)
        // Per-state phases are ROW-PAIRED: this thread computes the columns [16*odd, 16*odd + 16) of its own row AND of its
        // partner's row (lane ^ 1, same state), the partner the other 16 columns of both; halves are swapped by shuffles.
        constexpr bool KEEP_PP = N <= 8;                 // small n: keep the partner's attention weights in registers; large n: re-fetch
        float p[NMAX], pp[KEEP_PP ? NMAX : 1];           // attention weights of this row / of the partner's row
        const uint32_t half_off = odd ? 64u : 0u;        // byte offset of this thread's column half inside a feature row
        for (int l = 0; l < a.L; ++l) {
            const bool last = (l == a.L - 1);
            const bool robot_only = last && a.H == nullptr && a.S == nullptr && !SAVE;
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
                // ---- similarity logits of the two rows of the pair over this thread's 16 k-columns, then softmax of the own row ----
                {
                    uint32_t yr[32];
                    tmem_ld32(tl + C_D, yr);
                    if (SAVE && l == 0) {
                        float yv[32];
#pragma unroll
                        for (int c = 0; c < 32; ++c) yv[c] = __uint_as_float(yr[c]);
                        warp_save(a.sv.Y, 32, grow, saving, yv);
                    }
                    // own row's Y at my half / partner row's Y at my half (the partner sends the half it does not use itself)
                    f32x2 ya[8], yb[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t m0 = odd ? yr[16 + 2 * k] : yr[2 * k], m1 = odd ? yr[17 + 2 * k] : yr[2 * k + 1];
                        const uint32_t o0 = odd ? yr[2 * k] : yr[16 + 2 * k], o1 = odd ? yr[2 * k + 1] : yr[17 + 2 * k];
                        ya[k] = pack2u(m0, m1);
                        yb[k] = pack2u(__shfl_xor_sync(0xffffffffu, o0, 1), __shfl_xor_sync(0xffffffffu, o1, 1));
                    }
                    float mx = -INFINITY;
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        const uint32_t rp = row_ptr(xf_s, sbase + j) + half_off;
                        f32x2 da = 0ull, db = 0ull;              // partial logits (two lanes of a packed FFMA2 each)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            f32x2 xlo, xhi;
                            lds128s2(rp + c * 16, xlo, xhi);
                            fma2(da, ya[2 * c], xlo);
                            fma2(da, ya[2 * c + 1], xhi);
                            fma2(db, yb[2 * c], xlo);
                            fma2(db, yb[2 * c + 1], xhi);
                        }
                        float a0_, a1_, b0_, b1_;
                        unpack2(da, a0_, a1_);
                        unpack2(db, b0_, b1_);
                        // own row: my half + the partner's half (fp32 addition commutes: both lanes of a pair would get the same bits)
                        p[j] = (a0_ + a1_) + __shfl_xor_sync(0xffffffffu, b0_ + b1_, 1);
                        mx = fmaxf(mx, p[j]);
                    }
                    float sum = 0.f;
#pragma unroll
                    for (int j = 0; j < N; ++j) { p[j] = expf(p[j] - mx); sum += p[j]; }
                    const float rs = 1.f / sum;             // one division; p * (1/sum) differs from p / sum by <= 1 ulp
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        p[j] *= rs;
                        if (KEEP_PP) pp[j] = __shfl_xor_sync(0xffffffffu, p[j], 1);
                    }
                    if (a.A0 != nullptr && l == 0 && tile == 0 && s_loc == 0 && node < N) {
#pragma unroll
                        for (int j = 0; j < N; ++j) a.A0[node * n + j] = p[j];
                    }
                    if (saving && l == 0) {
#pragma unroll
                        for (int j = 0; j < N; ++j) a.sv.A[grow * n + j] = p[j];
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
                if (SAVE) {                                   // Z_l = H_{l-1} W_l (reassociated layer)
                    float zv[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) zv[c] = __uint_as_float(hw[c]);
                    warp_save(a.sv.M[l], 32, grow, saving, zv);
                }
            }
            tc_fence_before();
            group_sync();
            TC_MARK(9);   // after: group_sync()
            // value path, last layer: only node 0 is consumed -- the pair (node 0, node 1) computes that one row
            const bool pair_active = !robot_only || node < 2;
            const unsigned pmask = __ballot_sync(0xffffffffu, pair_active);          // both lanes of a pair are in or out together
            float rl[SAVE ? 32 : 1];
            if (pair_active) {
                f32x2 acc_a[8], acc_b[8];                    // my 16 columns of the own row / of the partner's row
#pragma unroll
                for (int c = 0; c < 8; ++c) { acc_a[c] = 0ull; acc_b[c] = 0ull; }
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const uint32_t rp = row_ptr(xf_s, sbase + j) + half_off;
                    const float pq = KEEP_PP ? pp[j] : __shfl_xor_sync(pmask, p[j], 1);
                    const f32x2 pj = pack2(p[j], p[j]), qj = pack2(pq, pq);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        f32x2 hlo, hhi;
                        lds128s2(rp + c * 16, hlo, hhi);
                        fma2(acc_a[2 * c], pj, hlo);
                        fma2(acc_a[2 * c + 1], pj, hhi);
                        fma2(acc_b[2 * c], qj, hlo);
                        fma2(acc_b[2 * c + 1], qj, hhi);
                    }
                }
                // swap: the partner receives its row's columns of my half, I receive my row's columns of the partner's half
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float m0, m1, s0_, s1_;
                    unpack2(acc_a[c], m0, m1);
                    unpack2(acc_b[c], s0_, s1_);
                    const float r0_ = __shfl_xor_sync(pmask, s0_, 1), r1_ = __shfl_xor_sync(pmask, s1_, 1);
                    const float lo0 = odd ? r0_ : m0, lo1 = odd ? r1_ : m1;       // columns 2c, 2c+1
                    const float hi0 = odd ? m0 : r0_, hi1 = odd ? m1 : r1_;       // columns 16+2c, 17+2c
                    if (SAVE) {              // relu(A Z_l) before the skip add (the relu mask of the backward)
                        rl[2 * c] = fmaxf(lo0, 0.f); rl[2 * c + 1] = fmaxf(lo1, 0.f);
                        rl[16 + 2 * c] = fmaxf(hi0, 0.f); rl[17 + 2 * c] = fmaxf(hi1, 0.f);
                    }
                    if (skip) {
                        x[2 * c] += fmaxf(lo0, 0.f); x[2 * c + 1] += fmaxf(lo1, 0.f);
                        x[16 + 2 * c] += fmaxf(hi0, 0.f); x[17 + 2 * c] += fmaxf(hi1, 0.f);
                    } else {
                        x[2 * c] = fmaxf(lo0, 0.f); x[2 * c + 1] = fmaxf(lo1, 0.f);
                        x[16 + 2 * c] = fmaxf(hi0, 0.f); x[17 + 2 * c] = fmaxf(hi1, 0.f);
                    }
                }
            }
            if (SAVE) {
                warp_save(a.sv.Rl[l], 32, grow, saving, reinterpret_cast<const float(&)[32]>(rl));
                warp_save(a.sv.Hl[l], 32, grow, saving, x);
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
                const float m0_ = fmaxf(__uint_as_float(hh[o + 0]) + b.x, 0.f), m1_ = fmaxf(__uint_as_float(hh[o + 1]) + b.y, 0.f);
                const float m2_ = fmaxf(__uint_as_float(hh[o + 2]) + b.z, 0.f), m3_ = fmaxf(__uint_as_float(hh[o + 3]) + b.w, 0.f);
                if (saving && a.sv.mh) *reinterpret_cast<float4*>(a.sv.mh + grow * 64 + 4 * k4) = make_float4(m0_, m1_, m2_, m3_);   // [B, n, 64]
                const f32x2 v01 = pack2(m0_, m1_);
                const f32x2 v23 = pack2(m2_, m3_);
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
                // dense SWIZZLE_128B staging in HBM order (state-major, the padding rows of odd n squeezed out); ONE 3-D tensor
                // store (32 floats x n nodes x SPT states) writes the tile, clipped to the batch by the tensor map
                const int srow = s_loc * N + node;
                if (row_used) {
                    const uint32_t rp = xf_s + srow * 128;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        sts128s(rp + (((c ^ srow) & 7) << 4), make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]));
                }
                fence_proxy_async();
                group_sync();
                if (gt == 0) {
                    tma_store_3d(&mapH, 0, 0, (int)s0, xf_s);
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
                    *reinterpret_cast<float4*>(dst + (size_t)idx * 4) = lds128s(row_ptr(xf_s, s * NP + i) + c * 16);
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
static size_t tp_smem_bytes(int L, bool motion, int G, bool save) {
    return 1024 + ((size_t)tc_graph_floats(L) + (motion ? TMOTION_FLOATS : 0) + (size_t)G * (XF_GROUP / 4) + (save ? (size_t)G * 4096 : 0)) * 4 + (2 + G) * 8 + 16;
}

template <int N, int G, bool SAVE>
static cudaError_t launch_tp(const GraphArgs& a, int num_sms, size_t max_smem, cudaStream_t st) {
    const size_t smem = tp_smem_bytes(a.L, a.mw != nullptr, G, SAVE);
    if (smem > max_smem) return cudaErrorInvalidConfiguration;
    GraphArgs b = a;
    if (pdl_early(a.B)) b.flags |= RGL_INTERNAL_PDL_EARLY;
    constexpr int spt = 128 / ((N + 1) & ~1);
    b.ntiles = (a.B + spt - 1) / spt;
    if (cudaError_t e = ensure_dyn_smem(graph_forward_tp_kernel<N, G, SAVE>, (int)max_smem)) return e;
    int per_sm = (int)((228 * 1024) / (smem + 1024));
    const int max_cta = G <= 2 ? 2 : 1;                       // matches __launch_bounds__ and the 512 TMEM columns of an SM
    if (per_sm > max_cta) per_sm = max_cta;
    if (per_sm < 1) per_sm = 1;
    const int want = (b.ntiles + G - 1) / G;
    const int grid = want < num_sms * per_sm ? want : num_sms * per_sm;
    // H leaves through a TMA tensor store when the driver offers cuTensorMapEncodeTiled (RGL_TC_TMA_OUT=0 disables: experiments)
    static const char* tma_env = getenv("RGL_TC_TMA_OUT");
    CUtensorMap mh;
    memset(&mh, 0, sizeof(mh));
    int tma_out = 0;
    if (a.H != nullptr && !(tma_env && tma_env[0] == '0')) tma_out = make_state_map(&mh, a.H, a.B, N, N, spt) ? 1 : 0;
    return launch_pdl(graph_forward_tp_kernel<N, G, SAVE>, dim3(grid), dim3(128 * G), smem, st, b, mh, tma_out);
}

template <int N>
static cudaError_t dispatch_tp(const GraphArgs& a, int num_sms, size_t max_smem, cudaStream_t st) {
    // same group-count policy as graph_forward_tc.cu (RGL_TC_GROUPS forces it: experiments only)
    static const char* force = getenv("RGL_TC_GROUPS");
    constexpr int spt = 128 / ((N + 1) & ~1);
    const int ntiles = (a.B + spt - 1) / spt;
    if (a.save)         // training forward with activation saves (TC layout: kernels.h GraphArgs::save == 2)
        return ntiles <= 2 * num_sms ? launch_tp<N, 1, true>(a, num_sms, max_smem, st) : launch_tp<N, 2, true>(a, num_sms, max_smem, st);
    int g = force ? atoi(force) : 0;
    if (g != 1 && g != 2 && g != 4) {
        if (a.mw != nullptr) g = ntiles <= 2 * num_sms ? 1 : 4;
        else g = (ntiles <= 2 * num_sms && !(a.flags & RGL_FLAG_THROUGHPUT)) ? 1 : 2;
    }
    if (g == 1) return launch_tp<N, 1, false>(a, num_sms, max_smem, st);
    if (g == 2) return launch_tp<N, 2, false>(a, num_sms, max_smem, st);
    return launch_tp<N, 4, false>(a, num_sms, max_smem, st);
}

// Node counts with a compile-time instantiation (Nh = 5, 10, 20: the BASELINE configurations); everything else runs on
// graph_forward_tc.cu.  Returns cudaErrorNotSupported when this kernel does not cover the shape.
cudaError_t run_graph_forward_tp(const GraphArgs& a, int num_sms, size_t max_smem, cudaStream_t st) {
    switch (a.Nh + 1) {
        case 6: return dispatch_tp<6>(a, num_sms, max_smem, st);
        case 11: return dispatch_tp<11>(a, num_sms, max_smem, st);
        case 21: return dispatch_tp<21>(a, num_sms, max_smem, st);
        default: return cudaErrorNotSupported;
    }
}

}  // namespace rgl
