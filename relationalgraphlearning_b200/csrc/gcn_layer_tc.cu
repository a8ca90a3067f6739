// Stand-alone GCN layer on tcgen05 / TMEM for sm_100a:  Hout = relu(A (X W)) (+ X),  A given ([B,n,n]) or
// A = softmax(X w_a X^T) computed in-kernel  (crowd_nav/policy/graph_model.py:119-128 and :65-66), on node features that
// already live in HBM.  This is the one unit of the path whose roofline is HBM (1 536 + 144 B of compulsory traffic per
// 6-node state for 14.6 kFLOP).  Same machine mapping as graph_forward_tc.cu:
//   group = 128 threads = one UMMA M-tile = SPT = 128 / n whole states, rows in HBM order (row = state * n + node), one
//   node row per thread / TMEM lane;  X W (and Y = X w_a, stacked along n) is one 3xTF32 tcgen05.mma chain with the A
//   operand in TMEM;  the per-state products run on the FMA pipe with the neighbours' rows in a swizzled shared-memory
//   buffer;  the output tile (one contiguous HBM block) is staged and written as 512 contiguous bytes per warp instruction.
//   The weight tiles (UMMA SWIZZLE_128B, hi / lo) are built in shared memory by the CTA from the nn.Parameter layout.
#include <stdlib.h>
#include "kernels.h"
#include "tc_common.cuh"
#include "tma_maps.cuh"

namespace rgl {

__device__ __forceinline__ uint32_t gl_row_ptr(uint32_t xf_s, int row) { return xf_s + row * 128 + ((row & 7) << 4); }

template <int N, int G>
__global__ void __launch_bounds__(128 * G, G <= 2 ? 2 : 1) gcn_layer_tc_kernel(const float* __restrict__ X, const float* __restrict__ Ag,
                                                                                 const float* __restrict__ Wg, const float* __restrict__ wag,
                                                                                 int B, int n_rt, int flags, float* __restrict__ Hout,
                                                                                 float* __restrict__ Aout, int ntiles) {
    constexpr int NMAX = N > 0 ? N : RGL_MAX_HUMANS + 1;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* tw = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));   // hi [64][32] | lo [64][32]
    float* xf_all = tw + 4096;
    uint64_t* bars = reinterpret_cast<uint64_t*>(xf_all + G * 4096);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + G);

    const int n = N > 0 ? N : n_rt;
    const int SPT = 128 / n;
    const int tid = threadIdx.x, lane = tid & 31, gt = tid & 127;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int grp = warp >> 2, wq = warp & 3;
    const bool skip = flags & RGL_FLAG_SKIP;
    const bool sim = Ag == nullptr;                      // attention computed in-kernel from w_a

    if (warp == 0) tmem_alloc(tslot, 128 * G);
    if (tid == 0) {
        for (int i = 0; i < G; ++i) mbar_init(bars + i, 1);
        fence_mbar_init();
    }
    // B-operand tiles: row nn of the tile holds column nn of the [k][n] parameter; with w_a: rows 0-31 = w_a^T, rows 32-63 = W^T
    for (int idx = tid; idx < (sim ? 2048 : 1024); idx += blockDim.x) {
        const int m = idx >> 10, k = (idx >> 5) & 31, nn = idx & 31;           // coalesced over nn
        const float w = (sim && m == 0) ? __ldg(wag + k * XD + nn) : __ldg(Wg + k * XD + nn);
        const int row = (sim ? m * 32 : 0) + nn;
        const int o = row * 32 + ((((k >> 2) ^ row) & 7) << 2) + (k & 3);
        const float hi = __uint_as_float((__float_as_uint(w) + 0x1000u) & 0xffffe000u);
        tw[o] = hi;
        tw[2048 + o] = w - hi;
    }
    fence_proxy_async();                                 // generic-proxy stores -> visible to the tensor core's operand fetch
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const uint32_t tbase = __shfl_sync(0xffffffffu, *tslot, 0);
    const uint32_t tg = tbase + grp * 128;
    const uint32_t tl = tg + ((uint32_t)(wq * 32) << 16);
    const uint32_t tw_s = __shfl_sync(0xffffffffu, smem_u32(tw), 0);
    const uint32_t xf_s = __shfl_sync(0xffffffffu, smem_u32(xf_all), 0) + grp * 16384;
    uint64_t* gbar = bars + grp;
    uint32_t par = 0;
    const bool issuer = wq == 0;
    auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" :: "r"(grp + 1) : "memory"); };
    auto mma_wait = [&]() { mbar_wait(gbar, par); par ^= 1; tc_fence_after(); };
    auto publish = [&]() { tmem_st_wait(); tc_fence_before(); group_sync(); };

    const int s_loc = gt / n;
    const bool row_used = gt < SPT * n;
    const uint32_t my_row = gl_row_ptr(xf_s, gt);
    const int srow0 = (row_used ? s_loc : 0) * n;        // first row of this thread's state (tail rows: state 0, reads stay inside the buffer)
    const int tstride = gridDim.x * G;

    for (int tile = blockIdx.x * G + grp; tile < ntiles; tile += tstride) {
        const long s0 = (long)tile * SPT;
        const int cnt = (int)min((long)SPT, (long)B - s0);
        const bool valid = row_used && s_loc < cnt;
        const long grow = s0 * n + gt;                   // global node row

        // ---- this thread's feature row (and attention row); next tile's rows are pulled into L2 meanwhile ----
        float x[32];
        {
            const float4* src = reinterpret_cast<const float4*>(X + grow * XD);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 v = valid ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
            }
        }
        float p[NMAX];
        if (!sim) {
#pragma unroll
            for (int j = 0; j < NMAX; ++j)
                if (N > 0 || j < n) p[j] = valid ? __ldg(Ag + grow * n + j) : 0.f;
        }
        if (tile + tstride < ntiles) {
            const long nrow = (long)(tile + tstride) * SPT * n + gt;
            if (row_used && nrow < (long)B * n) {
                asm volatile("prefetch.global.L2 [%0];" :: "l"(X + nrow * XD));
                if (!sim) asm volatile("prefetch.global.L2 [%0];" :: "l"(Ag + nrow * n));
            }
        }
        if (sim) {
#pragma unroll
            for (int c = 0; c < 8; ++c) sts128s(my_row ^ (c << 4), make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]));
        }
        st_split<32>(tl + 64, tl + 96, x);
        publish();
        if (issuer) {
            if (lane == 0) {
                tc_fence_after();
                // with w_a: one N = 64 chain, columns [0,32) = Y = X w_a, [32,64) = X W; otherwise N = 32 into columns [32,64)
                if (sim) issue_gemm<4>(tg, tg + 64, tg + 96, tw_s, tw_s + 8192, umma_idesc(128, 64), 0);
                else issue_gemm<4>(tg + 32, tg + 64, tg + 96, tw_s, tw_s + 8192, umma_idesc(128, 32), 0);
                umma_commit(gbar);
            }
            __syncwarp();
        }
        mma_wait();

        if (sim) {
            uint32_t yr[32];
            tmem_ld32(tl, yr);
            if (row_used) {
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < NMAX; ++j) {
                    if (N > 0 || j < n) {
                        const uint32_t rp = gl_row_ptr(xf_s, srow0 + j);
                        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 xv = lds128s(rp ^ (c << 4));
                            d0 = fmaf(__uint_as_float(yr[4 * c + 0]), xv.x, d0);
                            d1 = fmaf(__uint_as_float(yr[4 * c + 1]), xv.y, d1);
                            d2 = fmaf(__uint_as_float(yr[4 * c + 2]), xv.z, d2);
                            d3 = fmaf(__uint_as_float(yr[4 * c + 3]), xv.w, d3);
                        }
                        p[j] = (d0 + d1) + (d2 + d3);
                        mx = fmaxf(mx, p[j]);
                    }
                }
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < NMAX; ++j)
                    if (N > 0 || j < n) { p[j] = expf(p[j] - mx); sum += p[j]; }
#pragma unroll
                for (int j = 0; j < NMAX; ++j)
                    if (N > 0 || j < n) p[j] = p[j] / sum;
            }
            group_sync();                                // every read of the feature rows is done
        }
        if (Aout != nullptr && valid) {
#pragma unroll
            for (int j = 0; j < NMAX; ++j)
                if (N > 0 || j < n) Aout[grow * n + j] = p[j];
        }

        // ---- X W rows -> xf;  H' = relu(sum_j A[i][j] (X W)[j]) (+ X) ----
        {
            uint32_t hw[32];
            tmem_ld32(tl + 32, hw);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                sts128s(my_row ^ (c << 4), make_float4(__uint_as_float(hw[4 * c]), __uint_as_float(hw[4 * c + 1]),
                                                       __uint_as_float(hw[4 * c + 2]), __uint_as_float(hw[4 * c + 3])));
        }
        tc_fence_before();
        group_sync();
        float acc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = 0.f;
        if (row_used) {
#pragma unroll
            for (int j = 0; j < NMAX; ++j) {
                if (N > 0 || j < n) {
                    const uint32_t rp = gl_row_ptr(xf_s, srow0 + j);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 hv = lds128s(rp ^ (c << 4));
                        acc[4 * c + 0] = fmaf(p[j], hv.x, acc[4 * c + 0]);
                        acc[4 * c + 1] = fmaf(p[j], hv.y, acc[4 * c + 1]);
                        acc[4 * c + 2] = fmaf(p[j], hv.z, acc[4 * c + 2]);
                        acc[4 * c + 3] = fmaf(p[j], hv.w, acc[4 * c + 3]);
                    }
                }
            }
        }
        if (skip) {
#pragma unroll
            for (int c = 0; c < 32; ++c) x[c] += fmaxf(acc[c], 0.f);
        } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) x[c] = fmaxf(acc[c], 0.f);
        }

        // ---- stage the output rows, copy the tile out as one contiguous block ----
        group_sync();                                    // every read of the X W rows is done
#pragma unroll
        for (int c = 0; c < 8; ++c) sts128s(my_row ^ (c << 4), make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]));
        group_sync();
        float* dst = Hout + s0 * n * XD;
        const int chunks = cnt * n * 8;
        for (int idx = gt; idx < chunks; idx += 128)
            *reinterpret_cast<float4*>(dst + (size_t)idx * 4) = lds128s(gl_row_ptr(xf_s, idx >> 3) ^ ((idx & 7) << 4));
        // the next tile's first write to xf (sim) / TMEM comes after its own loads; order it after this tile's copy-out reads
        group_sync();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 128 * G);
}


// ---------------------------------------------------------------------------------------------------------------
// TMA variant: the input tile (one contiguous block of SPT*n rows) is fetched by a 2-D tensor copy into a swizzled
// shared-memory buffer, double-buffered one tile ahead, and the output tile leaves through a tensor store: no
// row-per-thread global accesses (those cost one L1 wavefront per thread and put the LSU pipe at 82%).
template <int N, int G>
__global__ void __launch_bounds__(128 * G, G <= 2 ? 2 : 1) gcn_layer_tma_kernel(const __grid_constant__ CUtensorMap mapX,
                                                                                  const __grid_constant__ CUtensorMap mapH,
                                                                                  const float* __restrict__ Ag, const float* __restrict__ Wg,
                                                                                  const float* __restrict__ wag, int B, int n_rt, int flags,
                                                                                  float* __restrict__ Aout, int ntiles) {
    constexpr int NMAX = N > 0 ? N : RGL_MAX_HUMANS + 1;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* tw = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));   // hi [64][32] | lo [64][32]
    float* buf_all = tw + 4096;                          // per group: in[0] | in[1] | xw   (3 x 16 KB, 1 KB aligned)
    uint64_t* bars = reinterpret_cast<uint64_t*>(buf_all + G * 3 * 4096);     // per group: mma, full[0], full[1]
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 3 * G);

    const int n = N > 0 ? N : n_rt;
    const int SPT = 128 / n;
    const int tid = threadIdx.x, lane = tid & 31, gt = tid & 127;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int grp = warp >> 2, wq = warp & 3;
    const bool skip = flags & RGL_FLAG_SKIP;
    const bool sim = Ag == nullptr;

    if (warp == 0) tmem_alloc(tslot, 128 * G);
    if (tid == 0) {
        for (int i = 0; i < 3 * G; ++i) mbar_init(bars + i, 1);
        fence_mbar_init();
    }
    for (int idx = tid; idx < (sim ? 2048 : 1024); idx += blockDim.x) {
        const int m = idx >> 10, k = (idx >> 5) & 31, nn = idx & 31;
        const float w = (sim && m == 0) ? __ldg(wag + k * XD + nn) : __ldg(Wg + k * XD + nn);
        const int row = (sim ? m * 32 : 0) + nn;
        const int o = row * 32 + ((((k >> 2) ^ row) & 7) << 2) + (k & 3);
        const float hi = __uint_as_float((__float_as_uint(w) + 0x1000u) & 0xffffe000u);
        tw[o] = hi;
        tw[2048 + o] = w - hi;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const uint32_t tbase = __shfl_sync(0xffffffffu, *tslot, 0);
    const uint32_t tg = tbase + grp * 128;
    const uint32_t tl = tg + ((uint32_t)(wq * 32) << 16);
    const uint32_t tw_s = __shfl_sync(0xffffffffu, smem_u32(tw), 0);
    const uint32_t buf_s = __shfl_sync(0xffffffffu, smem_u32(buf_all), 0) + grp * 49152;
    const uint32_t xw_s = buf_s + 32768;
    uint64_t* gbar = bars + 3 * grp;
    uint64_t* full = gbar + 1;
    uint32_t par = 0;
    const bool issuer = wq == 0;
    auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" :: "r"(grp + 1) : "memory"); };
    auto mma_wait = [&]() { mbar_wait(gbar, par); par ^= 1; tc_fence_after(); };
    auto publish = [&]() { tmem_st_wait(); tc_fence_before(); group_sync(); };

    const int s_loc = gt / n;
    const bool row_used = gt < SPT * n;
    const int srow0 = (row_used ? s_loc : 0) * n;        // tail rows read state 0: every read stays inside the row buffers
    const int tstride = gridDim.x * G;
    const int rows_tile = SPT * n;
    const uint32_t tile_bytes = (uint32_t)rows_tile * 128u;

    int tile = blockIdx.x * G + grp;
    if (gt == 0 && tile < ntiles) {
        mbar_arrive_expect_tx(full + 0, tile_bytes);
        tma_load_2d(buf_s, &mapX, 0, tile * rows_tile, full + 0);
    }
    for (int it = 0; tile < ntiles; tile += tstride, ++it) {
        const int b = it & 1;
        const uint32_t in_s = buf_s + b * 16384;
        const long s0 = (long)tile * SPT;
        const int cnt = (int)min((long)SPT, (long)B - s0);
        const bool valid = row_used && s_loc < cnt;
        const long grow = s0 * n + gt;

        // next tile's features -> the other input buffer (its last reader was the tensor store of tile it-1)
        if (gt == 0 && tile + tstride < ntiles) {
            tma_store_wait_read();
            mbar_arrive_expect_tx(full + (b ^ 1), tile_bytes);
            tma_load_2d(buf_s + (b ^ 1) * 16384, &mapX, 0, (tile + tstride) * rows_tile, full + (b ^ 1));
        }
        float p[NMAX];
        if (!sim) {
#pragma unroll
            for (int j = 0; j < NMAX; ++j)
                if (N > 0 || j < n) p[j] = valid ? __ldg(Ag + grow * n + j) : 0.f;
            if (tile + tstride < ntiles && row_used) {
                const long nrow = (long)(tile + tstride) * rows_tile + gt;
                if (nrow < (long)B * n) asm volatile("prefetch.global.L2 [%0];" :: "l"(Ag + nrow * n));
            }
        }
        mbar_wait(full + b, (it >> 1) & 1);
        const uint32_t my_in = gl_row_ptr(in_s, gt), my_xw = gl_row_ptr(xw_s, gt);
        float x[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 v = lds128s(my_in ^ (c << 4));
            x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
        }
        st_split<32>(tl + 64, tl + 96, x);
        publish();
        if (issuer) {
            // with w_a: one N = 64 chain, columns [0,32) = Y = X w_a, [32,64) = X W, issued by the elect.sync lane (direct
            // UTCHMMA).  A given: N = 32 into columns [32,64); for n <= 8 issued by lane 0 -- measured: with the faster elect.sync issue
            // this HBM-bound case drops from 6.2 to 5.6 TB/s (burstier DRAM traffic), every other case gains 5-10 %.
            if (sim) {
                if (elect_one()) {
                    tc_fence_after();
                    issue_gemm<4>(tg, tg + 64, tg + 96, tw_s, tw_s + 8192, umma_idesc(128, 64), 0);
                    umma_commit(gbar);
                }
            } else {
                bool me;
                if constexpr (N > 0 && N <= 8) me = lane == 0; else me = elect_one();      // (n = 11 / 21: elect.sync is the faster one)
                if (me) {
                    tc_fence_after();
                    issue_gemm<4>(tg + 32, tg + 64, tg + 96, tw_s, tw_s + 8192, umma_idesc(128, 32), 0);
                    umma_commit(gbar);
                }
            }
            __syncwarp();
        }
        mma_wait();

        if (sim) {
            uint32_t yr[32];
            tmem_ld32(tl, yr);
            if (row_used) {
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < NMAX; ++j) {
                    if (N > 0 || j < n) {
                        const uint32_t rp = gl_row_ptr(in_s, srow0 + j);
                        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 xv = lds128s(rp ^ (c << 4));
                            d0 = fmaf(__uint_as_float(yr[4 * c + 0]), xv.x, d0);
                            d1 = fmaf(__uint_as_float(yr[4 * c + 1]), xv.y, d1);
                            d2 = fmaf(__uint_as_float(yr[4 * c + 2]), xv.z, d2);
                            d3 = fmaf(__uint_as_float(yr[4 * c + 3]), xv.w, d3);
                        }
                        p[j] = (d0 + d1) + (d2 + d3);
                        mx = fmaxf(mx, p[j]);
                    }
                }
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < NMAX; ++j)
                    if (N > 0 || j < n) { p[j] = expf(p[j] - mx); sum += p[j]; }
#pragma unroll
                for (int j = 0; j < NMAX; ++j)
                    if (N > 0 || j < n) p[j] = p[j] / sum;
            }
        }
        if (Aout != nullptr && valid) {
#pragma unroll
            for (int j = 0; j < NMAX; ++j)
                if (N > 0 || j < n) Aout[grow * n + j] = p[j];
        }

        // ---- X W rows -> xw buffer;  H' = relu(sum_j A[i][j] (X W)[j]) (+ X) ----
        {
            uint32_t hw[32];
            tmem_ld32(tl + 32, hw);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                sts128s(my_xw ^ (c << 4), make_float4(__uint_as_float(hw[4 * c]), __uint_as_float(hw[4 * c + 1]),
                                                      __uint_as_float(hw[4 * c + 2]), __uint_as_float(hw[4 * c + 3])));
        }
        tc_fence_before();
        group_sync();                                    // X W rows visible; every similarity read of the input rows is done
        float acc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = 0.f;
        if (row_used) {
#pragma unroll
            for (int j = 0; j < NMAX; ++j) {
                if (N > 0 || j < n) {
                    const uint32_t rp = gl_row_ptr(xw_s, srow0 + j);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 hv = lds128s(rp ^ (c << 4));
                        acc[4 * c + 0] = fmaf(p[j], hv.x, acc[4 * c + 0]);
                        acc[4 * c + 1] = fmaf(p[j], hv.y, acc[4 * c + 1]);
                        acc[4 * c + 2] = fmaf(p[j], hv.z, acc[4 * c + 2]);
                        acc[4 * c + 3] = fmaf(p[j], hv.w, acc[4 * c + 3]);
                    }
                }
            }
        }
        if (skip) {
#pragma unroll
            for (int c = 0; c < 32; ++c) x[c] += fmaxf(acc[c], 0.f);
        } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) x[c] = fmaxf(acc[c], 0.f);
        }
        // ---- the output rows overwrite the (dead) input rows of this tile; one tensor store writes the block (rows beyond
        // the batch are clipped by the tensor map) ----
#pragma unroll
        for (int c = 0; c < 8; ++c) sts128s(my_in ^ (c << 4), make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]));
        fence_proxy_async();                             // generic-proxy stores -> visible to the TMA engine
        group_sync();                                    // also: every read of the X W rows is done (next tile may overwrite xw)
        if (gt == 0) tma_store_2d(&mapH, 0, tile * rows_tile, in_s);
    }
    if (gt == 0) tma_store_wait_all();                   // the stores must have read shared memory (and landed) before exit
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 128 * G);
}

template <int N, int G>
static cudaError_t launch_gl_tma(const float* X, const float* A, const float* W, const float* wa, int B, int n, int flags, float* Hout,
                                 float* Aout, int num_sms, size_t max_smem, cudaStream_t st) {
    const size_t smem = 1024 + (4096 + (size_t)G * 3 * 4096) * 4 + 3 * G * 8 + 16;
    if (smem > max_smem) return cudaErrorInvalidConfiguration;
    const int spt = 128 / n;
    CUtensorMap mx, mh;
    if (!make_row_map(&mx, X, (long)B * n, spt * n) || !make_row_map(&mh, Hout, (long)B * n, spt * n)) return cudaErrorNotSupported;
    if (cudaError_t e = ensure_dyn_smem(gcn_layer_tma_kernel<N, G>, (int)max_smem)) return e;
    const int ntiles = (B + spt - 1) / spt;
    int per_sm = (int)((228 * 1024) / (smem + 1024));
    const int max_cta = G <= 2 ? 2 : 1;
    if (per_sm > max_cta) per_sm = max_cta;
    if (per_sm < 1) per_sm = 1;
    const int want = (ntiles + G - 1) / G;
    const int grid = want < num_sms * per_sm ? want : num_sms * per_sm;
    gcn_layer_tma_kernel<N, G><<<grid, 128 * G, smem, st>>>(mx, mh, A, W, wa, B, n, flags, Aout, ntiles);
    return cudaGetLastError();
}

template <int N, int G>
static cudaError_t launch_gl(const float* X, const float* A, const float* W, const float* wa, int B, int n, int flags, float* Hout,
                             float* Aout, int num_sms, size_t max_smem, cudaStream_t st) {
    const size_t smem = 1024 + (4096 + (size_t)G * 4096) * 4 + G * 8 + 16;
    if (smem > max_smem) return cudaErrorInvalidConfiguration;
    if (cudaError_t e = ensure_dyn_smem(gcn_layer_tc_kernel<N, G>, (int)max_smem)) return e;
    const int spt = 128 / n;
    const int ntiles = (B + spt - 1) / spt;
    const int per_sm = G <= 2 ? 2 : 1;                   // 512 TMEM columns per SM = four 128-column groups
    const int want = (ntiles + G - 1) / G;
    const int grid = want < num_sms * per_sm ? want : num_sms * per_sm;
    gcn_layer_tc_kernel<N, G><<<grid, 128 * G, smem, st>>>(X, A, W, wa, B, n, flags, Hout, Aout, ntiles);
    return cudaGetLastError();
}

template <int N>
static cudaError_t dispatch_gl(const float* X, const float* A, const float* W, const float* wa, int B, int n, int flags, float* Hout,
                               float* Aout, int num_sms, size_t max_smem, cudaStream_t st) {
    static const char* force = getenv("RGL_TC_GROUPS");          // experiments only
    const int spt = 128 / n, ntiles = (B + spt - 1) / spt;
    int g = force ? atoi(force) : 0;
    if (g != 1 && g != 2 && g != 4) g = ntiles <= 2 * num_sms ? 1 : 2;
    // RGL_GCN_VARIANT=l (experiments only): row-per-thread global loads / staged copy-out instead of the TMA tensor copies
    static const char* variant = getenv("RGL_GCN_VARIANT");
    if (!(variant && variant[0] == 'l')) {
        // 48 KB of tile buffers per group: one CTA of four groups per SM in the steady state
        if (!force) g = ntiles <= num_sms ? 1 : (ntiles <= 2 * num_sms ? 2 : 4);
        cudaError_t e = g == 1 ? launch_gl_tma<N, 1>(X, A, W, wa, B, n, flags, Hout, Aout, num_sms, max_smem, st)
                      : g == 2 ? launch_gl_tma<N, 2>(X, A, W, wa, B, n, flags, Hout, Aout, num_sms, max_smem, st)
                               : launch_gl_tma<N, 4>(X, A, W, wa, B, n, flags, Hout, Aout, num_sms, max_smem, st);
        if (e != cudaErrorNotSupported) return e;       // no tensor-map entry point in this driver: fall through
    }
    if (g == 1) return launch_gl<N, 1>(X, A, W, wa, B, n, flags, Hout, Aout, num_sms, max_smem, st);
    if (g == 2) return launch_gl<N, 2>(X, A, W, wa, B, n, flags, Hout, Aout, num_sms, max_smem, st);
    return launch_gl<N, 4>(X, A, W, wa, B, n, flags, Hout, Aout, num_sms, max_smem, st);
}

cudaError_t run_gcn_layer_tc(const float* X, const float* A, const float* W, const float* wa, int B, int n, int flags, float* Hout,
                             float* Aout, int num_sms, size_t max_smem, cudaStream_t st) {
    switch (n) {
        case 6: return dispatch_gl<6>(X, A, W, wa, B, n, flags, Hout, Aout, num_sms, max_smem, st);
        case 11: return dispatch_gl<11>(X, A, W, wa, B, n, flags, Hout, Aout, num_sms, max_smem, st);
        case 21: return dispatch_gl<21>(X, A, W, wa, B, n, flags, Hout, Aout, num_sms, max_smem, st);
        default: return dispatch_gl<0>(X, A, W, wa, B, n, flags, Hout, Aout, num_sms, max_smem, st);
    }
}

}  // namespace rgl
