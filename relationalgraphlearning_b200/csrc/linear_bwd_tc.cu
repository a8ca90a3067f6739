// Linear-layer backward on the 5th-generation tensor cores (tcgen05 / TMEM) for the 32-wide layers of the training step
// (crowd_nav/utils/trainer.py:122-131 -> loss.backward() through the GCN layers, w_a and the second embedding layers).
//
//   y = x W over R rows, N = 32 output columns, K = 32 * KA input columns (KA = 1: GCN layers / w_a, KA = 2: w_r.2 / w_h.2)
//   data gradient     Gin[r, :] (+)= (G[r, :] . mask) W^T          one UMMA M = 128 rows, A = G hi|lo in TMEM (thread = row,
//                                                                   3xTF32), B = W^T tiles resident in shared memory
//   weight gradient   dW += X^T (G . mask)                          the ROWS of the tile are the contraction: both operands are
//                                                                   MN-major shared-memory tiles, A = [X hi ; X lo] (M = 64 KA),
//                                                                   B = G hi, then B = G lo (4 terms of the split product);
//                                                                   the accumulator stays in TMEM for the life of the CTA
//   bias gradient     db += column sums of G . mask                 from the split tiles in shared memory
//
// The mma.sync kernel of train_kernels.cu spends 40 % of its instructions re-splitting operands into tf32 hi / lo for every
// fragment that uses them (ncu: 736 warp-level splits and 7 400 warp instructions per 64-row tile, 5 % of them HMMA, issue
// slots 46 % busy, 2.5 TB/s).  Here every element is split ONCE by the thread that owns its row, the tensor core reads the
// operands from shared memory / TMEM by itself, and a 128-row tile costs about 2 000 warp instructions.
//
// MN-major tf32 operands exist only in the SWIZZLE_128B_BASE32B shared-memory layout (layout type 1): 128-byte rows,
// 4-row atoms of 512 B, 32-byte chunks XOR-ed with (row & 3); LBO = distance between the 32-column atoms along M,
// SBO = 512; M = 64 accumulators live in lanes 32 (m / 16) + m % 16.  All of it validated on hardware by tools/umma_probe.cu
// test 12 (profiles/r2_umma_mn_probe.md).
#include <stdlib.h>
#include "kernels.h"
#include "tc_common.cuh"
#include "train_common.cuh"

namespace rgl {

__device__ __forceinline__ void cp_async16_tc(uint32_t dst_s, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst_s), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_tc() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all_tc() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], kind::tf32, both operands through shared-memory descriptors
__device__ __forceinline__ void umma_tf32_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// MN-major descriptor, SWIZZLE_128B_BASE32B: LBO = 16 KB (next 32-column atom = next tile), SBO = 512 B (next 4 rows)
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(16384 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)1 << 61);
}
// byte offset of 16-byte chunk c16 of row r inside a [128][32] fp32 tile in that layout
__device__ __forceinline__ uint32_t sw32(int r, int c16) {
    return (uint32_t)r * 128u + (uint32_t)((((c16 >> 1) ^ r) & 3) << 5) + (uint32_t)((c16 & 1) << 4);
}
// K-major SWIZZLE_128B (the W^T tiles of the data gradient): chunk c16 of row r
__device__ __forceinline__ uint32_t sw128k(int r, int c16) { return (uint32_t)r * 128u + (uint32_t)(((c16 ^ r) & 7) << 4); }

__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
    const uint32_t h0 = (__float_as_uint(v.x) + 0x1000u) & 0xffffe000u, h1 = (__float_as_uint(v.y) + 0x1000u) & 0xffffe000u;
    const uint32_t h2 = (__float_as_uint(v.z) + 0x1000u) & 0xffffe000u, h3 = (__float_as_uint(v.w) + 0x1000u) & 0xffffe000u;
    hi = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(h2), __uint_as_float(h3));
    lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);       // exact; the tensor core drops the low 13 bits
}

// offsets (in floats) of rows first, first + STEP, ... of a grouped-row matrix: one division, then increments
template <int CNT, int STEP>
__device__ __forceinline__ void row_offsets(const Rows& R, int first, long long (&off)[CNT]) {
    if (R.rpg == 1) {
#pragma unroll
        for (int i = 0; i < CNT; ++i) off[i] = (long long)(first + STEP * i) * R.gstride;
        return;
    }
    int grp = first / R.rpg, rem = first - grp * R.rpg;
#pragma unroll
    for (int i = 0; i < CNT; ++i) {
        off[i] = (long long)grp * R.gstride + (long long)rem * R.ld;
        rem += STEP;
        while (rem >= R.rpg) { rem -= R.rpg; ++grp; }
    }
}

template <int KA>
__global__ void __launch_bounds__(128, KA == 1 ? 3 : 2) rows_linear_bwd_tc_kernel(const LinBwdArgs a) {
    constexpr int KL = 32 * KA;                                   // input columns of the layer
    constexpr uint32_t TCOLS = KA == 1 ? 128 : 256;
    constexpr uint32_t C_DD = 0, C_AH = KL, C_AL = KL + 32, C_DW = KL + 64;
    constexpr uint32_t TILE = 16384;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    // [W^T hi | W^T lo] (KL rows x 128 B each)   [G -> G hi] [mask -> G lo] [X -> X hi] x KA [X lo] x KA   [barriers]
    uint8_t* Wh = sm;
    uint8_t* Wl = Wh + KL * 128;
    uint8_t* Tg = Wl + KL * 128;
    uint8_t* Tm = Tg + TILE;
    uint8_t* Tx = Tm + TILE;
    uint8_t* Txl = Tx + KA * TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Txl + KA * TILE);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t sm_s = __shfl_sync(0xffffffffu, smem_u32(sm), 0);
    const uint32_t wh_s = sm_s, wl_s = wh_s + KL * 128, tg_s = wl_s + KL * 128, tm_s = tg_s + TILE, tx_s = tm_s + TILE, txl_s = tx_s + KA * TILE;
    const bool data_grad = a.W != nullptr && a.Gin.ptr != nullptr;
    const bool has_mask = a.mask.ptr != nullptr;
    // stage G (+ mask) and X of one tile: 8 consecutive threads copy the 8 chunks of a row; 16 rows per step
    auto stage = [&](int tile) {
        const int r0 = tile * 128;
        long long og[8], om[8];
        row_offsets<8, 16>(a.G, r0 + (tid >> 3), og);
        if (has_mask) row_offsets<8, 16>(a.mask, r0 + (tid >> 3), om);
        const int c = tid & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = (tid >> 3) + 16 * i;
            const bool in = r0 + r < a.R;
            cp_async16_tc(tg_s + sw32(r, c), in ? a.G.ptr + og[i] + 4 * c : a.G.ptr, in ? 16u : 0u);
            if (has_mask) cp_async16_tc(tm_s + sw32(r, c), in ? a.mask.ptr + om[i] + 4 * c : a.mask.ptr, in ? 16u : 0u);
        }
        if (a.dW) {
            constexpr int CPR = 8 * KA, RPI = 128 / CPR;       // chunks per row, rows per step
            long long ox[CPR];
            row_offsets<CPR, RPI>(a.Xin, r0 + tid / CPR, ox);
            const int cx = tid % CPR;
#pragma unroll
            for (int i = 0; i < CPR; ++i) {
                const int r = tid / CPR + RPI * i;
                const bool in = r0 + r < a.R;
                cp_async16_tc(tx_s + (cx >> 3) * TILE + sw32(r, cx & 7), in ? a.Xin.ptr + ox[i] + 4 * cx : a.Xin.ptr, in ? 16u : 0u);
            }
        }
        cp_async_commit_tc();
    };
    if (blockIdx.x < a.ntiles) stage(blockIdx.x);                // in flight under the TMEM allocation and the W^T staging

    if (warp == 0) tmem_alloc(tslot, TCOLS);
    if (tid == 0) { mbar_init(bars, 1); mbar_init(bars + 1, 1); fence_mbar_init(); }

    // ---- W^T tiles (B operand of the data gradient, K-major): row k = input column, 32 floats = the output columns ----
    if (a.W && a.Gin.ptr) {
        float w[8 * KA];                                          // every load in flight before the first use
#pragma unroll
        for (int i = 0; i < 8 * KA; ++i) {
            const int idx = tid + 128 * i;
            int k, nn;
            if (a.w_layout == 1) { k = idx >> 5; nn = idx & 31; } else { nn = idx / KL; k = idx - nn * KL; }
            w[i] = a.w_layout == 1 ? a.W[k * 32 + nn] : a.W[nn * KL + k];
        }
#pragma unroll
        for (int i = 0; i < 8 * KA; ++i) {
            const int idx = tid + 128 * i;
            int k, nn;
            if (a.w_layout == 1) { k = idx >> 5; nn = idx & 31; } else { nn = idx / KL; k = idx - nn * KL; }
            const uint32_t h = (__float_as_uint(w[i]) + 0x1000u) & 0xffffe000u;
            const uint32_t o = sw128k(k, nn >> 2) + (uint32_t)(nn & 3) * 4u;
            *reinterpret_cast<uint32_t*>(Wh + o) = h;
            *reinterpret_cast<float*>(Wl + o) = w[i] - __uint_as_float(h);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = __shfl_sync(0xffffffffu, *tslot, 0);
    const uint32_t tl = tbase + ((uint32_t)(warp * 32) << 16);
    float bacc = 0.f;
    uint32_t parity = 0;
    int it = 0;

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
        const int r0 = tile * 128;
        if (it > 0) stage(tile);
        // this thread's row of Gin (old values for +=), fetched under the staging copies
        const bool row_in = r0 + tid < a.R;
        float* gin_row = data_grad ? a.Gin.row(row_in ? r0 + tid : 0) : nullptr;
        float4 old[8 * KA];
        if (data_grad && a.accumulate && row_in) {
#pragma unroll
            for (int c = 0; c < 8 * KA; ++c) old[c] = reinterpret_cast<const float4*>(gin_row)[c];
        }
        cp_async_wait_all_tc();
        __syncthreads();

        // ---- thread = row: mask, split once; G hi|lo -> TMEM (A of the data gradient) and -> the tiles (B of the weight gradient) ----
        {
#pragma unroll
            for (int h = 0; h < 2; ++h) {                        // 16 columns at a time
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t o = sw32(tid, 4 * h + c);
                    float4 v = lds128s(tg_s + o);
                    if (has_mask) {
                        const float4 m = lds128s(tm_s + o);
                        v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
                    }
                    float4 vh, vl;
                    split4(v, vh, vl);
                    sts128s(tg_s + o, vh);
                    sts128s(tm_s + o, vl);
                    hi[4 * c] = __float_as_uint(vh.x); hi[4 * c + 1] = __float_as_uint(vh.y); hi[4 * c + 2] = __float_as_uint(vh.z); hi[4 * c + 3] = __float_as_uint(vh.w);
                    lo[4 * c] = __float_as_uint(vl.x); lo[4 * c + 1] = __float_as_uint(vl.y); lo[4 * c + 2] = __float_as_uint(vl.z); lo[4 * c + 3] = __float_as_uint(vl.w);
                }
                if (data_grad) {
                    tmem_st16(tl + C_AH + 16 * h, hi);
                    tmem_st16(tl + C_AL + 16 * h, lo);
                }
            }
            if (a.dW) {
#pragma unroll
                for (int c = 0; c < 8 * KA; ++c) {
                    const uint32_t o = (c >> 3) * TILE + sw32(tid, c & 7);
                    float4 vh, vl;
                    split4(lds128s(tx_s + o), vh, vl);
                    sts128s(tx_s + o, vh);
                    sts128s(txl_s + o, vl);
                }
            }
        }
        fence_proxy_async();
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();

        if (warp == 0) {
            if (elect_one()) {
                tc_fence_after();
                if (data_grad) {
                    issue_gemm<4>(tbase + C_DD, tbase + C_AH, tbase + C_AL, wh_s, wl_s, umma_idesc(128, KL), 0);
                    umma_commit(bars);
                }
                if (a.dW) {
                    constexpr uint32_t idesc_mn = umma_idesc(64 * KA, 32) | (1u << 15) | (1u << 16);
                    uint32_t acc = it > 0 ? 1u : 0u;
#pragma unroll 4
                    for (int ks = 0; ks < 16; ++ks) {
                        const uint64_t ad = umma_desc_mn(tx_s + ks * 1024);
                        umma_tf32_ss(tbase + C_DW, ad, umma_desc_mn(tm_s + ks * 1024), idesc_mn, acc);      // x G lo first (small terms)
                        umma_tf32_ss(tbase + C_DW, ad, umma_desc_mn(tg_s + ks * 1024), idesc_mn, 1);
                        acc = 1;
                    }
                    umma_commit(bars + 1);
                }
            }
            __syncwarp();
        }
        // ---- bias gradient: column sums of the masked tile (hi + lo), a quarter of the rows per warp ----
        if (a.db) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
            for (int r = warp * 32; r < warp * 32 + 32; ++r) {
                const uint32_t o = sw32(r, lane >> 2) + (uint32_t)(lane & 3) * 4u;
                float h, l;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(h) : "r"(tg_s + o));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l) : "r"(tm_s + o));
                s0 += h; s1 += l;
            }
            bacc += s0 + s1;
        }
        // ---- data-gradient epilogue: accumulator row -> (+ old) -> Gin ----
        if (data_grad) {
            mbar_wait_sleepy(bars, parity);
            tc_fence_after();
#pragma unroll
            for (int q = 0; q < KA; ++q) {
                uint32_t d[32];
                tmem_ld32(tl + C_DD + 32 * q, d);
                if (row_in) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float4 v = make_float4(__uint_as_float(d[4 * c]), __uint_as_float(d[4 * c + 1]), __uint_as_float(d[4 * c + 2]), __uint_as_float(d[4 * c + 3]));
                        if (a.accumulate) { const float4 o = old[8 * q + c]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                        reinterpret_cast<float4*>(gin_row)[8 * q + c] = v;
                    }
                }
            }
        }
        if (a.dW) mbar_wait_sleepy(bars + 1, parity);            // the tiles are free again
        parity ^= 1;
        tc_fence_before();
        __syncthreads();
    }

    // ---- flush: dW accumulator (hi rows + lo rows) and the bias sums ----
    if (a.dW && it > 0) {
        tc_fence_after();
        uint32_t d[32];
        tmem_ld32(tl + C_DW, d);
        float* scratch = reinterpret_cast<float*>(Tg);             // [2][KL][33]  (the G tiles are free: 32 KB)
        // M = 64 (KA = 1): rows m live in lanes 32 (m / 16) + m % 16; rows 0..31 = X hi features, 32..63 = X lo.  M = 128: lane = row.
        const bool holds = KA == 2 || lane < 16;
        const int m = KA == 2 ? tid : (warp * 16 + lane);
        if (holds) {
#pragma unroll
            for (int n = 0; n < 32; ++n) scratch[m * 33 + n] = __uint_as_float(d[n]);
        }
        __syncthreads();
        // every thread flushes 8 KA elements (hi part + lo part) as 2 KA vector reductions of four consecutive addresses of dW
#pragma unroll
        for (int i = 0; i < 2 * KA; ++i) {
            const int idx = (tid + 128 * i) * 4;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int f = a.w_layout == 0 ? ((idx + j) & (KL - 1)) : (idx >> 5), n = a.w_layout == 0 ? idx / KL : ((idx + j) & 31);
                v[j] = scratch[f * 33 + n] + scratch[(KL + f) * 33 + n];
            }
            float* dst = a.dW + idx;            // both layouts: element idx of the dense [.][.] matrix
            if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) atomicAdd(dst + j, v[j]);
            }
        }
    }
    if (a.db && it > 0) atomicAdd(a.db + lane, bacc);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TCOLS);
}

template <int KA>
static cudaError_t launch_bwd_tc(LinBwdArgs& a, int num_sms, size_t max_smem, cudaStream_t st) {
    constexpr size_t smem = 1024 + (size_t)2 * 32 * KA * 128 + (size_t)(2 + 2 * KA) * 16384 + 64;
    if (smem > max_smem) return cudaErrorInvalidConfiguration;
    if (cudaError_t e = ensure_dyn_smem(rows_linear_bwd_tc_kernel<KA>, (int)max_smem)) return e;
    a.ntiles = (a.R + 127) / 128;
    const int cap = num_sms * (KA == 1 ? 3 : 2);
    const int waves = (a.ntiles + cap - 1) / cap;                  // every CTA takes the same number of tiles (+-1)
    const int grid = (a.ntiles + waves - 1) / waves;
    rows_linear_bwd_tc_kernel<KA><<<grid, 128, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t run_linear_bwd_tc(LinBwdArgs& a, int num_sms, size_t max_smem, cudaStream_t st) {
    if (a.N != 32 || !a.vecG) return cudaErrorNotSupported;
    if (a.dW && !a.vecX) return cudaErrorNotSupported;
    if (a.Gin.ptr && ((reinterpret_cast<uintptr_t>(a.Gin.ptr) & 15u) || (a.Gin.ld & 3) || (a.Gin.gstride & 3))) return cudaErrorNotSupported;
    // Measured at C4 (tools/bwd_time.py, B = 8192, Nh = 10): K = 32 over B*n rows 19.6 us vs 24.6 us for the mma.sync kernel;
    // K = 64 (35 vs 28 us) and the B-row launches (one tile per CTA: 6.2 vs 6.0 us) stay on mma.sync unless RGL_BWD_VARIANT=t.
    // A 128-row tile costs 44 UMMA instructions of K = 8 (12 data-gradient + 32 weight-gradient) whose fixed issue cost, about
    // 1.5 us per tile on the SM's one tensor pipe, is what bounds this kernel (globaltimer trace: transform 1.7 us, MMA phase
    // 1.5 us, epilogue 0.6 us per tile; two tiles per CTA, three CTAs per SM).
    static const char* variant = getenv("RGL_BWD_VARIANT");
    const bool force = variant && variant[0] == 't';
    if (a.K == 32 && (force || (a.R + 127) / 128 > num_sms)) return launch_bwd_tc<1>(a, num_sms, max_smem, st);
    if (a.K == 64 && force) return launch_bwd_tc<2>(a, num_sms, max_smem, st);
    return cudaErrorNotSupported;
}

}  // namespace rgl
