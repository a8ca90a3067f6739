// Shared device helpers + packed-weight layout for the RGL sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/rgl_b200.h"

namespace rgl {

constexpr int XD  = RGL_X_DIM;          // 32
constexpr int HID = RGL_EMB_HIDDEN;     // 64
constexpr int RD  = RGL_ROBOT_DIM;      // 9
constexpr int HD  = RGL_HUMAN_DIM;      // 5
constexpr int LDX = XD + 4;             // padded smem row stride (floats): 36 = 4 mod 32 -> 8 consecutive
                                        // rows x 16 B hit 8 distinct bank groups (conflict-free LDS.128)

// ---- packed graph blob (floats), all matrices k-major W[k][N] -----------------------------------
// (PyTorch Linear stores [out,in]; the pack kernel transposes.  w_a / Ws are used as X @ W and are
//  already [k][N].)
constexpr int LDW = XD + 8;                     // row stride of the 32-column weight matrices: 40 = 8 mod 32, so the
                                                // mma B-fragment pattern (row k0+t, column n0+g) hits 32 distinct banks;
                                                // LDS.128 users only need a multiple of 4
constexpr int G_WR0 = 0;                       // [9][64]
constexpr int G_BR0 = G_WR0 + RD * HID;        // [64]
constexpr int G_WR1 = G_BR0 + HID;             // [64][LDW]
constexpr int G_BR1 = G_WR1 + HID * LDW;       // [32]
constexpr int G_WH0 = G_BR1 + XD;              // [5][64]
constexpr int G_BH0 = G_WH0 + HD * HID;        // [64]
constexpr int G_WH1 = G_BH0 + HID;             // [64][LDW]
constexpr int G_BH1 = G_WH1 + HID * LDW;       // [32]
constexpr int G_WA  = G_BH1 + XD;              // [32][LDW]
constexpr int G_WS  = G_WA + XD * LDW;         // num_layer x [32][LDW]
__host__ __device__ constexpr int graph_floats(int L) { return G_WS + L * XD * LDW; }

// ---- packed value blob: mlp(32,[32,100,100,1]); hidden 100 padded to 128 columns with zeros -----
constexpr int VH  = RGL_VALUE_HIDDEN;   // 100
constexpr int VHP = 128;
constexpr int V_W0 = 0;                        // [32][32]
constexpr int V_B0 = V_W0 + XD * XD;           // [32]
constexpr int V_W1 = V_B0 + XD;                // [32][128]
constexpr int V_B1 = V_W1 + XD * VHP;          // [128]
constexpr int V_W2 = V_B1 + VHP;               // [100][128]
constexpr int V_B2 = V_W2 + VH * VHP;          // [128]
constexpr int V_W3 = V_B2 + VHP;               // [128]
constexpr int V_B3 = V_W3 + VHP;               // [4] (1 used)
constexpr int VALUE_FLOATS = V_B3 + 4;

// ---- packed motion blob: mlp(32,[64,5]) ------------------------------------------------------------
constexpr int MH = RGL_MOTION_HIDDEN;   // 64
constexpr int M_W0 = 0;                        // [32][64] k-major
constexpr int M_B0 = M_W0 + XD * MH;           // [64]
constexpr int M_W1 = M_B0 + MH;                // [5][64]  (PyTorch layout: one row per output)
constexpr int M_B1 = M_W1 + HD * MH;           // [8] (5 used)
constexpr int MOTION_FLOATS = M_B1 + 8;

// ---- tensor-core (tcgen05) operand sections ------------------------------------------------------------
// Appended to the packed graph / motion blobs (rgl_pack_* writes both the FMA and the tensor-core sections).
// Every weight matrix is a B operand of tcgen05.mma kind::tf32: [n rows][32 k] K-major tiles in the UMMA
// SWIZZLE_128B canonical layout (row = 128 B, 16-byte chunk c of row r stored at chunk c ^ (r & 7)), split into a
// tf32 "hi" tile and a tf32 "lo" tile (w = hi + lo to ~2^-22) for the 3xTF32 product.  Offsets in floats.
constexpr int T_W0   = 0;                       // emb layer 1, robot and human weights concatenated along k, ONE [64 hidden][32] tile:
                                                //   k 0-15 = hi of [0-8 robot feats, 9-13 human feats, 14 robot bias, 15 human bias],
                                                //   k 16-31 = lo of the same 16 columns (the lo descriptor starts 64 B into the rows)
constexpr int T_W1   = T_W0 + 2048;             // emb layer 2, stacked along n: rows 0-31 = w_h.2.weight, rows 32-63 = w_r.2.weight;
                                                //   two k atoms: hi atoms at +0, +2048; lo atoms at +4096, +6144
constexpr int T_WA   = T_W1 + 8192;             // layer-0 tile, stacked along n: rows 0-31 = w_a^T, rows 32-63 = Ws[0]^T (one N=64 MMA gives
                                                //   Y = X w_a and X Ws[0] from a single A operand): hi [64][32], lo at +2048
constexpr int T_WS1  = T_WA + 4096;             // layers l >= 1: Ws[l]^T hi [32][32] at T_WS1 + (l-1)*2048, lo at +1024
__host__ __device__ constexpr int tc_bias_off(int L) { return T_WS1 + (L - 1) * 2048; }   // [0,32) = w_h.2.bias, [TC_RBIAS, +32) = w_r.2.bias
constexpr int TC_RBIAS = 36;                    // robot bias 144 B after the human bias (4 banks apart: a quarter-warp that holds robot
                                                // AND human rows reads both without a bank conflict; at +32 floats the two alias)
__host__ __device__ constexpr int tc_graph_floats(int L) { return tc_bias_off(L) + 256; }
__host__ __device__ constexpr int graph_tc_off(int L) { return (graph_floats(L) + 63) & ~63; }     // 256 B aligned
__host__ __device__ constexpr int graph_floats_total(int L) { return graph_tc_off(L) + tc_graph_floats(L); }
constexpr int TM_W0 = 0;                        // motion layer 1 (0.weight [64,32]): hi [64][32], lo at +2048
constexpr int TM_B0 = 4096;                     // [64]
constexpr int TM_W1 = TM_B0 + 64;               // 2.weight [5][64] (fp32, FMA pipe)
constexpr int TM_B1 = TM_W1 + 320;              // [8] (5 used)
constexpr int TMOTION_FLOATS = 4608;            // 18 KB
constexpr int MOTION_TC_OFF = (MOTION_FLOATS + 63) & ~63;
constexpr int MOTION_FLOATS_TOTAL = MOTION_TC_OFF + TMOTION_FLOATS;
// value network mlp(32,[32,100,100,1]) for value_head_tc.cu: 100-wide layers padded to N = 112 (UMMA N % 16 == 0), the
// K = 100 contraction of layer 2 split into four 32-wide k atoms (the last one carries 8 values: one k-step)
constexpr int TV_NP   = 112;
constexpr int TV_W0   = 0;                      // 0.weight [32][32]: hi, lo at +1024
constexpr int TV_BIAS = 2048;                   // b0 [32] | b1 [128] at +32 | b2 [128] at +160 | 6.weight [128] at +288 | 6.bias [4] at +416
constexpr int TV_W1   = TV_BIAS + 512;          // 2.weight [112][32] (rows 100-111 zero): hi, lo at +3584
constexpr int TV_W2   = TV_W1 + 2 * TV_NP * 32; // 4.weight: hi atoms [112][32] at +a*3584 (k = 32a .. 32a+31), lo atoms at +14336
constexpr int TVALUE_FLOATS = TV_W2 + 8 * TV_NP * 32;      // 38 400 floats = 150 KB
constexpr int VALUE_TC_OFF = (VALUE_FLOATS + 63) & ~63;
constexpr int VALUE_FLOATS_TOTAL = VALUE_TC_OFF + TVALUE_FLOATS;
static_assert((TV_W1 * 4) % 1024 == 0 && (TV_W2 * 4) % 1024 == 0 && (TV_NP * 128) % 1024 == 0, "UMMA tiles are 1 KB aligned");
static_assert(TM_B1 + 8 <= TMOTION_FLOATS && tc_graph_floats(1) % 256 == 0, "tensor-core blob sections are 1 KB multiples");

static_assert(graph_floats(2) - (2 * HID + 3 * XD) * (LDW - XD) == 8256, "graph parameter count (SURVEY.md 2b) + row padding");
static_assert(graph_floats(RGL_MAX_LAYERS) % 4 == 0 && VALUE_FLOATS % 4 == 0 && MOTION_FLOATS % 4 == 0, "16B sections");

// ---- programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// while its predecessor on the stream is still running; it must execute pdl_wait() before it reads anything the predecessor
// wrote (or writes anything the predecessor reads).  pdl_trigger() in the predecessor lets the dependent's launch latency and
// prologue overlap the predecessor's tail.  Both are no-ops for kernels launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- PTX: mbarrier + TMA bulk copy (cp.async.bulk -> SASS UBLKCP) ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// same wait with a suspend-time hint (ns): the thread sleeps in hardware until the phase completes (or the hint expires)
// instead of re-polling every few hundred cycles -- the MMA waits of the tcgen05 kernels are 1-2 us long
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion counted on an mbarrier (bytes % 16 == 0, both 16 B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float4 lds128(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void   sts128(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---- warp-level register-tiled GEMM micro-kernel -------------------------------------------------
// A warp owns a row block of RB = 8*RT rows.  lane = cg*8 + rg; thread rows = r0 + rg + 8q (q < RT),
// thread columns = cg*4 + 16m + {0..3} (m < NC4), i.e. the warp covers 16*NC4 output columns.
//   acc[q][4m+j] += sum_k x[row_q][k] * W[k][col]           x row-major in smem (stride ldx), K % 4 == 0
// Shared-memory traffic per 4 k: RT LDS.128 (8 distinct rows x 16 B, conflict-free for ldx = 4 mod 32)
// + 4*NC4 LDS.128 (4 distinct 16 B chunks, broadcast) for 16*RT*NC4 FFMA.
template <int RT, int NC4>
__device__ __forceinline__ void tile_gemm(float (&acc)[RT][NC4 * 4], const float* __restrict__ xrow0, int ldx,
                                          const float* __restrict__ w, int ldw, int K) {
    // xrow0 = &x[(r0 + rg) * ldx], w = &W[0][cg*4]
#pragma unroll 2
    for (int k4 = 0; k4 < K; k4 += 4) {
        float4 xv[RT];
#pragma unroll
        for (int q = 0; q < RT; ++q) xv[q] = lds128(xrow0 + q * 8 * ldx + k4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float4 wv[NC4];
#pragma unroll
            for (int m = 0; m < NC4; ++m) wv[m] = lds128(w + (k4 + kk) * ldw + 16 * m);
#pragma unroll
            for (int q = 0; q < RT; ++q) {
                const float xs = kk == 0 ? xv[q].x : kk == 1 ? xv[q].y : kk == 2 ? xv[q].z : xv[q].w;
#pragma unroll
                for (int m = 0; m < NC4; ++m) {
                    acc[q][4 * m + 0] = fmaf(xs, wv[m].x, acc[q][4 * m + 0]);
                    acc[q][4 * m + 1] = fmaf(xs, wv[m].y, acc[q][4 * m + 1]);
                    acc[q][4 * m + 2] = fmaf(xs, wv[m].z, acc[q][4 * m + 2]);
                    acc[q][4 * m + 3] = fmaf(xs, wv[m].w, acc[q][4 * m + 3]);
                }
            }
        }
    }
}


// Software-pipelined variant (K compile time, K/4 even): the operands of k-step s+1 are requested before the
// FFMAs of step s are issued, so one warp covers its own LDS latency.  The loop carries two k-steps per
// iteration (static double buffer); for the 32-row tiles (RT = 4) it is NOT unrolled further: a fully unrolled
// K=32 tile is 18 KB of SASS and five of them thrash the instruction cache (ncu: stall_no_inst 19%, icc hit 80%).
template <int RT, int NC4>
__device__ __forceinline__ void tile_load(float4 (&xv)[RT], float4 (&wv)[4][NC4], const float* __restrict__ xrow0, int ldx,
                                          const float* __restrict__ w, int ldw, int k4) {
#pragma unroll
    for (int q = 0; q < RT; ++q) xv[q] = lds128(xrow0 + q * 8 * ldx + k4);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int m = 0; m < NC4; ++m) wv[kk][m] = lds128(w + (k4 + kk) * ldw + 16 * m);
}
template <int RT, int NC4>
__device__ __forceinline__ void tile_fma(float (&acc)[RT][NC4 * 4], const float4 (&xv)[RT], const float4 (&wv)[4][NC4]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int q = 0; q < RT; ++q) {
            const float xs = kk == 0 ? xv[q].x : kk == 1 ? xv[q].y : kk == 2 ? xv[q].z : xv[q].w;
#pragma unroll
            for (int m = 0; m < NC4; ++m) {
                acc[q][4 * m + 0] = fmaf(xs, wv[kk][m].x, acc[q][4 * m + 0]);
                acc[q][4 * m + 1] = fmaf(xs, wv[kk][m].y, acc[q][4 * m + 1]);
                acc[q][4 * m + 2] = fmaf(xs, wv[kk][m].z, acc[q][4 * m + 2]);
                acc[q][4 * m + 3] = fmaf(xs, wv[kk][m].w, acc[q][4 * m + 3]);
            }
        }
    }
}
template <int RT, int NC4, int K>
__device__ __forceinline__ void tile_gemm_pf(float (&acc)[RT][NC4 * 4], const float* __restrict__ xrow0, int ldx,
                                             const float* __restrict__ w, int ldw) {
    static_assert(K % 8 == 0, "K/4 must be even");
    float4 xa[RT], xb[RT];
    float4 wa[4][NC4], wb[4][NC4];
    tile_load<RT, NC4>(xa, wa, xrow0, ldx, w, ldw, 0);
    if constexpr (RT <= 2) {
        // small tiles: straight-line code (measured faster: 388 vs 361 M states/s; 9 KB per instance)
#pragma unroll
        for (int k4 = 0; k4 < K; k4 += 8) {
            tile_load<RT, NC4>(xb, wb, xrow0, ldx, w, ldw, k4 + 4);
            tile_fma<RT, NC4>(acc, xa, wa);
            if (k4 + 8 < K) tile_load<RT, NC4>(xa, wa, xrow0, ldx, w, ldw, k4 + 8);
            tile_fma<RT, NC4>(acc, xb, wb);
        }
    } else {
#pragma unroll 1
        for (int k4 = 0; k4 < K; k4 += 8) {
            tile_load<RT, NC4>(xb, wb, xrow0, ldx, w, ldw, k4 + 4);
            tile_fma<RT, NC4>(acc, xa, wa);
            if (k4 + 8 < K) tile_load<RT, NC4>(xa, wa, xrow0, ldx, w, ldw, k4 + 8);
            tile_fma<RT, NC4>(acc, xb, wb);
        }
    }
}


// ---- tensor-core path: legacy warp-level mma.sync (SASS HMMA) with a 3xTF32 split --------------------------------
// fp32 operands are split x = hi + lo with hi = x rounded to a tf32 mantissa and lo = x - hi (exact in fp32), itself
// rounded to tf32; D += lo*Bhi + hi*Blo + hi*Bhi recovers fp32-level products
// (measured 4e-7 relative error on a K=32 dot of O(20) activations, tools/mma_peak.cu) with fp32 accumulation.
// Measured mma.sync m16n8k8 TF32 rate on this B200: 478 MAC/clk/SM (tcgen05 is 4x that; round-2 target).
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    // round-to-nearest on both parts (integer add of half a tf32 ulp before the low 13 bits are dropped): truncation
    // would bias every product the same way and the error would grow linearly in K (measured 2.5e-6 vs 3e-7 relative)
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;       // the tensor core ignores the low 13 bits
}

// Warp GEMM on tensor cores: acc[mt][nt][.] += X[16*MT rows][8*KS] * W[8*KS][32].
//   x: first of the warp's 16*MT rows (row-major, stride LDX); w: k-major weights, stride LDW (= 8 mod 32: conflict-free
//   B-fragment loads); accumulator fragment (m16n8): [0],[1] = row g, cols 2t,2t+1; [2],[3] = row g+8 (g = lane>>2, t = lane&3)
template <int KS, int MT>
__device__ __forceinline__ void mma_gemm_3xtf32(float (&acc)[MT][4][4], const float* __restrict__ x, const float* __restrict__ w, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        uint32_t ahi[MT][4], alo[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const float* xr = x + (mt * 16 + g) * LDX + ks * 8 + t;
            split_tf32(xr[0], ahi[mt][0], alo[mt][0]);
            split_tf32(xr[8 * LDX], ahi[mt][1], alo[mt][1]);
            split_tf32(xr[4], ahi[mt][2], alo[mt][2]);
            split_tf32(xr[8 * LDX + 4], ahi[mt][3], alo[mt][3]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32(w[(ks * 8 + t) * LDW + nt * 8 + g], bh0, bl0);
            split_tf32(w[(ks * 8 + t + 4) * LDW + nt * 8 + g], bh1, bl1);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                mma_tf32(acc[mt][nt], alo[mt], bh0, bh1);       // small terms first
                mma_tf32(acc[mt][nt], ahi[mt], bl0, bl1);
                mma_tf32(acc[mt][nt], ahi[mt], bh0, bh1);
            }
        }
    }
}
// accumulator-fragment helpers (base = first of the warp's 32 rows, stride LDX)
template <int MT>
__device__ __forceinline__ void cfrag_fill(float (&acc)[MT][4][4], const float* bias, int lane) {
    const int t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const float b0 = bias ? bias[nt * 8 + 2 * t] : 0.f, b1 = bias ? bias[nt * 8 + 2 * t + 1] : 0.f;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) { acc[mt][nt][0] = b0; acc[mt][nt][1] = b1; acc[mt][nt][2] = b0; acc[mt][nt][3] = b1; }
    }
}
template <bool RELU, int MT>
__device__ __forceinline__ void cfrag_store(float* base, const float (&acc)[MT][4][4], int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            float2 lo = make_float2(acc[mt][nt][0], acc[mt][nt][1]), hi = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
            if (RELU) { lo.x = fmaxf(lo.x, 0.f); lo.y = fmaxf(lo.y, 0.f); hi.x = fmaxf(hi.x, 0.f); hi.y = fmaxf(hi.y, 0.f); }
            *reinterpret_cast<float2*>(base + (mt * 16 + g) * LDX + nt * 8 + 2 * t) = lo;
            *reinterpret_cast<float2*>(base + (mt * 16 + g + 8) * LDX + nt * 8 + 2 * t) = hi;
        }
}

// Same tile, x read with scalar loads (tiny K not a multiple of 4: the raw 9- / 5-float states).
template <int RT, int NC4>
__device__ __forceinline__ void tile_gemm_smallk(float (&acc)[RT][NC4 * 4], const float* (&xrow)[RT],
                                                 const float* __restrict__ w, int ldw, int K) {
    for (int k = 0; k < K; ++k) {
        float4 wv[NC4];
#pragma unroll
        for (int m = 0; m < NC4; ++m) wv[m] = lds128(w + k * ldw + 16 * m);
#pragma unroll
        for (int q = 0; q < RT; ++q) {
            const float xs = xrow[q][k];
#pragma unroll
            for (int m = 0; m < NC4; ++m) {
                acc[q][4 * m + 0] = fmaf(xs, wv[m].x, acc[q][4 * m + 0]);
                acc[q][4 * m + 1] = fmaf(xs, wv[m].y, acc[q][4 * m + 1]);
                acc[q][4 * m + 2] = fmaf(xs, wv[m].z, acc[q][4 * m + 2]);
                acc[q][4 * m + 3] = fmaf(xs, wv[m].w, acc[q][4 * m + 3]);
            }
        }
    }
}

}  // namespace rgl
