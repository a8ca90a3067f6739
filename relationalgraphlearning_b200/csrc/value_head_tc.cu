// Value network V = mlp(32,[32,100,100,1])(E) on the 5th-generation tensor cores (tcgen05 / TMEM) of sm_100a.
//
// Replaces crowd_nav/policy/value_estimator.py:19 (value_network = mlp(...), crowd_nav/policy/helpers.py:5-13: ReLU after
// every layer but the last) for large batches; the fp32-FMA kernel of value_head.cu keeps the small batches and the
// training forward.
//
//   group  = 128 threads = one UMMA M-tile of 128 states; a thread owns one state end to end (TMEM lane = state).
//   GEMMs  = tcgen05.mma kind::tf32, M = 128, A operand in TMEM (3xTF32 split, tc_common.cuh), B = weight tiles resident
//            in shared memory (150 KB per CTA: the whole network, staged once by three cp.async.bulk copies).
//            layer 0: K = 32, N = 32.   layer 1: K = 32, N = 112 (100 padded).   layer 2: K = 100 as four k atoms
//            (32, 32, 32, 8), N = 112, the A operand double-buffered so that the split of atom q+1 overlaps the MMAs of
//            atom q.   layer 3 (100 -> 1): per-thread dot product on the FMA pipe.
//   TMEM   = 256 columns per group: [0,112) accumulator, [128,192) / [192,256) A hi|lo double buffer  ->  two groups per SM.
#include <stdlib.h>
#include "kernels.h"
#include "tc_common.cuh"

namespace rgl {


constexpr int VT_COLS = 256;
constexpr int VC_D = 0, VC_A0 = 128, VC_A1 = 192;      // A buffers: hi at +0, lo at +32

// SAVE: the training forward -- additionally writes the post-relu activations the backward reads
// (sv0 [B,32], sv1 / sv2 [B,128] with 100 valid columns; one thread = one row, 16-byte stores).
template <int G, bool SAVE>
__global__ void __launch_bounds__(128 * G, 1) value_head_tc_kernel(const float* __restrict__ E, int B, const float* __restrict__ vw,
                                                                    float* __restrict__ V, int ntiles, float* __restrict__ sv0,
                                                                    float* __restrict__ sv1, float* __restrict__ sv2) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* tw = reinterpret_cast<float*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    uint64_t* bars = reinterpret_cast<uint64_t*>(tw + TVALUE_FLOATS);      // [0..2] weight stages, [3+2g], [4+2g] group g
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 3 + 2 * G);

    const int tid = threadIdx.x, lane = tid & 31, gt = tid & 127;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int grp = warp >> 2, wq = warp & 3;

    if (warp == 0) tmem_alloc(tslot, VT_COLS * G);
    if (tid == 0) {
        for (int i = 0; i < 3 + 2 * G; ++i) mbar_init(bars + i, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();      // PDL: the prologue above may overlap the tail of the kernel that produces E (common.cuh)
    if (tid == 0) {
        const float* src = vw + VALUE_TC_OFF;
        mbar_arrive_expect_tx(bars + 0, TV_W1 * 4u);                                  // layer 0 + biases
        bulk_g2s(tw, src, TV_W1 * 4u, bars + 0);
        mbar_arrive_expect_tx(bars + 1, (TV_W2 - TV_W1) * 4u);                        // layer 1
        bulk_g2s(tw + TV_W1, src + TV_W1, (TV_W2 - TV_W1) * 4u, bars + 1);
        mbar_arrive_expect_tx(bars + 2, (TVALUE_FLOATS - TV_W2) * 4u);                // layer 2
        bulk_g2s(tw + TV_W2, src + TV_W2, (TVALUE_FLOATS - TV_W2) * 4u, bars + 2);
    }

    const uint32_t tbase = __shfl_sync(0xffffffffu, *tslot, 0);
    const uint32_t tg = tbase + grp * VT_COLS;
    const uint32_t tl = tg + ((uint32_t)(wq * 32) << 16);
    const uint32_t tw_s = __shfl_sync(0xffffffffu, smem_u32(tw), 0);
    // Two completion barriers per group, one per A buffer: an mbarrier parity wait tolerates a single outstanding phase, and
    // in layer 2 the commits of atoms q and q+1 are both in flight before the first of them is waited for.
    uint64_t* gbar0 = bars + 3 + 2 * grp;
    uint64_t* gbar1 = gbar0 + 1;
    uint32_t par0 = 0, par1 = 0;
    const bool issuer = wq == 0;
    auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" :: "r"(grp + 1) : "memory"); };
    auto mma_wait0 = [&]() { mbar_wait_sleepy(gbar0, par0); par0 ^= 1; tc_fence_after(); };
    auto mma_wait1 = [&]() { mbar_wait_sleepy(gbar1, par1); par1 ^= 1; tc_fence_after(); };
    auto publish = [&]() { tmem_st_wait(); tc_fence_before(); group_sync(); };
    const float* bias = tw + TV_BIAS;
    constexpr uint32_t W2A = TV_NP * 32 * 4;            // bytes per layer-2 k atom
    bool first = true;

    for (int tile = blockIdx.x * G + grp; tile < ntiles; tile += gridDim.x * G) {
        const long s = (long)tile * 128 + gt;
        const bool valid = s < B;

        // ---- layer 0: relu(W0 e + b0), K = 32, N = 32 ----
        {
            float e[32];
            const float4* src = reinterpret_cast<const float4*>(E + s * XD);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 v = valid ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                e[4 * c] = v.x; e[4 * c + 1] = v.y; e[4 * c + 2] = v.z; e[4 * c + 3] = v.w;
            }
            st_split<32>(tl + VC_A0, tl + VC_A0 + 32, e);
        }
        publish();
        if (issuer) {
            if (elect_one()) {
                if (first) mbar_wait(bars + 0, 0);
                tc_fence_after();
                issue_gemm<4>(tg + VC_D, tg + VC_A0, tg + VC_A0 + 32, tw_s + TV_W0 * 4, tw_s + (TV_W0 + 1024) * 4, umma_idesc(128, 32), 0);
                umma_commit(gbar0);
            }
            __syncwarp();
        }
        if (first) mbar_wait(bars + 0, 0);               // biases
        mma_wait0();

        // ---- layer 1: relu(W1 h0 + b1), K = 32, N = 112 ----
        {
            uint32_t d[32];
            tmem_ld32(tl + VC_D, d);
            float h[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 b = lds128(bias + 4 * c);
                h[4 * c] = fmaxf(__uint_as_float(d[4 * c]) + b.x, 0.f); h[4 * c + 1] = fmaxf(__uint_as_float(d[4 * c + 1]) + b.y, 0.f);
                h[4 * c + 2] = fmaxf(__uint_as_float(d[4 * c + 2]) + b.z, 0.f); h[4 * c + 3] = fmaxf(__uint_as_float(d[4 * c + 3]) + b.w, 0.f);
            }
            st_split<32>(tl + VC_A1, tl + VC_A1 + 32, h);
            if (SAVE && valid) {
                float4* o = reinterpret_cast<float4*>(sv0 + s * 32);
#pragma unroll
                for (int c = 0; c < 8; ++c) o[c] = make_float4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
            }
        }
        publish();
        if (issuer) {
            if (elect_one()) {
                if (first) mbar_wait(bars + 1, 0);
                tc_fence_after();
                issue_gemm<4>(tg + VC_D, tg + VC_A1, tg + VC_A1 + 32, tw_s + TV_W1 * 4, tw_s + (TV_W1 + TV_NP * 32) * 4, umma_idesc(128, TV_NP), 0);
                umma_commit(gbar1);
            }
            __syncwarp();
        }
        mma_wait1();

        // ---- layer 2: relu(W2 h1 + b2), K = 100 = 32 + 32 + 32 + 8, N = 112; the whole h1 row leaves TMEM first ----
        float h1[104];
        {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                uint32_t d[32];
                tmem_ld32(tl + VC_D + 32 * q, d);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 b = lds128(bias + 32 + 32 * q + 4 * c);
                    h1[32 * q + 4 * c] = fmaxf(__uint_as_float(d[4 * c]) + b.x, 0.f); h1[32 * q + 4 * c + 1] = fmaxf(__uint_as_float(d[4 * c + 1]) + b.y, 0.f);
                    h1[32 * q + 4 * c + 2] = fmaxf(__uint_as_float(d[4 * c + 2]) + b.z, 0.f); h1[32 * q + 4 * c + 3] = fmaxf(__uint_as_float(d[4 * c + 3]) + b.w, 0.f);
                }
            }
            uint32_t d8[8];
            tmem_ld8(tl + VC_D + 96, d8);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float4 b = lds128(bias + 32 + 96 + 4 * c);
                h1[96 + 4 * c] = fmaxf(__uint_as_float(d8[4 * c]) + b.x, 0.f); h1[96 + 4 * c + 1] = fmaxf(__uint_as_float(d8[4 * c + 1]) + b.y, 0.f);
                h1[96 + 4 * c + 2] = fmaxf(__uint_as_float(d8[4 * c + 2]) + b.z, 0.f); h1[96 + 4 * c + 3] = fmaxf(__uint_as_float(d8[4 * c + 3]) + b.w, 0.f);
            }
        }
        if (SAVE && valid) {
            float4* o = reinterpret_cast<float4*>(sv1 + s * 128);
#pragma unroll
            for (int c = 0; c < 26; ++c) o[c] = make_float4(h1[4 * c], h1[4 * c + 1], h1[4 * c + 2], h1[4 * c + 3]);
        }
        const uint32_t w2 = tw_s + TV_W2 * 4;
        // atom 0 -> A0 (free: layer 0 is done), atom 1 -> A1 (free: layer 1 is done), atom 2 -> A0 after atom 0's MMAs,
        // atom 3 (8 values) -> A1 after atom 1's MMAs.  One commit per atom; the waits are consumed in order.
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t acol = (q & 1) ? VC_A1 : VC_A0;
            if (q == 2) mma_wait0();                     // commit of atom q-2: its A buffer may be overwritten
            if (q == 3) mma_wait1();
            if (q < 3) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = h1[32 * q + j];
                st_split<32>(tl + acol, tl + acol + 32, v);
            } else {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = j < 8 ? h1[96 + j] : 0.f;
                st_split<16>(tl + acol, tl + acol + 32, v);
            }
            publish();
            if (issuer) {
                if (elect_one()) {
                    if (first && q == 0) mbar_wait(bars + 2, 0);
                    tc_fence_after();
                    if (q < 3) issue_gemm<4>(tg + VC_D, tg + acol, tg + acol + 32, w2 + q * W2A, w2 + (4 + q) * W2A, umma_idesc(128, TV_NP), q > 0);
                    else issue_gemm<1>(tg + VC_D, tg + acol, tg + acol + 32, w2 + q * W2A, w2 + (4 + q) * W2A, umma_idesc(128, TV_NP), 1);
                    umma_commit((q & 1) ? gbar1 : gbar0);
                }
                __syncwarp();
            }
        }
        mma_wait0();                                     // atom 2
        mma_wait1();                                     // atom 3: accumulator complete
        first = false;

        // ---- layer 3: V = w3 . relu(d + b2) + b3 (FMA pipe) ----
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            uint32_t d[32];
            tmem_ld32(tl + VC_D + 32 * q, d);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 b = lds128(bias + 160 + 32 * q + 4 * c), w = lds128(bias + 288 + 32 * q + 4 * c);
                const float4 r = make_float4(fmaxf(__uint_as_float(d[4 * c]) + b.x, 0.f), fmaxf(__uint_as_float(d[4 * c + 1]) + b.y, 0.f),
                                             fmaxf(__uint_as_float(d[4 * c + 2]) + b.z, 0.f), fmaxf(__uint_as_float(d[4 * c + 3]) + b.w, 0.f));
                v0 = fmaf(r.x, w.x, v0); v1 = fmaf(r.y, w.y, v1); v2 = fmaf(r.z, w.z, v2); v3 = fmaf(r.w, w.w, v3);
                if (SAVE && valid) reinterpret_cast<float4*>(sv2 + s * 128)[8 * q + c] = r;
            }
        }
        {
            uint32_t d8[8];
            tmem_ld8(tl + VC_D + 96, d8);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float4 b = lds128(bias + 160 + 96 + 4 * c), w = lds128(bias + 288 + 96 + 4 * c);
                const float4 r = make_float4(fmaxf(__uint_as_float(d8[4 * c]) + b.x, 0.f), fmaxf(__uint_as_float(d8[4 * c + 1]) + b.y, 0.f),
                                             fmaxf(__uint_as_float(d8[4 * c + 2]) + b.z, 0.f), fmaxf(__uint_as_float(d8[4 * c + 3]) + b.w, 0.f));
                v0 = fmaf(r.x, w.x, v0); v1 = fmaf(r.y, w.y, v1); v2 = fmaf(r.z, w.z, v2); v3 = fmaf(r.w, w.w, v3);
                if (SAVE && valid) reinterpret_cast<float4*>(sv2 + s * 128)[24 + c] = r;
            }
        }
        if (tile + gridDim.x * G >= ntiles) pdl_trigger();      // last tile of this CTA: the next kernel may start its prologue
        if (valid) V[s] = ((v0 + v1) + (v2 + v3)) + bias[416];
        tc_fence_before();                               // the next tile's MMAs overwrite the accumulator columns just read
    }

    if (tid == 0) { mbar_wait(bars + 0, 0); mbar_wait(bars + 1, 0); mbar_wait(bars + 2, 0); }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, VT_COLS * G);
}

template <int G, bool SAVE>
static cudaError_t launch_vtc(const float* E, int B, const float* vw, float* V, int num_sms, size_t max_smem, cudaStream_t st,
                              float* sv0, float* sv1, float* sv2) {
    const size_t smem = 1024 + (size_t)TVALUE_FLOATS * 4 + (3 + 2 * G) * 8 + 16;
    if (smem > max_smem) return cudaErrorInvalidConfiguration;
    if (cudaError_t e = ensure_dyn_smem(value_head_tc_kernel<G, SAVE>, (int)max_smem)) return e;
    const int ntiles = (B + 127) / 128;
    const int want = (ntiles + G - 1) / G;
    const int grid = want < num_sms ? want : num_sms;
    return launch_pdl(value_head_tc_kernel<G, SAVE>, dim3(grid), dim3(128 * G), smem, st, E, B, vw, V, ntiles, sv0, sv1, sv2);
}

cudaError_t run_value_head_tc(const float* E, int B, const float* vw, float* V, int num_sms, size_t max_smem, cudaStream_t st) {
    static const char* force = getenv("RGL_TC_VALUE_GROUPS");          // experiments only
    const int ntiles = (B + 127) / 128;
    const int g = force ? atoi(force) : (ntiles > num_sms ? 2 : 1);
    return g == 2 ? launch_vtc<2, false>(E, B, vw, V, num_sms, max_smem, st, nullptr, nullptr, nullptr)
                  : launch_vtc<1, false>(E, B, vw, V, num_sms, max_smem, st, nullptr, nullptr, nullptr);
}

// training forward: V plus the activation saves v0 [B,32], v1 / v2 [B,128] (columns >= 100 of v1 / v2: zeros up to 104, the
// rest untouched -- the backward reads 100)
cudaError_t run_value_head_tc_train(const float* E, int B, const float* vw, float* V, float* v0, float* v1, float* v2,
                                    int num_sms, size_t max_smem, cudaStream_t st) {
    const int ntiles = (B + 127) / 128;
    return ntiles > num_sms ? launch_vtc<2, true>(E, B, vw, V, num_sms, max_smem, st, v0, v1, v2)
                            : launch_vtc<1, true>(E, B, vw, V, num_sms, max_smem, st, v0, v1, v2);
}

}  // namespace rgl
