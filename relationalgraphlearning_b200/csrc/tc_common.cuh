// tcgen05 / TMEM building blocks shared by the tensor-core kernels (graph_forward_tc.cu, value_head_tc.cu):
// PTX wrappers, UMMA descriptors, TMEM load / store, the tf32 hi/lo operand split and the 3xTF32 MMA issue loop.
// Descriptor formats follow cute::UMMA::SmemDescriptor / InstrDescriptor (CUTLASS 3.8, cute/arch/mma_sm100_desc.hpp);
// every encoding used here is validated on hardware by tools/umma_probe.cu (profiles/r1_umma_probe.md).
#pragma once
#include "common.cuh"

namespace rgl {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32, issued by one thread for the whole CTA
__device__ __forceinline__ void umma_tf32_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor, sm_100 version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, tf32 x tf32, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#define RGL_R8(r, o)  "=r"(r[o]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
#define RGL_I8(r, o)  "r"(r[o]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7])
// this thread's TMEM lane, 32 consecutive columns -> registers (load + wait in one asm block: results are defined after it)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : RGL_R8(r, 0), RGL_R8(r, 8), RGL_R8(r, 16), RGL_R8(r, 24)
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[32], uint32_t (&q)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%64];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%65];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : RGL_R8(r, 0), RGL_R8(r, 8), RGL_R8(r, 16), RGL_R8(r, 24), RGL_R8(q, 0), RGL_R8(q, 8), RGL_R8(q, 16), RGL_R8(q, 24)
        : "r"(taddr), "r"(taddr + 32) : "memory");
}
// 16 / 8 columns (the column-sliced kernels: graph_forward_tq.cu)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : RGL_R8(r, 0), RGL_R8(r, 8) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16x2(uint32_t t0, uint32_t t1, uint32_t (&r)[16], uint32_t (&q)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%32];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%33];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : RGL_R8(r, 0), RGL_R8(r, 8), RGL_R8(q, 0), RGL_R8(q, 8) : "r"(t0), "r"(t1) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : RGL_R8(r, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" :: "r"(taddr), RGL_I8(r, 0) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), RGL_I8(r, 0), RGL_I8(r, 8) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 = two IEEE fp32 FMAs in one instruction, same results as two fmaf) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f32x2 pack2u(uint32_t a, uint32_t b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ void fma2(f32x2& d, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
// 16 bytes of shared memory as two fp32 pairs
__device__ __forceinline__ void lds128s2(uint32_t saddr, f32x2& lo, f32x2& hi) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(saddr));
}

// A-operand row: 16 / 32 values -> tf32 hi / lo columns of this thread's TMEM lane.  hi = x rounded to nearest tf32
// (integer add of half an ulp, then mask: cvt.rna.tf32 is emulated with 3 instructions on sm_100), lo = x - hi exactly;
// the tensor core drops the low 13 bits of lo (2^-23 relative to x).
template <int NV>
__device__ __forceinline__ void st_split(uint32_t t_hi, uint32_t t_lo, const float (&v)[NV]) {
#pragma unroll
    for (int b = 0; b < NV; b += 16) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#ifdef RGL_SPLIT_TRUNC      // experiment only: 2 instead of 3 instructions per element, +3 % throughput, 1.7x the error (H 5.2e-5 vs 3.1e-5 at scale 20)
            hi[j] = __float_as_uint(v[b + j]);                                             // the tensor core drops the low 13 bits
            lo[j] = __float_as_uint(v[b + j] - __uint_as_float(hi[j] & 0xffffe000u));
#else
            hi[j] = (__float_as_uint(v[b + j]) + 0x1000u) & 0xffffe000u;
            if (j & 1) {                                   // lo = x - hi for two elements with one packed subtract
                float l0, l1;
                unpack2(sub2(pack2(v[b + j - 1], v[b + j]), pack2u(hi[j - 1], hi[j])), l0, l1);
                lo[j - 1] = __float_as_uint(l0);
                lo[j] = __float_as_uint(l1);
            }
#endif
        }
        tmem_st16(t_hi + b, hi);
        tmem_st16(t_lo + b, lo);
    }
}

// 8-value variant (K = 16 operand shared by two threads)
__device__ __forceinline__ void st_split8(uint32_t t_hi, uint32_t t_lo, const float (&v)[8]) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        hi[j] = (__float_as_uint(v[j]) + 0x1000u) & 0xffffe000u;
        if (j & 1) {
            float l0, l1;
            unpack2(sub2(pack2(v[j - 1], v[j]), pack2u(hi[j - 1], hi[j])), l0, l1);
            lo[j - 1] = __float_as_uint(l0);
            lo[j] = __float_as_uint(l1);
        }
    }
    tmem_st8(t_hi, hi);
    tmem_st8(t_lo, lo);
}

// one lane of a converged warp, chosen by the hardware (elect.sync): the compiler knows a single thread runs the guarded code,
// so the uniform-datapath tcgen05.mma is predicated directly instead of being wrapped in a divergence loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// one elected thread: D[128 x N] (+)= A * W^T over KS k-steps of 8, 3xTF32 (small terms first).  Every operand is
// warp-uniform (derived from __shfl_sync(.., 0) values), so each MMA is one UTCHMMA with uniform-register operands.
template <int KS>
__device__ __forceinline__ void issue_gemm(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi, uint32_t w_lo, uint32_t idesc, uint32_t acc) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const uint64_t bh = umma_desc(w_hi + ks * 32), bl = umma_desc(w_lo + ks * 32);
        umma_tf32_ts(d, a_lo + ks * 8, bh, idesc, acc);
        umma_tf32_ts(d, a_hi + ks * 8, bl, idesc, 1);
        umma_tf32_ts(d, a_hi + ks * 8, bh, idesc, 1);
        acc = 1;
    }
}

// shared memory through 32-bit addresses (one LOP3 per swizzled access instead of 64-bit generic pointer arithmetic)
__device__ __forceinline__ float4 lds128s(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts128s(uint32_t saddr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
}  // namespace rgl
