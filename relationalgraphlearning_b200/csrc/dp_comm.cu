// Data-parallel gradient exchange for the value-net / state-predictor training step
// (crowd_nav/utils/trainer.py:122-131,143-149 in data-parallel form; SURVEY.md 8(e): ONE sum over a flat fp32 buffer
// of 22 813 gradients = 91 252 B per value step).
//
// The message is latency-bound on NVLink 5 / NVSwitch, so instead of a ring / tree library collective the step is ONE
// kernel over peer memory (one process per GPU, buffers shared with cudaIpc*):
//
//   every CTA owns a contiguous slice of the flat buffer and, for that slice,
//     1. PUSHES this rank's partial gradients into a receive slot in every peer's memory (plain remote stores:
//        fire-and-forget over NVLink, no read round trip),
//     2. publishes a per-(slice, rank) flag in every peer with st.release.sys and waits for the peers' flags
//        (ld.acquire.sys on LOCAL memory),
//     3. sums the `world` contributions in RANK ORDER (identical bits on every rank: the replicas never drift),
//        scales, writes the reduced gradient where the optimizer reads it and re-zeroes the accumulation slice for
//        the next backward.
//
// Receive slots and flags are double-buffered on the parity of a device-resident epoch counter, which makes one flag
// exchange per step sufficient: a peer can only start pushing epoch e+2 into the slot of epoch e after it has seen this
// rank's flag of epoch e+1, which this rank writes after it finished reading epoch e.  Everything the kernel needs
// between steps lives in device memory (no host-side argument changes), so the launch is CUDA-graph capturable.
// A wait that exceeds ~10 s (a peer died) sets a status word and returns instead of hanging the GPU.
#include <stdio.h>
#include <string.h>
#include "kernels.h"

namespace rgl {

constexpr int COMM_MAX_WORLD = 16;
constexpr int COMM_MAX_CTAS = 32;
constexpr int COMM_THREADS = 256;

struct CommDev {
    float* accum;                          // [nP]   local accumulation buffer (the backward kernels atomicAdd into it)
    float* recv;                           // [2][world][nP] local receive slots
    unsigned* flags;                       // [2][COMM_MAX_CTAS][COMM_MAX_WORLD] local flags
    unsigned* epoch;                       // [COMM_MAX_CTAS]
    unsigned* status;                      // [1]  0 = ok, 1 = a wait timed out
    float* peer_recv[COMM_MAX_WORLD];
    unsigned* peer_flags[COMM_MAX_WORLD];
    int rank, world;
    long long nP;                          // floats, multiple of 4
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(COMM_THREADS) grad_allreduce_push_kernel(const CommDev c, float* __restrict__ out, const float scale,
                                                                           const long long n4) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const long long per = (n4 + gridDim.x - 1) / gridDim.x;
    const long long lo = b * per, hi = (lo + per < n4) ? lo + per : n4;
    const unsigned e = c.epoch[b] + 1u;                 // every thread reads it; thread 0 stores the new value at the end
    const unsigned par = e & 1u;
    const float4* acc4 = reinterpret_cast<const float4*>(c.accum);
    const long long slot4 = c.nP >> 2;

    // 1. push this rank's slice into every peer's receive slot [par][rank]
    for (int p = 0; p < c.world; ++p) {
        if (p == c.rank) continue;
        float4* dst = reinterpret_cast<float4*>(c.peer_recv[p]) + ((long long)par * c.world + c.rank) * slot4;
        for (long long i = lo + tid; i < hi; i += COMM_THREADS) dst[i] = acc4[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. flag exchange: thread p talks to peer p
    if (tid < c.world && tid != c.rank) {
        st_release_sys(c.peer_flags[tid] + ((size_t)par * COMM_MAX_CTAS + b) * COMM_MAX_WORLD + c.rank, e);
        const unsigned* mine = c.flags + ((size_t)par * COMM_MAX_CTAS + b) * COMM_MAX_WORLD + tid;
        const unsigned long long t0 = global_ns();
        while ((int)(ld_acquire_sys(mine) - e) < 0) {
            if (global_ns() - t0 > 10000000000ull) { atomicExch(c.status, 1u); break; }
            __nanosleep(100);
        }
    }
    __syncthreads();
    // 3. rank-ordered sum of the slice; re-zero the accumulation slice for the next backward
    const float4* rcv = reinterpret_cast<const float4*>(c.recv) + (long long)par * c.world * slot4;
    float4* out4 = reinterpret_cast<float4*>(out);
    float4* accw = reinterpret_cast<float4*>(c.accum);
    for (long long i = lo + tid; i < hi; i += COMM_THREADS) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = 0; p < c.world; ++p) {
            const float4 v = (p == c.rank) ? acc4[i] : __ldcv(rcv + (long long)p * slot4 + i);     // remote-written: bypass L1
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        out4[i] = make_float4(s.x * scale, s.y * scale, s.z * scale, s.w * scale);
        accw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid == 0) c.epoch[b] = e;
}

}  // namespace rgl

// ------------------------------------------------------------------------------------------------------------------ C ABI
struct RglComm {
    rgl::CommDev d;
    void* base;                 // one cudaMalloc: [accum | recv | flags | epoch | status]
    size_t bytes;
    size_t off_recv, off_flags;
    void* peer_base[rgl::COMM_MAX_WORLD];
    long long n;                // floats in use
    int ctas;
    int dev;
};

namespace {
thread_local char g_comm_err[256] = "";
int cfail(int code, const char* msg, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess) snprintf(g_comm_err, sizeof(g_comm_err), "%s: %s", msg, cudaGetErrorString(e));
    else snprintf(g_comm_err, sizeof(g_comm_err), "%s", msg);
    return code;
}
}  // namespace

extern "C" {

const char* rgl_comm_last_error_string(void) { return g_comm_err; }

int rgl_comm_create(int rank, int world, long long nfloats, RglComm** out) {
    if (!out || world < 1 || world > rgl::COMM_MAX_WORLD || rank < 0 || rank >= world || nfloats < 1)
        return cfail(RGL_EINVAL, "rgl_comm_create: bad argument");
    RglComm* c = new RglComm();
    memset(c, 0, sizeof(*c));
    c->n = nfloats;
    const long long nP = (nfloats + 3) & ~3LL;
    const size_t accum_b = ((size_t)nP * 4 + 255) & ~(size_t)255;
    const size_t recv_b = ((size_t)2 * world * nP * 4 + 255) & ~(size_t)255;
    const size_t flags_b = (size_t)2 * rgl::COMM_MAX_CTAS * rgl::COMM_MAX_WORLD * 4;
    c->off_recv = accum_b;
    c->off_flags = accum_b + recv_b;
    c->bytes = c->off_flags + flags_b + rgl::COMM_MAX_CTAS * 4 + 256;
    cudaError_t e = cudaGetDevice(&c->dev);
    if (e == cudaSuccess) e = cudaMalloc(&c->base, c->bytes);
    if (e == cudaSuccess) e = cudaMemset(c->base, 0, c->bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { delete c; return cfail(RGL_ECUDA, "rgl_comm_create", e); }
    char* b = static_cast<char*>(c->base);
    c->d.accum = reinterpret_cast<float*>(b);
    c->d.recv = reinterpret_cast<float*>(b + c->off_recv);
    c->d.flags = reinterpret_cast<unsigned*>(b + c->off_flags);
    c->d.epoch = reinterpret_cast<unsigned*>(b + c->off_flags + flags_b);
    c->d.status = c->d.epoch + rgl::COMM_MAX_CTAS;
    c->d.rank = rank; c->d.world = world; c->d.nP = nP;
    c->d.peer_recv[rank] = c->d.recv;
    c->d.peer_flags[rank] = c->d.flags;
    // slices of >= 4 KB: enough CTAs to hide the NVLink store latency, few enough that all are co-resident
    int ctas = (int)((nP * 4 + 4095) / 4096);
    c->ctas = ctas < 1 ? 1 : (ctas > rgl::COMM_MAX_CTAS ? rgl::COMM_MAX_CTAS : ctas);
    *out = c;
    return RGL_OK;
}

int rgl_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int rgl_comm_ipc_handle(RglComm* c, void* handle_out) {
    if (!c || !handle_out) return cfail(RGL_EINVAL, "rgl_comm_ipc_handle: null argument");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, c->base);
    if (e != cudaSuccess) return cfail(RGL_ECUDA, "cudaIpcGetMemHandle", e);
    memcpy(handle_out, &h, sizeof(h));
    return RGL_OK;
}

int rgl_comm_open_peers(RglComm* c, const void* handles) {
    if (!c || !handles) return cfail(RGL_EINVAL, "rgl_comm_open_peers: null argument");
    const char* hb = static_cast<const char*>(handles);
    for (int p = 0; p < c->d.world; ++p) {
        if (p == c->d.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hb + (size_t)p * sizeof(h), sizeof(h));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return cfail(RGL_ECUDA, "cudaIpcOpenMemHandle (peer memory unavailable)", e);
        c->peer_base[p] = ptr;
        c->d.peer_recv[p] = reinterpret_cast<float*>(static_cast<char*>(ptr) + c->off_recv);
        c->d.peer_flags[p] = reinterpret_cast<unsigned*>(static_cast<char*>(ptr) + c->off_flags);
    }
    return RGL_OK;
}

float* rgl_comm_accum_ptr(RglComm* c) { return c ? c->d.accum : nullptr; }

int rgl_comm_allreduce(RglComm* c, float* out, float scale, rgl_stream_t stream) {
    if (!c || !out) return cfail(RGL_EINVAL, "rgl_comm_allreduce: null argument");
    if (reinterpret_cast<uintptr_t>(out) & 15u) return cfail(RGL_EALIGN, "rgl_comm_allreduce: out must be 16-byte aligned");
    for (int p = 0; p < c->d.world; ++p)
        if (!c->d.peer_recv[p]) return cfail(RGL_EINVAL, "rgl_comm_allreduce: peers not opened");
    rgl::grad_allreduce_push_kernel<<<c->ctas, rgl::COMM_THREADS, 0, (cudaStream_t)stream>>>(c->d, out, scale, c->d.nP >> 2);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? RGL_OK : cfail(RGL_ECUDA, "rgl_comm_allreduce", e);
}

int rgl_comm_status(RglComm* c, int* status) {
    if (!c || !status) return cfail(RGL_EINVAL, "rgl_comm_status: null argument");
    unsigned v = 0;
    cudaError_t e = cudaMemcpy(&v, c->d.status, sizeof(v), cudaMemcpyDeviceToHost);      // synchronises: call outside hot loops
    if (e != cudaSuccess) return cfail(RGL_ECUDA, "rgl_comm_status", e);
    *status = (int)v;
    return RGL_OK;
}

int rgl_comm_destroy(RglComm* c) {
    if (!c) return RGL_OK;
    cudaDeviceSynchronize();
    for (int p = 0; p < c->d.world; ++p)
        if (p != c->d.rank && c->peer_base[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
    if (c->base) cudaFree(c->base);
    delete c;
    return RGL_OK;
}

}  // extern "C"
