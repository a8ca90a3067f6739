// Phase trace of graph_forward_tc_kernel (built with -DRGL_TC_TRACE): average cycles per tile spent in each wait and in
// the thread work between waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DRGL_TC_TRACE -I relationalgraphlearning_b200/csrc \
//        -o tools/trace_tc tools/trace_tc.cu relationalgraphlearning_b200/csrc/graph_forward_tc.cu relationalgraphlearning_b200/csrc/pack_plan.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kernels.h"
namespace rgl { void tc_trace_read(unsigned long long* out); void tc_trace_reset(); }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)
static float* dev_rand(size_t n, float scale) {
    std::vector<float> h(n);
    for (auto& v : h) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * scale;
    float* d; CK(cudaMalloc(&d, n * 4)); CK(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice));
    return d;
}
int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 1 << 20, Nh = argc > 2 ? atoi(argv[2]) : 5, want_H = argc > 3 ? atoi(argv[3]) : 1;
    RglGraphParams p;
    p.wr0_w = dev_rand(64 * 9, 0.3f); p.wr0_b = dev_rand(64, 0.3f); p.wr1_w = dev_rand(32 * 64, 0.12f); p.wr1_b = dev_rand(32, 0.1f);
    p.wh0_w = dev_rand(64 * 5, 0.4f); p.wh0_b = dev_rand(64, 0.4f); p.wh1_w = dev_rand(32 * 64, 0.12f); p.wh1_b = dev_rand(32, 0.1f);
    p.w_a = dev_rand(32 * 32, 0.2f); p.Ws[0] = dev_rand(32 * 32, 0.2f); p.Ws[1] = dev_rand(32 * 32, 0.2f); p.num_layer = 2;
    float* blob; CK(cudaMalloc(&blob, rgl::graph_floats_total(2) * 4));
    CK(rgl::run_pack_graph(p, blob, 0));
    rgl::GraphArgs a = {};
    a.robot = dev_rand((size_t)B * 9, 3.f); a.humans = dev_rand((size_t)B * Nh * 5, 3.f);
    a.B = B; a.Nh = Nh; a.hb = 1; a.gw = blob; a.mw = nullptr; a.L = 2; a.flags = RGL_FLAG_SKIP;
    CK(cudaMalloc(&a.H, (size_t)B * (Nh + 1) * 32 * 4)); CK(cudaMalloc(&a.E, (size_t)B * 32 * 4));
    if (!want_H) a.H = nullptr;
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    int smem; CK(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0));
    for (int it = 0; it < 3; ++it) {
        rgl::tc_trace_reset();
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        CK(rgl::run_graph_forward_tc(a, sms, smem, 0));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        unsigned long long t[32]; rgl::tc_trace_read(t);
        const double tiles = (double)t[31];
        if (it < 2) continue;
        printf("B=%d Nh=%d H=%d: %.1f us, %.0f M states/s, traced tiles %.0f\n", B, Nh, want_H, ms * 1e3, B / (ms * 1e3), tiles);
        const char* names[15] = {"publish emb1 (barrier)", "wait emb1 MMA", "publish emb2a", "wait emb2a MMA", "publish emb2b", "wait emb2b MMA",
                                 "publish layer (x2)", "wait layer MMA (x2)", "sync after sim", "sync after HW sts (x2)", "sync layerwise",
                                 "publish motion", "wait motion", "sync before H stage", "sync after H stage"};
        double tot = 0;
        for (int k = 0; k < 15; ++k) { printf("  wait  %-26s %8.0f cycles/tile\n", names[k], t[k] / tiles); tot += t[k] / tiles; }
        printf("  work  before publish (split..)   %8.0f\n  work  tcgen05.wait::st            %8.0f\n  work  before mma_wait             %8.0f\n  work  before group_sync           %8.0f\n",
               t[20] / tiles, t[23] / tiles, t[21] / tiles, t[22] / tiles);
        tot += (t[20] + t[21] + t[22] + t[23]) / tiles;
        printf("  total accounted %.0f cycles per tile\n", tot);
    }
    return 0;
}
