// Internal launch interface between capi.cu and the kernel translation units.
#pragma once
#include "common.cuh"

namespace rgl {

struct GraphArgs {
    const float* robot;
    const float* humans;
    int B, Nh, hb;            // hb = humans_bcast
    const float* gw;          // packed graph blob
    const float* mw;          // packed motion blob (only when S is requested)
    int L, flags;
    float* H;
    float* E;
    float* S;
    float* A0;
    int ntiles;
    int use_tma;
};

cudaError_t run_graph_forward(const GraphArgs& a, int num_sms, size_t max_smem, cudaStream_t st);
cudaError_t run_value_head(const float* E, int B, const float* vw, float* V, int use_tma, int num_sms, cudaStream_t st);
cudaError_t run_gcn_layer(const float* X, const float* A, const float* W, const float* wa, int B, int n, int flags,
                          float* Hout, float* Aout, int num_sms, size_t max_smem, cudaStream_t st);
cudaError_t run_pack_graph(const RglGraphParams& p, float* out, cudaStream_t st);
cudaError_t run_pack_value(const RglValueParams& p, float* out, cudaStream_t st);
cudaError_t run_pack_motion(const RglMotionParams& p, float* out, cudaStream_t st);
cudaError_t run_plan_expand(const float* robot, const float* humans, int E, int Nh, int hb, const double* actions, int A, double dt,
                            float* next_robot, float* reward, cudaStream_t st);
cudaError_t run_plan_argmax(const float* reward, const float* V, int E, int A, float gamma_bar, float* value, int* best,
                            cudaStream_t st);

}  // namespace rgl
