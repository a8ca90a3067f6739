// Internal launch interface between capi.cu and the kernel translation units.
#pragma once
#include <mutex>
#include <stdlib.h>
#include <utility>
#include "common.cuh"

namespace rgl {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once per (kernel, device).  A small table
// keyed by the kernel's address (one table per translation unit), guarded by a mutex.
static inline cudaError_t ensure_dyn_smem_ptr(const void* kernel, int bytes) {
    struct Entry { const void* k; unsigned long long done; };
    static Entry table[64];
    static int count = 0;
    static std::mutex mu;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    Entry* ent = nullptr;
    for (int i = 0; i < count; ++i)
        if (table[i].k == kernel) { ent = &table[i]; break; }
    if (ent == nullptr && count < 64) { ent = &table[count++]; ent->k = kernel; ent->done = 0; }
    if (ent != nullptr && dev < 64 && (ent->done >> dev & 1ull)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && ent != nullptr && dev < 64) ent->done |= 1ull << dev;
    return e;
}
template <typename Kernel>
static inline cudaError_t ensure_dyn_smem(Kernel kernel, int bytes) { return ensure_dyn_smem_ptr(reinterpret_cast<const void*>(kernel), bytes); }

// Launch with the programmatic-stream-serialization attribute (PDL): only for kernels that call pdl_wait() before touching
// memory a predecessor may have written.  RGL_PDL=0 launches them fully serialised (experiments / debugging).
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    static const char* env = getenv("RGL_PDL");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (env && env[0] == '0') ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Internal flag bit of GraphArgs::flags (never part of the C ABI): trigger the programmatic launch of the next kernel at the START
// of this one instead of at its output phase.  Policy (pdl_early): RGL_PDL_EARLY=1 forces it, =0 forbids it.
#define RGL_INTERNAL_PDL_EARLY (1 << 30)
static inline bool pdl_early(int B) {
    static const char* env = getenv("RGL_PDL_EARLY");
    if (env) return env[0] == '1';
    (void)B;
    return false;
}

typedef RglGraphSave GraphSave;      // optional activation saves of the training forward (include/rgl_b200.h)

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
    int save;                 // 1 = sv holds pointers (training forward, fp32-FMA kernel); 2 = tcgen05 training forward (TC save layout,
                              // include/rgl_b200.h RGL_FLAG_TRAIN_TC); disables the robot-row-only last layer
    GraphSave sv;
};

cudaError_t run_graph_forward(const GraphArgs& a, int num_sms, size_t max_smem, cudaStream_t st);
// tcgen05 / TMEM variant (graph_forward_tc.cu): inference only (no activation saves)
cudaError_t run_graph_forward_tc(const GraphArgs& a, int num_sms, size_t max_smem, cudaStream_t st);
// row-paired tcgen05 variant (graph_forward_tp.cu): n = 6, 11, 21; cudaErrorNotSupported otherwise
cudaError_t run_graph_forward_tp(const GraphArgs& a, int num_sms, size_t max_smem, cudaStream_t st);
cudaError_t run_value_head(const float* E, int B, const float* vw, float* V, float* v0, float* v1, float* v2, int use_tma, int num_sms,
                           cudaStream_t st);
// tcgen05 value network (value_head_tc.cu): inference only
cudaError_t run_value_head_tc(const float* E, int B, const float* vw, float* V, int num_sms, size_t max_smem, cudaStream_t st);
cudaError_t run_value_head_tc_train(const float* E, int B, const float* vw, float* V, float* v0, float* v1, float* v2, int num_sms,
                                    size_t max_smem, cudaStream_t st);
cudaError_t run_linear_bwd(const RglRows* G, int N, const RglRows* mask, const RglRows* Xin, int K, const float* W, int w_layout,
                           const RglRows* Gin, int accumulate, float* dW, float* db, int R, int num_sms, size_t max_smem,
                           cudaStream_t st);
cudaError_t run_mlp2_bwd(const RglRows* G, const RglRows* mask, const RglRows* hidden, const float* W1, const RglRows* X0, int K0,
                         float* dW1, float* db1, float* dW0, float* db0, int R, int num_sms, size_t max_smem, cudaStream_t st);
cudaError_t run_attn_layer_bwd(const float* A, const float* Hprev, const float* gM, const float* gH, int skip, float* gHprev,
                               float* gA, int accumulate_gA, int B, int n, const float* mask, int up_rows, cudaStream_t st);
cudaError_t run_attn_sim_bwd(const float* A, const float* Z, const float* gM, const float* mask, int up_rows, const float* gA_in,
                             float* gZ, float* gA_out, const float* X, const float* Y, float* gY, float* gX, int gx_accumulate,
                             int B, int n, size_t max_smem, cudaStream_t st);
cudaError_t run_sim_bwd(const float* A, const float* gA, const float* X, const float* Y, float* gY, float* gX, int B, int n,
                        cudaStream_t st);
cudaError_t run_gcn_layer(const float* X, const float* A, const float* W, const float* wa, int B, int n, int flags,
                          float* Hout, float* Aout, int num_sms, size_t max_smem, cudaStream_t st);
// tcgen05 variant (gcn_layer_tc.cu)
cudaError_t run_gcn_layer_tc(const float* X, const float* A, const float* W, const float* wa, int B, int n, int flags,
                             float* Hout, float* Aout, int num_sms, size_t max_smem, cudaStream_t st);
cudaError_t run_pack_graph(const RglGraphParams& p, float* out, cudaStream_t st);
cudaError_t run_pack_value(const RglValueParams& p, float* out, cudaStream_t st);
cudaError_t run_pack_motion(const RglMotionParams& p, float* out, cudaStream_t st);
cudaError_t run_plan_expand(const float* robot, const float* humans, int E, int Nh, int hb, const double* actions, int A, double dt,
                            int unicycle, float* next_robot, float* reward, cudaStream_t st);
cudaError_t run_plan_argmax(const float* reward, const float* V, int E, int A, float gamma_bar, float* value, int* best,
                            const int* act_map, int* best_action, cudaStream_t st);
cudaError_t run_plan_select(const float* reward, const float* V, int E, int A, float gamma_bar, int width, const int* groups,
                            const float* next_robot, int* acts, float* child_rew, float* child_robot, float* value, cudaStream_t st);
cudaError_t run_plan_backup(const float* v, const float* nv, const float* rew, int E, int W, float gamma_bar, int depth,
                            float* ret_best, int* best, cudaStream_t st);

cudaError_t run_td_loss(const float* V, const float* reward, const float* Vnext, int B, float gamma_bar, float inv_count, float* loss,
                        float* gV, cudaStream_t st);
cudaError_t run_replay_gather(const float* store, const long long* idx, int B, int Nh, float* robot, float* humans, float* value,
                              float* reward, float* next_robot, float* next_humans, cudaStream_t st);
cudaError_t run_replay_push(float* store, long long slot, int Nh, const float* robot, const float* humans, const float* value,
                            const float* reward, const float* next_robot, const float* next_humans, cudaStream_t st);

}  // namespace rgl
