// extern "C" boundary of librgl_b200.so (declared in include/rgl_b200.h): argument validation, device
// queries, launches.  No torch types, no exceptions, no hidden synchronisation.
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include "kernels.h"

namespace {
thread_local char g_err[256] = "";

int fail(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
int fail_cuda(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return RGL_ECUDA;
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct DevInfo { int dev; int sms; size_t max_smem; };
// queried per call for the current device (cheap attribute reads; no global mutable state to race on)
int dev_info(DevInfo* d) {
    cudaError_t e = cudaGetDevice(&d->dev);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDevice");
    int v = 0;
    e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d->dev);
    if (e != cudaSuccess) return fail_cuda(e, "cudaDeviceGetAttribute(sm count)");
    d->sms = v;
    e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, d->dev);
    if (e != cudaSuccess) return fail_cuda(e, "cudaDeviceGetAttribute(smem optin)");
    d->max_smem = (size_t)v;
    return RGL_OK;
}
}  // namespace

extern "C" {

int rgl_version(void) { return RGL_B200_VERSION; }
const char* rgl_last_error_string(void) { return g_err; }

size_t rgl_packed_graph_floats(int num_layer) {
    if (num_layer < 1 || num_layer > RGL_MAX_LAYERS) return 0;
    return (size_t)rgl::graph_floats_total(num_layer);
}
size_t rgl_packed_value_floats(void) { return rgl::VALUE_FLOATS_TOTAL; }
size_t rgl_packed_motion_floats(void) { return rgl::MOTION_FLOATS_TOTAL; }

int rgl_pack_graph(const RglGraphParams* p, float* packed, rgl_stream_t stream) {
    if (!p || !packed) return fail(RGL_EINVAL, "rgl_pack_graph: null argument");
    if (p->num_layer < 1 || p->num_layer > RGL_MAX_LAYERS) return fail(RGL_EUNSUPPORTED, "rgl_pack_graph: num_layer out of range");
    if (!p->wr0_w || !p->wr0_b || !p->wr1_w || !p->wr1_b || !p->wh0_w || !p->wh0_b || !p->wh1_w || !p->wh1_b || !p->w_a)
        return fail(RGL_EINVAL, "rgl_pack_graph: null parameter tensor");
    for (int l = 0; l < p->num_layer; ++l)
        if (!p->Ws[l]) return fail(RGL_EINVAL, "rgl_pack_graph: null Ws tensor");
    cudaError_t e = rgl::run_pack_graph(*p, packed, (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_pack_graph");
}
int rgl_pack_value(const RglValueParams* p, float* packed, rgl_stream_t stream) {
    if (!p || !packed || !p->w0 || !p->b0 || !p->w1 || !p->b1 || !p->w2 || !p->b2 || !p->w3 || !p->b3)
        return fail(RGL_EINVAL, "rgl_pack_value: null argument");
    cudaError_t e = rgl::run_pack_value(*p, packed, (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_pack_value");
}
int rgl_pack_motion(const RglMotionParams* p, float* packed, rgl_stream_t stream) {
    if (!p || !packed || !p->w0 || !p->b0 || !p->w1 || !p->b1) return fail(RGL_EINVAL, "rgl_pack_motion: null argument");
    cudaError_t e = rgl::run_pack_motion(*p, packed, (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_pack_motion");
}

int rgl_graph_forward(const float* robot, const float* humans, int B, int Nh, int humans_bcast,
                      const float* graph_packed, int num_layer, int flags, const float* motion_packed,
                      float* H, float* E, float* S, float* A0, rgl_stream_t stream) {
    if (B == 0) return RGL_OK;          /* empty batch: nothing to do (empty tensors have null data pointers) */
    if (!robot || !humans || !graph_packed) return fail(RGL_EINVAL, "rgl_graph_forward: null input");
    if (!H && !E && !S) return fail(RGL_EINVAL, "rgl_graph_forward: no output requested");
    if (S && !motion_packed) return fail(RGL_EINVAL, "rgl_graph_forward: S requested without motion_packed");
    if (B < 0 || humans_bcast < 1) return fail(RGL_EINVAL, "rgl_graph_forward: bad batch / humans_bcast");
    if (Nh < 1 || Nh > RGL_MAX_HUMANS) return fail(RGL_EUNSUPPORTED, "rgl_graph_forward: human count outside [1,31]");
    if (num_layer < 1 || num_layer > RGL_MAX_LAYERS) return fail(RGL_EUNSUPPORTED, "rgl_graph_forward: num_layer out of range");
    if (flags & ~(RGL_FLAG_SKIP | RGL_FLAG_LAYERWISE | RGL_FLAG_THROUGHPUT | RGL_FLAG_FP32_FMA)) return fail(RGL_EINVAL, "rgl_graph_forward: unknown flag");
    if (!aligned16(graph_packed) || (motion_packed && !aligned16(motion_packed)))
        return fail(RGL_EALIGN, "rgl_graph_forward: packed weights must be 16-byte aligned");
    if ((H && !aligned16(H)) || (E && !aligned16(E))) return fail(RGL_EALIGN, "rgl_graph_forward: H/E must be 16-byte aligned");
    if (B == 0) return RGL_OK;
    DevInfo d;
    if (int rc = dev_info(&d)) return rc;
    rgl::GraphArgs a;
    a.robot = robot; a.humans = humans; a.B = B; a.Nh = Nh; a.hb = humans_bcast;
    a.gw = graph_packed; a.mw = S ? motion_packed : nullptr; a.L = num_layer; a.flags = flags;
    a.H = H; a.E = E; a.S = S; a.A0 = A0; a.ntiles = 0; a.save = 0;
    memset(&a.sv, 0, sizeof(a.sv));
    a.use_tma = (humans_bcast == 1 && aligned16(robot) && aligned16(humans)) ? 1 : 0;
    // Inference runs on the tcgen05 kernel (3xTF32 products, fp32 accumulation in TMEM); RGL_FLAG_FP32_FMA keeps every
    // product on the fp32 FMA pipe.  RGL_GRAPH_VARIANT (experiments only): 't' = tcgen05, anything else selects one of
    // the legacy FFMA / mma.sync variants of graph_forward.cu.
    static const char* variant = getenv("RGL_GRAPH_VARIANT");
    const bool tc = (variant ? (variant[0] == 't' || variant[0] == 'p') : true) && !(flags & RGL_FLAG_FP32_FMA);
    // 'p' = the row-paired tcgen05 kernel (graph_forward_tp.cu; compiled node counts n = 6, 11, 21).  Default policy from
    // the B200 measurements (gpurun_out/r2_qt_{t,p}.log, states/s at steady state, t -> p): H 1176 -> 1234 M and S 1013 ->
    // 1072 M at n = 6, H 307 -> 367 M, E 353 -> 389 M at n = 11; but the value path (E only), whose last layer runs for the
    // robot rows alone, is faster on the robot-first layout of graph_forward_tc.cu at n = 6 (1321 vs 1265 M) and n = 21.
    const bool e_only = !H && !S;
    const bool paired = tc && (variant ? variant[0] == 'p' : (!e_only || Nh + 1 == 11));
    cudaError_t e = cudaErrorNotSupported;
    if (paired) e = rgl::run_graph_forward_tp(a, d.sms, d.max_smem, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported)
        e = tc ? rgl::run_graph_forward_tc(a, d.sms, d.max_smem, (cudaStream_t)stream)
               : rgl::run_graph_forward(a, d.sms, d.max_smem, (cudaStream_t)stream);
    if (e == cudaErrorInvalidConfiguration) return fail(RGL_EUNSUPPORTED, "rgl_graph_forward: tile does not fit in shared memory");
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_graph_forward");
}

int rgl_value_head(const float* E, int B, const float* value_packed, float* V, rgl_stream_t stream) {
    if (B == 0) return RGL_OK;
    if (!E || !value_packed || !V || B < 0) return fail(RGL_EINVAL, "rgl_value_head: bad argument");
    if (!aligned16(value_packed)) return fail(RGL_EALIGN, "rgl_value_head: packed weights must be 16-byte aligned");
    if (B == 0) return RGL_OK;
    DevInfo d;
    if (int rc = dev_info(&d)) return rc;
    // Inference runs the value network on tcgen05 (3xTF32, value_head_tc.cu): faster than the fp32-FMA kernel at every batch
    // size (6.6 vs 8.1 us at B = 4096, 16 vs 92 us at B = 65536).  RGL_VALUE_VARIANT (experiments only): 't' / 'f' force one.
    static const char* variant = getenv("RGL_VALUE_VARIANT");
    const bool tc = aligned16(E) && (variant ? variant[0] == 't' : true);
    cudaError_t e = tc ? rgl::run_value_head_tc(E, B, value_packed, V, d.sms, d.max_smem, (cudaStream_t)stream)
                       : rgl::run_value_head(E, B, value_packed, V, nullptr, nullptr, nullptr, aligned16(E) ? 1 : 0, d.sms, (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_value_head");
}

int rgl_value_forward(const float* robot, const float* humans, int B, int Nh, int humans_bcast,
                      const float* graph_packed, int num_layer, int flags, const float* value_packed,
                      float* E_scratch, float* V, float* A0, rgl_stream_t stream) {
    if (B == 0) return RGL_OK;
    if (!E_scratch || !V) return fail(RGL_EINVAL, "rgl_value_forward: null output");
    int rc = rgl_graph_forward(robot, humans, B, Nh, humans_bcast, graph_packed, num_layer, flags, nullptr,
                               nullptr, E_scratch, nullptr, A0, stream);
    if (rc) return rc;
    return rgl_value_head(E_scratch, B, value_packed, V, stream);
}

int rgl_gcn_layer(const float* X, const float* A, const float* W, const float* w_a, int B, int n, int flags,
                  float* Hout, float* Aout, rgl_stream_t stream) {
    if (B == 0) return RGL_OK;
    if (!X || !W || !Hout || B < 0) return fail(RGL_EINVAL, "rgl_gcn_layer: bad argument");
    if (!A && !w_a) return fail(RGL_EINVAL, "rgl_gcn_layer: need A or w_a");
    if (n < 2 || n > RGL_MAX_HUMANS + 1) return fail(RGL_EUNSUPPORTED, "rgl_gcn_layer: n outside [2,32]");
    if (flags & ~RGL_FLAG_SKIP) return fail(RGL_EINVAL, "rgl_gcn_layer: unknown flag");
    if (!aligned16(X) || !aligned16(Hout) || !aligned16(W) || (w_a && !aligned16(w_a)))
        return fail(RGL_EALIGN, "rgl_gcn_layer: X/W/w_a/Hout must be 16-byte aligned");
    if (B == 0) return RGL_OK;
    DevInfo d;
    if (int rc = dev_info(&d)) return rc;
    // tcgen05 kernel (3xTF32 X W, gcn_layer_tc.cu) unless RGL_GCN_VARIANT=f (experiments only) selects the fp32-FMA kernel
    static const char* variant = getenv("RGL_GCN_VARIANT");
    const bool tc = variant ? variant[0] != 'f' : true;
    cudaError_t e = tc ? rgl::run_gcn_layer_tc(X, A, W, w_a, B, n, flags, Hout, Aout, d.sms, d.max_smem, (cudaStream_t)stream)
                       : rgl::run_gcn_layer(X, A, W, w_a, B, n, flags, Hout, Aout, d.sms, d.max_smem, (cudaStream_t)stream);
    if (e == cudaErrorInvalidConfiguration) return fail(RGL_EUNSUPPORTED, "rgl_gcn_layer: tile does not fit in shared memory");
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_gcn_layer");
}

int rgl_plan_expand(const float* robot, const float* humans, int E, int Nh, int humans_bcast, const double* actions, int A,
                    double time_step, int kinematics,
                    float* next_robot, float* reward, rgl_stream_t stream) {
    if (E == 0) return RGL_OK;
    if (!robot || !actions || E < 0 || A < 1 || Nh < 0 || humans_bcast < 1) return fail(RGL_EINVAL, "rgl_plan_expand: bad argument");
    if (kinematics != RGL_KIN_HOLONOMIC && kinematics != RGL_KIN_UNICYCLE) return fail(RGL_EINVAL, "rgl_plan_expand: unknown kinematics");
    if (reward && Nh > 0 && !humans) return fail(RGL_EINVAL, "rgl_plan_expand: reward needs humans");
    if (!next_robot && !reward) return fail(RGL_EINVAL, "rgl_plan_expand: no output requested");
    if ((long long)E * A > 0x7fffffffLL) return fail(RGL_EUNSUPPORTED, "rgl_plan_expand: E*A too large");
    cudaError_t e = rgl::run_plan_expand(robot, humans, E, Nh, humans_bcast, actions, A, time_step, kinematics == RGL_KIN_UNICYCLE,
                                         next_robot, reward, (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_plan_expand");
}

int rgl_plan_argmax(const float* reward, const float* V, int E, int A, float gamma_bar, float* value, int* best,
                    const int* act_map, int* best_action, rgl_stream_t stream) {
    if (E == 0) return RGL_OK;
    if (!reward || !V || E < 0 || A < 1 || (!value && !best && !best_action)) return fail(RGL_EINVAL, "rgl_plan_argmax: bad argument");
    cudaError_t e = rgl::run_plan_argmax(reward, V, E, A, gamma_bar, value, best, act_map, best_action, (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_plan_argmax");
}

int rgl_plan_select(const float* reward, const float* V, int E, int A, float gamma_bar, int width, const int* groups,
                    const float* next_robot, int* acts, float* child_reward, float* child_robot, float* value,
                    rgl_stream_t stream) {
    if (E == 0) return RGL_OK;
    if (!reward || !V || !acts || E < 0 || A < 1 || width < 1 || width > A) return fail(RGL_EINVAL, "rgl_plan_select: bad argument");
    if (child_robot && !next_robot) return fail(RGL_EINVAL, "rgl_plan_select: child_robot needs next_robot");
    if (A > 256) return fail(RGL_EUNSUPPORTED, "rgl_plan_select: more than 256 actions");
    cudaError_t e = rgl::run_plan_select(reward, V, E, A, gamma_bar, width, groups, next_robot, acts, child_reward, child_robot, value,
                                         (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_plan_select");
}

int rgl_plan_backup(const float* v, const float* next_v, const float* reward, int E, int W, float gamma_bar, int depth,
                    float* ret_best, int* best, rgl_stream_t stream) {
    if (E == 0) return RGL_OK;
    if (!v || !next_v || !reward || !ret_best || !best || E < 0 || W < 1 || depth < 2) return fail(RGL_EINVAL, "rgl_plan_backup: bad argument");
    cudaError_t e = rgl::run_plan_backup(v, next_v, reward, E, W, gamma_bar, depth, ret_best, best, (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_plan_backup");
}

int rgl_graph_forward_train(const float* robot, const float* humans, int B, int Nh, const float* graph_packed, int num_layer,
                            int flags, const float* motion_packed, const RglGraphSave* save, float* H, float* E, float* S,
                            rgl_stream_t stream) {
    if (B == 0) return RGL_OK;
    if (!robot || !humans || !graph_packed || !save) return fail(RGL_EINVAL, "rgl_graph_forward_train: null argument");
    if (flags & RGL_FLAG_LAYERWISE) return fail(RGL_EUNSUPPORTED, "rgl_graph_forward_train: layerwise graphs are not supported");
    if (flags & ~(RGL_FLAG_SKIP | RGL_FLAG_THROUGHPUT | RGL_FLAG_TRAIN_TC)) return fail(RGL_EINVAL, "rgl_graph_forward_train: unknown flag");
    if (B < 0 || Nh < 1 || Nh > RGL_MAX_HUMANS) return fail(RGL_EUNSUPPORTED, "rgl_graph_forward_train: bad batch / human count");
    if (num_layer < 1 || num_layer > RGL_MAX_LAYERS) return fail(RGL_EUNSUPPORTED, "rgl_graph_forward_train: num_layer out of range");
    if (!save->a1r || !save->a1h || !save->X || !save->Y || !save->A) return fail(RGL_EINVAL, "rgl_graph_forward_train: missing save buffer");
    for (int l = 0; l < num_layer; ++l)
        if (!save->M[l] || !save->Rl[l] || !save->Hl[l]) return fail(RGL_EINVAL, "rgl_graph_forward_train: missing per-layer save buffer");
    if (!aligned16(graph_packed) || !aligned16(save->a1r) || !aligned16(save->a1h) || !aligned16(save->X) || !aligned16(save->Y))
        return fail(RGL_EALIGN, "rgl_graph_forward_train: buffers must be 16-byte aligned");
    DevInfo d;
    if (int rc = dev_info(&d)) return rc;
    rgl::GraphArgs a;
    a.robot = robot; a.humans = humans; a.B = B; a.Nh = Nh; a.hb = 1;
    if (S && (!motion_packed || !aligned16(motion_packed))) return fail(RGL_EINVAL, "rgl_graph_forward_train: S needs 16-byte aligned motion_packed");
    if (!H && !E && !S) return fail(RGL_EINVAL, "rgl_graph_forward_train: no output requested");
    a.gw = graph_packed; a.mw = S ? motion_packed : nullptr; a.L = num_layer; a.flags = flags;
    a.H = H; a.E = E; a.S = S; a.A0 = nullptr; a.ntiles = 0; a.save = 1; a.sv = *save;
    a.use_tma = (aligned16(robot) && aligned16(humans)) ? 1 : 0;
    if (flags & RGL_FLAG_TRAIN_TC) {
        // tcgen05 training forward (graph_forward_tp.cu): hidden activations in ONE [B,n,64] buffer, M[l] = H_{l-1} W_l
        if (save->a1h != save->a1r + RGL_EMB_HIDDEN) return fail(RGL_EINVAL, "rgl_graph_forward_train: TC layout needs a1h == a1r + 64 ([B,n,64] buffer)");
        if (S && !save->mh) return fail(RGL_EINVAL, "rgl_graph_forward_train: S needs the mh save buffer");
        a.save = 2;
        a.flags = flags & ~RGL_FLAG_TRAIN_TC;
        cudaError_t e2 = rgl::run_graph_forward_tp(a, d.sms, d.max_smem, (cudaStream_t)stream);
        if (e2 == cudaErrorNotSupported) return fail(RGL_EUNSUPPORTED, "rgl_graph_forward_train: RGL_FLAG_TRAIN_TC covers Nh = 5, 10, 20");
        if (e2 == cudaErrorInvalidConfiguration) return fail(RGL_EUNSUPPORTED, "rgl_graph_forward_train: tile does not fit in shared memory");
        return e2 == cudaSuccess ? RGL_OK : fail_cuda(e2, "rgl_graph_forward_train");
    }
    cudaError_t e = rgl::run_graph_forward(a, d.sms, d.max_smem, (cudaStream_t)stream);
    if (e == cudaErrorInvalidConfiguration) return fail(RGL_EUNSUPPORTED, "rgl_graph_forward_train: tile does not fit in shared memory");
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_graph_forward_train");
}

int rgl_value_head_train(const float* E, int B, const float* value_packed, float* V, float* v0, float* v1, float* v2,
                         rgl_stream_t stream) {
    if (B == 0) return RGL_OK;
    if (!E || !value_packed || !V || !v0 || !v1 || !v2 || B < 0) return fail(RGL_EINVAL, "rgl_value_head_train: bad argument");
    if (!aligned16(value_packed)) return fail(RGL_EALIGN, "rgl_value_head_train: packed weights must be 16-byte aligned");
    DevInfo d;
    if (int rc = dev_info(&d)) return rc;
    // same kernel as inference (tcgen05, 3xTF32) with the activation saves written from the epilogues; the fp32-FMA kernel
    // keeps unaligned buffers and RGL_VALUE_VARIANT=f (experiments only)
    static const char* variant = getenv("RGL_VALUE_VARIANT");
    const bool tc = aligned16(E) && aligned16(v0) && aligned16(v1) && aligned16(v2) && !(variant && variant[0] == 'f');
    cudaError_t e = tc ? rgl::run_value_head_tc_train(E, B, value_packed, V, v0, v1, v2, d.sms, d.max_smem, (cudaStream_t)stream)
                       : rgl::run_value_head(E, B, value_packed, V, v0, v1, v2, aligned16(E) ? 1 : 0, d.sms, (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_value_head_train");
}

int rgl_linear_bwd(const RglRows* G, int N, const RglRows* mask, const RglRows* Xin, int K, const float* W, int w_layout,
                   const RglRows* Gin, int accumulate, float* dW, float* db, int R, rgl_stream_t stream) {
    if (R == 0) return RGL_OK;
    if (!G || !G->ptr || R < 0) return fail(RGL_EINVAL, "rgl_linear_bwd: null gradient");
    if (N < 1 || N > 128 || K < 1 || K > 128) return fail(RGL_EUNSUPPORTED, "rgl_linear_bwd: N, K must be in [1,128]");
    if (w_layout != 0 && w_layout != 1) return fail(RGL_EINVAL, "rgl_linear_bwd: bad w_layout");
    if (Gin && Gin->ptr && !W) return fail(RGL_EINVAL, "rgl_linear_bwd: data gradient needs W");
    if (dW && (!Xin || !Xin->ptr)) return fail(RGL_EINVAL, "rgl_linear_bwd: weight gradient needs Xin");
    if ((!Gin || !Gin->ptr) && !dW && !db) return fail(RGL_EINVAL, "rgl_linear_bwd: no output requested");
    DevInfo d;
    if (int rc = dev_info(&d)) return rc;
    cudaError_t e = rgl::run_linear_bwd(G, N, mask, Xin, K, W, w_layout, Gin, accumulate, dW, db, R, d.sms, d.max_smem, (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_linear_bwd");
}

int rgl_mlp2_bwd(const RglRows* G, const RglRows* mask, const RglRows* hidden, const float* W1, const RglRows* X0, int K0,
                 float* dW1, float* db1, float* dW0, float* db0, int R, rgl_stream_t stream) {
    if (R == 0) return RGL_OK;
    if (!G || !G->ptr || !hidden || !hidden->ptr || !X0 || !X0->ptr || !W1 || !dW1 || !dW0 || R < 0)
        return fail(RGL_EINVAL, "rgl_mlp2_bwd: bad argument");
    if (K0 < 1 || K0 > 16) return fail(RGL_EUNSUPPORTED, "rgl_mlp2_bwd: K0 must be in [1,16]");
    DevInfo d;
    if (int rc = dev_info(&d)) return rc;
    cudaError_t e = rgl::run_mlp2_bwd(G, mask, hidden, W1, X0, K0, dW1, db1, dW0, db0, R, d.sms, d.max_smem, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported) return fail(RGL_EUNSUPPORTED, "rgl_mlp2_bwd: shape not supported");
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_mlp2_bwd");
}

int rgl_attn_layer_bwd(const float* A, const float* Hprev, const float* gM, const float* gH, int skip, float* gHprev, float* gA,
                       int accumulate_gA, int B, int n, const float* mask, int up_rows, rgl_stream_t stream) {
    if (B == 0) return RGL_OK;
    if (!A || !Hprev || !gM || !gHprev || !gA || (skip && !gH) || B < 0 || n < 1 || n > 32)
        return fail(RGL_EINVAL, "rgl_attn_layer_bwd: bad argument");
    if (up_rows < 1 || up_rows > n) return fail(RGL_EINVAL, "rgl_attn_layer_bwd: up_rows must be in [1, n]");
    if (!aligned16(Hprev) || !aligned16(gM) || !aligned16(gHprev) || (gH && !aligned16(gH)) || (mask && !aligned16(mask)))
        return fail(RGL_EALIGN, "rgl_attn_layer_bwd: row buffers must be 16-byte aligned");
    cudaError_t e = rgl::run_attn_layer_bwd(A, Hprev, gM, gH, skip, gHprev, gA, accumulate_gA, B, n, mask, up_rows, (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_attn_layer_bwd");
}

int rgl_attn_sim_bwd(const float* A, const float* Z, const float* gM, const float* mask, int up_rows, const float* gA_in,
                     float* gZ, float* gA_out, const float* X, const float* Y, float* gY, float* gX, int gx_accumulate,
                     int B, int n, rgl_stream_t stream) {
    if (B == 0) return RGL_OK;
    if (!A || !Z || !gM || !gZ || B < 0 || n < 1 || n > 32) return fail(RGL_EINVAL, "rgl_attn_sim_bwd: bad argument");
    if (up_rows < 1 || up_rows > n) return fail(RGL_EINVAL, "rgl_attn_sim_bwd: up_rows must be in [1, n]");
    if (X ? (!Y || !gY || !gX) : !gA_out) return fail(RGL_EINVAL, "rgl_attn_sim_bwd: X needs Y, gY, gX; without X gA_out is required");
    if (!aligned16(Z) || !aligned16(gM) || !aligned16(gZ) || (mask && !aligned16(mask)) ||
        (X && (!aligned16(X) || !aligned16(Y) || !aligned16(gY) || !aligned16(gX))))
        return fail(RGL_EALIGN, "rgl_attn_sim_bwd: row buffers must be 16-byte aligned");
    DevInfo d;
    if (int rc = dev_info(&d)) return rc;
    cudaError_t e = rgl::run_attn_sim_bwd(A, Z, gM, mask, up_rows, gA_in, gZ, gA_out, X, Y, gY, gX, gx_accumulate, B, n, d.max_smem,
                                          (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_attn_sim_bwd");
}

int rgl_sim_bwd(const float* A, const float* gA, const float* X, const float* Y, float* gY, float* gX, int B, int n,
                rgl_stream_t stream) {
    if (B == 0) return RGL_OK;
    if (!A || !gA || !X || !Y || !gY || !gX || B < 0 || n < 1 || n > 32) return fail(RGL_EINVAL, "rgl_sim_bwd: bad argument");
    if (!aligned16(X) || !aligned16(Y) || !aligned16(gY) || !aligned16(gX)) return fail(RGL_EALIGN, "rgl_sim_bwd: row buffers must be 16-byte aligned");
    cudaError_t e = rgl::run_sim_bwd(A, gA, X, Y, gY, gX, B, n, (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_sim_bwd");
}

int rgl_td_loss(const float* V, const float* reward, const float* V_next, int B, float gamma_bar, float inv_count, float* loss,
                float* gV, rgl_stream_t stream) {
    if (B == 0) return RGL_OK;
    if (!V || !reward || !V_next || !loss || B < 0) return fail(RGL_EINVAL, "rgl_td_loss: bad argument");
    cudaError_t e = rgl::run_td_loss(V, reward, V_next, B, gamma_bar, inv_count, loss, gV, (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_td_loss");
}

int rgl_replay_record_floats(int Nh) { return Nh < 1 || Nh > RGL_MAX_HUMANS ? 0 : 2 * RGL_ROBOT_DIM + 2 * RGL_HUMAN_DIM * Nh + 2; }

int rgl_replay_push(float* store, long long slot, int Nh, const float* robot, const float* humans, const float* value,
                    const float* reward, const float* next_robot, const float* next_humans, rgl_stream_t stream) {
    if (!store || slot < 0 || !robot || !humans || !value || !reward || !next_robot || !next_humans)
        return fail(RGL_EINVAL, "rgl_replay_push: bad argument");
    if (Nh < 1 || Nh > RGL_MAX_HUMANS) return fail(RGL_EUNSUPPORTED, "rgl_replay_push: human count outside [1,31]");
    cudaError_t e = rgl::run_replay_push(store, slot, Nh, robot, humans, value, reward, next_robot, next_humans, (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_replay_push");
}

int rgl_replay_gather(const float* store, const long long* idx, int B, int Nh, float* robot, float* humans, float* value,
                      float* reward, float* next_robot, float* next_humans, rgl_stream_t stream) {
    if (B == 0) return RGL_OK;
    if (!store || !idx || B < 0 || !robot || !humans || !value || !reward || !next_robot || !next_humans)
        return fail(RGL_EINVAL, "rgl_replay_gather: bad argument");
    if (Nh < 1 || Nh > RGL_MAX_HUMANS) return fail(RGL_EUNSUPPORTED, "rgl_replay_gather: human count outside [1,31]");
    cudaError_t e = rgl::run_replay_gather(store, idx, B, Nh, robot, humans, value, reward, next_robot, next_humans, (cudaStream_t)stream);
    return e == cudaSuccess ? RGL_OK : fail_cuda(e, "rgl_replay_gather");
}

}  // extern "C"
