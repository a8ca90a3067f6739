/*
 * rgl_b200.h -- C ABI of the B200-native RGL hot path (librgl_b200.so).
 *
 * The reference (ChanganVR/RelationalGraphLearning) is pure Python/PyTorch and has no FFI; the
 * entry points below are what a binding for its hot path would bind.  Each one names the
 * reference interface it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous row-major fp32 unless said otherwise;
 *   - the caller owns every buffer; the library allocates nothing persistent;
 *   - `stream` is a cudaStream_t / CUstream handle (e.g. torch.cuda.current_stream().cuda_stream);
 *     all work is enqueued on it, there are no hidden synchronisations;
 *   - return value: 0 on success, a negative RGL_E* code otherwise; nothing throws or exits;
 *     rgl_last_error_string() gives a thread-local description of the last failure;
 *   - stateless and re-entrant: safe from several host threads / one process per GPU.
 *
 * State layouts (crowd_sim/envs/utils/state.py:27-28,51-52):
 *   robot [B,1,9]  = (px, py, vx, vy, radius, gx, gy, v_pref, theta)
 *   humans[B,Nh,5] = (px, py, vx, vy, radius)            n = Nh + 1 graph nodes, node 0 = robot
 */
#ifndef RGL_B200_H_
#define RGL_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGL_B200_VERSION 202          /* major*10000 + minor*100 + patch */

#define RGL_X_DIM        32           /* config.gcn.X_dim = final_state_dim (configs/icra_benchmark/config.py:103-108) */
#define RGL_EMB_HIDDEN   64           /* wr_dims[0] = wh_dims[0] */
#define RGL_ROBOT_DIM    9
#define RGL_HUMAN_DIM    5
#define RGL_VALUE_HIDDEN 100          /* value_network_dims = [32,100,100,1] (mp_separate.py:26) */
#define RGL_MOTION_HIDDEN 64          /* motion_predictor_dims = [64,5]      (mp_separate.py:25) */
#define RGL_MAX_LAYERS   4
#define RGL_MAX_HUMANS   31

/* error codes */
#define RGL_OK            0
#define RGL_EINVAL       -1           /* null pointer, bad size, unsupported flag combination */
#define RGL_EALIGN       -2           /* pointer not 16-byte aligned where required */
#define RGL_ECUDA        -3           /* a CUDA runtime call failed (see rgl_last_error_string) */
#define RGL_EUNSUPPORTED -4           /* shape outside what the kernels are built for */

/* flags for the graph kernels */
#define RGL_FLAG_SKIP       1         /* config.gcn.skip_connection  (graph_model.py:126-127) */
#define RGL_FLAG_LAYERWISE  2         /* config.gcn.layerwise_graph  (graph_model.py:120-122) */
#define RGL_FLAG_FP32_FMA   8         /* numerics: keep every GEMM on the fp32 FMA pipe (no 3xTF32 tensor-core split;
                                         the tensor path agrees with fp32 to ~3e-6 relative, the FMA path to ~3e-7) */
#define RGL_FLAG_TRAIN_TC   16        /* rgl_graph_forward_train only: run the training forward on the tcgen05 kernel (Nh = 5, 10, 20).
                                         Save layout: a1r = base of ONE [B,n,64] buffer (a1h must be a1r + 64), mh likewise
                                         [B,n,64], M[l] receives Z_l = H_{l-1} W_l (the layer is evaluated as relu(A (H W))),
                                         Rl[l] = relu(A Z_l), Hl[l] as usual; the backward then runs rgl_attn_layer_bwd (with
                                         mask = Rl[l], Hprev = Z_l) BEFORE rgl_linear_bwd */
#define RGL_FLAG_THROUGHPUT 4         /* scheduling hint: the caller keeps several independent launches in flight
                                         (multi-stream serving); prefer the two-CTAs-per-SM kernel variant */

typedef void* rgl_stream_t;

/* Pointers to the tensors of RGL.state_dict() in PyTorch's native layout, zero-copy
 * (crowd_nav/policy/graph_model.py:41-58; names in SURVEY.md §2b). Linear weights are [out,in]. */
typedef struct RglGraphParams {
    const float* wr0_w;   /* w_r.0.weight [64,9]  */
    const float* wr0_b;   /* w_r.0.bias   [64]    */
    const float* wr1_w;   /* w_r.2.weight [32,64] */
    const float* wr1_b;   /* w_r.2.bias   [32]    */
    const float* wh0_w;   /* w_h.0.weight [64,5]  */
    const float* wh0_b;   /* w_h.0.bias   [64]    */
    const float* wh1_w;   /* w_h.2.weight [32,64] */
    const float* wh1_b;   /* w_h.2.bias   [32]    */
    const float* w_a;     /* w_a          [32,32] */
    const float* Ws[RGL_MAX_LAYERS];   /* Ws.i [32,32] */
    int num_layer;
} RglGraphParams;

/* value_network state_dict (crowd_nav/policy/value_estimator.py:9): mlp(32,[32,100,100,1]) */
typedef struct RglValueParams {
    const float* w0; const float* b0;   /* 0.weight [32,32],   0.bias [32]  */
    const float* w1; const float* b1;   /* 2.weight [100,32],  2.bias [100] */
    const float* w2; const float* b2;   /* 4.weight [100,100], 4.bias [100] */
    const float* w3; const float* b3;   /* 6.weight [1,100],   6.bias [1]   */
} RglValueParams;

/* human_motion_predictor state_dict (crowd_nav/policy/state_predictor.py:17): mlp(32,[64,5]) */
typedef struct RglMotionParams {
    const float* w0; const float* b0;   /* 0.weight [64,32], 0.bias [64] */
    const float* w1; const float* b1;   /* 2.weight [5,64],  2.bias [5]  */
} RglMotionParams;

int         rgl_version(void);
const char* rgl_last_error_string(void);

/* ---- weight packing -------------------------------------------------------------------------
 * The kernels read weights from one contiguous k-major blob per module (staged into shared
 * memory with a single TMA bulk copy).  Packing is a tiny gather kernel; redo it after every
 * optimizer step / load_state_dict.  Sizes are in floats. */
size_t rgl_packed_graph_floats(int num_layer);
size_t rgl_packed_value_floats(void);
size_t rgl_packed_motion_floats(void);
int rgl_pack_graph (const RglGraphParams*  p, float* packed, rgl_stream_t stream);
int rgl_pack_value (const RglValueParams*  p, float* packed, rgl_stream_t stream);
int rgl_pack_motion(const RglMotionParams* p, float* packed, rgl_stream_t stream);

/* ---- fused graph forward ----------------------------------------------------------------------
 * Replaces RGL.forward (graph_model.py:99-130): embedding MLPs w_r/w_h, A = softmax(X w_a X^T),
 * num_layer x H' = relu(A H W_i) (+H), all inside one kernel, plus optionally the state-predictor
 * head (state_predictor.py:28,36).  Any non-null output is produced:
 *   H  [B,n,32]  final node features                      (RGL.forward return value)
 *   E  [B,32]    robot row H[:,0,:]                        (value_estimator.py:18)
 *   S  [B,Nh,5]  human_motion_predictor(H)[:,1:,:]         (needs motion_packed)
 *   A0 [n,n]     attention of state 0, first graph         (RGL.A, graph_model.py:116)
 * humans_bcast >= 1: state b reads humans[b / humans_bcast] (the planner evaluates many robot
 * actions against one human set; 1 = plain batch).  When only E is requested the last layer is
 * evaluated for the robot row only. */
int rgl_graph_forward(const float* robot, const float* humans, int B, int Nh, int humans_bcast,
                      const float* graph_packed, int num_layer, int flags,
                      const float* motion_packed,
                      float* H, float* E, float* S, float* A0,
                      rgl_stream_t stream);

/* Value head: V[B] = value_network(E[B,32]) (value_estimator.py:19). */
int rgl_value_head(const float* E, int B, const float* value_packed, float* V, rgl_stream_t stream);

/* Convenience: ValueEstimator.forward (value_estimator.py:11-20) = graph forward (E only) + value
 * head, two launches on `stream`; E_scratch [B,32] is caller-provided. */
int rgl_value_forward(const float* robot, const float* humans, int B, int Nh, int humans_bcast,
                      const float* graph_packed, int num_layer, int flags,
                      const float* value_packed, float* E_scratch, float* V, float* A0,
                      rgl_stream_t stream);

/* ---- stand-alone GCN layer ----------------------------------------------------------------------
 * One layer of graph_model.py:119-128 on node features already in HBM:
 *   Hout = relu((A X) W) (+ X if RGL_FLAG_SKIP), with A given ([B,n,n]) or, when A == NULL,
 *   computed in-kernel as softmax(X w_a X^T) (w_a [32,32] row-major, W [32,32] row-major,
 *   both exactly the nn.Parameter layout).  Aout (optional) receives the attention. */
int rgl_gcn_layer(const float* X, const float* A, const float* W, const float* w_a,
                  int B, int n, int flags, float* Hout, float* Aout, rgl_stream_t stream);

/* ---- batched look-ahead step (planner inner loop) -------------------------------------------------
 * For every (state e, action a): next robot state (state_predictor.py:41-60) and the reward estimate
 * (model_predictive_rl.py:304-357 + crowd_sim/envs/utils/utils.py:4-26), in one launch.
 * actions: DEVICE double [A,2] (model_predictive_rl.py:155-190 builds them in float64):
 *   RGL_KIN_HOLONOMIC (vx, vy)  ActionXY;   RGL_KIN_UNICYCLE (v, r)  ActionRot -- the next-state rule keeps the
 *   reference's indexing (the rotation is added to element 7 of the robot state, state_predictor.py:54).
 * Outputs: next_robot [E*A,1,9] fp32 (row e*A+a), reward [E*A] fp32.  State e reads humans[e / humans_bcast]
 * (look-ahead children of one parent share its predicted humans).  Reward arithmetic is float64 on the fp32 state values. */
#define RGL_KIN_HOLONOMIC 0
#define RGL_KIN_UNICYCLE  1
int rgl_plan_expand(const float* robot, const float* humans, int E, int Nh, int humans_bcast,
                    const double* actions, int A, double time_step, int kinematics,
                    float* next_robot, float* reward, rgl_stream_t stream);

/* value[e,a] = reward[e,a] + gamma_bar * V[e*A+a] (fp32, same op order as model_predictive_rl.py:227);
 * best[e] = first index of the maximum (strict '>' scan, model_predictive_rl.py:228-231), -1 if every value is NaN/-inf.
 * act_map (optional, DEVICE int32 [E,A]): the action each column stands for after action clipping;
 * best_action[e] = act_map[e,best[e]] (= best[e] without a map). */
int rgl_plan_argmax(const float* reward, const float* V, int E, int A, float gamma_bar,
                    float* value, int* best, const int* act_map, int* best_action, rgl_stream_t stream);

/* action_clip (model_predictive_rl.py:242-269) for E states at once: value = reward + gamma_bar * V as above, then the
 * `width` best actions per state -> acts [E,width] int32.  groups == NULL: descending value, ties by lower index;
 * groups (DEVICE int32 [A], action_group_index :168-181): sparse search (:252-263).  Also gathers what the next tree
 * level reads: child_reward [E,width] = reward[e,acts], child_robot [E*width,1,9] = next_robot[e*A+acts] (both optional);
 * value [E,A] optional. */
int rgl_plan_select(const float* reward, const float* V, int E, int A, float gamma_bar, int width, const int* groups,
                    const float* next_robot, int* acts, float* child_reward, float* child_robot, float* value,
                    rgl_stream_t stream);

/* V_planning backup (model_predictive_rl.py:293-298): ret[e,k] = v[e]/depth + (depth-1)/depth * (gamma_bar*next_v[e,k] + reward[e,k])
 * (fp32, every operation rounded as in the reference's tensor expression); ret_best[e] = max_k, best[e] = first argmax. */
int rgl_plan_backup(const float* v, const float* next_v, const float* reward, int E, int W, float gamma_bar, int depth,
                    float* ret_best, int* best, rgl_stream_t stream);


/* ---- training step (value estimator): forward with activation saves + backward building blocks -----------------
 * Replaces what autograd does under crowd_nav/utils/trainer.py:122-131 (loss.backward() through
 * ValueEstimator -> RGL).  The forward is the same fused kernels, additionally writing the activations the
 * backward needs; the backward is a short sequence of the three kernels below (see training.py). */
typedef struct RglGraphSave {          /* all state-major; NULL members are not written */
    float* a1r;                        /* [B,64]    relu(w_r.0 r)                              */
    float* a1h;                        /* [B,Nh,64] relu(w_h.0 h)                              */
    float* X;                          /* [B,n,32]  node embeddings                            */
    float* Y;                          /* [B,n,32]  X w_a                                      */
    float* A;                          /* [B,n,n]   attention                                  */
    float* M[RGL_MAX_LAYERS];          /* [B,n,32]  A H_{l-1}                                  */
    float* Rl[RGL_MAX_LAYERS];         /* [B,n,32]  relu(M_l W_l) (pre-skip)                   */
    float* Hl[RGL_MAX_LAYERS];         /* [B,n,32]  H_l                                        */
    float* mh;                         /* [B,Nh,64] relu(motion.0 H_L) (only with S; may be NULL) */
} RglGraphSave;

/* row r of a logical [R,width] matrix = ptr + (r / rows_per_group) * group_stride + (r % rows_per_group) * ld
 * (rows_per_group <= 1 and group_stride == 0: a plain matrix with leading dimension ld) */
typedef struct RglRows {
    float* ptr;
    int ld;
    int rows_per_group;
    long long group_stride;
} RglRows;

/* rgl_graph_forward with saves; layerwise graphs are not supported here (RGL_EUNSUPPORTED). H / E optional. */
int rgl_graph_forward_train(const float* robot, const float* humans, int B, int Nh,
                            const float* graph_packed, int num_layer, int flags,
                            const float* motion_packed, const RglGraphSave* save,
                            float* H, float* E, float* S, rgl_stream_t stream);
/* rgl_value_head with saves: v0 [B,32], v1 [B,128], v2 [B,128] (hidden activations, 100 wide zero-padded). */
int rgl_value_head_train(const float* E, int B, const float* value_packed, float* V,
                         float* v0, float* v1, float* v2, rgl_stream_t stream);
/* Backward of y = x W (+ b) over R rows: G [R,N] upstream gradient (multiplied by (mask > 0) when mask is given),
 * Xin [R,K] the layer input.  Any of: Gin [R,K] (= or +=) needs W; dW and db are ATOMICALLY ACCUMULATED (zero them
 * first).  w_layout 0: W/dW are [N][K] (nn.Linear.weight); 1: [K][N] (w_a, Ws).  N, K <= 128. */
int rgl_linear_bwd(const RglRows* G, int N, const RglRows* mask, const RglRows* Xin, int K,
                   const float* W, int w_layout, const RglRows* Gin, int accumulate,
                   float* dW, float* db, int R, rgl_stream_t stream);
/* Backward of BOTH Linear layers of a two-layer embedding MLP  x0 [R,K0] -> relu -> hidden [R,64] -> relu -> X [R,32]
 * (w_r / w_h, crowd_nav/policy/graph_model.py:41-42) in one launch; the hidden layer's gradient never reaches memory:
 *   G = gX . (mask > 0);  dW1 [32,64] += G^T hidden;  db1 [32] += colsum G;
 *   Hg = (G W1) . (hidden > 0);  dW0 [64,K0] += Hg^T x0;  db0 [64] += colsum Hg.        K0 <= 16; accumulators ATOMIC (zero first). */
int rgl_mlp2_bwd(const RglRows* G, const RglRows* mask, const RglRows* hidden, const float* W1, const RglRows* X0, int K0,
                 float* dW1, float* db1, float* dW0, float* db0, int R, rgl_stream_t stream);
/* gHprev[b,j,:] = (skip ? gH[b,j,:] : 0) + sum_i A[b,i,j] gM[b,i,:];  gA[b,i,j] (+)= gM[b,i,:] . Hprev[b,j,:]
 * mask (optional, [B,n,32]): gM is multiplied by (mask > 0) on load.
 * up_rows in [1,n]: only node rows i < up_rows of gM carry gradient; the others are taken as zero and never read
 * (1 for the top layer of the value step, whose head reads the robot row only: value_estimator.py:38; n otherwise). */
int rgl_attn_layer_bwd(const float* A, const float* Hprev, const float* gM, const float* gH, int skip,
                       float* gHprev, float* gA, int accumulate_gA, int B, int n, const float* mask, int up_rows,
                       rgl_stream_t stream);
/* Staged form of the two kernels around it for the RGL_FLAG_TRAIN_TC layer order (a CTA owns whole states and works out
 * of shared memory; one global round trip instead of n dependent ones):
 *   gZ[b,j,:] = sum_{i < up_rows} A[b,i,j] (gM[b,i,:] . (mask[b,i,:] > 0));   gA[b,i,j] = gA_in[b,i,j] (if given) + (gM.mask)[b,i,:] . Z[b,j,:]
 * Without X: gA is written to gA_out.  With X (layer 0): the similarity backward follows in the same kernel, gA never
 * leaves the chip:  gS = A (gA - rowsum(gA A)),  gY = gS X,  gX (gx_accumulate ? += : =) gS^T Y.   gX may alias gM. */
int rgl_attn_sim_bwd(const float* A, const float* Z, const float* gM, const float* mask, int up_rows, const float* gA_in,
                     float* gZ, float* gA_out, const float* X, const float* Y, float* gY, float* gX, int gx_accumulate,
                     int B, int n, rgl_stream_t stream);
/* softmax + similarity backward: gY = gS X, gX += gS^T Y with gS = A (gA - rowsum(gA A)) */
int rgl_sim_bwd(const float* A, const float* gA, const float* X, const float* Y, float* gY, float* gX,
                int B, int n, rgl_stream_t stream);

/* Temporal-difference loss of the value step (crowd_nav/utils/trainer.py:125-129) in one launch:
 *   target = reward + gamma_bar * V_next;  *loss += sum_b (V - target)^2 * inv_count  (zero *loss first; inv_count = 1 / global batch
 *   reproduces MSELoss(mean));  gV[b] = 2 (V - target) * inv_count (optional): the gradient loss.backward() hands to the value head. */
int rgl_td_loss(const float* V, const float* reward, const float* V_next, int B, float gamma_bar, float inv_count,
                float* loss, float* gV, rgl_stream_t stream);

/* ---- GPU-resident replay memory --------------------------------------------------------------------------------
 * crowd_nav/utils/memory.py:4-28 keeps python tuples of tiny tensors and torch's DataLoader collates 6 x batch_size of
 * them per minibatch (crowd_nav/utils/trainer.py:66-67,113-114).  Here a transition is one contiguous record of
 * rgl_replay_record_floats(Nh) = 20 + 10*Nh floats in a caller-owned device array `store` [capacity, rec]:
 *   [ robot 9 | humans 5*Nh | value 1 | reward 1 | next_robot 9 | next_humans 5*Nh ]
 * rgl_replay_push writes one transition (six DEVICE tensors) into `slot`; rgl_replay_gather builds a minibatch in one
 * launch: for b < B, record idx[b] (DEVICE int64) is scattered into robot[B,1,9], humans[B,Nh,5], value[B,1],
 * reward[B,1], next_robot[B,1,9], next_humans[B,Nh,5]. */
int rgl_replay_record_floats(int Nh);
int rgl_replay_push(float* store, long long slot, int Nh, const float* robot, const float* humans, const float* value,
                    const float* reward, const float* next_robot, const float* next_humans, rgl_stream_t stream);
int rgl_replay_gather(const float* store, const long long* idx, int B, int Nh, float* robot, float* humans, float* value,
                      float* reward, float* next_robot, float* next_humans, rgl_stream_t stream);

/* ---- data-parallel gradient exchange (one process per GPU, NVLink peer memory) ----------------------------------
 * Data-parallel form of crowd_nav/utils/trainer.py:122-131,143-149: every rank back-propagates its shard of the
 * minibatch, then ONE sum over a flat fp32 buffer holding every gradient (22 813 floats = 91 252 B for the value
 * estimator).  The reference has no multi-process code; this handle is the per-rank workspace the
 * boundary allows (rgl_*_create / _destroy).
 *   rgl_comm_create      allocates (cudaMalloc, zeroed) the local accumulation buffer, receive slots and flags;
 *   rgl_comm_ipc_handle  writes rgl_comm_handle_bytes() bytes that identify the allocation to other processes
 *                        (exchange them with any host-side all-gather, e.g. torch.distributed);
 *   rgl_comm_open_peers  maps every peer's allocation (handles = world x rgl_comm_handle_bytes(), rank order);
 *   rgl_comm_accum_ptr   DEVICE pointer of the local accumulation buffer [nfloats]: the backward kernels
 *                        (rgl_linear_bwd dW/db) accumulate straight into it;
 *   rgl_comm_allreduce   one kernel on `stream`: out[i] = scale * sum_over_ranks accum_r[i] (summed in rank order: the
 *                        same bits on every rank), then accum is re-zeroed.  No host synchronisation, CUDA-graph
 *                        capturable; every rank must launch it the same number of times;
 *   rgl_comm_status      0 = ok, 1 = a peer did not arrive within ~10 s (synchronises the device). */
typedef struct RglComm RglComm;
int    rgl_comm_create(int rank, int world, long long nfloats, RglComm** out);
int    rgl_comm_handle_bytes(void);
int    rgl_comm_ipc_handle(RglComm* c, void* handle_out);
int    rgl_comm_open_peers(RglComm* c, const void* handles);
float* rgl_comm_accum_ptr(RglComm* c);
int    rgl_comm_allreduce(RglComm* c, float* out, float scale, rgl_stream_t stream);
int    rgl_comm_status(RglComm* c, int* status);
int    rgl_comm_destroy(RglComm* c);
const char* rgl_comm_last_error_string(void);

#ifdef __cplusplus
}
#endif
#endif /* RGL_B200_H_ */
