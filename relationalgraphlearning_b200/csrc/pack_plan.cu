// Small helper kernels: weight packing (PyTorch state_dict layout -> k-major blobs) and the batched
// look-ahead step of the planner (next robot state + reward estimate + first-max argmax).
#include "kernels.h"

namespace rgl {

// dst[k*Np + o] = w[o*K + k] for o < N (PyTorch Linear weight [N,K]); columns N..Np-1 are zero-filled.
__device__ __forceinline__ void pack_linear_T(float* dst, const float* w, int N, int K, int Np, int t, int nt) {
    for (int idx = t; idx < K * Np; idx += nt) {
        const int k = idx / Np, o = idx - k * Np;
        dst[idx] = o < N ? w[o * K + k] : 0.f;
    }
}
__device__ __forceinline__ void pack_copy(float* dst, const float* src, int N, int Np, int t, int nt) {
    for (int idx = t; idx < Np; idx += nt) dst[idx] = idx < N ? src[idx] : 0.f;
}

// dst[k*Np + n] = src[k*N + n] (already k-major), padded columns zero
__device__ __forceinline__ void pack_rows(float* dst, const float* src, int K, int N, int Np, int t, int nt) {
    for (int idx = t; idx < K * Np; idx += nt) {
        const int k = idx / Np, n = idx - k * Np;
        dst[idx] = n < N ? src[k * N + n] : 0.f;
    }
}

// ---- tensor-core operand tiles (common.cuh "tensor-core operand sections") ----------------------------------
__device__ __forceinline__ float rna_tf32f(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// element (row, k) of a [rows][32] K-major SWIZZLE_128B tile; hi tile at `tile`, lo tile at `tile + lo_off`
__device__ __forceinline__ void tc_put(float* tile, int lo_off, int row, int k, float w) {
    const int idx = row * 32 + ((((k >> 2) ^ row) & 7) << 2) + (k & 3);
    const float hi = rna_tf32f(w);
    tile[idx] = hi;
    tile[lo_off + idx] = rna_tf32f(w - hi);
}

// One section per CTA (group): the sections are independent gathers of a few hundred to a few thousand floats, each ONE global
// round trip deep.  The first version ran them one after the other in every thread (12 dependent round trips, 8 us per
// launch at the head of every training step, where the weights change and the blob is re-packed); now they run side by side.
__global__ void pack_graph_kernel(RglGraphParams p, float* out) {
    float* tc = out + graph_tc_off(p.num_layer);
    int sec = blockIdx.x, sub = 0, nsub = 1;
    // sections 1 (embedding layer 2 tiles), 6 and 8 (k-major embedding layer 2) are split over several CTAs
    if (sec >= 1 && sec < 5) { sub = sec - 1; nsub = 4; sec = 1; }
    else if (sec >= 5) {
        sec -= 3;                                             // 2, 3, 4, 5, [6, 6], 7, [8, 8], 9, 10 + l
        if (sec == 6 || sec == 7) { sub = sec - 6; nsub = 2; sec = 6; }
        else if (sec == 8) sec = 7;
        else if (sec == 9 || sec == 10) { sub = sec - 9; nsub = 2; sec = 8; }
        else if (sec > 10) sec -= 2;
    }
    const int t = sub * blockDim.x + threadIdx.x, nt = nsub * blockDim.x;
    switch (sec) {
    case 0:
        for (int idx = t; idx < HID * 16; idx += nt) {          // emb layer 1: k-concatenated robot | human | biases; hi at k, lo at k + 16
            const int u = idx >> 4, k = idx & 15;
            const float w = k < RD ? p.wr0_w[u * RD + k] : k < RD + HD ? p.wh0_w[u * HD + (k - RD)] : k == 14 ? p.wr0_b[u] : p.wh0_b[u];
            const float hi = rna_tf32f(w);
            tc[T_W0 + u * 32 + ((((k >> 2) ^ u) & 7) << 2) + (k & 3)] = hi;
            tc[T_W0 + u * 32 + (((((k + 16) >> 2) ^ u) & 7) << 2) + (k & 3)] = rna_tf32f(w - hi);
        }
        break;
    case 1:
        for (int idx = t; idx < 64 * HID; idx += nt) {          // emb layer 2: n-stacked human (rows 0-31) | robot (rows 32-63)
            const int row = idx / HID, k = idx - row * HID;
            const float w = row < XD ? p.wh1_w[row * HID + k] : p.wr1_w[(row - XD) * HID + k];
            tc_put(tc + T_W1 + (k >> 5) * 2048, 4096, row, k & 31, w);
        }
        break;
    case 2:
        for (int idx = t; idx < XD * XD; idx += nt) {           // X @ W uses W[k][n]: B operand row n holds column n
            const int nn = idx >> 5, k = idx & 31;
            tc_put(tc + T_WA, 2048, nn, k, p.w_a[k * XD + nn]);
            tc_put(tc + T_WA, 2048, XD + nn, k, p.Ws[0][k * XD + nn]);
            for (int l = 1; l < p.num_layer; ++l) tc_put(tc + T_WS1 + (l - 1) * 2048, 1024, nn, k, p.Ws[l][k * XD + nn]);
        }
        break;
    case 3: {
        float* b = tc + tc_bias_off(p.num_layer);
        for (int idx = t; idx < 256; idx += nt)
            b[idx] = idx < XD ? p.wh1_b[idx] : (idx >= TC_RBIAS && idx < TC_RBIAS + XD) ? p.wr1_b[idx - TC_RBIAS] : 0.f;
        break;
    }
    case 4:                                                     // the four bias vectors of the k-major (fp32-FMA) section
        pack_copy(out + G_BR0, p.wr0_b, HID, HID, t, nt);
        pack_copy(out + G_BR1, p.wr1_b, XD, XD, t, nt);
        pack_copy(out + G_BH0, p.wh0_b, HID, HID, t, nt);
        pack_copy(out + G_BH1, p.wh1_b, XD, XD, t, nt);
        break;
    case 5: pack_linear_T(out + G_WR0, p.wr0_w, HID, RD, HID, t, nt); break;
    case 6: pack_linear_T(out + G_WR1, p.wr1_w, XD, HID, LDW, t, nt); break;
    case 7: pack_linear_T(out + G_WH0, p.wh0_w, HID, HD, HID, t, nt); break;
    case 8: pack_linear_T(out + G_WH1, p.wh1_w, XD, HID, LDW, t, nt); break;
    case 9: pack_rows(out + G_WA, p.w_a, XD, XD, LDW, t, nt); break;
    default: {
        const int l = sec - 10;
        if (l < p.num_layer) pack_rows(out + G_WS + l * XD * LDW, p.Ws[l], XD, XD, LDW, t, nt);
    }
    }
}

__global__ void pack_value_kernel(RglValueParams p, float* out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    pack_linear_T(out + V_W0, p.w0, XD, XD, XD, t, nt);
    pack_copy(out + V_B0, p.b0, XD, XD, t, nt);
    pack_linear_T(out + V_W1, p.w1, VH, XD, VHP, t, nt);
    pack_copy(out + V_B1, p.b1, VH, VHP, t, nt);
    pack_linear_T(out + V_W2, p.w2, VH, VH, VHP, t, nt);
    pack_copy(out + V_B2, p.b2, VH, VHP, t, nt);
    pack_copy(out + V_W3, p.w3, VH, VHP, t, nt);
    pack_copy(out + V_B3, p.b3, 1, 4, t, nt);
    // tensor-core section (common.cuh TV_*): Linear weights are [out,in] = B-operand rows already
    float* tc = out + VALUE_TC_OFF;
    for (int idx = t; idx < XD * XD; idx += nt) tc_put(tc + TV_W0, 1024, idx >> 5, idx & 31, p.w0[idx]);
    for (int idx = t; idx < TV_NP * 32; idx += nt) {
        const int r = idx >> 5, k = idx & 31;
        tc_put(tc + TV_W1, TV_NP * 32, r, k, r < VH ? p.w1[r * XD + k] : 0.f);
        for (int a = 0; a < 4; ++a) {
            const int kk = 32 * a + k;
            tc_put(tc + TV_W2 + a * TV_NP * 32, 4 * TV_NP * 32, r, k, (r < VH && kk < VH) ? p.w2[r * VH + kk] : 0.f);
        }
    }
    float* b = tc + TV_BIAS;
    for (int idx = t; idx < 512; idx += nt) {
        float v = 0.f;
        if (idx < 32) v = p.b0[idx];
        else if (idx < 160) v = idx - 32 < VH ? p.b1[idx - 32] : 0.f;
        else if (idx < 288) v = idx - 160 < VH ? p.b2[idx - 160] : 0.f;
        else if (idx < 416) v = idx - 288 < VH ? p.w3[idx - 288] : 0.f;
        else if (idx == 416) v = p.b3[0];
        b[idx] = v;
    }
}

__global__ void pack_motion_kernel(RglMotionParams p, float* out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    pack_linear_T(out + M_W0, p.w0, MH, XD, MH, t, nt);
    pack_copy(out + M_B0, p.b0, MH, MH, t, nt);
    pack_copy(out + M_W1, p.w1, HD * MH, HD * MH, t, nt);
    pack_copy(out + M_B1, p.b1, HD, 8, t, nt);
    float* tc = out + MOTION_TC_OFF;
    for (int idx = t; idx < MH * XD; idx += nt) tc_put(tc + TM_W0, 2048, idx >> 5, idx & 31, p.w0[idx]);
    pack_copy(tc + TM_B0, p.b0, MH, MH, t, nt);
    pack_copy(tc + TM_W1, p.w1, HD * MH, HD * MH, t, nt);
    pack_copy(tc + TM_B1, p.b1, HD, 8, t, nt);
}

// ---- planner: one thread per (state e, action a) ------------------------------------------------------
// next robot state: crowd_nav/policy/state_predictor.py:48-52 (holonomic: fp32 position + fp32(v*dt), velocity = action)
//                   and :53-58 (unicycle: the rotation is added to element 7 exactly as the reference does -- element 7
//                   is v_pref, the heading is element 8; SURVEY.md 5 -- then cos / sin of that element steer the step)
// reward:           crowd_nav/policy/model_predictive_rl.py:304-357, crowd_sim/envs/utils/utils.py:4-26, in float64
// actions: (vx, vy) holonomic | (v, r) unicycle
__global__ void plan_expand_kernel(const float* __restrict__ robot, const float* __restrict__ humans, int E, int Nh, int hb,
                                   const double* __restrict__ actions, int A, double dt, int unicycle,
                                   float* __restrict__ next_robot, float* __restrict__ reward) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();
    pdl_trigger();
    if (idx >= E * A) return;
    const int e = idx / A, a = idx - e * A;
    const float* r = robot + (size_t)e * RD;
    const double a0 = actions[2 * a], a1 = actions[2 * a + 1];
    if (next_robot) {
        float* o = next_robot + (size_t)idx * RD;
#pragma unroll
        for (int c = 4; c < RD; ++c) o[c] = r[c];
        if (!unicycle) {
            o[0] = __fadd_rn(r[0], (float)(a0 * dt));
            o[1] = __fadd_rn(r[1], (float)(a1 * dt));
            o[2] = (float)a0;
            o[3] = (float)a1;
        } else {
            const float h7 = __fadd_rn(r[7], (float)a1);
            const float c = cosf(h7), sn = sinf(h7);
            const float cv = __fmul_rn(c, (float)a0), sv = __fmul_rn(sn, (float)a0);
            o[7] = h7;
            o[0] = __fadd_rn(r[0], __fmul_rn(cv, (float)dt));
            o[1] = __fadd_rn(r[1], __fmul_rn(sv, (float)dt));
            o[2] = cv;
            o[3] = sv;
        }
    }
    if (reward) {
        const double rpx = r[0], rpy = r[1], rrad = r[4], gx = r[5], gy = r[6];
        // velocity of the robot under this action, world frame (unicycle: heading = theta + r, model_predictive_rl.py:319-321,337-340)
        double avx = a0, avy = a1;
        if (unicycle) {
            const double th = __dadd_rn(a1, (double)r[8]);
            avx = __dmul_rn(a0, cos(th));
            avy = __dmul_rn(a0, sin(th));
        }
        double dmin = INFINITY;
        bool collision = false;
        const float* h = humans + (size_t)(e / hb) * Nh * HD;
        for (int j = 0; j < Nh; ++j, h += HD) {
            const double px = (double)h[0] - rpx, py = (double)h[1] - rpy;
            const double vx = (double)h[2] - avx, vy = (double)h[3] - avy;
            const double ex = __dadd_rn(px, __dmul_rn(vx, dt)), ey = __dadd_rn(py, __dmul_rn(vy, dt));
            // distance from the origin to the segment (px,py)-(ex,ey)
            const double sx = ex - px, sy = ey - py;
            double d;
            if (sx == 0.0 && sy == 0.0) {
                d = sqrt(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)));
            } else {
                double u = __dadd_rn(__dmul_rn(-px, sx), __dmul_rn(-py, sy)) / __dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy));
                u = u > 1.0 ? 1.0 : (u < 0.0 ? 0.0 : u);
                const double x = __dadd_rn(px, __dmul_rn(u, sx)), y = __dadd_rn(py, __dmul_rn(u, sy));
                d = sqrt(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
            }
            d = d - (double)h[4] - rrad;
            if (d < 0.0) { collision = true; break; }
            if (d < dmin) dmin = d;
        }
        const double qx = __dadd_rn(rpx, __dmul_rn(avx, dt)) - gx, qy = __dadd_rn(rpy, __dmul_rn(avy, dt)) - gy;
        const bool reaching = sqrt(__dadd_rn(__dmul_rn(qx, qx), __dmul_rn(qy, qy))) < rrad;
        double rew;
        if (collision) rew = -0.25;
        else if (reaching) rew = 1.0;
        else if (dmin < 0.2) rew = __dmul_rn(__dmul_rn(dmin - 0.2, 0.5), dt);
        else rew = 0.0;
        reward[idx] = (float)rew;
    }
}

// value = fp32(reward) + fp32(gamma_bar) * V (two rounded fp32 ops, like the tensor expression in
// model_predictive_rl.py:227); best = first index of the maximum under a strict '>' scan (:228-231).
// act_map (optional, [E,A] int32): the actions the A columns stand for (after action clipping); best_action[e] =
// act_map[e, best[e]] (or best[e] without a map; -1 if no finite value).
__global__ void plan_argmax_kernel(const float* __restrict__ reward, const float* __restrict__ V, int E, int A, float gamma_bar,
                                   float* __restrict__ value, int* __restrict__ best, const int* __restrict__ act_map,
                                   int* __restrict__ best_action) {
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    pdl_wait();
    pdl_trigger();
    if (e >= E) return;
    float bv = -INFINITY;
    int bi = -1;
    for (int a = lane; a < A; a += 32) {
        const float v = __fadd_rn(reward[(size_t)e * A + a], __fmul_rn(gamma_bar, V[(size_t)e * A + a]));
        if (value) value[(size_t)e * A + a] = v;
        if (v > bv) { bv = v; bi = a; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
        if (best) best[e] = bi;
        if (best_action) best_action[e] = (bi >= 0 && act_map) ? act_map[(size_t)e * A + bi] : bi;
    }
}

// action_clip (model_predictive_rl.py:242-269) for one state per warp: value[a] = reward + gamma_bar * V (as above), then
// the `width` best actions --
//   groups == nullptr : descending value, ties by LOWER action index (the reference's argpartition order is unspecified;
//                       oracle/planner_oracle.py pins this one);
//   groups given      : sparse search (:252-263): walk by descending value, ties by HIGHER index (argsort()[::-1]), keep
//                       the first action of every not-yet-seen group.
// NaN values are never selected before any finite / infinite one.  Besides the kept indices acts[e,k] the kernel gathers
// what the next tree level consumes: child_rew[e,k] = reward[e,acts] and child_robot[e*width+k] = next_robot[e*A+acts].
constexpr int SEL_MAX_PER_LANE = 8;      // A <= 256
__global__ void plan_select_kernel(const float* __restrict__ reward, const float* __restrict__ V, int E, int A, float gamma_bar,
                                   int width, const int* __restrict__ groups, const float* __restrict__ next_robot,
                                   int* __restrict__ acts, float* __restrict__ child_rew, float* __restrict__ child_robot,
                                   float* __restrict__ value) {
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    pdl_wait();
    pdl_trigger();
    if (e >= E) return;
    float v[SEL_MAX_PER_LANE];
    unsigned taken = 0;                   // bit i: this lane's i-th action is no longer a candidate
#pragma unroll
    for (int i = 0; i < SEL_MAX_PER_LANE; ++i) {
        const int a = lane + 32 * i;
        v[i] = -INFINITY;
        if (a < A) {
            v[i] = __fadd_rn(reward[(size_t)e * A + a], __fmul_rn(gamma_bar, V[(size_t)e * A + a]));
            if (value) value[(size_t)e * A + a] = v[i];
        } else {
            taken |= 1u << i;
        }
    }
    for (int k = 0; k < width; ++k) {
        // this lane's best remaining candidate (NaN ranks below everything)
        float bv = 0.f;
        int bi = -1;
        bool bnan = true;
#pragma unroll
        for (int i = 0; i < SEL_MAX_PER_LANE; ++i) {
            if (taken >> i & 1u) continue;
            const int a = lane + 32 * i;
            const bool isn = v[i] != v[i];
            bool better;
            if (bi < 0) better = true;
            else if (isn != bnan) better = bnan;                                   // finite beats NaN
            else if (isn) better = groups ? a > bi : a < bi;
            else better = v[i] > bv || (v[i] == bv && (groups ? a > bi : a < bi));
            if (better) { bv = v[i]; bi = a; bnan = isn; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            const bool on = __shfl_xor_sync(0xffffffffu, (int)bnan, off) != 0;
            bool better;
            if (oi < 0) better = false;
            else if (bi < 0) better = true;
            else if (on != bnan) better = bnan;
            else if (on) better = groups ? oi > bi : oi < bi;
            else better = ov > bv || (ov == bv && (groups ? oi > bi : oi < bi));
            if (better) { bv = ov; bi = oi; bnan = on; }
        }
        // bi is warp-uniform now (every lane applied the same total order); -1 only if fewer candidates than width remain
        if (bi >= 0) {
            const int g = groups ? groups[bi] : 0;
#pragma unroll
            for (int i = 0; i < SEL_MAX_PER_LANE; ++i) {
                const int a = lane + 32 * i;
                if (a == bi || (groups && a < A && groups[a] == g)) taken |= 1u << i;
            }
        }
        const int sel = bi < 0 ? 0 : bi;
        if (lane == 0) {
            acts[(size_t)e * width + k] = sel;
            if (child_rew) child_rew[(size_t)e * width + k] = reward[(size_t)e * A + sel];
        }
        if (child_robot && lane < RD)
            child_robot[((size_t)e * width + k) * RD + lane] = next_robot[((size_t)e * A + sel) * RD + lane];
    }
}

// V_planning backup (model_predictive_rl.py:293,298): ret[e,k] = v[e] / depth + (depth-1)/depth * (gamma_bar * nv[e,k] + rew[e,k])
// evaluated as the reference's fp32 tensor expression (every operation rounded separately); best = first maximum.
__global__ void plan_backup_kernel(const float* __restrict__ v, const float* __restrict__ nv, const float* __restrict__ rew, int E,
                                   int W, float gamma_bar, float depth, float frac, float* __restrict__ ret_best,
                                   int* __restrict__ best) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();
    pdl_trigger();
    if (e >= E) return;
    const float base = __fdiv_rn(v[e], depth);
    float bv = 0.f;
    int bi = -1;
    for (int k = 0; k < W; ++k) {
        const float t = __fadd_rn(__fmul_rn(gamma_bar, nv[(size_t)e * W + k]), rew[(size_t)e * W + k]);
        const float r = __fadd_rn(base, __fmul_rn(frac, t));
        if (bi < 0 || (bv == bv && (r > bv || r != r))) { bv = r; bi = k; }     // np.argmax: first maximum, a NaN counts as the maximum
    }
    ret_best[e] = bv;
    best[e] = bi;
}

cudaError_t run_pack_graph(const RglGraphParams& p, float* out, cudaStream_t st) {
    pack_graph_kernel<<<15 + p.num_layer, 256, 0, st>>>(p, out);      // 15 section CTAs + one per GCN layer
    return cudaGetLastError();
}
cudaError_t run_pack_value(const RglValueParams& p, float* out, cudaStream_t st) {
    pack_value_kernel<<<16, 256, 0, st>>>(p, out);
    return cudaGetLastError();
}
cudaError_t run_pack_motion(const RglMotionParams& p, float* out, cudaStream_t st) {
    pack_motion_kernel<<<4, 256, 0, st>>>(p, out);
    return cudaGetLastError();
}
cudaError_t run_plan_expand(const float* robot, const float* humans, int E, int Nh, int hb, const double* actions, int A, double dt,
                            int unicycle, float* next_robot, float* reward, cudaStream_t st) {
    const int total = E * A;
    return launch_pdl(plan_expand_kernel, dim3((total + 127) / 128), dim3(128), 0, st, robot, humans, E, Nh, hb, actions, A, dt, unicycle, next_robot, reward);
}
cudaError_t run_plan_argmax(const float* reward, const float* V, int E, int A, float gamma_bar, float* value, int* best,
                            const int* act_map, int* best_action, cudaStream_t st) {
    return launch_pdl(plan_argmax_kernel, dim3((E + 3) / 4), dim3(128), 0, st, reward, V, E, A, gamma_bar, value, best, act_map, best_action);
}
cudaError_t run_plan_select(const float* reward, const float* V, int E, int A, float gamma_bar, int width, const int* groups,
                            const float* next_robot, int* acts, float* child_rew, float* child_robot, float* value, cudaStream_t st) {
    if (A > 32 * SEL_MAX_PER_LANE) return cudaErrorInvalidConfiguration;
    return launch_pdl(plan_select_kernel, dim3((E + 3) / 4), dim3(128), 0, st, reward, V, E, A, gamma_bar, width, groups, next_robot, acts, child_rew, child_robot, value);
}
cudaError_t run_plan_backup(const float* v, const float* nv, const float* rew, int E, int W, float gamma_bar, int depth,
                            float* ret_best, int* best, cudaStream_t st) {
    const float frac = (float)((double)(depth - 1) / (double)depth);
    return launch_pdl(plan_backup_kernel, dim3((E + 127) / 128), dim3(128), 0, st, v, nv, rew, E, W, gamma_bar, (float)depth, frac, ret_best, best);
}

}  // namespace rgl
