"""CPU oracle for the model-predictive planner tree -- TEST INFRASTRUCTURE ONLY (see rgl_oracle.py).

Restates, one batch-1 module forward at a time exactly like the reference does,

  build_action_space   crowd_nav/policy/model_predictive_rl.py:155-190
  predict              crowd_nav/policy/model_predictive_rl.py:192-240   (greedy branch)
  action_clip          crowd_nav/policy/model_predictive_rl.py:242-269
  V_planning           crowd_nav/policy/model_predictive_rl.py:271-302
  estimate_reward      crowd_nav/policy/model_predictive_rl.py:304-357
  point_to_segment     crowd_sim/envs/utils/utils.py:4-26
  tensor<->state       crowd_sim/envs/utils/state.py:64-92

The reference planner cannot run at depth>1 on this container's torch/numpy (SURVEY.md §5:
`np.array` over grad-carrying [1,1] tensors raises, `argpartition` misbehaves), so values are
scalarised with float() and the *intended* semantics are restated:
  * action_clip keeps the `width` best actions; order = descending value, ties by ascending
    action index (the reference's argpartition order is unspecified);
  * sparse search walks actions by descending value (ties: higher index first, mirroring
    `argsort()[::-1]`) and keeps the first action of each not-yet-seen group;
  * every argmax is "first maximum wins" (strict `>` in predict(), np.argmax in V_planning()).
Depth-1 behaviour is pinned against the reference's own predict()/estimate_reward() run in this
container (oracle/gen_golden.py -> tests/golden/planner_*.npz).

Reward arithmetic: float64 on the float32 state values (NumPy<2 value-based promotion gives
exactly this; under NEP-50 NumPy the reference evaluates look-ahead rewards in float32, a <=1e-6
absolute difference that the tests allow for).
"""
import math

import numpy as np
import torch

from oracle import rgl_oracle as O


def build_action_space(v_pref, speed_samples=5, rotation_samples=16, kinematics='holonomic',
                       sparse_rotation_samples=8, rotation_constraint=np.pi / 3):
    """Returns (actions float64 [A,2] = (vx, vy) holonomic | (v, r) unicycle, action_group_index list), speed-major."""
    holonomic = kinematics == 'holonomic'
    speeds = [(np.exp((i + 1) / speed_samples) - 1) / (np.e - 1) * v_pref for i in range(speed_samples)]
    if holonomic:
        rotations = np.linspace(0, 2 * np.pi, rotation_samples, endpoint=False)
    else:
        rotations = np.linspace(-rotation_constraint, rotation_constraint, rotation_samples)
    actions = [(0.0, 0.0)]
    groups = [0]
    for j, speed in enumerate(speeds):
        sg = 0 if j < 3 else 1
        for i, rot in enumerate(rotations):
            groups.append(sg * sparse_rotation_samples + i // 2)
            actions.append((speed * np.cos(rot), speed * np.sin(rot)) if holonomic else (speed, rot))
    return np.asarray(actions, dtype=np.float64), groups


def point_to_segment_dist(x1, y1, x2, y2, x3, y3):
    px = x2 - x1
    py = y2 - y1
    if px == 0 and py == 0:
        return float(np.linalg.norm((x3 - x1, y3 - y1)))
    u = ((x3 - x1) * px + (y3 - y1) * py) / (px * px + py * py)
    if u > 1:
        u = 1
    elif u < 0:
        u = 0
    x = x1 + u * px
    y = y1 + u * py
    return float(np.linalg.norm((x - x3, y - y3)))   # same library call as utils.py:26 (dot + sqrt)


def estimate_reward(robot, humans, action, time_step, kinematics='holonomic'):
    """robot: 9 floats, humans: [Nh][5] floats, action (vx, vy) | unicycle (v, r).  Returns python float."""
    rpx, rpy, _, _, rrad, gx, gy = [float(v) for v in robot[:7]]
    if kinematics == 'holonomic':
        avx, avy = float(action[0]), float(action[1])
    else:       # model_predictive_rl.py:319-321,337-340: heading = action.r + theta (theta = element 8)
        th = float(action[1]) + float(robot[8])
        avx, avy = float(action[0]) * np.cos(th), float(action[0]) * np.sin(th)
    dmin = float('inf')
    collision = False
    for h in humans:
        hpx, hpy, hvx, hvy, hrad = [float(v) for v in h[:5]]
        px = hpx - rpx
        py = hpy - rpy
        vx = hvx - avx
        vy = hvy - avy
        ex = px + vx * time_step
        ey = py + vy * time_step
        d = point_to_segment_dist(px, py, ex, ey, 0.0, 0.0) - hrad - rrad
        if d < 0:
            collision = True
            break
        elif d < dmin:
            dmin = d
    ex = rpx + avx * time_step
    ey = rpy + avy * time_step
    reaching_goal = np.linalg.norm(np.array((ex, ey)) - np.array([gx, gy])) < rrad
    if collision:
        return -0.25
    if reaching_goal:
        return 1
    if dmin < 0.2:
        return (dmin - 0.2) * 0.5 * time_step
    return 0


class OraclePlanner(object):
    """Batch-1, loop-per-action planner restating ModelPredictiveRL's greedy branch."""

    def __init__(self, graph_sd_v, value_sd, graph_sd_s, motion_sd, gamma=0.9, time_step=0.25, v_pref=1.0,
                 planning_depth=1, planning_width=1, do_action_clip=False, sparse_search=False,
                 speed_samples=5, rotation_samples=16, linear_state_predictor=False, graph_kw=None, kinematics='holonomic'):
        self.gv, self.vn, self.gs, self.mp = graph_sd_v, value_sd, graph_sd_s, motion_sd
        self.gamma, self.time_step, self.v_pref = gamma, time_step, v_pref
        self.depth, self.width = planning_depth, planning_width
        self.do_action_clip, self.sparse_search = do_action_clip, sparse_search
        self.linear = linear_state_predictor
        self.graph_kw = graph_kw or {}
        self.kinematics = kinematics
        self.actions, self.groups = build_action_space(v_pref, speed_samples, rotation_samples, kinematics)
        self.n_value_fwd = 0
        self.n_sp_fwd = 0

    def gamma_bar(self):
        return pow(self.gamma, self.time_step * self.v_pref)

    # --- the two module forwards, batch 1 ---------------------------------------------------
    def V(self, state):
        self.n_value_fwd += 1
        with torch.no_grad():
            return O.value_forward(self.gv, self.vn, state[0], state[1], **self.graph_kw)

    def SP(self, state, a):
        self.n_sp_fwd += 1
        vx, vy = self.actions[a]
        with torch.no_grad():
            nr = O.next_robot_state(state[0], vx, vy, self.time_step, self.kinematics)
            if self.linear:
                nh = O.linear_motion(state[1])
            else:
                nh = O.statepred_forward(self.gs, self.mp, state[0], state[1], **self.graph_kw)
        return (nr, nh)

    def R(self, state, a):
        robot = state[0].reshape(-1).numpy().astype(np.float64)
        humans = state[1].reshape(-1, 5).numpy().astype(np.float64)
        return estimate_reward(robot, humans, self.actions[a], self.time_step, self.kinematics)

    # --- tree -------------------------------------------------------------------------------
    def action_clip(self, state, width):
        vals = []
        for a in range(len(self.actions)):
            nxt = self.SP(state, a)
            ret = self.V(nxt)
            vals.append(float(self.R(state, a) + self.gamma_bar() * ret))
        vals = np.asarray(vals)
        if self.sparse_search:
            order = sorted(range(len(vals)), key=lambda i: (-vals[i], -i))
            seen, out = set(), []
            for i in order:
                if self.groups[i] not in seen:
                    out.append(i)
                    seen.add(self.groups[i])
                    if len(out) == width:
                        break
            return out, vals
        order = sorted(range(len(vals)), key=lambda i: (-vals[i], i))
        return order[:width], vals

    def V_planning(self, state, depth, width):
        v = self.V(state)
        if depth == 1:
            return v
        acts = self.action_clip(state, width)[0] if self.do_action_clip else list(range(len(self.actions)))
        best = None
        for a in acts:
            nxt = self.SP(state, a)
            r = self.R(state, a)
            nv = self.V_planning(nxt, depth - 1, self.width)
            ret = v / depth + (depth - 1) / depth * (self.gamma_bar() * nv + r)
            if best is None or float(ret) > float(best):
                best = ret
        return best

    def predict(self, robot, humans):
        """robot[1,1,9], humans[1,Nh,5] fp32.  Returns (best action index, best value, per-action values dict)."""
        state = (robot, humans)
        r = robot.reshape(-1).numpy().astype(np.float64)
        if np.linalg.norm((r[1] - r[6], r[0] - r[5])) < r[4]:
            return 0, None, {}      # reach_destination -> stop action (crowd_sim/envs/policy/policy.py:59-65)
        acts = self.action_clip(state, self.width)[0] if self.do_action_clip else list(range(len(self.actions)))
        best_a, best_v, table = None, float('-inf'), {}
        for a in acts:
            nxt = self.SP(state, a)
            ret = self.V_planning(nxt, self.depth, self.width)
            value = self.R(state, a) + self.gamma_bar() * ret
            table[a] = float(value)
            if float(value) > best_v:
                best_v, best_a = float(value), a
        return best_a, best_v, table
