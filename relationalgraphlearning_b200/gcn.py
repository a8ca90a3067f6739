"""GCN model-free policy -- drop-in for crowd_nav/policy/gcn.py:11-159 (`ValueNetwork`, `GCN`) with the
`MultiHumanRL` / `CADRL` machinery it inherits (crowd_nav/policy/multi_human_rl.py:12-130, crowd_nav/policy/cadrl.py:35-276).
SURVEY.md 8(f3): the same similarity + GCN math as the RGL graph model, on the agent-centric ("rotated") pairwise input.

`ValueNetwork` keeps the reference's constructor, parameter names (`w_r.*`, `w_h.*`, `w_a`, `w1`, `w2`, `value_net.*`)
and `forward(state[B,Nh,13]) -> value[B,1]`.  On CUDA, for the shipped configuration (embedded-gaussian similarity,
X_dim = 32, one or two layers), each GCN layer is ONE launch of the stand-alone layer kernel `rgl_gcn_layer`
(tcgen05 + TMA tensor copies, csrc/gcn_layer_tc.cu; layer 1 computes the attention in-kernel and returns it, layer 2 reuses
it); the two embedding MLPs (6 -> 64 -> 32, 7 -> 64 -> 32) and the 150-wide planning head are plain library GEMMs
(torch / cuBLAS).  Training, other similarity functions and CPU modules use the torch-op statement of the same math.

`GCN.predict` evaluates ALL actions in one batch (the reference loops over the 81 actions with one batch-1 model forward
each, multi_human_rl.py:39-65): propagate, reward, rotate and the value network run once on [A, Nh, .] tensors.
"""
import itertools
import logging

import numpy as np
import torch
import torch.nn as nn
from torch.nn import Parameter

from . import ops
from . import _torch_math as TM
from .helpers import mlp
from .simtypes import ActionRot, ActionXY, Policy


class ValueNetwork(nn.Module):
    def __init__(self, input_dim, self_state_dim, num_layer, X_dim, wr_dims, wh_dims, final_state_dim,
                 gcn2_w1_dim, planning_dims, similarity_function, layerwise_graph, skip_connection):
        super().__init__()
        self.similarity_function = similarity_function
        logging.info('self.similarity_func: {}'.format(self.similarity_function))
        human_state_dim = input_dim - self_state_dim
        self.self_state_dim = self_state_dim
        self.human_state_dim = human_state_dim
        self.num_layer = num_layer
        self.X_dim = X_dim
        self.layerwise_graph = layerwise_graph
        self.skip_connection = skip_connection
        self._dims = (list(wr_dims), list(wh_dims), final_state_dim, gcn2_w1_dim)

        # parameter creation order = the reference's (gcn.py:28-44): identical weights under the same seed
        self.w_r = mlp(self_state_dim, list(wr_dims), last_relu=True)
        self.w_h = mlp(human_state_dim, list(wh_dims), last_relu=True)
        if self.similarity_function == 'embedded_gaussian':
            self.w_a = Parameter(torch.randn(self.X_dim, self.X_dim))
        elif self.similarity_function == 'concatenation':
            self.w_a = mlp(2 * X_dim, [2 * X_dim, 1], last_relu=True)
        if num_layer == 1:
            self.w1 = Parameter(torch.randn(self.X_dim, final_state_dim))
        elif num_layer == 2:
            self.w1 = Parameter(torch.randn(self.X_dim, gcn2_w1_dim))
            self.w2 = Parameter(torch.randn(gcn2_w1_dim, final_state_dim))
        else:
            raise NotImplementedError
        self.value_net = mlp(final_state_dim, list(planning_dims))
        self._A_dev = None
        self._A_host = None

    # `.A`: attention of sample 0 (gcn.py:113); copied to the host when it is read
    @property
    def A(self):
        if self._A_host is None and self._A_dev is not None:
            self._A_host = self._A_dev.detach().cpu().numpy()
        return self._A_host

    @A.setter
    def A(self, value):
        self._A_host, self._A_dev = value, None

    def kernel_supported(self):
        wr, wh, fd, w1d = self._dims
        return (self.similarity_function == 'embedded_gaussian' and self.X_dim == 32 and wr[-1] == 32 and wh[-1] == 32
                and fd == 32 and (self.num_layer == 1 or w1d == 32))

    def _similarity(self, X):
        return TM.similarity(X, getattr(self, 'w_a', None), self.similarity_function)

    def _torch_forward(self, X):
        """gcn.py:115-140 as torch ops (training / unsupported configurations / CPU)."""
        A = self._similarity(X)
        self._A_dev, self._A_host = A[0].detach(), None
        h1 = torch.relu(torch.matmul(torch.matmul(A, X), self.w1))
        if self.num_layer == 1:
            return h1[:, 0, :]
        if self.skip_connection:
            h1 = h1 + X
        A2 = self._similarity(h1) if self.layerwise_graph else A
        h2 = torch.relu(torch.matmul(torch.matmul(A2, h1), self.w2))
        if self.skip_connection:
            h2 = h2 + h1
        return h2[:, 0, :]

    def forward(self, state_input):
        state = state_input[0] if isinstance(state_input, tuple) else state_input
        self_state = state[:, 0, :self.self_state_dim]
        human_states = state[:, :, self.self_state_dim:]
        X = torch.cat([self.w_r(self_state).unsqueeze(1), self.w_h(human_states)], dim=1)
        n = X.size(1)
        native = (X.is_cuda and self.kernel_supported() and 2 <= n <= 32 and X.size(0) > 0
                  and not ops._needs_grad(self, state))
        if not native:
            return self.value_net(self._torch_forward(X))
        X = X.contiguous()
        # layer 1: attention computed in-kernel from w_a (and returned), h1 = relu(A X w1) (+ X only in the two-layer net: gcn.py:117-124)
        skip1 = bool(self.skip_connection) and self.num_layer == 2
        h1, A = ops.gcn_layer(X, self.w1.detach(), w_a=self.w_a.detach(), skip=skip1, return_A=True)
        self._A_dev, self._A_host = A[0], None
        if self.num_layer == 1:
            return self.value_net(h1[:, 0, :])
        if self.layerwise_graph:
            h2 = ops.gcn_layer(h1, self.w2.detach(), w_a=self.w_a.detach(), skip=bool(self.skip_connection))
        else:
            h2 = ops.gcn_layer(h1, self.w2.detach(), A=A, skip=bool(self.skip_connection))
        return self.value_net(h2[:, 0, :])


class GCN(Policy):
    """crowd_nav/policy/gcn.py:129-159 + MultiHumanRL (multi_human_rl.py:8-130) + CADRL (cadrl.py:35-276), batched."""

    def __init__(self):
        super().__init__()
        self.name = 'GCN'
        self.trainable = True
        self.multiagent_training = None
        self.kinematics = None
        self.epsilon = None
        self.gamma = None
        self.sampling = None
        self.speed_samples = None
        self.rotation_samples = None
        self.query_env = None
        self.action_space = None
        self.rotation_constraint = None
        self.speeds = None
        self.rotations = None
        self.action_values = None
        self.with_om = None
        self.cell_num = None
        self.cell_size = None
        self.om_channel_size = None
        self.self_state_dim = 6
        self.human_state_dim = 7
        self.joint_state_dim = self.self_state_dim + self.human_state_dim

    # ---- configuration (gcn.py:134-156, cadrl.py:66-85) ----
    def configure(self, config):
        g = config.gcn
        self.multiagent_training = g.multiagent_training
        self.set_common_parameters(config)
        self.model = ValueNetwork(self.input_dim(), self.self_state_dim, g.num_layer, g.X_dim, g.wr_dims, g.wh_dims,
                                  g.final_state_dim, g.gcn2_w1_dim, g.planning_dims, g.similarity_function, g.layerwise_graph,
                                  g.skip_connection)
        logging.info('GCN layers: {}'.format(g.num_layer))
        logging.info('Policy: {}'.format(self.name))

    def set_common_parameters(self, config):
        self.gamma = config.rl.gamma
        self.kinematics = config.action_space.kinematics
        self.sampling = config.action_space.sampling
        self.speed_samples = config.action_space.speed_samples
        self.rotation_samples = config.action_space.rotation_samples
        self.query_env = config.action_space.query_env
        self.rotation_constraint = config.action_space.rotation_constraint
        self.cell_num = config.om.cell_num
        self.cell_size = config.om.cell_size
        self.om_channel_size = config.om.om_channel_size

    def set_device(self, device):
        self.device = device
        self.model.to(device)

    def set_epsilon(self, epsilon):
        self.epsilon = epsilon

    def get_matrix_A(self):
        return self.model.A

    def input_dim(self):
        return self.joint_state_dim + (self.cell_num ** 2 * self.om_channel_size if self.with_om else 0)

    # ---- action table (cadrl.py:93-113): ROTATION-major, unlike the model-predictive planner ----
    def build_action_space(self, v_pref):
        holonomic = self.kinematics == 'holonomic'
        speeds = [(np.exp((i + 1) / self.speed_samples) - 1) / (np.e - 1) * v_pref for i in range(self.speed_samples)]
        if holonomic:
            rotations = np.linspace(0, 2 * np.pi, self.rotation_samples, endpoint=False)
        else:
            rotations = np.linspace(-self.rotation_constraint, self.rotation_constraint, self.rotation_samples)
        action_space = [ActionXY(0, 0) if holonomic else ActionRot(0, 0)]
        for rotation, speed in itertools.product(rotations, speeds):
            if holonomic:
                action_space.append(ActionXY(speed * np.cos(rotation), speed * np.sin(rotation)))
            else:
                action_space.append(ActionRot(speed, rotation))
        self.speeds = speeds
        self.rotations = rotations
        self.action_space = action_space

    # ---- rotate (cadrl.py:241-276) for any leading shape [..., 14] -> [..., 13] ----
    def rotate(self, state):
        px, py, vx, vy, radius, gx, gy, v_pref, theta = [state[..., i] for i in range(9)]
        px1, py1, vx1, vy1, radius1 = [state[..., 9 + i] for i in range(5)]
        dx, dy = gx - px, gy - py
        rot = torch.atan2(dy, dx)
        cr, sr = torch.cos(rot), torch.sin(rot)
        dg = torch.sqrt(dx * dx + dy * dy)
        nvx = vx * cr + vy * sr
        nvy = vy * cr - vx * sr
        th = theta - rot if self.kinematics == 'unicycle' else torch.zeros_like(v_pref)
        nvx1 = vx1 * cr + vy1 * sr
        nvy1 = vy1 * cr - vx1 * sr
        npx1 = (px1 - px) * cr + (py1 - py) * sr
        npy1 = (py1 - py) * cr - (px1 - px) * sr
        da = torch.sqrt((px - px1) * (px - px1) + (py - py1) * (py - py1))
        return torch.stack([dg, v_pref, th, radius, nvx, nvy, npx1, npy1, nvx1, nvy1, radius1, da, radius + radius1], dim=-1)

    def transform(self, state):
        """JointState -> rotated [Nh, 13] on self.device (multi_human_rl.py:98-113)."""
        r = state.robot_state.to_tuple()
        rows = torch.tensor([list(r) + list(h.to_tuple()) for h in state.human_states], dtype=torch.float32, device=self.device)
        return self.rotate(rows)

    # ---- batched one-step look-ahead: every action at once ----
    def _expand(self, state):
        """-> (joint next states [A, Nh, 14] fp32 on the device, rewards float64 numpy [A]).
        propagate: cadrl.py:115-142; reward: multi_human_rl.py:73-96 (float64, like the reference's python floats)."""
        r = state.robot_state
        H = np.array([h.to_tuple() for h in state.human_states], dtype=np.float64)          # [Nh,5]
        A = len(self.action_space)
        dt = self.time_step
        if self.kinematics == 'holonomic':
            acts = np.array([[a.vx, a.vy] for a in self.action_space], dtype=np.float64)
            nvx, nvy = acts[:, 0], acts[:, 1]
            ntheta = np.full(A, r.theta, dtype=np.float64)
        else:
            acts = np.array([[a.v, a.r] for a in self.action_space], dtype=np.float64)
            ntheta = r.theta + acts[:, 1]
            nvx, nvy = acts[:, 0] * np.cos(ntheta), acts[:, 0] * np.sin(ntheta)
        npx, npy = r.px + nvx * dt, r.py + nvy * dt
        hpx, hpy = H[:, 0] + H[:, 2] * dt, H[:, 1] + H[:, 3] * dt                            # humans keep their velocity
        # reward of every action (float64)
        dist = np.sqrt((npx[:, None] - hpx[None, :]) ** 2 + (npy[:, None] - hpy[None, :]) ** 2) - r.radius - H[None, :, 4]
        collision = (dist < 0).any(axis=1)
        dmin = dist.min(axis=1)
        reaching = np.sqrt((npx - r.gx) ** 2 + (npy - r.gy) ** 2) < r.radius
        reward = np.where(collision, -0.25, np.where(reaching, 1.0, np.where(dmin < 0.2, (dmin - 0.2) * 0.5 * dt, 0.0)))
        nh = H.shape[0]
        joint = np.empty((A, nh, 14), dtype=np.float32)
        robot_next = np.stack([npx, npy, nvx, nvy, np.full(A, r.radius), np.full(A, r.gx), np.full(A, r.gy),
                               np.full(A, r.v_pref), ntheta], axis=1)
        joint[:, :, :9] = robot_next[:, None, :]
        joint[:, :, 9] = hpx[None, :]
        joint[:, :, 10] = hpy[None, :]
        joint[:, :, 11] = H[None, :, 2]
        joint[:, :, 12] = H[None, :, 3]
        joint[:, :, 13] = H[None, :, 4]
        return torch.from_numpy(joint).to(self.device), reward

    def predict(self, state):
        if self.phase is None or self.device is None:
            raise AttributeError('Phase, device attributes have to be set!')
        if self.phase == 'train' and self.epsilon is None:
            raise AttributeError('Epsilon attribute has to be set in training phase')
        if self.reach_destination(state):
            return ActionXY(0, 0) if self.kinematics == 'holonomic' else ActionRot(0, 0)
        if self.action_space is None:
            self.build_action_space(state.robot_state.v_pref)
        if not state.human_states:
            assert self.phase != 'train'
            return self.select_greedy_action(state.robot_state)
        if self.query_env:
            raise NotImplementedError('query_env (environment one-step look-ahead) is a simulator feature; use query_env = False')

        probability = np.random.random()
        if self.phase == 'train' and probability < self.epsilon:
            max_action = self.action_space[np.random.choice(len(self.action_space))]
        else:
            joint, reward = self._expand(state)
            with torch.no_grad():
                values = self.model(self.rotate(joint)).view(-1).double().cpu().numpy()
            gamma_bar = pow(self.gamma, self.time_step * state.robot_state.v_pref)
            vals = reward + gamma_bar * values
            self.action_values = [float(v) for v in vals]
            best, best_v = None, float('-inf')
            for i, v in enumerate(self.action_values):           # strict '>' : the first maximum wins (multi_human_rl.py:59)
                if v > best_v:
                    best_v, best = v, i
            if best is None:
                raise ValueError('Value network is not well trained. ')
            max_action = self.action_space[best]
        if self.phase == 'train':
            self.last_state = self.transform(state)
        return max_action

    def select_greedy_action(self, self_state):
        """No humans in sight: head for the goal, closest action in the table (cadrl.py:188-225)."""
        direction = np.arctan2(self_state.gy - self_state.py, self_state.gx - self_state.px)
        distance = np.linalg.norm((self_state.gy - self_state.py, self_state.gx - self_state.px))
        if self.kinematics == 'holonomic':
            speed = min(distance / self.time_step, self_state.v_pref)
            vx, vy = np.cos(direction) * speed, np.sin(direction) * speed
            diffs = [np.linalg.norm(np.array(a) - np.array((vx, vy))) for a in self.action_space]
            return self.action_space[int(np.argmin(diffs))]
        rotation = direction - self_state.theta
        if rotation < self.rotations[0]:
            return ActionRot(self.speeds[0], self.rotations[0])
        if rotation > self.rotations[-1]:
            return ActionRot(self.speeds[0], self.rotations[-1])
        speed = min(distance / self.time_step, self_state.v_pref)
        diffs = [np.linalg.norm(np.array((np.cos(a.r) * a.v, np.sin(a.r) * a.v)) - np.array((np.cos(rotation) * speed, np.sin(rotation) * a.v)))
                 for a in self.action_space]
        return self.action_space[int(np.argmin(diffs))]
