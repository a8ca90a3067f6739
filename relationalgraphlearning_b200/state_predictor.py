"""StatePredictor / LinearStatePredictor -- drop-ins for crowd_nav/policy/state_predictor.py:7-118.

forward(state, action, detach=False) -> [next_robot | None, next_humans[B,Nh,5]].  The human branch is
the fused graph kernel with the motion head; the robot kinematic step is batched (the reference raises
for B != 1, state_predictor.py:43-44) and accepts either one action for the whole batch or a [B,2]
tensor of (vx, vy).
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from . import _torch_math as TM
from .helpers import mlp


def compute_next_robot_state(robot_state, action, time_step, kinematics):
    """robot_state[B,1,9]; action: ActionXY/ActionRot-like or tensor [B,2].  state_predictor.py:41-60."""
    nxt = robot_state.clone()
    if isinstance(action, torch.Tensor):
        a0, a1 = action[:, 0].to(robot_state.dtype), action[:, 1].to(robot_state.dtype)
        d0, d1 = a0 * time_step, a1 * time_step
    else:
        a0, a1 = float(action[0]), float(action[1])
        d0, d1 = a0 * time_step, a1 * time_step       # python floats (float64 product, like the reference)
    if kinematics == 'holonomic':
        nxt[:, 0, 0] = robot_state[:, 0, 0] + d0
        nxt[:, 0, 1] = robot_state[:, 0, 1] + d1
        nxt[:, 0, 2] = a0
        nxt[:, 0, 3] = a1
    else:
        # The reference's unicycle branch adds the rotation to index 7 (= v_pref, not theta);
        # the same indexing is kept for behavioural parity.
        nxt[:, 0, 7] = robot_state[:, 0, 7] + a1
        c, s = torch.cos(nxt[:, 0, 7]), torch.sin(nxt[:, 0, 7])
        nxt[:, 0, 0] = robot_state[:, 0, 0] + c * a0 * time_step
        nxt[:, 0, 1] = robot_state[:, 0, 1] + s * a0 * time_step
        nxt[:, 0, 2] = c * a0
        nxt[:, 0, 3] = s * a0
    return nxt


class StatePredictor(nn.Module):
    def __init__(self, config, graph_model, time_step):
        super().__init__()
        self.trainable = True
        self.kinematics = config.action_space.kinematics
        self.graph_model = graph_model
        self._dims = list(config.model_predictive_rl.motion_predictor_dims)
        self.human_motion_predictor = mlp(config.gcn.X_dim, self._dims)
        self.time_step = time_step
        self._pack_cache = ops._PackCache()

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k not in ('_pack_cache', '_grad_sink'):
                setattr(new, k, copy.deepcopy(v, memo))
        new._pack_cache = ops._PackCache()
        return new

    def _train_params(self, detach=False):
        mp = list(self.human_motion_predictor.parameters())
        return mp if detach else self.graph_model.param_tensors() + mp

    def kernel_supported(self):
        g = self.graph_model
        return hasattr(g, 'kernel_supported') and g.kernel_supported() and self._dims == [64, 5]

    def run(self, robot, humans, humans_bcast=1, throughput=False):
        """One launch: graph kernel with the motion head, no autograd.  -> next_humans[B,Nh,5]."""
        mblob = ops.packed_motion(self.human_motion_predictor, self._pack_cache)
        return self.graph_model.run(robot, humans, humans_bcast=humans_bcast, motion_blob=mblob, want_S=True,
                                    throughput=throughput)['S']

    def _torch_humans(self, robot, humans, detach):
        emb = TM.graph_forward(self.graph_model, robot, humans)
        if detach:
            emb = emb.detach()
        return self.human_motion_predictor(emb)[:, 1:, :]

    def forward(self, state, action, detach=False):
        assert len(state[0].shape) == 3
        assert len(state[1].shape) == 3
        robot, humans = state
        ops.require_cuda_or_cpu_module(self, robot, humans, 'StatePredictor.forward')
        next_robot = None if action is None else self.compute_next_state(robot, action)
        if not robot.is_cuda or not self.kernel_supported() or not self.graph_model.shape_supported(humans):
            next_humans = self._torch_humans(robot, humans, detach)
        elif ops._needs_grad(self, robot, humans):
            from . import training
            if training.native_supported(self) and not (robot.requires_grad or humans.requires_grad):
                # fused forward with saves + hand-written backward (detach => only the motion head gets gradients)
                return [next_robot, training.statepred_forward_train(self, robot, humans, bool(detach))]
            params = list(self.human_motion_predictor.parameters())
            if not detach:
                params = self.graph_model.param_tensors() + params
            next_humans = ops.fused_with_autograd(lambda: self.run(robot, humans),
                                                  lambda: self._torch_humans(robot, humans, detach),
                                                  params, [robot, humans])
        else:
            next_humans = self.run(robot, humans)
        return [next_robot, next_humans]

    def compute_next_state(self, robot_state, action):
        return compute_next_robot_state(robot_state, action, self.time_step, self.kinematics)


class LinearStatePredictor(object):
    """Non-learned predictor (state_predictor.py:63-118): robot by kinematics, humans p += v
    (the reference adds v without the time-step factor, :115-116 -- kept)."""

    def __init__(self, config, time_step):
        self.trainable = False
        self.kinematics = config.action_space.kinematics
        self.time_step = time_step

    def __call__(self, state, action):
        assert len(state[0].shape) == 3
        assert len(state[1].shape) == 3
        return [self.compute_next_state(state[0], action), self.linear_motion_approximator(state[1])]

    def compute_next_state(self, robot_state, action):
        return compute_next_robot_state(robot_state, action, self.time_step, self.kinematics)

    @staticmethod
    def linear_motion_approximator(human_states):
        nxt = human_states.clone()
        nxt[..., 0] = human_states[..., 0] + human_states[..., 2]
        nxt[..., 1] = human_states[..., 1] + human_states[..., 3]
        return nxt
