"""Boundary types of the simulator that the planner touches.

If the reference's `crowd_sim` package is importable (the reference checkout is on sys.path, with
gym / rvo2 / ... installed) its own classes are used, so isinstance checks inside the simulator
hold.  Otherwise minimal stand-ins with the same fields are defined:
  ActionXY / ActionRot                      crowd_sim/envs/utils/action.py:3-4
  FullState / ObservableState / JointState  crowd_sim/envs/utils/state.py:4-79
  Policy                                    crowd_sim/envs/policy/policy.py:6-65
"""
from collections import namedtuple

import numpy as np
import torch

try:  # pragma: no cover - depends on the user's environment
    from crowd_sim.envs.policy.policy import Policy
    from crowd_sim.envs.utils.action import ActionRot, ActionXY
    from crowd_sim.envs.utils.state import FullState, JointState, ObservableState
    HAVE_CROWD_SIM = True
except Exception:  # noqa: BLE001 - any import failure (gym, rvo2, matplotlib ...) selects the stand-ins
    HAVE_CROWD_SIM = False

    ActionXY = namedtuple('ActionXY', ['vx', 'vy'])
    ActionRot = namedtuple('ActionRot', ['v', 'r'])

    class FullState(object):
        FIELDS = ('px', 'py', 'vx', 'vy', 'radius', 'gx', 'gy', 'v_pref', 'theta')

        def __init__(self, px, py, vx, vy, radius, gx, gy, v_pref, theta):
            for k, v in zip(self.FIELDS, (px, py, vx, vy, radius, gx, gy, v_pref, theta)):
                setattr(self, k, v)
            self.position = (px, py)
            self.goal_position = (gx, gy)
            self.velocity = (vx, vy)

        def to_tuple(self):
            return tuple(getattr(self, k) for k in self.FIELDS)

        def __add__(self, other):
            return other + self.to_tuple()

        def __str__(self):
            return ' '.join(str(x) for x in self.to_tuple())

        def get_observable_state(self):
            return ObservableState(self.px, self.py, self.vx, self.vy, self.radius)

    class ObservableState(object):
        FIELDS = ('px', 'py', 'vx', 'vy', 'radius')

        def __init__(self, px, py, vx, vy, radius):
            for k, v in zip(self.FIELDS, (px, py, vx, vy, radius)):
                setattr(self, k, v)
            self.position = (px, py)
            self.velocity = (vx, vy)

        def to_tuple(self):
            return tuple(getattr(self, k) for k in self.FIELDS)

        def __add__(self, other):
            return other + self.to_tuple()

        def __str__(self):
            return ' '.join(str(x) for x in self.to_tuple())

    class JointState(object):
        def __init__(self, robot_state, human_states):
            assert isinstance(robot_state, FullState)
            for h in human_states:
                assert isinstance(h, ObservableState)
            self.robot_state = robot_state
            self.human_states = human_states

        def to_tensor(self, add_batch_size=False, device=None):
            robot = torch.Tensor([self.robot_state.to_tuple()])
            humans = torch.Tensor([h.to_tuple() for h in self.human_states])
            if add_batch_size:
                robot, humans = robot.unsqueeze(0), humans.unsqueeze(0)
            if device is not None:
                robot, humans = robot.to(device), humans.to(device)
            return robot, humans

    class Policy(object):
        def __init__(self):
            self.trainable = False
            self.phase = None
            self.model = None
            self.device = None
            self.last_state = None
            self.time_step = None
            self.env = None

        def configure(self, config):
            raise NotImplementedError

        def set_phase(self, phase):
            self.phase = phase

        def set_device(self, device):
            self.device = device

        def set_env(self, env):
            self.env = env

        def set_time_step(self, time_step):
            self.time_step = time_step

        def get_model(self):
            return self.model

        def save_model(self, file):
            torch.save(self.model.state_dict(), file)

        def load_model(self, file):
            self.model.load_state_dict(torch.load(file))

        def get_state_dict(self):
            return self.model.state_dict()

        def load_state_dict(self, state_dict):
            self.model.load_state_dict(state_dict)

        def predict(self, state):
            raise NotImplementedError

        @staticmethod
        def reach_destination(state):
            r = state.robot_state
            return bool(np.linalg.norm((r.py - r.gy, r.px - r.gx)) < r.radius)


def joint_state_to_tensors(state, device):
    """(robot[1,1,9], humans[1,Nh,5]) fp32 on `device` (state.py:64-79 without its cuda:0-only quirk)."""
    robot = torch.tensor([[list(state.robot_state.to_tuple())]], dtype=torch.float32)
    humans = torch.tensor([[list(h.to_tuple()) for h in state.human_states]], dtype=torch.float32)
    return robot.to(device), humans.to(device)


def joint_states_to_tensors(states, device):
    """Vectorised JointState.to_tensor (state.py:64-79) for many environments that are stepped together:
    list of JointState (equal human counts) -> (robot[E,1,9], humans[E,Nh,5]) fp32 on `device`, staged through ONE pinned
    host buffer and one host->device copy per tensor instead of 2*E tiny tensors."""
    E = len(states)
    nh = len(states[0].human_states)
    flat = np.empty((E, 9 + 5 * nh), dtype=np.float32)
    for e, st in enumerate(states):
        assert len(st.human_states) == nh, 'states stepped together must have the same number of humans'
        flat[e, :9] = st.robot_state.to_tuple()
        for h, hs in enumerate(st.human_states):
            flat[e, 9 + 5 * h: 14 + 5 * h] = hs.to_tuple()
    host = torch.from_numpy(flat)
    dev = torch.device(device)
    if dev.type == 'cuda':
        host = host.pin_memory()
    both = host.to(dev, non_blocking=True)
    return both[:, :9].reshape(E, 1, 9).contiguous(), both[:, 9:].reshape(E, nh, 5).contiguous()
