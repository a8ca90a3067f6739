"""Seeded synthetic agent-graph states.

Layouts follow the reference state tuples (crowd_sim/envs/utils/state.py:27-28, 51-52):
  robot[B,1,9]  = (px, py, vx, vy, radius, gx, gy, v_pref, theta)
  humans[B,Nh,5] = (px, py, vx, vy, radius)
World frame, un-normalised (model_predictive_rl.py:366-370).  The distribution mimics the
`circle_crossing` scenario (crowd_sim/envs/crowd_sim.py:123-140; config.py:34,42-50).
"""
import math

import torch


def synthetic_states(batch, human_num, seed=1234, device='cpu', dtype=torch.float32):
    """Return (robot[B,1,9], humans[B,Nh,5]) drawn with torch.Generator().manual_seed(seed) on CPU."""
    g = torch.Generator().manual_seed(int(seed))
    robot = torch.empty(batch, 1, 9, dtype=torch.float32)
    humans = torch.empty(batch, human_num, 5, dtype=torch.float32)

    def uni(shape, lo, hi):
        return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    robot[:, 0, 0:2] = uni((batch, 2), -4.0, 4.0)
    speed = uni((batch,), 0.0, 1.0)
    ang = uni((batch,), 0.0, 2.0 * math.pi)
    robot[:, 0, 2] = speed * torch.cos(ang)
    robot[:, 0, 3] = speed * torch.sin(ang)
    robot[:, 0, 4] = 0.3
    robot[:, 0, 5] = 0.0
    robot[:, 0, 6] = 4.0
    robot[:, 0, 7] = 1.0
    robot[:, 0, 8] = math.pi / 2

    humans[:, :, 0:2] = uni((batch, human_num, 2), -5.0, 5.0)
    hspeed = uni((batch, human_num), 0.0, 1.0)
    hang = uni((batch, human_num), 0.0, 2.0 * math.pi)
    humans[:, :, 2] = hspeed * torch.cos(hang)
    humans[:, :, 3] = hspeed * torch.sin(hang)
    humans[:, :, 4] = 0.3
    return robot.to(device=device, dtype=dtype), humans.to(device=device, dtype=dtype)
