"""N>1 host logic on CPU: world_size-2 gloo processes (127.0.0.1) exercising the batch sharding and the single
flat-buffer gradient all-reduce.  The per-shard evaluator here is a small torch module -- the product modules are
CUDA-only -- because what is under test is the partitioning / collective plumbing, not the kernels."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from relationalgraphlearning_b200 import parallel as PAR
from relationalgraphlearning_b200.config import policy_config
from relationalgraphlearning_b200.graph_model import RGL
from relationalgraphlearning_b200.helpers import mlp
from relationalgraphlearning_b200 import _torch_math as TM
from relationalgraphlearning_b200.synthetic import synthetic_states


class TorchValue(torch.nn.Module):
    """Torch-op value estimator with the product's parameter structure (evaluates on CPU for this test only)."""

    def __init__(self, cfg):
        super().__init__()
        self.graph_model = RGL(cfg, 9, 5)
        self.value_network = mlp(32, [32, 100, 100, 1])

    def forward(self, state):
        return self.value_network(TM.graph_forward(self.graph_model, state[0], state[1])[:, 0, :])


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    torch.manual_seed(0)
    model = TorchValue(policy_config())
    target = TorchValue(policy_config())
    target.load_state_dict(model.state_dict())
    robot, humans = synthetic_states(B, 5, seed=3)
    nrobot, nhumans = synthetic_states(B, 5, seed=4)
    rewards = torch.linspace(-0.25, 1.0, B).unsqueeze(1)
    lo, hi = PAR.shard_range(B, rank, world)
    opt = torch.optim.SGD(model.parameters(), lr=0.01)
    red = PAR.FlatGradAllReducer(model.parameters())
    loss = PAR.dp_value_step(model, target, opt, red, robot[lo:hi], humans[lo:hi], rewards[lo:hi], nrobot[lo:hi], nhumans[lo:hi],
                             0.9 ** 0.25, B)
    tot = loss.clone()
    dist.all_reduce(tot)
    with torch.no_grad():
        v_local = model((robot[lo:hi], humans[lo:hi]))
    v_all = PAR.gather_results(v_local, B)
    if rank == 0:
        torch.save({'grad': red.buf.clone(), 'loss': tot, 'params': [p.detach().clone() for p in model.parameters()],
                    'v_all': v_all, 'numel': red.numel}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('B', [10, 7])
def test_dp_value_step_equals_single_process(tmp_path, B):
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(2, _free_port(), B, out), nprocs=2, join=True)
    got = torch.load(out)
    # single-process reference of the same step (trainer.py:122-131 with MSELoss(mean))
    torch.manual_seed(0)
    model = TorchValue(policy_config())
    target = TorchValue(policy_config())
    target.load_state_dict(model.state_dict())
    robot, humans = synthetic_states(B, 5, seed=3)
    nrobot, nhumans = synthetic_states(B, 5, seed=4)
    rewards = torch.linspace(-0.25, 1.0, B).unsqueeze(1)
    opt = torch.optim.SGD(model.parameters(), lr=0.01)
    opt.zero_grad()
    outv = model((robot, humans))
    tgt = rewards + 0.9 ** 0.25 * target((nrobot, nhumans)).detach()
    loss = torch.nn.functional.mse_loss(outv, tgt)
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    opt.step()
    assert got['numel'] == 22813                       # one 91 KB message per step
    assert torch.allclose(got['loss'], loss.detach(), rtol=1e-5, atol=1e-7)
    assert torch.allclose(got['grad'], flat, rtol=1e-4, atol=1e-6)
    for a, b in zip(got['params'], model.parameters()):
        assert torch.allclose(a, b.detach(), rtol=1e-4, atol=1e-6)
    with torch.no_grad():
        assert torch.allclose(got['v_all'], model((robot, humans)), rtol=1e-5, atol=1e-6)


def test_shard_ranges_cover_and_balance():
    for total in (0, 1, 7, 4096, 16385):
        for world in (1, 2, 3, 8):
            spans = [PAR.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
