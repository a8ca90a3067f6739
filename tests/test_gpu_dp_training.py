"""Data-parallel training step on real GPUs (-m gpu, needs >= 2 devices; skipped otherwise): NCCL flat-buffer
gradient all-reduce + native backward equals the single-GPU step on the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out):
    import copy
    from relationalgraphlearning_b200 import parallel as PAR
    from relationalgraphlearning_b200.config import policy_config
    from relationalgraphlearning_b200.graph_model import RGL
    from relationalgraphlearning_b200.synthetic import synthetic_states
    from relationalgraphlearning_b200.value_estimator import ValueEstimator
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(0)
    ve = ValueEstimator(policy_config(), RGL(policy_config(), 9, 5)).to(dev)
    tgt = copy.deepcopy(ve)
    robot, humans = synthetic_states(B, 5, seed=3, device=dev)
    nrobot, nhumans = synthetic_states(B, 5, seed=4, device=dev)
    rewards = torch.linspace(-0.25, 1.0, B, device=dev).unsqueeze(1)
    lo, hi = PAR.shard_range(B, rank, world)
    opt = torch.optim.SGD(ve.parameters(), lr=0.01)
    red = PAR.FlatGradAllReducer(ve.parameters())
    loss = PAR.dp_value_step(ve, tgt, opt, red, robot[lo:hi], humans[lo:hi], rewards[lo:hi], nrobot[lo:hi], nhumans[lo:hi],
                             0.9 ** 0.25, B)
    dist.all_reduce(loss)
    if rank == 0:
        torch.save({'params': [p.detach().cpu() for p in ve.parameters()], 'loss': loss.cpu(), 'grad': red.buf.cpu()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_dp_step_equals_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import copy
    from relationalgraphlearning_b200.config import policy_config
    from relationalgraphlearning_b200.graph_model import RGL
    from relationalgraphlearning_b200.synthetic import synthetic_states
    from relationalgraphlearning_b200.value_estimator import ValueEstimator
    B = 96
    out = str(tmp_path / 'dp.pt')
    mp.spawn(_worker, args=(2, _free_port(), B, out), nprocs=2, join=True)
    got = torch.load(out)
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    ve = ValueEstimator(policy_config(), RGL(policy_config(), 9, 5)).to(dev)
    tgt = copy.deepcopy(ve)
    robot, humans = synthetic_states(B, 5, seed=3, device=dev)
    nrobot, nhumans = synthetic_states(B, 5, seed=4, device=dev)
    rewards = torch.linspace(-0.25, 1.0, B, device=dev).unsqueeze(1)
    opt = torch.optim.SGD(ve.parameters(), lr=0.01)
    opt.zero_grad()
    with torch.no_grad():
        target = rewards + 0.9 ** 0.25 * tgt((nrobot, nhumans))
    loss = torch.nn.functional.mse_loss(ve((robot, humans)), target)
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in ve.parameters()]).cpu()
    opt.step()
    assert torch.allclose(got['loss'], loss.detach().cpu(), rtol=1e-5, atol=1e-7)
    assert torch.allclose(got['grad'], flat, rtol=2e-4, atol=1e-6)
    for a, b in zip(got['params'], ve.parameters()):
        assert torch.allclose(a, b.detach().cpu(), rtol=1e-4, atol=1e-6)
