"""Data-parallel training step on real GPUs (-m gpu, needs >= 2 devices; skipped otherwise): the native backward writes
every gradient into ONE flat buffer, one collective sums it across ranks -- the one-kernel push all-reduce over NVLink
peer memory (csrc/dp_comm.cu, backend 'p2p') or NCCL (backend 'nccl') -- and the result equals the single-GPU step on the
concatenated batch (crowd_nav/utils/trainer.py:122-131)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

STEPS = 3          # > 2: exercises both parities of the double-buffered receive slots and the in-kernel re-zeroing


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data(B, dev, k):
    from relationalgraphlearning_b200.synthetic import synthetic_states
    robot, humans = synthetic_states(B, 5, seed=3 + 10 * k, device=dev)
    nrobot, nhumans = synthetic_states(B, 5, seed=4 + 10 * k, device=dev)
    rewards = torch.linspace(-0.25, 1.0, B, device=dev).unsqueeze(1)
    return robot, humans, rewards, nrobot, nhumans


def _worker(rank, world, port, B, backend, graphed, out):
    import copy
    from relationalgraphlearning_b200 import parallel as PAR
    from relationalgraphlearning_b200.config import policy_config
    from relationalgraphlearning_b200.graph_model import RGL
    from relationalgraphlearning_b200.value_estimator import ValueEstimator
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(0)
    ve = ValueEstimator(policy_config(), RGL(policy_config(), 9, 5)).to(dev)
    tgt = copy.deepcopy(ve)
    lo, hi = PAR.shard_range(B, rank, world)
    opt = torch.optim.SGD(ve.parameters(), lr=0.01)
    red = PAR.FlatGrads(ve, backend=backend)
    assert red.backend == backend, (red.backend, red.fallback_reason)
    assert red.message_bytes == 91252
    losses, grads = [], []
    shards = [tuple(t[lo:hi].contiguous() for t in _data(B, dev, k)) for k in range(STEPS)]
    if graphed:
        # the whole step (forward, target forward, loss, native backward, peer-memory all-reduce, SGD) replayed from a CUDA graph
        static = [t.clone() for t in shards[0]]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ve((static[0], static[1]))          # pack weights / load kernels outside the capture (no parameter update)
        side.synchronize()
        dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            gl = PAR.dp_value_step(ve, tgt, opt, red, *static, 0.9 ** 0.25, B)
        for k in range(STEPS):
            for s, t in zip(static, shards[k]):
                s.copy_(t)
            g.replay()
            loss = gl.clone()
            dist.all_reduce(loss)
            losses.append(loss.cpu())
            grads.append(red.flat_grad().cpu())
    else:
        for k in range(STEPS):
            loss = PAR.dp_value_step(ve, tgt, opt, red, *shards[k], 0.9 ** 0.25, B)
            dist.all_reduce(loss)
            losses.append(loss.cpu())
            grads.append(red.flat_grad().cpu())
    assert red.status() == 0
    torch.save({'params': [p.detach().cpu() for p in ve.parameters()], 'loss': losses, 'grad': grads}, out + '.%d' % rank)
    dist.barrier()
    red.close()
    dist.destroy_process_group()


@pytest.mark.parametrize('backend,graphed', [('p2p', False), ('p2p', True), ('nccl', False)])
def test_two_gpu_dp_step_equals_single_gpu(tmp_path, backend, graphed):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import copy
    from relationalgraphlearning_b200.config import policy_config
    from relationalgraphlearning_b200.graph_model import RGL
    from relationalgraphlearning_b200.value_estimator import ValueEstimator
    B = 96
    out = str(tmp_path / 'dp.pt')
    mp.spawn(_worker, args=(2, _free_port(), B, backend, graphed, out), nprocs=2, join=True)
    got = [torch.load(out + '.%d' % r) for r in range(2)]
    # the replicas stay bit-identical: every rank sums the contributions in rank order
    for a, b in zip(got[0]['params'], got[1]['params']):
        assert torch.equal(a, b)
    for a, b in zip(got[0]['grad'], got[1]['grad']):
        assert torch.equal(a, b)
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    ve = ValueEstimator(policy_config(), RGL(policy_config(), 9, 5)).to(dev)
    tgt = copy.deepcopy(ve)
    opt = torch.optim.SGD(ve.parameters(), lr=0.01)
    for k in range(STEPS):
        robot, humans, rewards, nrobot, nhumans = _data(B, dev, k)
        opt.zero_grad()
        with torch.no_grad():
            target = rewards + 0.9 ** 0.25 * tgt((nrobot, nhumans))
        loss = torch.nn.functional.mse_loss(ve((robot, humans)), target)
        loss.backward()
        flat = torch.cat([p.grad.reshape(-1) for p in ve.parameters()]).cpu()
        opt.step()
        assert torch.allclose(got[0]['loss'][k], loss.detach().cpu(), rtol=1e-5, atol=1e-7), k
        scale = float(flat.abs().max())
        assert float((got[0]['grad'][k] - flat).abs().max()) <= 2e-4 * scale, k
    for a, b in zip(got[0]['params'], ve.parameters()):
        assert torch.allclose(a, b.detach().cpu(), rtol=1e-4, atol=1e-6)
