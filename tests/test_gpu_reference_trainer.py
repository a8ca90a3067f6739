"""The reference's OWN training-side caller over the drop-in modules (-m gpu): crowd_nav/utils/trainer.py `MPRLTrainer`
(optimize_epoch :63-108, optimize_batch :110-161) and crowd_nav/utils/memory.py `ReplayMemory`, vendored UNMODIFIED into
oracle/_ref by oracle/make_ref.py, drive `ValueEstimator` / `StatePredictor` on the GPU exactly as crowd_nav/train.py
does (DataLoader collate, nn.MSELoss, torch.optim.Adam, deepcopy'd target network).  The same trainer is then run over
CPU modules whose forward is the oracle restatement (autograd through oracle/rgl_oracle.py), with the same seeds, and
losses / parameters are compared after every call.

The reference's own modules cannot stand in on the CPU side: their in-place skip add makes loss.backward() raise on this
torch; the oracle's forward is bit-identical and uses the out-of-place add."""
import copy

import pytest
import torch
import torch.nn as nn

from oracle import make_ref
from oracle import rgl_oracle as O
from relationalgraphlearning_b200.config import policy_config
from relationalgraphlearning_b200.graph_model import RGL
from relationalgraphlearning_b200.helpers import mlp
from relationalgraphlearning_b200.replay import DeviceReplayMemory
from relationalgraphlearning_b200.state_predictor import StatePredictor
from relationalgraphlearning_b200.synthetic import synthetic_states
from relationalgraphlearning_b200.value_estimator import ValueEstimator

pytestmark = pytest.mark.gpu


class Writer(object):
    def __init__(self):
        self.scalars = []

    def add_scalar(self, tag, value, step):
        self.scalars.append((tag, float(value), step))


class OracleValue(nn.Module):
    """CPU checker: same parameter tree as ValueEstimator, forward = the oracle (differentiable torch ops)."""

    def __init__(self, cfg):
        super().__init__()
        self.graph_model = RGL(cfg, 9, 5)
        self.value_network = mlp(32, [32, 100, 100, 1])

    def forward(self, state):
        return O.value_forward(dict(self.graph_model.named_parameters()), dict(self.value_network.named_parameters()), state[0], state[1])


class OracleStatePredictor(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.trainable = True
        self.graph_model = RGL(cfg, 9, 5)
        self.human_motion_predictor = mlp(32, [64, 5])

    def forward(self, state, action, detach=False):
        gsd = dict(self.graph_model.named_parameters())
        H = O.rgl_forward(gsd, state[0], state[1])
        if detach:
            H = H.detach()
        return [None, O.mlp(H, dict(self.human_motion_predictor.named_parameters()), '')[:, 1:, :]]


def transitions(n, nh, seed):
    r, h = synthetic_states(n, nh, seed=seed)
    r2, h2 = synthetic_states(n, nh, seed=seed + 1)
    g = torch.Generator().manual_seed(seed + 2)
    val = torch.rand(n, 1, generator=g)
    rew = torch.rand(n, 1, generator=g) * 1.25 - 0.25
    return [(r[i], h[i], val[i], rew[i], r2[i], h2[i]) for i in range(n)]


def run_trainer(MPRLTrainer, ve, sp, memory, device, loader=None):
    w = Writer()
    tr = MPRLTrainer(ve, sp, memory, device, None, w, 64, 'Adam', 5, False, False, False, False)
    tr.set_learning_rate(1e-3)
    tr.update_target_model(ve)
    if loader is not None:
        tr.data_loader = loader
    log = []
    torch.manual_seed(11)
    tr.optimize_epoch(2)
    log += [s[1] for s in w.scalars]
    torch.manual_seed(12)
    log += list(tr.optimize_batch(3, 0))
    tr.update_target_model(ve)
    torch.manual_seed(13)
    log += list(tr.optimize_batch(2, 1))
    return log


def build_pair(device):
    cfg = policy_config()
    torch.manual_seed(5)
    ove, osp = OracleValue(cfg), OracleStatePredictor(cfg)
    ve = ValueEstimator(cfg, RGL(cfg, 9, 5))
    sp = StatePredictor(cfg, RGL(cfg, 9, 5), 0.25)
    ve.graph_model.load_state_dict(ove.graph_model.state_dict())
    ve.value_network.load_state_dict(ove.value_network.state_dict())
    sp.graph_model.load_state_dict(osp.graph_model.state_dict())
    sp.human_motion_predictor.load_state_dict(osp.human_motion_predictor.state_dict())
    return ove, osp, ve.to(device), sp.to(device)


def compare(log_gpu, log_cpu, mods_gpu, mods_cpu):
    assert len(log_gpu) == len(log_cpu) and len(log_gpu) >= 8
    for a, b in zip(log_gpu, log_cpu):
        assert abs(a - b) <= 2e-3 * max(abs(b), 1e-3), (log_gpu, log_cpu)
    for mg, mc in zip(mods_gpu, mods_cpu):
        for (k, pg), (_, pc) in zip(mg.named_parameters(), mc.named_parameters()):
            err = float((pg.detach().cpu() - pc.detach()).abs().max())
            assert err <= 2e-3 * max(float(pc.detach().abs().max()), 1e-2), (k, err)


def test_reference_trainer_runs_unmodified_over_the_dropin(cuda_device):
    if not make_ref.enable():
        pytest.skip('oracle/_ref not built (python oracle/make_ref.py needs the reference checkout)')
    from crowd_nav.utils.memory import ReplayMemory
    from crowd_nav.utils.trainer import MPRLTrainer
    ove, osp, ve, sp = build_pair(cuda_device)
    items = transitions(300, 5, 40)
    mem_cpu, mem_gpu = ReplayMemory(1000), ReplayMemory(1000)
    for it in items:
        mem_cpu.push(it)
        mem_gpu.push(tuple(x.to(cuda_device) for x in it))      # Explorer pushes device tensors (explorer.py:133-138)
    log_cpu = run_trainer(MPRLTrainer, ove, osp, mem_cpu, torch.device('cpu'))
    log_gpu = run_trainer(MPRLTrainer, ve, sp, mem_gpu, cuda_device)
    compare(log_gpu, log_cpu, (ve, sp), (ove, osp))


def test_reference_trainer_with_device_replay_memory(cuda_device):
    """Same trainer, minibatches from the GPU-resident replay memory (one gather launch instead of DataLoader collate):
    identical batch order under the same seeds, so the same losses / parameters as the CPU run over ReplayMemory."""
    if not make_ref.enable():
        pytest.skip('oracle/_ref not built')
    from crowd_nav.utils.memory import ReplayMemory
    from crowd_nav.utils.trainer import MPRLTrainer
    ove, osp, ve, sp = build_pair(cuda_device)
    items = transitions(300, 5, 40)
    mem_cpu = ReplayMemory(1000)
    mem_dev = DeviceReplayMemory(1000, 5, cuda_device)
    for it in items:
        mem_cpu.push(it)
        mem_dev.push(tuple(x.to(cuda_device) for x in it))
    assert len(mem_dev) == 300 and not mem_dev.is_full()
    got = mem_dev[7]
    for a, b in zip(got, items[7]):
        assert torch.equal(a.cpu().reshape(-1), b.reshape(-1))
    log_cpu = run_trainer(MPRLTrainer, ove, osp, mem_cpu, torch.device('cpu'))
    log_gpu = run_trainer(MPRLTrainer, ve, sp, mem_dev, cuda_device, loader=mem_dev.loader(64, shuffle=True))
    compare(log_gpu, log_cpu, (ve, sp), (ove, osp))


def test_device_replay_memory_semantics(cuda_device):
    mem = DeviceReplayMemory(8, 3, cuda_device)
    items = transitions(11, 3, 7)
    for it in items:
        mem.push(it)
    assert mem.is_full() and len(mem) == 8 and mem.position == 3          # ring: slots 0..2 hold items 8..10
    for slot, src in ((0, 8), (2, 10), (3, 3), (7, 7)):
        for a, b in zip(mem[slot], items[src]):
            assert torch.equal(a.cpu().reshape(-1), b.reshape(-1))
    idx = torch.tensor([7, 0, 0, 3], device=cuda_device)
    rb, hb, vb, wb, r2b, h2b = mem.gather(idx)
    assert rb.shape == (4, 1, 9) and hb.shape == (4, 3, 5) and vb.shape == (4, 1) and r2b.shape == (4, 1, 9)
    for row, src in enumerate((7, 8, 8, 3)):
        assert torch.equal(rb[row].cpu(), items[src][0]) and torch.equal(hb[row].cpu(), items[src][1])
        assert torch.equal(vb[row].cpu(), items[src][2]) and torch.equal(wb[row].cpu(), items[src][3])
        assert torch.equal(r2b[row].cpu(), items[src][4]) and torch.equal(h2b[row].cpu(), items[src][5])
    s = mem.sample(5)
    assert s[0].shape == (5, 1, 9)
    mem2 = DeviceReplayMemory(8, 3, cuda_device)
    cat = [torch.stack([it[j] for it in items[:5]]) for j in range(6)]
    mem2.push_batch(*[c.to(cuda_device) for c in cat])
    assert len(mem2) == 5 and torch.equal(mem2[4][1].cpu(), items[4][1])
    mem.clear()
    assert len(mem) == 0
    assert copy.copy(mem).capacity == 8
