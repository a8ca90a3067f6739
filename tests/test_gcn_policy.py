"""Model-free GCN policy (crowd_nav/policy/gcn.py, multi_human_rl.py, cadrl.py): value network and the
batched predict() against golden vectors minted from the reference policy (oracle/gen_golden.py gcn_case)."""
import numpy as np
import pytest
import torch

from conftest import assert_close_scaled, load_golden
from relationalgraphlearning_b200 import ops
from relationalgraphlearning_b200.config import Config, policy_config
from relationalgraphlearning_b200.gcn import GCN, ValueNetwork
from relationalgraphlearning_b200.simtypes import ActionXY, FullState, JointState, ObservableState

CASES = ['gcn_nh5', 'gcn_nh5_layerwise_noskip', 'gcn_nh3_l1']


def gcn_config(g):
    m = g['meta']
    cfg = policy_config(layerwise_graph=bool(m[5]), skip_connection=bool(m[6]), num_layer=int(m[7]))
    cfg.name = 'gcn'
    cfg.gcn.gcn2_w1_dim = 32
    cfg.gcn.planning_dims = [150, 100, 100, 1]
    cfg.om = Config(cell_num=4, cell_size=1, om_channel_size=3)
    return cfg


def make_policy(g, dev):
    pol = GCN()
    pol.configure(gcn_config(g))
    pol.model.load_state_dict(g['model'])
    pol.set_device(dev)
    pol.set_phase('test')
    pol.time_step = 0.25
    return pol


def load(case):
    return load_golden(case)


def joint_state(g, b):
    r = [float(x) for x in g['robot'][b, 0]]
    nh = int(g['meta'][1])
    return JointState(FullState(*r), [ObservableState(*[float(x) for x in g['humans'][b, h]]) for h in range(nh)])


@pytest.mark.parametrize('case', CASES)
def test_gcn_policy_cpu_matches_reference_golden(case):
    g = load(case)
    pol = make_policy(g, torch.device('cpu'))
    assert [tuple(k.shape) for k in pol.model.state_dict().values()] == [tuple(v.shape) for v in g['model'].values()]
    with torch.no_grad():
        v = pol.model(g['rotated'])
    assert_close_scaled(v, g['values'], 1e-6, case + ':values')
    assert np.abs(pol.model.A - g['A0'].numpy()).max() <= 1e-6
    pol.build_action_space(1.0)
    assert np.array_equal(np.array([[a.vx, a.vy] for a in pol.action_space]), np.asarray(g['actions']))
    for b in range(int(g['meta'][2])):
        a = pol.predict(joint_state(g, b))
        ref = np.asarray(g['action_values'][b], dtype=np.float64)
        assert np.abs(np.asarray(pol.action_values) - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())
        srt = np.sort(ref)[::-1]
        if srt[0] - srt[1] > 1e-6:
            assert a == pol.action_space[int(g['chosen'][b])]
        assert isinstance(a, ActionXY)
    # the rotated transform of a JointState: [Nh, 13]
    assert pol.transform(joint_state(g, 0)).shape == (int(g['meta'][1]), 13)


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
def test_gcn_policy_gpu_matches_reference_golden(case, cuda_device):
    """On CUDA every GCN layer is one launch of the stand-alone tcgen05 + TMA layer kernel (rgl_gcn_layer)."""
    g = load(case)
    pol = make_policy(g, cuda_device)
    l0 = ops.LAUNCHES
    with torch.no_grad():
        v = pol.model(g['rotated'].to(cuda_device))
    assert ops.LAUNCHES - l0 == int(g['meta'][7])               # one native launch per GCN layer
    assert_close_scaled(v, g['values'], 1e-5, case + ':values')
    e_ref = float((g['values'].double() - g['values64']).abs().max())
    e_got = float((v.double().cpu() - g['values64']).abs().max())
    assert e_got <= 4.0 * e_ref + 2e-6 * float(g['values64'].abs().max())
    assert np.abs(pol.model.A - g['A0'].numpy()).max() <= 1e-5
    pol.build_action_space(1.0)
    for b in range(int(g['meta'][2])):
        a = pol.predict(joint_state(g, b))
        ref = np.asarray(g['action_values'][b], dtype=np.float64)
        assert np.abs(np.asarray(pol.action_values) - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
        srt = np.sort(ref)[::-1]
        if srt[0] - srt[1] > 2e-5:
            assert a == pol.action_space[int(g['chosen'][b])]
    # training goes through the torch-op statement of the same math
    pol.model.zero_grad()
    pol.model(g['rotated'].to(cuda_device)).sum().backward()
    assert pol.model.w1.grad is not None and pol.model.w_a.grad is not None
