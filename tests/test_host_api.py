"""CPU-side checks (-m "not gpu"): module API / checkpoint compatibility with the reference, C-ABI exports."""
import copy
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from relationalgraphlearning_b200 import _lib
from relationalgraphlearning_b200.config import policy_config
from relationalgraphlearning_b200.graph_model import RGL
from relationalgraphlearning_b200.state_predictor import LinearStatePredictor, StatePredictor, compute_next_robot_state
from relationalgraphlearning_b200.value_estimator import ValueEstimator


def build(seed=0, **kw):
    cfg = policy_config(**kw)
    torch.manual_seed(seed)
    g1 = RGL(cfg, 9, 5)
    ve = ValueEstimator(cfg, g1)
    g2 = RGL(cfg, 9, 5)
    sp = StatePredictor(cfg, g2, 0.25)
    return cfg, g1, ve, g2, sp


def test_library_exports_every_declared_symbol():
    """Every function declared in include/rgl_b200.h is exported by librgl_b200.so (no compute calls here)."""
    hdr = open(os.path.join(ROOT, 'include', 'rgl_b200.h')).read()
    declared = set(re.findall(r'\b(rgl_[a-z_0-9]+)\s*\(', hdr))
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.lib()
    assert lib.rgl_version() == 202
    # FMA section (SURVEY.md 2b parameter inventory + 8 pad floats per 32-wide row, rounded to 64 floats) + the tcgen05
    # operand section (hi/lo tf32 tiles: 2048 + 8192 + 4096 + 2048 + 256 floats)
    assert lib.rgl_packed_graph_floats(2) == ((8256 + 224 * 8 + 63) // 64) * 64 + 16640
    assert lib.rgl_packed_graph_floats(0) == 0 and lib.rgl_packed_graph_floats(5) == 0


def test_argument_validation_without_gpu():
    lib = _lib.lib()
    rc = lib.rgl_graph_forward(None, None, 4, 5, 1, None, 2, 1, None, None, None, None, None, None)
    assert rc == -1 and b'null' in lib.rgl_last_error_string()
    rc = lib.rgl_value_head(None, 4, None, None, None)
    assert rc == -1
    assert lib.rgl_value_head(None, 0, None, None, None) == 0      # empty batch is a no-op


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_same_seed_same_weights_as_reference(seed):
    """Parameter names, shapes and creation order match the reference: torch.manual_seed(s) reproduces the
    exact tensors the reference modules were initialised with (golden fixtures hold the reference's)."""
    g = load_golden('fwd_nh5_s%d' % seed)
    _, g1, ve, g2, sp = build(seed)
    for mod, key in ((g1, 'graph1'), (ve.value_network, 'value'), (g2, 'graph2'), (sp.human_motion_predictor, 'motion')):
        sd = mod.state_dict()
        assert list(sd.keys()) == list(g[key].keys()) or set(sd.keys()) == set(g[key].keys())
        for k in sd:
            assert torch.equal(sd[k], g[key][k]), (key, k)


def test_state_dict_roundtrip_and_counts():
    _, g1, ve, g2, sp = build(3)
    assert sum(p.numel() for p in g1.parameters()) == 8256
    assert sum(p.numel() for p in ve.parameters()) == 22813
    assert sum(p.numel() for p in sp.parameters()) == 10693
    _, h1, we, _, _ = build(4)
    h1.load_state_dict(g1.state_dict())
    we.value_network.load_state_dict(ve.value_network.state_dict())
    for a, b in zip(h1.parameters(), g1.parameters()):
        assert torch.equal(a, b)
    assert [n for n, _ in ve.named_parameters()][:3] == ['graph_model.w_a', 'graph_model.w_r.0.weight', 'graph_model.w_r.0.bias']


def test_deepcopy_gives_independent_target_model():
    """trainer.py:40-41 deep-copies the value estimator as the target network."""
    _, g1, ve, _, _ = build(0)
    tgt = copy.deepcopy(ve)
    assert tgt.graph_model is not ve.graph_model
    assert tgt.graph_model._pack_cache is not ve.graph_model._pack_cache
    with torch.no_grad():
        ve.graph_model.w_a.add_(1.0)
    assert not torch.equal(tgt.graph_model.w_a, ve.graph_model.w_a)


def test_cpu_modules_accept_cpu_tensors_with_the_reference_math(monkeypatch):
    """modules whose parameters live on the CPU accept CPU tensors (torch-op statement of the reference
    math) -- outputs equal the reference-minted golden within fp32 round-off; RGL_FORBID_CPU=1 refuses instead."""
    from conftest import assert_close_scaled, load_golden
    g = load_golden('fwd_nh5_s0')
    _, g1, ve, g2, sp = build(0)
    g1.load_state_dict(g['graph1'])
    ve.value_network.load_state_dict(g['value'])
    g2.load_state_dict(g['graph2'])
    sp.human_motion_predictor.load_state_dict(g['motion'])
    with torch.no_grad():
        assert_close_scaled(g1((g['robot'], g['humans'])), g['H'], 1e-6, 'H')
        assert_close_scaled(ve((g['robot'], g['humans'])), g['V'], 1e-6, 'V')
        nr, S = sp((g['robot'], g['humans']), None)
        assert nr is None
        assert_close_scaled(S, g['S'], 1e-6, 'S')
    assert g1.A is not None and g1.A.shape == (6, 6)
    # gradients flow on the CPU path too (the reference trainer on a CPU policy)
    loss = ve((g['robot'], g['humans'])).sum()
    loss.backward()
    assert ve.graph_model.w_a.grad is not None and ve.value_network[0].weight.grad is not None
    monkeypatch.setenv('RGL_FORBID_CPU', '1')
    with pytest.raises(_lib.RglError):
        g1((g['robot'], g['humans']))
    with pytest.raises(_lib.RglError):
        ve((g['robot'], g['humans']))
    with pytest.raises(_lib.RglError):
        sp((g['robot'], g['humans']), None)


def test_batched_next_robot_state_rule():
    robot = torch.tensor([[[1.5, -2.25, 0.1, 0.2, 0.3, 0.0, 4.0, 1.0, 1.57]], [[0.0, 1.0, 0.0, 0.0, 0.3, 0.0, 4.0, 1.0, 1.57]]])
    out = compute_next_robot_state(robot, (0.3, -0.7), 0.25, 'holonomic')
    for b in range(2):
        exp = robot[b, 0].clone()
        exp[0] = exp[0] + 0.3 * 0.25
        exp[1] = exp[1] + (-0.7) * 0.25
        exp[2] = 0.3
        exp[3] = -0.7
        assert torch.equal(out[b, 0], exp)
    lin = LinearStatePredictor(policy_config(), 0.25)
    hs = torch.rand(2, 3, 5)
    nh = lin.linear_motion_approximator(hs)
    assert torch.equal(nh[..., 0], hs[..., 0] + hs[..., 2]) and torch.equal(nh[..., 4], hs[..., 4])


def test_argument_validation_of_the_round2_entry_points():
    """Planner select / backup, TD loss, replay and communicator entry points reject bad arguments before any launch."""
    lib = _lib.lib()
    assert lib.rgl_plan_select(None, None, 4, 11, 0.97, 2, None, None, None, None, None, None, None) == -1
    assert lib.rgl_plan_select(None, None, 0, 11, 0.97, 2, None, None, None, None, None, None, None) == 0      # empty: no-op
    assert lib.rgl_plan_backup(None, None, None, 4, 2, 0.97, 2, None, None, None) == -1
    assert lib.rgl_plan_expand(None, None, 4, 5, 1, None, 11, 0.25, 7, None, None, None) == -1                  # unknown kinematics / nulls
    assert lib.rgl_td_loss(None, None, None, 8, 0.97, 0.125, None, None, None) == -1
    assert lib.rgl_replay_record_floats(5) == 70 and lib.rgl_replay_record_floats(0) == 0 and lib.rgl_replay_record_floats(32) == 0
    assert lib.rgl_replay_gather(None, None, 4, 5, None, None, None, None, None, None, None) == -1
    h = ctypes.c_void_p()
    assert lib.rgl_comm_create(3, 2, 100, ctypes.byref(h)) == -1          # rank outside the world
    assert lib.rgl_comm_create(0, 64, 100, ctypes.byref(h)) == -1         # world too large
    assert b'rgl_comm_create' in lib.rgl_comm_last_error_string()
    assert lib.rgl_comm_handle_bytes() == 64
    assert lib.rgl_comm_destroy(None) == 0


def test_argument_validation_of_the_training_backward_entry_points():
    """The staged attention / similarity backward and the fused embedding-MLP backward reject bad arguments before any launch."""
    lib = _lib.lib()
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.POINTER(ctypes.c_float))
    # rgl_attn_sim_bwd(A, Z, gM, mask, up_rows, gA_in, gZ, gA_out, X, Y, gY, gX, gx_accumulate, B, n, stream)
    assert lib.rgl_attn_sim_bwd(None, None, None, None, 1, None, None, None, None, None, None, None, 0, 4, 6, None) == -1      # nulls
    assert lib.rgl_attn_sim_bwd(None, None, None, None, 1, None, None, None, None, None, None, None, 0, 0, 6, None) == 0       # empty: no-op
    assert lib.rgl_attn_sim_bwd(p, p, p, None, 7, None, p, p, None, None, None, None, 0, 4, 6, None) == -1                     # up_rows > n
    assert lib.rgl_attn_sim_bwd(p, p, p, None, 0, None, p, p, None, None, None, None, 0, 4, 6, None) == -1                     # up_rows < 1
    assert lib.rgl_attn_sim_bwd(p, p, p, None, 6, None, p, None, None, None, None, None, 0, 4, 6, None) == -1                  # no X and no gA_out
    assert lib.rgl_attn_sim_bwd(p, p, p, None, 6, None, p, None, p, None, None, None, 0, 4, 6, None) == -1                     # X without Y / gY / gX
    assert lib.rgl_attn_sim_bwd(p, p, p, None, 6, None, p, p, None, None, None, None, 0, 4, 33, None) == -1                    # n > 32
    # rgl_attn_layer_bwd(..., mask, up_rows, stream)
    assert lib.rgl_attn_layer_bwd(p, p, p, None, 0, p, p, 0, 4, 6, None, 0, None) == -1                                         # up_rows < 1
    # rgl_mlp2_bwd(G, mask, hidden, W1, X0, K0, dW1, db1, dW0, db0, R, stream)
    r = _lib.Rows()
    r.ptr, r.ld, r.rows_per_group, r.group_stride = ctypes.addressof(buf), 32, 0, 0
    assert lib.rgl_mlp2_bwd(None, None, None, None, None, 5, None, None, None, None, 8, None) == -1                            # nulls
    assert lib.rgl_mlp2_bwd(None, None, None, None, None, 5, None, None, None, None, 0, None) == 0                             # empty: no-op
    assert lib.rgl_mlp2_bwd(ctypes.byref(r), None, ctypes.byref(r), p, ctypes.byref(r), 17, p, p, p, p, 8, None) == -4          # K0 > 16: unsupported
    assert lib.rgl_mlp2_bwd(ctypes.byref(r), None, ctypes.byref(r), p, ctypes.byref(r), 0, p, p, p, p, 8, None) == -4
