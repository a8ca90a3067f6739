"""GPU parity of every launch shape of the tcgen05 graph kernel (-m gpu).

The number of 128-row groups per CTA (1 / 2 / 4) is normally chosen from the batch size and the requested outputs; the
experiment switch RGL_TC_GROUPS forces it, RGL_GRAPH_VARIANT selects the legacy fp32-FMA / mma.sync kernels, and
RGL_VALUE_VARIANT / RGL_TC_VALUE_GROUPS do the same for the value head; RGL_TC_TMA_OUT=0 selects the staged copy-out of H
instead of the TMA tensor stores.  They are read once per process, so each variant runs in a child process: graph / value / state-predictor outputs against the CPU
oracle on batches large enough that every group loops over several tiles, with a ragged last tile.
Tolerance: |x - ref| <= 1e-5 * max(|ref|, max|ref|)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, torch
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(root)r + '/tests')
from conftest import assert_close_scaled
from oracle import rgl_oracle as O
from relationalgraphlearning_b200.config import policy_config
from relationalgraphlearning_b200.graph_model import RGL
from relationalgraphlearning_b200.state_predictor import StatePredictor
from relationalgraphlearning_b200.synthetic import synthetic_states
from relationalgraphlearning_b200.value_estimator import ValueEstimator
dev = torch.device('cuda:0')
for nh, B, lw, skip in ((5, 20011, False, True), (1, 3000, False, True), (10, 9001, False, True), (20, 5003, False, True),
                        (3, 777, True, True), (7, 1234, False, False)):
    cfg = policy_config(layerwise_graph=lw, skip_connection=skip)
    torch.manual_seed(nh)
    g1 = RGL(cfg, 9, 5); ve = ValueEstimator(cfg, g1); g2 = RGL(cfg, 9, 5); sp = StatePredictor(cfg, g2, 0.25)
    sds = [{k: v.clone() for k, v in m.state_dict().items()} for m in (g1, ve.value_network, g2, sp.human_motion_predictor)]
    ve.to(dev); sp.to(dev)
    robot, humans = synthetic_states(B, nh, seed=nh + 40)
    with torch.no_grad():
        H = g1((robot.to(dev), humans.to(dev)))
        V = ve((robot.to(dev), humans.to(dev)))
        S = sp((robot.to(dev), humans.to(dev)), None)[1]
        torch.cuda.synchronize()
        kw = dict(layerwise_graph=lw, skip_connection=skip)
        Ho = O.rgl_forward(sds[0], robot, humans, **kw)
        Vo = O.value_forward(sds[0], sds[1], robot, humans, **kw)
        So = O.statepred_forward(sds[2], sds[3], robot, humans, **kw)
    tag = 'nh%%d_B%%d' %% (nh, B)
    assert_close_scaled(H, Ho, 1e-5, tag + ':H')
    assert_close_scaled(V, Vo, 1e-5, tag + ':V')
    assert_close_scaled(S, So, 1e-5, tag + ':S')
# stand-alone GCN layer (A given / computed in-kernel), several tiles per group and a ragged tail
from relationalgraphlearning_b200 import ops
for n, B in ((6, 30011), (11, 4001), (21, 4444), (2, 70001), (8, 515)):
    g = torch.Generator().manual_seed(n)
    X = torch.randn(B, n, 32, generator=g); W = torch.randn(32, 32, generator=g); wa = torch.randn(32, 32, generator=g) * 0.2
    A = torch.softmax(torch.randn(B, n, n, generator=g), dim=2)
    for skip in (False, True):
        ref = torch.relu(torch.matmul(torch.matmul(A, X), W)) + (X if skip else 0)
        got = ops.gcn_layer(X.to(dev), W.to(dev), A=A.to(dev), skip=skip)
        assert_close_scaled(got, ref, 1e-5, 'gcn_A n%%d skip%%d' %% (n, skip))
    Aref = torch.softmax(torch.matmul(torch.matmul(X, wa), X.transpose(1, 2)), dim=2)
    ref = torch.relu(torch.matmul(torch.matmul(Aref, X), W)) + X
    got, Agot = ops.gcn_layer(X.to(dev), W.to(dev), w_a=wa.to(dev), skip=True, return_A=True)
    assert_close_scaled(got, ref, 1e-5, 'gcn_wa n%%d' %% n)
    assert_close_scaled(Agot, Aref, 1e-5, 'gcn_wa A n%%d' %% n)
print('variant ok')
'''


@pytest.mark.parametrize('env', [{'RGL_TC_GROUPS': '1'}, {'RGL_TC_GROUPS': '2'}, {'RGL_TC_GROUPS': '4'}, {},
                                 {'RGL_GRAPH_VARIANT': 'm'}, {'RGL_GRAPH_VARIANT': '4', 'RGL_GCN_VARIANT': 'f'},
                                 {'RGL_VALUE_VARIANT': 't', 'RGL_TC_VALUE_GROUPS': '1'}, {'RGL_VALUE_VARIANT': 't', 'RGL_TC_VALUE_GROUPS': '2'},
                                 {'RGL_VALUE_VARIANT': 'f'}, {'RGL_TC_TMA_OUT': '0'}],
                         ids=['tc_g1', 'tc_g2', 'tc_g4', 'tc_auto', 'legacy_mma', 'legacy_ffma', 'value_tc_g1', 'value_tc_g2', 'value_ffma', 'staged_copy_out'])
def test_kernel_variant_against_oracle(env):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    e = dict(os.environ)
    e.pop('RGL_TC_GROUPS', None)
    e.pop('RGL_GRAPH_VARIANT', None)
    e.pop('RGL_VALUE_VARIANT', None)
    e.pop('RGL_GCN_VARIANT', None)
    e.pop('RGL_TC_TMA_OUT', None)
    e.pop('RGL_TC_VALUE_GROUPS', None)
    e.update(env)
    res = subprocess.run([sys.executable, '-c', CHILD % {'root': ROOT}], env=e, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0 and 'variant ok' in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
