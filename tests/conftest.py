import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
    """Golden fixture minted from the reference by oracle/gen_golden.py -> dict of tensors + state-dicts."""
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    out = {'graph1': {}, 'value': {}, 'graph2': {}, 'motion': {}, 'model': {}}
    for k in z.files:
        if '/' in k:
            grp, key = k.split('/', 1)
            out[grp][key] = torch.from_numpy(z[k])
        elif z[k].dtype.kind == 'U':
            out[k] = str(z[k])
        else:
            out[k] = torch.from_numpy(z[k]) if z[k].dtype.kind == 'f' else z[k]
    return out


FWD_CASES = ['fwd_nh5_s0', 'fwd_nh5_s1', 'fwd_nh5_s2', 'fwd_nh5_b1', 'fwd_nh10_s0', 'fwd_nh20_s0', 'fwd_nh1_s0',
             'fwd_nh5_trained', 'fwd_nh5_adversarial', 'fwd_nh5_layerwise_noskip', 'fwd_nh5_layerwise_skip']


SIM_CASES = ['fwd_nh5_sim_' + s for s in ('gaussian', 'cosine', 'cosine_softmax', 'concatenation', 'squared', 'equal_attention', 'diagonal')]


def graph_kw(g):
    kw = dict(layerwise_graph=bool(g['meta'][3]), skip_connection=bool(g['meta'][4]))
    if 'similarity' in g:
        kw['similarity_function'] = g['similarity']
    return kw


def assert_close_scaled(got, ref, rel=1e-5, name=''):
    """SURVEY.md §8(c) comparator: |x - ref| <= rel * max(|ref|, s), s = per-tensor scale max|ref| (>= 1e-3)."""
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    s = max(float(ref.abs().max()), 1e-3)
    err = (got - ref).abs()
    bound = rel * torch.clamp(ref.abs(), min=s)
    worst = float((err / bound).max()) if err.numel() else 0.0
    assert worst <= 1.0, '%s: max err %.3e at scale %.3e (%.2fx the %.0e bound)' % (name, float(err.max()), s, worst, rel)
    return float(err.max()) / s


@pytest.fixture(scope='session')
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')
