"""Host-buffer streaming front end (-m gpu): results through HostStream (pinned host -> H2D -> kernels -> D2H) equal
the direct device call, for graph-replayed and eager submissions."""
import pytest
import torch

from conftest import load_golden
from relationalgraphlearning_b200.hostio import HostStream
from test_gpu_parity import modules_from_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('kind', ['graph', 'value', 'statepred'])
def test_hoststream_matches_direct_call(kind, cuda_device):
    g = load_golden('fwd_nh5_s0')
    g1, ve, g2, sp = modules_from_golden(g, cuda_device)
    module = {'graph': g1, 'value': ve, 'statepred': sp}[kind]
    B = 64
    hs = HostStream(kind, module, B, 5, cuda_device, depth=3)
    batches = []
    for i in range(5):
        r = (g['robot'] + 0.01 * i).pin_memory()
        h = (g['humans'] - 0.02 * i).pin_memory()
        batches.append((r, h))
    results = {}
    for rep in range(3):                      # second/third pass replay the captured per-buffer graphs
        for i, (r, h) in enumerate(batches):
            slot = hs.submit(r, h)
            results[(rep, i)] = hs.result(slot).clone()
    hs.drain()
    with torch.no_grad():
        for i, (r, h) in enumerate(batches):
            rd, hd = r.to(cuda_device), h.to(cuda_device)
            if kind == 'graph':
                ref = g1((rd, hd))
            elif kind == 'value':
                ref = ve((rd, hd))
            else:
                ref = sp((rd, hd), None)[1]
            for rep in range(3):
                assert torch.equal(results[(rep, i)], ref.cpu()), (kind, rep, i)
    # unpinned host tensors take the eager path and still give the same answer
    slot = hs.submit(batches[0][0].clone(), batches[0][1].clone())
    assert torch.equal(hs.result(slot), results[(0, 0)])
    assert hs.h2d_bytes == B * 136 and hs.d2h_bytes == results[(0, 0)].numel() * 4


@pytest.mark.parametrize('kind', ['graph', 'value', 'statepred'])
def test_hoststream_sees_weight_updates_between_submits(kind, cuda_device):
    """A replayed per-buffer graph must use the CURRENT weights: the packed blobs are refreshed before every replay
    (optimizer step / load_state_dict between submissions)."""
    g = load_golden('fwd_nh5_s1')
    g1, ve, g2, sp = modules_from_golden(g, cuda_device)
    module = {'graph': g1, 'value': ve, 'statepred': sp}[kind]
    hs = HostStream(kind, module, 64, 5, cuda_device, depth=2)
    r, h = g['robot'].pin_memory(), g['humans'].pin_memory()

    def direct():
        with torch.no_grad():
            rd, hd = r.to(cuda_device), h.to(cuda_device)
            if kind == 'graph':
                return g1((rd, hd)).cpu()
            return (ve((rd, hd)) if kind == 'value' else sp((rd, hd), None)[1]).cpu()

    for _ in range(4):                        # both slots have captured graphs for this buffer pair
        before = hs.result(hs.submit(r, h)).clone()
    assert torch.equal(before, direct())
    with torch.no_grad():                     # in-place update, as an optimizer step does (bumps the version counters)
        for p in module.parameters():
            p.mul_(0.9)
    after = hs.result(hs.submit(r, h)).clone()
    assert not torch.equal(after, before)
    assert torch.equal(after, direct())
    assert hs.overwritten == 0
