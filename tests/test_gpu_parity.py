"""GPU parity (-m gpu): the sm_100a kernels, called through the C ABI (ctypes), against
  (1) the golden vectors minted from the reference (tests/golden), fp32 and fp64, and
  (2) the CPU oracle on fresh seeded inputs, plus size-independent properties at full batch sizes.
Tolerance: |x - ref| <= 1e-5 * max(|ref|, max|ref|)  (north_star: 1e-5 rel fp32; SURVEY.md 8(c) floor)."""
import numpy as np
import pytest
import torch

from conftest import FWD_CASES, SIM_CASES, assert_close_scaled, graph_kw, load_golden
from oracle import rgl_oracle as O
from relationalgraphlearning_b200 import _lib, ops
from relationalgraphlearning_b200.config import policy_config
from relationalgraphlearning_b200.graph_model import RGL
from relationalgraphlearning_b200.state_predictor import StatePredictor
from relationalgraphlearning_b200.synthetic import synthetic_states
from relationalgraphlearning_b200.value_estimator import ValueEstimator

pytestmark = pytest.mark.gpu
REL = 1e-5


def modules_from_golden(g, dev):
    kw = graph_kw(g)
    cfg = policy_config(layerwise_graph=kw['layerwise_graph'], skip_connection=kw['skip_connection'],
                        similarity_function=kw.get('similarity_function', 'embedded_gaussian'))
    g1 = RGL(cfg, 9, 5)
    ve = ValueEstimator(cfg, g1)
    g2 = RGL(cfg, 9, 5)
    sp = StatePredictor(cfg, g2, 0.25)
    g1.load_state_dict(g['graph1'])
    ve.value_network.load_state_dict(g['value'])
    g2.load_state_dict(g['graph2'])
    sp.human_motion_predictor.load_state_dict(g['motion'])
    for m in (ve, sp):
        m.to(dev)
    return g1, ve, g2, sp


@pytest.mark.parametrize('case', FWD_CASES)
def test_forward_matches_reference_golden(case, cuda_device):
    g = load_golden(case)
    g1, ve, g2, sp = modules_from_golden(g, cuda_device)
    robot, humans = g['robot'].to(cuda_device), g['humans'].to(cuda_device)
    with torch.no_grad():
        H = g1((robot, humans))
        V = ve((robot, humans))
        nr, S = sp((robot, humans), None)
    torch.cuda.synchronize()
    assert nr is None
    assert_close_scaled(H, g['H'], REL, case + ':H')
    assert_close_scaled(V, g['V'], REL, case + ':V')
    assert_close_scaled(S, g['S'], REL, case + ':S')
    if g['A0'].numel():
        assert_close_scaled(torch.from_numpy(g1.A), g['A0'], REL, case + ':A0')
    else:
        assert g1.A is None
    # second criterion of SURVEY.md 8(c): not further from the fp64 truth than k x the reference's own fp32 error
    for got, r32, r64, nm in ((H, g['H'], g['H64'], 'H'), (V, g['V'], g['V64'], 'V'), (S, g['S'], g['S64'], 'S')):
        scale = float(r64.abs().max())
        e_ref = float((r32.double() - r64).abs().max())
        e_got = float((got.double().cpu() - r64).abs().max())
        assert e_got <= 4.0 * e_ref + 2e-6 * scale, (case, nm, e_got, e_ref)


@pytest.mark.parametrize('case', SIM_CASES)
def test_other_similarity_functions_match_reference_golden(case, cuda_device):
    """The seven similarity functions the fused kernel is not specialised for (graph_model.py:67-93) run as torch ops on
    the GPU; fixtures minted from the reference modules."""
    g = load_golden(case)
    g1, ve, g2, sp = modules_from_golden(g, cuda_device)
    assert not g1.kernel_supported()
    robot, humans = g['robot'].to(cuda_device), g['humans'].to(cuda_device)
    with torch.no_grad():
        H, V, S = g1((robot, humans)), ve((robot, humans)), sp((robot, humans), None)[1]
    assert_close_scaled(H, g['H'], REL, case + ':H')
    assert_close_scaled(V, g['V'], REL, case + ':V')
    assert_close_scaled(S, g['S'], REL, case + ':S')
    # ... and train (autograd through the same torch ops)
    ve((robot, humans)).sum().backward()
    assert all(p.grad is not None for p in ve.value_network.parameters())


@pytest.mark.parametrize('nh', [0, 32, 40])
def test_human_counts_outside_the_kernel_range(nh, cuda_device):
    """Nh = 0 or > 31: the modules still compute (torch ops on the GPU), like the reference does for any Nh."""
    g = load_golden('fwd_nh5_s0')
    g1, ve, g2, sp = modules_from_golden(g, cuda_device)
    robot, humans = synthetic_states(9, max(nh, 1), seed=5)
    humans = humans[:, :nh]
    with torch.no_grad():
        H, V, S = g1((robot.to(cuda_device), humans.to(cuda_device))), ve((robot.to(cuda_device), humans.to(cuda_device))), \
            sp((robot.to(cuda_device), humans.to(cuda_device)), None)[1]
        assert_close_scaled(H, O.rgl_forward(g['graph1'], robot, humans), REL, 'H')
        assert_close_scaled(V, O.value_forward(g['graph1'], g['value'], robot, humans), REL, 'V')
        assert S.shape == (9, nh, 5)
        if nh:
            assert_close_scaled(S, O.statepred_forward(g['graph2'], g['motion'], robot, humans), REL, 'S')


def test_c2_every_state_against_the_oracle_with_error_histogram(cuda_device, capsys):
    """BASELINE C2 (B = 4096, Nh = 5): ALL 4096 states against the CPU oracle, under the scaled bound of the comparator AND
    reported element-wise: the distribution of |x - ref| / max(|ref|, 1e-3 * scale) (what '1e-5 rel' means element by
    element; elements near zero are measured against 1e-3 of the tensor scale)."""
    g = load_golden('fwd_nh5_s0')
    g1, ve, g2, sp = modules_from_golden(g, cuda_device)
    robot, humans = synthetic_states(4096, 5, seed=1234)
    with torch.no_grad():
        H, V, S = g1((robot.to(cuda_device), humans.to(cuda_device))), ve((robot.to(cuda_device), humans.to(cuda_device))), \
            sp((robot.to(cuda_device), humans.to(cuda_device)), None)[1]
        torch.set_num_threads(8)
        Ho = O.rgl_forward(g['graph1'], robot, humans)
        Vo = O.value_forward(g['graph1'], g['value'], robot, humans)
        So = O.statepred_forward(g['graph2'], g['motion'], robot, humans)
    lines = []
    for nm, got, ref in (('H', H, Ho), ('V', V, Vo), ('S', S, So)):
        assert_close_scaled(got, ref, REL, 'C2 all states ' + nm)
        got, ref = got.double().cpu(), ref.double()
        scale = float(ref.abs().max())
        rel = (got - ref).abs() / torch.clamp(ref.abs(), min=1e-3 * scale)
        qs = torch.quantile(rel.flatten()[:: max(1, rel.numel() // 1000000)], torch.tensor([0.5, 0.9, 0.99, 0.999], dtype=torch.float64))
        frac = float((rel <= 1e-5).double().mean())
        lines.append('%s: scale %.3g  max|err|/scale %.2e  elementwise rel p50 %.1e p90 %.1e p99 %.1e p99.9 %.1e max %.1e  frac<=1e-5 %.5f' %
                     (nm, scale, float((got - ref).abs().max()) / scale, *[float(q) for q in qs], float(rel.max()), frac))
        assert float(qs[0]) <= 1e-5, lines[-1]              # the typical element is within 1e-5 element-wise; the tail is reported
    with capsys.disabled():
        print('\n[C2 error histogram, 4096 states vs oracle]\n' + '\n'.join(lines))
    import os
    try:
        os.makedirs(os.path.join(os.path.dirname(__file__), '..', 'gpurun_out'), exist_ok=True)
        with open(os.path.join(os.path.dirname(__file__), '..', 'gpurun_out', 'c2_error_histogram.txt'), 'w') as f:
            f.write('\n'.join(lines) + '\n')
    except OSError:
        pass


@pytest.mark.parametrize('B', [0, 1, 2, 31, 33, 100, 257])
def test_ragged_batches_and_partial_tiles(B, cuda_device):
    g = load_golden('fwd_nh5_s1')
    g1, ve, g2, sp = modules_from_golden(g, cuda_device)
    robot, humans = synthetic_states(max(B, 1), 5, seed=99)
    robot, humans = robot[:B], humans[:B]
    with torch.no_grad():
        H = g1((robot.to(cuda_device), humans.to(cuda_device)))
        V = ve((robot.to(cuda_device), humans.to(cuda_device)))
        S = sp((robot.to(cuda_device), humans.to(cuda_device)), None)[1]
        Ho = O.rgl_forward(g['graph1'], robot, humans)
        Vo = O.value_forward(g['graph1'], g['value'], robot, humans)
        So = O.statepred_forward(g['graph2'], g['motion'], robot, humans)
    assert H.shape == (B, 6, 32) and V.shape == (B, 1) and S.shape == (B, 5, 5)
    if B:
        assert_close_scaled(H, Ho, REL, 'H')
        assert_close_scaled(V, Vo, REL, 'V')
        assert_close_scaled(S, So, REL, 'S')


def test_unaligned_inputs_take_the_non_tma_path(cuda_device):
    """A storage offset of one float breaks the 16-byte alignment TMA needs; results must not change."""
    g = load_golden('fwd_nh5_s0')
    g1, ve, _, _ = modules_from_golden(g, cuda_device)
    robot, humans = g['robot'].to(cuda_device), g['humans'].to(cuda_device)
    rbuf = torch.zeros(robot.numel() + 1, device=cuda_device)
    hbuf = torch.zeros(humans.numel() + 1, device=cuda_device)
    rbuf[1:] = robot.reshape(-1)
    hbuf[1:] = humans.reshape(-1)
    r2, h2 = rbuf[1:].view_as(robot), hbuf[1:].view_as(humans)
    assert r2.data_ptr() % 16 != 0
    with torch.no_grad():
        assert torch.equal(ve((robot, humans)), ve((r2, h2)))
        assert torch.equal(g1((robot, humans)), g1((r2, h2)))


def test_humans_broadcast_matches_materialised_batch(cuda_device):
    """Planner layout: A actions share one human set (humans_bcast=A)."""
    g = load_golden('fwd_nh5_s0')
    _, ve, _, sp = modules_from_golden(g, cuda_device)
    E, A = 7, 11
    robot, humans = synthetic_states(E * A, 5, seed=5)
    humans = humans[:E]
    robot, humans = robot.to(cuda_device), humans.to(cuda_device)
    full = humans.repeat_interleave(A, dim=0)
    with torch.no_grad():
        assert torch.equal(ve.run(robot, humans, humans_bcast=A), ve.run(robot, full))
        assert torch.equal(sp.run(robot, humans, humans_bcast=A), sp.run(robot, full))


@pytest.mark.parametrize('nh', [1, 2, 3, 7, 10, 15, 20])
def test_human_counts_against_oracle(nh, cuda_device):
    g = load_golden('fwd_nh5_s2')
    g1, ve, g2, sp = modules_from_golden(g, cuda_device)
    robot, humans = synthetic_states(70, nh, seed=nh)
    with torch.no_grad():
        H = g1((robot.to(cuda_device), humans.to(cuda_device)))
        V = ve((robot.to(cuda_device), humans.to(cuda_device)))
        S = sp((robot.to(cuda_device), humans.to(cuda_device)), None)[1]
        assert_close_scaled(H, O.rgl_forward(g['graph1'], robot, humans), REL, 'H')
        assert_close_scaled(V, O.value_forward(g['graph1'], g['value'], robot, humans), REL, 'V')
        assert_close_scaled(S, O.statepred_forward(g['graph2'], g['motion'], robot, humans), REL, 'S')


@pytest.mark.parametrize('num_layer', [1, 3])
def test_layer_counts(num_layer, cuda_device):
    cfg = policy_config(num_layer=num_layer)
    torch.manual_seed(11)
    g1 = RGL(cfg, 9, 5)
    ve = ValueEstimator(cfg, g1)
    sd_g = {k: v.clone() for k, v in g1.state_dict().items()}
    sd_v = {k: v.clone() for k, v in ve.value_network.state_dict().items()}
    ve.to(cuda_device)
    robot, humans = synthetic_states(65, 5, seed=3)
    with torch.no_grad():
        assert_close_scaled(g1((robot.to(cuda_device), humans.to(cuda_device))), O.rgl_forward(sd_g, robot, humans), REL, 'H')
        assert_close_scaled(ve((robot.to(cuda_device), humans.to(cuda_device))), O.value_forward(sd_g, sd_v, robot, humans), REL, 'V')


@pytest.mark.parametrize('given_A', [True, False])
@pytest.mark.parametrize('n,B', [(6, 64), (6, 1000), (11, 77), (21, 33), (2, 5)])
def test_gcn_layer_kernel(n, B, given_A, cuda_device):
    gen = torch.Generator().manual_seed(n * 1000 + B)
    X = torch.randn(B, n, 32, generator=gen)
    W = torch.randn(32, 32, generator=gen)
    wa = torch.randn(32, 32, generator=gen) * 0.2
    A = torch.softmax(torch.matmul(torch.matmul(X, wa), X.permute(0, 2, 1)), dim=2)
    for skip in (False, True):
        ref = torch.relu(torch.matmul(torch.matmul(A, X), W))
        if skip:
            ref = ref + X
        if given_A:
            out = ops.gcn_layer(X.to(cuda_device), W.to(cuda_device), A=A.to(cuda_device), skip=skip)
        else:
            out, Aout = ops.gcn_layer(X.to(cuda_device), W.to(cuda_device), w_a=wa.to(cuda_device), skip=skip, return_A=True)
            assert_close_scaled(Aout, A, REL, 'A')
        assert_close_scaled(out, ref, REL, 'gcn n=%d B=%d' % (n, B))


def test_full_size_properties_c2(cuda_device):
    """BASELINE config C2 (B=4096, Nh=5): properties that need no oracle at that size."""
    g = load_golden('fwd_nh5_s0')
    g1, ve, g2, sp = modules_from_golden(g, cuda_device)
    B = 4096
    robot, humans = synthetic_states(B, 5, seed=1234, device=cuda_device)
    with torch.no_grad():
        H, V, S = g1((robot, humans)), ve((robot, humans)), sp((robot, humans), None)[1]
        # (1) states are independent: any batch permutation / split gives bit-identical rows
        perm = torch.randperm(B, device=cuda_device)
        assert torch.equal(ve((robot[perm], humans[perm])), V[perm])
        assert torch.equal(g1((robot[perm], humans[perm])), H[perm])
        # chunks run with other groups-per-CTA / tile boundaries than the full batch: equal within the parity tolerance
        # (rows are independent, so in practice bit-identical); the fp32-FMA path agrees with the 3xTF32 path
        parts = [ve((robot[i:i + 1000], humans[i:i + 1000])) for i in range(0, B, 1000)]
        assert_close_scaled(torch.cat(parts), V, REL, 'V split')
        g1.fp32_fma = True
        Vf = ve((robot, humans))
        parts = [ve((robot[i:i + 2048], humans[i:i + 2048])) for i in range(0, B, 2048)]
        assert_close_scaled(torch.cat(parts), Vf, REL, 'V split fp32')
        assert_close_scaled(Vf, V, REL, 'fp32 FFMA vs 3xTF32')
        g1.fp32_fma = False
        # (2) the value head sees only the robot row: E from the H path equals the E-only path
        E = g1.run(robot, humans, want_E=True)['E']
        assert_close_scaled(E, H[:, 0, :], REL, 'E vs H[:,0]')      # (the E-only path may run another kernel variant than the H path)
        # (3) permuting humans permutes the predicted humans and leaves V unchanged up to summation order
        hp = torch.tensor([3, 0, 4, 1, 2], device=cuda_device)
        S2 = sp((robot, humans[:, hp]), None)[1]
        assert_close_scaled(S2, S[:, hp], REL, 'S perm')
        assert_close_scaled(ve((robot, humans[:, hp])), V, REL, 'V perm')
        # (4) spot-check 128 states against the oracle
        idx = torch.arange(0, B, 32)
        rc, hc = robot[idx].cpu(), humans[idx].cpu()
        assert_close_scaled(V[idx], O.value_forward(g['graph1'], g['value'], rc, hc), REL, 'V spot')
        assert_close_scaled(H[idx], O.rgl_forward(g['graph1'], rc, hc), REL, 'H spot')
        assert torch.isfinite(H).all() and torch.isfinite(S).all()



@pytest.mark.parametrize('case,nh,B', [('fwd_nh10_s0', 10, 8192), ('fwd_nh20_s0', 20, 16384)])
def test_full_size_properties_c4_c5_shapes(case, nh, B, cuda_device):
    """BASELINE configs C4 / C5 shapes (Nh = 10, B = 8192; Nh = 20, B = 16384): size-independent properties + oracle spot checks."""
    g = load_golden(case)
    g1, ve, g2, sp = modules_from_golden(g, cuda_device)
    robot, humans = synthetic_states(B, nh, seed=4321 + nh, device=cuda_device)
    with torch.no_grad():
        H, V, S = g1((robot, humans)), ve((robot, humans)), sp((robot, humans), None)[1]
        assert torch.isfinite(H).all() and torch.isfinite(V).all() and torch.isfinite(S).all()
        # states are independent: a batch permutation permutes the rows bit for bit
        perm = torch.randperm(B, device=cuda_device)
        assert torch.equal(ve((robot[perm], humans[perm])), V[perm])
        assert torch.equal(g1((robot[perm], humans[perm])), H[perm])
        # ragged split (tile boundaries move)
        cut = B // 3 + 7
        parts = torch.cat([g1((robot[:cut], humans[:cut])), g1((robot[cut:], humans[cut:]))])
        assert_close_scaled(parts, H, REL, 'H split')
        # the value head sees only the robot row; the E-only path (last layer for the robot rows only) matches the H path
        assert_close_scaled(g1.run(robot, humans, want_E=True)['E'], H[:, 0, :], REL, 'E vs H[:,0]')
        # humans_bcast: every state of a group of 4 reads the humans of the group's first state
        hb = humans[::4].contiguous()
        Vb = ve.run(robot, hb, humans_bcast=4)
        Vm = ve((robot, hb.repeat_interleave(4, dim=0)))
        assert torch.equal(Vb.view(-1), Vm.view(-1))
        # permuting the humans permutes S and leaves V unchanged up to summation order
        hp = torch.randperm(nh, device=cuda_device)
        assert_close_scaled(sp((robot, humans[:, hp]), None)[1], S[:, hp], REL, 'S perm')
        assert_close_scaled(ve((robot, humans[:, hp])), V, REL, 'V perm')
        # oracle spot check on 64 states
        idx = torch.arange(0, B, B // 64)
        rc, hc = robot[idx].cpu(), humans[idx].cpu()
        assert_close_scaled(V[idx], O.value_forward(g['graph1'], g['value'], rc, hc), REL, 'V spot')
        assert_close_scaled(H[idx], O.rgl_forward(g['graph1'], rc, hc), REL, 'H spot')
        assert_close_scaled(S[idx], O.statepred_forward(g['graph2'], g['motion'], rc, hc), REL, 'S spot')


def test_weight_update_invalidates_packed_blob(cuda_device):
    g = load_golden('fwd_nh5_s0')
    g1, ve, _, _ = modules_from_golden(g, cuda_device)
    robot, humans = g['robot'].to(cuda_device), g['humans'].to(cuda_device)
    with torch.no_grad():
        v0 = ve((robot, humans)).clone()
        ve.value_network[6].bias.add_(0.5)
        v1 = ve((robot, humans))
        assert_close_scaled(v1, v0 + 0.5, REL, 'bias shift')
        g1.w_a.mul_(0.5)
        sd = {k: v.cpu() for k, v in g1.state_dict().items()}
        sv = {k: v.cpu() for k, v in ve.value_network.state_dict().items()}
        assert_close_scaled(ve((robot, humans)), O.value_forward(sd, sv, g['robot'], g['humans']), REL, 'after w_a update')


def test_autograd_matches_oracle_gradients(cuda_device):
    """trainer.py:122-131 semantics: MSE loss gradients w.r.t. every parameter (out-of-place skip add)."""
    g = load_golden('fwd_nh5_s1')
    g1, ve, g2, sp = modules_from_golden(g, cuda_device)
    robot, humans = g['robot'], g['humans']
    target = torch.linspace(-0.2, 0.8, robot.size(0)).unsqueeze(1)
    loss = torch.nn.functional.mse_loss(ve((robot.to(cuda_device), humans.to(cuda_device))), target.to(cuda_device))
    loss.backward()
    sd_g = {k: v.clone().requires_grad_(True) for k, v in g['graph1'].items()}
    sd_v = {k: v.clone().requires_grad_(True) for k, v in g['value'].items()}
    lo = torch.nn.functional.mse_loss(O.value_forward(sd_g, sd_v, robot, humans), target)
    lo.backward()
    assert abs(float(loss) - float(lo)) <= 1e-5 * max(1.0, abs(float(lo)))
    for name, p in ve.graph_model.named_parameters():
        assert_close_scaled(p.grad, sd_g[name].grad, 1e-4, 'grad ' + name)
    for name, p in ve.value_network.named_parameters():
        assert_close_scaled(p.grad, sd_v[name].grad, 1e-4, 'grad value ' + name)
    # state predictor, detach=True: graph parameters get no gradient (state_predictor.py:29-30)
    out = sp((robot.to(cuda_device), humans.to(cuda_device)), None, detach=True)[1]
    out.sum().backward()
    assert all(p.grad is None for p in sp.graph_model.parameters())
    assert all(p.grad is not None for p in sp.human_motion_predictor.parameters())
