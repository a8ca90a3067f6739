"""Native training step (-m gpu): fused forward with saves + hand-written backward vs oracle autograd gradients
(the oracle is the reference math with the out-of-place skip add; crowd_nav/utils/trainer.py:122-131)."""
import pytest
import torch

from conftest import assert_close_scaled, load_golden
from oracle import rgl_oracle as O
from relationalgraphlearning_b200 import ops, training
from relationalgraphlearning_b200.config import policy_config
from relationalgraphlearning_b200.graph_model import RGL
from relationalgraphlearning_b200.synthetic import synthetic_states
from relationalgraphlearning_b200.value_estimator import ValueEstimator

pytestmark = pytest.mark.gpu


def build(seed, dev, **kw):
    cfg = policy_config(**kw)
    torch.manual_seed(seed)
    g = RGL(cfg, 9, 5)
    ve = ValueEstimator(cfg, g)
    sd_g = {k: v.clone() for k, v in g.state_dict().items()}
    sd_v = {k: v.clone() for k, v in ve.value_network.state_dict().items()}
    ve.to(dev)
    return ve, sd_g, sd_v


@pytest.mark.parametrize('nh,B,kw', [(5, 64, {}), (5, 100, {}), (10, 257, {}), (20, 33, {}), (3, 50, {}), (10, 2048, {}),
                                     (5, 64, dict(num_layer=1)), (5, 64, dict(num_layer=3)), (5, 96, dict(skip_connection=False))])
def test_native_backward_matches_oracle_autograd(nh, B, kw, cuda_device):
    ve, sd_g, sd_v = build(7, cuda_device, **kw)
    assert training.native_supported(ve)
    robot, humans = synthetic_states(B, nh, seed=B)
    target = torch.linspace(-0.25, 1.0, B).unsqueeze(1)
    before = ops.LAUNCHES
    out = ve((robot.to(cuda_device), humans.to(cuda_device)))
    loss = torch.nn.functional.mse_loss(out, target.to(cuda_device))
    loss.backward()
    assert ops.LAUNCHES - before >= 10          # the hand-written kernels ran (no torch-op recompute)
    gkw = dict(skip_connection=kw.get('skip_connection', True))
    pg = {k: v.clone().requires_grad_(True) for k, v in sd_g.items()}
    pv = {k: v.clone().requires_grad_(True) for k, v in sd_v.items()}
    ref = O.value_forward(pg, pv, robot, humans, **gkw)
    lo = torch.nn.functional.mse_loss(ref, target)
    lo.backward()
    assert_close_scaled(out, ref, 1e-5, 'V')
    for name, p in ve.graph_model.named_parameters():
        assert p.grad is not None, name
        assert_close_scaled(p.grad, pg[name].grad, 2e-4, 'grad graph ' + name)
    for name, p in ve.value_network.named_parameters():
        assert_close_scaled(p.grad, pv[name].grad, 2e-4, 'grad value ' + name)


def test_training_steps_track_the_oracle(cuda_device):
    """20 Adam steps of the value regression (trainer.py:79-85) on the GPU vs the same steps with the oracle on CPU."""
    ve, sd_g, sd_v = build(3, cuda_device)
    pg = {k: v.clone().requires_grad_(True) for k, v in sd_g.items()}
    pv = {k: v.clone().requires_grad_(True) for k, v in sd_v.items()}
    opt_gpu = torch.optim.Adam(ve.parameters(), lr=1e-3)
    order = [n for n, _ in ve.named_parameters()]
    cpu_params = [pg[n[len('graph_model.'):]] if n.startswith('graph_model.') else pv[n[len('value_network.'):]] for n in order]
    opt_cpu = torch.optim.Adam(cpu_params, lr=1e-3)
    robot, humans = synthetic_states(128, 5, seed=11)
    target = (robot[:, 0, 0:1] * 0.1 + 0.2)
    rd, hd, td = robot.to(cuda_device), humans.to(cuda_device), target.to(cuda_device)
    for step in range(20):
        opt_gpu.zero_grad()
        lg = torch.nn.functional.mse_loss(ve((rd, hd)), td)
        lg.backward()
        opt_gpu.step()
        opt_cpu.zero_grad()
        lc = torch.nn.functional.mse_loss(O.value_forward(pg, pv, robot, humans), target)
        lc.backward()
        opt_cpu.step()
        assert abs(float(lg.detach()) - float(lc.detach())) <= 1e-3 * max(1e-3, abs(float(lc.detach()))), (step, float(lg), float(lc))
    assert float(lg.detach()) < 0.9 * float(torch.nn.functional.mse_loss(O.value_forward(sd_g, sd_v, robot, humans), target))


def test_trainer_style_value_update_with_target_network(cuda_device):
    """optimize_batch semantics (trainer.py:122-131): target = r + gamma_bar * deepcopy(model)(next)."""
    import copy
    ve, _, _ = build(5, cuda_device)
    tgt = copy.deepcopy(ve)
    opt = torch.optim.Adam(ve.parameters(), lr=1e-3)
    robot, humans = synthetic_states(100, 5, seed=1, device=cuda_device)
    nrobot, nhumans = synthetic_states(100, 5, seed=2, device=cuda_device)
    rewards = torch.rand(100, 1, device=cuda_device)
    losses = []
    for _ in range(5):
        opt.zero_grad()
        out = ve((robot, humans))
        target = rewards + pow(0.9, 0.25) * tgt((nrobot, nhumans))
        loss = torch.nn.functional.mse_loss(out, target)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert all(torch.isfinite(p).all() for p in ve.parameters())
    assert losses[-1] < losses[0]


@pytest.mark.parametrize('nh,B,detach', [(5, 64, False), (5, 100, True), (10, 130, False), (20, 33, False)])
def test_state_predictor_native_backward(nh, B, detach, cuda_device):
    """trainer.py:143-149: MSE between predicted and actual next human states; detach=True trains only the motion head."""
    from relationalgraphlearning_b200.state_predictor import StatePredictor
    cfg = policy_config()
    torch.manual_seed(9)
    g = RGL(cfg, 9, 5)
    sp = StatePredictor(cfg, g, 0.25)
    sd_g = {k: v.clone() for k, v in g.state_dict().items()}
    sd_m = {k: v.clone() for k, v in sp.human_motion_predictor.state_dict().items()}
    sp.to(cuda_device)
    robot, humans = synthetic_states(B, nh, seed=4)
    nxt = humans + 0.25 * torch.randn_like(humans)
    before = ops.LAUNCHES
    out = sp((robot.to(cuda_device), humans.to(cuda_device)), None, detach=detach)[1]
    loss = torch.nn.functional.mse_loss(out, nxt.to(cuda_device))
    loss.backward()
    assert ops.LAUNCHES - before >= (3 if detach else 10)
    pg = {k: v.clone().requires_grad_(True) for k, v in sd_g.items()}
    pm = {k: v.clone().requires_grad_(True) for k, v in sd_m.items()}
    H = O.rgl_forward(pg, robot, humans)
    if detach:
        H = H.detach()
    ref = O.mlp(H, pm, '')[:, 1:, :]
    torch.nn.functional.mse_loss(ref, nxt).backward()
    assert_close_scaled(out, ref, 1e-5, 'S')
    for name, p in sp.human_motion_predictor.named_parameters():
        assert_close_scaled(p.grad, pm[name].grad, 2e-4, 'grad motion ' + name)
    for name, p in sp.graph_model.named_parameters():
        if detach:
            assert p.grad is None
        else:
            assert_close_scaled(p.grad, pg[name].grad, 2e-4, 'grad graph ' + name)


def test_fused_td_loss_matches_torch(cuda_device):
    """rgl_td_loss (target, MSE over the global batch and dLoss/dV in one launch) against the tensor expression of
    crowd_nav/utils/trainer.py:125-129."""
    from relationalgraphlearning_b200 import training
    g = torch.Generator().manual_seed(5)
    B = 1000
    V = torch.randn(B, 1, generator=g).to(cuda_device).requires_grad_(True)
    rew = (torch.rand(B, 1, generator=g) * 1.25 - 0.25).to(cuda_device)
    vn = torch.randn(B, 1, generator=g).to(cuda_device)
    gamma = 0.9 ** 0.25
    loss = training.td_loss(V, rew, vn, gamma, 4 * B)               # count = global batch of a 4-rank step
    (3.0 * loss).backward()
    V2 = V.detach().clone().requires_grad_(True)
    ref = ((V2 - (rew + gamma * vn)) ** 2).sum() / float(4 * B)
    (3.0 * ref).backward()
    assert abs(float(loss) - float(ref)) <= 1e-6 * max(1.0, abs(float(ref)))
    assert torch.allclose(V.grad, V2.grad, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize('n,B,up,sim,acc', [(11, 130, 11, True, True), (11, 64, 1, False, False), (6, 77, 6, True, False),
                                            (21, 40, 21, False, True), (32, 9, 32, True, True), (1, 50, 1, True, False)])
def test_staged_attn_sim_backward_matches_the_row_kernels(n, B, up, sim, acc, cuda_device):
    """rgl_attn_sim_bwd (states staged in shared memory, similarity backward fused) against rgl_attn_layer_bwd + rgl_sim_bwd."""
    from relationalgraphlearning_b200 import _lib
    dev = cuda_device
    gen = torch.Generator().manual_seed(n * 1000 + B)
    rnd = lambda *s: torch.randn(*s, generator=gen).to(dev)   # noqa: E731
    A = torch.softmax(rnd(B, n, n), dim=2).contiguous()
    Z, gM, X, Y, mask = rnd(B, n, 32), rnd(B, n, 32), rnd(B, n, 32), rnd(B, n, 32), rnd(B, n, 32)
    gM[:, up:, :] = 0                                         # the contract of up_rows
    gA_in = rnd(B, n, n)
    gX0 = rnd(B, n, 32)
    lib, st, P = _lib.lib(), _lib.stream_ptr(dev), _lib.ptr
    with torch.cuda.device(dev):
        # row kernels
        gZ_ref, gA_ref = torch.empty_like(Z), (gA_in.clone() if acc else torch.empty_like(gA_in))
        _lib.check(lib.rgl_attn_layer_bwd(P(A), P(Z), P(gM), None, 0, P(gZ_ref), P(gA_ref), 1 if acc else 0, B, n, P(mask), n, st), 'attn')
        gY_ref, gX_ref = torch.empty_like(X), gX0.clone()
        if sim:
            _lib.check(lib.rgl_sim_bwd(P(A), P(gA_ref), P(X), P(Y), P(gY_ref), P(gX_ref), B, n, st), 'sim')
        # staged kernel
        gZ, gA_out, gY, gX = torch.empty_like(Z), torch.full_like(gA_in, float('nan')), torch.empty_like(X), gX0.clone()
        _lib.check(lib.rgl_attn_sim_bwd(P(A), P(Z), P(gM), P(mask), up, P(gA_in) if acc else None, P(gZ), None if sim else P(gA_out),
                                        P(X) if sim else None, P(Y) if sim else None, P(gY) if sim else None, P(gX) if sim else None,
                                        1, B, n, st), 'attn_sim')
    torch.cuda.synchronize(dev)
    assert_close_scaled(gZ, gZ_ref, 1e-5, 'gZ')
    if sim:
        assert_close_scaled(gY, gY_ref, 1e-5, 'gY')
        assert_close_scaled(gX, gX_ref, 1e-5, 'gX')
    else:
        assert_close_scaled(gA_out, gA_ref, 1e-5, 'gA')


_TC_BWD_SCRIPT = r"""
import sys, torch
sys.path.insert(0, %r)
from relationalgraphlearning_b200 import training as T
dev = torch.device('cuda:0')
torch.manual_seed(5)
worst = 0.0
# (rows, K, layout, mask, bias, accumulate, grouped): N = 32 always (the tcgen05 kernel's shapes)
for R, K, layout, mask, bias, accum, grouped in [(1000, 32, 1, False, False, True, False), (333, 32, 0, True, True, False, False),
                                                 (4100, 64, 0, True, True, False, True), (129, 64, 1, False, False, True, False),
                                                 (20000, 32, 1, False, False, True, False)]:
    n = 11
    if grouped:      # rows 1..n-1 of every state of a [B, n, .] tensor (the human rows)
        Bq = R // (n - 1); R = Bq * (n - 1)
        Gf, Mf, Xf = torch.randn(Bq, n, 32, device=dev), torch.randn(Bq, n, 32, device=dev), torch.randn(Bq, n, K, device=dev)
        G, M, X = Gf[:, 1:].reshape(R, 32), Mf[:, 1:].reshape(R, 32), Xf[:, 1:].reshape(R, K)
        rG, rM, rX = T._rows(Gf, 32, n - 1, n * 32, offset=32), T._rows(Mf, 32, n - 1, n * 32, offset=32), T._rows(Xf, K, n - 1, n * K, offset=K)
    else:
        G, M, X = torch.randn(R, 32, device=dev), torch.randn(R, 32, device=dev), torch.randn(R, K, device=dev)
        rG, rM, rX = T._rows(G, 32), T._rows(M, 32), T._rows(X, K)
    W = torch.randn(32, K, device=dev) if layout == 0 else torch.randn(K, 32, device=dev)
    Gin0 = torch.randn(R, K, device=dev)
    Gin, dW, db = Gin0.clone(), torch.zeros_like(W), torch.zeros(32, device=dev)
    T._linear_bwd(rG, 32, rX, K, R, W=W, w_layout=layout, mask=rM if mask else None, Gin=T._rows(Gin, K), accumulate=accum,
                  dW=dW, db=db if bias else None, dev=dev)
    torch.cuda.synchronize()
    Gm = (G * (M > 0)) if mask else G
    Gd, Xd, Wd = Gm.double(), X.double(), W.double()
    ref_in = (Gd @ Wd) if layout == 0 else (Gd @ Wd.t())
    if accum: ref_in = ref_in + Gin0.double()
    ref_dW = (Gd.t() @ Xd) if layout == 0 else (Xd.t() @ Gd)
    for got, ref in ((Gin, ref_in), (dW, ref_dW)) + (((db, Gd.sum(0)),) if bias else ()):
        worst = max(worst, float((got.double() - ref).abs().max() / ref.abs().max()))
print('WORST %%.3e' %% worst)
"""


def test_tcgen05_linear_backward_forced_on_every_shape(cuda_device):
    """linear_bwd_tc.cu takes only the big K = 32 launches by default; RGL_BWD_VARIANT=t (read once per process, hence the
    subprocess) runs it on K = 32 / 64, with mask, bias, += and grouped rows, against float64 matmuls."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RGL_BWD_VARIANT='t')
    out = subprocess.run([sys.executable, '-c', _TC_BWD_SCRIPT % root], env=env, capture_output=True, text=True, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    worst = float(out.stdout.strip().split('WORST')[-1])
    assert worst < 2e-6, out.stdout                       # 3xTF32: fp32-level agreement with the float64 reference


@pytest.mark.parametrize('R,K0,grouped', [(700, 5, False), (4000, 9, False), (23000, 5, True), (64, 9, True)])
def test_fused_embedding_mlp_backward_matches_float64(R, K0, grouped, cuda_device):
    """rgl_mlp2_bwd (both Linear layers of w_r / w_h in one launch, hidden gradient kept on chip) against float64 autograd of
    relu(relu(x0 W0^T + b0) W1^T + b1) -- graph_model.py:41-42, helpers.py:5-13 with last_relu=True."""
    import ctypes
    from relationalgraphlearning_b200 import _lib
    dev = cuda_device
    gen = torch.Generator().manual_seed(R + K0)
    rnd = lambda *s: torch.randn(*s, generator=gen)   # noqa: E731
    W0, b0, W1, b1 = rnd(64, K0) * 0.5, rnd(64) * 0.1, rnd(32, 64) * 0.2, rnd(32) * 0.1
    n = 11
    if grouped:      # the rows are nodes 1..n-1 of every state inside [B, n, .] tensors, as the human branch addresses them
        Bq = max(1, R // (n - 1)); R = Bq * (n - 1)
    x0, gX = rnd(R, K0), rnd(R, 32)
    # float64 reference
    p = [t.double().requires_grad_(True) for t in (W0, b0, W1, b1)]
    hid = torch.relu(x0.double() @ p[0].t() + p[1])
    X = torch.relu(hid @ p[2].t() + p[3])
    X.backward(gX.double())
    hid32, X32 = hid.detach().float(), X.detach().float()
    if grouped:
        def embed(t, w):      # place rows into nodes 1..n-1 of a [Bq, n, w] tensor (node 0 = garbage the kernel must not touch)
            full = torch.full((Bq, n, w), 7.0)
            full[:, 1:, :] = t.view(Bq, n - 1, w)
            return full.to(dev)
        gXd, Xd, hd = embed(gX, 32), embed(X32, 32), embed(hid32, 64)
        rG, rM = training._rows(gXd, 32, n - 1, n * 32, offset=32), training._rows(Xd, 32, n - 1, n * 32, offset=32)
        rH = training._rows(hd, 64, n - 1, n * 64, offset=64)
    else:
        gXd, Xd, hd = gX.to(dev), X32.to(dev), hid32.to(dev)
        rG, rM, rH = training._rows(gXd, 32), training._rows(Xd, 32), training._rows(hd, 64)
    x0d = x0.to(dev).contiguous()
    W1d = W1.to(dev)
    dW1, db1, dW0, db0 = (torch.zeros(32, 64, device=dev), torch.zeros(32, device=dev), torch.zeros(64, K0, device=dev),
                          torch.zeros(64, device=dev))
    r0 = training._rows(x0d, K0)
    with torch.cuda.device(dev):
        rc = _lib.lib().rgl_mlp2_bwd(ctypes.byref(rG), ctypes.byref(rM), ctypes.byref(rH), _lib.ptr(W1d), ctypes.byref(r0), K0,
                                     _lib.ptr(dW1), _lib.ptr(db1), _lib.ptr(dW0), _lib.ptr(db0), R, _lib.stream_ptr(dev))
    _lib.check(rc, 'rgl_mlp2_bwd')
    torch.cuda.synchronize(dev)
    for got, ref, what in ((dW1, p[2].grad, 'dW1'), (db1, p[3].grad, 'db1'), (dW0, p[0].grad, 'dW0'), (db0, p[1].grad, 'db0')):
        err = float((got.double().cpu() - ref).abs().max() / ref.abs().max())
        assert err < 2e-5, (what, err)            # 3xTF32 products, fp32 accumulation over R rows
