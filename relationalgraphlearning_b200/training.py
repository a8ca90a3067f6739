"""Native training steps: fused forward with activation saves + hand-written backward.

`value_forward_train(ve, robot, humans)` / `statepred_forward_train(sp, robot, humans, detach)` are what
`ValueEstimator.forward` / `StatePredictor.forward` dispatch to when gradients are required
(crowd_nav/utils/trainer.py:80,95,123,144 followed by loss.backward()).  Forward = the same fused sm_100a kernels
as inference, additionally writing the activations the backward needs; backward = the kernel sequence below
(csrc/train_kernels.cu), producing the gradient of every parameter.

  value head    4 x rgl_linear_bwd                    (gV -> gE, dW/db of the 4 Linear layers)
  motion head   2 x rgl_linear_bwd                    (gS -> gH_L on the human rows)
  per layer     rgl_linear_bwd (W_l, relu mask)  ->  rgl_attn_layer_bwd (A^T gM, gA += gM H^T)          [fp32-FMA forward]
                rgl_attn_layer_bwd (relu mask, Z_l)  ->  rgl_linear_bwd (W_l)                      [tcgen05 forward: relu(A (H W))]
  similarity    rgl_sim_bwd (softmax + Y X^T)    ->  rgl_linear_bwd (w_a)
  embedding     2 x rgl_linear_bwd per agent kind (robot rows / human rows of the [B,n,32] gradient, grouped-row view)
"""
import ctypes

import torch

from . import _lib, ops


def _rows(t, ld=None, rows_per_group=0, group_stride=0, offset=0):
    r = _lib.Rows()
    r.ptr = t.data_ptr() + 4 * offset
    r.ld = int(ld if ld is not None else t.size(-1))
    r.rows_per_group = int(rows_per_group)
    r.group_stride = int(group_stride)
    return r


def _linear_bwd(G, N, Xin, K, R, W=None, w_layout=0, mask=None, Gin=None, accumulate=False, dW=None, db=None, dev=None):
    rc = _lib.lib().rgl_linear_bwd(ctypes.byref(G), N, ctypes.byref(mask) if mask is not None else None,
                                   ctypes.byref(Xin) if Xin is not None else None, K,
                                   _lib.ptr(W) if W is not None else None, w_layout,
                                   ctypes.byref(Gin) if Gin is not None else None, 1 if accumulate else 0,
                                   _lib.ptr(dW) if dW is not None else None, _lib.ptr(db) if db is not None else None,
                                   R, _lib.stream_ptr(dev))
    _lib.check(rc, 'rgl_linear_bwd')
    ops._count(1)


def _mlp2_bwd(G, mask, hidden, X0, K0, R, mlp, Gf, dev):
    """Both Linear layers of an embedding MLP (mlp[0]: K0 -> 64, mlp[2]: 64 -> 32) in one launch (rgl_mlp2_bwd)."""
    rc = _lib.lib().rgl_mlp2_bwd(ctypes.byref(G), ctypes.byref(mask), ctypes.byref(hidden), _lib.ptr(mlp[2].weight), ctypes.byref(X0), K0,
                                 _lib.ptr(Gf(mlp[2].weight)), _lib.ptr(Gf(mlp[2].bias)), _lib.ptr(Gf(mlp[0].weight)), _lib.ptr(Gf(mlp[0].bias)),
                                 R, _lib.stream_ptr(dev))
    _lib.check(rc, 'rgl_mlp2_bwd')
    ops._count(1)


def _carve(sizes, dev, zero=False):
    """One flat fp32 allocation carved into 16-byte aligned 1-D views (1 alloc / 1 fill instead of one per tensor)."""
    flat = (torch.zeros if zero else torch.empty)(sum((x + 3) & ~3 for x in sizes), dtype=torch.float32, device=dev)
    views, off = [], 0
    for x in sizes:
        views.append(flat[off:off + x])
        off += (x + 3) & ~3
    return views


def _train_tc(n):
    """The training forward runs on the tcgen05 kernel (graph_forward_tp.cu, RGL_FLAG_TRAIN_TC) for the compiled node counts;
    RGL_TRAIN_VARIANT=f keeps the fp32-FMA kernel (experiments only)."""
    import os
    return n in (6, 11, 21) and os.environ.get('RGL_TRAIN_VARIANT', 't')[0] != 'f'


def _graph_forward_train(g, robot, humans, extra_sizes, motion_blob=None, want_E=False, want_S=False):
    """Fused graph forward with saves.  Returns (saves dict, extra views, E or None, S or None).

    Two save layouts (sv['tc']):
      fp32-FMA kernel   a1r [B,64], a1h [B,Nh,64], M_l = A H_{l-1}, mh [B,Nh,64]         layer = relu((A H) W)
      tcgen05 kernel    a1 [B,n,64] (robot row 0, humans rows 1..Nh), M_l = Z_l = H_{l-1} W_l, mh [B,n,64]   layer = relu(A (H W))
    sv['a1r'] / sv['a1h'] / sv['mh'] are RglRows views, so the backward addresses either layout the same way."""
    B, Nh = robot.size(0), humans.size(1)
    n, L, dev = Nh + 1, g.num_layer, robot.device
    tc = _train_tc(n)
    hid = [B * n * 64, 0] if tc else [B * 64, B * Nh * 64]
    sizes = hid + [B * n * 32, B * n * 32, B * n * n] + [B * n * 32] * (3 * L) + \
        [B * 32 if want_E else 0, B * Nh * 5 if want_S else 0, (B * n * 64 if tc else B * Nh * 64) if want_S else 0] + list(extra_sizes)
    v = _carve(sizes, dev)
    sv = dict(tc=tc, X=v[2].view(B, n, 32), Y=v[3].view(B, n, 32), A=v[4].view(B, n, n),
              M=[v[5 + l].view(B, n, 32) for l in range(L)], Rl=[v[5 + L + l].view(B, n, 32) for l in range(L)],
              Hl=[v[5 + 2 * L + l].view(B, n, 32) for l in range(L)])
    E = v[5 + 3 * L].view(B, 32) if want_E else None
    S = v[6 + 3 * L].view(B, Nh, 5) if want_S else None
    mh = v[7 + 3 * L] if want_S else None
    sv['_keep'] = (v[0], v[1], mh)
    if tc:
        sv['a1r'] = _rows(v[0], 64, 1, n * 64)                         # row 0 of every state
        sv['a1h'] = _rows(v[0], 64, Nh, n * 64, offset=64)             # rows 1..Nh
        sv['mh'] = _rows(mh, 64, Nh, n * 64, offset=64) if want_S else None
        a1r_ptr, a1h_ptr = v[0].data_ptr(), v[0].data_ptr() + 64 * 4
    else:
        sv['a1r'] = _rows(v[0], 64)
        sv['a1h'] = _rows(v[1], 64)
        sv['mh'] = _rows(mh, 64) if want_S else None
        a1r_ptr, a1h_ptr = v[0].data_ptr(), v[1].data_ptr()
    cs = _lib.GraphSave()
    cs.a1r, cs.a1h = a1r_ptr, a1h_ptr
    for k in ('X', 'Y', 'A'):
        setattr(cs, k, sv[k].data_ptr())
    for l in range(L):
        cs.M[l], cs.Rl[l], cs.Hl[l] = sv['M'][l].data_ptr(), sv['Rl'][l].data_ptr(), sv['Hl'][l].data_ptr()
    cs.mh = mh.data_ptr() if want_S else None
    flags = (g.flags() & ~_lib.FLAG_FP32_FMA) | (_lib.FLAG_TRAIN_TC if tc else 0)
    with torch.cuda.device(dev):
        rc = _lib.lib().rgl_graph_forward_train(_lib.ptr(robot), _lib.ptr(humans), B, Nh, _lib.ptr(ops.packed_graph(g, force=ops.capturing())), L, flags,
                                                _lib.ptr(motion_blob) if want_S else None, ctypes.byref(cs), None,
                                                _lib.ptr(E), _lib.ptr(S), _lib.stream_ptr(dev))
    _lib.check(rc, 'rgl_graph_forward_train')
    ops._count(1)
    return sv, v[8 + 3 * L:], E, S


def _attn_bwd(A, Hprev, gM, gH, skip, gHprev, gA, accumulate_gA, B, n, dev, mask=None, up_rows=None):
    rc = _lib.lib().rgl_attn_layer_bwd(_lib.ptr(A), _lib.ptr(Hprev), _lib.ptr(gM), _lib.ptr(gH) if gH is not None else None, 1 if skip else 0,
                                       _lib.ptr(gHprev), _lib.ptr(gA), 1 if accumulate_gA else 0, B, n,
                                       _lib.ptr(mask) if mask is not None else None, n if up_rows is None else up_rows, _lib.stream_ptr(dev))
    _lib.check(rc, 'rgl_attn_layer_bwd')
    ops._count(1)


def _graph_backward(g, sv, robot, humans, gH, G, dev, top_rows=None):
    """Back-propagate gH = dLoss/dH_L [B,n,32] through the GCN layers, the similarity and the embeddings.
    G(p) returns the (zero-initialised) gradient buffer of parameter p; gradients are accumulated into them.
    top_rows: only the first top_rows node rows of gH are non-zero (1 for the value head, which reads the robot row)."""
    B, Nh = robot.size(0), humans.size(1)
    n, L = Nh + 1, g.num_layer
    skip = bool(g.skip_connection)
    lib = _lib.lib()
    t = _carve([B * n * n, B * n * 32, B * n * 32, 0, 0] + [B * n * 32] * L, dev)
    gA, gM, gY = t[0].view(B, n, n), t[1].view(B, n, 32), t[2].view(B, n, 32)
    if sv['tc']:
        # H_l = relu(A Z) (+ H_{l-1}), Z = H_{l-1} W_l:   gZ = A^T (gH . mask),  gA += (gH . mask) Z^T,
        #                                                 dW_l = H_{l-1}^T gZ,    gH_{l-1} = gZ W_l^T (+ gH)
        # One staged kernel per layer (rgl_attn_sim_bwd); at layer 0 it also runs the similarity backward, so gA never
        # reaches memory there: it writes gY and the similarity term of gX, and the linear backward of W_0 accumulates
        # onto it.  Skip connection: gH has been consumed by the staged kernel, gZ W^T is accumulated into it in place.
        for l in range(L - 1, -1, -1):
            Hprev = sv['X'] if l == 0 else sv['Hl'][l - 1]
            sim = l == 0
            gHp = gH if skip else t[5 + l].view(B, n, 32)
            rc = lib.rgl_attn_sim_bwd(_lib.ptr(sv['A']), _lib.ptr(sv['M'][l]), _lib.ptr(gH), _lib.ptr(sv['Rl'][l]),
                                      top_rows if (l == L - 1 and top_rows) else n, _lib.ptr(gA) if l != L - 1 else None,
                                      _lib.ptr(gM), None if sim else _lib.ptr(gA),
                                      _lib.ptr(sv['X']) if sim else None, _lib.ptr(sv['Y']) if sim else None,
                                      _lib.ptr(gY) if sim else None, _lib.ptr(gHp) if sim else None, 1 if skip else 0,
                                      B, n, _lib.stream_ptr(dev))
            _lib.check(rc, 'rgl_attn_sim_bwd')
            ops._count(1)
            _linear_bwd(_rows(gM, 32), 32, _rows(Hprev, 32), 32, B * n, W=g.Ws[l], w_layout=1, Gin=_rows(gHp, 32),
                        accumulate=skip or sim, dW=G(g.Ws[l]), dev=dev)
            gH = gHp
        gX = gH
    else:
        for l in range(L - 1, -1, -1):
            Hprev = sv['X'] if l == 0 else sv['Hl'][l - 1]
            gHp = t[5 + l].view(B, n, 32)
            _linear_bwd(_rows(gH, 32), 32, _rows(sv['M'][l], 32), 32, B * n, W=g.Ws[l], w_layout=1, mask=_rows(sv['Rl'][l], 32),
                        Gin=_rows(gM, 32), dW=G(g.Ws[l]), dev=dev)
            _attn_bwd(sv['A'], Hprev, gM, gH, skip, gHp, gA, l != L - 1, B, n, dev)
            gH = gHp
        gX = gH                                         # gradient w.r.t. X from the layer stack
        rc = lib.rgl_sim_bwd(_lib.ptr(sv['A']), _lib.ptr(gA), _lib.ptr(sv['X']), _lib.ptr(sv['Y']), _lib.ptr(gY), _lib.ptr(gX),
                             B, n, _lib.stream_ptr(dev))
        _lib.check(rc, 'rgl_sim_bwd')
        ops._count(1)
    _linear_bwd(_rows(gY, 32), 32, _rows(sv['X'], 32), 32, B * n, W=g.w_a, w_layout=1, Gin=_rows(gX, 32), accumulate=True,
                dW=G(g.w_a), dev=dev)
    # embeddings: robot rows (node 0) and human rows (nodes 1..Nh) of gX, addressed in place as grouped rows.
    # The robot branch (B rows: less than one wave of CTAs) runs on a side stream next to the human branch (B*Nh rows);
    # inside a captured step it becomes a parallel branch of the graph.
    cur, side = torch.cuda.current_stream(dev), _bwd_side_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        _mlp2_bwd(_rows(gX, 32, 1, n * 32), _rows(sv['X'], 32, 1, n * 32), sv['a1r'], _rows(robot, 9), 9, B, g.w_r, G, dev)
    _mlp2_bwd(_rows(gX, 32, Nh, n * 32, offset=32), _rows(sv['X'], 32, Nh, n * 32, offset=32), sv['a1h'], _rows(humans, 5), 5, B * Nh,
              g.w_h, G, dev)
    cur.wait_stream(side)           # every buffer the side branch touched outlives this join


_BWD_SIDE = {}


def _bwd_side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _BWD_SIDE:
        _BWD_SIDE[key] = torch.cuda.Stream(dev)
    return _BWD_SIDE[key]


def _grad_buffers(plist, extra, dev, sink=None):
    """Gradient buffers of `plist` + `extra` zeroed floats.  Returns (dict id(p)->view, extra view, direct).

    With a flat-gradient sink attached to the module (parallel.FlatGrads) that covers every parameter, the views are
    slices of the sink's accumulation buffer: the backward kernels accumulate straight into the buffer the gradient
    all-reduce / optimizer reads (direct = True; autograd then receives no parameter gradients).  Otherwise a fresh
    zero-filled flat buffer is carved and returned through autograd."""
    if sink is not None and sink.covers(plist) and sink.dev == dev:
        return {id(p): sink.accum_view(p) for p in plist}, torch.zeros(extra, dtype=torch.float32, device=dev), True
    v = _carve([p.numel() for p in plist] + [extra], dev, zero=True)
    return {id(p): v[i].view(p.shape) for i, p in enumerate(plist)}, v[len(plist)], False


def _saved(ctx, what):
    if ctx.sv is None:
        raise RuntimeError('%s: the saved activations were freed by the first backward pass; run the forward again '
                           '(retain_graph=True is not supported by the native training path)' % what)


class _ValueTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ve, robot, humans, *params):
        g = ve.graph_model
        robot, humans = ops._check_state(robot, humans)
        B, dev = robot.size(0), robot.device
        # the value-network weights are (re)packed on a side stream while the graph forward runs
        cur, side = torch.cuda.current_stream(dev), _bwd_side_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            vblob = ops.packed_value(ve.value_network, ve._pack_cache, force=ops.capturing())
        sv, ex, E, _ = _graph_forward_train(g, robot, humans, [B, B * 32, B * 128, B * 128], want_E=True)
        cur.wait_stream(side)
        V, v0, v1, v2 = ex[0].view(B, 1), ex[1].view(B, 32), ex[2].view(B, 128), ex[3].view(B, 128)
        with torch.cuda.device(dev):
            rc = _lib.lib().rgl_value_head_train(_lib.ptr(E), B, _lib.ptr(vblob),
                                                 _lib.ptr(V), _lib.ptr(v0), _lib.ptr(v1), _lib.ptr(v2), _lib.stream_ptr(dev))
        _lib.check(rc, 'rgl_value_head_train')
        ops._count(1)
        ctx.ve, ctx.sv, ctx.acts = ve, sv, (robot, humans, E, v0, v1, v2)
        return V.clone()

    @staticmethod
    def backward(ctx, gV):
        _saved(ctx, 'ValueEstimator backward')
        ve, sv = ctx.ve, ctx.sv
        g, vn = ve.graph_model, ve.value_network
        robot, humans, E, v0, v1, v2 = ctx.acts
        B, n, dev = robot.size(0), humans.size(1) + 1, robot.device
        if not any(ctx.needs_input_grad[3:]):            # every parameter frozen (e.g. a target network called under grad)
            ctx.sv = ctx.acts = None
            return (None,) * (3 + len(ve._train_params()))
        gV = gV.contiguous().float()
        gp, gHflat, direct = _grad_buffers(list(g.parameters()) + list(vn.parameters()), B * n * 32, dev, getattr(ve, '_grad_sink', None))
        G = lambda p: gp[id(p)]   # noqa: E731
        with torch.cuda.device(dev), torch.no_grad():
            t = _carve([B * 128, B * 128, B * 32], dev)
            g2, g1, g0 = t[0].view(B, 128), t[1].view(B, 128), t[2].view(B, 32)
            gH = gHflat.view(B, n, 32)                    # gradient w.r.t. H_L: only the robot row is non-zero
            _linear_bwd(_rows(gV, 1), 1, _rows(v2, 128), 100, B, W=vn[6].weight, Gin=_rows(g2, 128), dW=G(vn[6].weight), db=G(vn[6].bias), dev=dev)
            _linear_bwd(_rows(g2, 128), 100, _rows(v1, 128), 100, B, W=vn[4].weight, mask=_rows(v2, 128), Gin=_rows(g1, 128),
                        dW=G(vn[4].weight), db=G(vn[4].bias), dev=dev)
            _linear_bwd(_rows(g1, 128), 100, _rows(v0, 32), 32, B, W=vn[2].weight, mask=_rows(v1, 128), Gin=_rows(g0, 32),
                        dW=G(vn[2].weight), db=G(vn[2].bias), dev=dev)
            _linear_bwd(_rows(g0, 32), 32, _rows(E, 32), 32, B, W=vn[0].weight, mask=_rows(v0, 32),
                        Gin=_rows(gH, 32, 1, n * 32), dW=G(vn[0].weight), db=G(vn[0].bias), dev=dev)
            _graph_backward(g, sv, robot, humans, gH, G, dev, top_rows=1)
        grads = [None if direct else gp[id(p)] for p in ve._train_params()]
        ctx.sv = ctx.acts = None
        return (None, None, None) + tuple(grads)


class _StatePredTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sp, detach, robot, humans, *params):
        g = sp.graph_model
        robot, humans = ops._check_state(robot, humans)
        mblob = ops.packed_motion(sp.human_motion_predictor, sp._pack_cache, force=ops.capturing())
        sv, _, _, S = _graph_forward_train(g, robot, humans, [], motion_blob=mblob, want_S=True)
        ctx.sp, ctx.sv, ctx.acts, ctx.detach = sp, sv, (robot, humans), detach
        return S.clone()

    @staticmethod
    def backward(ctx, gS):
        _saved(ctx, 'StatePredictor backward')
        sp, sv, detach = ctx.sp, ctx.sv, ctx.detach
        g, mp = sp.graph_model, sp.human_motion_predictor
        robot, humans = ctx.acts
        B, Nh, dev = robot.size(0), humans.size(1), robot.device
        n, L = Nh + 1, g.num_layer
        if not any(ctx.needs_input_grad[4:]):
            ctx.sv = ctx.acts = None
            return (None,) * (4 + len(sp._train_params(detach)))
        gS = gS.contiguous().float()
        plist = list(mp.parameters()) + ([] if detach else list(g.parameters()))
        gp, gHflat, direct = _grad_buffers(plist, B * n * 32, dev, getattr(sp, '_grad_sink', None))
        G = lambda p: gp[id(p)]   # noqa: E731
        with torch.cuda.device(dev), torch.no_grad():
            gmh = _carve([B * Nh * 64], dev)[0].view(B * Nh, 64)
            gH = gHflat.view(B, n, 32)                    # gradient w.r.t. H_L: human rows only (the head drops node 0)
            HL = sv['Hl'][L - 1]
            _linear_bwd(_rows(gS, 5), 5, sv['mh'], 64, B * Nh, W=mp[2].weight, Gin=_rows(gmh, 64), dW=G(mp[2].weight),
                        db=G(mp[2].bias), dev=dev)
            _linear_bwd(_rows(gmh, 64), 64, _rows(HL, 32, Nh, n * 32, offset=32), 32, B * Nh, W=None if detach else mp[0].weight,
                        mask=sv['mh'], Gin=None if detach else _rows(gH, 32, Nh, n * 32, offset=32),
                        dW=G(mp[0].weight), db=G(mp[0].bias), dev=dev)
            if not detach:
                _graph_backward(g, sv, robot, humans, gH, G, dev)
        grads = [None if direct else gp[id(p)] for p in sp._train_params(detach)]
        ctx.sv = ctx.acts = None
        return (None, None, None, None) + tuple(grads)


def native_supported(module):
    g = module.graph_model
    return module.kernel_supported() and not g.layerwise_graph


def value_forward_train(ve, robot, humans):
    return _ValueTrain.apply(ve, robot, humans, *ve._train_params())


def statepred_forward_train(sp, robot, humans, detach):
    return _StatePredTrain.apply(sp, detach, robot, humans, *sp._train_params(detach))


class _TDLoss(torch.autograd.Function):
    """loss = sum_b (V - (reward + gamma_bar * V_next))^2 / count in ONE launch (trainer.py:125-129; MSELoss(mean) when
    count = the global batch); the same launch stores dLoss/dV, so backward is a scale by the incoming gradient."""

    @staticmethod
    def forward(ctx, V, reward, V_next, gamma_bar, count):
        V, reward, V_next = ops._f32c(V), ops._f32c(reward), ops._f32c(V_next)
        B, dev = V.numel(), V.device
        buf = torch.zeros(1 + B, dtype=torch.float32, device=dev)          # [loss | gV]
        with torch.cuda.device(dev):
            rc = _lib.lib().rgl_td_loss(_lib.ptr(V), _lib.ptr(reward), _lib.ptr(V_next), B, float(gamma_bar), 1.0 / float(count),
                                        _lib.ptr(buf), ctypes.c_void_p(buf.data_ptr() + 4), _lib.stream_ptr(dev))
        _lib.check(rc, 'rgl_td_loss')
        ops._count(1)
        ctx.gV = buf[1:].view_as(V)
        return buf[0]

    @staticmethod
    def backward(ctx, gloss):
        return ctx.gV * gloss, None, None, None, None


def td_loss(V, reward, V_next, gamma_bar, count):
    """Fused TD loss of the value step (differentiable in V only; reward / V_next are treated as constants, as in the
    data-parallel step where the target network runs under no_grad)."""
    return _TDLoss.apply(V, reward, V_next, gamma_bar, count)
