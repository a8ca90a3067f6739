"""Native training step of the value estimator: fused forward with activation saves + hand-written backward.

`value_forward_train(ve, robot, humans)` is what `ValueEstimator.forward` dispatches to when gradients are
required (crowd_nav/utils/trainer.py:80,123 followed by loss.backward()).  Forward = the same fused sm_100a
kernels as inference, additionally writing the activations the backward needs; backward = the kernel sequence
below (csrc/train_kernels.cu), producing the gradient of every parameter of `graph_model` and `value_network`.

  value head   4 x rgl_linear_bwd                     (gV -> gE, dW/db of the 4 Linear layers)
  per layer    rgl_linear_bwd (W_l, relu mask)  ->  rgl_attn_layer_bwd (A^T gM, gA += gM H^T)
  similarity   rgl_sim_bwd (softmax + Y X^T)    ->  rgl_linear_bwd (w_a)
  embedding    2 x rgl_linear_bwd per agent kind (robot rows / human rows of the [B,n,32] gradient, grouped-row view)
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
    lib = _lib.lib()
    rc = lib.rgl_linear_bwd(ctypes.byref(G), N, ctypes.byref(mask) if mask is not None else None,
                            ctypes.byref(Xin) if Xin is not None else None, K,
                            _lib.ptr(W) if W is not None else None, w_layout,
                            ctypes.byref(Gin) if Gin is not None else None, 1 if accumulate else 0,
                            _lib.ptr(dW) if dW is not None else None, _lib.ptr(db) if db is not None else None,
                            R, _lib.stream_ptr(dev))
    _lib.check(rc, 'rgl_linear_bwd')
    ops._count(1)


class _ValueTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ve, robot, humans, *params):
        g = ve.graph_model
        robot, humans = ops._check_state(robot, humans)
        B, Nh = robot.size(0), humans.size(1)
        n, L, dev = Nh + 1, g.num_layer, robot.device
        # one flat allocation for every saved activation (views below), instead of ~15 small tensors
        sizes = [B * 64, B * Nh * 64, B * n * 32, B * n * 32, B * n * n] + [B * n * 32] * (3 * L) + [B * 32, B, B * 32, B * 128, B * 128]
        flat = torch.empty(sum((x + 3) & ~3 for x in sizes), dtype=torch.float32, device=dev)
        views, off = [], 0
        for x in sizes:
            views.append(flat[off:off + x])
            off += (x + 3) & ~3
        sv = dict(a1r=views[0].view(B, 64), a1h=views[1].view(B, Nh, 64), X=views[2].view(B, n, 32), Y=views[3].view(B, n, 32),
                  A=views[4].view(B, n, n), M=[views[5 + l].view(B, n, 32) for l in range(L)],
                  Rl=[views[5 + L + l].view(B, n, 32) for l in range(L)], Hl=[views[5 + 2 * L + l].view(B, n, 32) for l in range(L)])
        E, V, v0, v1, v2 = (views[5 + 3 * L].view(B, 32), views[6 + 3 * L].view(B, 1), views[7 + 3 * L].view(B, 32),
                            views[8 + 3 * L].view(B, 128), views[9 + 3 * L].view(B, 128))
        cs = _lib.GraphSave()
        for k in ('a1r', 'a1h', 'X', 'Y', 'A'):
            setattr(cs, k, sv[k].data_ptr())
        for l in range(L):
            cs.M[l], cs.Rl[l], cs.Hl[l] = sv['M'][l].data_ptr(), sv['Rl'][l].data_ptr(), sv['Hl'][l].data_ptr()
        lib = _lib.lib()
        with torch.cuda.device(dev):
            rc = lib.rgl_graph_forward_train(_lib.ptr(robot), _lib.ptr(humans), B, Nh, _lib.ptr(ops.packed_graph(g)), L,
                                             g.flags(), ctypes.byref(cs), None, _lib.ptr(E), _lib.stream_ptr(dev))
            _lib.check(rc, 'rgl_graph_forward_train')
            rc = lib.rgl_value_head_train(_lib.ptr(E), B, _lib.ptr(ops.packed_value(ve.value_network, ve._pack_cache)),
                                          _lib.ptr(V), _lib.ptr(v0), _lib.ptr(v1), _lib.ptr(v2), _lib.stream_ptr(dev))
            _lib.check(rc, 'rgl_value_head_train')
        ops._count(2)
        ctx.ve, ctx.sv, ctx.acts = ve, sv, (robot, humans, E, v0, v1, v2)
        ctx.nparams = len(params)
        return V.clone()

    @staticmethod
    def backward(ctx, gV):
        ve, sv = ctx.ve, ctx.sv
        g = ve.graph_model
        robot, humans, E, v0, v1, v2 = ctx.acts
        B, Nh = robot.size(0), humans.size(1)
        n, L, dev = Nh + 1, g.num_layer, robot.device
        skip = bool(g.skip_connection)
        gV = gV.contiguous().float()
        vn = ve.value_network
        plist = list(g.parameters()) + list(vn.parameters())
        # one zero-filled flat buffer: [all parameter gradients | gH_L], one uninitialised flat buffer for the temporaries
        psz = [(p.numel() + 3) & ~3 for p in plist]
        zflat = torch.zeros(sum(psz) + B * n * 32, dtype=torch.float32, device=dev)
        gp, off = {}, 0
        for p, sz in zip(plist, psz):
            gp[id(p)] = zflat[off:off + p.numel()].view(p.shape)
            off += sz
        gH0 = zflat[off:].view(B, n, 32)
        tsz = [B * 128, B * 128, B * 32, B * n * n, B * n * 32, B * n * 32, B * 64, B * Nh * 64] + [B * n * 32] * L
        tflat = torch.empty(sum((x + 3) & ~3 for x in tsz), dtype=torch.float32, device=dev)
        tv, off = [], 0
        for x in tsz:
            tv.append(tflat[off:off + x])
            off += (x + 3) & ~3
        G = lambda p: gp[id(p)]   # noqa: E731
        with torch.cuda.device(dev), torch.no_grad():
            # ---------------- value head: V = L6(relu(L4(relu(L2(relu(L0(E))))))) ----------------
            g2, g1, g0 = tv[0].view(B, 128), tv[1].view(B, 128), tv[2].view(B, 32)
            gH = gH0                                      # gradient w.r.t. H_L: only the robot row is non-zero
            _linear_bwd(_rows(gV, 1), 1, _rows(v2, 128), 100, B, W=vn[6].weight, Gin=_rows(g2, 128), dW=G(vn[6].weight), db=G(vn[6].bias), dev=dev)
            _linear_bwd(_rows(g2, 128), 100, _rows(v1, 128), 100, B, W=vn[4].weight, mask=_rows(v2, 128), Gin=_rows(g1, 128),
                        dW=G(vn[4].weight), db=G(vn[4].bias), dev=dev)
            _linear_bwd(_rows(g1, 128), 100, _rows(v0, 32), 32, B, W=vn[2].weight, mask=_rows(v1, 128), Gin=_rows(g0, 32),
                        dW=G(vn[2].weight), db=G(vn[2].bias), dev=dev)
            _linear_bwd(_rows(g0, 32), 32, _rows(E, 32), 32, B, W=vn[0].weight, mask=_rows(v0, 32),
                        Gin=_rows(gH, 32, 1, n * 32), dW=G(vn[0].weight), db=G(vn[0].bias), dev=dev)
            # ---------------- GCN layers, last to first ----------------
            gA = tv[3].view(B, n, n)
            gM = tv[4].view(B, n, 32)
            lib = _lib.lib()
            for l in range(L - 1, -1, -1):
                Hprev = sv['X'] if l == 0 else sv['Hl'][l - 1]
                _linear_bwd(_rows(gH, 32), 32, _rows(sv['M'][l], 32), 32, B * n, W=g.Ws[l], w_layout=1, mask=_rows(sv['Rl'][l], 32),
                            Gin=_rows(gM, 32), dW=G(g.Ws[l]), dev=dev)
                gHp = tv[8 + l].view(B, n, 32)
                rc = lib.rgl_attn_layer_bwd(_lib.ptr(sv['A']), _lib.ptr(Hprev), _lib.ptr(gM), _lib.ptr(gH), 1 if skip else 0,
                                            _lib.ptr(gHp), _lib.ptr(gA), 0 if l == L - 1 else 1, B, n, _lib.stream_ptr(dev))
                _lib.check(rc, 'rgl_attn_layer_bwd')
                ops._count(1)
                gH = gHp
            gX = gH                                         # gradient w.r.t. X from the layer stack
            # ---------------- similarity ----------------
            gY = tv[5].view(B, n, 32)
            rc = lib.rgl_sim_bwd(_lib.ptr(sv['A']), _lib.ptr(gA), _lib.ptr(sv['X']), _lib.ptr(sv['Y']), _lib.ptr(gY), _lib.ptr(gX),
                                 B, n, _lib.stream_ptr(dev))
            _lib.check(rc, 'rgl_sim_bwd')
            ops._count(1)
            _linear_bwd(_rows(gY, 32), 32, _rows(sv['X'], 32), 32, B * n, W=g.w_a, w_layout=1, Gin=_rows(gX, 32), accumulate=True,
                        dW=G(g.w_a), dev=dev)
            # ---------------- embeddings: robot rows (node 0) and human rows (nodes 1..Nh) of gX ----------------
            ga_r, ga_h = tv[6].view(B, 64), tv[7].view(B * Nh, 64)
            _linear_bwd(_rows(gX, 32, 1, n * 32), 32, _rows(sv['a1r'], 64), 64, B, W=g.w_r[2].weight, mask=_rows(sv['X'], 32, 1, n * 32),
                        Gin=_rows(ga_r, 64), dW=G(g.w_r[2].weight), db=G(g.w_r[2].bias), dev=dev)
            _linear_bwd(_rows(ga_r, 64), 64, _rows(robot, 9), 9, B, mask=_rows(sv['a1r'], 64), dW=G(g.w_r[0].weight), db=G(g.w_r[0].bias), dev=dev)
            _linear_bwd(_rows(gX, 32, Nh, n * 32, offset=32), 32, _rows(sv['a1h'], 64), 64, B * Nh, W=g.w_h[2].weight,
                        mask=_rows(sv['X'], 32, Nh, n * 32, offset=32), Gin=_rows(ga_h, 64), dW=G(g.w_h[2].weight), db=G(g.w_h[2].bias), dev=dev)
            _linear_bwd(_rows(ga_h, 64), 64, _rows(humans, 5), 5, B * Nh, mask=_rows(sv['a1h'], 64), dW=G(g.w_h[0].weight),
                        db=G(g.w_h[0].bias), dev=dev)
        grads = [gp[id(p)] for p in ve._train_params()]
        ctx.sv = ctx.acts = None
        return (None, None, None) + tuple(grads)


def native_supported(ve):
    g = ve.graph_model
    return ve.kernel_supported() and not g.layerwise_graph


def value_forward_train(ve, robot, humans):
    return _ValueTrain.apply(ve, robot, humans, *ve._train_params())
