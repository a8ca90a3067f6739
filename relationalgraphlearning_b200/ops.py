"""Tensor-level wrappers over the C ABI (include/rgl_b200.h): packing caches, forward ops, autograd glue.

Every function here launches hand-written sm_100a kernels through ctypes on the current CUDA stream of
the inputs' device.  There is no alternative compute path for CUDA tensors: if librgl_b200.so is missing
or a call fails an exception propagates.
"""
import ctypes
import weakref

import torch

from . import _lib
from . import _torch_math as TM

LAUNCHES = 0        # kernels launched through this module (bench.py reports it as gpu_launches)


def _count(n):
    global LAUNCHES
    LAUNCHES += n


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# ---------------------------------------------------------------- weight packing ----------------
def capturing():
    """True while the current CUDA stream is being captured into a CUDA graph.  The TRAINING forwards then re-pack their
    weights unconditionally: a captured optimizer step changes the parameters on every replay, but the host-side
    (data_ptr, _version) check only runs at capture time -- the pack kernel has to be part of the graph."""
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


class _PackCache(object):
    """Packed k-major blob of one module, rebuilt when any parameter changed.

    Change detection = (data_ptr, _version) of every parameter: optimizers, load_state_dict and
    .to(device) all bump one of the two.  Writes through `.data` do not; call mark_dirty() then.
    """

    def __init__(self):
        self.key = None
        self.blob = None
        self.vkey = None          # fast path: (_version of every parameter) + data_ptr of the first one

    def mark_dirty(self):
        self.key = None
        self.vkey = None

    def fresh(self, params):
        """Cheap check used on per-batch hot paths (hostio.HostStream): True if no parameter's version counter moved since
        the blob was packed and the first parameter still lives at the same address (optimizer steps and load_state_dict bump
        the counters; .to(device) moves the storage)."""
        vk = self.vkey
        if vk is None or self.blob is None or len(vk) != len(params) + 1 or vk[0] != params[0].data_ptr():
            return False
        for i, p in enumerate(params):
            if p._version != vk[i + 1]:
                return False
        return True

    def get(self, params, nfloats, pack_fn, force=False):
        key = tuple((p.data_ptr(), p._version) for p in params)
        if force or key != self.key or self.blob is None or self.blob.device != params[0].device:
            dev = params[0].device
            if self.blob is None or self.blob.device != dev or self.blob.numel() != nfloats:
                self.blob = torch.empty(nfloats, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                pack_fn(self.blob)
            _count(1)
            self.key = key
        self.vkey = (params[0].data_ptr(),) + tuple(p._version for p in params)
        return self.blob


def _graph_param_list(rgl):
    return [rgl.w_r[0].weight, rgl.w_r[0].bias, rgl.w_r[2].weight, rgl.w_r[2].bias,
            rgl.w_h[0].weight, rgl.w_h[0].bias, rgl.w_h[2].weight, rgl.w_h[2].bias, rgl.w_a] + list(rgl.Ws)


def packed_graph(rgl, force=False):
    params = _graph_param_list(rgl)          # (the cache key reads data_ptr / _version only: no detach() on the fast path)
    L = len(rgl.Ws)
    lib = _lib.lib()

    def pack(blob):
        gp = _lib.GraphParams()
        names = ['wr0_w', 'wr0_b', 'wr1_w', 'wr1_b', 'wh0_w', 'wh0_b', 'wh1_w', 'wh1_b', 'w_a']
        keep = [_f32c(p.detach()) for p in params]
        for nme, t in zip(names, keep[:9]):
            setattr(gp, nme, t.data_ptr())
        for i in range(L):
            gp.Ws[i] = keep[9 + i].data_ptr()
        gp.num_layer = L
        _lib.check(lib.rgl_pack_graph(ctypes.byref(gp), _lib.ptr(blob), _lib.stream_ptr(blob.device)), 'rgl_pack_graph')

    return rgl._pack_cache.get(params, int(lib.rgl_packed_graph_floats(L)), pack, force)


def packed_value(seq, cache, force=False):
    params = [seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias, seq[4].weight, seq[4].bias,
              seq[6].weight, seq[6].bias]
    lib = _lib.lib()

    def pack(blob):
        vp = _lib.ValueParams()
        keep = [_f32c(p.detach()) for p in params]
        for nme, t in zip(['w0', 'b0', 'w1', 'b1', 'w2', 'b2', 'w3', 'b3'], keep):
            setattr(vp, nme, t.data_ptr())
        _lib.check(lib.rgl_pack_value(ctypes.byref(vp), _lib.ptr(blob), _lib.stream_ptr(blob.device)), 'rgl_pack_value')

    return cache.get(params, int(lib.rgl_packed_value_floats()), pack, force)


def packed_motion(seq, cache, force=False):
    params = [seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias]
    lib = _lib.lib()

    def pack(blob):
        mp = _lib.MotionParams()
        keep = [_f32c(p.detach()) for p in params]
        for nme, t in zip(['w0', 'b0', 'w1', 'b1'], keep):
            setattr(mp, nme, t.data_ptr())
        _lib.check(lib.rgl_pack_motion(ctypes.byref(mp), _lib.ptr(blob), _lib.stream_ptr(blob.device)), 'rgl_pack_motion')

    return cache.get(params, int(lib.rgl_packed_motion_floats()), pack, force)


def require_cuda_or_cpu_module(module, robot, humans, what):
    """Device policy of the drop-in modules.  CUDA tensors ALWAYS run the sm_100a kernels (or raise: a missing library is
    never papered over).  CPU tensors are accepted only by a module whose parameters live on the CPU -- a policy the
    reference's caller never moved to a GPU -- and are then evaluated with the torch-op statement of the reference math
    (_torch_math.py) on the CPU, as SURVEY.md 8(b) asks; RGL_FORBID_CPU=1 turns that into an error."""
    if robot.is_cuda and humans.is_cuda:
        _lib.lib()                      # raises if librgl_b200.so is missing
        return
    import os
    p = next(module.parameters(), None)
    if robot.is_cuda != humans.is_cuda or (p is not None and p.is_cuda):
        raise _lib.RglError('%s: module and inputs must be on the same device (module on %s, robot on %s, humans on %s)' %
                            (what, p.device if p is not None else '?', robot.device, humans.device))
    if os.environ.get('RGL_FORBID_CPU'):
        raise _lib.RglError('%s: CPU tensors refused (RGL_FORBID_CPU is set)' % what)


# ---------------------------------------------------------------- raw forward ops ----------------
def _check_state(robot, humans):
    if not (robot.is_cuda and humans.is_cuda):
        raise _lib.RglError('relationalgraphlearning_b200 computes on CUDA devices only: move the module and its '
                            'inputs to a CUDA device (there is no CPU path)')
    assert robot.dim() == 3 and humans.dim() == 3 and robot.size(1) == 1 and robot.size(2) == 9 and humans.size(2) == 5
    return _f32c(robot), _f32c(humans)


def graph_forward_raw(gblob, num_layer, flags, robot, humans, humans_bcast=1, mblob=None,
                      want_H=False, want_E=False, want_S=False, want_A0=False, out_H=None):
    """One launch of the fused graph kernel.  Returns dict with the requested outputs.  `out_H` (optional): a preallocated
    [B,n,32] fp32 buffer for H -- device memory, or pinned host memory (the kernel then streams H over PCIe itself:
    zero-copy, unified virtual addressing)."""
    robot, humans = _check_state(robot, humans)
    B, Nh = robot.size(0), humans.size(1)
    assert humans.size(0) * humans_bcast >= B
    n = Nh + 1
    dev = robot.device
    out = {}
    if want_H:
        if out_H is not None:
            assert out_H.dtype == torch.float32 and out_H.is_contiguous() and tuple(out_H.shape) == (B, n, 32)
            assert out_H.is_cuda or out_H.is_pinned(), 'out_H must be device or pinned host memory'
        out['H'] = out_H if out_H is not None else torch.empty(B, n, 32, dtype=torch.float32, device=dev)
    if want_E:
        out['E'] = torch.empty(B, 32, dtype=torch.float32, device=dev)
    if want_S:
        out['S'] = torch.empty(B, Nh, 5, dtype=torch.float32, device=dev)
    if want_A0:
        out['A0'] = torch.empty(n, n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().rgl_graph_forward(_lib.ptr(robot), _lib.ptr(humans), B, Nh, humans_bcast, _lib.ptr(gblob),
                                          num_layer, flags, _lib.ptr(mblob) if want_S else None,
                                          _lib.ptr(out.get('H')), _lib.ptr(out.get('E')), _lib.ptr(out.get('S')),
                                          _lib.ptr(out.get('A0')), _lib.stream_ptr(dev))
    _lib.check(rc, 'rgl_graph_forward')
    _count(1 if B > 0 else 0)
    return out


def value_head_raw(vblob, E):
    E = _f32c(E)
    B = E.size(0)
    V = torch.empty(B, 1, dtype=torch.float32, device=E.device)
    with torch.cuda.device(E.device):
        rc = _lib.lib().rgl_value_head(_lib.ptr(E), B, _lib.ptr(vblob), _lib.ptr(V), _lib.stream_ptr(E.device))
    _lib.check(rc, 'rgl_value_head')
    _count(1 if B > 0 else 0)
    return V


def gcn_layer(X, W, A=None, w_a=None, skip=False, return_A=False):
    """Hout = relu((A X) W) (+X); A given [B,n,n] or computed as softmax(X w_a X^T)."""
    X = _f32c(X)
    B, n, _ = X.shape
    Hout = torch.empty_like(X)
    Aout = torch.empty(B, n, n, dtype=torch.float32, device=X.device) if return_A else None
    with torch.cuda.device(X.device):
        rc = _lib.lib().rgl_gcn_layer(_lib.ptr(X), _lib.ptr(_f32c(A)) if A is not None else None, _lib.ptr(_f32c(W)),
                                      _lib.ptr(_f32c(w_a)) if w_a is not None else None, B, n,
                                      _lib.FLAG_SKIP if skip else 0, _lib.ptr(Hout), _lib.ptr(Aout), _lib.stream_ptr(X.device))
    _lib.check(rc, 'rgl_gcn_layer')
    _count(1 if B > 0 else 0)
    return (Hout, Aout) if return_A else Hout


def plan_expand(robot, humans, actions, time_step, humans_bcast=1, want_next=True, want_reward=True, kinematics='holonomic'):
    """robot[E,1,9], humans[E/humans_bcast,Nh,5], actions double[A,2] ((vx,vy) holonomic | (v,r) unicycle)
    -> next_robot[E*A,1,9], reward[E*A]."""
    robot, humans = _check_state(robot, humans)
    E, Nh, A = robot.size(0), humans.size(1), actions.size(0)
    assert actions.dtype == torch.float64 and actions.is_cuda and actions.is_contiguous()
    assert humans.size(0) * humans_bcast >= E, 'humans must cover every state (humans.size(0) * humans_bcast >= E)'
    nxt = torch.empty(E * A, 1, 9, dtype=torch.float32, device=robot.device) if want_next else None
    rew = torch.empty(E * A, dtype=torch.float32, device=robot.device) if want_reward else None
    kin = _lib.KIN_HOLONOMIC if kinematics == 'holonomic' else _lib.KIN_UNICYCLE
    with torch.cuda.device(robot.device):
        rc = _lib.lib().rgl_plan_expand(_lib.ptr(robot), _lib.ptr(humans), E, Nh, humans_bcast, _lib.ptr(actions), A, float(time_step),
                                        kin, _lib.ptr(nxt), _lib.ptr(rew), _lib.stream_ptr(robot.device))
    _lib.check(rc, 'rgl_plan_expand')
    _count(1 if E > 0 else 0)
    return nxt, rew


def plan_argmax(reward, V, E, A, gamma_bar, want_value=True, act_map=None):
    """value[E,A] = reward + gamma_bar * V; best[E] = first maximum (-1: none); with act_map[E,A] (int32) also
    best_action[E] = act_map[e, best[e]].  Returns (value, best, best_action)."""
    reward, V = _f32c(reward), _f32c(V)
    value = torch.empty(E, A, dtype=torch.float32, device=V.device) if want_value else None
    best = torch.empty(E, dtype=torch.int32, device=V.device)
    best_action = torch.empty(E, dtype=torch.int32, device=V.device) if act_map is not None else None
    if act_map is not None:
        assert act_map.dtype == torch.int32 and act_map.is_contiguous() and act_map.numel() == E * A
    with torch.cuda.device(V.device):
        rc = _lib.lib().rgl_plan_argmax(_lib.ptr(reward), _lib.ptr(V), E, A, float(gamma_bar), _lib.ptr(value),
                                        _lib.ptr(best), _lib.ptr(act_map), _lib.ptr(best_action), _lib.stream_ptr(V.device))
    _lib.check(rc, 'rgl_plan_argmax')
    _count(1 if E > 0 else 0)
    return value, best, (best_action if act_map is not None else best)


def plan_select(reward, V, E, A, gamma_bar, width, groups=None, next_robot=None, want_value=False):
    """action_clip for E states: -> (acts[E,width] int32, child_reward[E,width], child_robot[E*width,1,9] | None, value[E,A] | None)."""
    reward, V = _f32c(reward), _f32c(V)
    dev = V.device
    acts = torch.empty(E, width, dtype=torch.int32, device=dev)
    crew = torch.empty(E, width, dtype=torch.float32, device=dev)
    crob = torch.empty(E * width, 1, 9, dtype=torch.float32, device=dev) if next_robot is not None else None
    value = torch.empty(E, A, dtype=torch.float32, device=dev) if want_value else None
    if groups is not None:
        assert groups.dtype == torch.int32 and groups.is_contiguous() and groups.numel() == A
    with torch.cuda.device(dev):
        rc = _lib.lib().rgl_plan_select(_lib.ptr(reward), _lib.ptr(V), E, A, float(gamma_bar), int(width), _lib.ptr(groups),
                                        _lib.ptr(_f32c(next_robot)) if next_robot is not None else None, _lib.ptr(acts), _lib.ptr(crew),
                                        _lib.ptr(crob), _lib.ptr(value), _lib.stream_ptr(dev))
    _lib.check(rc, 'rgl_plan_select')
    _count(1 if E > 0 else 0)
    return acts, crew, crob, value


def plan_backup(v, next_v, reward, E, W, gamma_bar, depth):
    """V_planning backup: (ret_best[E], best[E] int32) over ret[e,k] = v/depth + (depth-1)/depth * (gamma_bar*next_v + reward)."""
    v, next_v, reward = _f32c(v), _f32c(next_v), _f32c(reward)
    ret = torch.empty(E, dtype=torch.float32, device=v.device)
    best = torch.empty(E, dtype=torch.int32, device=v.device)
    with torch.cuda.device(v.device):
        rc = _lib.lib().rgl_plan_backup(_lib.ptr(v), _lib.ptr(next_v), _lib.ptr(reward), E, W, float(gamma_bar), int(depth),
                                        _lib.ptr(ret), _lib.ptr(best), _lib.stream_ptr(v.device))
    _lib.check(rc, 'rgl_plan_backup')
    _count(1 if E > 0 else 0)
    return ret, best


# ---------------------------------------------------------------- autograd glue -------------------
def _needs_grad(*tensors_or_modules):
    if not torch.is_grad_enabled():
        return False
    for t in tensors_or_modules:
        if isinstance(t, torch.Tensor):
            if t.requires_grad:
                return True
        else:
            for p in t.parameters():
                if p.requires_grad:
                    return True
    return False


class _FusedForward(torch.autograd.Function):
    """Forward = sm_100a kernels; backward = autograd recompute with torch ops on the same GPU
    (placeholder until the fused backward kernels land; see _torch_math.py)."""

    @staticmethod
    def forward(ctx, run_kernels, run_torch, nparams, *tensors):
        ctx.run_torch = run_torch
        ctx.nparams = nparams
        ctx.save_for_backward(*tensors)
        return run_kernels()

    @staticmethod
    def backward(ctx, *grad_outs):
        tensors = ctx.saved_tensors
        with torch.enable_grad():
            outs = ctx.run_torch()
            if isinstance(outs, torch.Tensor):
                outs = (outs,)
            wanted = [t for t, need in zip(tensors, ctx.needs_input_grad[3:]) if need]
            grads = torch.autograd.grad(outs, wanted, grad_outs, allow_unused=True)
        it = iter(grads)
        res = [next(it) if need else None for need in ctx.needs_input_grad[3:]]
        return (None, None, None) + tuple(res)


def fused_with_autograd(run_kernels, run_torch, params, inputs):
    """run_kernels(): no-grad CUDA kernel path; run_torch(): differentiable torch restatement over the SAME
    parameter / input tensors (they are passed to the Function so autograd routes the gradients)."""
    tensors = list(params) + list(inputs)
    return _FusedForward.apply(run_kernels, run_torch, len(params), *tensors)
