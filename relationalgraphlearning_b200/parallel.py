"""Multi-GPU plumbing for the RGL hot path: one process per GPU, torch.distributed (NCCL on GPUs).

The path shards along the batch of independent states:
  * inference / planning: contiguous shard per rank, NO collective on the data path (an optional all_gather of
    the small per-state results when one rank needs them all);
  * training (crowd_nav/utils/trainer.py:118-132,143-149 in data-parallel form): local forward/backward on the
    shard, then ONE all-reduce over a single flat fp32 buffer holding every gradient (22 813 floats = 91 KB for
    the value estimator, 10 693 for the state predictor) -- latency-bound on NVLink/NVSwitch, so one call, not
    per-tensor buckets -- then the identical optimizer step on every rank.
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous, balanced [lo, hi) of `total` items for `rank` (first `total % world` ranks get one more)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_states(robot, humans, rank=None, world=None):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_range(robot.size(0), rank, world)
    return robot[lo:hi], humans[lo:hi]


def gather_results(local, total, group=None):
    """all_gather of per-state results with uneven shards -> tensor [total, ...] on every rank."""
    world = dist.get_world_size(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.size(0)] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)


class FlatGradAllReducer(object):
    """Generic sum-all-reduce of the gradients autograd left in `p.grad`, through one flat buffer (one collective per
    optimizer step).  Used for modules without the native backward (CPU / gloo tests, layerwise graphs); the product
    path is FlatGrads below, where no gather / scatter copies exist at all."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self.buf = None
        self.views = None

    def step_begin(self, optimizer=None):
        if optimizer is not None:
            optimizer.zero_grad()

    def reduce(self):
        dev = self.params[0].device
        if self.buf is None or self.buf.device != dev:
            self.buf = torch.zeros(self.numel, dtype=torch.float32, device=dev)
            self.views, off = [], 0
            for p in self.params:
                self.views.append(self.buf[off:off + p.numel()].view_as(p))
                off += p.numel()
        have = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        if len(have) < len(self.params):
            for v, p in zip(self.views, self.params):
                if p.grad is None:
                    v.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group)
        for v, p in zip(self.views, self.params):
            p.grad = v                      # the optimizer reads the reduced gradients in place (no scatter copies)
        return self.buf


class _DevMem(object):
    """Raw device memory as a __cuda_array_interface__ exporter (torch.as_tensor wraps it without copying)."""

    def __init__(self, ptr, nfloats):
        self.__cuda_array_interface__ = {'shape': (nfloats,), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}


class FlatGrads(object):
    """Every gradient of a ValueEstimator / StatePredictor in ONE flat fp32 buffer, written there by the backward kernels.

      accum   where training.py's native backward accumulates (rgl_linear_bwd dW / db atomics): `module._grad_sink`
      grad    the reduced gradients; every `p.grad` is a view into it, so the optimizer reads them in place
      reduce  world == 1: nothing (grad IS accum).  world > 1: ONE collective over the flat buffer --
              backend 'p2p'  : rgl_comm_allreduce, the one-kernel push all-reduce over NVLink peer memory
                               (csrc/dp_comm.cu; accum lives in the cudaIpc-shared workspace; CUDA-graph capturable),
              backend 'nccl' : torch.distributed.all_reduce on `accum` in place (also capturable).
    'auto' picks p2p when the peer-memory rendezvous succeeds on every rank and NCCL otherwise.
    """

    def __init__(self, module, group=None, backend='auto', params=None):
        from . import _lib
        self.module = module
        self.params = [p for p in (params if params is not None else module.parameters()) if p.requires_grad]
        self.group = group
        self.dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.offsets, off = {}, 0
        for p in self.params:
            self.offsets[id(p)] = off
            off += (p.numel() + 3) & ~3                  # 16-byte aligned views
        self.nfloats = off
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.comm = None
        self.backend = 'local'
        self.fallback_reason = None
        if self.world > 1:
            want = backend
            if want in ('auto', 'p2p') and self.dev.type == 'cuda':
                try:
                    self._open_p2p(_lib)
                    self.backend = 'p2p'
                except Exception as e:  # noqa: BLE001
                    if want == 'p2p':
                        raise
                    self.fallback_reason = str(e)
            if self.backend != 'p2p':
                self.backend = 'nccl'
        if self.backend == 'p2p':
            ptr = _lib.lib().rgl_comm_accum_ptr(self.comm)
            self._mem = _DevMem(ptr, self.nfloats)
            self.accum = torch.as_tensor(self._mem, device=self.dev)
            self.grad = torch.zeros(self.nfloats, dtype=torch.float32, device=self.dev)
        else:
            self.accum = torch.zeros(self.nfloats, dtype=torch.float32, device=self.dev)
            self.grad = self.accum
        self.accum_views = {id(p): self.accum[self.offsets[id(p)]:self.offsets[id(p)] + p.numel()].view_as(p) for p in self.params}
        self.grad_views = {id(p): self.grad[self.offsets[id(p)]:self.offsets[id(p)] + p.numel()].view_as(p) for p in self.params}
        self.message_bytes = self.numel * 4
        module._grad_sink = self
        self._install()

    # ---- peer-memory rendezvous: create the workspace, all-gather the 64-byte cudaIpc handles, map the peers ----
    def _open_p2p(self, _lib):
        import ctypes
        lib = _lib.lib()
        with torch.cuda.device(self.dev):
            h = ctypes.c_void_p()
            _lib.check(lib.rgl_comm_create(self.rank, self.world, self.nfloats, ctypes.byref(h)), 'rgl_comm_create', comm=True)
            nb = lib.rgl_comm_handle_bytes()
            mine = (ctypes.c_ubyte * nb)()
            _lib.check(lib.rgl_comm_ipc_handle(h, mine), 'rgl_comm_ipc_handle', comm=True)
            t = torch.tensor(list(mine), dtype=torch.uint8, device=self.dev)
            outs = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(outs, t, group=self.group)
            allh = (ctypes.c_ubyte * (nb * self.world))(*[int(x) for o in outs for x in o.cpu().tolist()])
            rc = lib.rgl_comm_open_peers(h, allh)
            ok = torch.tensor([1 if rc == 0 else 0], device=self.dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)      # all ranks agree on the backend
            if int(ok) == 0:
                msg = lib.rgl_comm_last_error_string()
                lib.rgl_comm_destroy(h)
                raise RuntimeError('peer-memory rendezvous failed on some rank: %s' % (msg.decode() if msg else ''))
            self.comm = h

    def _install(self):
        for p in self.params:
            v = self.grad_views[id(p)]
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v

    def covers(self, plist):
        return all(id(p) in self.offsets for p in plist)

    def accum_view(self, p):
        return self.accum_views[id(p)]

    def step_begin(self, optimizer=None):
        """Replaces optimizer.zero_grad(): one fill of the flat buffer (p2p: the all-reduce kernel already re-zeroed it)."""
        if self.backend != 'p2p':
            self.accum.zero_()
        self._install()

    def reduce(self):
        from . import _lib
        if self.backend == 'p2p':
            with torch.cuda.device(self.dev):
                _lib.check(_lib.lib().rgl_comm_allreduce(self.comm, _lib.ptr(self.grad), 1.0, _lib.stream_ptr(self.dev)),
                           'rgl_comm_allreduce', comm=True)
            from . import ops
            ops._count(1)
        elif self.backend == 'nccl':
            dist.all_reduce(self.accum, op=dist.ReduceOp.SUM, group=self.group)
        return self.grad

    def flat_grad(self):
        """Dense copy of the reduced gradients in parameter order (tests / checks)."""
        return torch.cat([self.grad_views[id(p)].reshape(-1) for p in self.params])

    def status(self):
        """0 = ok; 1 = a peer did not arrive at an all-reduce within ~10 s (synchronises)."""
        if self.comm is None:
            return 0
        import ctypes
        from . import _lib
        st = ctypes.c_int(0)
        _lib.check(_lib.lib().rgl_comm_status(self.comm, ctypes.byref(st)), 'rgl_comm_status', comm=True)
        return int(st.value)

    def close(self):
        if getattr(self.module, '_grad_sink', None) is self:
            self.module._grad_sink = None
        if self.comm is not None:
            from . import _lib
            for p in self.params:
                p.grad = None
            self.accum_views = self.accum = self._mem = None
            _lib.lib().rgl_comm_destroy(self.comm)
            self.comm = None


_SIDE_STREAMS = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device)
    return _SIDE_STREAMS[key]


def dp_value_step(value_estimator, target_model, optimizer, reducer, robot, humans, rewards, next_robot, next_humans,
                  gamma_bar, global_batch):
    """One data-parallel value-network step on this rank's shard (trainer.py:122-131).  `reducer`: FlatGrads (native
    backward writes the flat buffer directly) or FlatGradAllReducer (generic).

    MSELoss(mean) over the GLOBAL batch = sum over ranks of (local squared-error sum / global_batch), so each rank
    back-propagates `sum((out - target)^2) / global_batch` and the gradients are summed across ranks.
    Returns the local loss contribution (a tensor; sum over ranks = the global mean loss).
    """
    reducer.step_begin(optimizer)
    if robot.is_cuda:
        # the target-network forward does not depend on the online forward: it runs on a side stream (a parallel branch of
        # the CUDA graph when the step is captured) and joins before the loss
        cur = torch.cuda.current_stream(robot.device)
        side = _side_stream(robot.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            v_next = target_model((next_robot, next_humans))
        out = value_estimator((robot, humans))
        cur.wait_stream(side)
        v_next.record_stream(cur)
    else:
        out = value_estimator((robot, humans))
        with torch.no_grad():
            v_next = target_model((next_robot, next_humans))
    if out.is_cuda:
        from . import training
        loss = training.td_loss(out, rewards, v_next, gamma_bar, global_batch)       # one launch: target, loss, dLoss/dV
    else:
        loss = ((out - (rewards + gamma_bar * v_next)) ** 2).sum() / float(global_batch)
    loss.backward()
    reducer.reduce()
    optimizer.step()
    return loss.detach()


def dp_state_predictor_step(state_predictor, optimizer, reducer, robot, humans, next_humans, global_batch, detach=False):
    """Data-parallel state-predictor step (trainer.py:143-149): MSE over the global [B,Nh,5] prediction."""
    reducer.step_begin(optimizer)
    _, est = state_predictor((robot, humans), None, detach=detach)
    denom = float(global_batch) * est.size(1) * est.size(2)
    loss = ((est - next_humans) ** 2).sum() / denom
    loss.backward()
    reducer.reduce()
    optimizer.step()
    return loss.detach()
