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
    """Sum-all-reduce of all gradients of `params` through one flat buffer (one collective per optimizer step)."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self.buf = None

    def reduce(self):
        dev = self.params[0].device
        if self.buf is None or self.buf.device != dev:
            self.buf = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.buf[off:off + n].zero_()
            else:
                self.buf[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group)
        off = 0
        for p in self.params:
            n = p.numel()
            g = self.buf[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
        return self.buf


def dp_value_step(value_estimator, target_model, optimizer, reducer, robot, humans, rewards, next_robot, next_humans,
                  gamma_bar, global_batch):
    """One data-parallel value-network step on this rank's shard (trainer.py:122-131).

    MSELoss(mean) over the GLOBAL batch = sum over ranks of (local squared-error sum / global_batch), so each rank
    back-propagates `sum((out - target)^2) / global_batch` and the gradients are summed across ranks.
    Returns the local loss contribution (a tensor; sum over ranks = the global mean loss).
    """
    optimizer.zero_grad()
    out = value_estimator((robot, humans))
    with torch.no_grad():
        target = rewards + gamma_bar * target_model((next_robot, next_humans))
    loss = ((out - target) ** 2).sum() / float(global_batch)
    loss.backward()
    reducer.reduce()
    optimizer.step()
    return loss.detach()


def dp_state_predictor_step(state_predictor, optimizer, reducer, robot, humans, next_humans, global_batch, detach=False):
    """Data-parallel state-predictor step (trainer.py:143-149): MSE over the global [B,Nh,5] prediction."""
    optimizer.zero_grad()
    _, est = state_predictor((robot, humans), None, detach=detach)
    denom = float(global_batch) * est.size(1) * est.size(2)
    loss = ((est - next_humans) ** 2).sum() / denom
    loss.backward()
    reducer.reduce()
    optimizer.step()
    return loss.detach()
