#!/usr/bin/env python
"""bench.py -- RGL graph-forward states/sec (batch 4096, 5 humans) on N B200s, one rank per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload graph|value|statepred|train|plan]

A step = one pass of the hot path over one batch of `--batch` synthetic states (SURVEY.md 8(d) distribution).
  value      device-resident throughput: inputs already in HBM, the K steps of a pass replayed from CUDA graphs, timed with
             CUDA events on the launching stream, max over ranks.  A pass is repeated until >= 50 ms are timed (`replays`);
             the per-pass median / min / max are reported next to the total.
  e2e        same metric through the host-buffer API (relationalgraphlearning_b200.hostio.HostStream):
             every step copies its states from pinned host memory and reads the result back; every result is consumed.
  roofline   the dominant kernel against the roof that binds it (fp32-FMA + tensor pipe: 80 FLOP/B), with the HBM
             fraction on algorithmic bytes alongside.
  extra      the other BASELINE configs in the same run: train_c4 (value-net training step B=8192 Nh=10 per GPU, ONE flat
             gradient all-reduce of 91 252 B per step when N > 1), plan_c3 / plan_c5 (look-ahead trees), value path,
             stand-alone GCN layer (the HBM-bound unit), steady state.
  cpu_baseline / --impl reference: the reference's own modules (oracle/_ref, vendored by oracle/make_ref.py; kind
             "reference") or, if absent, the bit-pinned oracle port (kind "port") on the host cores.
L2 policy: inputs rotate through a pool larger than L2 (126 MB); consecutive passes replay DIFFERENT graphs that walk
different slices of the pool, so no pass re-reads states (or re-writes outputs) that are still L2-resident.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

L2_BYTES = 126 * 1024 * 1024
FMA_PEAK_TFLOPS = 72.6         # measured on this pool's B200 with tools/fma_peak.cu (36.3 TFMA/s, 124.8 FMA/clk/SM)
MIN_TIMED_MS = 50.0


def algorithmic(workload, nh):
    """(bytes, flops) per state: compulsory HBM I/O and 2*MAC."""
    n = nh + 1
    mac_graph = 2624 + 2368 * nh + (1024 * n + 32 * n * n) * 3
    b_in = 36 + 20 * nh
    if workload == 'graph':
        return b_in + 128 * n, 2 * mac_graph
    if workload == 'value':
        return b_in + 4, 2 * (mac_graph + 14324)
    return b_in + 20 * nh, 2 * (mac_graph + 2368 * n)


def pool_batches(B, nh):
    return max(8, (int(1.15 * L2_BYTES) + B * (36 + 20 * nh) - 1) // (B * (36 + 20 * nh)))


def build_modules(seed=0):
    from relationalgraphlearning_b200.config import policy_config
    from relationalgraphlearning_b200.graph_model import RGL
    from relationalgraphlearning_b200.state_predictor import StatePredictor
    from relationalgraphlearning_b200.value_estimator import ValueEstimator
    cfg = policy_config()
    torch.manual_seed(seed)
    g1 = RGL(cfg, 9, 5)
    ve = ValueEstimator(cfg, g1)
    g2 = RGL(cfg, 9, 5)
    sp = StatePredictor(cfg, g2, 0.25)
    return g1, ve, g2, sp


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.stop_flag = index, period, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {'hw_slowdown': nv.nvmlClocksThrottleReasonHwSlowdown,
                 'hw_thermal_slowdown': nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 'sw_thermal_slowdown': nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 'sw_power_cap': nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s)}


# --------------------------------------------------------------------------------------------------------- CPU legs
def _malloc_best_case():
    # The CPU path allocates ~20 MB of temporaries per step; with glibc's default thresholds every one of them is mmap'ed and
    # page-faulted afresh (measured 2.3-2.6x slower).  Give the CPU leg its best case: keep freed memory in the heap.
    try:
        import ctypes
        libc = ctypes.CDLL('libc.so.6')
        libc.mallopt(-3, 1 << 30)       # M_MMAP_THRESHOLD
        libc.mallopt(-1, 1 << 30)       # M_TRIM_THRESHOLD
    except Exception:  # noqa: BLE001
        pass


def cpu_forward_step_fn(workload):
    """(step(robot, humans), kind): the reference's own modules from oracle/_ref when vendored (kind 'reference'), else the
    bit-pinned oracle port (kind 'port').  Same seed => same weights as build_modules(0) (same construction order)."""
    from oracle import make_ref
    if make_ref.enable():
        from crowd_nav.configs.icra_benchmark.mp_separate import PolicyConfig
        from crowd_nav.policy.graph_model import RGL
        from crowd_nav.policy.state_predictor import StatePredictor
        from crowd_nav.policy.value_estimator import ValueEstimator
        import logging
        logging.disable(logging.INFO)
        cfg = PolicyConfig()
        torch.manual_seed(0)
        g1 = RGL(cfg, 9, 5)
        ve = ValueEstimator(cfg, g1)
        g2 = RGL(cfg, 9, 5)
        sp = StatePredictor(cfg, g2, 0.25)
        fn = {'graph': lambda r, h: g1((r, h)), 'value': lambda r, h: ve((r, h)), 'statepred': lambda r, h: sp((r, h), None)[1]}[workload]
        return fn, 'reference'
    from oracle import rgl_oracle as O
    g1, ve, g2, sp = build_modules(0)
    sd = [{k: v.detach().clone() for k, v in m.state_dict().items()} for m in (g1, ve.value_network, g2, sp.human_motion_predictor)]
    fn = {'graph': lambda r, h: O.rgl_forward(sd[0], r, h), 'value': lambda r, h: O.value_forward(sd[0], sd[1], r, h),
          'statepred': lambda r, h: O.statepred_forward(sd[2], sd[3], r, h)}[workload]
    return fn, 'port'


def cpu_reference_rate(workload, batch, nh, steps, warmup, budget_s=150.0):
    """states/s of the reference CPU path on the host cores -> dict."""
    from relationalgraphlearning_b200.synthetic import synthetic_states
    _malloc_best_case()
    fn, kind = cpu_forward_step_fn(workload)
    pool = [synthetic_states(batch, nh, seed=100 + i) for i in range(8)]
    avail = host_cores()

    def step(i):
        r, h = pool[i % len(pool)]
        with torch.no_grad():
            return fn(r, h)

    # torchrun exports OMP_NUM_THREADS=1; the baseline gets every host core it can use.  Small batched ops do not
    # always scale to all cores, so the thread count is calibrated and the FASTEST setting is the one reported.
    best_t, cores = None, avail
    for nt in sorted({avail, max(1, avail // 2), max(1, avail // 4), min(avail, 16), min(avail, 8)}, reverse=True):
        torch.set_num_threads(nt)
        for i in range(3):
            step(i)
        t0 = time.perf_counter()
        for i in range(5):
            step(i)
        dt = (time.perf_counter() - t0) / 5
        if best_t is None or dt < best_t:
            best_t, cores = dt, nt
    torch.set_num_threads(cores)
    for i in range(max(warmup, 3)):
        step(i)
    t0 = time.perf_counter()
    step(0)
    t1 = time.perf_counter() - t0
    k = steps
    sample = '%d steps x full batch %d' % (k, batch)
    if t1 * k > budget_s:
        k = max(10, int(budget_s / t1))
        sample = '%d of %d steps x full batch %d (bounded to ~%.0f s of CPU work)' % (k, steps, batch, budget_s)
    t0 = time.perf_counter()
    for i in range(k):
        step(i)
    dt = time.perf_counter() - t0
    return {'value': batch * k / dt, 'unit': 'states/s', 'cores': cores, 'cores_available': avail, 'kind': kind,
            'sample': sample, 'ms_per_step': dt / k * 1e3, 'steps': k,
            'threads_note': 'cores = torch threads of the fastest calibrated setting; cores_available = host cores visible to the process'}


# --------------------------------------------------------------------------------------------------------- helpers
class Dist(object):
    def __init__(self, rank, world, local):
        self.rank, self.world, self.local = rank, world, local
        self.dev = torch.device('cuda', local)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_(self, values):
        t = torch.tensor(values, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def sum_(self, values):
        t = torch.tensor(values, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t]


def timed_passes(D, launch_pass, est_ms, min_ms=MIN_TIMED_MS, max_passes=400):
    """Repeat `launch_pass(p)` until >= min_ms are timed (same count on every rank).  Returns (total_ms, [per-pass ms], R)."""
    est = D.max_([est_ms])[0]
    R = int(min(max_passes, max(3, math.ceil(min_ms / max(est, 1e-3)))))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(R + 1)]
    D.barrier()
    ev[0].record()
    for p in range(R):
        launch_pass(p)
        ev[p + 1].record()
    D.barrier()
    per = [ev[p].elapsed_time(ev[p + 1]) for p in range(R)]
    return ev[0].elapsed_time(ev[R]), per, R


def spread(per):
    s = sorted(per)
    return {'median_ms': s[len(s) // 2], 'min_ms': s[0], 'max_ms': s[-1], 'passes': len(s)}


# --------------------------------------------------------------------------------------------------------- train (C4)
def measure_train(D, B, nh, K, W, backend='auto', cpu_leg=False, check=True):
    """BASELINE configs[3]: value-net training step (forward + target forward + MSE + native backward + Adam), `B` samples
    per GPU; the backward kernels accumulate every gradient into ONE flat buffer and, when world > 1, one collective sums
    it across ranks per step (crowd_nav/utils/trainer.py:122-131 in data-parallel form).  The whole step, collective
    included, is replayed from a CUDA graph."""
    import copy
    from relationalgraphlearning_b200 import ops, parallel
    from relationalgraphlearning_b200.synthetic import synthetic_states
    dev, rank, world = D.dev, D.rank, D.world
    g1, ve, _, _ = build_modules(0)
    ve.to(dev)
    target = copy.deepcopy(ve)
    init = copy.deepcopy(ve) if (check and world > 1) else None
    red = parallel.FlatGrads(ve, backend=backend)
    opt = torch.optim.Adam(ve.parameters(), lr=1e-3, fused=True, capturable=True)
    gamma_bar = pow(0.9, 0.25)

    def make(rk, i):
        r, h = synthetic_states(B, nh, seed=10 + i + 100 * rk, device=dev)
        r2, h2 = synthetic_states(B, nh, seed=50 + i + 100 * rk, device=dev)
        gen = torch.Generator().manual_seed(7 + i + 100 * rk)
        return (r, h, (torch.rand(B, 1, generator=gen) * 1.25 - 0.25).to(dev), r2, h2)

    pool = [make(rank, i) for i in range(8)]

    def step(i):
        return parallel.dp_value_step(ve, target, opt, red, *pool[i % len(pool)], gamma_bar, B * world)

    # ---- DP check (world > 1): the first DP step's loss and all-reduced gradient against the single-GPU step on the
    # concatenated global batch, from the same initial weights ----
    dp_check = None
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        loss0 = step(0)
        if init is not None:
            import torch.distributed as dist
            ltot = loss0.clone()
            dist.all_reduce(ltot)
            got = red.flat_grad().clone()
            if rank == 0:
                alld = [make(rk, 0) for rk in range(world)]
                cat = [torch.cat([d[j] for d in alld], dim=0) for j in range(5)]
                out = init((cat[0], cat[1]))
                with torch.no_grad():
                    tgt = cat[2] + gamma_bar * target((cat[3], cat[4]))
                lref = torch.nn.functional.mse_loss(out, tgt)
                lref.backward()
                ref = torch.cat([p.grad.reshape(-1) for p in init.parameters()])
                gerr = float((got - ref).abs().max() / ref.abs().max())
                lerr = float((ltot - lref.detach()).abs() / lref.detach().abs())
                dp_check = {'grad_max_err_over_max': gerr, 'loss_rel_err': lerr, 'ok': bool(gerr <= 2e-4 and lerr <= 1e-5),
                            'what': 'first DP step (loss all-reduced, flat gradient after the collective) vs the single-GPU '
                                    'step on the concatenated %d-sample batch' % (B * world)}
                assert dp_check['ok'], dp_check
                del alld, cat, out, tgt, ref
            init = None
            dist.barrier()                                   # the other ranks wait here while rank 0 checks
        for i in range(1, W):
            step(i)
    torch.cuda.synchronize()
    D.barrier()
    G = max(d for d in range(1, 9) if K % d == 0)
    graphs, captured = [], True
    l0 = ops.LAUNCHES
    try:
        for j in range(2):                                   # two graphs walking different pool entries
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    for i in range(G):
                        loss = step(W + j * G + i)
            graphs.append(g)
    except Exception as e:  # noqa: BLE001  (a collective that cannot be captured: run the steps eagerly)
        captured, graphs = False, []
        capture_error = str(e).splitlines()[0][:200]
        torch.cuda.synchronize()
    launches_per_step = (ops.LAUNCHES - l0) / (2.0 * G) if captured else None
    reps = K // G

    def launch_pass(p):
        nonlocal loss
        if captured:
            for q in range(reps):
                graphs[(p * reps + q) % 2].replay()
        else:
            with torch.cuda.stream(side):
                for i in range(K):
                    loss = step(W + i)
            torch.cuda.current_stream().wait_stream(side)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    e0.record()
    launch_pass(0)
    e1.record()
    D.barrier()
    total_ms, per, R = timed_passes(D, launch_pass, e0.elapsed_time(e1))
    total_ms = D.max_([total_ms])[0]
    status = red.status()
    final_loss = float(loss)
    out = None
    if rank == 0:
        out = {'metric': 'RGL value-net training samples/sec (batch %d per GPU, %d humans)' % (B, nh),
               'value': world * B * K * R / (total_ms * 1e-3), 'unit': 'samples/s', 'n_gpus': world, 'steps': K, 'warmup': W,
               'replays': R, 'ms_per_step': total_ms / (K * R), 'pass': spread(per), 'higher_is_better': True, 'scaling': 'weak',
               'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
               'gpu_launches': None if launches_per_step is None else int(round(launches_per_step * K * R)),
               'final_loss': final_loss, 'graph_captured': captured,
               'grad_allreduce_bytes': red.message_bytes if world > 1 else 0,
               'grad_allreduce_backend': red.backend, 'comm_status': status,
               'config': {'workload': 'value-net train step B=%d Nh=%d (BASELINE configs[3]): fused fwd+saves, native bwd into one '
                                      'flat gradient buffer, Adam' % (B, nh),
                          'parallelism': 'dp%d, one flat all-reduce per step (%s)' % (world, red.backend)}}
        if red.fallback_reason:
            out['grad_allreduce_fallback_reason'] = red.fallback_reason
        if not captured:
            out['capture_error'] = capture_error
        if dp_check is not None:
            out['dp_check'] = dp_check
        if cpu_leg:
            from oracle import rgl_oracle as O
            sd = [{k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()} for m in (g1, ve.value_network)]
            sdt = [{k: v.detach().cpu().clone() for k, v in m.state_dict().items()} for m in (g1, ve.value_network)]
            copt = torch.optim.Adam(list(sd[0].values()) + list(sd[1].values()), lr=1e-3)
            cpool = [tuple(t.cpu() for t in p) for p in pool[:2]]
            torch.set_num_threads(min(16, host_cores()))     # small autograd ops do not scale past ~16 threads

            def cstep(i):
                r, h, rew, r2, h2 = cpool[i % 2]
                copt.zero_grad()
                out_ = O.value_forward(sd[0], sd[1], r, h)
                tgt_ = rew + gamma_bar * O.value_forward(sdt[0], sdt[1], r2, h2)
                torch.nn.functional.mse_loss(out_, tgt_).backward()
                copt.step()
            for i in range(2):
                cstep(i)
            t0 = time.perf_counter()
            n = 0
            while time.perf_counter() - t0 < 8.0:
                cstep(n)
                n += 1
            dt = time.perf_counter() - t0
            out['cpu_baseline'] = {'value': B * n / dt, 'unit': 'samples/s', 'cores': torch.get_num_threads(), 'cores_available': host_cores(),
                                   'kind': 'port', 'sample': '%d full train steps (~8 s)' % n, 'ms_per_step': dt / n * 1e3,
                                   'note': 'the reference modules cannot back-propagate with skip_connection=True on this torch '
                                           '(in-place skip add, SURVEY.md 5): the oracle port with the out-of-place add is timed'}
    del graphs
    red.close()
    return out


# --------------------------------------------------------------------------------------------------------- planner (C3 / C5)
def measure_plan(D, E, nh, K, W, depth, width, speed_samples, rotation_samples, graphed=True, cpu_leg=False, label=''):
    """BASELINE configs[2] / [4]: d-step, w-width look-ahead of E root states per GPU (sharded over ranks, no collective).
    Reports root states planned per second and rollout states (value / state-predictor evaluations) per second."""
    from relationalgraphlearning_b200 import ops
    from relationalgraphlearning_b200.config import policy_config
    from relationalgraphlearning_b200.model_predictive_rl import ModelPredictiveRL
    from relationalgraphlearning_b200.synthetic import synthetic_states
    dev, rank, world = D.dev, D.rank, D.world
    cfg = policy_config(planning_depth=depth, planning_width=width, do_action_clip=depth > 1 or width > 1,
                        speed_samples=speed_samples, rotation_samples=rotation_samples)
    torch.manual_seed(0)
    pol = ModelPredictiveRL()
    pol.time_step = 0.25
    pol.configure(cfg)
    pol.set_time_step(0.25)
    pol.set_device(dev)
    pol.set_phase('test')
    pol.build_action_space(1.0)
    pool = [synthetic_states(E, nh, seed=900 + i + 100 * rank, device=dev) for i in range(4)]
    # predict() replays the whole look-ahead from a CUDA graph captured once per input shape (ModelPredictiveRL.use_cuda_graphs,
    # the policy's default); graphed=False issues the launches one by one from Python instead.
    run = (lambda r, h: pol.predict_batch_graphed(r, h)[0]) if graphed else pol.predict_batch
    pol.stat_value_states = pol.stat_sp_states = 0
    l0 = ops.LAUNCHES
    pol.predict_batch(*pool[0])                      # one eager pass: launches and rollout states of one step
    launches_per_step = ops.LAUNCHES - l0
    rollout = pol.stat_value_states + pol.stat_sp_states
    for i in range(W):
        run(*pool[i % 4])

    def launch_pass(p):
        for i in range(K):
            run(*pool[(p * K + i) % 4])

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    e0.record()
    launch_pass(0)
    e1.record()
    D.barrier()
    total_ms, per, R = timed_passes(D, launch_pass, e0.elapsed_time(e1), max_passes=100)
    total_ms = D.max_([total_ms])[0]
    out = None
    if rank == 0:
        out = {'metric': 'model_predictive_rl look-ahead: root states planned/sec (d=%d, w=%d, %d actions, %d humans)' %
                         (depth, width, len(pol.action_space), nh),
               'value': world * E * K * R / (total_ms * 1e-3), 'unit': 'root states/s', 'n_gpus': world, 'steps': K, 'warmup': W,
               'replays': R, 'ms_per_step': total_ms / (K * R), 'pass': spread(per), 'higher_is_better': True, 'scaling': 'weak',
               'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'gpu_launches': launches_per_step * K * R,
               'launches_per_step': launches_per_step, 'graph_replay': graphed,
               'rollout_states_per_step': rollout, 'rollout_states_per_s': world * rollout * K * R / (total_ms * 1e-3),
               'config': {'workload': '%splanner tree d=%d w=%d, %d root states per GPU' % (label, depth, width, E),
                          'parallelism': 'dp%d (root states sharded, no collective)' % world}}
        if cpu_leg:
            from oracle import planner_oracle as P
            sd = pol.get_state_dict()
            cpu = lambda d: {k: v.detach().cpu() for k, v in d.items()}   # noqa: E731
            orc = P.OraclePlanner(cpu(sd['graph_model1']), cpu(sd['value_network']), cpu(sd['graph_model2']), cpu(sd['motion_predictor']),
                                  planning_depth=depth, planning_width=width, do_action_clip=cfg.model_predictive_rl.do_action_clip,
                                  speed_samples=speed_samples, rotation_samples=rotation_samples)
            torch.set_num_threads(1)          # batch-1 forwards: more threads only add overhead
            r, h = pool[0][0].cpu(), pool[0][1].cpu()
            t0 = time.perf_counter()
            nroots = 0
            while time.perf_counter() - t0 < 8.0 and nroots < E:
                orc.predict(r[nroots:nroots + 1], h[nroots:nroots + 1])
                nroots += 1
            dt = time.perf_counter() - t0
            out['cpu_baseline'] = {'value': nroots / dt, 'unit': 'root states/s', 'cores': 1, 'cores_available': host_cores(), 'kind': 'port',
                                   'sample': '%d root states through the batch-1 oracle tree (~8 s)' % nroots}
    pol._graphs = {}
    return out


# --------------------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='graph', choices=['graph', 'value', 'statepred', 'train', 'plan'])
    ap.add_argument('--roots', type=int, default=1024)
    ap.add_argument('--plan-eager', action='store_true', help='plan workload: launch from Python instead of replaying the captured CUDA graph')
    ap.add_argument('--depth', type=int, default=2)
    ap.add_argument('--width', type=int, default=2)
    ap.add_argument('--speed-samples', type=int, default=2)
    ap.add_argument('--rotation-samples', type=int, default=5)
    ap.add_argument('--batch', type=int, default=4096)
    ap.add_argument('--humans', type=int, default=5)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='graph workload: skip the train_c4 / plan_c3 / plan_c5 extras')
    ap.add_argument('--dp-comm', default='auto', choices=['auto', 'p2p', 'nccl'], help='train: gradient all-reduce backend')
    ap.add_argument('--zero-copy', action='store_true', help='e2e: the kernel writes H into pinned host memory itself')
    ap.add_argument('--host-depth', type=int, default=4, help='e2e: batches in flight through hostio.HostStream')
    ap.add_argument('--streams', type=int, default=4, help='CUDA streams the independent steps are issued on')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    B, nh, K, W = args.batch, args.humans, args.steps, max(args.warmup, 3)
    abytes, aflops = algorithmic(args.workload if args.workload in ('graph', 'value', 'statepred') else 'graph', nh)
    metric = 'RGL graph-forward states/sec (batch %d, %d humans)' % (B, nh)
    pool_n = pool_batches(B, nh)
    config = {'workload': 'rgl_%s_forward B=%d Nh=%d 2-layer GCN fp32 (BASELINE configs[1])' % (args.workload, B, nh),
              'batch_per_gpu': B, 'humans': nh, 'parallelism': 'dp%d (batch sharded, no collective)' % world,
              'l2_policy': 'inputs rotate through a pool > 126 MB L2; consecutive passes replay graphs over different pool slices',
              'streams': args.streams, 'input_pool_mb': round(pool_n * B * (36 + 20 * nh) / 1e6, 1)}

    if args.impl == 'reference':
        if rank != 0:
            return
        if args.workload in ('plan', 'train'):
            print(json.dumps({'impl': 'reference', 'unavailable': '%s workload: the CPU leg is reported inside the ours arm (cpu_baseline)' % args.workload}))
            return
        c = cpu_reference_rate(args.workload, B, nh, K, W)
        print(json.dumps({'impl': 'reference', 'metric': metric, 'value': c['value'], 'unit': 'states/s', 'n_gpus': args.gpus,
                          'steps': c['steps'], 'warmup': W, 'ms_per_step': c['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
                          'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
                          'cpu_baseline': {k: c[k] for k in ('value', 'unit', 'cores', 'cores_available', 'kind', 'sample')},
                          'e2e': {'value': c['value'], 'unit': 'states/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                          'gpu_launches': 0}))
        return

    import torch.distributed as dist
    from relationalgraphlearning_b200 import ops
    from relationalgraphlearning_b200.hostio import HostStream
    from relationalgraphlearning_b200.synthetic import synthetic_states

    assert torch.cuda.is_available(), 'bench.py (impl ours) needs a CUDA device'
    D = Dist(rank, world, local)
    dev = D.dev
    torch.cuda.set_device(dev)
    if world > 1:
        # one rank per GPU: give every rank its own slice of the host cores (the host-buffer path is fed by the CPU)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]) or set(cores))
        except Exception:  # noqa: BLE001
            pass
        dist.init_process_group('nccl', device_id=dev)

    if args.workload == 'train':
        out = measure_train(D, B, nh, K, W, backend=args.dp_comm, cpu_leg=(world == 1 and not args.no_cpu_baseline))
        if rank == 0:
            print(json.dumps(out))
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload == 'plan':
        out = measure_plan(D, args.roots, nh, K, W, args.depth, args.width, args.speed_samples, args.rotation_samples,
                           graphed=not args.plan_eager, cpu_leg=(world == 1 and not args.no_cpu_baseline))
        if rank == 0:
            print(json.dumps(out))
        if world > 1:
            dist.destroy_process_group()
        return

    g1, ve, g2, sp = build_modules(0)
    ve.to(dev)
    sp.to(dev)
    module = {'graph': g1, 'value': ve, 'statepred': sp}[args.workload]
    thr = args.streams > 1

    def run_step(robot, humans):
        if args.workload == 'graph':
            return g1.run(robot, humans, want_H=True, throughput=thr)['H']
        return module.run(robot, humans, throughput=thr)

    # ---- input pool larger than L2 (distinct states per rank) ----
    rb, hb = synthetic_states(pool_n * B, nh, seed=1234 + rank)
    robots = [rb[i * B:(i + 1) * B].contiguous() for i in range(pool_n)]
    humans = [hb[i * B:(i + 1) * B].contiguous() for i in range(pool_n)]
    robots_d = [r.to(dev) for r in robots]
    humans_d = [h.to(dev) for h in humans]

    # ---- warm-up (eager) + CUDA-graph capture of the passes ----
    with torch.no_grad():
        for i in range(W):
            run_step(robots_d[i % pool_n], humans_d[i % pool_n])
    torch.cuda.synchronize()
    G = K
    if K > 500:
        G = max(d for d in range(1, 501) if K % d == 0)
    reps = K // G
    n_graphs = int(min(16, max(1, math.ceil(pool_n / G))))
    # Steps are independent batches, so consecutive steps are issued round-robin on `--streams` CUDA streams (forked
    # from / joined to the capture stream): the tail of step i overlaps the head of step i+1 on the GPU.
    nstreams = max(1, args.streams)
    side = torch.cuda.Stream()
    branches = [torch.cuda.Stream() for _ in range(nstreams)]

    def capture(run, first, count, multi):
        g = torch.cuda.CUDAGraph()
        keep = []
        with torch.no_grad(), torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                if multi:
                    fork = torch.cuda.Event()
                    fork.record(side)
                    for b in branches:
                        b.wait_event(fork)
                for i in range(count):
                    idx = (first + i) % pool_n
                    if multi:
                        with torch.cuda.stream(branches[i % nstreams]):
                            keep.append(run(robots_d[idx], humans_d[idx]))
                    else:
                        keep.append(run(robots_d[idx], humans_d[idx]))
                    if len(keep) >= 64:     # rotate output buffers: a 64-deep ring (> L2 for the H output) instead of `count` live tensors
                        keep = keep[32:]
                if multi:
                    for b in branches:
                        ev = torch.cuda.Event()
                        ev.record(b)
                        side.wait_event(ev)
        return g, keep

    with torch.no_grad(), torch.cuda.stream(side):
        run_step(robots_d[0], humans_d[0])
    torch.cuda.synchronize()
    l0 = ops.LAUNCHES
    graphs = [capture(run_step, W + j * G, G, True) for j in range(n_graphs)]
    launches_per_step = (ops.LAUNCHES - l0) / float(G * n_graphs)
    for g, _ in graphs:                 # untimed: uploads every graph (and leaves only the LAST graphs' slices in L2)
        g.replay()
    torch.cuda.synchronize()

    def launch_pass(p):
        for q in range(reps):
            graphs[(p * reps + q) % n_graphs][0].replay()

    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    e0.record()
    launch_pass(0)
    e1.record()
    D.barrier()
    t_ms, per_pass, R = timed_passes(D, lambda p: launch_pass(p + 1), e0.elapsed_time(e1))
    pass_stats = spread(per_pass)

    # ---- dominant kernel alone: back-to-back launches on ONE stream (no overlap between launches) walking the whole input
    # pool (> L2), CUDA events on that stream -> average launch duration for the roofline ----
    KG = int(min(pool_n, 300))
    kgraph, keep3 = capture(run_step, 0, KG, False)
    with torch.cuda.stream(side):
        kgraph.replay()
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(side)
        for _ in range(5):
            kgraph.replay()
        k1.record(side)
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / (5 * KG)          # per step on one stream (graph workload: exactly one kernel per step)
    del kgraph, keep3

    # ---- end to end through the host-buffer API: pinned host -> device -> kernels -> pinned host; every result is consumed
    # by the host before its slot is reused (back-pressure: at most `depth` batches in flight) ----
    npin = min(pool_n, 48)
    robots_p = [r.pin_memory() for r in robots[:npin]]
    humans_p = [h.pin_memory() for h in humans[:npin]]
    depth = args.host_depth

    def e2e_run(hstream, steps):
        for i in range(max(W, npin * depth)):            # warm-up also captures the per-(slot, buffer) graphs
            hstream.result(hstream.submit(robots_p[i % npin], humans_p[i % npin]))
        hstream.drain()
        D.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(hstream.streams[0])
        checksum, inflight = 0.0, []
        t0 = time.perf_counter()
        nsub = 0
        while True:
            if nsub < steps:
                inflight.append(hstream.submit(robots_p[nsub % npin], humans_p[nsub % npin]))
                nsub += 1
            if len(inflight) >= depth or (nsub >= steps and inflight):
                checksum += float(hstream.result_numpy(inflight.pop(0)).flat[0])     # the host reads every result
            if nsub >= steps and not inflight:
                break
        hstream.drain()
        for s in hstream.streams[1:]:
            hstream.streams[0].wait_stream(s)
        s1.record(hstream.streams[0])
        D.barrier()
        return s0.elapsed_time(s1), (time.perf_counter() - t0) * 1e3, checksum

    K_e2e = max(K, 400)                                  # >= 25 ms of PCIe traffic at the ~65 us a 3.1 MB result takes
    hs = HostStream(args.workload, module, B, nh, dev, depth=depth, zero_copy_out=args.zero_copy)
    e2e_ms, e2e_host_ms, _ = e2e_run(hs, K_e2e)

    # the box's aggregate device->host ceiling with every rank copying at once (what bounds e2e for the 3.1 MB H result)
    d2h_gbs = 0.0
    if args.workload == 'graph':
        n = nh + 1
        # destinations rotate through 16 pinned buffers (50 MB per rank): like the real result stream, the copies land in host
        # DRAM, not in a last-level-cache-resident pair of buffers
        src = [torch.empty(B, n, 32, device=dev) for _ in range(2)]
        dst = [torch.empty(B, n, 32).pin_memory() for _ in range(16)]
        cs = [torch.cuda.Stream() for _ in range(2)]
        for rep in range(2):
            D.barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for s in cs:
                s.wait_event(c0)
            for i in range(200):
                with torch.cuda.stream(cs[i % 2]):
                    dst[i % 16].copy_(src[i % 2], non_blocking=True)
            for s in cs:
                torch.cuda.current_stream().wait_stream(s)
            c1.record()
            D.barrier()
            d2h_gbs = 200 * B * n * 128 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del src, dst
    sampler.stop_flag = True
    sampler.join()

    # secondary numbers in the same run: the policy-facing call ValueEstimator.forward (graph forward + value head -> V),
    # whose result is 16 KB per step instead of the 3.1 MB H tensor, device-resident and end to end
    extra_ms = [0.0, 0.0]
    if args.workload == 'graph':
        with torch.no_grad():
            for i in range(W):
                ve.run(robots_d[i % pool_n], humans_d[i % pool_n], throughput=thr)
        torch.cuda.synchronize()
        vrun = lambda r, h: ve.run(r, h, throughput=thr)     # noqa: E731
        Gv = min(K, 250)
        nv = int(min(8, max(1, math.ceil(pool_n / Gv))))
        vgraphs = [capture(vrun, j * Gv, Gv, True) for j in range(nv)]
        for g, _ in vgraphs:
            g.replay()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier()
        a0.record()
        vgraphs[0][0].replay()
        a1.record()
        D.barrier()
        vt, _, vR = timed_passes(D, lambda p: vgraphs[(p + 1) % nv][0].replay(), a0.elapsed_time(a1))
        extra_ms[0] = vt / (vR * Gv)
        del vgraphs
        hv = HostStream('value', ve, B, nh, dev, depth=depth)
        v_ms, _, _ = e2e_run(hv, K_e2e)
        extra_ms[1] = v_ms / K_e2e
    # the one HBM-bound unit of the path: a stand-alone GCN layer on features resident in HBM, A given.
    # 1 536 + 144 B per 6-node state; working set (1 M states = 1.7 GB) far beyond L2.
    gcn_ms = 0.0
    if args.workload == 'graph':
        Bg, n_ = 1 << 20, nh + 1
        Xg = torch.randn(Bg, n_, 32, device=dev)
        Wg = torch.randn(32, 32, device=dev)
        Ag = torch.softmax(torch.randn(Bg, n_, n_, device=dev), dim=2)
        with torch.no_grad():
            for _ in range(3):
                ops.gcn_layer(Xg, Wg, A=Ag, skip=True)
            D.barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(10):
                ops.gcn_layer(Xg, Wg, A=Ag, skip=True)
            c1.record()
        D.barrier()
        gcn_ms = c0.elapsed_time(c1) / 10
        del Xg, Ag
    # the same fused kernel in its steady state: one launch over 1 Mi states (950 MB of algorithmic traffic, >> L2)
    steady_ms = 0.0
    if args.workload == 'graph':
        rs, hs_ = synthetic_states(1 << 16, nh, seed=77 + rank, device=dev)
        rs, hs_ = rs.repeat(16, 1, 1), hs_.repeat(16, 1, 1)
        with torch.no_grad():
            for _ in range(2):
                keep4 = g1.run(rs, hs_, want_H=True)
            D.barrier()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record()
            for _ in range(5):
                keep4 = g1.run(rs, hs_, want_H=True)
            d1.record()
        D.barrier()
        steady_ms = d0.elapsed_time(d1) / 5
        del rs, hs_, keep4
    t_ms, e2e_ms, extra0, extra1, gcn_ms, steady_ms = D.max_([t_ms, e2e_ms, extra_ms[0], extra_ms[1], gcn_ms, steady_ms])
    extra_ms = [extra0, extra1]
    d2h_total = D.sum_([d2h_gbs])[0]
    del graphs, robots_d, humans_d
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, same run, every rank participates (C4 train step with the gradient all-reduce; C3 / C5 trees) ----
    extras = {}
    if args.workload == 'graph' and not args.no_extras:
        cpu_leg = world == 1 and not args.no_cpu_baseline
        extras['train_c4'] = measure_train(D, 8192, 10, 40, 5, backend=args.dp_comm, cpu_leg=cpu_leg)
        extras['plan_c3'] = measure_plan(D, 1024, 5, 20, 3, 2, 2, 2, 5, cpu_leg=cpu_leg, label='C3: ')
        extras['plan_c5'] = measure_plan(D, 2048, 20, 5, 2, 3, 2, 5, 16, cpu_leg=False, label='C5: 16 384 rollout roots over 8 GPUs = 2 048 per GPU; ')

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:  # noqa: BLE001
            pass
        hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
        per_launch_s = kernel_ms * 1e-3
        ach_gbs = abytes * B / per_launch_s / 1e9
        ach_tf = aflops * B / per_launch_s / 1e12
        traffic, traffic_note = None, 'no ncu capture for this workload / batch'
        tj = {}
        try:
            tj = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
            key = '%s_b%d_nh%d' % (args.workload, B, nh)
            if key in tj:
                traffic, traffic_note = tj[key]['dram_bytes'], tj[key]['note']
        except Exception:  # noqa: BLE001
            pass
        total_steps = K * R
        e2e_value = world * B * K_e2e / (e2e_ms * 1e-3)
        out = {
            'metric': metric, 'value': world * B * total_steps / (t_ms * 1e-3), 'unit': 'states/s', 'n_gpus': world, 'steps': K,
            'warmup': W, 'replays': R, 'ms_per_step': t_ms / total_steps, 'pass': pass_stats,
            'timing': 'R = %d passes of K = %d steps (>= %.0f ms timed), CUDA events, max over ranks; consecutive passes replay %d '
                      'different CUDA graphs over different input-pool slices' % (R, K, MIN_TIMED_MS, n_graphs),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': config,
            'e2e': {'value': e2e_value, 'unit': 'states/s', 'h2d_bytes_per_step': hs.h2d_bytes,
                    'd2h_bytes_per_step': hs.d2h_bytes, 'ms_per_step': e2e_ms / K_e2e, 'steps': K_e2e,
                    'api': 'hostio.HostStream.submit / result (pinned host -> H2D -> kernels -> D2H -> pinned host, %d batches in flight, '
                           'every result read by the host, CUDA-graph replay)' % depth},
            'gpu_launches': int(round(launches_per_step * total_steps)),
            'roofline': {'bound': 'fp32_fma+tensor', 'achieved': ach_tf, 'peak': FMA_PEAK_TFLOPS, 'unit': 'TFLOP/s', 'frac': ach_tf / FMA_PEAK_TFLOPS,
                         'peak_source': 'measured fp32-FMA peak of this pool (tools/fma_peak.cu); the shared-weight GEMMs (90 %% of the MACs) run '
                                        'on tcgen05 as 3xTF32, so the fraction can exceed 1 (measured bf16 tensor peak: %s TFLOP/s)' % peaks.get('bf16_tflops'),
                         'traffic': traffic, 'traffic_note': traffic_note,
                         'kernel': {'graph': 'graph_forward_tp_kernel<6,2>', 'value': 'graph_forward_tc_kernel (+ value_head_tc_kernel)',
                                    'statepred': 'graph_forward_tp_kernel'}[args.workload] if nh == 5 else 'graph_forward_tp/tc_kernel',
                         'launch_us': kernel_ms * 1e3,
                         'launch_timing': 'CUDA events around %d back-to-back launches on one stream (graph replay over the whole > L2 input pool)' % (5 * KG),
                         'algorithmic_bytes_per_state': abytes, 'algorithmic_flops_per_state': aflops,
                         'hbm': {'achieved': ach_gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach_gbs / hbm_peak,
                                 'peak_source': 'measured (MEASURED_PEAKS.json hbm_gbs)' if peaks else 'fallback',
                                 'note': '%d FLOP/B: not an HBM-bound unit; the HBM-bound unit of the path is extra.gcn_layer' % (aflops // abytes)}},
            'clocks': sampler.summary(),
        }
        if d2h_total > 0:
            d2h_used = e2e_value * (nh + 1) * 128 / 1e9
            out['e2e']['d2h_ceiling'] = {'aggregate_gbs': d2h_total, 'used_gbs': d2h_used, 'frac': d2h_used / d2h_total,
                                         'how': 'every rank copies 3.1 MB device->pinned-host buffers back to back at the same time '
                                                '(2 streams, 200 copies, 16 rotating destinations = 50 MB per rank): the host fabric bound of the H read-back'}
        if extra_ms[0] > 0:
            out['extra'] = {'value_path': {'call': 'ValueEstimator.forward = graph forward (E only) + value head -> V[B,1]',
                                           'value': world * B / (extra_ms[0] * 1e-3), 'e2e': world * B / (extra_ms[1] * 1e-3),
                                           'unit': 'states/s', 'd2h_bytes_per_step': B * 4}}
        if gcn_ms > 0:
            n_ = nh + 1
            gb = (2 * 128 * n_ + 4 * n_ * n_) * (1 << 20) / (gcn_ms * 1e-3) / 1e9
            out.setdefault('extra', {})['gcn_layer'] = {
                'call': 'rgl_gcn_layer: H\' = relu(A (X W)) + X on features in HBM, A given, B = 1 Mi states per GPU (gcn_layer_tma_kernel)',
                'value': world * (1 << 20) / (gcn_ms * 1e-3), 'unit': 'layer-states/s', 'launch_us': gcn_ms * 1e3,
                'roofline': {'bound': 'hbm', 'achieved': gb, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gb / hbm_peak,
                             'algorithmic_bytes_per_state': 2 * 128 * n_ + 4 * n_ * n_,
                             'traffic': (tj.get('gcn_layer_b1048576_n%d' % n_) or {}).get('dram_bytes'),
                             'note': 'the only unit of the path that sits at the HBM / FMA ridge (8.7 FLOP/B)'}}
        if steady_ms > 0:
            gb = abytes * (1 << 20) / (steady_ms * 1e-3) / 1e9
            tf = aflops * (1 << 20) / (steady_ms * 1e-3) / 1e12
            out.setdefault('extra', {})['steady_state'] = {
                'call': 'the same fused kernel, one launch over B = 1 Mi states per GPU (launch latency amortised)',
                'value': world * (1 << 20) / (steady_ms * 1e-3), 'unit': 'states/s', 'launch_us': steady_ms * 1e3,
                'roofline': {'bound': 'fp32_fma+tensor', 'achieved': tf, 'peak': FMA_PEAK_TFLOPS, 'unit': 'TFLOP/s', 'frac': tf / FMA_PEAK_TFLOPS,
                             'hbm': {'achieved': gb, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gb / hbm_peak}}}
        for k, v in extras.items():
            if v is not None:
                out.setdefault('extra', {})[k] = v
        if world == 1 and not args.no_cpu_baseline:
            c = cpu_reference_rate(args.workload, B, nh, min(K, 2000), 3, budget_s=15.0)
            out['cpu_baseline'] = {k: c[k] for k in ('value', 'unit', 'cores', 'cores_available', 'kind', 'sample', 'ms_per_step')}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
