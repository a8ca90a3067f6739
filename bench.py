#!/usr/bin/env python
"""bench.py -- RGL graph-forward states/sec (batch 4096, 5 humans) on N B200s, one rank per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload graph|value|statepred]

A step = one pass of the hot path over one batch of `--batch` synthetic states (SURVEY.md 8(d) distribution).
  value      device-resident throughput: inputs already in HBM, K steps replayed from CUDA graphs, timed with
             CUDA events on the launching stream, max over ranks.
  e2e        same metric through the host-buffer API (relationalgraphlearning_b200.hostio.HostStream):
             every step copies its states from pinned host memory and reads the result back.
  roofline   HBM fraction of the dominant kernel on algorithmic bytes (+ fp32-FMA fraction, the binding one).
  cpu_baseline / --impl reference: the CPU oracle port of the reference path (oracle/rgl_oracle.py: same ATen
             ops as crowd_nav/policy/graph_model.py) on the host cores.
Inputs rotate through a pool larger than L2 (126 MB) so no step re-reads L2-resident states.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

L2_BYTES = 126 * 1024 * 1024
FMA_PEAK_TFLOPS = 72.6         # measured on this pool's B200 with tools/fma_peak.cu (36.3 TFMA/s, 124.8 FMA/clk/SM)


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


def cpu_reference_rate(workload, batch, nh, steps, warmup, budget_s=150.0):
    """states/s of the oracle port (the reference's ATen op sequence) on the host cores."""
    from oracle import rgl_oracle as O
    from relationalgraphlearning_b200.synthetic import synthetic_states
    # The oracle allocates ~20 MB of temporaries per step; with glibc's default thresholds every one of them is mmap'ed and
    # page-faulted afresh (measured 2.3-2.6x slower).  Give the CPU leg its best case: keep freed memory in the heap.
    try:
        import ctypes
        libc = ctypes.CDLL('libc.so.6')
        libc.mallopt(-3, 1 << 30)       # M_MMAP_THRESHOLD
        libc.mallopt(-1, 1 << 30)       # M_TRIM_THRESHOLD
    except Exception:  # noqa: BLE001
        pass
    g1, ve, g2, sp = build_modules(0)
    sd = [{k: v.detach().clone() for k, v in m.state_dict().items()} for m in (g1, ve.value_network, g2, sp.human_motion_predictor)]
    pool = [synthetic_states(batch, nh, seed=100 + i) for i in range(8)]
    avail = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)

    def step(i):
        r, h = pool[i % len(pool)]
        with torch.no_grad():
            if workload == 'graph':
                return O.rgl_forward(sd[0], r, h)
            if workload == 'value':
                return O.value_forward(sd[0], sd[1], r, h)
            return O.statepred_forward(sd[2], sd[3], r, h)

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
    return batch * k / dt, cores, sample, dt / k * 1e3, k


def train_bench(args, rank, world, local):
    """BASELINE configs[3]: value-net training step (forward + target forward + MSE + backward + Adam), batch per GPU
    `--batch`, one flat gradient all-reduce per step when world > 1.  Extra workload; not the headline metric."""
    import copy
    import torch.distributed as dist
    from relationalgraphlearning_b200 import ops, parallel
    from relationalgraphlearning_b200.synthetic import synthetic_states
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B, nh, K, W = args.batch, args.humans, args.steps, max(args.warmup, 3)
    g1, ve, _, _ = build_modules(0)
    ve.to(dev)
    target = copy.deepcopy(ve)
    opt = torch.optim.Adam(ve.parameters(), lr=1e-3, fused=True, capturable=True)
    red = parallel.FlatGradAllReducer(ve.parameters())
    pool = []
    for i in range(8):
        r, h = synthetic_states(B, nh, seed=10 + i + 100 * rank, device=dev)
        r2, h2 = synthetic_states(B, nh, seed=50 + i + 100 * rank, device=dev)
        pool.append((r, h, torch.rand(B, 1, device=dev) * 1.25 - 0.25, r2, h2))
    gamma_bar = pow(0.9, 0.25)

    def step(i):
        r, h, rew, r2, h2 = pool[i % len(pool)]
        if world > 1:
            return parallel.dp_value_step(ve, target, opt, red, r, h, rew, r2, h2, gamma_bar, B * world)
        opt.zero_grad()
        out = ve((r, h))
        with torch.no_grad():
            tgt = rew + gamma_bar * target((r2, h2))
        loss = torch.nn.functional.mse_loss(out, tgt)
        loss.backward()
        opt.step()
        return loss.detach()

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(W):
            step(i)
    torch.cuda.synchronize()
    # single GPU: the whole step (pack, fused forward, target forward, loss, native backward, fused Adam) is captured into a
    # CUDA graph of G steps and replayed (no Python / launch overhead in the timed region).  With world > 1 the steps run
    # eagerly: the NCCL all-reduce is issued from Python between backward and the optimizer step.
    use_graph = world == 1
    G = max(d for d in range(1, 9) if K % d == 0) if use_graph else 1
    l0 = ops.LAUNCHES
    if use_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for i in range(G):
                    loss = step(W + i)
        launches_per_graph = ops.LAUNCHES - l0
        graph.replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if use_graph:
        for _ in range(K // G):
            graph.replay()
    else:
        for i in range(K):
            loss = step(W + i)
        launches_per_graph = ops.LAUNCHES - l0
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t[0])
        out = {'metric': 'RGL value-net training samples/sec (batch %d per GPU, %d humans)' % (B, nh),
               'value': world * B * K / (ms * 1e-3), 'unit': 'samples/s', 'n_gpus': world, 'steps': K, 'warmup': W,
               'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
               'data': 'synthetic', 'gpu_launches': launches_per_graph * (K // G) if use_graph else launches_per_graph, 'final_loss': float(loss),
               'graph_captured': use_graph,
               'config': {'workload': 'value-net train step B=%d Nh=%d (BASELINE configs[3]): fused fwd+saves, native bwd, Adam' % (B, nh),
                          'grad_allreduce_bytes': red.numel * 4 if world > 1 else 0,
                          'parallelism': 'dp%d, one flat all-reduce per step' % world}}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import rgl_oracle as O
            sd = [{k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()} for m in (g1, ve.value_network)]
            sdt = [{k: v.detach().cpu().clone() for k, v in m.state_dict().items()} for m in (g1, ve.value_network)]
            copt = torch.optim.Adam(list(sd[0].values()) + list(sd[1].values()), lr=1e-3)
            cpool = [tuple(t.cpu() for t in p) for p in pool[:2]]
            torch.set_num_threads(min(16, len(os.sched_getaffinity(0))))     # small autograd ops do not scale past ~16 threads

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
            while time.perf_counter() - t0 < 10.0:
                cstep(n)
                n += 1
            dt = time.perf_counter() - t0
            out['cpu_baseline'] = {'value': B * n / dt, 'unit': 'samples/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                                   'sample': '%d full train steps (~10 s)' % n, 'ms_per_step': dt / n * 1e3}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def plan_bench(args, rank, world, local):
    """BASELINE configs[2] / [4]: d-step, w-width look-ahead of `--roots` root states per GPU (sharded over ranks, no
    collective).  Reports root states planned per second and rollout states (value / state-predictor evaluations) per
    second; CPU leg = the batch-1 oracle tree (oracle/planner_oracle.py) on a bounded number of roots."""
    import torch.distributed as dist
    from relationalgraphlearning_b200 import ops
    from relationalgraphlearning_b200.config import policy_config
    from relationalgraphlearning_b200.model_predictive_rl import ModelPredictiveRL
    from relationalgraphlearning_b200.synthetic import synthetic_states
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    E, nh, K, W = args.roots, args.humans, args.steps, max(args.warmup, 3)
    cfg = policy_config(planning_depth=args.depth, planning_width=args.width, do_action_clip=args.depth > 1 or args.width > 1,
                        speed_samples=args.speed_samples, rotation_samples=args.rotation_samples)
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
    # the policy's default); --plan-eager issues the launches one by one from Python instead.
    graphed = not args.plan_eager
    run = (lambda r, h: pol.predict_batch_graphed(r, h)[0]) if graphed else pol.predict_batch
    pol.stat_value_states = pol.stat_sp_states = 0
    l0 = ops.LAUNCHES
    pol.predict_batch(*pool[0])                      # one eager pass: launches and rollout states of one step
    launches_per_step = ops.LAUNCHES - l0
    rollout_per_step = pol.stat_value_states + pol.stat_sp_states
    for i in range(W):
        run(*pool[i % 4])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        best = run(*pool[i % 4])
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t[0])
        rollout = rollout_per_step
        out = {'metric': 'model_predictive_rl look-ahead: root states planned/sec (d=%d, w=%d, %d actions, %d humans)' %
                         (args.depth, args.width, len(pol.action_space), nh),
               'value': world * E * K / (ms * 1e-3), 'unit': 'root states/s', 'n_gpus': world, 'steps': K, 'warmup': W,
               'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
               'data': 'synthetic', 'gpu_launches': launches_per_step * K, 'graph_replay': graphed,
               'rollout_states_per_step': rollout, 'rollout_states_per_s': world * rollout * K / (ms * 1e-3),
               'config': {'workload': 'planner tree d=%d w=%d, %d root states per GPU' % (args.depth, args.width, E),
                          'parallelism': 'dp%d (root states sharded, no collective)' % world}}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import planner_oracle as P
            sd = pol.get_state_dict()
            cpu = lambda d: {k: v.detach().cpu() for k, v in d.items()}   # noqa: E731
            orc = P.OraclePlanner(cpu(sd['graph_model1']), cpu(sd['value_network']), cpu(sd['graph_model2']), cpu(sd['motion_predictor']),
                                  planning_depth=args.depth, planning_width=args.width, do_action_clip=cfg.model_predictive_rl.do_action_clip,
                                  speed_samples=args.speed_samples, rotation_samples=args.rotation_samples)
            torch.set_num_threads(1)          # batch-1 forwards: more threads only add overhead
            r, h = pool[0][0].cpu(), pool[0][1].cpu()
            t0 = time.perf_counter()
            nroots = 0
            while time.perf_counter() - t0 < 10.0 and nroots < E:
                orc.predict(r[nroots:nroots + 1], h[nroots:nroots + 1])
                nroots += 1
            dt = time.perf_counter() - t0
            out['cpu_baseline'] = {'value': nroots / dt, 'unit': 'root states/s', 'cores': 1, 'kind': 'port',
                                   'sample': '%d root states through the batch-1 oracle tree (~10 s)' % nroots}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
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
    ap.add_argument('--zero-copy', action='store_true', help='e2e: the kernel writes H into pinned host memory itself')
    ap.add_argument('--host-depth', type=int, default=4, help='e2e: batches in flight through hostio.HostStream')
    ap.add_argument('--streams', type=int, default=4, help='CUDA streams the independent steps are issued on')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.workload == 'plan':
        if args.impl == 'reference':
            print(json.dumps({'impl': 'reference', 'unavailable': 'plan workload: the CPU leg is reported inside the ours arm (cpu_baseline)'}))
            return
        return plan_bench(args, rank, world, local)
    if args.workload == 'train':
        if args.impl == 'reference':
            print(json.dumps({'impl': 'reference', 'unavailable': 'train workload: the CPU leg is reported inside the ours arm (cpu_baseline)'}))
            return
        return train_bench(args, rank, world, local)
    B, nh, K, W = args.batch, args.humans, args.steps, max(args.warmup, 3)
    abytes, aflops = algorithmic(args.workload, nh)
    metric = 'RGL graph-forward states/sec (batch %d, %d humans)' % (B, nh)
    config = {'workload': 'rgl_%s_forward B=%d Nh=%d 2-layer GCN fp32 (BASELINE configs[1])' % (args.workload, B, nh),
              'batch_per_gpu': B, 'humans': nh, 'parallelism': 'dp%d (batch sharded, no collective)' % world,
              'l2_policy': 'inputs rotate through a pool > 126 MB L2',
              'streams': args.streams}

    if args.impl == 'reference':
        if rank != 0:
            return
        rate, cores, sample, ms, k = cpu_reference_rate(args.workload, B, nh, K, W)
        print(json.dumps({'impl': 'reference', 'metric': metric, 'value': rate, 'unit': 'states/s', 'n_gpus': args.gpus,
                          'steps': k, 'warmup': W, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
                          'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
                          'cpu_baseline': {'value': rate, 'unit': 'states/s', 'cores': cores, 'kind': 'port', 'sample': sample},
                          'e2e': {'value': rate, 'unit': 'states/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                          'gpu_launches': 0}))
        return

    import torch.distributed as dist
    from relationalgraphlearning_b200 import ops
    from relationalgraphlearning_b200.hostio import HostStream
    from relationalgraphlearning_b200.synthetic import synthetic_states

    assert torch.cuda.is_available(), 'bench.py (impl ours) needs a CUDA device'
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    g1, ve, g2, sp = build_modules(0)
    ve.to(dev)
    sp.to(dev)
    module = {'graph': g1, 'value': ve, 'statepred': sp}[args.workload]

    def run_step(robot, humans):
        if args.workload == 'graph':
            return g1.run(robot, humans, want_H=True, throughput=args.streams > 1)['H']
        return module.run(robot, humans, throughput=args.streams > 1)

    # ---- input pool larger than L2 (distinct states per rank) ----
    pool_n = max(8, (int(1.15 * L2_BYTES) + B * (36 + 20 * nh) - 1) // (B * (36 + 20 * nh)))
    rb, hb = synthetic_states(pool_n * B, nh, seed=1234 + rank)
    robots = [rb[i * B:(i + 1) * B].contiguous() for i in range(pool_n)]
    humans = [hb[i * B:(i + 1) * B].contiguous() for i in range(pool_n)]
    robots_d = [r.to(dev) for r in robots]
    humans_d = [h.to(dev) for h in humans]
    config['input_pool_mb'] = round(pool_n * B * (36 + 20 * nh) / 1e6, 1)

    # ---- warm-up (eager) + CUDA-graph capture of the steps ----
    with torch.no_grad():
        for i in range(W):
            run_step(robots_d[i % pool_n], humans_d[i % pool_n])
    torch.cuda.synchronize()
    G = K
    if K > 500:
        G = max(d for d in range(1, 501) if K % d == 0)
    reps = K // G
    l0 = ops.LAUNCHES
    # Steps are independent batches, so consecutive steps are issued round-robin on `--streams` CUDA streams (forked
    # from / joined to the capture stream): the tail of step i overlaps the head of step i+1 on the GPU.
    nstreams = max(1, args.streams)
    side = torch.cuda.Stream()
    branches = [torch.cuda.Stream() for _ in range(nstreams)]
    graph = torch.cuda.CUDAGraph()
    keep = []
    with torch.no_grad(), torch.cuda.stream(side):
        run_step(robots_d[0], humans_d[0])
        torch.cuda.synchronize()
        l0 = ops.LAUNCHES
        with torch.cuda.graph(graph, stream=side):
            fork = torch.cuda.Event()
            fork.record(side)
            for b in branches:
                b.wait_event(fork)
            for i in range(G):
                with torch.cuda.stream(branches[i % nstreams]):
                    out = run_step(robots_d[(W + i) % pool_n], humans_d[(W + i) % pool_n])
                    keep.append(out)
                if len(keep) >= 64:     # rotate output buffers: a 64-deep ring (> L2 for the H output) instead of K live tensors
                    keep = keep[32:]
            for b in branches:
                ev = torch.cuda.Event()
                ev.record(b)
                side.wait_event(ev)
    launches_per_step = (ops.LAUNCHES - l0) / G
    graph.replay()                      # untimed: uploads the graph, K more warm steps
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    barrier()
    t_ms = e0.elapsed_time(e1)

    # ---- dominant kernel alone: back-to-back launches on ONE stream (no overlap between launches), CUDA events on that
    # stream -> average launch duration for the roofline (a multi-stream step time would understate it) ----
    kgraph = torch.cuda.CUDAGraph()
    KG = min(G, 200)
    keep3 = []
    with torch.no_grad(), torch.cuda.stream(side):
        with torch.cuda.graph(kgraph, stream=side):
            for i in range(KG):
                keep3.append(run_step(robots_d[(W + i) % pool_n], humans_d[(W + i) % pool_n]))
                if len(keep3) >= 64:
                    keep3 = keep3[32:]
        kgraph.replay()
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(side)
        for _ in range(5):
            kgraph.replay()
        k1.record(side)
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / (5 * KG)          # per step on one stream (graph workload: exactly one kernel per step)

    # ---- end to end through the host-buffer API: pinned host -> device -> kernels -> pinned host ----
    npin = min(pool_n, 48)
    robots_p = [r.pin_memory() for r in robots[:npin]]
    humans_p = [h.pin_memory() for h in humans[:npin]]
    depth = args.host_depth
    hs = HostStream(args.workload, module, B, nh, dev, depth=depth, zero_copy_out=args.zero_copy)
    for i in range(max(W, npin * depth)):            # warm-up also captures the per-(slot, buffer) graphs
        hs.submit(robots_p[i % npin], humans_p[i % npin])
    hs.drain()
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    s0.record(hs.streams[0])
    checksum = 0.0
    for i in range(K):
        slot = hs.submit(robots_p[i % npin], humans_p[i % npin])
        if i % 256 == 255:
            checksum += float(hs.result(slot).view(-1)[0])      # the host consumes results while the stream runs
    hs.drain()
    s1.record(hs.streams[0])
    barrier()
    e2e_ms = s0.elapsed_time(s1)                # device timestamps: before the first copy .. after the last result landed
    e2e_host_ms = (time.perf_counter() - t_host0) * 1e3
    sampler.stop_flag = True
    sampler.join()

    # secondary numbers in the same run: the policy-facing call ValueEstimator.forward (graph forward + value head -> V),
    # whose result is 16 KB per step instead of the 3.1 MB H tensor, device-resident and end to end
    extra_ms = [0.0, 0.0]
    if args.workload == 'graph':
        K2 = min(K, 1000)
        with torch.no_grad():
            for i in range(W):
                ve.run(robots_d[i % pool_n], humans_d[i % pool_n], throughput=args.streams > 1)
        g2_ = torch.cuda.CUDAGraph()
        keep2 = []
        with torch.no_grad(), torch.cuda.stream(side):
            with torch.cuda.graph(g2_, stream=side):
                fork = torch.cuda.Event()
                fork.record(side)
                for b in branches:
                    b.wait_event(fork)
                for i in range(min(K2, 250)):
                    with torch.cuda.stream(branches[i % nstreams]):
                        keep2.append(ve.run(robots_d[i % pool_n], humans_d[i % pool_n], throughput=args.streams > 1))
                for b in branches:
                    ev = torch.cuda.Event()
                    ev.record(b)
                    side.wait_event(ev)
        g2_.replay()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(K2 // min(K2, 250)):
            g2_.replay()
        a1.record()
        barrier()
        K2 = (K2 // min(K2, 250)) * min(K2, 250)
        extra_ms[0] = a0.elapsed_time(a1) / K2
        hv = HostStream('value', ve, B, nh, dev, depth=depth)
        for i in range(npin * depth):
            hv.submit(robots_p[i % npin], humans_p[i % npin])
        hv.drain()
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(hv.streams[0])
        for i in range(K2):
            hv.submit(robots_p[i % npin], humans_p[i % npin])
        hv.drain()
        b1.record(hv.streams[0])
        barrier()
        extra_ms[1] = b0.elapsed_time(b1) / K2
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
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(10):
                ops.gcn_layer(Xg, Wg, A=Ag, skip=True)
            c1.record()
        barrier()
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
            barrier()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record()
            for _ in range(5):
                keep4 = g1.run(rs, hs_, want_H=True)
            d1.record()
        barrier()
        steady_ms = d0.elapsed_time(d1) / 5
        del rs, hs_, keep4
    times = torch.tensor([t_ms, e2e_ms] + extra_ms + [gcn_ms, steady_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_ms, e2e_ms = float(times[0]), float(times[1])
    extra_ms = [float(times[2]), float(times[3])]
    gcn_ms = float(times[4])
    steady_ms = float(times[5])

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
        out = {
            'metric': metric, 'value': world * B * K / (t_ms * 1e-3), 'unit': 'states/s', 'n_gpus': world, 'steps': K,
            'warmup': W, 'ms_per_step': t_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': config,
            'e2e': {'value': world * B * K / (e2e_ms * 1e-3), 'unit': 'states/s', 'h2d_bytes_per_step': hs.h2d_bytes,
                    'd2h_bytes_per_step': hs.d2h_bytes, 'ms_per_step': e2e_ms / K,
                    'api': 'hostio.HostStream.submit (pinned host -> H2D -> kernels -> D2H -> pinned host, %d streams, CUDA-graph replay)' % depth},
            'gpu_launches': int(round(launches_per_step * K)),
            'roofline': {'bound': 'hbm', 'achieved': ach_gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach_gbs / hbm_peak,
                         'traffic': traffic, 'traffic_note': traffic_note,
                         'peak_source': 'measured (MEASURED_PEAKS.json hbm_gbs)' if peaks else 'fallback',
                         'kernel': 'graph_forward_tc_kernel' if args.workload == 'graph' else 'graph_forward_tc_kernel (+ value_head_kernel)',
                         'launch_us': kernel_ms * 1e3,
                         'launch_timing': 'CUDA events around %d back-to-back launches on one stream (graph replay)' % (5 * KG),
                         'algorithmic_bytes_per_state': abytes, 'algorithmic_flops_per_state': aflops,
                         'compute': {'achieved_tflops': ach_tf,
                                     'note': 'algorithmic fp32 FLOP/s (%d FLOP/B: not an HBM-bound unit); the shared-weight GEMMs run on '
                                             'tcgen05 as 3xTF32 (3 tensor MACs per algorithmic MAC), the per-state similarity / softmax / '
                                             'A.H work on the fp32 FMA pipe (measured FMA peak %.1f TFLOP/s)' % (aflops // abytes, FMA_PEAK_TFLOPS)}},
            'clocks': sampler.summary(),
        }
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
            out.setdefault('extra', {})['steady_state'] = {
                'call': 'the same fused kernel, one launch over B = 1 Mi states per GPU (launch latency amortised)',
                'value': world * (1 << 20) / (steady_ms * 1e-3), 'unit': 'states/s', 'launch_us': steady_ms * 1e3,
                'roofline': {'bound': 'hbm', 'achieved': gb, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gb / hbm_peak,
                             'note': 'latency / LSU-bound, not HBM-bound (80 FLOP/B): DESIGN.md section 3'}}
        if world == 1 and not args.no_cpu_baseline:
            rate, cores, sample, ms, k = cpu_reference_rate(args.workload, B, nh, min(K, 2000), 3, budget_s=15.0)
            out['cpu_baseline'] = {'value': rate, 'unit': 'states/s', 'cores': cores, 'kind': 'port', 'sample': sample,
                                   'ms_per_step': ms}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
