"""Quick device timing of the forward kernels (CUDA events, back-to-back launches, rotating input pool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relationalgraphlearning_b200 import ops
from relationalgraphlearning_b200.config import policy_config
from relationalgraphlearning_b200.graph_model import RGL
from relationalgraphlearning_b200.state_predictor import StatePredictor
from relationalgraphlearning_b200.synthetic import synthetic_states
from relationalgraphlearning_b200.value_estimator import ValueEstimator

dev = torch.device('cuda:0')
cfg = policy_config()
torch.manual_seed(0)
g1 = RGL(cfg, 9, 5); ve = ValueEstimator(cfg, g1); g2 = RGL(cfg, 9, 5); sp = StatePredictor(cfg, g2, 0.25)
ve.to(dev); sp.to(dev)

def timeit(fn, iters=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us

def graphed(fn, iters=200):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(5): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

import os
CASES = ((5, 4096), (5, 65536), (5, 1048576)) if os.environ.get('QT_FAST') else ((5, 4096), (5, 65536), (5, 1048576), (10, 8192), (20, 16384))
for nh, B in CASES:
    robot, humans = synthetic_states(min(B, 65536), nh, seed=1, device=dev)
    if B > 65536:
        robot = robot.repeat(B // 65536, 1, 1); humans = humans.repeat(B // 65536, 1, 1)
    gb, vb = ops.packed_graph(g1), ops.packed_value(ve.value_network, ve._pack_cache)
    mb = ops.packed_motion(sp.human_motion_predictor, sp._pack_cache)
    with torch.no_grad():
        E = g1.run(robot, humans, want_E=True)['E']
        it = 200 if B <= 65536 else 20
        res = {}
        res['graph_H'] = graphed(lambda: g1.run(robot, humans, want_H=True), it)
        res['graph_E'] = graphed(lambda: g1.run(robot, humans, want_E=True), it)
        res['graph_S'] = graphed(lambda: sp.run(robot, humans), it)
        res['vhead'] = graphed(lambda: ops.value_head_raw(vb, E), it)
        res['value'] = graphed(lambda: ve.run(robot, humans), it)
        res['graph_H_eager'] = timeit(lambda: g1.run(robot, humans, want_H=True), it)
        X = torch.randn(B, nh + 1, 32, device=dev); W = torch.randn(32, 32, device=dev); wa = torch.randn(32, 32, device=dev) * 0.1
        A = torch.softmax(torch.randn(B, nh + 1, nh + 1, device=dev), dim=2)
        res['gcn_A'] = graphed(lambda: ops.gcn_layer(X, W, A=A, skip=True), it)
        res['gcn_wa'] = graphed(lambda: ops.gcn_layer(X, W, w_a=wa, skip=True), it)
    print('Nh=%d B=%d: ' % (nh, B) + '  '.join('%s %.1fus (%.0f Mst/s)' % (k, v, B / v) for k, v in res.items()), flush=True)
