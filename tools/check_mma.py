"""Accuracy + timing of the tensor-core (mma.sync 3xTF32) graph kernel variant vs the FFMA variant and the oracle.
Run as:  RGL_GRAPH_VARIANT=m python tools/check_mma.py   (and =4 for the FFMA baseline)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import rgl_oracle as O
from relationalgraphlearning_b200.config import policy_config
from relationalgraphlearning_b200.graph_model import RGL
from relationalgraphlearning_b200.state_predictor import StatePredictor
from relationalgraphlearning_b200.synthetic import synthetic_states
from relationalgraphlearning_b200.value_estimator import ValueEstimator

dev = torch.device('cuda:0')
print('variant', os.environ.get('RGL_GRAPH_VARIANT'))
worst = 0.0
for seed, scale in ((0, None), (1, 0.2), (2, 4.0)):
    cfg = policy_config()
    torch.manual_seed(seed)
    g1 = RGL(cfg, 9, 5); ve = ValueEstimator(cfg, g1); g2 = RGL(cfg, 9, 5); sp = StatePredictor(cfg, g2, 0.25)
    if scale is not None:
        with torch.no_grad():
            for g in (g1, g2):
                g.w_a.mul_(scale)
                for w in g.Ws: w.mul_(scale)
    sds = [{k: v.clone() for k, v in m.state_dict().items()} for m in (g1, ve.value_network, g2, sp.human_motion_predictor)]
    ve.to(dev); sp.to(dev)
    for B in (4096, 2400 + 7):
        robot, humans = synthetic_states(B, 5, seed=B + seed)
        with torch.no_grad():
            H = g1((robot.to(dev), humans.to(dev))).cpu(); V = ve((robot.to(dev), humans.to(dev))).cpu()
            S = sp((robot.to(dev), humans.to(dev)), None)[1].cpu()
            idx = torch.arange(0, B, 7)
            r, h = robot[idx], humans[idx]
            Ho = O.rgl_forward(sds[0], r, h); Vo = O.value_forward(sds[0], sds[1], r, h); So = O.statepred_forward(sds[2], sds[3], r, h)
            H64 = O.rgl_forward(O.to_double(sds[0]), r.double(), h.double())
        for nm, got, ref in (('H', H[idx], Ho), ('V', V[idx], Vo), ('S', S[idx], So)):
            sc = float(ref.abs().max()); err = float((got - ref).abs().max())
            worst = max(worst, err / sc)
            print('seed %d scale %s B %d %s: max|err| %.3e  scale %.3e  rel %.2e' % (seed, scale, B, nm, err, sc, err / sc))
        print('   vs fp64: ours %.3e  ref-fp32 %.3e' % (float((H[idx].double() - H64).abs().max()), float((Ho.double() - H64).abs().max())))
print('WORST rel', worst, 'PASS' if worst <= 1e-5 else 'FAIL')

cfg = policy_config(); torch.manual_seed(0)
g1 = RGL(cfg, 9, 5).to(dev)
def graphed(fn, iters=100):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for B in (4096, 65536, 1048576):
    robot, humans = synthetic_states(min(B, 65536), 5, seed=1, device=dev)
    if B > 65536: robot = robot.repeat(B // 65536, 1, 1); humans = humans.repeat(B // 65536, 1, 1)
    with torch.no_grad():
        t = graphed(lambda: g1.run(robot, humans, want_H=True), 100 if B <= 65536 else 10)
        t2 = graphed(lambda: g1.run(robot, humans, want_E=True), 100 if B <= 65536 else 10)
    print('B=%d graph_H %.1f us (%.0f M/s)  graph_E %.1f us (%.0f M/s)' % (B, t, B / t, t2, B / t2))
