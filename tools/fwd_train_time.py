"""Warm timing of the training forward with activation saves (fp32-FMA kernel vs tcgen05 kernel) at the C4 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relationalgraphlearning_b200 import training as T
from relationalgraphlearning_b200.config import policy_config
from relationalgraphlearning_b200.graph_model import RGL
from relationalgraphlearning_b200.synthetic import synthetic_states
dev = torch.device('cuda:0')
B, nh = int(os.environ.get('B', 8192)), int(os.environ.get('NH', 10))
torch.manual_seed(0)
g = RGL(policy_config(), 9, 5).to(dev)
robot, humans = synthetic_states(B, nh, seed=1, device=dev)
for variant in ('f', 't'):
    os.environ['RGL_TRAIN_VARIANT'] = variant
    def run():
        return T._graph_forward_train(g, robot, humans, [], want_E=True)
    for _ in range(5): run()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(20): keep = run()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    n = nh + 1
    save_mb = B * n * (64 + 32 * 2 + n + 32 * 3 * 2) * 4 / 1e6
    us = e0.elapsed_time(e1) / 20 * 1e3
    print('variant %s: %.1f us per forward (%.0f MB of saves -> %.2f TB/s)' % (variant, us, save_mb, save_mb / us / 1e3 * 1e3 / 1e3))
