"""Value-head latency / throughput sweep (CUDA-graph replay of back-to-back launches); run with RGL_VALUE_VARIANT=t|f."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relationalgraphlearning_b200 import ops
from relationalgraphlearning_b200.config import policy_config
from relationalgraphlearning_b200.graph_model import RGL
from relationalgraphlearning_b200.value_estimator import ValueEstimator
dev = torch.device('cuda:0')
torch.manual_seed(0)
ve = ValueEstimator(policy_config(), RGL(policy_config(), 9, 5)).to(dev)
vb = ops.packed_value(ve.value_network, ve._pack_cache)
res = []
for B in (512, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
    E = torch.randn(B, 32, device=dev)
    with torch.no_grad():
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(5): ops.value_head_raw(vb, E)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(200): ops.value_head_raw(vb, E)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        res.append('B=%d %.2fus' % (B, e0.elapsed_time(e1) / 200 * 1e3))
print(os.environ.get('RGL_VALUE_VARIANT', 'auto'), os.environ.get('RGL_TC_VALUE_GROUPS', ''), '  '.join(res))
