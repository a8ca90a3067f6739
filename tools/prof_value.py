"""Run the tcgen05 value head a few times at a given batch (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relationalgraphlearning_b200 import ops
from relationalgraphlearning_b200.config import policy_config
from relationalgraphlearning_b200.graph_model import RGL
from relationalgraphlearning_b200.value_estimator import ValueEstimator
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device('cuda:0')
torch.manual_seed(0)
ve = ValueEstimator(policy_config(), RGL(policy_config(), 9, 5)).to(dev)
vb = ops.packed_value(ve.value_network, ve._pack_cache)
E = torch.randn(B, 32, device=dev)
with torch.no_grad():
    for _ in range(6):
        ops.value_head_raw(vb, E)
torch.cuda.synchronize()
