"""Run the stand-alone GCN layer a few times (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relationalgraphlearning_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device('cuda:0')
X = torch.randn(B, n, 32, device=dev); W = torch.randn(32, 32, device=dev)
A = torch.softmax(torch.randn(B, n, n, device=dev), dim=2)
for _ in range(5):
    ops.gcn_layer(X, W, A=A, skip=True)
torch.cuda.synchronize()
