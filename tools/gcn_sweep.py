"""Stand-alone GCN layer throughput (A given / w_a), B = 1 Mi states; run with RGL_TC_GROUPS / RGL_GCN_VARIANT."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relationalgraphlearning_b200 import ops
dev = torch.device('cuda:0')
out = []
for n in (6, 11, 21):
    B = (1 << 20) * 6 // n
    X = torch.randn(B, n, 32, device=dev); W = torch.randn(32, 32, device=dev); wa = torch.randn(32, 32, device=dev) * 0.1
    A = torch.softmax(torch.randn(B, n, n, device=dev), dim=2)
    for name, fn in (('A', lambda: ops.gcn_layer(X, W, A=A, skip=True)), ('wa', lambda: ops.gcn_layer(X, W, w_a=wa, skip=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        byt = B * (2 * 128 * n + (4 * n * n if name == 'A' else 0))
        out.append('n=%d %s %.0fus %.0f Mst/s %.2f TB/s' % (n, name, us, B / us, byt / us / 1e6))
print(os.environ.get('RGL_TC_GROUPS', 'auto'), ' | '.join(out))
