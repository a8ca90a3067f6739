"""Warm timing (CUDA events, back-to-back launches) of rgl_linear_bwd for the shapes of the C4 training step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relationalgraphlearning_b200 import training as T
dev = torch.device('cuda:0')
B, nh = 8192, 10
n = nh + 1
# (name, rows, N, K, bias, mask, w_layout, accumulate) as the step calls them (training._graph_backward / _ValueTrain.backward)
shapes = [('value L3 (1,100)', B, 1, 100, True, False, 0, False), ('value L2 (100,100)', B, 100, 100, True, True, 0, False),
          ('value L1 (100,32)', B, 100, 32, True, True, 0, False), ('value L0 (32,32)', B, 32, 32, True, True, 0, False),
          ('gcn layer (32,32) R=B*n', B * n, 32, 32, False, False, 1, True), ('w_a (32,32) R=B*n', B * n, 32, 32, False, False, 1, True),
          ('w_h.2 (32,64) R=B*nh', B * nh, 32, 64, True, True, 0, False), ('w_h.0 (64,5) R=B*nh', B * nh, 64, 5, True, True, 0, False),
          ('w_r.2 (32,64) R=B', B, 32, 64, True, True, 0, False), ('w_r.0 (64,9) R=B', B, 64, 9, True, True, 0, False)]
only = os.environ.get('BWD_ONLY')
nodw, nodx = os.environ.get('BWD_NODW'), os.environ.get('BWD_NODX')     # experiments: data gradient only / weight gradient only
for si, (name, R, N, K, bias, mask, layout, accum) in enumerate(shapes):
    if only is not None and si != int(only):
        continue
    ldN = 128 if N == 100 else max(N, 1)
    ldK = 128 if K == 100 else K
    G = torch.randn(R, ldN, device=dev)
    M = torch.randn(R, ldN, device=dev)
    X = torch.randn(R, ldK, device=dev)
    W = torch.randn(N, K, device=dev) if layout == 0 else torch.randn(K, N, device=dev)
    Gin = torch.zeros(R, ldK, device=dev)
    dW = torch.zeros_like(W)
    db = torch.zeros(N, device=dev) if bias else None
    def run():
        T._linear_bwd(T._rows(G, ldN), N, T._rows(X, ldK), K, R, W=W, w_layout=layout, mask=T._rows(M, ldN) if mask else None,
                      Gin=T._rows(Gin, ldK) if (K > 9 and not nodx) else None, accumulate=accum, dW=None if nodw else dW,
                      db=None if nodw else db, dev=dev)
    for _ in range(5): run()
    torch.cuda.synchronize()
    if only is not None:
        for _ in range(5): run()
        torch.cuda.synchronize()
        continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(50): run()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    traffic = R * ((2 if mask else 1) * N + (3 if accum else 2) * K) * 4 / 1e6
    print('%-28s R=%6d  %7.1f us   (%.1f MB -> %.2f TB/s)' % (name, R, us, traffic, traffic / us), flush=True)
