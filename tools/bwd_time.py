"""Warm timing (CUDA events, back-to-back launches) of rgl_linear_bwd for the shapes of the C4 training step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relationalgraphlearning_b200 import training as T
dev = torch.device('cuda:0')
B, nh = 8192, 10
n = nh + 1
shapes = [('value L3 (1,100)', B, 1, 100, True), ('value L2 (100,100)', B, 100, 100, True), ('value L1 (100,32)', B, 100, 32, True),
          ('value L0 (32,32)', B, 32, 32, True), ('gcn layer (32,32) R=B*n', B * n, 32, 32, False), ('w_a (32,32) R=B*n', B * n, 32, 32, False),
          ('w_h.2 (32,64) R=B*nh', B * nh, 32, 64, True), ('w_h.0 (64,5) R=B*nh', B * nh, 64, 5, True), ('w_r.2 (32,64) R=B', B, 32, 64, True),
          ('w_r.0 (64,9) R=B', B, 64, 9, True)]
only = os.environ.get('BWD_ONLY')
for si, (name, R, N, K, bias) in enumerate(shapes):
    if only is not None and si != int(only):
        continue
    ldN = 128 if N == 100 else max(N, 1)
    ldK = 128 if K == 100 else K
    G = torch.randn(R, ldN, device=dev)
    M = torch.randn(R, ldN, device=dev)
    X = torch.randn(R, ldK, device=dev)
    W = torch.randn(N, K, device=dev)
    Gin = torch.empty(R, ldK, device=dev)
    dW = torch.zeros(N, K, device=dev)
    db = torch.zeros(N, device=dev) if bias else None
    def run():
        T._linear_bwd(T._rows(G, ldN), N, T._rows(X, ldK), K, R, W=W, mask=T._rows(M, ldN), Gin=T._rows(Gin, ldK), dW=dW, db=db, dev=dev)
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
    traffic = R * (2 * N + 2 * K) * 4 / 1e6
    print('%-28s R=%6d  %7.1f us   (%.1f MB -> %.0f GB/s)' % (name, R, us, traffic, traffic / us * 1e3 / 1e3), flush=True)
