"""Differentiable torch-op statement of the path, on whatever device its tensors live.  Used for

  * graph configurations the sm_100a kernels are not specialised for: the seven similarity functions other than
    'embedded_gaussian' (graph_model.py:67-93; no shipped config selects them), non-default layer widths, human
    counts outside 1..31 -- on the GPU, checked against reference-minted goldens (tests/test_gpu_parity.py);
  * the backward pass of layerwise-graph configurations (autograd recompute on the inputs' device);
  * modules whose parameters live on the CPU and are called with CPU tensors (the reference's callers
    may keep a policy on the CPU) -- ops.require_cuda_or_cpu_module states the policy.

It is never a substitute for a missing librgl_b200.so: CUDA tensors always go through the kernels, and every op on
that path raises if the library cannot be loaded.
Math follows crowd_nav/policy/graph_model.py:63-130, value_estimator.py:18-19, state_predictor.py:28,36.
"""
import torch


def similarity(X, w_a, kind):
    XT = X.transpose(1, 2)
    if kind == 'embedded_gaussian':
        return torch.softmax(torch.matmul(torch.matmul(X, w_a), XT), dim=2)
    if kind == 'gaussian':
        return torch.softmax(torch.matmul(X, XT), dim=2)
    if kind in ('cosine', 'cosine_softmax'):
        A = torch.matmul(X, XT)
        mag = torch.norm(A, dim=2, keepdim=True)
        A = A / torch.matmul(mag, mag.transpose(1, 2))
        return torch.softmax(A, dim=2) if kind == 'cosine_softmax' else A
    if kind == 'squared':
        A = torch.matmul(X, XT)
        A = A * A
        return A / A.sum(dim=2, keepdim=True)
    n = X.size(1)
    if kind == 'equal_attention':
        return (torch.ones(n, n, dtype=X.dtype, device=X.device) / n).expand(X.size(0), n, n)
    if kind == 'diagonal':
        return torch.eye(n, dtype=X.dtype, device=X.device).expand(X.size(0), n, n)
    if kind == 'concatenation':
        B = X.size(0)
        pair = torch.cat([X.unsqueeze(2).expand(B, n, n, X.size(2)), X.unsqueeze(1).expand(B, n, n, X.size(2))], dim=3)
        return w_a(pair.reshape(B, n * n, -1)).reshape(B, n, n)
    raise NotImplementedError(kind)


def graph_forward(rgl, robot, humans, return_A=False):
    """rgl: an RGL module (parameters used directly so autograd reaches them)."""
    X = torch.cat([rgl.w_r(robot), rgl.w_h(humans)], dim=1)
    w_a = getattr(rgl, 'w_a', None)
    A = A_first = None
    if not rgl.layerwise_graph:
        A = A_first = similarity(X, w_a, rgl.similarity_function)
    H = X
    for i in range(rgl.num_layer):
        if rgl.layerwise_graph:
            A = similarity(H, w_a, rgl.similarity_function)
            if A_first is None:
                A_first = A
        nxt = torch.relu(torch.matmul(torch.matmul(A, H), rgl.Ws[i]))
        H = nxt + H if rgl.skip_connection else nxt
    return (H, A_first) if return_A else H
