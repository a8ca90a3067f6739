"""NumPy restatement of the RGL hot path -- TEST INFRASTRUCTURE ONLY (see oracle/rgl_oracle.py for the rules).

An independent second oracle: the same math as `oracle/rgl_oracle.py` written with explicit einsum / exp / sum instead of
the ATen ops, so that a mistake shared by the torch restatement and the reference's own op sequence cannot hide.  It is not
bit-comparable with the reference (NumPy sums in another order); it is pinned in fp64 against the fp64 golden outputs
minted from the reference (`tests/golden/*.npz`: H64 / V64 / S64, agreement ~1e-12) and in fp32 within the parity
tolerance (`tests/test_oracle_golden.py`).

  mlp                 crowd_nav/policy/helpers.py:5-13
  rgl_forward         crowd_nav/policy/graph_model.py:63-66 (embedded_gaussian), :99-130
  value_forward       crowd_nav/policy/value_estimator.py:11-20
  statepred_forward   crowd_nav/policy/state_predictor.py:20-39
"""
import numpy as np


def mlp(x, sd, prefix='', last_relu=False):
    keys = []
    i = 0
    while (prefix + '%d.weight' % i) in sd:
        keys.append(i)
        i += 2
    for k, i in enumerate(keys):
        w, b = sd[prefix + '%d.weight' % i], sd[prefix + '%d.bias' % i]
        x = np.einsum('...k,ok->...o', x, w) + b            # Linear: y = x W^T + b, W stored [out, in]
        if k != len(keys) - 1 or last_relu:
            x = np.maximum(x, 0)
    return x


def softmax_rows(a):
    e = np.exp(a - a.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def rgl_forward(sd, robot, humans, skip_connection=True, layerwise_graph=False, return_A=False):
    """H[B,n,32] from robot[B,1,9], humans[B,Nh,5]; sd = RGL state-dict as numpy arrays."""
    X = np.concatenate([mlp(robot, sd, 'w_r.', last_relu=True), mlp(humans, sd, 'w_h.', last_relu=True)], axis=1)

    def attention(F_):
        return softmax_rows(np.einsum('bik,kl,bjl->bij', F_, sd['w_a'], F_))

    A = attention(X)
    A0 = A
    H = X
    num_layer = sum(1 for k in sd if k.startswith('Ws.'))
    for l in range(num_layer):
        if layerwise_graph and l > 0:
            A = attention(H)
        nxt = np.maximum(np.einsum('bij,bjk,kl->bil', A, H, sd['Ws.%d' % l]), 0)
        H = nxt + H if skip_connection else nxt
    return (H, A0) if return_A else H


def value_forward(graph_sd, value_sd, robot, humans, **kw):
    return mlp(rgl_forward(graph_sd, robot, humans, **kw)[:, 0, :], value_sd)


def statepred_forward(graph_sd, motion_sd, robot, humans, **kw):
    return mlp(rgl_forward(graph_sd, robot, humans, **kw), motion_sd)[:, 1:, :]
