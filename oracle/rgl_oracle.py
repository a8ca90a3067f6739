"""CPU oracle for the RGL hot path -- TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it.  Nothing under
`relationalgraphlearning_b200/` imports it.

It is a functional (state-dict in, tensors out) restatement of the reference's CPU PyTorch math,
written against the reference call sites so the same ATen ops run in the same order:

  mlp                 crowd_nav/policy/helpers.py:5-13      (Linear, ReLU after every layer except the
                                                             last unless last_relu)
  similarity          crowd_nav/policy/graph_model.py:63-97
  rgl_forward         crowd_nav/policy/graph_model.py:99-130
  value_forward       crowd_nav/policy/value_estimator.py:11-20
  statepred_forward   crowd_nav/policy/state_predictor.py:20-39
  next_robot_state    crowd_nav/policy/state_predictor.py:41-60   (batched here; reference is B=1 only)
  linear_motion       crowd_nav/policy/state_predictor.py:110-118

Parity pin: the reference ships NO golden vectors for this path; the pin is
the reference modules themselves, imported from /root/reference by `oracle/gen_golden.py`, whose
outputs are committed under tests/golden/ and against which this restatement is checked
bit-for-bit (`tests/test_oracle_golden.py`).

All functions are dtype-generic: pass float64 tensors to get the fp64 arbiter used by the
tolerance rule of SURVEY.md §8(c).
"""
import torch
import torch.nn.functional as F


def mlp(x, sd, prefix, last_relu=False):
    """nn.Sequential(Linear, ReLU, Linear, ...) addressed by its state-dict keys
    '<prefix>{2i}.weight' / '<prefix>{2i}.bias' (helpers.py:5-13)."""
    idx = []
    i = 0
    while (prefix + '%d.weight' % i) in sd:
        idx.append(i)
        i += 2
    for n, i in enumerate(idx):
        x = F.linear(x, sd[prefix + '%d.weight' % i], sd[prefix + '%d.bias' % i])
        if n != len(idx) - 1 or last_relu:
            x = torch.relu(x)
    return x


def similarity(X, sd, kind='embedded_gaussian'):
    """Normalised pairwise similarity A[B,n,n] (graph_model.py:63-97)."""
    XT = X.permute(0, 2, 1)
    if kind == 'embedded_gaussian':
        return torch.softmax(torch.matmul(torch.matmul(X, sd['w_a']), XT), dim=2)
    if kind == 'gaussian':
        return torch.softmax(torch.matmul(X, XT), dim=2)
    if kind in ('cosine', 'cosine_softmax'):
        A = torch.matmul(X, XT)
        mag = torch.norm(A, dim=2, keepdim=True)
        A = A / torch.matmul(mag, mag.permute(0, 2, 1))
        return torch.softmax(A, dim=2) if kind == 'cosine_softmax' else A
    if kind == 'squared':
        A = torch.matmul(X, XT)
        A = A * A
        return A / torch.sum(A, dim=2, keepdim=True)
    n = X.size(1)
    if kind == 'concatenation':
        # A[i][j] = w_a([X_i | X_j]) with w_a = mlp(2*X_dim, [2*X_dim, 1], last_relu=True), no normalisation (:81-86)
        import itertools
        idx = torch.LongTensor([p for p in itertools.product(range(n), repeat=2)]).reshape(-1)
        pair = torch.index_select(X, 1, idx).reshape(-1, n * n, X.size(2) * 2)
        return mlp(pair, sd, 'w_a.', last_relu=True).reshape(-1, n, n)
    if kind == 'equal_attention':
        return (torch.ones(n, n, dtype=X.dtype) / n).expand(X.size(0), n, n)
    if kind == 'diagonal':
        return torch.eye(n, dtype=X.dtype).expand(X.size(0), n, n)
    raise NotImplementedError(kind)


def rgl_forward(sd, robot, humans, skip_connection=True, layerwise_graph=False,
                similarity_function='embedded_gaussian', return_A=False):
    """H[B,n,32] from (robot[B,1,9], humans[B,Nh,5]) (graph_model.py:99-130).

    `sd` = RGL.state_dict(): w_r.{0,2}.{weight,bias}, w_h.{0,2}.{weight,bias}, w_a, Ws.{i}.
    The skip add is out-of-place; forward values are identical to the reference's in-place add
    (SURVEY.md §5).
    """
    X = torch.cat([mlp(robot, sd, 'w_r.', last_relu=True), mlp(humans, sd, 'w_h.', last_relu=True)], dim=1)
    num_layer = 0
    while ('Ws.%d' % num_layer) in sd:
        num_layer += 1
    A = None
    if not layerwise_graph:
        A = similarity(X, sd, similarity_function)
    A_first = A
    H = X
    for i in range(num_layer):
        if layerwise_graph:
            A = similarity(H, sd, similarity_function)
            if A_first is None:
                A_first = A
        nxt = torch.relu(torch.matmul(torch.matmul(A, H), sd['Ws.%d' % i]))
        if skip_connection:
            nxt = nxt + H
        H = nxt
    if return_A:
        return H, A_first
    return H


def value_forward(graph_sd, value_sd, robot, humans, **graph_kw):
    """V[B,1] = value_network(RGL(state)[:, 0, :]) (value_estimator.py:11-20)."""
    H = rgl_forward(graph_sd, robot, humans, **graph_kw)
    return mlp(H[:, 0, :], value_sd, '')


def statepred_forward(graph_sd, motion_sd, robot, humans, **graph_kw):
    """next_humans[B,Nh,5] = human_motion_predictor(RGL(state))[:, 1:, :] (state_predictor.py:28,36)."""
    H = rgl_forward(graph_sd, robot, humans, **graph_kw)
    return mlp(H, motion_sd, '')[:, 1:, :]


def next_robot_state(robot, vx, vy, time_step, kinematics='holonomic'):
    """Batched robot kinematic step (state_predictor.py:41-60, holonomic branch :48-52).

    robot[B,1,9]; vx, vy broadcastable to [B].  The reference multiplies python floats
    (float64) and writes the product into an fp32 tensor element: p32 + fp32(a*dt) is done as an
    fp32 tensor + python-scalar add, i.e. fp32(p) + fp32(a*dt) rounded to fp32.
    """
    if kinematics != 'holonomic':
        return next_robot_state_unicycle(robot, vx, vy, time_step)
    out = robot.clone()
    dt = float(time_step)
    vx = torch.as_tensor(vx, dtype=torch.float64)
    vy = torch.as_tensor(vy, dtype=torch.float64)
    out[:, 0, 0] = robot[:, 0, 0] + (vx * dt).to(robot.dtype)
    out[:, 0, 1] = robot[:, 0, 1] + (vy * dt).to(robot.dtype)
    out[:, 0, 2] = vx.to(robot.dtype)
    out[:, 0, 3] = vy.to(robot.dtype)
    return out


def next_robot_state_unicycle(robot, v, r, time_step):
    """Unicycle branch of state_predictor.py:53-58 for robot[1,1,9], ActionRot (v, r) as python floats.

    The reference adds the rotation to element 7 (v_pref; the heading is element 8 -- SURVEY.md 5; kept), then
    `np.cos(next_state[7]) * action.v * self.time_step`: np.cos of a 0-dim fp32 tensor comes back as a 0-dim fp32
    TENSOR (Tensor.__array_wrap__), so the products are fp32 tensor-times-python-scalar operations."""
    import numpy as np
    assert robot.shape[0] == 1
    out = robot.clone().squeeze()
    out[7] = out[7] + float(r)
    out[0] = out[0] + np.cos(out[7]) * float(v) * float(time_step)
    out[1] = out[1] + np.sin(out[7]) * float(v) * float(time_step)
    out[2] = np.cos(out[7]) * float(v)
    out[3] = np.sin(out[7]) * float(v)
    return out.unsqueeze(0).unsqueeze(0)


def linear_motion(humans):
    """LinearStatePredictor human step: p += v (no time-step factor; state_predictor.py:110-118)."""
    out = humans.clone()
    out[..., 0] = humans[..., 0] + humans[..., 2]
    out[..., 1] = humans[..., 1] + humans[..., 3]
    return out


def to_double(sd):
    return {k: v.double() for k, v in sd.items()}
