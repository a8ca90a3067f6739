"""The oracle restatement must reproduce the REFERENCE's own outputs (tests/golden, minted by
oracle/gen_golden.py from /root/reference) bit-for-bit in fp32, and to fp64 round-off in fp64."""
import numpy as np
import pytest
import torch

from conftest import FWD_CASES, graph_kw, load_golden
from oracle import planner_oracle as P
from oracle import rgl_oracle as O
from oracle import rgl_oracle_np as ON


@pytest.mark.parametrize('case', FWD_CASES)
def test_forward_bit_exact_vs_reference(case):
    g = load_golden(case)
    torch.set_num_threads(1)
    kw = graph_kw(g)
    with torch.no_grad():
        H, A = O.rgl_forward(g['graph1'], g['robot'], g['humans'], return_A=True, **kw)
        V = O.value_forward(g['graph1'], g['value'], g['robot'], g['humans'], **kw)
        S = O.statepred_forward(g['graph2'], g['motion'], g['robot'], g['humans'], **kw)
    assert torch.equal(H, g['H'])
    assert torch.equal(V, g['V'])
    assert torch.equal(S, g['S'])
    if g['A0'].numel():   # the reference records .A only when layerwise_graph is False (graph_model.py:114-116)
        assert np.array_equal(A[0].numpy(), g['A0'].numpy())


@pytest.mark.parametrize('case', FWD_CASES)
def test_forward_fp64_arbiter(case):
    g = load_golden(case)
    kw = graph_kw(g)
    r, h = g['robot'].double(), g['humans'].double()
    with torch.no_grad():
        H = O.rgl_forward(O.to_double(g['graph1']), r, h, **kw)
        V = O.value_forward(O.to_double(g['graph1']), O.to_double(g['value']), r, h, **kw)
        S = O.statepred_forward(O.to_double(g['graph2']), O.to_double(g['motion']), r, h, **kw)
    for got, ref in ((H, g['H64']), (V, g['V64']), (S, g['S64'])):
        assert float((got - ref).abs().max()) <= 1e-10 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize('case', FWD_CASES)
def test_numpy_restatement_matches_reference(case):
    """The independent NumPy restatement (einsum / exp / sum, no ATen ops) against the reference-minted fixtures: fp64 to
    round-off against the fp64 outputs, fp32 within the parity tolerance (1e-5 of the tensor scale) against the fp32 ones."""
    g = load_golden(case)
    kw = graph_kw(g)

    def np_sd(sd, dt):
        return {k: v.numpy().astype(dt) for k, v in sd.items()}

    for dt, keys, tol in ((np.float64, ('H64', 'V64', 'S64'), 1e-10), (np.float32, ('H', 'V', 'S'), 1e-5)):
        r, h = g['robot'].numpy().astype(dt), g['humans'].numpy().astype(dt)
        g1, g2 = np_sd(g['graph1'], dt), np_sd(g['graph2'], dt)
        H, A0 = ON.rgl_forward(g1, r, h, return_A=True, **kw)
        V = ON.value_forward(g1, np_sd(g['value'], dt), r, h, **kw)
        S = ON.statepred_forward(g2, np_sd(g['motion'], dt), r, h, **kw)
        for got, key in ((H, keys[0]), (V, keys[1]), (S, keys[2])):
            ref = g[key].numpy().astype(np.float64)
            assert got.shape == ref.shape and got.dtype == dt
            assert np.abs(got.astype(np.float64) - ref).max() <= tol * max(1e-3, np.abs(ref).max()), (case, key, dt)
        if dt == np.float32 and g['A0'].numel():
            assert np.abs(A0[0] - g['A0'].numpy()).max() <= 1e-5


def test_action_space_matches_reference():
    g = load_golden('planner_d1_nh5')
    acts, groups = P.build_action_space(1.0)
    assert acts.shape == (81, 2)
    assert np.array_equal(acts, g['actions'])
    assert list(groups) == list(g['action_group_index'])


def test_planner_depth1_matches_reference():
    """Reference predict() at depth 1 (81 SP + 81 VE batch-1 forwards) vs the restated tree."""
    g = load_golden('planner_d1_nh5')
    torch.set_num_threads(1)
    pl = P.OraclePlanner(g['graph1'], g['value'], g['graph2'], g['motion'])
    for b in range(g['robot'].shape[0]):
        robot, humans = g['robot'][b:b + 1], g['humans'][b:b + 1]
        a, v, table = pl.predict(robot, humans)
        assert a == int(g['chosen'][b])
        if not table:          # reach_destination short-circuit
            assert a == 0
            continue
        ref_vals = g['values'][b].numpy() if hasattr(g['values'], 'numpy') else g['values'][b]
        got_vals = np.array([table[i] for i in range(81)])
        # rewards: the reference evaluates the root reward in float64 on python floats -> exact match
        rew = np.array([pl.R((robot, humans), i) for i in range(81)], dtype=np.float64)
        ref_rew = np.asarray(g['rewards'][b], dtype=np.float64)
        assert np.array_equal(rew, ref_rew)
        assert np.array_equal(got_vals, np.asarray(ref_vals, dtype=np.float64))
        assert a == int(g['chosen'][b])


def test_reward_branches_covered():
    g = load_golden('planner_d1_nh5')
    rew = np.asarray(g['rewards'])
    assert (rew == -0.25).any() and (rew == 1).any() and (rew == 0).any()
    assert ((rew < 0) & (rew > -0.25)).any()


def test_next_robot_state_matches_reference_rule():
    robot = torch.tensor([[[1.5, -2.25, 0.1, 0.2, 0.3, 0.0, 4.0, 1.0, 1.57]]])
    out = O.next_robot_state(robot, 0.3, -0.7, 0.25)
    exp = robot.clone().squeeze()
    exp[0] = exp[0] + 0.3 * 0.25
    exp[1] = exp[1] + (-0.7) * 0.25
    exp[2] = 0.3
    exp[3] = -0.7
    assert torch.equal(out.squeeze(), exp)


TREE_CASES = ['planner_d2w2_nh5', 'planner_d2w2_a81_nh5', 'planner_d3w2_nh5', 'planner_d2w3_sparse_nh5']


def tree_kw(g):
    m = g['meta']
    return dict(planning_depth=int(m[4]), planning_width=int(m[5]), do_action_clip=True, sparse_search=bool(m[6]),
                speed_samples=int(m[7]), rotation_samples=int(m[8]))


@pytest.mark.parametrize('case', TREE_CASES)
def test_planner_depth_gt1_matches_patched_reference(case):
    """Depth > 1 look-ahead with action clipping: the restated tree vs the reference planner run through the one-line patch
    of oracle/make_ref.py (`values.append(float(value))`, model_predictive_rl.py:250).  Same kept action set at the root,
    same per-action values (float32 look-ahead rewards under NEP-50 numpy: <= 1e-7), same chosen action."""
    g = load_golden(case)
    torch.set_num_threads(1)
    pl = P.OraclePlanner(g['graph1'], g['value'], g['graph2'], g['motion'], **tree_kw(g))
    for b in range(g['robot'].shape[0]):
        a, v, table = pl.predict(g['robot'][b:b + 1], g['humans'][b:b + 1])
        kept = [int(x) for x in g['kept'][b]]
        assert sorted(table) == sorted(kept), (case, b)
        ref = dict(zip(kept, np.asarray(g['values'][b], dtype=np.float64)))
        for k in kept:
            assert abs(table[k] - ref[k]) <= 1e-7, (case, b, k)
        assert a == int(g['chosen'][b])
        assert int(g['traj'][b][0]) == a


def test_vendored_reference_modules_reproduce_the_goldens():
    """oracle/_ref (oracle/make_ref.py: the reference's own modules, vendored unmodified; what bench.py's reference arm
    times) reproduces the committed fixtures bit for bit -- the fixtures were minted from /root/reference itself."""
    from oracle import make_ref
    if not make_ref.enable():
        pytest.skip('oracle/_ref not built (needs the reference checkout: python oracle/make_ref.py)')
    from crowd_nav.configs.icra_benchmark.mp_separate import PolicyConfig
    from crowd_nav.policy.graph_model import RGL
    from crowd_nav.policy.state_predictor import StatePredictor
    from crowd_nav.policy.value_estimator import ValueEstimator
    for case in ('fwd_nh5_s0', 'fwd_nh20_s0'):
        g = load_golden(case)
        cfg = PolicyConfig()
        g1 = RGL(cfg, 9, 5)
        ve = ValueEstimator(cfg, g1)
        g2 = RGL(cfg, 9, 5)
        sp = StatePredictor(cfg, g2, 0.25)
        g1.load_state_dict(g['graph1'])
        ve.value_network.load_state_dict(g['value'])
        g2.load_state_dict(g['graph2'])
        sp.human_motion_predictor.load_state_dict(g['motion'])
        with torch.no_grad():
            assert torch.equal(g1((g['robot'], g['humans'])), g['H'])
            assert torch.equal(ve((g['robot'], g['humans'])), g['V'])
            assert torch.equal(sp((g['robot'], g['humans']), None)[1], g['S'])


def test_planner_unicycle_matches_reference():
    """ActionRot / unicycle branch (model_predictive_rl.py:204,319-321,337-340; state_predictor.py:53-58, including its
    element-7 indexing): the reference's own depth-1 predict() vs the restated tree, bit for bit."""
    g = load_golden('planner_d1_unicycle_nh5')
    torch.set_num_threads(1)
    pl = P.OraclePlanner(g['graph1'], g['value'], g['graph2'], g['motion'], kinematics='unicycle')
    assert np.array_equal(pl.actions, np.asarray(g['actions']))
    for b in range(g['robot'].shape[0]):
        robot, humans = g['robot'][b:b + 1], g['humans'][b:b + 1]
        a, v, table = pl.predict(robot, humans)
        assert a == int(g['chosen'][b])
        if not table:
            continue
        assert np.array_equal(np.array([table[i] for i in range(81)]), np.asarray(g['values'][b], dtype=np.float64))
        rew = np.array([pl.R((robot, humans), i) for i in range(81)], dtype=np.float64)
        assert np.array_equal(rew, np.asarray(g['rewards'][b], dtype=np.float64))


from conftest import SIM_CASES  # noqa: E402


@pytest.mark.parametrize('case', SIM_CASES)
def test_similarity_variants_bit_exact_vs_reference(case):
    """The seven other similarity functions (graph_model.py:67-93), fixtures minted from the reference modules."""
    g = load_golden(case)
    torch.set_num_threads(1)
    kw = graph_kw(g)
    with torch.no_grad():
        H = O.rgl_forward(g['graph1'], g['robot'], g['humans'], **kw)
        V = O.value_forward(g['graph1'], g['value'], g['robot'], g['humans'], **kw)
        S = O.statepred_forward(g['graph2'], g['motion'], g['robot'], g['humans'], **kw)
    assert torch.equal(H, g['H']) and torch.equal(V, g['V']) and torch.equal(S, g['S'])
