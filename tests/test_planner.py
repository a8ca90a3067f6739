"""Planner: host-side API checks on CPU, batched look-ahead vs the oracle tree and the reference golden on GPU."""
import numpy as np
import pytest
import torch

from conftest import assert_close_scaled, load_golden
from oracle import planner_oracle as P
from relationalgraphlearning_b200 import ops
from relationalgraphlearning_b200.config import policy_config
from relationalgraphlearning_b200.model_predictive_rl import ModelPredictiveRL
from relationalgraphlearning_b200.simtypes import ActionXY, FullState, JointState, ObservableState
from relationalgraphlearning_b200.synthetic import synthetic_states


def make_policy(g, dev, **cfg_kw):
    pol = ModelPredictiveRL()
    pol.time_step = 0.25
    pol.configure(policy_config(**cfg_kw))
    pol.set_time_step(0.25)
    sd = {'graph_model1': g['graph1'], 'graph_model2': g['graph2'], 'value_network': g['value'], 'motion_predictor': g['motion']}
    if cfg_kw.get('share_graph_model'):
        sd = {'graph_model': g['graph1'], 'value_network': g['value'], 'motion_predictor': g['motion']}
    if cfg_kw.get('linear_state_predictor'):
        sd = {'graph_model': g['graph1'], 'value_network': g['value']}
    pol.load_state_dict(sd)
    pol.set_device(dev)
    pol.set_phase('test')
    return pol


def joint_state(robot, humans, b):
    r = [float(x) for x in robot[b, 0]]
    return JointState(FullState(*r), [ObservableState(*[float(x) for x in humans[b, h]]) for h in range(humans.size(1))])


# ------------------------------------------------------------------ CPU: API surface
def test_policy_api_surface_and_checkpoint_keys():
    pol = ModelPredictiveRL()
    assert pol.name == 'ModelPredictiveRL' and pol.trainable and pol.multiagent_training
    pol.time_step = 0.25
    pol.configure(policy_config())
    assert set(pol.get_state_dict()) == {'graph_model1', 'graph_model2', 'value_network', 'motion_predictor'}
    assert len(pol.model) == 4 and pol.get_model() is pol.value_estimator
    pol2 = ModelPredictiveRL()
    pol2.time_step = 0.25
    pol2.configure(policy_config(share_graph_model=True))
    assert set(pol2.get_state_dict()) == {'graph_model', 'value_network', 'motion_predictor'}
    assert pol2.value_estimator.graph_model is pol2.state_predictor.graph_model
    pol3 = ModelPredictiveRL()
    pol3.time_step = 0.25
    pol3.configure(policy_config(linear_state_predictor=True))
    assert set(pol3.get_state_dict()) == {'graph_model', 'value_network'} and not pol3.state_predictor.trainable
    pol.set_time_step(0.5)
    assert pol.state_predictor.time_step == 0.5 and abs(pol.get_normalized_gamma() - 0.9 ** 0.5) < 1e-15
    for attr in ('epsilon', 'action_values', 'traj', 'planning_depth', 'planning_width', 'do_action_clip', 'kinematics',
                 'last_state', 'env', 'phase', 'device'):
        assert hasattr(pol, attr)
    with pytest.raises(AttributeError):
        pol.predict(None)


def test_action_space_matches_reference_golden():
    g = load_golden('planner_d1_nh5')
    pol = ModelPredictiveRL()
    pol.time_step = 0.25
    pol.configure(policy_config())
    pol.build_action_space(1.0)
    tab = np.array([[a.vx, a.vy] for a in pol.action_space])
    assert np.array_equal(tab, g['actions'])
    assert list(pol.action_group_index) == list(g['action_group_index'])


def test_checkpoint_interchange_with_reference_layout(tmp_path):
    g = load_golden('fwd_nh5_s0')
    ref_ckpt = {'graph_model1': g['graph1'], 'graph_model2': g['graph2'], 'value_network': g['value'], 'motion_predictor': g['motion']}
    f = tmp_path / 'rl_model.pth'
    torch.save(ref_ckpt, f)                      # what the reference's save_model writes (model_predictive_rl.py:148-149)
    pol = ModelPredictiveRL()
    pol.time_step = 0.25
    pol.configure(policy_config())
    pol.device = torch.device('cpu')
    pol.load_model(str(f))
    out = tmp_path / 'resaved.pth'
    pol.save_model(str(out))
    back = torch.load(str(out))
    for grp in ref_ckpt:
        assert list(back[grp].keys()) == list(ref_ckpt[grp].keys())
        for k in ref_ckpt[grp]:
            assert torch.equal(back[grp][k], ref_ckpt[grp][k])


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_plan_expand_matches_oracle_rewards_and_kinematics(cuda_device):
    g = load_golden('planner_d1_nh5')
    acts = torch.from_numpy(np.asarray(g['actions'])).to(cuda_device)
    robot, humans = g['robot'].to(cuda_device), g['humans'].to(cuda_device)
    nxt, rew = ops.plan_expand(robot, humans, acts, 0.25)
    E, A = robot.size(0), acts.size(0)
    rew = rew.view(E, A).cpu().double().numpy()
    # golden rewards come from the reference's own estimate_reward (float64): the device computes float64 too
    ref = np.asarray(g['rewards'], dtype=np.float64)
    assert np.abs(rew - ref).max() <= 1e-7
    assert ((ref == -0.25) == (rew == -0.25)).all() and ((ref == 1) == (rew == 1)).all()
    from oracle import rgl_oracle as O
    for a in (0, 5, 40, 80):
        exp = O.next_robot_state(g['robot'], float(g['actions'][a][0]), float(g['actions'][a][1]), 0.25)
        assert torch.equal(nxt.view(E, A, 1, 9)[:, a].cpu(), exp)


@pytest.mark.gpu
def test_depth1_predict_matches_reference_golden(cuda_device):
    """81 + 81 batch-1 forwards of the reference vs 5 launches here: same per-action values, same action."""
    g = load_golden('planner_d1_nh5')
    pol = make_policy(g, cuda_device)
    robot, humans = g['robot'].to(cuda_device), g['humans'].to(cuda_device)
    best, det = pol.predict_batch(robot, humans, return_details=True)
    vals = det['value'].cpu().double().numpy()
    ref = np.asarray(g['values'], dtype=np.float64)
    assert np.abs(vals - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    for b in range(robot.size(0)):
        a = pol.predict(joint_state(g['robot'], g['humans'], b))
        want = int(g['chosen'][b])
        order = np.sort(ref[b])[::-1]
        if want != 0 or order[0] - order[1] > 1e-5:      # chosen==0 may be the reach_destination short-circuit
            assert a == pol.action_space[want], (b, a, want)
        assert isinstance(a, ActionXY)
        if pol.traj is not None and want != 0:
            assert pol.traj[0][1] == a and pol.traj[-1][1] is None


@pytest.mark.gpu
@pytest.mark.parametrize('cfg', [dict(planning_depth=2, planning_width=2, do_action_clip=True),
                                 dict(planning_depth=3, planning_width=2, do_action_clip=True),
                                 dict(planning_depth=2, planning_width=3, do_action_clip=True, sparse_search=True),
                                 dict(planning_depth=1, planning_width=4, do_action_clip=True),
                                 dict(planning_depth=2, planning_width=2, do_action_clip=True, linear_state_predictor=True),
                                 dict(planning_depth=2, planning_width=1, do_action_clip=False, speed_samples=2, rotation_samples=5)])
def test_tree_planning_matches_oracle(cfg, cuda_device):
    """Level-batched tree vs the batch-1 oracle tree (oracle/planner_oracle.py) on several root states."""
    g = load_golden('fwd_nh5_s1')
    pol = make_policy(g, cuda_device, **cfg)
    okw = dict(planning_depth=cfg['planning_depth'], planning_width=cfg['planning_width'], do_action_clip=cfg['do_action_clip'],
               sparse_search=cfg.get('sparse_search', False), linear_state_predictor=cfg.get('linear_state_predictor', False),
               speed_samples=cfg.get('speed_samples', 5), rotation_samples=cfg.get('rotation_samples', 16))
    orc = P.OraclePlanner(g['graph1'], g['value'], g['graph2'], g['motion'], **okw)
    robot, humans = synthetic_states(5, 5, seed=321)
    pol.build_action_space(1.0)
    best, det = pol.predict_batch(robot.to(cuda_device), humans.to(cuda_device), return_details=True)
    torch.set_num_threads(4)
    for b in range(robot.size(0)):
        a, v, table = orc.predict(robot[b:b + 1], humans[b:b + 1])
        if not table:
            continue
        # compare the value of every action the oracle evaluated at the root
        if det['acts'] is None:
            got = {i: float(det['value'][b, i]) for i in range(det['value'].size(1))}
        else:
            got = {int(det['acts'][b, k]): float(det['value'][b, k]) for k in range(det['acts'].size(1))}
        srt = sorted(table.values(), reverse=True)
        clear = len(srt) < 2 or srt[0] - srt[1] > 2e-5
        if set(got) == set(table):
            for i in table:
                assert abs(got[i] - table[i]) <= 2e-5 * max(1.0, abs(table[i])), (cfg, b, i, got[i], table[i])
            if clear:
                assert int(best[b]) == a
        else:
            # a near-tie flipped the clipped set; the chosen value must still be within tolerance of the oracle's best
            assert abs(max(got.values()) - v) <= 1e-4, (cfg, b, got, table)


@pytest.mark.gpu
def test_unicycle_predict_matches_reference_golden(cuda_device):
    """ActionRot kinematics: per-action values, rewards and chosen actions of the reference's depth-1 predict()."""
    from relationalgraphlearning_b200.simtypes import ActionRot
    g = load_golden('planner_d1_unicycle_nh5')
    pol = make_policy(g, cuda_device, kinematics='unicycle')
    pol.build_action_space(1.0)
    assert np.array_equal(np.array([[a.v, a.r] for a in pol.action_space]), np.asarray(g['actions']))
    robot, humans = g['robot'].to(cuda_device), g['humans'].to(cuda_device)
    best, det = pol.predict_batch(robot, humans, return_details=True)
    ref = np.asarray(g['values'], dtype=np.float64)
    assert np.abs(det['value'].cpu().double().numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    assert np.abs(det['rew'].cpu().double().numpy() - np.asarray(g['rewards'], dtype=np.float64)).max() <= 1e-6
    for b in range(robot.size(0)):
        order = np.sort(ref[b])[::-1]
        a = pol.predict(joint_state(g['robot'], g['humans'], b))
        assert isinstance(a, ActionRot)
        if int(g['chosen'][b]) != 0 or order[0] - order[1] > 1e-5:
            assert a == pol.action_space[int(g['chosen'][b])], b
    r = pol.estimate_reward(joint_state(g['robot'], g['humans'], 1), pol.action_space[7])
    assert abs(r - float(g['rewards'][1][7])) <= 1e-7


TREE_CASES = ['planner_d2w2_nh5', 'planner_d2w2_a81_nh5', 'planner_d3w2_nh5', 'planner_d2w3_sparse_nh5']


@pytest.mark.gpu
@pytest.mark.parametrize('case', TREE_CASES)
@pytest.mark.parametrize('graphed', [False, True])
def test_tree_planning_matches_patched_reference_golden(case, graphed, cuda_device):
    """Depth > 1 pin: golden vectors minted by the reference planner itself (one-line patch, oracle/make_ref.py) --
    kept action set after action_clip, value of every kept action, chosen action and the best trajectory's action
    sequence (model_predictive_rl.py:212-233, 242-302)."""
    g = load_golden(case)
    m = g['meta']
    pol = make_policy(g, cuda_device, planning_depth=int(m[4]), planning_width=int(m[5]), do_action_clip=True,
                      sparse_search=bool(m[6]), speed_samples=int(m[7]), rotation_samples=int(m[8]))
    pol.use_cuda_graphs = graphed
    pol.build_action_space(1.0)
    robot, humans = g['robot'].to(cuda_device), g['humans'].to(cuda_device)
    best, det = pol.predict_batch(robot, humans, return_details=True)
    for b in range(robot.size(0)):
        kept = [int(x) for x in g['kept'][b]]
        ref = dict(zip(kept, np.asarray(g['values'][b], dtype=np.float64)))
        got = {int(det['acts'][b, k]): float(det['value'][b, k]) for k in range(det['acts'].size(1))}
        srt = sorted(ref.values(), reverse=True)
        assert set(got) == set(ref), (case, b, got, ref)
        for k in kept:
            assert abs(got[k] - ref[k]) <= 1e-5 * max(1.0, abs(ref[k])), (case, b, k, got[k], ref[k])
        if srt[0] - srt[1] > 2e-5:
            assert int(best[b]) == int(g['chosen'][b]), (case, b)
            a = pol.predict(joint_state(g['robot'], g['humans'], b))
            assert a == pol.action_space[int(g['chosen'][b])]
            want = [int(x) for x in g['traj'][b] if x >= 0]
            have = [pol.action_space.index(t[1]) for t in pol.get_traj() if t[1] is not None]
            assert have == want, (case, b, have, want)


@pytest.mark.gpu
def test_batched_roots_equal_single_roots(cuda_device):
    g = load_golden('fwd_nh5_s0')
    pol = make_policy(g, cuda_device, planning_depth=2, planning_width=2, do_action_clip=True)
    pol.build_action_space(1.0)
    robot, humans = synthetic_states(33, 5, seed=8, device=cuda_device)
    best, det = pol.predict_batch(robot, humans, return_details=True)
    for b in (0, 7, 32):
        b1, d1 = pol.predict_batch(robot[b:b + 1], humans[b:b + 1], return_details=True)
        assert int(b1[0]) == int(best[b])
        assert torch.equal(d1['value'][0], det['value'][b])


@pytest.mark.gpu
def test_reference_shaped_helpers(cuda_device):
    g = load_golden('planner_d1_nh5')
    pol = make_policy(g, cuda_device)
    pol.build_action_space(1.0)
    st = joint_state(g['robot'], g['humans'], 0)
    r = pol.estimate_reward(st, pol.action_space[3])
    assert abs(r - float(g['rewards'][0][3])) <= 1e-7
    state = (g['robot'][:1].to(cuda_device), g['humans'][:1].to(cuda_device))
    v, traj = pol.V_planning(state, 1, 1)
    assert v.shape == (1, 1) and traj[0][1] is None
    clipped = pol.action_clip(state, pol.action_space, 3)
    assert len(clipped) == 3 and all(isinstance(a, ActionXY) for a in clipped)
    rt, ht = pol.transform(st)
    assert rt.shape == (1, 9) and ht.shape == (5, 5) and rt.device.type == 'cuda'


@pytest.mark.gpu
def test_graphed_predict_equals_eager_and_tracks_weight_updates(cuda_device):
    g = load_golden('planner_d1_nh5')
    pol = make_policy(g, cuda_device, planning_depth=2, planning_width=2, do_action_clip=True)
    pol.build_action_space(1.0)
    robot, humans = g['robot'][:1].to(cuda_device), g['humans'][:1].to(cuda_device)
    be, de = pol.predict_batch(robot, humans, return_details=True)
    for rep in range(3):
        bg, dg = pol.predict_batch_graphed(robot, humans)
        assert int(bg[0]) == int(be[0]) and torch.equal(dg['value'], de['value'])
    # other inputs through the same captured graph
    r2, h2 = g['robot'][4:5].to(cuda_device), g['humans'][4:5].to(cuda_device)
    b2, d2 = pol.predict_batch_graphed(r2, h2)
    e2, f2 = pol.predict_batch(r2, h2, return_details=True)
    assert int(b2[0]) == int(e2[0]) and torch.equal(d2['value'], f2['value'])
    v2 = d2['value'].clone()          # graph outputs are static tensors: the next replay overwrites them
    # a parameter update must be visible to the replayed graph (packed blobs are refreshed in place)
    with torch.no_grad():
        pol.value_estimator.value_network[6].bias.add_(0.25)
    b3, d3 = pol.predict_batch_graphed(r2, h2)
    e3, f3 = pol.predict_batch(r2, h2, return_details=True)
    assert torch.equal(d3['value'], f3['value']) and not torch.equal(d3['value'], v2)
    # predict() uses the graphed path and still builds the trajectory
    a = pol.predict(joint_state(g['robot'], g['humans'], 0))
    assert a == pol.action_space[int(be[0])] and pol.traj[0][1] == a


@pytest.mark.gpu
def test_predict_many_equals_per_state_predict(cuda_device):
    """Vectorised front end: many environments stepped together give the actions predict() gives one state at a time."""
    g = load_golden('planner_d2w2_nh5')
    pol = make_policy(g, cuda_device, planning_depth=2, planning_width=2, do_action_clip=True, speed_samples=2, rotation_samples=5)
    pol.build_action_space(1.0)
    robot, humans = synthetic_states(40, 5, seed=66)
    robot[7, 0, 0:2] = robot[7, 0, 5:7]                       # already at the goal: short-circuit to the stop action
    states = [joint_state(robot, humans, b) for b in range(40)]
    many = pol.predict_many(states)
    assert many[7] == ActionXY(0, 0)
    for b in (0, 7, 13, 39):
        assert pol.predict(states[b]) == many[b]
    # train phase: epsilon-greedy per state, last_state holds every transformed state
    pol.set_phase('train')
    pol.set_epsilon(0.5)
    np.random.seed(3)
    acts = pol.predict_many(states)
    assert len(acts) == 40 and all(a in pol.action_space for a in acts)
    assert len(pol.last_state) == 40 and pol.last_state[5][0].shape == (1, 9) and pol.last_state[5][1].shape == (5, 5)
    rt, ht = pol.transform(states[5])
    assert torch.equal(pol.last_state[5][0], rt) and torch.equal(pol.last_state[5][1], ht)
