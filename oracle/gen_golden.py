"""Mint golden input/output vectors FROM THE REFERENCE ITSELF -- run in the build container only.

    python oracle/gen_golden.py            # needs /root/reference; writes tests/golden/*.npz

The reference (pure Python, `/root/reference`) cannot travel to the GPU box, and ships no golden
vectors of its own for this path.  This script imports the reference's unmodified
modules (`crowd_nav.policy.graph_model.RGL`, `value_estimator.ValueEstimator`,
`state_predictor.StatePredictor`, and -- with stubs for the absent simulator deps gym /
matplotlib / rvo2 / socialforce -- `model_predictive_rl.ModelPredictiveRL`), seeds them as
SURVEY.md §8(d) prescribes, runs them on CPU in fp32 and fp64, and stores weights, inputs and
outputs as small .npz fixtures.  Nothing here is copied from the reference; it is only executed.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get('RGL_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from relationalgraphlearning_b200.synthetic import synthetic_states  # noqa: E402


def load_ref_config(name):
    path = os.path.join(REF, 'crowd_nav', 'configs', 'icra_benchmark', name + '.py')
    spec = importlib.util.spec_from_file_location('config', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sd_np(sd, prefix):
    return {prefix + k: v.detach().cpu().numpy() for k, v in sd.items()}


def build_modules(seed, layerwise=False, skip=True, scale=None, similarity='embedded_gaussian'):
    from crowd_nav.policy.graph_model import RGL
    from crowd_nav.policy.value_estimator import ValueEstimator
    from crowd_nav.policy.state_predictor import StatePredictor
    cfg = load_ref_config('mp_separate').PolicyConfig()
    cfg.gcn.layerwise_graph = layerwise
    cfg.gcn.skip_connection = skip
    cfg.gcn.similarity_function = similarity
    torch.manual_seed(seed)
    g1 = RGL(cfg, 9, 5)
    ve = ValueEstimator(cfg, g1)
    g2 = RGL(cfg, 9, 5)
    sp = StatePredictor(cfg, g2, 0.25)
    if scale is not None:
        with torch.no_grad():
            for g in (g1, g2):
                if isinstance(getattr(g, 'w_a', None), torch.nn.Parameter):
                    g.w_a.mul_(scale)
                for w in g.Ws:
                    w.mul_(scale)
    return cfg, g1, ve, g2, sp


def forward_case(name, seed, nh, batch, layerwise=False, skip=True, scale=None, data_seed=1234, similarity='embedded_gaussian'):
    cfg, g1, ve, g2, sp = build_modules(seed, layerwise, skip, scale, similarity)
    robot, humans = synthetic_states(batch, nh, seed=data_seed)
    out = {}
    with torch.no_grad():
        H = g1((robot, humans)).clone()
        A = np.zeros((0, 0), np.float32) if g1.A is None else np.array(g1.A, copy=True)   # layerwise: reference never sets .A
        V = ve((robot, humans))
        S = sp((robot, humans), None)[1]
        g1.double(); ve.double(); g2.double(); sp.double()
        try:
            H64 = g1((robot.double(), humans.double())).clone()
            V64 = ve((robot.double(), humans.double()))
            S64 = sp((robot.double(), humans.double()), None)[1]
        except RuntimeError:
            # equal_attention / diagonal build a float32 A whatever the input dtype (graph_model.py:91-93): the reference
            # itself cannot run them in fp64; the fp32 outputs stand in
            H64, V64, S64 = H.double(), V.double(), S.double()
        g1.float(); ve.float(); g2.float(); sp.float()
    out.update(sd_np(g1.state_dict(), 'graph1/'))
    out.update(sd_np(ve.value_network.state_dict(), 'value/'))
    out.update(sd_np(g2.state_dict(), 'graph2/'))
    out.update(sd_np(sp.human_motion_predictor.state_dict(), 'motion/'))
    out.update(robot=robot.numpy(), humans=humans.numpy(), H=H.numpy(), A0=A, V=V.numpy(), S=S.numpy(),
               H64=H64.numpy(), V64=V64.numpy(), S64=S64.numpy(),
               meta=np.array([seed, nh, batch, int(layerwise), int(skip), data_seed], dtype=np.int64),
               scale=np.array([1.0 if scale is None else scale]))
    if similarity != 'embedded_gaussian':
        out['similarity'] = np.array(similarity)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name, 'V range', float(V.min()), float(V.max()), '|H|max', float(H.abs().max()))


def stub_sim_deps():
    for name in ('gym', 'gym.envs', 'gym.envs.registration', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.lines',
                 'matplotlib.animation', 'matplotlib.patches', 'matplotlib.collections', 'rvo2', 'socialforce',
                 'tensorboardX', 'git'):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules['gym'].Env = object
    sys.modules['gym.envs.registration'].register = lambda **kw: None
    sys.modules['matplotlib'].lines = sys.modules['matplotlib.lines']
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['matplotlib'].animation = sys.modules['matplotlib.animation']
    sys.modules['matplotlib'].patches = sys.modules['matplotlib.patches']
    sys.modules['matplotlib'].collections = sys.modules['matplotlib.collections']
    sys.modules['matplotlib'].use = lambda *a, **k: None
    sys.modules['matplotlib.collections'].PatchCollection = object


def planner_case(name, seed, nh, n_states, data_seed=77, kinematics='holonomic'):
    """Depth-1 predict() of the unmodified reference planner on n_states synthetic JointStates."""
    stub_sim_deps()
    from crowd_nav.policy.model_predictive_rl import ModelPredictiveRL
    from crowd_sim.envs.utils.state import FullState, ObservableState, JointState
    cfg = load_ref_config('mp_separate').PolicyConfig()
    cfg.action_space.kinematics = kinematics
    torch.manual_seed(seed)
    pol = ModelPredictiveRL()
    pol.configure(cfg)
    pol.set_time_step(0.25)
    pol.set_device(torch.device('cpu'))
    pol.set_phase('test')
    robot, humans = synthetic_states(n_states, nh, seed=data_seed)
    # bring a few states near collision / goal so every reward branch is exercised
    humans[1, 0, 0:2] = robot[1, 0, 0:2] + torch.tensor([0.55, 0.0])
    humans[2, 1, 0:2] = robot[2, 0, 0:2] + torch.tensor([0.0, 0.75])
    robot[3, 0, 0:2] = torch.tensor([0.05, 3.8])
    chosen, values, rewards, vnext = [], [], [], []
    with torch.no_grad():
        for b in range(n_states):
            r = [float(x) for x in robot[b, 0]]
            st = JointState(FullState(*r), [ObservableState(*[float(x) for x in humans[b, h]]) for h in range(nh)])
            act = pol.predict(st)
            idx = [i for i, a in enumerate(pol.action_space) if a == act][0]
            chosen.append(idx)
            vals, rews, vns = [], [], []
            st_t = st.to_tensor(add_batch_size=True, device=pol.device)
            for a in pol.action_space:
                nxt = pol.state_predictor(st_t, a)
                v, _ = pol.V_planning(nxt, 1, 1)
                rew = pol.estimate_reward(st, a)
                rews.append(float(rew))
                vns.append(float(v))
                vals.append(float(rew + pol.get_normalized_gamma() * v))
            values.append(vals); rewards.append(rews); vnext.append(vns)
            # look-ahead style reward: state given as a tensor tuple (tensor_to_joint_state path)
    actions = np.array([[a[0], a[1]] for a in pol.action_space], dtype=np.float64)     # (vx, vy) or, unicycle, (v, r)
    sd = pol.get_state_dict()
    out = {}
    out.update(sd_np(sd['graph_model1'], 'graph1/'))
    out.update(sd_np(sd['value_network'], 'value/'))
    out.update(sd_np(sd['graph_model2'], 'graph2/'))
    out.update(sd_np(sd['motion_predictor'], 'motion/'))
    out.update(robot=robot.numpy(), humans=humans.numpy(), chosen=np.array(chosen), values=np.array(values),
               rewards=np.array(rewards), vnext=np.array(vnext), actions=actions,
               action_group_index=np.array(pol.action_group_index),
               meta=np.array([seed, nh, n_states, data_seed], dtype=np.int64))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name, 'chosen', chosen)


def load_patched_planner():
    """crowd_nav/policy/model_predictive_rl_d.py from oracle/_ref: the reference planner with ONE line changed
    (`values.append(float(value))`, model_predictive_rl.py:250; recipe and rationale in oracle/make_ref.py) so that
    action_clip / the depth > 1 recursion run on this torch / numpy."""
    from oracle import make_ref
    if not make_ref.available():
        make_ref.make(quiet=True)
    path = os.path.join(make_ref.OUT, 'crowd_nav', 'policy', 'model_predictive_rl_d.py')
    spec = importlib.util.spec_from_file_location('crowd_nav.policy.model_predictive_rl_d', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ModelPredictiveRL


def planner_tree_case(name, seed, nh, n_states, depth, width, sparse=False, speed_samples=5, rotation_samples=16, data_seed=91):
    """predict() of the reference planner (patched copy, see load_patched_planner) with action clipping at depth > 1:
    root clipped action set, per-kept-action reward / look-ahead return / value, the chosen action and the action sequence
    of the best trajectory (model_predictive_rl.py:212-233, 242-302)."""
    stub_sim_deps()
    ModelPredictiveRL = load_patched_planner()
    from crowd_sim.envs.utils.state import FullState, ObservableState, JointState
    cfg = load_ref_config('mp_separate').PolicyConfig()
    mp = cfg.model_predictive_rl
    mp.planning_depth, mp.planning_width, mp.do_action_clip = depth, width, True
    if sparse:
        mp.sparse_search = True
    cfg.action_space.speed_samples, cfg.action_space.rotation_samples = speed_samples, rotation_samples
    torch.manual_seed(seed)
    pol = ModelPredictiveRL()
    pol.configure(cfg)
    pol.set_time_step(0.25)
    pol.set_device(torch.device('cpu'))
    pol.set_phase('test')
    robot, humans = synthetic_states(n_states, nh, seed=data_seed)
    humans[1, 0, 0:2] = robot[1, 0, 0:2] + torch.tensor([0.6, 0.1])          # near-collision root
    robot[2, 0, 0:2] = torch.tensor([0.1, 3.7])                               # near-goal root
    chosen, kept, values, rewards, rets, trajs = [], [], [], [], [], []
    with torch.no_grad():
        for b in range(n_states):
            r = [float(x) for x in robot[b, 0]]
            st = JointState(FullState(*r), [ObservableState(*[float(x) for x in humans[b, h]]) for h in range(nh)])
            act = pol.predict(st)
            index = {a: i for i, a in enumerate(pol.action_space)}
            chosen.append(index[act])
            trajs.append([index[a] for (_, a, _) in pol.get_traj() if a is not None])
            st_t = st.to_tensor(add_batch_size=True, device=pol.device)
            clipped = pol.action_clip(st_t, pol.action_space, pol.planning_width)
            kept.append([index[a] for a in clipped])
            vals, rews, rts = [], [], []
            for a in clipped:
                nxt = pol.state_predictor(st_t, a)
                ret, _ = pol.V_planning(nxt, pol.planning_depth, pol.planning_width)
                rew = pol.estimate_reward(st, a)
                rews.append(float(rew))
                rts.append(float(ret))
                vals.append(float(rew + pol.get_normalized_gamma() * ret))
            values.append(vals); rewards.append(rews); rets.append(rts)
    actions = np.array([[a.vx, a.vy] for a in pol.action_space], dtype=np.float64)
    sd = pol.get_state_dict()
    out = {}
    out.update(sd_np(sd['graph_model1'], 'graph1/'))
    out.update(sd_np(sd['value_network'], 'value/'))
    out.update(sd_np(sd['graph_model2'], 'graph2/'))
    out.update(sd_np(sd['motion_predictor'], 'motion/'))
    tl = max(len(t) for t in trajs)
    out.update(robot=robot.numpy(), humans=humans.numpy(), chosen=np.array(chosen), kept=np.array(kept), values=np.array(values),
               rewards=np.array(rewards), rets=np.array(rets), traj=np.array([t + [-1] * (tl - len(t)) for t in trajs]),
               actions=actions, action_group_index=np.array(pol.action_group_index),
               meta=np.array([seed, nh, n_states, data_seed, depth, width, int(sparse), speed_samples, rotation_samples], dtype=np.int64))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name, 'chosen', chosen, 'kept', kept, 'traj', trajs)


def gcn_case(name, seed, nh, n_states, batch=24, data_seed=55, layerwise=False, skip=True, num_layer=2):
    """The model-free GCN policy (crowd_nav/policy/gcn.py + multi_human_rl.py + cadrl.py, config rgl.py): value-network
    outputs on a batch of rotated joint states, and predict() -- all 81 action values and the chosen action -- on
    n_states JointStates."""
    stub_sim_deps()
    from crowd_nav.policy.gcn import GCN
    from crowd_sim.envs.utils.state import FullState, ObservableState, JointState
    cfg = load_ref_config('rgl').PolicyConfig()
    cfg.gcn.layerwise_graph, cfg.gcn.skip_connection, cfg.gcn.num_layer = layerwise, skip, num_layer
    torch.manual_seed(seed)
    pol = GCN()
    pol.configure(cfg)
    pol.set_device(torch.device('cpu'))
    pol.set_phase('test')
    pol.time_step = 0.25
    with torch.no_grad():
        pol.model.w_a.mul_(0.2); pol.model.w1.mul_(0.2)
        if num_layer == 2:
            pol.model.w2.mul_(0.2)
    robot, humans = synthetic_states(max(n_states, batch), nh, seed=data_seed)
    # value network on rotated joint states of the first `batch` synthetic states
    joint = torch.cat([robot[:batch].expand(batch, nh, 9), humans[:batch]], dim=2)            # [batch, nh, 14]
    rotated = torch.stack([pol.rotate(joint[b]) for b in range(batch)])                       # [batch, nh, 13]
    with torch.no_grad():
        values = pol.model(rotated)
        A0 = np.array(pol.model.A, copy=True)
        pol.model.double()
        values64 = pol.model(rotated.double())
        pol.model.float()
    chosen, avals = [], []
    with torch.no_grad():
        for b in range(n_states):
            r = [float(x) for x in robot[b, 0]]
            st = JointState(FullState(*r), [ObservableState(*[float(x) for x in humans[b, h]]) for h in range(nh)])
            act = pol.predict(st)
            chosen.append([i for i, a in enumerate(pol.action_space) if a == act][0])
            avals.append([float(v) for v in pol.action_values])
    out = sd_np(pol.model.state_dict(), 'model/')
    out.update(robot=robot.numpy(), humans=humans.numpy(), rotated=rotated.numpy(), values=values.numpy(), values64=values64.numpy(),
               A0=A0, chosen=np.array(chosen), action_values=np.array(avals),
               actions=np.array([[a.vx, a.vy] for a in pol.action_space], dtype=np.float64),
               meta=np.array([seed, nh, n_states, batch, data_seed, int(layerwise), int(skip), num_layer], dtype=np.int64))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name, 'chosen', chosen, 'V range', float(values.min()), float(values.max()))


SIMILARITIES = ['gaussian', 'cosine', 'cosine_softmax', 'concatenation', 'squared', 'equal_attention', 'diagonal']


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    forward_case('fwd_nh5_s0', 0, 5, 64)
    forward_case('fwd_nh5_s1', 1, 5, 64)
    forward_case('fwd_nh5_s2', 2, 5, 64)
    forward_case('fwd_nh5_b1', 0, 5, 1, data_seed=5)            # C1: batch 1
    forward_case('fwd_nh10_s0', 0, 10, 48)
    forward_case('fwd_nh20_s0', 0, 20, 40)
    forward_case('fwd_nh1_s0', 0, 1, 33)                       # smallest graph
    forward_case('fwd_nh5_trained', 1, 5, 64, scale=0.2)       # trained-like logits
    forward_case('fwd_nh5_adversarial', 2, 5, 64, scale=4.0)   # softmax saturation
    forward_case('fwd_nh5_layerwise_noskip', 0, 5, 64, layerwise=True, skip=False)   # BasePolicyConfig defaults
    forward_case('fwd_nh5_layerwise_skip', 1, 5, 64, layerwise=True, skip=True)
    # the seven other similarity functions of graph_model.py:67-93 (trained-like scale keeps the un-normalised ones finite)
    for sim in SIMILARITIES:
        forward_case('fwd_nh5_sim_' + sim, 3, 5, 16, layerwise=(sim in ('gaussian', 'squared')), scale=0.2, similarity=sim)
    planner_case('planner_d1_nh5', 0, 5, 8)
    gcn_case('gcn_nh5', 0, 5, 6)                                  # model-free GCN policy (rgl.py)
    gcn_case('gcn_nh5_layerwise_noskip', 1, 5, 3, layerwise=True, skip=False)
    gcn_case('gcn_nh3_l1', 2, 3, 3, num_layer=1)
    planner_case('planner_d1_unicycle_nh5', 1, 5, 6, data_seed=78, kinematics='unicycle')     # ActionRot branch (:204,:319-321,:337-340)
    # depth > 1 look-ahead with action clipping, through the one-line-patched copy of the reference planner
    planner_tree_case('planner_d2w2_nh5', 0, 5, 6, 2, 2, speed_samples=2, rotation_samples=5)           # BASELINE C3: 11 actions
    planner_tree_case('planner_d2w2_a81_nh5', 1, 5, 4, 2, 2)                                            # mp_separate_dp: 81 actions
    planner_tree_case('planner_d3w2_nh5', 2, 5, 4, 3, 2, speed_samples=2, rotation_samples=5)
    planner_tree_case('planner_d2w3_sparse_nh5', 0, 5, 4, 2, 3, sparse=True)                            # sparse search (:252-263)


if __name__ == '__main__':
    main()
