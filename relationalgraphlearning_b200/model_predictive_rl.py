"""ModelPredictiveRL -- drop-in for crowd_nav/policy/model_predictive_rl.py:15-370.

Same class name, attributes, configure()/predict()/state-dict API (so train.py / test.py / Explorer /
CrowdSim see the object they expect), but the d-step / w-width look-ahead is evaluated LEVEL-BATCHED on
the GPU instead of one batch-1 module forward per (node, action):

  reference (model_predictive_rl.py:222-231)        here
  ---------------------------------------------     -------------------------------------------------
  for action in action_space:                       next_humans = SP(state)            1 launch (the human
      next = state_predictor(state, action)            branch ignores the action, state_predictor.py:28,36)
      ret  = V_planning(next, depth, width)          next_robot, reward = rgl_plan_expand     1 launch
      r    = estimate_reward(state, action)          V = VE(next_robot[A], humans shared)     2 launches
      value = r + gamma_bar * ret; strict '>'        rgl_plan_argmax                          1 launch

and the same per tree level for depth > 1 (action_clip + recursion, :242-302), for any number of root
states at once (`predict_batch`).  Depth-1 planning = 5 launches instead of 162 module forwards.

Deviations, all documented in DESIGN.md: the reference planner does not run at depth > 1 on current
torch/numpy without a one-line change (`float(value)` at :250, oracle/make_ref.py); the
semantics implemented here -- top-`width` by value with ties broken by lower action index, first-maximum
argmax -- are pinned by golden vectors minted from that patched reference (tests/golden/planner_d*).
Look-ahead rewards are evaluated in float64 on the fp32 state values.  Both kinematics are implemented
(ActionXY holonomic, ActionRot unicycle).
"""
import logging

import numpy as np
import torch

from . import ops
from .graph_model import RGL
from .simtypes import ActionRot, ActionXY, Policy, joint_state_to_tensors, joint_states_to_tensors
from .state_predictor import LinearStatePredictor, StatePredictor
from .value_estimator import ValueEstimator


class ModelPredictiveRL(Policy):
    def __init__(self):
        super().__init__()
        self.name = 'ModelPredictiveRL'
        self.trainable = True
        self.multiagent_training = True
        self.kinematics = None
        self.epsilon = None
        self.gamma = None
        self.sampling = None
        self.speed_samples = None
        self.rotation_samples = None
        self.action_space = None
        self.rotation_constraint = None
        self.speeds = None
        self.rotations = None
        self.action_values = None
        self.robot_state_dim = 9
        self.human_state_dim = 5
        self.v_pref = 1
        self.share_graph_model = None
        self.value_estimator = None
        self.linear_state_predictor = None
        self.state_predictor = None
        self.planning_depth = None
        self.planning_width = None
        self.do_action_clip = None
        self.sparse_search = None
        self.sparse_speed_samples = 2
        self.sparse_rotation_samples = 8
        self.action_group_index = []
        self.traj = None
        self._actions_dev = None          # device copy of the action table (float64 [A,2])
        self._groups_dev = None
        self.stat_value_states = 0        # states pushed through the value estimator / state predictor (bench counters)
        self.stat_sp_states = 0
        self.use_cuda_graphs = True       # predict(): replay a captured CUDA graph of the whole look-ahead per (E, Nh) shape
        self._graphs = {}
        self._num_groups = 1

    # ------------------------------------------------------------------ configuration (:48-105)
    def configure(self, config):
        self.set_common_parameters(config)
        mp = config.model_predictive_rl
        self.planning_depth = mp.planning_depth
        self.do_action_clip = mp.do_action_clip
        if hasattr(mp, 'sparse_search'):
            self.sparse_search = mp.sparse_search
        self.planning_width = mp.planning_width
        self.share_graph_model = mp.share_graph_model
        self.linear_state_predictor = mp.linear_state_predictor

        if self.linear_state_predictor:
            self.state_predictor = LinearStatePredictor(config, self.time_step)
            graph_model = RGL(config, self.robot_state_dim, self.human_state_dim)
            self.value_estimator = ValueEstimator(config, graph_model)
            self.model = [graph_model, self.value_estimator.value_network]
        elif self.share_graph_model:
            graph_model = RGL(config, self.robot_state_dim, self.human_state_dim)
            self.value_estimator = ValueEstimator(config, graph_model)
            self.state_predictor = StatePredictor(config, graph_model, self.time_step)
            self.model = [graph_model, self.value_estimator.value_network, self.state_predictor.human_motion_predictor]
        else:
            graph_model1 = RGL(config, self.robot_state_dim, self.human_state_dim)
            self.value_estimator = ValueEstimator(config, graph_model1)
            graph_model2 = RGL(config, self.robot_state_dim, self.human_state_dim)
            self.state_predictor = StatePredictor(config, graph_model2, self.time_step)
            self.model = [graph_model1, graph_model2, self.value_estimator.value_network,
                          self.state_predictor.human_motion_predictor]

        self._graphs = {}                      # captured look-aheads are bound to the previous modules' weight blobs
        self._actions_dev = None
        logging.info('Planning depth: {}'.format(self.planning_depth))
        logging.info('Planning width: {}'.format(self.planning_width))
        logging.info('Sparse search: {}'.format(self.sparse_search))
        if self.planning_depth > 1 and not self.do_action_clip:
            logging.warning('Performing d-step planning without action space clipping!')

    def set_common_parameters(self, config):
        self.gamma = config.rl.gamma
        self.kinematics = config.action_space.kinematics
        self.sampling = config.action_space.sampling
        self.speed_samples = config.action_space.speed_samples
        self.rotation_samples = config.action_space.rotation_samples
        self.rotation_constraint = config.action_space.rotation_constraint

    def set_device(self, device):
        self.device = device
        for model in self.model:
            model.to(device)
        self._actions_dev = None
        self._graphs = {}

    def set_epsilon(self, epsilon):
        self.epsilon = epsilon

    def set_time_step(self, time_step):
        self.time_step = time_step
        self.state_predictor.time_step = time_step

    def get_normalized_gamma(self):
        return pow(self.gamma, self.time_step * self.v_pref)

    def get_model(self):
        return self.value_estimator

    # ------------------------------------------------------------------ checkpoints (:110-153)
    def get_state_dict(self):
        if self.state_predictor.trainable:
            if self.share_graph_model:
                return {'graph_model': self.value_estimator.graph_model.state_dict(),
                        'value_network': self.value_estimator.value_network.state_dict(),
                        'motion_predictor': self.state_predictor.human_motion_predictor.state_dict()}
            return {'graph_model1': self.value_estimator.graph_model.state_dict(),
                    'graph_model2': self.state_predictor.graph_model.state_dict(),
                    'value_network': self.value_estimator.value_network.state_dict(),
                    'motion_predictor': self.state_predictor.human_motion_predictor.state_dict()}
        return {'graph_model': self.value_estimator.graph_model.state_dict(),
                'value_network': self.value_estimator.value_network.state_dict()}

    def get_traj(self):
        return self.traj

    def load_state_dict(self, state_dict):
        if self.state_predictor.trainable:
            if self.share_graph_model:
                self.value_estimator.graph_model.load_state_dict(state_dict['graph_model'])
            else:
                self.value_estimator.graph_model.load_state_dict(state_dict['graph_model1'])
                self.state_predictor.graph_model.load_state_dict(state_dict['graph_model2'])
            self.value_estimator.value_network.load_state_dict(state_dict['value_network'])
            self.state_predictor.human_motion_predictor.load_state_dict(state_dict['motion_predictor'])
        else:
            self.value_estimator.graph_model.load_state_dict(state_dict['graph_model'])
            self.value_estimator.value_network.load_state_dict(state_dict['value_network'])

    def save_model(self, file):
        torch.save(self.get_state_dict(), file)

    def load_model(self, file):
        self.load_state_dict(torch.load(file, map_location=self.device))

    # ------------------------------------------------------------------ action table (:155-190)
    def build_action_space(self, v_pref):
        holonomic = self.kinematics == 'holonomic'
        speeds = [(np.exp((i + 1) / self.speed_samples) - 1) / (np.e - 1) * v_pref for i in range(self.speed_samples)]
        if holonomic:
            rotations = np.linspace(0, 2 * np.pi, self.rotation_samples, endpoint=False)
        else:
            rotations = np.linspace(-self.rotation_constraint, self.rotation_constraint, self.rotation_samples)
        action_space = [ActionXY(0, 0) if holonomic else ActionRot(0, 0)]
        groups = [0]
        for j, speed in enumerate(speeds):
            speed_index = 0 if j < 3 else 1               # two coarse speed groups for sparse search
            for i, rotation in enumerate(rotations):
                groups.append(speed_index * self.sparse_rotation_samples + i // 2)
                if holonomic:
                    action_space.append(ActionXY(speed * np.cos(rotation), speed * np.sin(rotation)))
                else:
                    action_space.append(ActionRot(speed, rotation))
        self.speeds = speeds
        self.rotations = rotations
        self.action_space = action_space
        self.action_group_index = groups
        self._actions_dev = None

    def _action_table(self):
        """Device copy of the action table, float64 [A,2]: (vx, vy) holonomic | (v, r) unicycle (ActionRot)."""
        if self._actions_dev is None or self._actions_dev.device != torch.device(self.device):
            tab = np.array([[a[0], a[1]] for a in self.action_space], dtype=np.float64)
            self._actions_dev = torch.from_numpy(tab).to(self.device)
            self._groups_dev = torch.tensor(self.action_group_index, dtype=torch.int32, device=self.device)
            self._num_groups = max(self.action_group_index) + 1
            self._graphs = {}
        return self._actions_dev

    # ------------------------------------------------------------------ batched tree
    def _expand(self, robot, humans, hb):
        """All actions of every state: (next_robot[N*A,1,9], reward[N,A]) in one launch."""
        acts = self._action_table()
        nxt, rew = ops.plan_expand(robot, humans, acts, self.time_step, humans_bcast=hb, kinematics=self.kinematics)
        return nxt, rew.view(robot.size(0), acts.size(0))

    def _next_humans(self, robot, humans, hb):
        """Predicted humans of every state (independent of the action).  -> [N,Nh,5]"""
        self.stat_sp_states += robot.size(0)
        if self.linear_state_predictor:
            nh = LinearStatePredictor.linear_motion_approximator(humans)
            return nh.repeat_interleave(hb, dim=0) if hb > 1 else nh
        return self.state_predictor.run(robot, humans, humans_bcast=hb)

    def _value(self, robot, humans, hb):
        self.stat_value_states += robot.size(0)
        return self.value_estimator.run(robot, humans, humans_bcast=hb).view(-1)

    def _clip(self, robot, humans, hb, width, want_value=False):
        """action_clip (:242-269) for N states at once: 5 launches (state predictor, expand, graph + value head over the
        N*A children, select).  -> dict(acts[N,width] int32, rew[N,width], robot[N*width,1,9], nh, all_rew[N,A], nxt, value)."""
        N, A = robot.size(0), len(self.action_space)
        nh = self._next_humans(robot, humans, hb)
        nxt, rew = self._expand(robot, humans, hb)
        V = self._value(nxt, nh, A)
        groups = None
        if self.sparse_search:
            assert width <= self._num_groups, 'sparse search keeps at most one action per group'
            groups = self._groups_dev
        acts, crew, crob, value = ops.plan_select(rew, V, N, A, self.get_normalized_gamma(), width, groups=groups,
                                                   next_robot=nxt, want_value=want_value)
        return dict(acts=acts, rew=crew, robot=crob, nh=nh, all_rew=rew, nxt=nxt, value=value)

    def _V_planning(self, robot, humans, hb, depth, width):
        """V_planning (:271-302) for N states.  Returns (ret[N], node) where node holds what get_traj needs."""
        N, A = robot.size(0), len(self.action_space)
        v = self._value(robot, humans, hb)
        if depth == 1:
            return v, None
        if self.do_action_clip:
            c = self._clip(robot, humans, hb, width)
            acts, child_rew, child_robot, nh, w = c['acts'], c['rew'], c['robot'], c['nh'], width
        else:
            nh = self._next_humans(robot, humans, hb)
            child_robot, child_rew = self._expand(robot, humans, hb)
            acts, w = None, A
        nv, child = self._V_planning(child_robot, nh, w, depth - 1, self.planning_width)
        ret, best = ops.plan_backup(v, nv, child_rew, N, w, self.get_normalized_gamma(), depth)     # first maximum wins (np.argmax, :298)
        node = dict(acts=acts, best=best, rew=child_rew, child_robot=child_robot, child_humans=nh, child=child, w=w)
        return ret, node

    def predict_batch(self, robot, humans, return_details=False):
        """Greedy branch of predict() (:212-233) for E root states at once.

        robot[E,1,9], humans[E,Nh,5] on self.device -> best action index [E] (int32, -1 if no finite value).
        Every step is a hand-written kernel: no ATen op runs between the input tensors and the chosen actions.
        """
        if self.action_space is None:
            self.build_action_space(self.v_pref)
        self._action_table()
        E, A = robot.size(0), len(self.action_space)
        with torch.no_grad():
            if self.do_action_clip:
                c = self._clip(robot, humans, 1, self.planning_width)
                acts, nh, rew, nxt, w = c['acts'], c['nh'], c['all_rew'], c['nxt'], self.planning_width
                ret, node = self._V_planning(c['robot'], nh, w, self.planning_depth, self.planning_width)
                value, best, best_action = ops.plan_argmax(c['rew'].reshape(-1), ret, E, w, self.get_normalized_gamma(), act_map=acts)
            else:
                acts = None
                nh = self._next_humans(robot, humans, 1)
                nxt, rew = self._expand(robot, humans, 1)
                ret, node = self._V_planning(nxt, nh, A, self.planning_depth, self.planning_width)
                value, best, best_action = ops.plan_argmax(rew.reshape(-1), ret, E, A, self.get_normalized_gamma())
        if return_details:
            return best_action, dict(value=value, best=best, acts=acts, rew=rew, next_robot=nxt, next_humans=nh, node=node, ret=ret)
        return best_action

    def _refresh_packed_weights(self):
        """Re-pack any module whose parameters changed (the captured graphs read the packed blobs in place)."""
        ve = self.value_estimator
        ops.packed_graph(ve.graph_model)
        ops.packed_value(ve.value_network, ve._pack_cache)
        sp = self.state_predictor
        if getattr(sp, 'trainable', False):
            ops.packed_graph(sp.graph_model)
            ops.packed_motion(sp.human_motion_predictor, sp._pack_cache)

    def predict_batch_graphed(self, robot, humans):
        """predict_batch(..., return_details=True) replayed from a CUDA graph captured once per input shape: the
        ~10-40 launches of a look-ahead become one cudaGraphLaunch (single-state planning latency, BASELINE C1)."""
        if self.action_space is None:
            self.build_action_space(self.v_pref)
        self._action_table()
        # everything the capture bakes in as a host scalar, a flag or a pointer is part of the key
        sp_g = getattr(self.state_predictor, 'graph_model', None)
        key = (robot.size(0), humans.size(1), self.planning_depth, self.planning_width, bool(self.do_action_clip),
               bool(self.sparse_search), float(self.time_step), str(robot.device), float(self.get_normalized_gamma()),
               self.kinematics, bool(self.linear_state_predictor), id(self.value_estimator), id(self.state_predictor),
               self.value_estimator.graph_model.flags(), sp_g.flags() if sp_g is not None else 0, len(self.action_space))
        self._refresh_packed_weights()
        entry = self._graphs.get(key)
        if entry is None:
            sr, sh = robot.clone(), humans.clone()
            side = torch.cuda.Stream(robot.device)
            side.wait_stream(torch.cuda.current_stream(robot.device))
            with torch.cuda.stream(side):
                self.predict_batch(sr, sh)                         # warm-up outside capture
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                best, det = self.predict_batch(sr, sh, return_details=True)
            entry = (g, sr, sh, best, det)
            if len(self._graphs) < 64:
                self._graphs[key] = entry
        g, sr, sh, best, det = entry
        sr.copy_(robot)
        sh.copy_(humans)
        g.replay()
        return best, det

    # ------------------------------------------------------------------ predict (:192-240)
    def predict(self, state):
        if self.phase is None or self.device is None:
            raise AttributeError('Phase, device attributes have to be set!')
        if self.phase == 'train' and self.epsilon is None:
            raise AttributeError('Epsilon attribute has to be set in training phase')

        if self.reach_destination(state):
            return ActionXY(0, 0) if self.kinematics == 'holonomic' else ActionRot(0, 0)
        if self.action_space is None:
            self.build_action_space(state.robot_state.v_pref)

        probability = np.random.random()
        if self.phase == 'train' and probability < self.epsilon:
            max_action = self.action_space[np.random.choice(len(self.action_space))]
        else:
            robot, humans = joint_state_to_tensors(state, self.device)
            if self.use_cuda_graphs:
                best, det = self.predict_batch_graphed(robot, humans)
            else:
                best, det = self.predict_batch(robot, humans, return_details=True)
            b = int(best[0])
            if b < 0:
                raise ValueError('Value network is not well trained.')
            max_action = self.action_space[b]
            if self.phase != 'train':
                self.traj = self._build_traj(robot, humans, det, b)

        if self.phase == 'train':
            self.last_state = self.transform(state)
        return max_action

    def predict_many(self, states):
        """Vectorised front end: predict() for MANY environments stepped together -- what
        Explorer.run_k_episodes / Robot.act (crowd_nav/utils/explorer.py:42-53, crowd_sim/envs/utils/robot.py:9-15) do one
        JointState at a time.  Same per-state semantics as predict(): reach_destination short-circuit, epsilon-greedy in
        the train phase (one np.random draw per state, in order), greedy look-ahead otherwise -- but every greedy state goes
        through ONE batched look-ahead (predict_batch).  Returns the list of actions; in the train phase `last_state`
        becomes the list of transformed states (one per environment)."""
        if self.phase is None or self.device is None:
            raise AttributeError('Phase, device attributes have to be set!')
        if self.phase == 'train' and self.epsilon is None:
            raise AttributeError('Epsilon attribute has to be set in training phase')
        if self.action_space is None:
            self.build_action_space(states[0].robot_state.v_pref)
        stop = ActionXY(0, 0) if self.kinematics == 'holonomic' else ActionRot(0, 0)
        actions = [None] * len(states)
        greedy = []
        for i, st in enumerate(states):
            if self.reach_destination(st):
                actions[i] = stop
                continue
            if self.phase == 'train' and np.random.random() < self.epsilon:
                actions[i] = self.action_space[np.random.choice(len(self.action_space))]
            else:
                greedy.append(i)
        if greedy:
            robot, humans = joint_states_to_tensors([states[i] for i in greedy], self.device)
            best = self.predict_batch(robot, humans).cpu().tolist()
            for i, b in zip(greedy, best):
                if b < 0:
                    raise ValueError('Value network is not well trained.')
                actions[i] = self.action_space[b]
        if self.phase == 'train':
            robot, humans = joint_states_to_tensors(states, self.device)
            self.last_state = [(robot[i], humans[i]) for i in range(len(states))]      # transform() of every state
        return actions

    def _build_traj(self, robot, humans, det, b):
        """[(state, action, reward), ...] along the best branch, as the reference's max_traj (:231, :295-300)."""
        k = int(det['best'][0])                       # position inside the evaluated action list
        traj = [((robot.clone(), humans.clone()), self.action_space[b], float(det['rew'][0, b]))]
        idx = b if det['acts'] is None else k         # row of the child in the level-1 batch
        # (the tensors are cloned: with CUDA-graph replay `det` aliases graph-private buffers the next predict() overwrites)
        state = (det['next_robot'][b:b + 1].clone(), det['next_humans'][0:1].clone())
        node = det['node']
        while node is not None:
            kb = int(node['best'][idx])
            a = kb if node['acts'] is None else int(node['acts'][idx, kb])
            traj.append((state, self.action_space[a], float(node['rew'][idx, kb])))
            w = node['w']
            state = (node['child_robot'][idx * w + kb: idx * w + kb + 1].clone(), node['child_humans'][idx: idx + 1].clone())
            idx = idx * w + kb
            node = node['child']
        traj.append((state, None, None))
        return traj

    # ------------------------------------------------------------------ reference-shaped helpers
    def action_clip(self, state, action_space, width, depth=1):
        """Reference signature (:242): state = (robot[1,1,9], humans[1,Nh,5]) -> list of `width` actions."""
        if self.action_space is None:
            self.build_action_space(self.v_pref)
        self._action_table()
        with torch.no_grad():
            acts = self._clip(state[0], state[1], 1, width)['acts']
        return [action_space[int(i)] for i in acts[0]]

    def V_planning(self, state, depth, width):
        """Reference signature (:271): returns (value[1,1], trajectory placeholder)."""
        if self.action_space is None:
            self.build_action_space(self.v_pref)
        self._action_table()
        with torch.no_grad():
            ret, _ = self._V_planning(state[0], state[1], 1, depth, width)
        return ret.view(1, 1), [(state, None, None)]

    def estimate_reward(self, state, action):
        """Reference signature (:304): one (state, action) pair -> python float (evaluated on the GPU)."""
        if isinstance(state, (list, tuple)):
            robot, humans = state
        else:
            robot, humans = joint_state_to_tensors(state, self.device)
        act = torch.tensor([[action[0], action[1]]], dtype=torch.float64, device=robot.device)      # ActionXY (vx, vy) | ActionRot (v, r)
        _, rew = ops.plan_expand(robot, humans, act, self.time_step, want_next=False, kinematics=self.kinematics)
        return float(rew[0])

    def transform(self, state):
        """JointState -> (robot[1,9], humans[Nh,5]) tensors on self.device (:359-370)."""
        robot = torch.Tensor([state.robot_state.to_tuple()]).to(self.device)
        humans = torch.Tensor([h.to_tuple() for h in state.human_states]).to(self.device)
        return robot, humans
