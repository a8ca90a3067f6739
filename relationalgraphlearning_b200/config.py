"""Attribute-bag configs for the RGL hot path.

The reference loads a python file of config classes with importlib
(crowd_nav/train.py:66-70) and hands `PolicyConfig()` to `policy.configure`.
Only attribute access is used (`config.gcn.X_dim`, `config.model_predictive_rl.planning_depth`, ...),
so any object with the same attribute tree works.  `policy_config()` builds the tree the shipped
`mp_*` configs produce (crowd_nav/configs/icra_benchmark/config.py:59-114 and mp_separate.py:9-28)
so that tests / bench on a box without the reference checkout have identical hyper-parameters.
"""
import math


class Config(object):
    """Plain attribute bag (same role as crowd_nav/configs/icra_benchmark/config.py:9-11)."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def __repr__(self):
        return 'Config(%s)' % ', '.join('%s=%r' % kv for kv in sorted(vars(self).items()))


def policy_config(num_layer=2, X_dim=32, similarity_function='embedded_gaussian', layerwise_graph=False,
                  skip_connection=True, planning_depth=1, planning_width=1, do_action_clip=False,
                  share_graph_model=False, linear_state_predictor=False, sparse_search=None,
                  speed_samples=5, rotation_samples=16, kinematics='holonomic', gamma=0.9):
    """PolicyConfig equivalent of the shipped model_predictive_rl configs.

    Defaults = mp_separate.py:9-28 on top of BasePolicyConfig (config.py:59-114).
    `mp_separate_dp.py` = planning_depth=2, planning_width=2, do_action_clip=True.
    """
    c = Config()
    c.name = 'model_predictive_rl'
    c.rl = Config(gamma=gamma)
    c.action_space = Config(kinematics=kinematics, speed_samples=speed_samples,
                            rotation_samples=rotation_samples, sampling='exponential', query_env=False,
                            rotation_constraint=math.pi / 3)
    c.gcn = Config(multiagent_training=True, num_layer=num_layer, X_dim=X_dim,
                   wr_dims=[64, X_dim], wh_dims=[64, X_dim], final_state_dim=X_dim,
                   similarity_function=similarity_function, layerwise_graph=layerwise_graph,
                   skip_connection=skip_connection)
    c.model_predictive_rl = Config(linear_state_predictor=linear_state_predictor,
                                   planning_depth=planning_depth, planning_width=planning_width,
                                   do_action_clip=do_action_clip, motion_predictor_dims=[64, 5],
                                   value_network_dims=[32, 100, 100, 1],
                                   share_graph_model=share_graph_model)
    if sparse_search is not None:
        c.model_predictive_rl.sparse_search = sparse_search
    return c
