"""Policy registry mirroring crowd_nav/policy/policy_factory.py:9-13 for the path this package replaces."""
from .gcn import GCN
from .model_predictive_rl import ModelPredictiveRL

policy_factory = {'model_predictive_rl': ModelPredictiveRL, 'gcn': GCN}


def install_into_reference():
    """Swap the B200 implementation in under the reference's own registry key, so the reference's
    train.py / test.py (`policy_factory[policy_config.name]()`, train.py:87-88) build it unchanged."""
    from crowd_nav.policy.policy_factory import policy_factory as ref_factory
    ref_factory['model_predictive_rl'] = ModelPredictiveRL
    ref_factory['gcn'] = GCN
    return ref_factory
