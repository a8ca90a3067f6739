"""relationalgraphlearning_b200 -- B200-native (sm_100a) implementation of the RGL policy hot path:
graph-model forward, value / state-predictor heads and the batched look-ahead of the
model_predictive_rl planner, behind the reference's own module / policy API.
See DESIGN.md for the scope and INTEGRATION.md for how it plugs into the reference."""
from .config import Config, policy_config  # noqa: F401
from .gcn import GCN  # noqa: F401
from .graph_model import RGL  # noqa: F401
from .helpers import mlp  # noqa: F401
from .model_predictive_rl import ModelPredictiveRL  # noqa: F401
from .policy_factory import policy_factory  # noqa: F401
from .state_predictor import LinearStatePredictor, StatePredictor  # noqa: F401
from .value_estimator import ValueEstimator  # noqa: F401

__version__ = '0.1.0'
