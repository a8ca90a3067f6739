"""ValueEstimator -- drop-in for crowd_nav/policy/value_estimator.py:5-20.

forward(state) -> V[B,1] = value_network(graph_model(state)[:, 0, :]).  On the kernel path the graph
kernel emits only the robot row of the last layer (E[B,32]) and the value-head kernel consumes it.
"""
import torch.nn as nn

from . import _lib, ops
from . import _torch_math as TM
from .helpers import mlp


class ValueEstimator(nn.Module):
    def __init__(self, config, graph_model):
        super().__init__()
        self.graph_model = graph_model
        self._dims = list(config.model_predictive_rl.value_network_dims)
        self.value_network = mlp(config.gcn.X_dim, self._dims)
        self._pack_cache = ops._PackCache()

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k not in ('_pack_cache', '_grad_sink'):
                setattr(new, k, copy.deepcopy(v, memo))
        new._pack_cache = ops._PackCache()
        return new

    def _train_params(self):
        return self.graph_model.param_tensors() + list(self.value_network.parameters())

    def kernel_supported(self):
        g = self.graph_model
        return hasattr(g, 'kernel_supported') and g.kernel_supported() and self._dims == [32, 100, 100, 1]

    def run(self, robot, humans, humans_bcast=1, throughput=False):
        """Two launches (graph kernel, value head), no autograd."""
        E = self.graph_model.run(robot, humans, humans_bcast=humans_bcast, want_E=True, throughput=throughput)['E']
        return ops.value_head_raw(ops.packed_value(self.value_network, self._pack_cache), E)

    def forward(self, state):
        assert len(state[0].shape) == 3
        assert len(state[1].shape) == 3
        robot, humans = state
        ops.require_cuda_or_cpu_module(self, robot, humans, 'ValueEstimator.forward')
        if not robot.is_cuda or not self.kernel_supported() or not self.graph_model.shape_supported(humans):
            return self.value_network(self.graph_model(state)[:, 0, :])
        if ops._needs_grad(self, robot, humans):
            from . import training
            if training.native_supported(self) and not (robot.requires_grad or humans.requires_grad):
                return training.value_forward_train(self, robot, humans)      # fused forward + hand-written backward
            params = self.graph_model.param_tensors() + list(self.value_network.parameters())
            return ops.fused_with_autograd(lambda: self.run(robot, humans),
                                           lambda: self.value_network(TM.graph_forward(self.graph_model, robot, humans)[:, 0, :]),
                                           params, [robot, humans])
        return self.run(robot, humans)
