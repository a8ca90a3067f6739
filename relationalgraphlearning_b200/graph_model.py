"""RGL relational graph encoder -- drop-in for crowd_nav/policy/graph_model.py:10-130.

Same constructor signature, same parameter names/shapes (w_r.*, w_h.*, w_a, Ws.*), same forward
contract `forward((robot[B,1,9], humans[B,Nh,5])) -> H[B,Nh+1,32]`; the math runs in the fused
sm_100a kernel (csrc/graph_forward.cu) through the C ABI.
"""
import logging

import torch
import torch.nn as nn
from torch.nn import Parameter

from . import _lib, ops
from . import _torch_math as TM
from .helpers import mlp


class RGL(nn.Module):
    def __init__(self, config, robot_state_dim=9, human_state_dim=5):
        super().__init__()
        g = config.gcn
        self.multiagent_training = g.multiagent_training
        self.similarity_function = g.similarity_function
        self.robot_state_dim = robot_state_dim
        self.human_state_dim = human_state_dim
        self.num_layer = g.num_layer
        self.X_dim = g.X_dim
        self.layerwise_graph = g.layerwise_graph
        self.skip_connection = g.skip_connection
        self._wr_dims, self._wh_dims, self._final_dim = list(g.wr_dims), list(g.wh_dims), g.final_state_dim

        logging.info('Similarity_func: {}'.format(self.similarity_function))
        logging.info('Layerwise_graph: {}'.format(self.layerwise_graph))
        logging.info('Skip_connection: {}'.format(self.skip_connection))
        logging.info('Number of layers: {}'.format(self.num_layer))

        # parameter creation order matches the reference so torch.manual_seed(s) gives identical weights
        self.w_r = mlp(robot_state_dim, self._wr_dims, last_relu=True)
        self.w_h = mlp(human_state_dim, self._wh_dims, last_relu=True)
        if self.similarity_function == 'embedded_gaussian':
            self.w_a = Parameter(torch.randn(self.X_dim, self.X_dim))
        elif self.similarity_function == 'concatenation':
            self.w_a = mlp(2 * self.X_dim, [2 * self.X_dim, 1], last_relu=True)
        self.Ws = nn.ParameterList()
        for i in range(self.num_layer):
            out_dim = self._final_dim if (i == self.num_layer - 1 and i != 0) else self.X_dim
            self.Ws.append(Parameter(torch.randn(self.X_dim, out_dim)))

        self._pack_cache = ops._PackCache()
        self._A_dev = None
        self._A_host = None
        # numerics switch: large inference batches run their shared-weight GEMMs on the tensor cores with a 3xTF32 split
        # (~3e-6 relative to fp32); set True to keep everything on the fp32 FMA pipe (~3e-7)
        self.fp32_fma = False

    # ---- `.A`: attention matrix of the first sample, for visualisation (graph_model.py:116).  The
    # reference copies it to the host on every forward; here the copy happens when it is read.
    @property
    def A(self):
        if self._A_host is None and self._A_dev is not None:
            self._A_host = self._A_dev.detach().cpu().numpy()
        return self._A_host

    @A.setter
    def A(self, value):
        self._A_host, self._A_dev = value, None

    def __deepcopy__(self, memo):
        # nn.Module deepcopy copies __dict__; give the copy its own cache objects
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ('_pack_cache', '_A_dev', '_A_host'):   # fp32_fma is a plain bool and is copied
                continue
            setattr(new, k, copy.deepcopy(v, memo))
        new._pack_cache = ops._PackCache()
        new._A_dev = None
        new._A_host = None
        return new

    def mark_dirty(self):
        """Call after writing parameters through `.data` (not seen by the version counters)."""
        self._pack_cache.mark_dirty()

    def kernel_supported(self):
        return (self.similarity_function == 'embedded_gaussian' and self.X_dim == 32 and self._wr_dims == [64, 32]
                and self._wh_dims == [64, 32] and self._final_dim == 32 and self.robot_state_dim == 9
                and self.human_state_dim == 5 and 1 <= self.num_layer <= _lib.MAX_LAYERS)

    def flags(self):
        return ((_lib.FLAG_SKIP if self.skip_connection else 0) | (_lib.FLAG_LAYERWISE if self.layerwise_graph else 0)
                | (_lib.FLAG_FP32_FMA if self.fp32_fma else 0))

    def param_tensors(self):
        return ops._graph_param_list(self)

    def run(self, robot, humans, humans_bcast=1, motion_blob=None, want_H=False, want_E=False, want_S=False, throughput=False, out_H=None):
        """Kernel launch (no autograd).  Records the device-side attention of sample 0 for `.A`."""
        want_A0 = not self.layerwise_graph
        out = ops.graph_forward_raw(ops.packed_graph(self), self.num_layer,
                                    self.flags() | (_lib.FLAG_THROUGHPUT if throughput else 0), robot, humans,
                                    humans_bcast=humans_bcast, mblob=motion_blob, want_H=want_H, want_E=want_E,
                                    want_S=want_S, want_A0=want_A0, out_H=out_H)
        if want_A0:
            self._A_dev, self._A_host = out['A0'], None
        return out

    def shape_supported(self, humans):
        """The fused kernels hold one state's nodes in one 128-row tile: 1..31 humans (include/rgl_b200.h RGL_MAX_HUMANS)."""
        return 1 <= humans.size(1) <= _lib.MAX_HUMANS

    def forward(self, state):
        robot, humans = state
        ops.require_cuda_or_cpu_module(self, robot, humans, 'RGL.forward')
        # torch-op statement of the same math (_torch_math.py), on the tensors' own device: CPU modules (the reference's
        # callers may keep a policy on the CPU, crowd_nav/train.py:82), the seven non-default similarity functions,
        # non-default layer widths and human counts outside the kernels' 1..31
        if not robot.is_cuda or not self.kernel_supported() or not self.shape_supported(humans):
            H, A = TM.graph_forward(self, robot, humans, return_A=True)
            if not self.layerwise_graph:
                self._A_dev, self._A_host = A[0].detach(), None
            return H
        if ops._needs_grad(self, robot, humans):
            return ops.fused_with_autograd(lambda: self.run(robot, humans, want_H=True)['H'],
                                           lambda: TM.graph_forward(self, robot, humans),
                                           self.param_tensors(), [robot, humans])
        return self.run(robot, humans, want_H=True)['H']
