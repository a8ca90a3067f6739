"""ctypes binding of librgl_b200.so -- the only way the Python modules reach the sm_100a kernels.

There is deliberately NO fallback: if the library is missing or a call fails, an exception is raised
(the product path must fail loudly rather than silently compute somewhere else).
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'librgl_b200.so')

MAX_LAYERS = 4
MAX_HUMANS = 31
FLAG_SKIP = 1
FLAG_LAYERWISE = 2
FLAG_THROUGHPUT = 4
FLAG_FP32_FMA = 8
FLAG_TRAIN_TC = 16
KIN_HOLONOMIC = 0
KIN_UNICYCLE = 1

c_float_p = ctypes.c_void_p


class GraphParams(ctypes.Structure):
    _fields_ = [('wr0_w', c_float_p), ('wr0_b', c_float_p), ('wr1_w', c_float_p), ('wr1_b', c_float_p),
                ('wh0_w', c_float_p), ('wh0_b', c_float_p), ('wh1_w', c_float_p), ('wh1_b', c_float_p),
                ('w_a', c_float_p), ('Ws', c_float_p * MAX_LAYERS), ('num_layer', ctypes.c_int)]


class ValueParams(ctypes.Structure):
    _fields_ = [('w0', c_float_p), ('b0', c_float_p), ('w1', c_float_p), ('b1', c_float_p),
                ('w2', c_float_p), ('b2', c_float_p), ('w3', c_float_p), ('b3', c_float_p)]


class MotionParams(ctypes.Structure):
    _fields_ = [('w0', c_float_p), ('b0', c_float_p), ('w1', c_float_p), ('b1', c_float_p)]


class GraphSave(ctypes.Structure):
    _fields_ = [('a1r', c_float_p), ('a1h', c_float_p), ('X', c_float_p), ('Y', c_float_p), ('A', c_float_p),
                ('M', c_float_p * MAX_LAYERS), ('Rl', c_float_p * MAX_LAYERS), ('Hl', c_float_p * MAX_LAYERS), ('mh', c_float_p)]


class Rows(ctypes.Structure):
    _fields_ = [('ptr', c_float_p), ('ld', ctypes.c_int), ('rows_per_group', ctypes.c_int), ('group_stride', ctypes.c_longlong)]


EXPORTS = {
    'rgl_version': (ctypes.c_int, []),
    'rgl_last_error_string': (ctypes.c_char_p, []),
    'rgl_packed_graph_floats': (ctypes.c_size_t, [ctypes.c_int]),
    'rgl_packed_value_floats': (ctypes.c_size_t, []),
    'rgl_packed_motion_floats': (ctypes.c_size_t, []),
    'rgl_pack_graph': (ctypes.c_int, [ctypes.POINTER(GraphParams), c_float_p, ctypes.c_void_p]),
    'rgl_pack_value': (ctypes.c_int, [ctypes.POINTER(ValueParams), c_float_p, ctypes.c_void_p]),
    'rgl_pack_motion': (ctypes.c_int, [ctypes.POINTER(MotionParams), c_float_p, ctypes.c_void_p]),
    'rgl_graph_forward': (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         c_float_p, ctypes.c_int, ctypes.c_int, c_float_p,
                                         c_float_p, c_float_p, c_float_p, c_float_p, ctypes.c_void_p]),
    'rgl_value_head': (ctypes.c_int, [c_float_p, ctypes.c_int, c_float_p, c_float_p, ctypes.c_void_p]),
    'rgl_value_forward': (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         c_float_p, ctypes.c_int, ctypes.c_int, c_float_p, c_float_p, c_float_p,
                                         c_float_p, ctypes.c_void_p]),
    'rgl_gcn_layer': (ctypes.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, c_float_p, c_float_p, ctypes.c_void_p]),
    'rgl_graph_forward_train': (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_int, ctypes.c_int, c_float_p, ctypes.c_int,
                                               ctypes.c_int, c_float_p, ctypes.POINTER(GraphSave), c_float_p, c_float_p, c_float_p,
                                               ctypes.c_void_p]),
    'rgl_value_head_train': (ctypes.c_int, [c_float_p, ctypes.c_int, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                            ctypes.c_void_p]),
    'rgl_linear_bwd': (ctypes.c_int, [ctypes.POINTER(Rows), ctypes.c_int, ctypes.POINTER(Rows), ctypes.POINTER(Rows), ctypes.c_int,
                                      c_float_p, ctypes.c_int, ctypes.POINTER(Rows), ctypes.c_int, c_float_p, c_float_p,
                                      ctypes.c_int, ctypes.c_void_p]),
    'rgl_mlp2_bwd': (ctypes.c_int, [ctypes.POINTER(Rows), ctypes.POINTER(Rows), ctypes.POINTER(Rows), c_float_p, ctypes.POINTER(Rows),
                                    ctypes.c_int, c_float_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_void_p]),
    'rgl_attn_layer_bwd': (ctypes.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, c_float_p, c_float_p,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, c_float_p, ctypes.c_int, ctypes.c_void_p]),
    'rgl_attn_sim_bwd': (ctypes.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, c_float_p, c_float_p, c_float_p,
                                        c_float_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p]),
    'rgl_sim_bwd': (ctypes.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_void_p]),
    'rgl_plan_expand': (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_double, ctypes.c_int, c_float_p, c_float_p, ctypes.c_void_p]),
    'rgl_plan_argmax': (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                       c_float_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'rgl_plan_select': (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                       ctypes.c_void_p, c_float_p, ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_void_p]),
    'rgl_plan_backup': (ctypes.c_int, [c_float_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                       c_float_p, ctypes.c_void_p, ctypes.c_void_p]),
    'rgl_td_loss': (ctypes.c_int, [c_float_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_float, ctypes.c_float, c_float_p, c_float_p,
                                   ctypes.c_void_p]),
    # GPU-resident replay memory (csrc/replay.cu)
    'rgl_replay_record_floats': (ctypes.c_int, [ctypes.c_int]),
    'rgl_replay_push': (ctypes.c_int, [c_float_p, ctypes.c_longlong, ctypes.c_int, c_float_p, c_float_p, c_float_p, c_float_p,
                                       c_float_p, c_float_p, ctypes.c_void_p]),
    'rgl_replay_gather': (ctypes.c_int, [c_float_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_float_p, c_float_p, c_float_p,
                                         c_float_p, c_float_p, c_float_p, ctypes.c_void_p]),
    # data-parallel gradient exchange over NVLink peer memory (csrc/dp_comm.cu)
    'rgl_comm_create': (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.POINTER(ctypes.c_void_p)]),
    'rgl_comm_handle_bytes': (ctypes.c_int, []),
    'rgl_comm_ipc_handle': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'rgl_comm_open_peers': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'rgl_comm_accum_ptr': (ctypes.c_void_p, [ctypes.c_void_p]),
    'rgl_comm_allreduce': (ctypes.c_int, [ctypes.c_void_p, c_float_p, ctypes.c_float, ctypes.c_void_p]),
    'rgl_comm_status': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]),
    'rgl_comm_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'rgl_comm_last_error_string': (ctypes.c_char_p, []),
}

_lib = None


class RglError(RuntimeError):
    pass


def lib():
    """Load librgl_b200.so (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RglError('librgl_b200.so not found at %s; build it with '
                           '`python -m relationalgraphlearning_b200.build`' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)     # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what, comm=False):
    if rc != 0:
        msg = lib().rgl_comm_last_error_string() if comm else lib().rgl_last_error_string()
        raise RglError('%s failed (rc=%d): %s' % (what, rc, msg.decode() if msg else ''))


def ptr(t):
    """Device pointer of a contiguous fp32/fp64/int32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    # pinned host memory is device-addressable too (unified virtual addressing): zero-copy outputs
    assert (t.is_cuda or t.is_pinned()) and t.is_contiguous(), 'librgl_b200 takes contiguous CUDA (or pinned host) tensors'
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
