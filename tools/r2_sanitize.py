"""One small value-net training step (tcgen05 forward with saves, staged attention / similarity backward, fused embedding-MLP
backward, mma.sync and -- with RGL_BWD_VARIANT=t -- tcgen05 linear backward) for compute-sanitizer runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relationalgraphlearning_b200.config import policy_config
from relationalgraphlearning_b200.graph_model import RGL
from relationalgraphlearning_b200.synthetic import synthetic_states
from relationalgraphlearning_b200.value_estimator import ValueEstimator

dev = torch.device('cuda:0')
for nh, B in ((10, 300), (5, 130)):
    cfg = policy_config()
    torch.manual_seed(3)
    ve = ValueEstimator(cfg, RGL(cfg, 9, 5)).to(dev)
    robot, humans = synthetic_states(B, nh, seed=B)
    out = ve((robot.to(dev), humans.to(dev)))
    loss = torch.nn.functional.mse_loss(out, torch.zeros_like(out))
    loss.backward()
    torch.cuda.synchronize()
    print('nh', nh, 'B', B, 'loss %.6f' % float(loss), 'grad norm %.6f' % float(sum(p.grad.norm() for p in ve.parameters())))
