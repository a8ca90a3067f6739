"""Run the graph kernel a few times at a given batch (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relationalgraphlearning_b200.config import policy_config
from relationalgraphlearning_b200.graph_model import RGL
from relationalgraphlearning_b200.synthetic import synthetic_states
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
nh = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device('cuda:0')
torch.manual_seed(0)
g1 = RGL(policy_config(), 9, 5).to(dev)
robot, humans = synthetic_states(min(B, 65536), nh, seed=1, device=dev)
if B > 65536:
    robot = robot.repeat(B // 65536, 1, 1); humans = humans.repeat(B // 65536, 1, 1)
with torch.no_grad():
    for _ in range(6):
        g1.run(robot, humans, want_H=True)
torch.cuda.synchronize()
