"""Three launches of the fused rollout kernel on the bench workload (for `ncu -k regex:policy_tc4 -s 2 -c 1`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import workload_tables
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE
from mansy_immersivevideostreaming_b200.policy import PolicyNet, mansy_state_dict_shapes, seeded_state_dict
from mansy_immersivevideostreaming_b200.rollout import PolicyRollout
from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tables = workload_tables(ViewportTiler(device=0).chunk_masks, n)
sim = BatchSimulator(tables, n, OBS_MODE_MANSY, REWARD_QOE, seed=0, worker_num=n)
shapes = mansy_state_dict_shapes()
policy = PolicyNet(seeded_state_dict(shapes[0], 1), seeded_state_dict(shapes[1], 2), OBS_MODE_MANSY)
roll = PolicyRollout(sim, policy, 27, seed=1234)
for _ in range(3):
    roll.run(steps)
    torch.cuda.synchronize()
print("done")
