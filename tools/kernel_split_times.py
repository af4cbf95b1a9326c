"""Per-kernel durations of the two-kernel rollout step (events around every launch) by env count and policy split."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import workload_tables
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE
from mansy_immersivevideostreaming_b200.policy import PolicyNet, mansy_state_dict_shapes, seeded_state_dict
from mansy_immersivevideostreaming_b200.rollout import PolicyRollout
from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler

sizes = [int(x) for x in sys.argv[1:]] or [4096, 8192]
shapes = mansy_state_dict_shapes()
policy = PolicyNet(seeded_state_dict(shapes[0], 1), seeded_state_dict(shapes[1], 2), OBS_MODE_MANSY)
tiler = ViewportTiler(device=0)
for n in sizes:
    tables = workload_tables(tiler.chunk_masks, n)
    sim = BatchSimulator(tables, n, OBS_MODE_MANSY, REWARD_QOE, seed=0, worker_num=n)
    slabs = max(4, -(-(320 << 20) // (n * sim.obs_stride * 4)))
    roll = PolicyRollout(sim, policy, slabs, seed=1234)
    roll.reserve_timing(200)
    for split in (1, 2, 4):
        try:
            policy.set_tc_split(split)
        except Exception:      # noqa: BLE001
            continue
        roll.run(20, timed=True)
        torch.cuda.synchronize()
        roll.run(200, timed=True)
        torch.cuda.synchronize()
        pm, sm, k = roll.kernel_ms()
        print(f"{n:6d} envs split {split}: policy {pm / k * 1e3:6.2f} us, step {sm / k * 1e3:6.2f} us", flush=True)
    sim.close()
    del roll
    torch.cuda.empty_cache()
