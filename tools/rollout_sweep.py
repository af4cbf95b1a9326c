"""Rollout step time (policy + sample + simulator step) by env count: fused cluster kernel vs two launches per step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import workload_tables
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE
from mansy_immersivevideostreaming_b200.policy import PolicyNet, mansy_state_dict_shapes, seeded_state_dict
from mansy_immersivevideostreaming_b200.rollout import PolicyRollout
from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler

sizes = [int(x) for x in sys.argv[1:]] or [4096, 8192, 16384, 32768]
shapes = mansy_state_dict_shapes()
policy = PolicyNet(seeded_state_dict(shapes[0], 1), seeded_state_dict(shapes[1], 2), OBS_MODE_MANSY)
tiler = ViewportTiler(device=0)
for n in sizes:
    tables = workload_tables(tiler.chunk_masks, n)
    sim = BatchSimulator(tables, n, OBS_MODE_MANSY, REWARD_QOE, seed=0, worker_num=n)
    slabs = max(4, -(-(320 << 20) // (n * sim.obs_stride * 4)))
    roll = PolicyRollout(sim, policy, slabs, seed=1234)
    out = []
    for fused, split in ((True, 0), (False, 0), (False, 1), (False, 4)):
        policy.set_tc_split(split)
        roll.run(30, fused=fused)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        roll.run(300, fused=fused)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 300 * 1e3
        out.append(f"{'fused' if fused else 'two-kernel split ' + str(split)} {us:6.1f} us = {n / us:5.1f} M/s")
    print(f"{n:6d} envs: " + "   ".join(out), flush=True)
    sim.close()
    del roll
    torch.cuda.empty_cache()
