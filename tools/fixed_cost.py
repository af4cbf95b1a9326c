"""Where the per-rollout fixed cost of the bench region goes: the region of bench.py (start event, ONE fused launch of K steps,
the per-rollout gather, stop event) for a range of K -> intercept / slope, with and without the gather, and with the GPU kept
busy by a spin kernel while the host enqueues (launch latency hidden).  Stands in for an nsys trace (not installed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import workload_tables
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE
from mansy_immersivevideostreaming_b200.policy import PolicyNet, mansy_state_dict_shapes, seeded_state_dict
from mansy_immersivevideostreaming_b200.rollout import PeerGroup, PolicyRollout, gather_episode_stats
from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler

n = 4096
tables = workload_tables(ViewportTiler(device=0).chunk_masks, n)
sim = BatchSimulator(tables, n, OBS_MODE_MANSY, REWARD_QOE, seed=0, worker_num=n)
shapes = mansy_state_dict_shapes()
policy = PolicyNet(seeded_state_dict(shapes[0], 1), seeded_state_dict(shapes[1], 2), OBS_MODE_MANSY)
roll = PolicyRollout(sim, policy, 27, seed=1234)
peers = PeerGroup(n, 0)
roll.run(5)
gather_episode_stats(sim, peers=peers)
torch.cuda.synchronize()
Ks = [1, 2, 5, 10, 20, 40, 80, 160]


def region(K, gather, busy):
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if busy:
            torch.cuda._sleep(200_000)          # ~100 us of GPU work queued ahead: the launches below are enqueued meanwhile
        e0.record()
        roll.run(K)
        if gather:
            gather_episode_stats(sim, peers=peers)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3)
    return best


for gather, busy in ((True, False), (False, False), (True, True), (False, True)):
    us = [region(K, gather, busy) for K in Ks]
    slope, icpt = np.polyfit(Ks[2:], us[2:], 1)
    print(f"gather={gather!s:5} launches enqueued behind queued GPU work={busy!s:5}: " + "  ".join(f"K={k}: {u:7.1f}" for k, u in zip(Ks, us)))
    print(f"    fit over K >= 5: {slope:6.2f} us per step + {icpt:6.1f} us per rollout", flush=True)
