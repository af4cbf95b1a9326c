"""A/B of the host-storage rollout: actions handed to the host by a store kernel (zero copy) vs cudaMemcpyAsync."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload_tables  # noqa: E402
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE  # noqa: E402
from mansy_immersivevideostreaming_b200.policy import PolicyNet, mansy_state_dict_shapes, seeded_state_dict  # noqa: E402
from mansy_immersivevideostreaming_b200.rollout import PolicyRollout  # noqa: E402
from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler  # noqa: E402

n = 4096
tables = workload_tables(ViewportTiler().chunk_masks, n)
a, c = mansy_state_dict_shapes()
policy = PolicyNet(seeded_state_dict(a, 1), seeded_state_dict(c, 2), OBS_MODE_MANSY)
if "--after-fused" in sys.argv:          # what bench.py does before its e2e section
    sim0 = BatchSimulator(tables, n, OBS_MODE_MANSY, REWARD_QOE, seed=0)
    slabs = -(-(320 << 20) // (n * sim0.obs_stride * 4))
    roll0 = PolicyRollout(sim0, policy, slabs, seed=1234)
    if "--reserve" in sys.argv:
        roll0.reserve_timing(2000)
    roll0.run(50)
    roll0.run(2000)
    if "--timed" in sys.argv:
        roll0.run(2000, timed=True)
    torch.cuda.synchronize()
    print("fused rollout done", flush=True)
for zc in (False, True, False, True):
    sim = BatchSimulator(tables, n, OBS_MODE_MANSY, REWARD_QOE, seed=0)
    roll = PolicyRollout(sim, policy, 4, seed=1234)
    host = roll.make_host_buffers(host_slabs=8)
    roll.run_host(5, host, zero_copy=zc)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    roll.run_host(100, host, zero_copy=zc)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"zero_copy={zc}: {dt / 100 * 1e6:.1f} us per step, {n * 100 / dt:.3e} chunk-steps/s", flush=True)
    sim.close()
