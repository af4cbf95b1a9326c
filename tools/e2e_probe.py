"""Host-storage rollout: per-call time of consecutive short calls (where does a 20-step call spend its time?)."""
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
for dev_slabs, host_slabs, warm in ((4, 8, 5), (4, 8, 8), (8, 8, 8), (4, 4, 5)):
    sim = BatchSimulator(tables, n, OBS_MODE_MANSY, REWARD_QOE, seed=0)
    roll = PolicyRollout(sim, policy, dev_slabs, seed=1234)
    host = roll.make_host_buffers(host_slabs=host_slabs)
    roll.run_host(warm, host)
    torch.cuda.synchronize()
    out = []
    for steps in (20, 20, 20, 100, 20):
        t0 = time.perf_counter()
        roll.run_host(steps, host)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out.append(f"{steps}: {dt / steps * 1e6:.0f} us/step")
    print(f"device slabs {dev_slabs}, host slabs {host_slabs}, warm-up {warm}:  " + "   ".join(out), flush=True)
    sim.close()
