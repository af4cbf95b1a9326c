"""Device timing of the MPC expert: decisions/s at horizon 1..4 for N environments (BASELINE configs[1] tables)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload_tables  # noqa: E402
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE  # noqa: E402
from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    tables = workload_tables(ViewportTiler().chunk_masks, n)
    sim = BatchSimulator(tables, n, OBS_MODE_MANSY, REWARD_QOE, seed=0)
    sim.reset()
    sim.rollout_random(10, seed=3)           # desynchronise trace positions / buffers
    for h in (1, 2, 3, 4):
        sim.expert_actions(h)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5 if h < 4 else 3
        e0.record()
        for _ in range(reps):
            sim.expert_actions(h)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"horizon {h}: {n} decisions in {ms:.3f} ms = {n / ms * 1e3:.3e} decisions/s, {n * 15 ** h / ms * 1e3:.3e} sequences/s", flush=True)


if __name__ == "__main__":
    main()
