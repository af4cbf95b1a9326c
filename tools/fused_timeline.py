"""In-kernel timeline of the fused rollout kernel (policy + sample + simulator step per 128-env cluster) on the
bench workload -- a profiling aid, run on a GPU box:  python tools/fused_timeline.py [n_envs] [cta ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import workload_tables
from mansy_immersivevideostreaming_b200._capi import check
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE
from mansy_immersivevideostreaming_b200.policy import PolicyNet, mansy_state_dict_shapes, seeded_state_dict
from mansy_immersivevideostreaming_b200.rollout import PolicyRollout
from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctas = [int(x) for x in sys.argv[2:]] or [0, 1, 2, 3]
tables = workload_tables(ViewportTiler(device=0).chunk_masks, n)
sim = BatchSimulator(tables, n, OBS_MODE_MANSY, REWARD_QOE, seed=0, worker_num=n)
shapes = mansy_state_dict_shapes()
policy = PolicyNet(seeded_state_dict(shapes[0], 1), seeded_state_dict(shapes[1], 2), OBS_MODE_MANSY)
roll = PolicyRollout(sim, policy, 27, seed=1234)
roll.run(20)
torch.cuda.synchronize()
tl = torch.zeros(512, dtype=torch.int64, device="cuda")
names = {493: "sim loads issued (state, slot, prefetch)", 494: "bar.sync 1 (partials published)", 495: "logits summed",
         502: "weights gathered + summed",
         489: "step begin", 480: "partial D2 done", 482: "partials stored (L2 exchange)", 483: "cluster sync 1",
         484: "hidden slice + heads, D3 pushed", 485: "cluster sync 2", 486: "own 32 rows finished (sample)", 490: "state loaded",
         491: "step_env done", 492: "observation row + state stored", 487: "simulator phase done (fences)", 488: "cluster sync 3"}
for cta in ctas:
    tl.zero_()
    check(sim.lib.mansy_debug_fused_timeline(tl.data_ptr(), cta))
    torch.cuda.synchronize()
    time.sleep(0.02)          # an idle GPU before the launch, as in bench.py's timed region
    roll.run(4)
    torch.cuda.synchronize()
    t = tl.cpu().numpy()
    t0 = t[489]
    which = os.environ.get("MANSY_TC_TIMELINE_STEP")
    print(f"--- CTA {cta} (cluster rank {cta % 4}), {'step ' + which if which else 'second rollout step'} of the launch, cycles since step begin"
          f" (kernel prologue ended {t0 - t[511]} cycles before it)")
    for j in range(40):
        if t[j] == 0:
            break
        print(f"  job {j:2d}: TMA issue {t[j]-t0:6d}  operands {t[128+j]-t0:6d}  MMAs issued {t[256+j]-t0:6d}")
    for key in (480, 482, 483, 484, 485, 493, 486, 490, 491, 492, 487, 488):
        print(f"  {names[key]:34s} {t[key]-t0:7d}")
    if t[496]:      # -DMANSY_STEP_PROFILE build: inside step_env (thread 256 of CTA 0, last stamped step)
        for j, nm in enumerate(["entry", "gathers issued+summed (own)", "group sums done", "trace walk done", "qoe done", "history slot done"]):
            print(f"  step_env {nm:30s} {t[496+j]-t[496]:7d}")
check(sim.lib.mansy_debug_fused_timeline(None, 0))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
roll.run(2000)
e1.record()
torch.cuda.synchronize()
print("fused: us per rollout step:", e0.elapsed_time(e1) / 2000 * 1e3)
