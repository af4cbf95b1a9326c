"""Fused rollout kernel at several env counts against the two-kernel loop, with progress prints (a debugging aid: run under
`timeout`).  python tools/fused_smoke.py [n ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mansy_immersivevideostreaming_b200 import synth
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, OBS_MODE_SIMPLE, REWARD_QOE, SimConfig
from mansy_immersivevideostreaming_b200.policy import PolicyNet, mansy_state_dict_shapes, seeded_state_dict, simple_state_dict_shapes
from mansy_immersivevideostreaming_b200.rollout import PolicyRollout
from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler

sizes = [int(x) for x in sys.argv[1:]] or [1, 160, 4096, 8269]

# progress buffer (debug builds, -DMANSY_MBAR_WATCHDOG): host-mapped, dumped by a watchdog thread if a launch does not return
import threading
import time
from mansy_immersivevideostreaming_b200 import _capi
lib = _capi.load_library()
prog = torch.zeros(148 * 16, dtype=torch.int32).pin_memory()
have_prog = lib.mansy_debug_progress(prog.data_ptr()) == 0
beat = [time.time(), "start"]


def watchdog():
    while True:
        time.sleep(2)
        if time.time() - beat[0] > 25:
            print(f"WATCHDOG: no progress for 25 s in '{beat[1]}'", flush=True)
            if have_prog:
                p = prog.numpy().reshape(148, 16)
                for cta in range(148):
                    if p[cta].any():
                        print(f"  cta {cta:3d} (rank {cta % 4}): " + " ".join(f"{v >> 16}:{v & 0xFFFF}" for v in p[cta]), flush=True)
            os._exit(3)


threading.Thread(target=watchdog, daemon=True).start()


def mark(what):
    beat[0], beat[1] = time.time(), what

tables0 = synth.make_synthetic_tables(ViewportTiler(SimConfig()).chunk_masks, n_videos=3, n_users=4, n_traces=5, seed=21, trace_len_range=(40, 90))
for kind in (OBS_MODE_MANSY, OBS_MODE_SIMPLE):
    shapes = mansy_state_dict_shapes() if kind == OBS_MODE_MANSY else simple_state_dict_shapes()
    for n in sizes:
        tables = tables0.with_samples(synth.per_env_samples(tables0, n))
        rolls = []
        for _ in range(2):
            policy = PolicyNet(seeded_state_dict(shapes[0], 1), seeded_state_dict(shapes[1], 2), kind)
            rolls.append(PolicyRollout(BatchSimulator(tables, n, kind, REWARD_QOE, seed=3), policy, 4, seed=77))
        a, b = rolls
        mark(f"kind {kind} n {n} two-kernel")
        print(f"kind {kind} n {n}: two-kernel ...", flush=True)
        b.run(9, fused=False)
        torch.cuda.synchronize()
        mark(f"kind {kind} n {n} fused")
        print("   fused ...", flush=True)
        a.run(4)
        torch.cuda.synchronize()
        mark(f"kind {kind} n {n} fused continued")
        print("   fused (continued) ...", flush=True)
        a.run(5)
        torch.cuda.synchronize()
        same = all(torch.equal(getattr(a.buf, k), getattr(b.buf, k)) for k in ("obs", "actions", "reward", "done", "value", "logp"))
        print(f"   equal: {same}  stats equal: {torch.equal(a.sim.episode_stats(), b.sim.episode_stats())}  error flags {a.sim.error_flag()} {b.sim.error_flag()}", flush=True)
        if not same:
            for k in ("obs", "actions", "reward", "done", "value", "logp"):
                x, y = getattr(a.buf, k), getattr(b.buf, k)
                if not torch.equal(x, y):
                    d = (x != y)
                    idx = d.nonzero()[:5].tolist()
                    print(f"      {k}: {int(d.sum())} differing elements, first at {idx}", flush=True)
        for r in rolls:
            r.sim.close()
            r.policy.close()
print("fused_smoke done")
