"""Generate ``tests/golden/expert_kat.npz`` from the UNMODIFIED reference ``ExpertEnv`` -- build container only.

    python -m oracle.make_golden_expert

TEST INFRASTRUCTURE.  Writes a small synthetic dataset in the reference's on-disk formats, runs the reference's MPC
expert (bitrate_selection/envs/expert_env.py: ``reset`` / ``choose_action`` / ``step``) over whole episodes with
horizons 1..4 (4 = the reference's default, expert_env.py / run_expert.py: a handful of decisions, 50 625 Python roll-outs
each), and asserts that the restatement (``oracle.sim_oracle.expert_choose_action`` on ``OracleEnv``)
chooses the same action at every decision in the numeric chain the reference runs in under this container's numpy
(float32).  The fixture stores the float64-chain decisions (what the CUDA kernel follows, like the
step kernel) together with the float32-chain ones the reference produced, plus the tables.
"""
from __future__ import annotations

import os
import tempfile

import numpy as np

from mansy_immersivevideostreaming_b200 import synth
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE, SimConfig
from oracle import sim_oracle as so
from oracle.ref_loader import load_reference, silence_prints

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = SimConfig()


def main() -> None:
    ref = load_reference()
    assert ref.expert_env is not None, "envs.expert_env did not import"
    root = tempfile.mkdtemp(prefix="mansy_expert_")
    tables = synth.make_synthetic_tables(lambda g, p: so.chunk_masks(g, p, CFG), n_videos=2, n_users=2, n_traces=3, seed=11,
                                         trace_len_range=(40, 90), short_tail_frac=0.5)
    cfg_path = synth.write_reference_layout(tables, root)
    config = ref.common.get_config_from_yml(cfg_path)
    qoe_weights = [[float(x) for x in w] for w in tables.qoe_w]
    samples = [tuple(int(x) for x in s) for s in tables.samples[::5][:4]]
    tb = tables.with_samples(np.asarray(samples, dtype=np.int32))
    out = {}
    for horizon in (1, 2, 3, 4):
        ref.expert_env.ExpertEnv.init = False            # class-level cache flag (expert_env.py:18)
        with silence_prints():
            env = ref.expert_env.ExpertEnv(config, "Synth", "SynthNet", qoe_weights, samples, root,
                                           os.path.join(root, f"cache_{horizon}.pkl"), os.path.join(root, f"log_{horizon}.csv"),
                                           CFG.startup_download, horizon, True, "train", 0)
        o32 = so.OracleEnv(tb, OBS_MODE_MANSY, REWARD_QOE, "f32", worker_id=0, worker_num=1)
        o64 = so.OracleEnv(tb, OBS_MODE_MANSY, REWARD_QOE, "f64", worker_id=0, worker_num=1)
        acts32, acts64, vals64, eps = [], [], [], []
        n_ep = 4 if horizon < 3 else (2 if horizon == 3 else 1)
        max_decisions = 10 if horizon == 4 else 10 ** 9
        for ep in range(n_ep):
            with silence_prints():
                env.reset()
            o32.reset(); o64.reset()
            done = False
            while not done and len(acts32) < max_decisions:
                a_ref = int(env.choose_action())
                a32 = so.expert_choose_action(o32, horizon)
                a64, v64 = so.expert_choose_action(o64, horizon, return_value=True)
                assert a_ref == a32, (horizon, ep, a_ref, a32)
                acts32.append(a32); acts64.append(a64); vals64.append(v64); eps.append(ep)
                _, r_ref, done, _ = env.step(a_ref)
                _, r32, d32, _ = o32.step(a_ref)
                o64.step(a_ref)                               # teacher-forced with the reference's action
                assert float(r_ref) == float(r32) and done == d32
        agree = float(np.mean(np.asarray(acts32) == np.asarray(acts64)))
        print(f"horizon {horizon}: {len(acts32)} decisions equal to the reference (f32 chain); f64 chain agrees on {agree * 100:.1f}%")
        out[f"h{horizon}_actions_ref"] = np.asarray(acts32, np.int32)
        out[f"h{horizon}_actions_f64"] = np.asarray(acts64, np.int32)
        out[f"h{horizon}_value_f64"] = np.asarray(vals64, np.float64)
        out[f"h{horizon}_episode"] = np.asarray(eps, np.int32)
    out.update(tb.to_npz_dict())
    path = os.path.join(ROOT, "tests", "golden", "expert_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
