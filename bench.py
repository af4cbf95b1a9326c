#!/usr/bin/env python
"""Benchmark of the streaming-simulator hot path (BASELINE.json: "simulated chunk-steps/sec").

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm (CPU port, host cores)

Workload at every N (weak scaling): BASELINE.json configs[1] -- 4,096 parallel MANSY envs per GPU,
PPO rollout = policy forward + Categorical sample + lock-step simulator step, observations written
straight into the rollout buffer the learner reads.  One "step" = one lock-step chunk-step of all
envs of the job.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "simulated chunk-steps/sec"
UNIT = "chunk-steps/s"
ENVS_PER_GPU = 4096
# algorithmic HBM bytes per MANSY chunk-step with materialised observation
# dram__bytes_read.sum + dram__bytes_write.sum of ONE 40-step launch of the fused kernel at 4096 envs, per rollout
# step (ncu --set full, profiles/r02zz_fused_kernel_40steps_ncu.txt: 47.9 MB read + 486.5 MB written / 40; r02p: 14.43 MB)
FUSED_DRAM_BYTES_PER_STEP_4096 = 13_358_572
# the same for one launch of step_kernel<MANSY> at 1 048 576 envs (profiles/r02p_step_kernel_1048576_ncu.txt: 413.6 MB read
# + 3 433.1 MB written; algorithmic 3 513 B x 1 048 576 = 3 683.6 MB)
STEP_DRAM_BYTES_MANSY_1M = 3_846_669_864
BYTES_PER_STEP_MANSY = 3513
FLOP_PER_STEP_POLICY = 2 * 425_472        # 0.851 MFLOP, shared FeatureNet evaluated once


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=2000)
    p.add_argument("--warmup", type=int, default=50)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--diverse-qoe", action="store_true",
                   help="BASELINE configs[2]: one Dirichlet(1,1,1)*9 QoE preference vector per environment (65,536 envs = "
                        "--envs-per-gpu 8192 on 8 GPUs) instead of the 4 default weights")
    p.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    p.add_argument("--no-extras", action="store_true",
                   help="only the headline rollout line: skip configs[2] (8,192 diverse-QoE envs per GPU), configs[3] (simulator-only "
                        "sweep at 65,536 / 1,048,576 envs) and configs[4] (MTIO masks feeding the environments)")
    p.add_argument("--nccl-gather", action="store_true",
                   help="exchange the episode totals with torch.distributed (NCCL) instead of the NVLink peer-memory kernel")
    p.add_argument("--rank-times", action="store_true",
                   help="diagnostic: also record when each rank's rollout kernel ended (an extra event inside the timed region)")
    p.add_argument("--workload", default="rollout", choices=["rollout", "mtio"],
                   help="rollout = BASELINE configs[1] (the bench line the driver records); mtio = BASELINE configs[4], the "
                        "MTIO viewport-prediction inference feeding predicted tile masks to the environments")
    p.add_argument("--mtio-samples", type=int, default=16384)
    p.add_argument("--no-mtio", action="store_true", help="leave the viewport_prediction section out of the rollout line")
    p.add_argument("--profile-sim", type=int, default=0,
                   help="only run a few simulator-only launches at this env count (for ncu captures); prints no bench line")
    return p.parse_args()


# ---------------------------------------------------------------------------------------------
# shared: synthetic workload
# ---------------------------------------------------------------------------------------------
def workload_tables(mask_fn, n_slots, diverse_qoe=False):
    from mansy_immersivevideostreaming_b200 import synth
    t = synth.make_synthetic_tables(mask_fn, n_videos=24, n_users=60, n_chunks=60, n_traces=40, seed=20260101)
    if diverse_qoe:      # Q = number of environments, weights ~ Dirichlet(1,1,1) * 9; environment k uses weight k
        import numpy as np
        t = t.with_samples(t.samples, qoe_w=synth.diverse_qoe_weights(n_slots))
        s = synth.per_env_samples(t, n_slots)
        s[:, 3] = np.arange(n_slots, dtype=s.dtype)
        return t.with_samples(s)
    return t.with_samples(synth.per_env_samples(t, n_slots))


# ---------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (oracle/_ref archive or /root/reference) on all host cores
# ---------------------------------------------------------------------------------------------
def _ref_worker(args):
    """One process = one reference environment driven exactly like the reference's own loop (run_mansy.py:161-175):
    the reference's MANSYEnv (unmodified file) stepped with actions sampled from the reference's Actor (models/mansy.py,
    torch on the CPU, batch of 1, Categorical(logits) as run_mansy.py:228-229), resets included.  Returns per interval
    (steps, seconds) for `intervals` back-to-back intervals of `seconds` each."""
    worker, n_workers, cfg_path, intervals, seconds, with_policy, kind = args
    import numpy as np
    out = []
    if kind == "reference":
        import torch
        from torch.distributions import Categorical
        torch.set_num_threads(1)
        from oracle.ref_loader import load_reference, silence_prints
        from mansy_immersivevideostreaming_b200.policy import mansy_state_dict_shapes, seeded_state_dict
        ref = load_reference()
        config = ref.common.get_config_from_yml(cfg_path)
        qoe = config.qoe_split["train"]
        log = os.path.join(os.path.dirname(cfg_path), f"train_log_{worker}.csv")
        with silence_prints():
            env = ref.mansy_env.MANSYEnv(config, "Synth", "SynthNet", qoe, None, 0.5, log, config.startup_download,
                                         mode="train", seed=worker, worker_num=n_workers, device="cpu")
        env.seed(worker)
        M = ref.models_mansy
        fn = M.FeatureNet(config.past_k, config.tile_total_num, len(config.video_rates), 128, device="cpu")
        actor = M.Actor(fn, 1280, 128, config.action_space, "cpu")
        shapes, _ = mansy_state_dict_shapes()
        actor.load_state_dict({k: torch.from_numpy(v) for k, v in seeded_state_dict(shapes, 1).items()})
        torch.manual_seed(1234 + worker)
        rng = np.random.default_rng(worker)
        state = env.reset()
        with torch.no_grad():
            for _ in range(intervals):
                steps, t0 = 0, time.perf_counter()
                while time.perf_counter() - t0 < seconds:
                    for _ in range(10):
                        if with_policy:
                            for key, value in state.items():                 # run_mansy.py:167-168
                                state[key] = np.expand_dims(value, 0)
                            logits, _ = actor(state)
                            action = Categorical(logits=logits).sample().item()
                        else:
                            action = int(rng.integers(0, 15))
                        state, _, done, _ = env.step(action)
                        steps += 1
                        if done:
                            state = env.reset()
                out.append((steps, time.perf_counter() - t0))
        return out
    # fallback when no reference archive travelled with the repo: the oracle port (restatement) of the simulator
    from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE
    from mansy_immersivevideostreaming_b200.synth import synthetic_actions
    from mansy_immersivevideostreaming_b200.tables import SimTables
    from oracle import sim_oracle as so
    tables = SimTables.from_npz_dict(np.load(cfg_path))
    env = so.OracleEnv(tables, OBS_MODE_MANSY, REWARD_QOE, "f64", worker_id=worker, worker_num=n_workers)
    env.reset()
    actions = [int(a) for a in synthetic_actions(8192, worker, seed=1234)]
    total = 0
    for _ in range(intervals):
        steps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            for _ in range(50):
                _, _, done, _ = env.step(actions[total & 8191])
                steps += 1; total += 1
                if done:
                    env.reset()
        out.append((steps, time.perf_counter() - t0))
    return out


def cpu_arm(intervals: int, seconds: float, cores: int, with_policy: bool = True):
    """Chunk-steps/s of the reference on `cores` processes (one environment each) for `intervals` intervals.
    Returns (list of per-interval aggregate rates, description dict)."""
    import multiprocessing as mp
    import numpy as np
    from oracle import sim_oracle as so
    from oracle.ref_loader import code_available
    from mansy_immersivevideostreaming_b200 import synth
    from mansy_immersivevideostreaming_b200.config import SimConfig
    cfg = SimConfig()
    # a slice of the bench workload (same generators and seed): 4 videos x 8 users x 40 traces, 4 default QoE weights
    t = synth.make_synthetic_tables(lambda g, p: so.chunk_masks(g, p, cfg), n_videos=4, n_users=8, n_chunks=60,
                                    n_traces=40, seed=20260101)
    root = tempfile.mkdtemp(prefix="mansy_cpu_")
    kind = "reference" if code_available() else "port"
    if kind == "reference":
        path = synth.write_reference_layout(t, root)          # the reference's own on-disk formats + config.yml
    else:
        t = t.with_samples(synth.per_env_samples(t, max(cores, 64)))
        path = os.path.join(root, "tables.npz")
        np.savez(path, **t.to_npz_dict())
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_ref_worker, [(w, cores, path, intervals, seconds, with_policy, kind) for w in range(cores)])
    wall = time.perf_counter() - t0
    rates = [sum(r[i][0] / r[i][1] for r in res) for i in range(intervals)]
    steps = [sum(r[i][0] for r in res) for i in range(intervals)]
    what = ("the UNMODIFIED reference classes (oracle/_ref archive): MANSYEnv.step/reset (envs/mansy_env.py, resets re-read the "
            "manifest / viewport / trace files as simulators/simulator.py:30-38 does)"
            + (" + Actor forward and Categorical sample per step (models/mansy.py, torch CPU, batch 1: the loop of "
               "run_mansy.py:161-175)" if with_policy else " with uniform random actions (simulator only)")
            if kind == "reference" else
            "Python float64 oracle port of MANSYEnv.step/reset (no reference archive on this box), simulator only")
    return rates, {"cores": cores, "kind": kind, "wall_s": wall,
                   "sample": f"{steps[-1]} chunk-steps per interval: {cores} processes x 1 env, {what}, {seconds:.1f} s per "
                             f"interval on a 4-video x 8-user x 40-trace slice of the workload written in the reference's "
                             f"on-disk formats"}


def cpu_baseline(seconds: float, cores: int):
    """`cpu_baseline` of our line: one bounded sample of the reference arm + the simulator-only rate beside it."""
    rates, d = cpu_arm(1, seconds, cores, with_policy=True)
    sim_rates, _ = cpu_arm(1, max(2.0, seconds / 3), cores, with_policy=False)
    d.update(value=rates[0], unit=UNIT, simulator_only_value=sim_rates[0])
    return d


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    W, K = max(args.warmup, 0), max(args.steps, 1)
    # a "step" of this arm = one bounded interval of all-core stepping; the whole run is capped at ~100 s of stepping
    seconds = max(0.5, min(4.0, 100.0 / (W + K)))
    rates, d = cpu_arm(W + K, seconds, cores, with_policy=True)
    timed = rates[W:]
    value = sum(timed) / len(timed)
    n_env = args.envs_per_gpu * args.gpus
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * n_env / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 simulator scalars / f32 observations / f32 policy",
        "data": "synthetic",
        "config": {"workload": f"mansy_ppo_rollout_{args.envs_per_gpu}_envs_per_gpu", "envs": n_env,
                   "envs_per_gpu": args.envs_per_gpu,
                   "how": "CPU arm: one reference env per host core, policy forward + sample + env.step per chunk-step; "
                          "ms_per_step = time the host cores need for one lock-step of all envs at the measured rate"},
        "cpu_baseline": dict(d, value=value, unit=UNIT),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.path = os.path.join(tempfile.mkdtemp(prefix="mansy_clk_"), "clocks.csv")
        self.proc = None
        self.idx = device_index

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "20"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def _max_over_ranks(x: float, world: int) -> float:
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


RANK_TIMES = False       # --rank-times
LAST_RANK_MS = None      # per rank [ms to the end of the rollout launches, ms to the end of the exchange] of the last timed_rollout


def _all_ranks(values, world):
    """[world][len(values)] floats, every rank's `values` (torch.distributed all_gather; world 1: just this rank's)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [[float(v) for v in values]]
    t = torch.tensor(values, dtype=torch.float64, device="cuda")
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [[float(x) for x in o.cpu()] for o in out]


def make_peers(args, n_local, device):
    """The NVLink peer-memory group for the per-rollout exchange, or None (-> torch.distributed all-gather over NCCL) when
    asked for with --nccl-gather or when some rank cannot map a peer's mailbox: PeerGroup raises on EVERY rank in that
    case, so the whole job takes the same path."""
    from mansy_immersivevideostreaming_b200.rollout import PeerGroup
    if args.nccl_gather:
        return None
    try:
        return PeerGroup(n_local, device)
    except RuntimeError as exc:
        if int(os.environ.get("RANK", "0")) == 0:
            print(f"bench: peer-memory gather unavailable ({exc}); using the NCCL all-gather", file=sys.stderr)
        return None


def timed_rollout(roll, sim, peers, K, W, world):
    """W warm-up steps, then EXACTLY K rollout steps + the per-rollout gather between CUDA events on the launching
    stream, bracketed by barrier + synchronize on both sides; the device-side peer barrier right before the start
    event lines the GPUs up so host launch skew is not part of any rank's region.  Returns (ms max over ranks, stats,
    launches of our kernels inside the region)."""
    import torch
    import torch.distributed as dist
    from mansy_immersivevideostreaming_b200 import _capi
    from mansy_immersivevideostreaming_b200.rollout import gather_episode_stats
    lib = _capi.load_library()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    roll.run(W)
    gather_episode_stats(sim, peers=peers)       # a warm-up rollout includes its gather
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if peers is not None:
        peers.barrier()                          # device-side: all GPUs pass this point within microseconds
    mid = torch.cuda.Event(enable_timing=True) if RANK_TIMES else None
    launches0 = lib.mansy_kernel_launches()
    start.record()
    roll.run(K)
    if mid is not None:
        mid.record()         # --rank-times: when this rank's rollout launches were done (an event between the rollout kernel and
                             # the exchange kernel costs ~2 us and their overlap, so it is not in the default line)
    stats = gather_episode_stats(sim, peers=peers)
    stop.record()
    launches = lib.mansy_kernel_launches() - launches0
    barrier()
    global LAST_RANK_MS
    LAST_RANK_MS = _all_ranks([start.elapsed_time(mid) if mid is not None else float("nan"), start.elapsed_time(stop)], world)
    return _max_over_ranks(start.elapsed_time(stop), world), stats, int(launches)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from mansy_immersivevideostreaming_b200 import _capi
    from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE
    from mansy_immersivevideostreaming_b200.policy import PolicyNet, mansy_state_dict_shapes, seeded_state_dict
    from mansy_immersivevideostreaming_b200.rollout import PeerGroup, PolicyRollout, summarise_stats
    from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world != 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _capi.load_library()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()                           # before warm-up: the timed regions are shorter than one sampling period

    n_local = args.envs_per_gpu
    n_global = n_local * world
    tiler = ViewportTiler(device=local)
    tables = workload_tables(tiler.chunk_masks, n_global, args.diverse_qoe)   # masks come from the CUDA viewport->tiles kernel
    sim = BatchSimulator(tables, n_local, OBS_MODE_MANSY, REWARD_QOE, seed=0, worker_num=n_global,
                         env_offset=rank * n_local, device=local)
    actor_shapes, critic_shapes = mansy_state_dict_shapes()
    policy = PolicyNet(seeded_state_dict(actor_shapes, 1), seeded_state_dict(critic_shapes, 2), OBS_MODE_MANSY, device=local)
    slab_bytes = n_local * sim.obs_stride * 4
    slabs = max(4, -(-(320 << 20) // slab_bytes))                # rollout buffer > 2.5 x L2 (126 MB)
    roll = PolicyRollout(sim, policy, slabs, seed=1234)
    peers = make_peers(args, n_local, local)     # world 1: a plain pack kernel; world > 1: one NVLink push kernel

    W = max(args.warmup, 3)
    K = args.steps
    roll.reserve_timing(K)
    elapsed_ms, stats, launches = timed_rollout(roll, sim, peers, K, W, world)
    rank_ms = LAST_RANK_MS
    # per-kernel durations for the rooflines: the same K steps again with CUDA events around every launch on the
    # launching stream (events between the launches serialise them, so this pass is not the one `value` is from)
    roll.run(3, timed=True)       # first launches of the stand-alone kernels in this process (module load) stay out of the averages
    roll.run(K, timed=True)
    torch.cuda.synchronize()
    policy_sum, step_sum, timed_steps = roll.kernel_ms()
    policy_ms, step_ms = policy_sum / timed_steps, step_sum / timed_steps
    summary = summarise_stats(stats)

    # ---- e2e: the same rollout with HOST storage (tianshou's replay buffer is numpy): per step D2H actions +
    # sync, H2D actions, step, D2H next observation / reward / done / logp / value + sync -----------------------
    e2e_sim = BatchSimulator(tables, n_local, OBS_MODE_MANSY, REWARD_QOE, seed=0, worker_num=n_global,
                             env_offset=rank * n_local, device=local)
    e2e_roll = PolicyRollout(e2e_sim, policy, 4, seed=1234)
    host = e2e_roll.make_host_buffers(host_slabs=8)              # 8 x 12.9 MB pinned ring
    e2e_steps = max(20, min(K, 100))
    e2e_roll.run_host(max(5, host["obs"].shape[0]), host)        # warm-up: every slab of the pinned ring has been written once
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_roll.run_host(e2e_steps, host)
    torch.cuda.synchronize()
    e2e_s = _max_over_ranks(time.perf_counter() - t0, world)
    h2d, d2h = e2e_roll.host_bytes_per_step()
    # the ceiling of that path: plain device -> pinned-host copies of the same 12.9 MB slabs into the same ring, all ranks
    # at once (the box's host link is shared by its GPUs)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_slab = e2e_roll.buf.obs[0]
    for i in range(3):
        host["obs"][i % host["obs"].shape[0]].copy_(dev_slab, non_blocking=True)
    torch.cuda.synchronize()
    c0.record()
    for i in range(20):
        host["obs"][i % host["obs"].shape[0]].copy_(dev_slab, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    d2h_ceiling_gbs = 20 * dev_slab.numel() * 4 / (_max_over_ranks(c0.elapsed_time(c1), world) * 1e-3) / 1e9    # per rank
    e2e_sim.close()
    del e2e_roll, host

    # ---- the other BASELINE configs, each a short measurement of its own (every rank takes part) ----------------
    extra = {}
    if not args.no_extras:
        sim.close()
        del roll
        torch.cuda.empty_cache()
        extra["configs[2]"] = config3_section(args, tiler, policy, rank, world, local)
        extra["configs[3]"] = {"what": "simulator-only sweep, hashed in-kernel actions, observation materialised, one launch per "
                                       "lock-step, fresh slab per step; every rank runs its own shard, value = all ranks",
                               "entries": simulator_sweep(tables, local, world)}
        extra["configs[4]"] = config5_section(args, tables, policy, tiler, rank, world, local)
    clk = clocks.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_gbs, tflops, peak_src = measured_peaks()
    value = n_global * K / (elapsed_ms * 1e-3)
    fused = launches <= 2                        # the fused rollout kernel + the gather kernel
    # fused kernel: the step's algorithmic bytes (3 513 B) + the policy's outputs (action, logp, value,
    # logits row: 76 B); the observation rows the policy re-reads were written one phase earlier and are not counted
    fused_bytes = n_local * (BYTES_PER_STEP_MANSY + 76)
    fused_gbs = fused_bytes / (elapsed_ms / K * 1e-3) / 1e9
    step_gbs = n_local * BYTES_PER_STEP_MANSY / (step_ms * 1e-3) / 1e9
    pol_tflops = n_local * FLOP_PER_STEP_POLICY / (policy_ms * 1e-3) / 1e12
    tf32_peak = tflops / 2.0      # TF32 dense rate = 1/2 of bf16 (nominal 1.1 vs 2.25 PFLOP/s); bf16 figure measured
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 simulator scalars / f32 observations / tf32 policy (fp32 accumulate)", "data": "synthetic",
        "config": {"workload": f"mansy_ppo_rollout_{n_local}_envs_per_gpu" + ("_diverse_qoe" if args.diverse_qoe else ""),
                   "envs": n_global, "envs_per_gpu": n_local,
                   "obs_row_bytes": sim.obs_stride * 4,
                   "l2": f"observations stream into a {slabs}-slab rollout buffer of {slabs * slab_bytes >> 20} MiB "
                         "(> L2 126 MB): every step writes a slab last touched >2.5 L2-sizes ago",
                   "tables": "24 videos x 60 chunks, 1440 viewport pairs, 40 traces (SURVEY.md 8(d))",
                   "timed": ("ONE launch of the fused cluster kernel: K x (tcgen05 split-K policy forward + sample + simulator "
                             "step per 128-env tile) (mansy_rollout_policy)" if fused else
                             "K x (tcgen05 policy forward+sample launch, simulator step launch) driven from C "
                             "(mansy_rollout_policy, programmatic dependent launch)")
                            + (" + the per-rollout exchange of episode totals: one kernel storing into every peer's mailbox over "
                               "NVLink (mansy_peer_allgather_stats)" if peers is not None else
                               " + 1 NCCL all-gather of episode totals"),
                   "kernel_timing": "roofline launch durations: a second pass of the same K steps with CUDA events "
                                    "around every launch on the launching stream (serialised launches)"},
        "roofline": ({"bound": "hbm", "achieved": fused_gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": fused_gbs / hbm_gbs,
                      "traffic": FUSED_DRAM_BYTES_PER_STEP_4096 * K if n_local == 4096 else None,
                      "traffic_source": "ncu --set full of one 40-step launch, scaled by K (profiles/r02zz_fused_kernel_40steps_ncu.txt)",
                      "kernel": "policy_tc4_kernel<fused> (policy + sample + simulator step, K steps per launch)",
                      "bytes_per_launch": fused_bytes * K, "avg_launch_ms": elapsed_ms, "peak_source": peak_src,
                      "note": "4096 envs move 14.7 MB per step (2.2 us of HBM time): the step is latency-bound, see "
                              "roofline_step_kernel / configs[3] for the stand-alone kernel at HBM-filling sizes"}
                     if fused else
                     {"bound": "hbm", "achieved": step_gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": step_gbs / hbm_gbs,
                      "traffic": None, "kernel": "step_kernel<MANSY>", "bytes_per_launch": n_local * BYTES_PER_STEP_MANSY,
                      "avg_launch_ms": step_ms, "peak_source": peak_src}),
        "roofline_step_kernel": {"bound": "hbm", "achieved": step_gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": step_gbs / hbm_gbs,
                                 "kernel": "step_kernel<MANSY> (stand-alone, second pass)",
                                 "bytes_per_launch": n_local * BYTES_PER_STEP_MANSY, "avg_launch_ms": step_ms},
        "roofline_policy": {"bound": "tensor", "achieved": pol_tflops, "peak": tf32_peak, "unit": "TFLOP/s",
                            "frac": pol_tflops / tf32_peak, "kernel": "policy_tc4_kernel (tcgen05 kind::tf32, split-K 4-CTA clusters)" if 4 * ((n_local + 127) // 128) <= 3 * 148
                            else "policy_tc_kernel (tcgen05 kind::tf32)",
                            "avg_launch_ms": policy_ms, "flop_per_launch": n_local * FLOP_PER_STEP_POLICY,
                            "peak_source": "1/2 x measured bf16 cuBLAS burst (MEASURED_PEAKS.json); TF32 runs at half the bf16 rate"},
        "e2e": {"value": n_global * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps,
                "d2h_achieved_GBps_per_gpu": d2h * e2e_steps / e2e_s / 1e9,
                "d2h_ceiling_GBps_per_gpu": d2h_ceiling_gbs,
                "frac_of_d2h_ceiling": (d2h * e2e_steps / e2e_s / 1e9) / d2h_ceiling_gbs,
                "d2h_ceiling_how": "20 back-to-back cudaMemcpyAsync of one 12.9 MB observation slab into the same pinned ring, every rank at "
                                   "the same time, slowest rank",
                "path": "mansy_rollout_policy_host: per step policy launch, actions to the host (store kernel into the mapped "
                        "pinned buffer) + sync, H2D actions, step launch, reward+done+logp+value stored by one kernel into the mapped pinned buffers, D2H copy of the observation slab into the pinned host ring "
                        "on a copy stream (overlaps the next step); all copies complete inside the timed region"},
        "gpu_launches": int(launches),
        "per_rank_ms": {"rollout_launches_done": [round(r[0], 4) for r in rank_ms] if RANK_TIMES else None,
                        "exchange_done": [round(r[1], 4) for r in rank_ms],
                        "what": "device time from the start event on each rank: ms_per_step uses the maximum of the second list"},
        "clocks": clk,
        "rollout_summary": summary,
    }
    line.update(extra)
    if not args.no_extras and not args.no_mtio:
        line["expert_mpc"] = expert_section(tables, local, n_local)
        line["viewport_prediction"] = mtio_section(local, args.mtio_samples,
                                                   cpu_seconds=0.0 if (world > 1 or args.no_cpu_baseline) else 6.0)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.cpu_baseline_seconds, os.cpu_count() or 1)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def config3_section(args, tiler, policy, rank, world, local):
    """BASELINE configs[2]: 8,192 envs per GPU, one Dirichlet QoE preference vector per environment (65,536 envs at 8 GPUs)."""
    import torch
    from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE
    from mansy_immersivevideostreaming_b200.rollout import PeerGroup, PolicyRollout
    from mansy_immersivevideostreaming_b200.simulator import BatchSimulator
    n_local = 8192
    n_global = n_local * world
    tables = workload_tables(tiler.chunk_masks, n_global, diverse_qoe=True)
    sim = BatchSimulator(tables, n_local, OBS_MODE_MANSY, REWARD_QOE, seed=0, worker_num=n_global,
                         env_offset=rank * n_local, device=local)
    slabs = max(4, -(-(320 << 20) // (n_local * sim.obs_stride * 4)))
    roll = PolicyRollout(sim, policy, slabs, seed=1234)
    peers = make_peers(args, n_local, local)
    K = max(20, min(args.steps, 200))
    ms, _, launches = timed_rollout(roll, sim, peers, K, 5, world)
    out = {"workload": "mansy_ppo_rollout_8192_envs_per_gpu_diverse_qoe", "envs": n_global, "envs_per_gpu": n_local, "steps": K,
           "ms_per_step": ms / K, "value": n_global * K / (ms * 1e-3), "unit": UNIT, "gpu_launches": launches,
           "qoe_vectors": n_global}
    sim.close()
    if peers is not None:
        peers.close()
    del roll
    torch.cuda.empty_cache()
    return out


def config5_section(args, tables, policy, tiler, rank, world, local):
    """BASELINE configs[4]: MTIO viewport prediction feeding the environments, 16,384 envs over the job's GPUs.  Per
    lock-step and GPU: the predict.py mask pipeline (5 autoregressive steps -> 5 predicted points -> tile masks + IoU) for
    the 16384/N environments of the shard, then one policy + simulator step of those environments."""
    import numpy as np
    import torch
    from mansy_immersivevideostreaming_b200 import mtio as mo
    from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE
    from mansy_immersivevideostreaming_b200.mtio import ViewportTransformerMTIO
    from mansy_immersivevideostreaming_b200.rollout import PolicyRollout
    from mansy_immersivevideostreaming_b200.simulator import BatchSimulator
    from mansy_immersivevideostreaming_b200 import synth
    n_global = args.mtio_samples
    n = max(128, n_global // world)
    net = ViewportTransformerMTIO(device=f"cuda:{local}", max_batch=n).load_state_dict(mo.seeded_mtio_state_dict(3, bias=True))
    hist, cur = mo.synthetic_history(n, 4 + rank)
    h, c = torch.from_numpy(hist).cuda(local), torch.from_numpy(cur).cuda(local)
    gt = torch.from_numpy(np.mod(cur + np.cumsum(np.random.default_rng(1 + rank).normal(0, 0.03, size=(n, 15, 2)), axis=1), 1.0)
                          .astype(np.float32)).cuda(local)
    t = tables.with_samples(synth.per_env_samples(tables, n * world))
    sim = BatchSimulator(t, n, OBS_MODE_MANSY, REWARD_QOE, seed=0, worker_num=n * world, env_offset=rank * n, device=local)
    roll = PolicyRollout(sim, policy, 4, seed=1234)
    for _ in range(3):
        net.predict_chunk_masks(h, c, gt, tiler)
        roll.run(1)
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
    reps = 10
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(reps):
        net.predict_chunk_masks(h, c, gt, tiler)
    e1.record()
    for _ in range(reps):
        net.predict_chunk_masks(h, c, gt, tiler)
        roll.run(1)
    e2.record()
    torch.cuda.synchronize()
    ms_masks = _max_over_ranks(e0.elapsed_time(e1) / reps, world)
    ms_chain = _max_over_ranks(e1.elapsed_time(e2) / reps, world)
    out = {"workload": f"mtio_masks_then_rollout_step_{n}_envs_per_gpu", "envs": n * world, "envs_per_gpu": n,
           "mask_pipeline_ms": ms_masks, "mask_pipeline_samples_per_s": n * world / (ms_masks * 1e-3),
           "chain_ms_per_step": ms_chain, "value": n * world / (ms_chain * 1e-3), "unit": "chunk-steps/s with an on-line viewport prediction per step"}
    sim.close(); net.close()
    del roll
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------
# MTIO viewport prediction (BASELINE configs[4])
# ---------------------------------------------------------------------------------------------
def mtio_flop_per_sample(T=5, Tm=3, F=15, n_enc=2, n_dec=2, d=512):
    """Multiply-adds x 2 of the key/value-cached formulation; attention itself is < 1 %."""
    mac = n_enc * T * 6 * d * d + T * 3 * d * d + n_dec * Tm * 2 * d * d + F * n_dec * 8 * d * d
    return 2 * mac


def mtio_cpu_baseline(seconds: float = 10.0):
    """The numpy oracle port of ViewportTransformerMTIO.sample (re-decodes the prefix every step like the reference,
    BLAS threads as numpy sees fit) on a bounded sample."""
    from oracle import mtio_oracle as mo
    sd = mo.seeded_mtio_state_dict(3, bias=True)
    n, done = 64, 0
    hist, cur = mo.synthetic_history(n, 4)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        mo.sample(sd, hist, cur, 15)
        done += n
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "viewport samples/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{done} samples in batches of {n} through the numpy restatement of model.sample (15 autoregressive "
                      f"steps, whole-prefix decoding as mtio.py:120-123; numpy/BLAS threading)", "wall_s": dt}


def mtio_section(device_index: int, n: int, reps: int = 5, cpu_seconds: float = 0.0):
    """16,384 viewport histories -> 15 predicted points each (model.sample), and the predict.py mask pipeline."""
    import numpy as np
    import torch
    from mansy_immersivevideostreaming_b200.mtio import ViewportTransformerMTIO
    from mansy_immersivevideostreaming_b200.simulator import ViewportTiler
    from mansy_immersivevideostreaming_b200 import mtio as mo      # seeded weights / synthetic walks (numpy)
    _, tflops, _ = measured_peaks()
    sd = mo.seeded_mtio_state_dict(3, bias=True)
    net = ViewportTransformerMTIO(device=f"cuda:{device_index}", max_batch=n).load_state_dict(sd)
    hist, cur = mo.synthetic_history(n, 4)
    h, c = torch.from_numpy(hist).cuda(device_index), torch.from_numpy(cur).cuda(device_index)
    for _ in range(3):
        net.sample(h, c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        net.sample(h, c)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    net.sample(h, c, timed=True)
    torch.cuda.synchronize()
    kms, cnt = net.kernel_ms()
    flop = n * mtio_flop_per_sample()
    gemm_tflops = flop / (kms[0] * 1e-3) / 1e12
    # predict.py pipeline: 5 steps, OR of the tile masks of 5 points, IoU against the ground truth (a13-a16)
    tiler = ViewportTiler(device=device_index)
    gt = torch.from_numpy(np.mod(cur + np.cumsum(np.random.default_rng(1).normal(0, 0.03, size=(n, 15, 2)), axis=1), 1.0)
                          .astype(np.float32)).cuda(device_index)
    net.predict_chunk_masks(h, c, gt, tiler)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        net.predict_chunk_masks(h, c, gt, tiler)
    e1.record()
    torch.cuda.synchronize()
    ms_masks = e0.elapsed_time(e1) / reps
    # end to end: pinned host histories in, predictions out
    hp, cp = torch.from_numpy(hist).pin_memory(), torch.from_numpy(cur).pin_memory()
    out = torch.empty((n, 15, 2), dtype=torch.float32).pin_memory()
    net.sample_host(hp, cp, out)
    t0 = time.perf_counter()
    for _ in range(reps):
        net.sample_host(hp, cp, out)
    e2e_s = (time.perf_counter() - t0) / reps
    sec = {
        "workload": f"mtio_sample_{n} (his 5, fut 15, d_model 512, 2+2 layers: predict.py defaults)",
        "metric": "viewport samples/s", "value": n / (ms * 1e-3), "ms_per_batch": ms, "dtype": "tf32 (fp32 accumulate)",
        "flop_per_sample": mtio_flop_per_sample(),
        "kernel_ms": {"gemm": kms[0], "attention": kms[1], "other": kms[2]},
        "kernel_launches": {"gemm": int(cnt[0]), "attention": int(cnt[1]), "other": int(cnt[2])},
        "roofline": {"bound": "tensor", "achieved": gemm_tflops, "peak": tflops / 2.0, "unit": "TFLOP/s",
                     "frac": gemm_tflops / (tflops / 2.0), "kernel": "mtio_gemm_kernel (tcgen05 kind::tf32; all 191 launches of a batch)",
                     "avg_launch_ms": kms[0] / max(int(cnt[0]), 1),
                     "peak_source": "1/2 x measured bf16 cuBLAS burst (MEASURED_PEAKS.json); TF32 runs at half the bf16 rate",
                     "traffic": None},
        "mask_pipeline": {"value": n / (ms_masks * 1e-3), "unit": "viewport samples/s", "ms_per_batch": ms_masks,
                          "what": "5 of 15 autoregressive steps + tile masks + IoU (predict.py:33-48 uses the first 5 points)"},
        "e2e": {"value": n / e2e_s, "unit": "viewport samples/s", "h2d_bytes_per_step": n * 48, "d2h_bytes_per_step": n * 120},
    }
    if cpu_seconds > 0:
        sec["cpu_baseline"] = mtio_cpu_baseline(cpu_seconds)
    net.close()
    return sec


def expert_section(tables, device_index: int, n: int):
    """MPC expert: decisions/s of ExpertEnv.choose_action for n environments, horizons 2 and 4."""
    import torch
    from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE
    from mansy_immersivevideostreaming_b200.simulator import BatchSimulator
    sim = BatchSimulator(tables, n, OBS_MODE_MANSY, REWARD_QOE, seed=0, device=device_index)
    sim.reset()
    sim.rollout_random(10, seed=3)            # desynchronise trace positions / buffers
    out = {"envs": n, "what": "ExpertEnv.choose_action (expert_env.py:358-422): exhaustive search over 15^horizon action sequences"}
    for h in (2, 4):
        sim.expert_actions(h)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            sim.expert_actions(h)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out[f"horizon_{h}"] = {"ms": ms, "decisions_per_s": n / (ms * 1e-3), "sequences_per_s": n * 15 ** h / (ms * 1e-3)}
    sim.close()
    return out


def run_mtio(args):
    """Stand-alone line for the MTIO workload (one process per GPU, samples sharded, no collective)."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    sec = mtio_section(local, args.mtio_samples, reps=max(args.steps if args.steps < 100 else 10, 3),
                       cpu_seconds=0.0 if (args.no_cpu_baseline or world > 1 or rank != 0) else args.cpu_baseline_seconds)
    clk = clocks.stop() if rank == 0 else None
    ms = sec["ms_per_batch"]
    if world > 1:
        tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    if rank == 0:
        line = {"metric": "viewport samples/s", "value": world * args.mtio_samples / (ms * 1e-3), "unit": "viewport samples/s",
                "n_gpus": world, "steps": max(args.steps if args.steps < 100 else 10, 3), "warmup": 3, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": sec["dtype"], "data": "synthetic",
                "config": {"workload": sec["workload"], "samples_per_gpu": args.mtio_samples,
                           "l2": "activations + key/value cache of a batch: 3.3 GB (>> L2)"},
                "roofline": sec["roofline"], "e2e": sec["e2e"], "mask_pipeline": sec["mask_pipeline"],
                "kernel_ms": sec["kernel_ms"], "gpu_launches": sum(sec["kernel_launches"].values()), "clocks": clk}
        if "cpu_baseline" in sec:
            line["cpu_baseline"] = sec["cpu_baseline"]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def simulator_sweep(tables, device_index, world=1, sizes=(65536, 1048576)):
    """Simulator-only kernel at HBM-filling env counts (hashed in-kernel actions, observation
    materialised, one launch per step, fresh slab per step), both observation layouts."""
    import torch
    from mansy_immersivevideostreaming_b200 import synth
    from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, OBS_MODE_SIMPLE, REWARD_QOE
    from mansy_immersivevideostreaming_b200.simulator import BatchSimulator
    hbm_gbs, _, _ = measured_peaks()
    out = []
    for mode, name, bytes_per_step in ((OBS_MODE_MANSY, "mansy", 3513), (OBS_MODE_SIMPLE, "simple_rl", 1805)):
        for n in sizes:
            t = tables.with_samples(synth.per_env_samples(tables, n))
            sim = BatchSimulator(t, n, mode, REWARD_QOE, seed=0, device=device_index)
            slab = n * sim.obs_stride * 4
            slabs = max(2, min(8, -(-(320 << 20) // slab)))
            obs = torch.empty((slabs, n, sim.obs_stride), dtype=torch.float32, device=sim.device)
            rew = torch.empty(n, dtype=torch.float32, device=sim.device)
            done = torch.empty(n, dtype=torch.uint8, device=sim.device)
            sim.reset(None, obs[0])
            for k in range(5):
                sim.rollout_random(1, seed=5, step0=k, obs=obs[k % slabs], reward=rew, done=done)
            torch.cuda.synchronize()
            reps = 30
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(reps):
                sim.rollout_random(1, seed=5, step0=5 + k, obs=obs[k % slabs], reward=rew, done=done)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gbs = n * bytes_per_step / (ms * 1e-3) / 1e9              # this GPU's kernel against this GPU's HBM peak
            ms_all = _max_over_ranks(ms, world)
            out.append({"env": name, "envs_per_gpu": n, "ms_per_step": ms_all, "chunk_steps_per_s": world * n / (ms_all * 1e-3),
                        "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / hbm_gbs, "bytes_per_chunk_step": bytes_per_step,
                        "kernel": "step_kernel<%s>" % ("MANSY" if mode == OBS_MODE_MANSY else "SIMPLE"),
                        "traffic": STEP_DRAM_BYTES_MANSY_1M if (mode == OBS_MODE_MANSY and n == 1048576) else None})
            sim.close()
            del obs
            torch.cuda.empty_cache()
    return out


def profile_sim(n):
    """A handful of simulator-only launches at n envs (MANSY obs), for `ncu -k regex:step_kernel`."""
    import torch
    from mansy_immersivevideostreaming_b200 import synth
    from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE
    from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler
    tables = workload_tables(ViewportTiler().chunk_masks, n)
    sim = BatchSimulator(tables, n, OBS_MODE_MANSY, REWARD_QOE, seed=0)
    slabs = max(2, min(8, -(-(320 << 20) // (n * sim.obs_stride * 4))))
    obs = torch.empty((slabs, n, sim.obs_stride), dtype=torch.float32, device=sim.device)
    rew = torch.empty(n, dtype=torch.float32, device=sim.device)
    done = torch.empty(n, dtype=torch.uint8, device=sim.device)
    sim.reset(None, obs[0])
    for k in range(12):
        sim.rollout_random(1, seed=5, step0=k, obs=obs[k % slabs], reward=rew, done=done)
    torch.cuda.synchronize()


def main():
    args = parse_args()
    global RANK_TIMES
    RANK_TIMES = bool(args.rank_times)
    if args.profile_sim:
        profile_sim(args.profile_sim)
        return
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "mtio":
        run_mtio(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
