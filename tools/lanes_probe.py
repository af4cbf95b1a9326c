"""Probe: the rollout of N environments as G independent lanes (shards of the env range, each with its own simulator handle,
policy handle and CUDA stream) on ONE GPU -- policy of one lane overlapping the simulator step of another.
    python tools/lanes_probe.py 8192 2 4      # envs, lanes, policy split (0 auto / 1 / 4)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import workload_tables
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE
from mansy_immersivevideostreaming_b200.policy import PolicyNet, mansy_state_dict_shapes, seeded_state_dict
from mansy_immersivevideostreaming_b200.rollout import PolicyRollout
from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
shapes = mansy_state_dict_shapes()
tiler = ViewportTiler(device=0)
tables = workload_tables(tiler.chunk_masks, n)
K = 300
for G, split, fused in [(1, 0, True), (2, 1, False), (2, 4, False), (2, 0, True), (3, 1, False), (3, 4, False)]:
    per = n // G // 128 * 128
    if per * G != n:
        continue
    sims, rolls, pols, streams = [], [], [], []
    for g in range(G):
        sim = BatchSimulator(tables, per, OBS_MODE_MANSY, REWARD_QOE, seed=0, worker_num=n, env_offset=g * per)
        pol = PolicyNet(seeded_state_dict(shapes[0], 1), seeded_state_dict(shapes[1], 2), OBS_MODE_MANSY)
        pol.set_tc_split(split)
        slabs = max(4, -(-(320 << 20) // (n * sim.obs_stride * 4)))
        rolls.append(PolicyRollout(sim, pol, slabs, seed=1234))
        sims.append(sim); pols.append(pol); streams.append(torch.cuda.Stream())
    torch.cuda.synchronize()

    def go(k):
        # round-robin in short bursts so that the host feeds every lane's stream from the start
        left = k
        while left > 0:
            b = min(left, 10)
            for g in range(G):
                with torch.cuda.stream(streams[g]):
                    rolls[g].run(b, fused=fused)
            left -= b
    go(30)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    go(K)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / K * 1e3
    import time
    torch.cuda.synchronize()
    h0 = time.perf_counter(); go(K); h1 = time.perf_counter()
    torch.cuda.synchronize()
    host_us = (h1 - h0) / K * 1e6
    # the same loop as a CUDA graph (no host launch cost): capture 20 steps on the lanes' streams, replay 15 times
    cap = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    gus = float("nan")
    try:
        with torch.cuda.stream(cap):
            graph.capture_begin(capture_error_mode="relaxed")
            ev = torch.cuda.Event(); ev.record(cap)
            for sg in streams:
                sg.wait_event(ev)
            go(20)
            for sg in streams:
                cap.wait_stream(sg)
            graph.capture_end()
        torch.cuda.synchronize()
        graph.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(15):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        gus = e0.elapsed_time(e1) / 300 * 1e3
    except Exception as exc:      # noqa: BLE001
        print("graph capture failed:", str(exc)[:200])
    print(f"{n} envs, {G} lanes x {per}, split {split}, {'fused' if fused else 'PDL two-kernel'}: {us:6.1f} us per step = {n / us:6.1f} M chunk-steps/s; host enqueue {host_us:5.1f} us per step; as a graph {gus:6.1f} us = {n / gus:6.1f} M/s", flush=True)
    for sim, pol in zip(sims, pols):
        sim.close(); pol.close()
    del rolls, sims, pols
    torch.cuda.empty_cache()
