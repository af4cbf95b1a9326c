"""Quick device timing of the MTIO inference path: samples/s, per-kernel-class breakdown, error vs the oracle.

    python tools/mtio_bench.py [batch] [reps]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mansy_immersivevideostreaming_b200.mtio import ViewportTransformerMTIO  # noqa: E402
from oracle import mtio_oracle as mo  # noqa: E402  (checker only)

FLOP_PER_SAMPLE = None


def flops_per_sample(T=5, Tm=3, F=15, n_enc=2, n_dec=2, d=512):
    mac = n_enc * T * (3 * d * d + d * d + 2 * d * d)          # qkv, out, ffn
    mac += T * 3 * d * d                                        # distill conv
    mac += n_dec * Tm * 2 * d * d                               # memory k / v
    mac += F * n_dec * (3 * d * d + d * d + d * d + d * d + 2 * d * d)
    return 2 * mac


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    sd = mo.seeded_mtio_state_dict(3, bias=True)
    net = ViewportTransformerMTIO(device="cuda:0", max_batch=n).load_state_dict(sd)
    hist, cur = mo.synthetic_history(n, 4)
    h, c = torch.from_numpy(hist).cuda(), torch.from_numpy(cur).cuda()
    if os.environ.get("MTIO_BENCH_PROFILE"):      # lean run for ncu: one warm-up pass, one profiled pass
        net.sample(h, c)
        torch.cuda.synchronize()
        net.sample(h, c)
        torch.cuda.synchronize()
        return
    k = min(n, 64)
    want = mo.sample(sd, hist[:k], cur[:k], 15)
    for fp32 in (True, False):
        net.fp32 = fp32
        got = net.sample(h[:k], c[:k]).cpu().numpy()
        print(f"{'fp32' if fp32 else 'tf32'} max abs err vs oracle ({k} samples): {np.abs(got - want).max():.3e}", flush=True)
    fl = flops_per_sample()
    for fp32 in (False, True):
        net.fp32 = fp32
        for _ in range(2):
            net.sample(h, c)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps if not fp32 else 1):
            net.sample(h, c)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps if not fp32 else 1)
        print(f"{'fp32' if fp32 else 'tf32'} batch {n}: {ms:.3f} ms  {n / ms * 1e3:.3e} samples/s  {n * fl / ms / 1e9:.1f} TFLOP/s "
              f"({fl / 1e6:.1f} MFLOP/sample)", flush=True)
        net.sample(h, c, timed=True)
        torch.cuda.synchronize()
        kms, cnt = net.kernel_ms()
        print(f"   serialised: gemm {kms[0]:.3f} ms / {cnt[0]} launches, attention {kms[1]:.3f} ms / {cnt[1]}, other {kms[2]:.3f} ms / {cnt[2]}",
              flush=True)
    net.fp32 = False
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    net.sample(h, c, steps=5)
    e0.record()
    for _ in range(reps):
        net.sample(h, c, steps=5)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"tf32 batch {n}, 5 of 15 steps (what predict.py's masks use): {ms:.3f} ms  {n / ms * 1e3:.3e} samples/s", flush=True)
    # host round trip
    net.fp32 = False
    hp, cp = torch.from_numpy(hist).pin_memory(), torch.from_numpy(cur).pin_memory()
    out = torch.empty((n, 15, 2), dtype=torch.float32).pin_memory()
    net.sample_host(hp, cp, out)
    t0 = time.perf_counter()
    for _ in range(reps):
        net.sample_host(hp, cp, out)
    dt = (time.perf_counter() - t0) / reps
    print(f"host buffers: {dt * 1e3:.3f} ms  {n / dt:.3e} samples/s", flush=True)


if __name__ == "__main__":
    main()
