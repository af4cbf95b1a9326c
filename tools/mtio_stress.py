"""Run-to-run determinism stress of the MTIO path: N passes over 16,384 samples must equal the first bit for bit."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mansy_immersivevideostreaming_b200.mtio import ViewportTransformerMTIO, seeded_mtio_state_dict, synthetic_history  # noqa: E402

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 40
sd = seeded_mtio_state_dict(29, bias=True)
n = 16384
hist, cur = synthetic_history(n, 40)
h, c = torch.from_numpy(hist).cuda(), torch.from_numpy(cur).cuda()
whole = ViewportTransformerMTIO(device="cuda:0", max_batch=n).load_state_dict(sd)
parts = ViewportTransformerMTIO(device="cuda:0", max_batch=4096).load_state_dict(sd)
ref = whole.sample(h, c)
bad = 0
for i in range(passes):
    got = (whole if i % 2 == 0 else parts).sample(h, c)
    d = (got != ref).any(dim=2).any(dim=1).sum().item()
    bad += d
    if d:
        print("pass", i, "rows that differ:", d, flush=True)
print(f"{passes} passes, rows that differ in total: {bad}")
