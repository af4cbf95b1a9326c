"""Print the in-kernel timeline of policy_tc_kernel (CTA 0) -- a profiling aid, run on a GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY
from mansy_immersivevideostreaming_b200.policy import PolicyNet, mansy_state_dict_shapes, seeded_state_dict
from mansy_immersivevideostreaming_b200._capi import check

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
split = int(sys.argv[2]) if len(sys.argv) > 2 else 0
shapes = mansy_state_dict_shapes()
net = PolicyNet(seeded_state_dict(shapes[0], 1), seeded_state_dict(shapes[1], 2), OBS_MODE_MANSY)
net.set_tc_split(split)
obs = torch.rand((n, 784), device="cuda")
logits = torch.empty((n, 16), device="cuda"); value = torch.empty(n, device="cuda")
act = torch.empty(n, dtype=torch.int32, device="cuda"); logp = torch.empty(n, device="cuda")
tl = torch.zeros(512, dtype=torch.int64, device="cuda")
for rep in range(3):
    tl.zero_()
    check(net.lib.mansy_policy_forward_tc_timeline(net._h, obs.data_ptr(), 784, n, logits.data_ptr(), value.data_ptr(), act.data_ptr(),
                                                   logp.data_ptr(), 1, rep, 0, None, None, tl.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
t = tl.cpu().numpy()
t0 = t[511]
print("job  issue  arrive  mma_issued   (cycles since kernel start)")
for j in range(128):
    if t[j] == 0: break
    print(f"{j:3d} {t[j]-t0:7d} {t[128+j]-t0:7d} {t[256+j]-t0:7d}   lat={t[128+j]-t[j]}")
print("epilogue begin/end per branch:")
for i in range(11):
    print(i, t[384+2*i]-t0, t[385+2*i]-t0)
names = {480: "partial D2 done", 482: "partials stored (L2)", 483: "cluster sync 1", 484: "hidden + heads, D3 pushed",
         485: "cluster sync 2", 486: "own 32 rows written"}
if t[480]:
    print("split-K phases (CTA MANSY_TC_TIMELINE_CTA, cycles since kernel start):")
    for k, nm in names.items():
        print(f"  {nm:28s} {t[k]-t0:7d}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for rep in range(50):
    net.forward_tc(obs, logits, value, act, logp, seed=1, step=rep)
e1.record(); torch.cuda.synchronize()
print("avg launch us (back-to-back, L2-warm):", e0.elapsed_time(e1) / 50 * 1e3)
