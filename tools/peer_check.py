"""2+ rank check of the NVLink peer-memory exchange (csrc/mansy_peer.cu); run under torch.distributed.run on a multi-GPU
box:  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mansy_immersivevideostreaming_b200 import synth
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, REWARD_QOE, SimConfig
from mansy_immersivevideostreaming_b200.rollout import PeerGroup, gather_episode_stats
from mansy_immersivevideostreaming_b200.simulator import BatchSimulator, ViewportTiler

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1024
t = synth.make_synthetic_tables(ViewportTiler(SimConfig(), device=local).chunk_masks, n_videos=3, n_users=4, n_traces=5, seed=21,
                                trace_len_range=(40, 90))
t = t.with_samples(synth.per_env_samples(t, n * world))
sim = BatchSimulator(t, n, OBS_MODE_MANSY, REWARD_QOE, seed=0, worker_num=n * world, env_offset=rank * n, device=local)
sim.reset()
peers = PeerGroup(n, local)
for it in range(6):
    sim.rollout_random(20 + it, seed=5 + it, step0=100 * it)
    got = peers.gather_episode_stats(sim).clone()
    want = gather_episode_stats(sim)                       # torch.distributed / NCCL path on the same totals
    torch.cuda.synchronize()
    assert got.shape == (n * world, 6) and torch.equal(got, want), (rank, it)
    assert torch.equal(got[rank * n:(rank + 1) * n], sim.episode_totals())
    peers.barrier()
# shard invariance of the gathered array: every rank holds the same bytes
ref = got.clone()
dist.broadcast(ref, src=0)
assert torch.equal(ref, got)
# latency of the exchange, device-timed
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
peers.barrier(); torch.cuda.synchronize()
e0.record()
for _ in range(50):
    peers.gather_episode_stats(sim)
e1.record()
torch.cuda.synchronize()
assert not peers.timed_out()
if rank == 0:
    print(f"peer_check ok: world {world}, {n} envs per rank, {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per gather")
peers.close()
dist.destroy_process_group()
