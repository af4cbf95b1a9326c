"""Concurrent device -> pinned-host bandwidth of all ranks with the pinned ring allocated (a) wherever the process happens to
run, (b) on the NUMA node of the rank's GPU (set_mempolicy before the allocation).  Run under torch.distributed.run."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mansy_immersivevideostreaming_b200 import numa

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
node = numa.gpu_numa_node(local)
info = {"rank": rank, "gpu_node": node, "nodes": numa.online_nodes(), "allowed_cpus": len(os.sched_getaffinity(0)),
        "cpus_of_node": {n: len(numa.node_cpus(n) & os.sched_getaffinity(0)) for n in numa.online_nodes()},
        "mems_allowed": numa.mems_allowed()}
if rank == 0:
    os.system("nvidia-smi topo -m 2>&1 | head -14")
print(info, flush=True)


def measure(tag, bind):
    slab = torch.empty((4096, 784), dtype=torch.float32, device="cuda").normal_()
    with numa.memory_on_node(node if bind else None) as ok:
        ring = torch.empty((8, 4096, 784), dtype=torch.float32).pin_memory()
    where = numa.pages_node(ring.data_ptr(), ring.numel() * 4)
    for i in range(8):
        ring[i].copy_(slab, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40):
        ring[i % 8].copy_(slab, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = 40 * slab.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    print(f"rank {rank} {tag}: {gbs:6.1f} GB/s  (policy applied: {ok}, pages on nodes {where})", flush=True)
    if world > 1:
        dist.barrier()
    del ring


measure("default", False)
measure("bound to the GPU's node", True)
measure("default again", False)
if world > 1:
    dist.destroy_process_group()
