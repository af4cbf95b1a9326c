"""Rollout driver: policy forward -> Categorical sample -> lock-step simulator step, all on the device.

Replaces the host loop of tianshou's ``Collector.collect`` around ``policy(batch)`` /
``env.step(act)`` (bitrate_selection/run_mansy.py:161-175 for the test loop):
observations are written by the step kernel straight into the rollout buffer the learner reads
(no per-step host<->device copies, no ``buffer.add`` copy), actions are sampled on the device, and
finished episodes are reset inside the step kernel.

Multi-GPU: environments are sharded by contiguous index ranges (no collective on the step path);
``gather_episode_stats`` is the single all-gather per rollout of fixed-size per-env statistics.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from ._capi import (ROLLOUT_FP32_POLICY, ROLLOUT_NO_PDL, ROLLOUT_NO_ZERO_COPY, ROLLOUT_TIME_KERNELS, ROLLOUT_TWO_KERNELS, Rollout, RolloutHost,
                    check)
from .config import OBS_MODE_NONE
from .policy import PolicyNet
from .simulator import BatchSimulator

# columns of BatchSimulator.episode_stats() exchanged per rollout (include/mansy_b200.h MANSY_STAT_TOT_*)
STAT_COLUMNS = (6, 7, 8, 9, 10, 11)     # sum qoe, qoe1, qoe2, qoe3, steps, episodes


class RolloutBuffers:
    """Ring of ``slabs`` observation slabs ``[slabs, N, stride]`` plus per-step scalars."""

    def __init__(self, sim: BatchSimulator, slabs: int):
        dev, n = sim.device, sim.n_envs
        self.slabs = int(slabs)
        self.obs = torch.empty((self.slabs, n, sim.obs_stride), dtype=torch.float32, device=dev)
        self.actions = torch.empty((self.slabs, n), dtype=torch.int32, device=dev)
        self.logp = torch.empty((self.slabs, n), dtype=torch.float32, device=dev)
        self.value = torch.empty((self.slabs, n), dtype=torch.float32, device=dev)
        self.reward = torch.empty((self.slabs, n), dtype=torch.float32, device=dev)
        self.done = torch.empty((self.slabs, n), dtype=torch.uint8, device=dev)
        self.logits = torch.empty((n, 16), dtype=torch.float32, device=dev)

    def bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.obs, self.actions, self.logp, self.value, self.reward, self.done))


class PolicyRollout:
    """obs[t] -> logits/value/action (one tcgen05 launch) -> simulator step -> obs[t+1] (one launch).

    ``run`` drives the loop from C (``mansy_rollout_policy``: no Python between launches); ``step`` is the
    same step from Python; ``run_host`` keeps the rollout storage in pinned host memory
    (``mansy_rollout_policy_host``), the data flow of tianshou's Collector with numpy buffers."""

    def __init__(self, sim: BatchSimulator, policy: PolicyNet, slabs: int, seed: int = 0, tensor_cores: bool = True):
        if sim.obs_mode == OBS_MODE_NONE:
            raise ValueError("the policy consumes observation rows")
        if slabs < 2:
            raise ValueError("a rollout needs at least 2 observation slabs")
        self.sim, self.policy, self.seed = sim, policy, int(seed)
        self.tensor_cores = bool(tensor_cores)
        self.buf = RolloutBuffers(sim, slabs)
        self.t = 0                                   # global step counter (keys the action sampler)
        sim.reset(None, self.buf.obs[0])
        b = self.buf
        self._c = Rollout(b.obs.data_ptr(), b.obs.stride(1), b.slabs, b.actions.data_ptr(), b.logp.data_ptr(),
                          b.value.data_ptr(), b.reward.data_ptr(), b.done.data_ptr(), b.logits.data_ptr())

    def _flags(self, timed: bool = False, fused: bool = True, pdl: bool = True) -> int:
        return ((0 if self.tensor_cores else ROLLOUT_FP32_POLICY) | (ROLLOUT_TIME_KERNELS if timed else 0) |
                (0 if fused else ROLLOUT_TWO_KERNELS) | (0 if pdl else ROLLOUT_NO_PDL))

    def step(self) -> None:
        b, s = self.buf, self.t % self.buf.slabs
        nxt = (self.t + 1) % b.slabs
        if self.tensor_cores:
            self.policy.forward_tc_sim(self.sim, b.obs[s], b.logits, b.value[s], b.actions[s], b.logp[s], seed=self.seed, step=self.t)
        else:
            self.policy.forward(b.obs[s], b.logits, b.value[s])
            self.policy.sample(b.logits, self.seed, self.t, self.sim.env_offset, b.actions[s], b.logp[s])
        self.sim.step(b.actions[s], auto_reset=True, obs=b.obs[nxt], reward=b.reward[s], done=b.done[s])
        self.t += 1

    def run(self, n_steps: int, timed: bool = False, fused: bool = True, pdl: bool = True) -> None:
        """``n_steps`` rollout steps launched from C on the current stream (asynchronous).  Batches of up to two 128-env tiles per resident 4-CTA cluster (8 448 envs on B200) run as ONE
        launch of the fused policy+step cluster kernel unless ``fused=False`` / ``timed=True`` (two launches per
        step, programmatic dependent launch unless ``pdl=False``)."""
        check(self.sim.lib.mansy_rollout_policy(self.sim._h, self.policy._h, C.byref(self._c), int(n_steps), self.t,
                                                self.seed, self._flags(timed, fused, pdl), self.sim._stream()))
        self.t += int(n_steps)

    def reserve_timing(self, n_steps: int) -> None:
        """Create the events of a later ``run(n_steps, timed=True)`` now (outside any timed region)."""
        check(self.sim.lib.mansy_rollout_reserve_timing(self.sim._h, int(n_steps)))

    def kernel_ms(self) -> Tuple[float, float, int]:
        """(sum policy ms, sum step ms, steps) of the last ``run(..., timed=True)``; call after a synchronize."""
        pm, sm, n = C.c_double(0), C.c_double(0), C.c_int32(0)
        check(self.sim.lib.mansy_rollout_kernel_ms(self.sim._h, C.byref(pm), C.byref(sm), C.byref(n)))
        return pm.value, sm.value, n.value

    def make_host_buffers(self, host_slabs: int = 2) -> Dict[str, torch.Tensor]:
        """Pinned host ring for ``run_host``, allocated on the memory node of this simulator's GPU when the machine has
        more than one and says which (``numa.memory_on_node``; a no-op otherwise)."""
        from . import numa
        n, st = self.sim.n_envs, self.sim.obs_stride
        pin = lambda *shape, dtype: torch.empty(shape, dtype=dtype).pin_memory()   # noqa: E731
        with numa.memory_on_node(numa.gpu_numa_node(self.sim.device_index)):
            return {"obs": pin(host_slabs, n, st, dtype=torch.float32), "actions": pin(host_slabs, n, dtype=torch.int32),
                    "logp": pin(host_slabs, n, dtype=torch.float32), "value": pin(host_slabs, n, dtype=torch.float32),
                    "reward": pin(host_slabs, n, dtype=torch.float32), "done": pin(host_slabs, n, dtype=torch.uint8)}

    def run_host(self, n_steps: int, host: Dict[str, torch.Tensor], zero_copy: bool = True) -> None:
        """Rollout with host storage: per step actions to the host + sync, H2D actions, step, D2H results + sync.
        ``zero_copy``: the actions reach the (device-mapped) pinned host buffer through a store kernel instead of a
        cudaMemcpyAsync that would queue behind the previous step's observation copy."""
        hc = RolloutHost(host["obs"].shape[0], host["obs"].data_ptr(), host["actions"].data_ptr(), host["logp"].data_ptr(),
                         host["value"].data_ptr(), host["reward"].data_ptr(), host["done"].data_ptr())
        check(self.sim.lib.mansy_rollout_policy_host(self.sim._h, self.policy._h, C.byref(self._c), C.byref(hc), int(n_steps),
                                                     self.t, self.seed, self._flags() | (0 if zero_copy else ROLLOUT_NO_ZERO_COPY),
                                                     self.sim._stream()))
        self.t += int(n_steps)

    def host_bytes_per_step(self) -> Tuple[int, int]:
        """(H2D, D2H) bytes one ``run_host`` step moves."""
        n = self.sim.n_envs
        return 4 * n, n * (self.sim.obs_stride * 4 + 4 + 4 + 4 + 4 + 1)


def all_gather_stats(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather equally sized per-rank ``[N_local, C]`` blocks into ``[N_global, C]`` in rank order
    (rank r owns global envs ``[r*N_local, (r+1)*N_local)``).  NCCL on GPUs, gloo in the CPU tests."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    local = local.contiguous()
    out = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out


def broadcast_state_dict(sd, src: int = 0, device=None, group=None):
    """The optional weight broadcast of SURVEY 8(e): every rank gets rank ``src``'s policy (or MTIO) state dict before a
    rollout, as ONE flat float32 buffer (1.31 M parameters = 5.2 MB for the MANSY policy) through
    ``torch.distributed.broadcast`` (NCCL on GPUs, gloo in the CPU tests).  Keys, shapes and order must agree on all
    ranks (they come from the same architecture); returns ``{name: float32 numpy array}`` for ``PolicyNet`` / ``load_state_dict``."""
    import numpy as np
    import torch.distributed as dist
    names = list(sd.keys())
    arrs = [np.ascontiguousarray(np.asarray(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v, dtype=np.float32))
            for v in sd.values()]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(zip(names, arrs))
    flat = torch.from_numpy(np.concatenate([a.reshape(-1) for a in arrs]))
    if device is not None:
        flat = flat.to(device)
    dist.broadcast(flat, src=src, group=group)
    flat = flat.cpu().numpy()
    out, off = {}, 0
    for name, a in zip(names, arrs):
        out[name] = flat[off:off + a.size].reshape(a.shape).copy()
        off += a.size
    return out


class _DevArray:
    """``__cuda_array_interface__`` over library-owned device memory (zero-copy ``torch.as_tensor``)."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
        self._owner = owner           # keeps the allocation alive


class PeerGroup:
    """The job's GPUs as a peer group: the per-rollout all-gather of episode totals as ONE kernel that stores into
    every peer's mailbox over NVLink (CUDA IPC mapped memory) and a device-side barrier -- ``csrc/mansy_peer.cu``.
    The 64-byte IPC handles are exchanged once through ``torch.distributed`` (any backend); afterwards nothing on the
    gather path goes through a collective library.  ``world == 1`` needs no process group."""

    def __init__(self, n_local: int, device: int, group=None):
        import torch.distributed as dist
        from . import _capi
        self.lib = _capi.load_library()
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0
        self.n_local, self.device = int(n_local), torch.device("cuda", device)
        h = C.c_void_p()
        check(self.lib.mansy_peer_create(self.world, self.rank, self.n_local * _capi.TOTALS_DOUBLES * 8, device, C.byref(h)))
        self._h = h
        if self.world > 1:
            mine = (C.c_uint8 * _capi.PEER_HANDLE_BYTES)()
            check(self.lib.mansy_peer_export(self._h, mine))
            every = [None] * self.world
            dist.all_gather_object(every, bytes(mine), group=group)
            blob = (C.c_uint8 * (_capi.PEER_HANDLE_BYTES * self.world)).from_buffer_copy(b"".join(every))
            # every rank has mapped every mailbox before the first store -- or every rank raises: a rank that cannot
            # open a peer's handle (no P2P route) must not leave the others waiting on its epoch flags
            rc = self.lib.mansy_peer_connect(self._h, blob)
            why = self.lib.mansy_last_error().decode() if rc else ""
            failed = [None] * self.world
            dist.all_gather_object(failed, (self.rank, why) if rc else None, group=group)
            failed = [f for f in failed if f is not None]
            if failed:
                self.close()
                raise RuntimeError(f"peer-memory mailboxes unavailable on rank(s) {[f[0] for f in failed]}: {failed[0][1]}")

    def barrier(self, stream: Optional[int] = None) -> None:
        """Device-side barrier over the group's GPUs on ``stream`` (no host synchronisation)."""
        check(self.lib.mansy_peer_barrier(self._h, stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream))

    def gather_episode_stats(self, sim: BatchSimulator) -> torch.Tensor:
        """``[world * N_local, 6]`` float64 in rank order, a zero-copy view of this rank's mailbox (valid until the
        gather after the next one)."""
        ptr = C.c_void_p()
        check(self.lib.mansy_peer_allgather_stats(self._h, sim._h, sim._stream(), C.byref(ptr)))
        return torch.as_tensor(_DevArray(ptr.value, (self.world * self.n_local, 6), "<f8", self), device=self.device)

    def timed_out(self) -> bool:
        f = C.c_int32(0)
        check(self.lib.mansy_peer_timed_out(self._h, C.byref(f)))
        return bool(f.value)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.mansy_peer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_episode_stats(sim: BatchSimulator, group=None, peers: Optional[PeerGroup] = None,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The one exchange of a rollout: all-gather ``[N_local, 6]`` float64 episode totals (sum qoe, qoe1, qoe2, qoe3,
    steps, episodes per env) into ``[N_global, 6]`` on every rank.  With ``peers`` it is one kernel over NVLink peer
    memory (:class:`PeerGroup`); otherwise the totals are packed by one kernel and, when a process group is up,
    exchanged with ``torch.distributed.all_gather_into_tensor`` into the preallocated ``out`` (NCCL on GPUs)."""
    if peers is not None:
        return peers.gather_episode_stats(sim)
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return sim.episode_totals(out)
    local = sim.episode_totals()
    if out is None:
        out = torch.empty((dist.get_world_size(group) * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out


def summarise_stats(stats: torch.Tensor) -> Dict[str, float]:
    """Rollout summary from gathered stats: mean per-step QoE terms and episode count."""
    tot = stats.sum(dim=0).tolist()
    steps = max(tot[4], 1.0)
    return {"mean_qoe": tot[0] / steps, "mean_qoe1": tot[1] / steps, "mean_qoe2": tot[2] / steps,
            "mean_qoe3": tot[3] / steps, "steps": tot[4], "episodes": tot[5]}


def shard_range(n_global: int, rank: int, world: int):
    """Contiguous env-index range of ``rank``."""
    per = n_global // world
    if per * world != n_global:
        raise ValueError("n_global must divide evenly over the ranks")
    return rank * per, per
