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

from typing import Dict, Optional

import torch

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
    def __init__(self, sim: BatchSimulator, policy: PolicyNet, slabs: int, seed: int = 0):
        if sim.obs_mode == OBS_MODE_NONE:
            raise ValueError("the policy consumes observation rows")
        self.sim, self.policy, self.seed = sim, policy, int(seed)
        self.buf = RolloutBuffers(sim, slabs)
        self.t = 0                                   # global step counter (keys the action sampler)
        sim.reset(None, self.buf.obs[0])

    def step(self) -> None:
        """obs[t] -> logits/value -> action -> simulator -> obs[t+1] (3 kernel launches)."""
        b, s = self.buf, self.t % self.buf.slabs
        nxt = (self.t + 1) % b.slabs
        self.policy.forward(b.obs[s], b.logits, b.value[s])
        self.policy.sample(b.logits, self.seed, self.t, self.sim.env_offset, b.actions[s], b.logp[s])
        self.sim.step(b.actions[s], auto_reset=True, obs=b.obs[nxt], reward=b.reward[s], done=b.done[s])
        self.t += 1

    def run(self, n_steps: int) -> None:
        for _ in range(n_steps):
            self.step()


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


def gather_episode_stats(sim: BatchSimulator, group=None) -> torch.Tensor:
    """The one collective of a rollout: all-gather ``[N_local, 6]`` float64 episode statistics
    (sum qoe, qoe1, qoe2, qoe3, steps, episodes per env) into ``[N_global, 6]`` on every rank."""
    return all_gather_stats(sim.episode_stats()[:, list(STAT_COLUMNS)].contiguous(), group)


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
