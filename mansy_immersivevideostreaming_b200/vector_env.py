"""Vectorised environment with the reference's gym / tianshou-facing behaviour.

``B200VectorEnv`` duck-types tianshou 0.4.8's ``BaseVectorEnv`` as the reference's scripts use it
(bitrate_selection/run_mansy.py:44-56, run_simple_rl.py:38-50): ``len()``, ``reset(id)``,
``step(action, id)``, ``seed``, ``close``, and the Collector-driven reset protocol (the venv
returns the terminal observation; the caller resets the finished ids).  tianshou itself is
pinned by the reference (README.md:21) but is neither vendored nor installed here, so this
surface follows its documented contract -- "parity unpinned" at that boundary.

Per-episode CSV rows follow ``MANSYEnv._log`` (envs/mansy_env.py:271-290) so the reference's
``read_log_file`` (utils/common.py:196-218) keeps working.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .config import OBS_MODE_MANSY, OBS_MODE_NONE, OBS_MODE_SIMPLE, REWARD_QOE, SimConfig
from .simulator import BatchSimulator, obs_segments_for, obs_views
from .tables import SimTables

LOG_HEADER = "video,user,trace,qoe_w1,qoe_w2,qoe_w3,qoe,qoe1,qoe2,qoe3\n"


class _Discrete:
    """Stand-in for gym.spaces.Discrete when gym is not installed."""

    def __init__(self, n: int):
        self.n = int(n)

    def sample(self) -> int:
        return int(np.random.randint(self.n))


def episode_log_line(tables: SimTables, sample_id: int, sum_qoe: float, sum_q1: float, sum_q2: float, sum_q3: float,
                     steps: int) -> str:
    """One ``_log`` row (envs/mansy_env.py:278-285)."""
    v, u, t, q = (int(x) for x in tables.samples[sample_id])
    w = tables.qoe_w[q]
    wsum = float(w[0]) + float(w[1]) + float(w[2])
    qoe = round(sum_qoe / steps / wsum, 5)
    return (f"{int(tables.video_ids[v])},{int(tables.user_ids[u])},{int(tables.trace_ids[t])},"
            f"{w[0]},{w[1]},{w[2]},{qoe},{round(sum_q1 / steps, 5)},{round(sum_q2 / steps, 5)},{round(sum_q3 / steps, 5)}\n")


class B200VectorEnv:
    def __init__(self, tables: SimTables, env_num: int, obs_mode: int = OBS_MODE_MANSY, reward_mode: int = REWARD_QOE,
                 seed: int = 0, worker_num: Optional[int] = None, log_path: Optional[str] = None, device: int = 0,
                 output: str = "numpy", env_offset: int = 0):
        assert output in ("numpy", "torch")
        if obs_mode == OBS_MODE_NONE:
            raise ValueError("a vector env materialises observations; use BatchSimulator for the fused path")
        self.sim = BatchSimulator(tables, env_num, obs_mode, reward_mode, seed=seed, worker_num=worker_num,
                                  env_offset=env_offset, device=device)
        self.tables = tables
        self.env_num = int(env_num)
        self.obs_mode = obs_mode
        self.output = output
        self.log_path = log_path
        self.is_async = False
        n_act = tables.cfg.action_space if obs_mode == OBS_MODE_MANSY else len(tables.cfg.video_rates)
        self.action_space = [_Discrete(n_act) for _ in range(self.env_num)]   # simple_rl_env.py:33 quirk kept
        self._obs = self.sim.new_obs()                                         # [N, stride] current observations
        self._logged = np.zeros(self.env_num, dtype=np.int64)                  # episodes of each env already in the CSV
        self._closed = False

    def __len__(self) -> int:
        return self.env_num

    def sample_count(self) -> int:
        return self.tables.n_samples

    # -- helpers -----------------------------------------------------------
    def _ids(self, id) -> Optional[np.ndarray]:
        """tianshou passes an int, a list or an array of env indices; out-of-range ids raise ``IndexError`` like the
        reference's ``self.workers[i]`` list indexing, duplicates ``ValueError`` (BatchSimulator._ids)."""
        if id is None:
            return None
        if np.isscalar(id):
            id = [id]
        ids = np.asarray(id, dtype=np.int64).reshape(-1)
        if ids.size and (ids.min() < 0 or ids.max() >= self.env_num):
            raise IndexError(f"env id out of range [0, {self.env_num})")
        if np.unique(ids).size != ids.size:
            raise ValueError("env ids must be unique")
        return ids.astype(np.int32)

    def _package(self, rows: torch.Tensor):
        if self.output == "torch":
            return obs_views(rows, self.obs_mode)
        host = rows.cpu().numpy()
        return {k: host[:, off:off + int(np.prod(shape))].reshape((host.shape[0],) + tuple(shape)).copy()
                for k, off, shape in obs_segments_for(self.obs_mode)}

    # -- tianshou-facing API -------------------------------------------------
    def seed(self, seed=None) -> List[int]:
        s = 0 if seed is None else int(seed if np.isscalar(seed) else seed[0])
        self.sim.seed(s)
        return [s + i for i in range(self.env_num)]

    def reset(self, id=None):
        ids = self._ids(id)
        if ids is None:
            self.sim.reset(None, self._obs)
            return self._package(self._obs)
        rows = self.sim.reset(ids)
        self._obs[torch.as_tensor(ids, dtype=torch.long, device=self._obs.device)] = rows
        return self._package(rows)

    def step(self, action, id=None):
        ids = self._ids(id)
        act = torch.as_tensor(np.asarray(action, dtype=np.int32).reshape(-1)) if not isinstance(action, torch.Tensor) else action
        numpy_out = self.output == "numpy"
        # numpy callers get the reward as the float64 the reference returns (qoe.py:33), not its float32 rounding
        aux = self.sim.new_aux(self.env_num if ids is None else len(ids)) if numpy_out else None
        rows, rew, done = self.sim.step(act.to(self.sim.device), env_ids=ids, auto_reset=False, aux=aux)
        idx_h = np.arange(self.env_num) if ids is None else ids.astype(np.int64)
        idx = torch.as_tensor(idx_h, dtype=torch.long, device=rows.device)
        self._obs[idx] = rows
        done_h = done.cpu().numpy().astype(bool)
        if done_h.any():
            self._log_finished(idx_h[done_h])
        info = np.array([{"env_id": int(i)} for i in idx_h], dtype=object)
        if not numpy_out:
            return self._package(rows), rew, done.bool(), info
        return self._package(rows), aux[:, 13].cpu().numpy(), done_h, info     # MANSY_AUX_REWARD (0 for an env already finished)

    def choose_actions(self, horizon: int = 4) -> np.ndarray:
        """The MPC expert's action for every environment (``ExpertEnv.choose_action``, envs/expert_env.py:358-422)."""
        a = self.sim.expert_actions(horizon)
        return a if self.output == "torch" else a.cpu().numpy()

    def _log_finished(self, env_ids: Sequence[int]) -> None:
        if self.log_path is None:
            return
        ids = torch.as_tensor(np.asarray(env_ids, dtype=np.int64), device=self.sim.device)
        stats = self.sim.episode_stats()[ids].cpu().numpy()       # rows of the finished envs only
        if not os.path.exists(self.log_path):
            with open(self.log_path, "w", encoding="utf-8") as fh:
                fh.write(LOG_HEADER)
        with open(self.log_path, "a", encoding="utf-8") as fh:
            for e, s in zip(env_ids, stats):
                # a finished env stepped again before its reset reports done once more without a new episode:
                # one CSV row per FINISHED EPISODE (MANSY_STAT_TOT_EPISODES), like `_log` (mansy_env.py:225)
                if int(s[11]) <= self._logged[int(e)]:
                    continue
                self._logged[int(e)] = int(s[11])
                fh.write(episode_log_line(self.tables, int(s[5]), s[0], s[1], s[2], s[3], int(s[4])))

    def close(self) -> None:
        if not self._closed:
            self.sim.close()
            self._closed = True


class SingleEnv:
    """N = 1 gym-style environment with the reference's old-gym return conventions: ``reset()``
    returns the observation dict only and ``step`` a 4-tuple (envs/mansy_env.py:99,154,248).
    The returned dict is the env's own ``state`` object, updated in place, as in the reference
    (callers such as run_mansy.py:167-168 mutate it)."""

    def __init__(self, tables: SimTables, obs_mode: int, reward_mode: int, log_path: Optional[str], seed: int = 0,
                 worker_num: int = 1, device: int = 0):
        self.tables = tables
        self.worker_num = int(worker_num)
        self._venv = B200VectorEnv(tables, 1, obs_mode, reward_mode, seed=0, worker_num=self.worker_num,
                                   log_path=log_path, device=device, output="numpy")
        self._venv.sim.seed(int(seed) % self.worker_num)
        self.action_space = self._venv.action_space[0]
        self.state: Optional[Dict[str, np.ndarray]] = None
        self.sample_id = -1
        self.current_video = self.current_user = self.current_trace = self.current_qoe_weight = None

    def _after_reset(self) -> None:
        st = self._venv.sim.episode_state_host()
        self.sample_id = int(st["sample_id"][0])
        v, u, t, q = (int(x) for x in self.tables.samples[self.sample_id])
        self.current_video = int(self.tables.video_ids[v])
        self.current_user = int(self.tables.user_ids[u])
        self.current_trace = int(self.tables.trace_ids[t])
        self.current_qoe_weight = self.tables.qoe_w[q].copy()

    def reset(self, seed=None, options=None):
        obs = self._venv.reset()
        squeezed = {k: v[0] for k, v in obs.items()}
        if self.state is None:
            self.state = squeezed
        else:
            self.state.update(squeezed)
        self._after_reset()
        return self.state

    def step(self, action):
        obs, rew, done, _ = self._venv.step([int(action)])
        self.state.update({k: v[0] for k, v in obs.items()})
        return self.state, float(rew[0]), bool(done[0]), {}

    def choose_action(self, horizon: Optional[int] = None) -> int:
        """``ExpertEnv.choose_action`` (envs/expert_env.py:358-422)."""
        return int(self._venv.sim.expert_actions(int(horizon if horizon is not None else getattr(self, "horizon", 4))).cpu()[0])

    def sample_count(self) -> int:
        return self.tables.n_samples

    def seed(self, seed):
        np.random.seed(seed)
        self._venv.sim.seed(int(seed) % self.worker_num)    # mansy_env.py:253-256

    def close(self):
        self._venv.close()
