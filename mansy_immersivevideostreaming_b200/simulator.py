"""Batched simulator handle: thin Python over the C ABI, with torch owning device memory.

``BatchSimulator`` is N lock-stepped copies of the reference's
``Simulator`` + ``MANSYEnv``/``SimpleRLEnv`` state machine (bitrate_selection/simulators/
simulator.py:15-108, envs/mansy_env.py:99-248, envs/simple_rl_env.py:76-160).  PyTorch is only
used for device buffers and streams; all compute is in ``csrc/libmansy_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _capi
from ._capi import AUX_DOUBLES, STATS_DOUBLES, Cfg, MansyError, Out, Tables, check
from .config import (MANSY_OBS_SEGMENTS, MANSY_OBS_STRIDE, OBS_MODE_MANSY, OBS_MODE_NONE, OBS_MODE_SIMPLE,
                     REWARD_QOE, SIMPLE_OBS_SEGMENTS, SIMPLE_OBS_STRIDE, SimConfig)
from .tables import SimTables


def _np_ptr(a: np.ndarray) -> int:
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def _require_cuda(device: int) -> None:
    if not torch.cuda.is_available():
        raise MansyError("no CUDA device: the B200 simulator has no CPU fallback")
    if device >= torch.cuda.device_count():
        raise MansyError(f"cuda:{device} does not exist")


def obs_stride_for(obs_mode: int) -> int:
    return {OBS_MODE_MANSY: MANSY_OBS_STRIDE, OBS_MODE_SIMPLE: SIMPLE_OBS_STRIDE, OBS_MODE_NONE: 0}[obs_mode]


def obs_segments_for(obs_mode: int):
    return {OBS_MODE_MANSY: MANSY_OBS_SEGMENTS, OBS_MODE_SIMPLE: SIMPLE_OBS_SEGMENTS, OBS_MODE_NONE: ()}[obs_mode]


def obs_views(obs: torch.Tensor, obs_mode: int) -> Dict[str, torch.Tensor]:
    """Zero-copy views of packed observation rows ``[..., stride]`` with the reference's per-key
    shapes (envs/mansy_env.py:136-150): e.g. ``throughput`` -> ``[..., 1, 8]``."""
    out = {}
    lead = obs.shape[:-1]
    for key, off, shape in obs_segments_for(obs_mode):
        n = int(np.prod(shape))
        out[key] = obs[..., off:off + n].reshape(*lead, *shape)
    return out


class BatchSimulator:
    def __init__(self, tables: SimTables, n_envs: int, obs_mode: int = OBS_MODE_MANSY, reward_mode: int = REWARD_QOE,
                 seed: int = 0, worker_num: Optional[int] = None, env_offset: int = 0, device: int = 0):
        _require_cuda(device)
        self.lib = _capi.load_library()
        self.tables = tables
        self.cfg: SimConfig = tables.cfg
        self.n_envs = int(n_envs)
        self.obs_mode = obs_mode
        self.reward_mode = reward_mode
        self.device_index = device
        self.device = torch.device("cuda", device)
        self.obs_stride = obs_stride_for(obs_mode)
        self.env_offset = int(env_offset)
        self.worker_num = int(self.n_envs if worker_num is None else worker_num)

        t = Tables()
        keep = []           # host arrays must outlive mansy_create
        for name in ("size", "quality", "video_time", "vp_gt", "vp_pred", "vp_acc", "vp_start", "vp_end", "trace",
                     "trace_len", "qoe_w", "samples"):
            arr = np.ascontiguousarray(getattr(tables, name))
            keep.append(arr)
            setattr(t, name, _np_ptr(arr))
        t.n_videos, t.n_chunks = tables.n_videos, tables.n_chunks
        t.n_users, t.n_vp_chunks = tables.n_users, tables.n_vp_chunks
        t.n_traces, t.trace_stride = tables.n_traces, tables.trace.shape[1]
        t.n_qoe, t.n_samples = tables.qoe_w.shape[0], tables.n_samples
        c = Cfg()
        c.n_envs, c.env_offset, c.worker_num, c.seed = self.n_envs, self.env_offset, self.worker_num, int(seed)
        c.obs_mode, c.reward_mode = obs_mode, reward_mode
        for i, r in enumerate(self.cfg.video_rates):
            c.video_rates[i] = int(r)
        c.startup_download, c.chunk_length = self.cfg.startup_download, self.cfg.chunk_length
        c.max_size, c.max_throughput = self.cfg.max_size, self.cfg.max_throughput
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.mansy_create(C.byref(t), C.byref(c), device, C.byref(handle)))
        self._h = handle

    # ------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.mansy_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def new_obs(self, rows: Optional[int] = None) -> torch.Tensor:
        return torch.empty((self.n_envs if rows is None else rows, self.obs_stride), dtype=torch.float32,
                           device=self.device)

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
        return None if t is None else t.data_ptr()

    def _ids(self, env_ids) -> Optional[torch.Tensor]:
        """Device int32 ids of the envs a call touches.  Ids must be unique and in ``[0, n_envs)``: the kernels
        index state / history / statistics rows with them (two lane groups on one id would race on its 128-byte
        state line), so the host side raises ``IndexError`` / ``ValueError`` where the reference's list indexing would."""
        if env_ids is None:
            return None
        host = env_ids.detach().cpu().numpy() if isinstance(env_ids, torch.Tensor) else np.asarray(env_ids)
        host = host.reshape(-1).astype(np.int64)
        if host.size > self.n_envs:
            raise ValueError(f"{host.size} env ids for {self.n_envs} environments")
        if host.size and (host.min() < 0 or host.max() >= self.n_envs):
            raise IndexError(f"env id out of range [0, {self.n_envs})")
        if np.unique(host).size != host.size:
            raise ValueError("env ids must be unique")
        return torch.as_tensor(host.astype(np.int32), device=self.device).contiguous()

    # ------------------------------------------------------------------
    def seed(self, seed: int) -> None:
        """envs/mansy_env.py:253-256 for a vector env: env k gets worker_id (seed+k) % worker_num."""
        check(self.lib.mansy_seed(self._h, int(seed), self._stream()))

    def reset(self, env_ids=None, obs: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        ids = self._ids(env_ids)
        n = self.n_envs if ids is None else int(ids.numel())
        if obs is None and self.obs_mode != OBS_MODE_NONE:
            obs = self.new_obs(n)
        if n == 0:                 # an empty id list touches nothing (and has no device pointer to pass)
            return obs
        check(self.lib.mansy_reset(self._h, self._ptr(ids), n, self._ptr(obs),
                                   obs.stride(0) if obs is not None else 0, self._stream()))
        return obs

    def step(self, actions: torch.Tensor, env_ids=None, auto_reset: bool = False,
             obs: Optional[torch.Tensor] = None, reward: Optional[torch.Tensor] = None,
             done: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
             versions: Optional[torch.Tensor] = None, materialise_obs: bool = True):
        """One lock-step chunk-step.  ``actions`` int32 on the device.  Returns (obs, reward, done)."""
        ids = self._ids(env_ids)
        n = self.n_envs if ids is None else int(ids.numel())
        actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
        if actions.numel() != n:
            raise ValueError("one action per stepped env")
        if obs is None and materialise_obs and self.obs_mode != OBS_MODE_NONE:
            obs = self.new_obs(n)
        if reward is None:
            reward = torch.empty(n, dtype=torch.float32, device=self.device)
        if done is None:
            done = torch.empty(n, dtype=torch.uint8, device=self.device)
        if n == 0:                 # an empty id list touches nothing (and has no device pointer to pass)
            return obs, reward, done
        o = Out()
        o.obs, o.obs_stride = self._ptr(obs), (obs.stride(0) if obs is not None else 0)
        o.reward, o.done = self._ptr(reward), self._ptr(done)
        o.aux, o.tile_versions = self._ptr(aux), self._ptr(versions)
        check(self.lib.mansy_step(self._h, actions.data_ptr(), self._ptr(ids), n, 1 if auto_reset else 0,
                                  C.byref(o), self._stream()))
        return obs, reward, done

    def new_aux(self, rows: Optional[int] = None) -> torch.Tensor:
        return torch.zeros((self.n_envs if rows is None else rows, AUX_DOUBLES), dtype=torch.float64, device=self.device)

    def new_versions(self, rows: Optional[int] = None) -> torch.Tensor:
        return torch.zeros((self.n_envs if rows is None else rows, 64), dtype=torch.uint8, device=self.device)

    def rollout_random(self, n_steps: int, seed: int = 1234, step0: int = 0, obs: Optional[torch.Tensor] = None,
                       reward: Optional[torch.Tensor] = None, done: Optional[torch.Tensor] = None,
                       aux: Optional[torch.Tensor] = None, per_step_outputs: bool = False) -> None:
        """``n_steps`` steps inside ONE kernel launch with hashed in-kernel actions and auto-reset
        (state stays in registers between steps).  Buffers are ``[T, N, ...]`` when
        ``per_step_outputs`` else ``[N, ...]`` (overwritten every step)."""
        o = Out()
        if obs is not None:
            o.obs, o.obs_stride = obs.data_ptr(), obs.stride(-2)
        o.reward, o.done, o.aux = self._ptr(reward), self._ptr(done), self._ptr(aux)
        check(self.lib.mansy_rollout_random(self._h, int(n_steps), int(seed), int(step0),
                                            1 if per_step_outputs else 0, C.byref(o), self._stream()))

    def episode_stats(self) -> torch.Tensor:
        """[N, 16] float64: last finished episode + totals (include/mansy_b200.h MANSY_STAT_*)."""
        out = torch.empty((self.n_envs, STATS_DOUBLES), dtype=torch.float64, device=self.device)
        check(self.lib.mansy_episode_stats(self._h, out.data_ptr(), self._stream()))
        return out

    def episode_totals(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[N, 6] float64 running totals (sum qoe, qoe1, qoe2, qoe3, steps, episodes): the per-rollout send buffer,
        packed by one small kernel (``mansy_episode_totals``)."""
        if out is None:
            out = torch.empty((self.n_envs, _capi.TOTALS_DOUBLES), dtype=torch.float64, device=self.device)
        check(self.lib.mansy_episode_totals(self._h, out.data_ptr(), self._stream()))
        return out

    def set_outcome_table(self, enable: bool) -> None:
        """Steps read chunk bytes / viewport quality / intra-chunk variance of (pair, chunk, action) from the table built
        at create (default) or gather them from the size / quality tables again (``mansy_set_outcome_table``)."""
        check(self.lib.mansy_set_outcome_table(self._h, 1 if enable else 0))

    def stats_clear(self) -> None:
        check(self.lib.mansy_stats_clear(self._h, self._stream()))

    def episode_state_host(self) -> np.ndarray:
        """Structured numpy view of the per-env state records (``_capi.ENV_STATE_FIELDS``)."""
        raw = torch.empty((self.n_envs, 128), dtype=torch.uint8, device=self.device)
        check(self.lib.mansy_state_snapshot(self._h, raw.data_ptr(), self._stream()))
        return raw.cpu().numpy().view(np.dtype(_capi.ENV_STATE_FIELDS)).reshape(self.n_envs)

    def expert_actions(self, horizon: int = 4, actions: Optional[torch.Tensor] = None, return_value: bool = False):
        """``ExpertEnv.choose_action`` (envs/expert_env.py:358-422) for every environment at its current state: the
        first action of the best of the ``15 ** horizon`` sequences over the next chunks (``mansy_expert_actions``).
        Returns int32 ``[N]`` on the device (and the winning QoE sums, float64 ``[N]``, with ``return_value``)."""
        if actions is None:
            actions = torch.empty(self.n_envs, dtype=torch.int32, device=self.device)
        value = torch.empty(self.n_envs, dtype=torch.float64, device=self.device) if return_value else None
        check(self.lib.mansy_expert_actions(self._h, int(horizon), actions.data_ptr(), self._ptr(value), self._stream()))
        return (actions, value) if return_value else actions

    def error_flag(self) -> int:
        flag = C.c_int32(0)
        check(self.lib.mansy_error_flag(self._h, C.byref(flag)))
        return int(flag.value)

    # ---- host-buffer path (what a numpy-facing caller uses; bench.py's e2e) ----------------
    def make_host_buffers(self) -> Dict[str, torch.Tensor]:
        n = self.n_envs
        return {
            "actions": torch.empty(n, dtype=torch.int32).pin_memory(),
            "obs": torch.empty((n, max(self.obs_stride, 1)), dtype=torch.float32).pin_memory(),
            "reward": torch.empty(n, dtype=torch.float32).pin_memory(),
            "done": torch.empty(n, dtype=torch.uint8).pin_memory(),
        }

    def step_host(self, host: Dict[str, torch.Tensor], auto_reset: bool = True) -> None:
        """H2D actions -> step -> D2H obs/reward/done, synchronised (mansy_step_host)."""
        obs_ptr = host["obs"].data_ptr() if self.obs_mode != OBS_MODE_NONE else None
        check(self.lib.mansy_step_host(self._h, host["actions"].data_ptr(), 1 if auto_reset else 0, obs_ptr,
                                       host["reward"].data_ptr(), host["done"].data_ptr(), self._stream()))

    def reset_host(self, host: Dict[str, torch.Tensor]) -> None:
        obs_ptr = host["obs"].data_ptr() if self.obs_mode != OBS_MODE_NONE else None
        check(self.lib.mansy_reset_host(self._h, obs_ptr, self._stream()))


class ViewportTiler:
    """Field-of-view -> tile masks on the GPU (viewport_prediction/utils/common.py:46-58 and the
    per-chunk OR + IoU of viewport_prediction/predict.py:33-48)."""

    def __init__(self, cfg: Optional[SimConfig] = None, device: int = 0):
        _require_cuda(device)
        self.lib = _capi.load_library()
        self.cfg = cfg or SimConfig()
        self.cfg.validate()
        self.device = torch.device("cuda", device)

    def chunk_masks_device(self, gt_xy: torch.Tensor, pred_xy: Optional[torch.Tensor] = None):
        gt_xy = gt_xy.to(device=self.device, dtype=torch.float32).contiguous()
        n, points = gt_xy.shape[0], gt_xy.shape[1]
        gt = torch.empty(n, dtype=torch.int64, device=self.device)
        pred = acc = None
        if pred_xy is not None:
            pred_xy = pred_xy.to(device=self.device, dtype=torch.float32).contiguous()
            pred = torch.empty(n, dtype=torch.int64, device=self.device)
            acc = torch.empty(n, dtype=torch.float64, device=self.device)
        c = self.cfg
        check(self.lib.mansy_viewport_tiles(gt_xy.data_ptr(), None if pred_xy is None else pred_xy.data_ptr(), n,
                                            points, c.video_width, c.video_height, c.fov_width, c.fov_height,
                                            gt.data_ptr(), None if pred is None else pred.data_ptr(),
                                            None if acc is None else acc.data_ptr(),
                                            torch.cuda.current_stream(self.device).cuda_stream))
        return gt, pred, acc

    def chunk_masks(self, gt_xy: np.ndarray, pred_xy: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """numpy in / numpy out (uint64 masks, float64 IoU): the ``mask_fn`` of synth.make_synthetic_tables."""
        gt, pred, acc = self.chunk_masks_device(torch.from_numpy(np.ascontiguousarray(gt_xy)),
                                                torch.from_numpy(np.ascontiguousarray(pred_xy)))
        return (gt.cpu().numpy().view(np.uint64), pred.cpu().numpy().view(np.uint64), acc.cpu().numpy())

    def allocate_tile_versions(self, masks: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
        """bitrate_selection/utils/common.py:101-119,142-193 for standalone masks -> uint8 [n, 64]."""
        masks = masks.to(device=self.device, dtype=torch.int64).contiguous()
        actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
        out = torch.empty((masks.numel(), 64), dtype=torch.uint8, device=self.device)
        rates = (C.c_int32 * 5)(*self.cfg.video_rates)
        check(self.lib.mansy_allocate_tile_versions(masks.data_ptr(), actions.data_ptr(), masks.numel(),
                                                    C.byref(rates), out.data_ptr(),
                                                    torch.cuda.current_stream(self.device).cuda_stream))
        return out
