"""CPU restatement of the reference's streaming-simulator hot path -- TEST INFRASTRUCTURE ONLY.

Not part of the product; see ``oracle/__init__.py`` for who may import this and for the parity
pinning status (PINNED against the unmodified reference run in the build container and against
the shipped ground-truth masks; vectors in ``tests/golden/``).

Every function cites the reference file:line it restates (paths relative to the reference root).
The restatement works on the dense :class:`SimTables` instead of re-reading files and is scalar
Python on purpose: it is the checker, not something to be fast.

Two numeric chains are provided:

* ``chain="f64"`` -- the reference under its *pinned* stack (numpy 1.24.3 legacy promotion):
  builtin ``sum`` over float32 arrays accumulates in float64 and the QoE / reward scalars are
  float64.  This is the primary oracle; the CUDA kernels follow it.
* ``chain="f32"`` -- the reference as it executes under numpy >= 2 (NEP 50), i.e. what the
  imported reference computes in this container: the same expressions stay in float32.  It
  exists so the restatement can be compared *bit for bit* with the imported reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# Shared with the product package: only the plain containers the tests hand to both sides (constants of config.yml and the
# dense table arrays) and the mode enums.  Nothing computed: the action table below is the oracle's own restatement.
from mansy_immersivevideostreaming_b200.config import OBS_MODE_MANSY, OBS_MODE_SIMPLE, REWARD_QOE, REWARD_QOE_NORM, SimConfig
from mansy_immersivevideostreaming_b200.tables import SimTables

F32 = np.float32

# bitrate_selection/utils/common.py:101-119: action -> (rate_in, rate_out) version indices, the if-chain written out.
# (The kernels carry their own packed copy, csrc/mansy_core.cuh action_to_rates, and the host package another,
# config.ACTION_TABLE; tests/test_oracle_golden.py checks all of them against the reference-generated allocate_kat.npz.)
ACTION_TABLE: Tuple[Tuple[int, int], ...] = ((1, 0), (2, 0), (3, 0), (4, 0), (2, 1), (3, 1), (4, 1), (3, 2), (4, 2), (4, 3),
                                             (0, 0), (1, 1), (2, 2), (3, 3), (4, 4))


# ---------------------------------------------------------------------------
# a13-a16: field-of-view -> tile mask
# ---------------------------------------------------------------------------
def find_block(p: int, b: int) -> int:
    """viewport_prediction/utils/common.py:37-43 (one axis): a point on a tile boundary belongs
    to the lower tile, except the origin."""
    k = p // b
    if p > 0 and p % b == 0:
        k -= 1
    return k


def _axis_intervals(lo: int, hi: int, length: int) -> List[Tuple[int, int]]:
    """One axis of viewport_prediction/utils/common.py:83-127.  The nine coded cases are the
    cross product of {inside, wrap-low, wrap-high} on each axis; the order of the regions does
    not matter because they are OR-ed into the mask."""
    if lo >= 0 and hi <= length:
        return [(lo, hi)]
    if lo < 0 and hi <= length:
        return [(0, hi), (lo % length, length)]
    if lo >= 0 and hi > length:
        return [(0, hi % length), (lo, length)]
    # the reference has no branch for a FoV wider than the frame: `regions` stays unbound
    raise ValueError("FoV wraps on both sides of an axis (reference raises UnboundLocalError)")


def fov_tile_mask(x: int, y: int, cfg: SimConfig) -> np.ndarray:
    """viewport_prediction/utils/common.py:46-58 -> uint8[tile_num_height][tile_num_width]."""
    if not (0 <= x <= cfg.video_width and 0 <= y <= cfg.video_height):
        raise ValueError("viewport centre outside the frame (contract: normalised centre in [0,1])")
    mask = np.zeros((cfg.tile_num_height, cfg.tile_num_width), dtype=np.uint8)
    hw, hh = cfg.fov_width // 2, cfg.fov_height // 2
    xs = _axis_intervals(x - hw, x + hw, cfg.video_width)
    ys = _axis_intervals(y - hh, y + hh, cfg.video_height)
    for (x1, x2) in xs:
        for (y1, y2) in ys:
            tx1, tx2 = find_block(x1, cfg.tile_width), find_block(x2, cfg.tile_width)
            ty1, ty2 = find_block(y1, cfg.tile_height), find_block(y2, cfg.tile_height)
            mask[ty1:ty2 + 1, tx1:tx2 + 1] = 1
    return mask


def mask_bits(mask: np.ndarray) -> int:
    out = 0
    for t in np.nonzero(np.asarray(mask).reshape(-1))[0]:
        out |= 1 << int(t)
    return out


def centre_to_pixels(vx, vy, cfg: SimConfig, chain: str = "f64") -> Tuple[int, int]:
    """viewport_prediction/predict.py:40,43: ``int(value * config.video_width)`` with a float32
    value; the product is float64 under the pinned numpy (exact), float32 under numpy >= 2."""
    if chain == "f64":
        return int(float(F32(vx)) * cfg.video_width), int(float(F32(vy)) * cfg.video_height)
    return int(F32(vx) * F32(cfg.video_width)), int(F32(vy) * F32(cfg.video_height))


def chunk_masks(gt_xy: np.ndarray, pred_xy: np.ndarray, cfg: Optional[SimConfig] = None,
                chain: str = "f64") -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """viewport_prediction/predict.py:33-48: per chunk, OR the tile masks of ``frequency`` ground-
    truth points and of ``frequency`` predicted points; accuracy = IoU as float64.

    ``gt_xy``/``pred_xy``: float32 [n_chunks][frequency][2] normalised centres.
    Returns (gt uint64[n], pred uint64[n], acc float64[n]).
    """
    cfg = cfg or SimConfig()
    n = gt_xy.shape[0]
    gt = np.zeros(n, dtype=np.uint64)
    pred = np.zeros(n, dtype=np.uint64)
    acc = np.zeros(n, dtype=np.float64)
    for i in range(n):
        g = p = 0
        for j in range(gt_xy.shape[1]):
            x, y = centre_to_pixels(gt_xy[i, j, 0], gt_xy[i, j, 1], cfg, chain)
            g |= mask_bits(fov_tile_mask(x, y, cfg))
            x, y = centre_to_pixels(pred_xy[i, j, 0], pred_xy[i, j, 1], cfg, chain)
            p |= mask_bits(fov_tile_mask(x, y, cfg))
        gt[i], pred[i] = g, p
        acc[i] = bin(g & p).count("1") / bin(g | p).count("1")
    return gt, pred, acc


# ---------------------------------------------------------------------------
# a1-a2: action -> per-tile bitrate versions
# ---------------------------------------------------------------------------
def action_to_rates(action: int) -> Tuple[int, int]:
    """bitrate_selection/utils/common.py:101-119 (unknown actions keep the initial (0, 0))."""
    if 0 <= int(action) < len(ACTION_TABLE):
        return ACTION_TABLE[int(action)]
    return (0, 0)


def _closest_version(rates: Sequence[int], rate: int) -> int:
    """bitrate_selection/utils/common.py:170-180."""
    best, gap = 0, abs(rates[0] - rate)
    for i in range(len(rates)):
        g = abs(rates[i] - rate)
        if g < gap:
            best, gap = i, g
        elif g == gap and rates[i] < rates[best]:
            best = i
    return best


def tile_scales(pred_bits: int, w: int = 8, h: int = 8) -> np.ndarray:
    """bitrate_selection/utils/common.py:149-168: multi-source breadth-first search from the
    predicted-viewport tiles over the 8 neighbours with wrap-around on both axes; ``scale`` is
    the BFS level.  Restated as level-by-level frontier expansion (same levels as the FIFO
    queue).  Empty mask -> every scale stays 0."""
    scale = np.zeros((h, w), dtype=np.int32)
    seen = np.zeros((h, w), dtype=bool)
    frontier = [(t // w, t % w) for t in range(w * h) if (pred_bits >> t) & 1]
    for r, c in frontier:
        seen[r, c] = True
    level = 0
    while frontier:
        level += 1
        nxt = []
        for r, c in frontier:
            for dr in (-1, 0, 1):
                for dc in (-1, 0, 1):
                    if dr == 0 and dc == 0:
                        continue
                    rr, cc = (r + dr) % h, (c + dc) % w
                    if not seen[rr, cc]:
                        seen[rr, cc] = True
                        scale[rr, cc] = level
                        nxt.append((rr, cc))
        frontier = nxt
    return scale.reshape(-1)


def allocate_tile_versions(rate_in: int, rate_out: int, pred_bits: int, rates: Sequence[int]) -> np.ndarray:
    """bitrate_selection/utils/common.py:142-193 -> tile_rate_versions int32[64] (the
    "chosen-tile indices"; the second return value of the reference is unused by its callers)."""
    scale = tile_scales(pred_bits)
    out = np.zeros(64, dtype=np.int32)
    out[scale == 0] = rate_in
    for s in range(1, int(scale.max()) + 1):
        out[scale == s] = _closest_version(rates, rates[rate_out] // s)
    return out


# ---------------------------------------------------------------------------
# a4-a5: bandwidth trace walk and playback buffer
# ---------------------------------------------------------------------------
def trace_download(size, thr: np.ndarray, length: int, cur_idx: int, cur_time: float) -> Tuple[float, int, float]:
    """bitrate_selection/simulators/network.py:22-35.  Returns (download_time, cur_idx, cur_time).
    Pure Python float (float64) in both numeric chains."""
    start = cur_time
    while size > 0:
        remain = (math.floor(cur_time + 1) - cur_time) * float(thr[cur_idx])
        if size >= remain:
            cur_idx = (cur_idx + 1) % length
            cur_time = math.floor(cur_time + 1)
            size -= remain
        else:
            cur_time += size / float(thr[cur_idx])
            size = 0
    return cur_time - start, cur_idx, cur_time


def buffer_push(buf: float, chunk_length: int, download_time: float) -> Tuple[float, float]:
    """bitrate_selection/simulators/buffer.py:8-15 -> (rebuffer_time, new_buffer)."""
    if download_time > buf:
        return download_time - buf, chunk_length
    return 0.0, buf - download_time + chunk_length


# ---------------------------------------------------------------------------
# a8: QoE
# ---------------------------------------------------------------------------
def qoe_f64(gt_bits: int, tile_q: np.ndarray, rebuffer: float, prev_vq, w: np.ndarray, max_q: int):
    """bitrate_selection/utils/qoe.py:22-34 under numpy 1.24 legacy promotion (App. A.6):
    builtin ``sum`` accumulates float32 terms in float64, left to right."""
    m = [(gt_bits >> t) & 1 for t in range(64)]
    s_m = 0.0
    s_mq = 0.0
    for t in range(64):
        s_mq += float(F32(m[t]) * F32(tile_q[t]))
        s_m += float(m[t])
    vq = s_mq / s_m
    vq32 = F32(vq)                     # float32 array minus float64 scalar: scalar is demoted
    s_dev = 0.0
    for t in range(64):
        s_dev += float(F32(m[t]) * F32(abs(F32(tile_q[t]) - vq32)))
    intra = (s_dev / s_m) / max_q
    vqn = vq / max_q
    inter = abs(vqn - prev_vq) if prev_vq is not None else 0.0
    q1, q2, q3 = vqn, float(rebuffer), intra + inter
    qoe = float(w[0]) * q1 - float(w[1]) * q2 - float(w[2]) * q3
    return qoe, q1, q2, q3, vqn


def qoe_f32(gt_bits: int, tile_q: np.ndarray, rebuffer: float, prev_vq, w: np.ndarray, max_q: int):
    """Same expressions under numpy >= 2 (NEP 50): everything stays float32."""
    m = [F32((gt_bits >> t) & 1) for t in range(64)]
    s_m = F32(0)
    s_mq = F32(0)
    for t in range(64):
        s_mq = F32(s_mq + F32(m[t] * F32(tile_q[t])))
        s_m = F32(s_m + m[t])
    vq = F32(s_mq / s_m)
    s_dev = F32(0)
    for t in range(64):
        s_dev = F32(s_dev + F32(m[t] * F32(abs(F32(F32(tile_q[t]) - vq)))))
    intra = F32(F32(s_dev / s_m) / F32(max_q))
    vqn = F32(vq / F32(max_q))
    inter = F32(abs(F32(vqn - prev_vq))) if prev_vq is not None else F32(0.0)
    q1, q2, q3 = vqn, float(rebuffer), F32(intra + inter)
    qoe = F32(F32(F32(w[0]) * q1) - F32(F32(w[1]) * F32(q2))) - F32(F32(w[2]) * q3)
    return F32(qoe), q1, q2, q3, vqn


# ---------------------------------------------------------------------------
# a3, a6, a7, a9-a12: the environments
# ---------------------------------------------------------------------------
class OracleEnv:
    """One environment: ``MANSYEnv`` (obs_mode=OBS_MODE_MANSY, bitrate_selection/envs/
    mansy_env.py:99-248) or ``SimpleRLEnv`` (OBS_MODE_SIMPLE, envs/simple_rl_env.py:76-160) on
    top of the shared simulator core (simulators/simulator.py:15-108).

    ``worker_id``/``worker_num`` reproduce the sample striding of ``seed``/``reset``
    (mansy_env.py:100-101,253-256): every reset uses ``sample_id = worker_id`` and then advances
    ``worker_id = (worker_id + worker_num) % sample_len``.
    """

    def __init__(self, tables: SimTables, obs_mode: int = OBS_MODE_MANSY, reward_mode: int = REWARD_QOE,
                 chain: str = "f64", worker_id: int = 0, worker_num: int = 1):
        assert chain in ("f64", "f32")
        self.t = tables
        self.cfg = tables.cfg
        self.obs_mode = obs_mode
        self.reward_mode = reward_mode
        self.chain = chain
        self.worker_num = int(worker_num)
        self.worker_id = int(worker_id) % self.worker_num
        self.sample_id = -1
        self.episodes: List[Dict] = []       # finished-episode records (the rows `_log` appends)
        self._log = [[], [], [], []]
        self.state = None
        self.done = True

    # -- simulator getters (simulator.py:48-86) ---------------------------
    def _chunk_tables(self, chunk: int):
        return self.t.size[self.video, chunk], self.t.quality[self.video, chunk]

    def _viewport(self, chunk: int):
        j = chunk - int(self.t.vp_start[self.pair])            # hmdtrace.py:16-19
        return int(self.t.vp_gt[self.pair, j]), int(self.t.vp_pred[self.pair, j]), float(self.t.vp_acc[self.pair, j])

    # -- reset (mansy_env.py:99-152 / simple_rl_env.py:76-111) ------------
    def reset(self):
        cfg = self.cfg
        self.sample_id = self.worker_id
        self.worker_id = (self.worker_id + self.worker_num) % self.t.n_samples
        vi, ui, ti, qi = (int(x) for x in self.t.samples[self.sample_id])
        self.video, self.user, self.trace, self.qoe_idx = vi, ui, ti, qi
        self.pair = vi * self.t.n_users + ui
        self.w = self.t.qoe_w[qi].astype(np.float32)
        # Simulator.__init__ (simulator.py:28-45)
        self.buf = float(cfg.chunk_length * 3)                 # buffer.py:6
        self.cur_idx, self.cur_time = 0, 0.0                   # network.py:19-20
        self.start_chunk = int(self.t.vp_start[self.pair])
        self.end_chunk = min(int(self.t.vp_end[self.pair]), int(self.t.video_time[vi]) - 1)
        self.next_chunk = cfg.startup_download + 1
        self.prev_vq = None                                    # qoe.py:18
        self.ep_step = 0
        self.done = False
        self.gt_bits, self.pred_bits, acc = self._viewport(self.next_chunk)
        self.last_acc = acc
        self.obs_chunk = self.next_chunk                       # chunk whose tables the obs shows
        k = cfg.past_k
        self.h_thr = np.zeros(k, F32); self.h_rin = np.zeros(k, F32); self.h_rout = np.zeros(k, F32)
        self.h_acc = np.zeros(k, F32); self.h_vq = np.zeros(k, F32); self.h_var = np.zeros(k, F32)
        self.h_reb = np.zeros(k, F32)
        self.one_hot = np.zeros(cfg.action_space, F32)
        self.last_bitrates = np.zeros(2, F32)
        self.rebuffer_obs = F32(0)
        self.state = self._obs()
        return self.state

    def _obs(self) -> Dict[str, np.ndarray]:
        cfg = self.cfg
        size, quality = self._chunk_tables(self.obs_chunk)
        size_n = (size.astype(np.float32) / F32(cfg.max_size)).astype(np.float32)       # common.py:45-47
        pred = np.array([(self.pred_bits >> t) & 1 for t in range(64)], dtype=np.float32)
        if self.obs_mode == OBS_MODE_SIMPLE:
            return {"throughput": self.h_thr.reshape(1, -1).copy(), "chunk_sizes": size_n,
                    "rebuffer": np.array([self.rebuffer_obs], F32), "last_bitrates": self.last_bitrates.copy(),
                    "pred_viewport": pred}
        qual_n = (quality.astype(np.float32) / F32(cfg.video_rates[-1])).astype(np.float32)  # common.py:40-42
        if self.chain == "f64":   # builtin sum is float64 under legacy promotion, then demoted by the array op
            wsum = F32(float(self.w[0]) + float(self.w[1]) + float(self.w[2]))
        else:
            wsum = F32(F32(F32(self.w[0]) + F32(self.w[1])) + F32(self.w[2]))
        return {
            "throughput": self.h_thr.reshape(1, -1).copy(),
            "next_chunk_size": size_n,
            "next_chunk_quality": qual_n,
            "pred_viewport": pred.reshape(1, -1),
            "rates_inside": self.h_rin.reshape(1, -1).copy(),
            "rates_outside": self.h_rout.reshape(1, -1).copy(),
            "viewport_acc": self.h_acc.reshape(1, -1).copy(),
            "buffer": np.array([F32(F32(self.buf) / F32(cfg.startup_download))], F32),
            "qoe_weight": (self.w / wsum).astype(np.float32),                          # common.py:55-57
            "action_one_hot": self.one_hot.copy(),
            "past_viewport_qualities": self.h_vq.reshape(1, -1).copy(),
            "past_quality_variances": self.h_var.reshape(1, -1).copy(),
            "past_rebuffering": self.h_reb.reshape(1, -1).copy(),
        }

    @staticmethod
    def _push(h: np.ndarray, v) -> None:
        """np.roll(x, 1); x[0, 0] = v (mansy_env.py:192-206): newest value at index 0."""
        h[1:] = h[:-1].copy()
        h[0] = F32(v)

    # -- step (mansy_env.py:154-248 / simple_rl_env.py:113-160) ------------
    def step(self, action: int):
        assert not self.done, "step() after the episode ended (the reference would index past its tables)"
        cfg = self.cfg
        rates = cfg.video_rates
        rate_in, rate_out = action_to_rates(action)
        versions = allocate_tile_versions(rate_in, rate_out, self.pred_bits, rates)
        size, quality = self._chunk_tables(self.next_chunk)
        # Simulator.simulate_download (simulator.py:88-108)
        chunk_size = int(sum(int(size[versions[t], t]) for t in range(64)))
        tile_q = np.array([quality[versions[t], t] for t in range(64)], dtype=np.float32)
        dl, self.cur_idx, self.cur_time = trace_download(chunk_size, self.t.trace[self.trace],
                                                         int(self.t.trace_len[self.trace]), self.cur_idx, self.cur_time)
        rebuf, self.buf = buffer_push(self.buf, cfg.chunk_length, dl)
        self.next_chunk += 1
        over = self.next_chunk > self.end_chunk
        qoe_fn = qoe_f64 if self.chain == "f64" else qoe_f32
        qoe, q1, q2, q3, self.prev_vq = qoe_fn(self.gt_bits, tile_q, rebuf, self.prev_vq, self.w, rates[-1])

        if self.chain == "f64":
            wsum = float(self.w[0]) + float(self.w[1]) + float(self.w[2])
            reward = qoe / wsum if self.reward_mode == REWARD_QOE_NORM else qoe
        else:
            wsum = F32(F32(F32(self.w[0]) + F32(self.w[1])) + F32(self.w[2]))
            reward = F32(qoe / wsum) if self.reward_mode == REWARD_QOE_NORM else qoe
        for lst, v in zip(self._log, (qoe, q1, q2, q3)):
            lst.append(float(v))

        self.one_hot = np.zeros(cfg.action_space, F32)
        if 0 <= int(action) < cfg.action_space:
            self.one_hot[int(action)] = 1.0
        self._push(self.h_thr, (chunk_size / dl) / cfg.max_throughput)
        self._push(self.h_acc, self.last_acc)
        self._push(self.h_rin, rates[rate_in] / rates[-1])
        self._push(self.h_rout, rates[rate_out] / rates[-1])
        self._push(self.h_vq, q1)
        self._push(self.h_reb, q2 / cfg.startup_download)
        self._push(self.h_var, q3)
        self.rebuffer_obs = F32(q2)                                            # simple_rl_env.py:136
        self.last_bitrates = (np.array([rates[rate_in], rates[rate_out]], F32) / F32(rates[-1])).astype(F32)
        self.ep_step += 1

        aux = {"versions": versions, "chunk_size": chunk_size, "download_time": dl, "rebuffer": rebuf,
               "buffer": self.buf, "qoe": float(qoe), "qoe1": float(q1), "qoe2": float(q2), "qoe3": float(q3),
               "cur_idx": self.cur_idx, "cur_time": self.cur_time, "next_chunk": self.next_chunk,
               "sample_id": self.sample_id, "gt_bits": self.gt_bits, "pred_bits": self.pred_bits}
        if over:
            self.done = True
            self._finish_episode()
        else:
            self.obs_chunk = self.next_chunk
            self.gt_bits, self.pred_bits, self.last_acc = self._viewport(self.next_chunk)
        self.state = self._obs()
        return self.state, reward, over, aux

    # -- episode log row (mansy_env.py:271-290) ----------------------------
    def _finish_episode(self):
        n = len(self._log[0])
        wsum = float(sum(float(x) for x in self.w))
        means = [sum(l) / len(l) for l in self._log]
        self.episodes.append({
            "video": int(self.t.video_ids[self.video]), "user": int(self.t.user_ids[self.user]),
            "trace": int(self.t.trace_ids[self.trace]), "w": tuple(float(x) for x in self.w),
            "qoe": round(means[0] / wsum, 5), "qoe1": round(means[1], 5), "qoe2": round(means[2], 5),
            "qoe3": round(means[3], 5), "steps": n, "sample_id": self.sample_id,
            "sums": tuple(sum(l) for l in self._log),
        })
        for l in self._log:
            l.clear()


def flatten_obs(obs: Dict[str, np.ndarray], obs_mode: int) -> np.ndarray:
    """Pack an oracle/reference observation dict into the product's padded row layout."""
    from mansy_immersivevideostreaming_b200.config import (MANSY_OBS_SEGMENTS, MANSY_OBS_STRIDE,
                                                           SIMPLE_OBS_SEGMENTS, SIMPLE_OBS_STRIDE)
    segs, stride = ((MANSY_OBS_SEGMENTS, MANSY_OBS_STRIDE) if obs_mode == OBS_MODE_MANSY
                    else (SIMPLE_OBS_SEGMENTS, SIMPLE_OBS_STRIDE))
    row = np.zeros(stride, dtype=np.float32)
    for key, off, shape in segs:
        n = int(np.prod(shape))
        row[off:off + n] = np.asarray(obs[key], dtype=np.float32).reshape(-1)
    return row


class OracleVectorEnv:
    """N independent oracle envs with the vector-env sample striding (env k of a vector env
    seeded with s gets worker_id (s+k) % worker_num)."""

    def __init__(self, tables: SimTables, n_envs: int, obs_mode: int = OBS_MODE_MANSY,
                 reward_mode: int = REWARD_QOE, chain: str = "f64", seed: int = 0,
                 worker_num: Optional[int] = None, env_offset: int = 0):
        worker_num = n_envs if worker_num is None else worker_num
        self.envs = [OracleEnv(tables, obs_mode, reward_mode, chain, worker_id=(seed + env_offset + k) % worker_num,
                               worker_num=worker_num) for k in range(n_envs)]
        self.obs_mode = obs_mode

    def reset(self, ids=None):
        ids = range(len(self.envs)) if ids is None else ids
        return np.stack([flatten_obs(self.envs[i].reset(), self.obs_mode) for i in ids])

    def step(self, actions, auto_reset: bool = False):
        rows, rews, dones, auxs = [], [], [], []
        for env, a in zip(self.envs, actions):
            obs, r, d, aux = env.step(int(a))
            if d and auto_reset:
                obs = env.reset()
            rows.append(flatten_obs(obs, self.obs_mode)); rews.append(float(r)); dones.append(d); auxs.append(aux)
        return np.stack(rows), np.asarray(rews, dtype=np.float64), np.asarray(dones, dtype=bool), auxs


# ---------------------------------------------------------------------------
# the MPC expert (bitrate_selection/envs/expert_env.py:358-422)
# ---------------------------------------------------------------------------
def expert_profile(env: "OracleEnv", chunk: int, action: int):
    """One entry of the expert's cache (expert_env.py:121-160, ``chunk_pred_*``): tiles allocated from the PREDICTED
    viewport, quality statistics taken over the ACTUAL viewport (simulator.py:146-158).  Returns
    ``(chunk_size, viewport_quality, intra_viewport_quality_variance)`` (not normalised) in the env's numeric chain."""
    rates = env.cfg.video_rates
    rate_in, rate_out = action_to_rates(action)
    gt_bits, pred_bits, _ = env._viewport(chunk)
    versions = allocate_tile_versions(rate_in, rate_out, pred_bits, rates)
    size, quality = env._chunk_tables(chunk)
    chunk_size = int(sum(int(size[versions[t], t]) for t in range(64)))
    tile_q = np.array([quality[versions[t], t] for t in range(64)], dtype=np.float32)
    if env.chain == "f64":
        m = [(gt_bits >> t) & 1 for t in range(64)]
        s_m = s_mq = 0.0
        for t in range(64):
            s_mq += float(F32(m[t]) * F32(tile_q[t]))
            s_m += float(m[t])
        vq = s_mq / s_m
        vq32 = F32(vq)
        s_dev = 0.0
        for t in range(64):
            s_dev += float(F32(m[t]) * F32(abs(F32(tile_q[t]) - vq32)))
        return chunk_size, vq, s_dev / s_m
    m = [F32((gt_bits >> t) & 1) for t in range(64)]
    s_m = s_mq = F32(0)
    for t in range(64):
        s_mq = F32(s_mq + F32(m[t] * F32(tile_q[t])))
        s_m = F32(s_m + m[t])
    vq = F32(s_mq / s_m)
    s_dev = F32(0)
    for t in range(64):
        s_dev = F32(s_dev + F32(m[t] * F32(abs(F32(F32(tile_q[t]) - vq)))))
    return chunk_size, vq, F32(s_dev / s_m)


def _expert_qoe(env: "OracleEnv", vq, intra, rebuffer: float, prev_vq):
    """``QoEModelExpert.calculate_qoe_with_given_quality`` (utils/qoe.py:50-60)."""
    max_q = env.cfg.video_rates[-1]
    w = env.w
    if env.chain == "f64":
        vqn = vq / max_q
        intra_n = intra / max_q
        inter = abs(vqn - prev_vq) if prev_vq is not None else 0.0
        q3 = intra_n + inter
        return float(w[0]) * vqn - float(w[1]) * float(rebuffer) - float(w[2]) * q3, vqn
    vqn = F32(vq / F32(max_q))
    intra_n = F32(intra / F32(max_q))
    inter = F32(abs(F32(vqn - prev_vq))) if prev_vq is not None else F32(0.0)
    q3 = F32(intra_n + inter)
    qoe = F32(F32(F32(w[0]) * vqn) - F32(F32(w[1]) * F32(rebuffer))) - F32(F32(w[2]) * q3)
    return F32(qoe), vqn


def expert_choose_action(env: "OracleEnv", horizon: int, return_value: bool = False):
    """``ExpertEnv.choose_action`` (expert_env.py:358-422): exhaustive search over action sequences of the next
    ``min(horizon, chunks left)`` chunks on virtual downloads from the current trace / buffer state
    (simulator.py:125-144); the first sequence (lowest index, first digit = first action) with the largest QoE sum wins.
    The reference enumerates ``15 ** horizon`` sequences even when fewer chunks are left; the extra digits are unused
    there and ties resolve to the lowest index, which is the enumeration of ``15 ** H`` done here."""
    H = min(int(horizon), env.end_chunk - env.next_chunk + 1)
    if H <= 0:
        return (0, 0.0) if return_value else 0
    prof = [[expert_profile(env, env.next_chunk + t, a) for a in range(15)] for t in range(H)]
    thr = env.t.trace[env.trace]
    tlen = int(env.t.trace_len[env.trace])
    best, best_i = float("-inf"), 0
    for i in range(15 ** H):
        cur_idx, cur_time, buf, prev = env.cur_idx, env.cur_time, env.buf, env.prev_vq
        qoe_sum = 0
        tmp = i
        for t in range(H):
            a = tmp % 15
            tmp //= 15
            size, vq, intra = prof[t][a]
            dl, cur_idx, cur_time = trace_download(size, thr, tlen, cur_idx, cur_time)
            rebuf, buf = buffer_push(buf, env.cfg.chunk_length, dl)
            qoe, prev = _expert_qoe(env, vq, intra, rebuf, prev)
            qoe_sum = qoe_sum + qoe if env.chain == "f64" else F32(qoe_sum + qoe)
        if best < qoe_sum:
            best, best_i = qoe_sum, i
    return (best_i % 15, float(best)) if return_value else best_i % 15
