"""Synthetic datasets shaped like the reference's shipped data.

There is no network for datasets, so benchmarks and parity tests run on synthetic traces.
Two outputs are supported from the same arrays:

* :func:`make_synthetic_tables` -> dense :class:`SimTables` for the CUDA simulator, and
* :func:`write_reference_layout` -> the reference's own on-disk formats (manifest JSON,
  viewport pickles, bandwidth pickles, a ``config.yml``), so the unmodified reference runs on
  *identical* synthetic traces when goldens are generated (oracle/make_golden.py).

Tile masks are part of the hot path (viewport -> tiles, viewport_prediction/utils/common.py:
46-58), so this module does not compute them itself: ``mask_fn`` maps 5 Hz viewport centres to
(gt_mask, pred_mask, accuracy).  The product passes the CUDA kernel
(``ViewportTiler.chunk_masks``); CPU tests pass the oracle.
"""
from __future__ import annotations

import json
import os
import pickle
from typing import Callable, Optional, Sequence, Tuple

import numpy as np

from .config import SimConfig
from .tables import SimTables, environment_samples, environment_test_samples, u64_to_masks

# shipped-data statistics
_BASE_SIZE = np.array([60138, 130403, 146066, 155970, 161012], dtype=np.float64)
_ROW_PROFILE = np.array([0.36, 0.60, 0.80, 1.21, 1.36, 1.41, 1.21, 1.03], dtype=np.float64)
DEFAULT_QOE = ((7, 1, 1), (1, 7, 1), (1, 1, 7), (3, 3, 3))       # config.yml:142

MaskFn = Callable[[np.ndarray, np.ndarray], Tuple[np.ndarray, np.ndarray, np.ndarray]]


def synth_sizes(rng: np.random.Generator, n_videos: int, n_chunks: int, cfg: SimConfig) -> np.ndarray:
    R, TT = len(cfg.video_rates), cfg.tile_total_num
    base = np.interp(np.arange(R), np.linspace(0, R - 1, len(_BASE_SIZE)), _BASE_SIZE)
    rowprof = _ROW_PROFILE[(np.arange(TT) // cfg.tile_num_width) % len(_ROW_PROFILE)]
    rowprof = rowprof / rowprof.mean()
    noise = rng.lognormal(mean=0.0, sigma=0.35, size=(n_videos, n_chunks, R, TT))
    size = np.rint(base[None, None, :, None] * rowprof[None, None, None, :] * noise)
    return np.clip(size, 3800, 990000).astype(np.int32)


def synth_traces(rng: np.random.Generator, n_traces: int, lo: int = 166, hi: int = 758,
                 zero_frac: float = 0.013) -> Tuple[np.ndarray, np.ndarray]:
    """Integer bytes/s traces: log-normal, mean ~3.78e6, median ~3.6e6, ~1.3 % zero seconds."""
    lens = rng.integers(lo, hi + 1, size=n_traces).astype(np.int32)
    L = int(lens.max())
    sigma = np.sqrt(2.0 * np.log(3.78e6 / 3.6e6))
    thr = np.rint(rng.lognormal(mean=np.log(3.6e6), sigma=sigma, size=(n_traces, L)))
    # slow fading so stalls / multi-second downloads happen as in the 4G logs
    fade = np.exp(0.9 * np.sin(np.arange(L)[None, :] / rng.uniform(6, 40, size=(n_traces, 1))
                               + rng.uniform(0, 6.28, size=(n_traces, 1))))
    thr = np.rint(np.clip(thr * fade, 0, 13.9e6))
    thr[rng.random(size=thr.shape) < zero_frac] = 0
    for t in range(n_traces):
        thr[t, lens[t]:] = 0
        if not np.any(thr[t, : lens[t]] > 0):
            thr[t, 0] = 3.6e6
    return thr.astype(np.float64), lens


def synth_viewport_centres(rng: np.random.Generator, n_pairs: int, n_vp_chunks: int, freq: int = 5,
                           step_sigma: float = 0.03, pred_sigma: float = 0.05
                           ) -> Tuple[np.ndarray, np.ndarray]:
    """5 Hz viewport centres on the unit torus: float32 [P][CV][freq][2] for gt and prediction."""
    steps = rng.normal(0.0, step_sigma, size=(n_pairs, n_vp_chunks * freq, 2))
    start = rng.random(size=(n_pairs, 1, 2))
    gt = np.mod(start + np.cumsum(steps, axis=1), 1.0)
    pred = np.mod(gt + rng.normal(0.0, pred_sigma, size=gt.shape), 1.0)
    gt = gt.astype(np.float32).reshape(n_pairs, n_vp_chunks, freq, 2)
    pred = pred.astype(np.float32).reshape(n_pairs, n_vp_chunks, freq, 2)
    # float32 rounding of values just below 1.0 can give exactly 1.0: allowed (closed interval)
    return gt, pred


def make_synthetic_tables(mask_fn: MaskFn, cfg: Optional[SimConfig] = None, n_videos: int = 24,
                          n_users: int = 60, n_chunks: int = 60, n_traces: int = 40,
                          qoe_w: Optional[Sequence[Sequence[float]]] = None, mode: str = "train",
                          seed: int = 20260101, vp_first_chunk: int = 3, vp_last_chunk: int = 56,
                          short_tail_frac: float = 0.06, trace_len_range: Tuple[int, int] = (166, 758),
                          return_centres: bool = False):
    """Dense synthetic tables following SURVEY.md section 8(d).

    Seeds: ``numpy.random.default_rng(seed + table_id)`` with table ids 0 sizes, 1 traces,
    2 viewports, 3 tails.
    """
    cfg = cfg or SimConfig()
    cfg.validate()
    size = synth_sizes(np.random.default_rng(seed + 0), n_videos, n_chunks, cfg)
    quality = np.broadcast_to(np.asarray(cfg.video_rates, dtype=np.float32)[None, None, :, None],
                              size.shape).copy()      # quality == bitrate (dataset_preprocess/video.py:95)
    video_time = np.full(n_videos, n_chunks, dtype=np.int32)
    video_time[8::9] = n_chunks - 2                    # a few shorter videos (config.yml:38-64: 58 s)
    trace, trace_len = synth_traces(np.random.default_rng(seed + 1), n_traces, *trace_len_range)

    P = n_videos * n_users
    CV = vp_last_chunk - vp_first_chunk + 1
    gt_xy, pred_xy = synth_viewport_centres(np.random.default_rng(seed + 2), P, CV, cfg.frequency)
    gt, pred, acc = mask_fn(gt_xy.reshape(P * CV, cfg.frequency, 2), pred_xy.reshape(P * CV, cfg.frequency, 2))
    gt = np.asarray(gt, dtype=np.uint64).reshape(P, CV)
    pred = np.asarray(pred, dtype=np.uint64).reshape(P, CV)
    acc = np.asarray(acc, dtype=np.float64).reshape(P, CV)
    vp_start = np.full(P, vp_first_chunk, dtype=np.int32)
    vp_end = np.full(P, vp_last_chunk, dtype=np.int32)
    rng_tail = np.random.default_rng(seed + 3)
    short = rng_tail.random(P) < short_tail_frac          # shipped data: 92/1440 files end early (8..55)
    lo_end = cfg.startup_download + 3
    vp_end[short] = rng_tail.integers(lo_end, vp_last_chunk, size=int(short.sum()))
    beyond = np.arange(CV)[None, :] > (vp_end - vp_start)[:, None]    # entries a shorter pickle would not hold
    gt[beyond] = 0
    pred[beyond] = 0
    acc[beyond] = 0.0

    qoe = np.asarray(DEFAULT_QOE if qoe_w is None else qoe_w, dtype=np.float32).reshape(-1, 3)
    if mode == "test":
        samples = environment_test_samples(n_videos, n_users, n_traces, qoe.shape[0])
    else:
        samples = environment_samples(n_videos, n_users, n_traces, qoe.shape[0])
    tables = SimTables(cfg=cfg, size=size, quality=quality, video_time=video_time, vp_gt=gt, vp_pred=pred,
                       vp_acc=acc, vp_start=vp_start, vp_end=vp_end, trace=trace, trace_len=trace_len,
                       qoe_w=qoe, samples=samples, n_users=n_users,
                       video_ids=np.arange(1, n_videos + 1), user_ids=np.arange(1, n_users + 1),
                       trace_ids=np.arange(n_traces))
    if return_centres:
        return tables, gt_xy, pred_xy
    return tables


def diverse_qoe_weights(n: int, seed: int = 20260101 + 7) -> np.ndarray:
    """Dirichlet(1,1,1)*9 preference vectors for the "diverse preferences" config."""
    rng = np.random.default_rng(seed)
    return (rng.dirichlet((1.0, 1.0, 1.0), size=n) * 9.0).astype(np.float32)


def per_env_samples(tables: SimTables, n_envs: int, seed: int = 20260101 + 11) -> np.ndarray:
    """Random (video, user, trace, qoe) tuples, one per env (large-N throughput sweeps)."""
    rng = np.random.default_rng(seed)
    s = np.stack([rng.integers(0, tables.n_videos, n_envs), rng.integers(0, tables.n_users, n_envs),
                  rng.integers(0, tables.n_traces, n_envs), rng.integers(0, tables.qoe_w.shape[0], n_envs)],
                 axis=1)
    return s.astype(np.int32)


def synthetic_actions(n_envs: int, step: int, seed: int = 1234, n_actions: int = 15,
                      env_offset: int = 0) -> np.ndarray:
    """Counter-based action stream keyed by (seed, env, step): identical on every side.

    A 64-bit mix (splitmix64 finaliser) of the key, reduced modulo ``n_actions``; the CUDA
    rollout kernel and the C oracle compute the same function.
    """
    env = np.arange(env_offset, env_offset + n_envs, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
             + env * np.uint64(0xBF58476D1CE4E5B9) + np.uint64(step) * np.uint64(0x94D049BB133111EB))
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z % np.uint64(n_actions)).astype(np.int32)


# ---------------------------------------------------------------------------
# Reference on-disk layout writer (formats: SURVEY.md App. B)
# ---------------------------------------------------------------------------
def write_reference_layout(tables: SimTables, root: str, dataset: str = "Synth", network_dataset: str = "SynthNet",
                           mode_splits: Optional[dict] = None) -> str:
    """Write ``tables`` in the reference's formats under ``root`` and return the config.yml path.

    The generated config follows the reference's schema (config.yml) so
    ``utils.common.get_config_from_yml(path)`` loads it unchanged.
    """
    cfg = tables.cfg
    vids = [int(v) for v in tables.video_ids]
    uids = [int(u) for u in tables.user_ids]
    tids = [int(t) for t in tables.trace_ids]
    man_dir = os.path.join(root, "datasets", dataset, "video_manifests")
    vp_dir = os.path.join(root, "datasets", dataset, "viewports", "prediction")
    net_dir = os.path.join(root, "datasets", "network", network_dataset)
    for d in (man_dir, vp_dir, net_dir):
        os.makedirs(d, exist_ok=True)

    for vi, v in enumerate(vids):
        chunks = {}
        for c in range(tables.n_chunks):
            chunks[str(c)] = {"size": tables.size[vi, c].astype(int).tolist(),
                              "quality": [[int(q) for q in row] for row in tables.quality[vi, c]]}
        manifest = {"Video_Time": int(tables.video_time[vi]), "Chunk_Count": tables.n_chunks, "Chunk_Time": 1,
                    "Available_Bitrates": list(cfg.video_rates), "Chunks": chunks}
        with open(os.path.join(man_dir, f"video{v}.json"), "w", encoding="utf-8") as fh:
            json.dump(manifest, fh)

    gt_masks = u64_to_masks(tables.vp_gt)
    pred_masks = u64_to_masks(tables.vp_pred)
    for vi, v in enumerate(vids):
        os.makedirs(os.path.join(vp_dir, f"video{v}"), exist_ok=True)
        for ui, u in enumerate(uids):
            p = vi * tables.n_users + ui
            entries = []
            for j in range(int(tables.vp_end[p] - tables.vp_start[p] + 1)):
                entries.append((int(tables.vp_start[p]) + j, gt_masks[p, j].copy(), pred_masks[p, j].copy(),
                                np.float64(tables.vp_acc[p, j])))
            with open(os.path.join(vp_dir, f"video{v}", f"user{u}.pkl"), "wb") as fh:
                pickle.dump(entries, fh)

    network_info = {}
    for ti, t in enumerate(tids):
        name = f"synth_trace_{t:04d}.pkl"
        network_info[t] = name
        L = int(tables.trace_len[ti])
        row = tables.trace[ti, :L]
        if np.all(row == np.rint(row)):
            data = [(i, int(row[i])) for i in range(L)]
        else:
            data = [(i, float(row[i])) for i in range(L)]
        with open(os.path.join(net_dir, name), "wb") as fh:
            pickle.dump(data, fh)

    splits = mode_splits or {}
    doc = {
        "datasets_base_dir": os.path.join(root, "datasets") + os.sep,
        "raw_datasets_dir": {dataset: f"raw/{dataset}/"},
        "raw_network_datasets_dir": {network_dataset: f"raw_network/{network_dataset}/"},
        "viewport_datasets_dir": {dataset: f"{dataset}/viewports/"},
        "video_datasets_dir": {dataset: f"{dataset}/video_manifests/"},
        "network_datasets_dir": {network_dataset: f"network/{network_dataset}"},
        "results_base_dir": os.path.join(root, "results") + os.sep,
        "vp_results_dir": "viewport_prediction", "bs_results_dir": "bitrate_selection",
        "models_base_dir": os.path.join(root, "models") + os.sep,
        "vp_models_dir": "viewport_prediction", "bs_models_dir": "bitrate_selection",
        "datasets_list": [dataset], "network_datasets_list": [network_dataset],
        "video_num": {dataset: len(vids)}, "user_num": {dataset: len(uids)},
        "tile_num_width": cfg.tile_num_width, "tile_num_height": cfg.tile_num_height,
        "tile_total_num": cfg.tile_total_num, "video_width": cfg.video_width, "video_height": cfg.video_height,
        "chunk_length": cfg.chunk_length, "video_rates": list(cfg.video_rates),
        "network_info": {network_dataset: network_info},
        "network_split": {network_dataset: {m: splits.get("traces", tids) for m in ("train", "valid", "test")}},
        "video_split": {dataset: {m: splits.get("videos", vids) for m in ("train", "valid", "test")}},
        "user_split": {dataset: {m: splits.get("users", uids) for m in ("train", "valid", "test")}},
        "qoe_split": {m: [[float(x) for x in w] for w in tables.qoe_w] for m in ("train", "valid", "test")},
        "trim_head": cfg.trim_head, "trim_tail": 15, "frequency": cfg.frequency, "sample_step": 5,
        "startup_download": cfg.startup_download, "max_size": cfg.max_size,
        "max_throughput": cfg.max_throughput, "past_k": cfg.past_k, "action_space": cfg.action_space,
    }
    import yaml
    path = os.path.join(root, "config.yml")
    with open(path, "w", encoding="utf8") as fh:
        yaml.safe_dump(doc, fh)
    return path
