"""Shared plumbing of the drop-in envs: reference datasets -> device tables (cached per process)."""
from __future__ import annotations

import dataclasses
from typing import Dict, Tuple

from ...config import SimConfig
from ...tables import SimTables, pack_from_reference_layout

_TABLE_CACHE: Dict[Tuple, SimTables] = {}


def tables_for(config, dataset: str, network_dataset: str, qoe_weights, mode: str, startup_download: int) -> SimTables:
    """What ``Simulator.__init__`` re-reads from disk at every reset (simulators/simulator.py:30-38) is
    packed once per (dataset, split, weights) and cached: manifests, viewport pickles, bandwidth pickles."""
    videos = list(config.video_split[dataset][mode])                # envs/mansy_env.py:44-46
    users = list(config.user_split[dataset][mode])
    traces = list(config.network_split[network_dataset][mode])
    qkey = tuple(tuple(float(x) for x in w) for w in qoe_weights)
    key = (config.video_datasets_dir[dataset], config.viewport_datasets_dir[dataset],
           config.network_datasets_dir[network_dataset], tuple(videos), tuple(users), tuple(traces), qkey, mode,
           int(startup_download))
    if key not in _TABLE_CACHE:
        sim_cfg = dataclasses.replace(SimConfig.from_reference_config(config), startup_download=int(startup_download))
        sim_cfg.validate()
        _TABLE_CACHE[key] = pack_from_reference_layout(config, dataset, network_dataset, videos, users, traces,
                                                       [list(w) for w in qkey], mode, sim_cfg=sim_cfg)
    return _TABLE_CACHE[key]


def device_index(device) -> int:
    """The reference passes ``device`` for the identifier net only ('cpu' / 'cuda' / 'cuda:1'); the
    simulator always runs on a GPU, so 'cpu' maps to cuda:0."""
    s = str(device)
    return int(s.split(":")[1]) if s.startswith("cuda:") else 0
