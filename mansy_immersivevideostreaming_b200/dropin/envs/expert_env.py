"""``ExpertEnv`` with the reference's interface (bitrate_selection/envs/expert_env.py), CUDA-backed.

``run_expert.py:66-80,91-108`` drives it as ``reset()`` -> ``choose_action()`` -> ``step(action)`` until done.  The
MPC search (expert_env.py:358-422: ``15 ** horizon`` virtual roll-outs per decision in Python) is one kernel launch
(``mansy_expert_actions``); the cache of per-chunk statistics the reference pickles (``cache_path``,
expert_env.py:93-107,121-167) is not needed -- the kernel derives the few entries a decision uses on the fly --
so ``cache_path`` / ``refresh_cache`` / ``demos_dir`` are accepted and ignored.  ``samples`` is the explicit list of
(video, user, trace, qoe) index tuples the reference passes (run_expert.py:49-64); episodes walk it in order
(expert_env.py:217-221).
"""
from __future__ import annotations

import numpy as np

from ...config import OBS_MODE_MANSY, REWARD_QOE
from ...vector_env import SingleEnv
from ._common import device_index, tables_for


class ExpertEnv(SingleEnv):
    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 50}

    def __init__(self, config, dataset, network_dataset, qoe_weights, samples, demos_dir, cache_path, log_path,
                 startup_download, horizon, refresh_cache=True, mode='train', seed=0, device='cpu'):
        self.config, self.dataset, self.network_dataset = config, dataset, network_dataset
        self.qoe_weights, self.demos_dir, self.log_path = qoe_weights, demos_dir, log_path
        self.startup_download, self.horizon, self.mode = startup_download, int(horizon), mode
        tables = tables_for(config, dataset, network_dataset, qoe_weights, mode, startup_download)
        tables = tables.with_samples(np.asarray([tuple(int(x) for x in s) for s in samples], dtype=np.int32).reshape(-1, 4))
        super().__init__(tables, OBS_MODE_MANSY, REWARD_QOE, log_path, seed=0, worker_num=1, device=device_index(device))
        self.videos, self.users, self.traces = tables.video_ids, tables.user_ids, tables.trace_ids
        self.samples = samples
        self.sample_len = tables.n_samples

    def seed(self, seed):                    # expert_env.py:334-337 only seeds numpy (its worker_id is never used)
        np.random.seed(seed)
