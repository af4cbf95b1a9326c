"""``MANSYEnv`` with the reference's interface (bitrate_selection/envs/mansy_env.py), CUDA-backed.

Constructor arguments, ``reset``/``step``/``seed``/``sample_count``/``close`` and the attributes the
reference's scripts touch (``action_space.n``, ``current_video`` ...) keep their meaning; ``reset`` returns
the observation dict only and ``step`` a 4-tuple, like the old-gym reference (mansy_env.py:99,154,248).
"""
from __future__ import annotations

from ...config import OBS_MODE_MANSY, REWARD_QOE, REWARD_QOE_NORM
from ...vector_env import SingleEnv
from ._common import device_index, tables_for


class MANSYEnv(SingleEnv):
    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 50}

    def __init__(self, config, dataset, network_dataset, qoe_weights, identifier, lamb, log_path,
                 startup_download, mode='train', seed=0, worker_num=1, device='cpu', use_identifier=False):
        assert mode in ['train', 'valid', 'test']                       # mansy_env.py:22
        self.config, self.dataset, self.network_dataset = config, dataset, network_dataset
        self.qoe_weights, self.identifier, self.lamb = qoe_weights, identifier, lamb
        self.log_path, self.startup_download, self.mode = log_path, startup_download, mode
        self.use_identifier = use_identifier
        tables = tables_for(config, dataset, network_dataset, qoe_weights, mode, startup_download)
        # mansy_env.py:168-177: plain QoE unless training with the identifier (then qoe / sum(w))
        reward_mode = REWARD_QOE_NORM if (mode == 'train' and use_identifier) else REWARD_QOE
        super().__init__(tables, OBS_MODE_MANSY, reward_mode, log_path, seed=seed, worker_num=worker_num,
                         device=device_index(device))
        self.videos, self.users, self.traces = tables.video_ids, tables.user_ids, tables.trace_ids
        self.samples = tables.samples
        self.sample_len = tables.n_samples
