"""``SimpleRLEnv`` with the reference's interface (bitrate_selection/envs/simple_rl_env.py), CUDA-backed."""
from __future__ import annotations

from ...config import OBS_MODE_SIMPLE, REWARD_QOE, REWARD_QOE_NORM
from ...vector_env import SingleEnv
from ._common import device_index, tables_for


class SimpleRLEnv(SingleEnv):
    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 50}

    def __init__(self, config, dataset, network_dataset, qoe_weights, log_path, startup_download, mode='train',
                 seed=0, worker_num=1, device='cpu'):
        assert mode in ['train', 'valid', 'test']                       # simple_rl_env.py:18
        self.config, self.dataset, self.network_dataset = config, dataset, network_dataset
        self.qoe_weights, self.log_path, self.startup_download, self.mode = qoe_weights, log_path, startup_download, mode
        tables = tables_for(config, dataset, network_dataset, qoe_weights, mode, startup_download)
        reward_mode = REWARD_QOE_NORM if mode == 'train' else REWARD_QOE     # simple_rl_env.py:124-127
        super().__init__(tables, OBS_MODE_SIMPLE, reward_mode, log_path, seed=seed, worker_num=worker_num,
                         device=device_index(device))
        self.videos, self.users, self.traces = tables.video_ids, tables.user_ids, tables.trace_ids
        self.samples = tables.samples
        self.sample_len = tables.n_samples
