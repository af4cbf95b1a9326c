"""Simulator constants of the streaming hot path.

Values mirror the reference's ``config.yml`` (file:line cited per field).  The CUDA
kernels are specialised for the 8x8 tile grid / 5 bitrate versions / past_k 8 /
15 actions of the reference defaults; ``SimConfig.validate`` rejects anything else
loudly instead of silently computing something different.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence, Tuple

# Observation row layouts (float32 words).  Every segment starts on a 16-byte
# boundary so the kernels can use 128-bit stores and torch can hand out zero-copy
# views with the reference's shapes (bitrate_selection/envs/mansy_env.py:136-150).
MANSY_OBS_SEGMENTS: Tuple[Tuple[str, int, Tuple[int, ...]], ...] = (
    # key, offset (floats), per-env shape
    ("throughput", 0, (1, 8)),
    ("next_chunk_size", 8, (5, 64)),
    ("next_chunk_quality", 328, (5, 64)),
    ("pred_viewport", 648, (1, 64)),
    ("rates_inside", 712, (1, 8)),
    ("rates_outside", 720, (1, 8)),
    ("viewport_acc", 728, (1, 8)),
    ("past_viewport_qualities", 736, (1, 8)),
    ("past_quality_variances", 744, (1, 8)),
    ("past_rebuffering", 752, (1, 8)),
    ("action_one_hot", 760, (15,)),
    ("qoe_weight", 776, (3,)),
    ("buffer", 779, (1,)),
)
MANSY_OBS_FLOATS = 779          # payload: 13 arrays, 779 float32 (SURVEY.md quotes 777; the shapes sum to 779)
MANSY_OBS_STRIDE = 784          # padded row: 3136 B = 98 full 32-B sectors

# bitrate_selection/envs/simple_rl_env.py:103-109
SIMPLE_OBS_SEGMENTS: Tuple[Tuple[str, int, Tuple[int, ...]], ...] = (
    ("throughput", 0, (1, 8)),
    ("chunk_sizes", 8, (5, 64)),
    ("pred_viewport", 328, (64,)),
    ("last_bitrates", 392, (2,)),
    ("rebuffer", 394, (1,)),
)
SIMPLE_OBS_FLOATS = 395
SIMPLE_OBS_STRIDE = 400         # 1600 B = 50 full sectors

OBS_MODE_NONE = 0               # fused / no observation materialised
OBS_MODE_MANSY = 1
OBS_MODE_SIMPLE = 2

# reward modes (bitrate_selection/envs/mansy_env.py:168-177, simple_rl_env.py:124-127)
REWARD_QOE = 0                  # reward = qoe
REWARD_QOE_NORM = 1             # reward = qoe / sum(w)

# bitrate_selection/utils/common.py:101-119 -- action -> (rate_in, rate_out); anything
# outside 0..14 keeps the function's initial (0, 0).
ACTION_TABLE: Tuple[Tuple[int, int], ...] = (
    (1, 0), (2, 0), (3, 0), (4, 0), (2, 1), (3, 1), (4, 1), (3, 2), (4, 2), (4, 3),
    (0, 0), (1, 1), (2, 2), (3, 3), (4, 4),
)


@dataclass(frozen=True)
class SimConfig:
    """Constants of one simulator instance (config.yml:68-75,153-157)."""

    tile_num_width: int = 8            # config.yml:68
    tile_num_height: int = 8           # config.yml:69
    video_width: int = 2560            # config.yml:71
    video_height: int = 1440           # config.yml:72
    chunk_length: int = 1              # config.yml:74
    video_rates: Tuple[int, ...] = (1, 5, 8, 16, 35)   # config.yml:75
    startup_download: int = 5          # config.yml:153
    max_size: int = 500000             # config.yml:154
    max_throughput: int = 5000000      # config.yml:155
    past_k: int = 8                    # config.yml:156
    action_space: int = 15             # config.yml:157
    fov_width: int = 600               # viewport_prediction/utils/common.py:47
    fov_height: int = 300
    frequency: int = 5                 # config.yml:149
    trim_head: int = 15                # config.yml:147

    @property
    def tile_total_num(self) -> int:
        return self.tile_num_width * self.tile_num_height

    @property
    def tile_width(self) -> int:        # derived as in viewport_prediction/utils/results.py:36-39
        return self.video_width // self.tile_num_width

    @property
    def tile_height(self) -> int:
        return self.video_height // self.tile_num_height

    def validate(self) -> None:
        if (self.tile_num_width, self.tile_num_height) != (8, 8):
            raise ValueError("sm_100a kernels are specialised for the reference's 8x8 tile grid")
        if len(self.video_rates) != 5:
            raise ValueError("sm_100a kernels are specialised for 5 bitrate versions")
        if self.past_k != 8 or self.action_space != 15:
            raise ValueError("sm_100a kernels are specialised for past_k=8, action_space=15")
        if list(self.video_rates) != sorted(self.video_rates):
            raise ValueError("video_rates must ascend (max quality is video_rates[-1])")

    @classmethod
    def from_reference_config(cls, config) -> "SimConfig":
        """Build from the reference's Munch config (bitrate_selection/utils/common.py:13-37)."""
        return cls(
            tile_num_width=int(config.tile_num_width), tile_num_height=int(config.tile_num_height),
            video_width=int(config.video_width), video_height=int(config.video_height),
            chunk_length=int(config.chunk_length), video_rates=tuple(int(r) for r in config.video_rates),
            startup_download=int(config.startup_download), max_size=int(config.max_size),
            max_throughput=int(config.max_throughput), past_k=int(config.past_k),
            action_space=int(config.action_space),
            frequency=int(getattr(config, "frequency", 5)), trim_head=int(getattr(config, "trim_head", 15)),
        )


def closest_rate_version(rates: Sequence[int], rate: int) -> int:
    """Version whose bitrate is closest to ``rate``; ties go to the lower bitrate.

    Host-side construction of the allocation LUT the kernels index; follows
    bitrate_selection/utils/common.py:170-180.
    """
    best, gap = 0, abs(rates[0] - rate)
    for i, r in enumerate(rates):
        g = abs(r - rate)
        if g < gap or (g == gap and r < rates[best]):
            best, gap = i, g
    return best


def rate_out_lut(rates: Sequence[int], max_scale: int = 4) -> Tuple[Tuple[int, ...], ...]:
    """LUT[rate_out][scale] for scale 1..max_scale (index 0 unused -> rate_out).

    bitrate_selection/utils/common.py:186-190: version closest to
    ``rates[rate_out] // scale``.  On an 8x8 torus the largest Chebyshev distance is 4.
    """
    lut = []
    for ro in range(len(rates)):
        row = [ro]
        for s in range(1, max_scale + 1):
            row.append(closest_rate_version(rates, rates[ro] // s))
        lut.append(tuple(row))
    return tuple(lut)
