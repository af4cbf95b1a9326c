"""``envs.mansy_env`` of the reference (bitrate_selection/envs/mansy_env.py), CUDA-backed."""
from mansy_immersivevideostreaming_b200.dropin.envs.mansy_env import MANSYEnv  # noqa: F401
