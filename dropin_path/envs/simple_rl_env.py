"""``envs.simple_rl_env`` of the reference (bitrate_selection/envs/simple_rl_env.py), CUDA-backed."""
from mansy_immersivevideostreaming_b200.dropin.envs.simple_rl_env import SimpleRLEnv  # noqa: F401
