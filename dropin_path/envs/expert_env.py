"""``envs.expert_env`` of the reference (bitrate_selection/envs/expert_env.py), CUDA-backed."""
from mansy_immersivevideostreaming_b200.dropin.envs.expert_env import ExpertEnv  # noqa: F401
