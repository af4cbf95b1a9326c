from .mansy_env import MANSYEnv            # noqa: F401
from .simple_rl_env import SimpleRLEnv     # noqa: F401
from .expert_env import ExpertEnv          # noqa: F401
