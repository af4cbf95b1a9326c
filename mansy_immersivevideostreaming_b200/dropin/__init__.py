"""Drop-in mirrors of the reference's environment modules, backed by the CUDA simulator.

``dropin/envs/mansy_env.py`` and ``dropin/envs/simple_rl_env.py`` export ``MANSYEnv`` / ``SimpleRLEnv`` with
the reference's constructor signatures and gym-style methods (bitrate_selection/envs/mansy_env.py:19-20,
envs/simple_rl_env.py:15-16), so ``run_mansy.py`` / ``run_simple_rl.py`` keep working when their import line
is pointed here.  There is no CPU fallback: constructing an env without a CUDA device raises.
"""
