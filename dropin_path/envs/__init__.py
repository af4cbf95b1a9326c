"""PYTHONPATH-first shadow of the reference's ``envs`` package (bitrate_selection/envs/).

    cd <reference>/bitrate_selection
    PYTHONPATH=<this repo>/dropin_path python run_mansy.py --test ...

The reference's ``envs/`` directory has no ``__init__.py`` (an implicit namespace package), so a *regular* package
of the same name anywhere on ``sys.path`` wins the import even though the script directory comes first:
``run_mansy.py:16`` (``from envs.mansy_env import MANSYEnv``), ``run_simple_rl.py:16`` and ``run_expert.py`` then bind
the CUDA-backed classes without an edit.  ``utils.*``, ``models.*`` and ``simulators.*`` keep resolving to the
reference's own files.  Only the directory holding this package has to be on ``PYTHONPATH``; the repo root is
added here so ``mansy_immersivevideostreaming_b200`` resolves too.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.append(_ROOT)
