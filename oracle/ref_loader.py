"""Import the UNMODIFIED reference from /root/reference -- TEST INFRASTRUCTURE, build container only.

The reference is pure Python but needs ``gym``, ``munch`` and ``prettytable``, none of which is
installed here.  Minimal stand-ins for exactly the names the reference
touches are registered in ``sys.modules`` before the import; the reference's own files are used
where they lie and are never copied.  ``/root/reference`` does not exist on the GPU box: nothing
that runs there (``-m gpu`` tests, ``smoke()``, ``bench.py``) may call this module -- it is used
by ``oracle/make_golden.py`` and by container-only cross-check tests that skip when the
reference tree is absent.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MANSY_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "bitrate_selection", "envs"))


def _install_stubs() -> None:
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env:                      # gym.Env: only subclassed, super().__init__() called
            def __init__(self, *a, **k):
                pass

        spaces = types.ModuleType("gym.spaces")

        class Discrete:
            def __init__(self, n):
                self.n = n

        spaces.Discrete = Discrete
        gym.Env = Env
        gym.spaces = spaces
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    if "munch" not in sys.modules:
        munch = types.ModuleType("munch")

        class Munch(dict):              # attribute access over a dict is all the reference uses
            __getattr__ = dict.__getitem__
            __setattr__ = dict.__setitem__

        munch.Munch = Munch
        sys.modules["munch"] = munch
    if "prettytable" not in sys.modules:
        pt = types.ModuleType("prettytable")

        class PrettyTable:
            def __init__(self):
                self.field_names, self.rows = [], []

            def add_row(self, row):
                self.rows.append(row)

            def __str__(self):
                return "\n".join(",".join(map(str, r)) for r in [self.field_names] + self.rows)

        pt.PrettyTable = PrettyTable
        sys.modules["prettytable"] = pt


class _Ref:
    """Handles to the imported reference modules."""


_cached = None


def load_reference() -> "_Ref":
    """Import the reference's bitrate_selection and viewport_prediction modules."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    ref = _Ref()
    bs = os.path.join(REFERENCE_ROOT, "bitrate_selection")
    # The reference uses top-level package names (envs, simulators, utils, models); import the
    # bitrate-selection side first, then re-import `utils.common` of viewport_prediction under a
    # private name so both geometries are reachable.
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k.split(".")[0] in ("envs", "simulators", "utils", "models")}
    sys.path.insert(0, bs)
    try:
        ref.mansy_env = importlib.import_module("envs.mansy_env")
        ref.simple_rl_env = importlib.import_module("envs.simple_rl_env")
        try:                             # needs tqdm (present in this image); only the expert golden uses it
            ref.expert_env = importlib.import_module("envs.expert_env")
        except ImportError:
            ref.expert_env = None
        ref.common = importlib.import_module("utils.common")
        ref.qoe = importlib.import_module("utils.qoe")
        ref.simulator = importlib.import_module("simulators.simulator")
        ref.models_mansy = importlib.import_module("models.mansy")
        ref.models_simple = importlib.import_module("models.simple_rl")
    finally:
        sys.path.remove(bs)
    bs_mods = {k: sys.modules.pop(k) for k in list(sys.modules)
               if k.split(".")[0] in ("envs", "simulators", "utils", "models")}
    ref._bs_mods = bs_mods      # keep alive
    vp_common = os.path.join(REFERENCE_ROOT, "viewport_prediction", "utils", "common.py")
    spec = importlib.util.spec_from_file_location("_ref_vp_common", vp_common)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ref.vp_common = mod
    sys.modules.update(saved)
    _cached = ref
    return ref


def silence_prints():
    """The reference env prints 'Use Identifier: ...' on construction."""
    import contextlib
    import io
    return contextlib.redirect_stdout(io.StringIO())
