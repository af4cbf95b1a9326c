"""Import the UNMODIFIED reference -- TEST INFRASTRUCTURE (tests, smoke(), bench.py's reference / cpu_baseline legs).

The reference is pure Python but needs ``gym``, ``munch`` and ``prettytable``, none of which is
installed here.  Minimal stand-ins for exactly the names the reference
touches are registered in ``sys.modules`` before the import.  Two sources, in this order:

* ``/root/reference`` (build container): the reference's own files, used where they lie -- code AND datasets
  (``oracle/make_golden*.py`` need the datasets);
* ``oracle/_ref/mansy_reference.zip`` (``oracle/make_ref.py``: byte-for-byte members, git-ignored, shipped to the
  GPU box by gpurun): code and the small shipped fixtures only, imported straight from the archive.  This is what
  the GPU box uses, where ``/root/reference`` does not exist.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MANSY_REFERENCE_ROOT", "/root/reference")
ARCHIVE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "mansy_reference.zip")


def reference_available() -> bool:
    """The full reference tree (code + datasets) is present (build container)."""
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "bitrate_selection", "envs"))


def archive_available() -> bool:
    return os.path.isfile(ARCHIVE)


def code_available() -> bool:
    return reference_available() or archive_available()


def code_root() -> str:
    """Directory (or zip path) that holds ``bitrate_selection/`` and ``viewport_prediction/``."""
    if reference_available():
        return REFERENCE_ROOT
    if archive_available():
        return ARCHIVE
    raise RuntimeError(f"reference not found: neither {REFERENCE_ROOT} nor {ARCHIVE} (python -m oracle.make_ref) exists")


def read_member(rel: str) -> bytes:
    """Bytes of a reference file: ``rel`` relative to the reference root, or ``fixtures/<name>`` for the shipped
    example-run files listed in oracle/make_ref.py."""
    if archive_available():
        import zipfile
        with zipfile.ZipFile(ARCHIVE) as z:
            return z.read(rel)
    if reference_available():
        if rel.startswith("fixtures/"):
            from oracle.make_ref import _members
            rel = {name: src for src, name in _members()}[rel]
        with open(os.path.join(REFERENCE_ROOT, rel), "rb") as fh:
            return fh.read()
    raise RuntimeError("reference not available")


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
    root = code_root()
    _install_stubs()
    ref = _Ref()
    ref.root = root
    bs = root + "/bitrate_selection"
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
    mod = types.ModuleType("_ref_vp_common")        # viewport_prediction/utils/common.py under a private name
    mod.__file__ = root + "/viewport_prediction/utils/common.py"
    src = read_member("viewport_prediction/utils/common.py") if root == ARCHIVE else open(mod.__file__, "rb").read()
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    ref.vp_common = mod
    sys.modules.update(saved)
    _cached = ref
    return ref


def silence_prints():
    """The reference env prints 'Use Identifier: ...' on construction."""
    import contextlib
    import io
    return contextlib.redirect_stdout(io.StringIO())
