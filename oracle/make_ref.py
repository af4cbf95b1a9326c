"""Pack the UNMODIFIED reference sources of the hot path into ``oracle/_ref/`` -- TEST INFRASTRUCTURE.

    python -m oracle.make_ref            # build container only (needs /root/reference)

The reference is pure Python, so "building" it for the GPU box is a byte-for-byte copy: the files are read
where they lie under ``/root/reference`` and stored (uncompressed members, sha256 in ``MANIFEST.json``) in
``oracle/_ref/mansy_reference.zip``.  ``oracle/_ref/`` is git-ignored (nothing of the reference enters the
history) but not gpurun-ignored, so the archive travels to the GPU box like the built ``.so``.  Python
imports straight from the archive (``oracle/ref_loader.py``); only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s reference / ``cpu_baseline`` legs may use it, never the product path.

Members:
  bitrate_selection/{envs,simulators,utils,models}/*.py, run_mansy.py, run_simple_rl.py   
  viewport_prediction/utils/common.py                                                     (a13-a15)
  config.yml
  fixtures/{best_policy.pth,best_identifier.pth,train_log.csv,valid_log.csv,results.csv}  
"""
from __future__ import annotations

import glob
import hashlib
import json
import os
import sys
import zipfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ROOT = os.environ.get("MANSY_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(ROOT, "oracle", "_ref")
ARCHIVE = os.path.join(OUT_DIR, "mansy_reference.zip")
MANIFEST = os.path.join(OUT_DIR, "MANIFEST.json")

_RUN = "epochs_1_bs_512_lr_0.0005_gamma_0.95_seed_5_ent_0.02_useid_True_lambda_0.5_ilr_0.0001_iur_2_bc_False"
_MODELS = f"models/bitrate_selection/mansy/Jin2022_4G/qoe0_1_2_3/{_RUN}"
_RESULTS = f"results/bitrate_selection/mansy/Jin2022_4G/seen_qoe0_1_2_3/{_RUN}"


def _members():
    """(path inside the reference tree, member name in the archive)."""
    out = []
    for sub in ("envs", "simulators", "utils", "models"):
        for p in sorted(glob.glob(os.path.join(REFERENCE_ROOT, "bitrate_selection", sub, "*.py"))):
            rel = os.path.relpath(p, REFERENCE_ROOT)
            out.append((rel, rel))
    for rel in ("bitrate_selection/run_mansy.py", "bitrate_selection/run_simple_rl.py",
                "viewport_prediction/utils/common.py", "config.yml"):
        out.append((rel, rel))
    for name in ("best_policy.pth", "best_identifier.pth", "train_log.csv", "valid_log.csv"):
        out.append((f"{_MODELS}/{name}", f"fixtures/{name}"))
    out.append((f"{_RESULTS}/results.csv", "fixtures/results.csv"))
    return out


def reference_present() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "bitrate_selection", "envs"))


def build_ref(force: bool = False) -> str:
    """Write the archive (idempotent: skipped when the manifest already matches the reference tree)."""
    if not reference_present():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    members = _members()
    digests = {}
    for src, name in members:
        with open(os.path.join(REFERENCE_ROOT, src), "rb") as fh:
            digests[name] = hashlib.sha256(fh.read()).hexdigest()
    if not force and os.path.exists(ARCHIVE) and os.path.exists(MANIFEST):
        try:
            if json.load(open(MANIFEST)).get("sha256") == digests:
                return ARCHIVE
        except Exception:
            pass
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = ARCHIVE + ".tmp"
    with zipfile.ZipFile(tmp, "w", compression=zipfile.ZIP_STORED) as z:
        dirs = set()
        for _, name in members:        # explicit directory members: zipimport finds namespace packages (envs/, utils/,
            parts = name.split("/")[:-1]   # models/ have no __init__.py in the reference) only through them
            for i in range(1, len(parts) + 1):
                dirs.add("/".join(parts[:i]) + "/")
        for d in sorted(dirs):
            z.writestr(zipfile.ZipInfo(d), b"")
        for src, name in members:
            z.write(os.path.join(REFERENCE_ROOT, src), arcname=name)
    os.replace(tmp, ARCHIVE)
    with open(MANIFEST, "w") as fh:
        json.dump({"source": REFERENCE_ROOT, "what": "unmodified reference files (test infrastructure, not product source)",
                   "sha256": digests}, fh, indent=1, sort_keys=True)
    return ARCHIVE


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
