"""In-tree build of ``csrc/libmansy_b200.so`` for sm_100a (nvcc cross-compiles without a GPU).

    python -m mansy_immersivevideostreaming_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from typing import List

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
LIB = os.path.join(CSRC, "libmansy_b200.so")
SOURCES = ["mansy_sim.cu", "mansy_policy.cu", "mansy_policy_tc.cu", "mansy_mtio.cu"]
HEADERS = ["mansy_core.cuh", "mansy_sim.cuh", "mansy_step.cuh", "mansy_policy.cuh", "mansy_tc.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only; no PTX for other targets
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",                              # self-contained: no dependency on torch's libcudart
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    lib_m = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(INCLUDE, "mansy_b200.h")]
    return any(os.path.getmtime(d) > lib_m for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    cmd: List[str] = [find_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-I", CSRC]
    cmd += os.environ.get("MANSY_NVCC_EXTRA", "").split()          # tuning experiments, e.g. -DMANSY_STEP_MIN_BLOCKS=6
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stdout + proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
