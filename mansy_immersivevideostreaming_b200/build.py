"""In-tree build of ``csrc/libmansy_b200.so`` for sm_100a (nvcc cross-compiles without a GPU).

    python -m mansy_immersivevideostreaming_b200.build [--force] [--verbose]

Every ``.cu`` is compiled to its own object (in parallel, rebuilt only when it or a header changed) and the
objects are linked into the shared library the ctypes layer loads.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from typing import List

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
LIB = os.path.join(CSRC, "libmansy_b200.so")
OBJ_DIR = os.path.join(CSRC, "build")
SOURCES = ["mansy_sim.cu", "mansy_policy.cu", "mansy_policy_tc.cu", "mansy_mtio.cu", "mansy_peer.cu"]
HEADERS = ["mansy_core.cuh", "mansy_sim.cuh", "mansy_step.cuh", "mansy_policy.cuh", "mansy_tc.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only; no PTX for other targets
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-shared", "-cudart", "static"]         # self-contained: no dependency on torch's libcudart


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _deps() -> List[str]:
    return [os.path.join(CSRC, f) for f in HEADERS] + [os.path.join(INCLUDE, "mansy_b200.h")]


def _flags_tag(extra: List[str]) -> str:
    return hashlib.sha1(" ".join(NVCC_FLAGS + extra).encode()).hexdigest()[:10]


def build_library(force: bool = False, verbose: bool = False) -> str:
    extra = os.environ.get("MANSY_NVCC_EXTRA", "").split()          # tuning experiments, e.g. -DMANSY_STEP_MIN_BLOCKS=6
    tag = _flags_tag(extra)
    srcs = [os.path.join(CSRC, f) for f in SOURCES] + _deps()
    tag_file = os.path.join(OBJ_DIR, "linked.tag")      # which flag set the library on disk was linked from
    linked = open(tag_file).read().strip() if os.path.exists(tag_file) else ""
    if (not force and not extra and linked in ("", tag) and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in srcs)):
        return LIB            # e.g. on the GPU box: the library travels with the snapshot, the objects do not
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_m = max(os.path.getmtime(d) for d in _deps())
    nvcc = None
    jobs = []
    objs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, f"{os.path.splitext(src)[0]}.{tag}.o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(sp), hdr_m):
            nvcc = nvcc or find_nvcc()
            cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", INCLUDE, "-I", CSRC, "-c", sp, "-o", obj]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            jobs.append(cmd)
    if not jobs and linked == tag and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(o) for o in objs):
        return LIB

    def run(cmd):
        return cmd, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=max(1, len(jobs))) as ex:
        for cmd, proc in ex.map(run, jobs):
            if proc.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
            if verbose:
                sys.stderr.write(proc.stdout + proc.stderr)
    nvcc = nvcc or find_nvcc()
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", *LINK_FLAGS, "-Xcompiler", "-fPIC", "-o", LIB, *objs]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    with open(tag_file, "w") as fh:
        fh.write(tag)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
