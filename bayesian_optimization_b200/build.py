"""Build libb200bo.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build() and by hand:
``python -m bayesian_optimization_b200.build``.  The .so is git-ignored but travels with gpurun."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200bo.so")
HOSTMATH = os.path.join(HERE, "libb200bo_hostmath.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def sources():
    out = [os.path.join(HERE, "..", "include", "b200bo.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h", ".cpp")):
            out.append(os.path.join(CSRC, f))
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = sources()
    if force or not _newer(LIB, srcs):
        dev = ["-DB200BO_DEV_KERNELS"] if os.environ.get("B200BO_DEV_KERNELS") == "1" else []  # + the superseded generations 2, 3
        dev += ["-D" + d for d in os.environ.get("B200BO_EXTRA_DEFINES", "").split() if d]             # A/B builds
        cmd = [nvcc, *NVCC_FLAGS, *dev, "-o", LIB, os.path.join(CSRC, "b200bo.cu")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(os.path.join(HERE, "build.log"), "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + log[-6000:])
        if verbose:
            print(log)
    if force or not _newer(HOSTMATH, srcs):
        gxx = shutil.which("g++") or "g++"
        cmd = [gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", HOSTMATH, os.path.join(CSRC, "hostmath.cpp")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
