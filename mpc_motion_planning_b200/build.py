"""Build libmpcb200.so (sm_100a) in-tree with nvcc.  `python -m mpc_motion_planning_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libmpcb200.so")
SRCS = sorted(os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))) + [
    os.path.join(HERE, "..", "include", "mpcb200.h")
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in SRCS)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        nvcc = os.environ.get("NVCC", "nvcc")
        extra = os.environ.get("MPCB_NVCC_EXTRA", "").split()  # e.g. -DMPCB_W0=12 for tuning experiments
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, os.path.join(HERE, "csrc", "mpcb_api.cu")]
        subprocess.check_call(cmd, cwd=HERE)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
