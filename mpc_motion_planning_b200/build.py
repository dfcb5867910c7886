"""Build libmpcb200.so (sm_100a) in-tree with nvcc.  `python -m mpc_motion_planning_b200.build`.

The kernel families are separate translation units (csrc/mpcb_variants.cu compiled once per
-DMPCB_FAMILY=k) so that the compile runs on all host cores; objects go to _obj/ (git-ignored)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libmpcb200.so")
OBJ = os.path.join(HERE, "_obj")
SRCS = sorted(os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))) + [
    os.path.join(HERE, "..", "include", "mpcb200.h")
]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
N_FAMILIES = 12  # csrc/mpcb_variants.cu: 10 kinematic families + the dynamic bicycle (rows aligned, rows as shipped)
N_LANE_FAMILIES = 10  # csrc/mpcb_lane.cu: the lane-per-scenario engine, kinematic families with plain rows


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in SRCS)


SO_DEBUG = os.path.join(HERE, "libmpcb200_debug.so")
DEBUG_FAMILIES = (0, 2, 5)


def build_debug_slots() -> str:
    """libmpcb200_debug.so: the kinematic warp kernels (rate rows + 1-2 obstacles, plain and discrete-CBF rows, and the
    row-free family) with -DMPCB_DEBUG_SLOTS - the slot-ownership checker that stands in for compute-sanitizer."""
    if os.path.exists(SO_DEBUG) and all(os.path.getmtime(s) <= os.path.getmtime(SO_DEBUG) for s in SRCS if os.path.exists(s)):
        return SO_DEBUG
    nvcc = os.environ.get("NVCC", "nvcc")
    flags = NVCC_FLAGS + ["-DMPCB_DEBUG_SLOTS"]
    os.makedirs(OBJ, exist_ok=True)
    # only the families the checker test exercises are recompiled (row-free, rate row + one obstacle, discrete-CBF rows);
    # the other families and the lane engine (no aliased slots) are linked from the regular objects - KParams is the same
    build()
    jobs = [(os.path.join(HERE, "csrc", "mpcb_api.cu"), os.path.join(OBJ, "dbg_api.o"), [])]
    jobs += [(os.path.join(HERE, "csrc", "mpcb_variants.cu"), os.path.join(OBJ, f"dbg_family{k}.o"), [f"-DMPCB_FAMILY={k}"]) for k in DEBUG_FAMILIES]
    regular = [os.path.join(OBJ, f"family{k}.o") for k in range(N_FAMILIES) if k not in DEBUG_FAMILIES]
    regular += [os.path.join(OBJ, f"lane{k}.o") for k in range(N_LANE_FAMILIES)]

    def compile_one(job):
        src, obj, defs = job
        return job, subprocess.run([nvcc] + flags + defs + ["-c", "-o", obj, src], cwd=HERE, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=max(1, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, jobs))
    for (src, obj, defs), r in results:
        if r.returncode != 0:
            sys.stderr.write(f"---- nvcc {' '.join(defs)} {os.path.basename(src)}\n{r.stdout}{r.stderr}")
            raise subprocess.CalledProcessError(r.returncode, r.args)
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO_DEBUG] + [obj for _, obj, _ in jobs] + regular, cwd=HERE)
    return SO_DEBUG


def build(force: bool = False, verbose: bool = False) -> str:
    if not (force or stale()):
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("MPCB_NVCC_EXTRA", "").split()  # e.g. -DMPCB_W0=12 for tuning experiments
    flags = NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else [])
    os.makedirs(OBJ, exist_ok=True)
    jobs = [(os.path.join(HERE, "csrc", "mpcb_api.cu"), os.path.join(OBJ, "api.o"), [])]
    jobs += [(os.path.join(HERE, "csrc", "mpcb_variants.cu"), os.path.join(OBJ, f"family{k}.o"), [f"-DMPCB_FAMILY={k}"]) for k in range(N_FAMILIES)]
    jobs += [(os.path.join(HERE, "csrc", "mpcb_lane.cu"), os.path.join(OBJ, f"lane{k}.o"), [f"-DMPCB_LANE_FAMILY={k}"]) for k in range(N_LANE_FAMILIES)]

    def compile_one(job):
        src, obj, defs = job
        r = subprocess.run([nvcc] + flags + defs + ["-c", "-o", obj, src], cwd=HERE, capture_output=True, text=True)
        return job, r

    with ThreadPoolExecutor(max_workers=max(1, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, jobs))
    for (src, obj, defs), r in results:
        if verbose or r.returncode != 0:
            sys.stderr.write(f"---- nvcc {' '.join(defs)} {os.path.basename(src)}\n{r.stdout}{r.stderr}")
        if r.returncode != 0:
            raise subprocess.CalledProcessError(r.returncode, r.args)
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO] + [obj for _, obj, _ in jobs], cwd=HERE)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--debug-slots" in sys.argv:
        print(build_debug_slots())
