"""Build libqtrack.so (the C-ABI library of include/qtrack.h) for sm_100a with nvcc.

The library is built in-tree (lqr_quadcopter_test_b200/libqtrack.so) so that it travels with
the repository snapshot to the GPU box; nvcc cross-compiles without a GPU.  Objects are
rebuilt only when a source or header is newer.
"""
from __future__ import annotations

import os
import pathlib
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = pathlib.Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = PKG / "libqtrack.so"
SOURCES = ["qt_api.cu", "qt_env.cu", "qt_rollout.cu", "qt_dare.cu", "qt_reduce.cu", "qt_policy.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libqtrack.so cannot be built (there is no CPU fallback)")
    return exe


def _newest_header() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(INCLUDE.glob("*.h"))
    return max(h.stat().st_mtime for h in hdrs)


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None) -> pathlib.Path:
    nvcc = _nvcc()
    obj_dir = PKG / "build"
    obj_dir.mkdir(exist_ok=True)
    hdr_time = _newest_header()
    flags = NVCC_FLAGS + (extra_flags or [])
    jobs = []
    for src in SOURCES:
        s = CSRC / src
        o = obj_dir / (src + ".o")
        stale = force or not o.exists() or o.stat().st_mtime < max(s.stat().st_mtime, hdr_time)
        if stale:
            jobs.append([nvcc, *flags, "-c", str(s), "-o", str(o)])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(run, jobs))
    objs = [str(obj_dir / (s + ".o")) for s in SOURCES]
    if jobs or not LIB.exists():
        run([nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
