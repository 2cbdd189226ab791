"""Builds libcaldera_b200.so (sm_100a only) in-tree with nvcc.

    python -m ee274_convexcaldera_llm_quantization_b200.build [--force]

One translation unit per csrc/*.cu, compiled in parallel to objects under csrc/_build and
linked into ee274_convexcaldera_llm_quantization_b200/libcaldera_b200.so.  No torch headers,
no JIT cache: the .so is a plain C-ABI library (include/caldera_b200.h) and travels with
the tree.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(PKG, "libcaldera_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, measure: bool = False) -> str:
    """measure=True builds libcaldera_b200_measure.so instead: the same sources with -DCB_MEASURE, i.e. with the timing
    stamps, knock-out switch and probe entry points that scripts/probe_*.py use (never loaded by the package itself)."""
    if measure:
        return _build(force, verbose, os.path.join(CSRC, "_build_measure"), os.path.join(PKG, "libcaldera_b200_measure.so"),
                      FLAGS + ["-DCB_MEASURE"])
    return _build(force, verbose, OBJ, LIB, FLAGS)


def _build(force, verbose, OBJ, LIB, FLAGS) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(os.path.dirname(PKG), "include", "*.h"))
    jobs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run([NVCC] + FLAGS + ["-c", s, "-o", o], capture_output=True, text=True)
        return s, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {s}")
            with open(os.path.join(OBJ, os.path.basename(s)[:-3] + ".ptxas.log"), "w") as f:
                # registers / spills / shared memory per kernel; compile times vary from run to run and are dropped
                f.write("".join(line for line in r.stderr.splitlines(keepends=True) if "Compile time" not in line))
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if force or jobs or _stale(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, measure="--measure" in sys.argv))
