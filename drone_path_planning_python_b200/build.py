"""Build libmst.so in-tree with nvcc for sm_100a (no JIT cache: the .so must travel with
the repository snapshot to the GPU box)."""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libmst.so")
SOURCES = ["api.cu", "solve_banded_lu.cu", "solve_condensed.cu", "sample.cu", "collide.cu",
           "formation.cu", "pipeline_fused.cu", "pipeline_onepass.cu", "pipeline_cull.cu", "csv.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--extended-lambda", "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    lib_m = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(PKG), "include", "mst.h"))
    return any(os.path.getmtime(d) > lib_m for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into libmst.so; returns the library path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + srcs
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(PKG, "build_ptxas.log")
    with open(log, "w") as fh:
        # registers / spills / shared memory per kernel (-Xptxas -v); compile times dropped so that
        # the tracked log only changes when the code does
        text = "\n".join(l for l in (proc.stdout + proc.stderr).splitlines() if "Compile time" not in l)
        fh.write(" ".join(cmd) + "\n" + text + "\n")
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (see %s)" % log)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose="-v" in sys.argv))
