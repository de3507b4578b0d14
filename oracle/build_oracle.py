"""Compile the oracle's C restatement (TEST INFRASTRUCTURE ONLY).

``oracle/collision_oracle.c`` -> ``oracle/_build/libcollision_oracle.so`` with plain gcc.
The reference itself is pure Python and its collision arithmetic lives in third-party FCL,
which is not in the reference tree, so there is nothing to compile into ``oracle/_ref``;
DESIGN.md records that.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libcollision_oracle.so")
SRC = os.path.join(HERE, "collision_oracle.c")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"],
                       check=True)
    return LIB


REF_SRC = "/root/reference/src/optimizations"
REF_DIR = os.path.join(HERE, "_ref")


def populate_ref() -> bool:
    """Place the UNMODIFIED reference trajectory package under the git-ignored ``oracle/_ref/`` (it
    then travels to the GPU box with the snapshot, where /root/reference does not exist).  The
    reference is pure Python: there is nothing to compile.  Only ``bench.py``'s CPU arm imports it
    (``import_ref``).  Returns whether ``oracle/_ref/optimizations`` is available."""
    import shutil
    dst = os.path.join(REF_DIR, "optimizations")
    if os.path.isdir(REF_SRC):
        os.makedirs(REF_DIR, exist_ok=True)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(REF_SRC, dst, ignore=shutil.ignore_patterns("__pycache__"))
    return os.path.isfile(os.path.join(dst, "calculatingTrajectories.py"))


def import_ref():
    """Import the unmodified reference package from ``oracle/_ref`` (None when absent).  matplotlib /
    mpl_toolkits are not in the image and only used by its plotting helper, so empty stand-in
    modules are registered first (as oracle/make_golden.py does); no reference file is touched."""
    import importlib.util
    import sys
    import types
    import warnings
    init = os.path.join(REF_DIR, "optimizations", "__init__.py")
    if not os.path.isfile(init):
        return None
    warnings.filterwarnings("ignore")
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(sys.modules["mpl_toolkits.mplot3d"], "Axes3D"):
        sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("_reference_optimizations", init,
                                                  submodule_search_locations=[os.path.dirname(init)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_reference_optimizations"] = mod
    spec.loader.exec_module(mod)
    import importlib
    ct = importlib.import_module("_reference_optimizations.calculatingTrajectories")
    return mod, ct


_lib = None


def c_collide_poses(robot_tris, env_tris, poses, prune: bool = True) -> np.ndarray:
    """ctypes front end of ``oracle_collide_poses`` (same contract as
    ``collision_oracle.collide_poses`` without the margin)."""
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_collide_poses.restype = ctypes.c_int
        _lib.oracle_collide_poses.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_void_p]
    robot = np.ascontiguousarray(robot_tris, dtype=np.float64).reshape(-1, 9)
    env = np.ascontiguousarray(env_tris, dtype=np.float64).reshape(-1, 9)
    poses = np.ascontiguousarray(poses, dtype=np.float64)
    hit = np.zeros(poses.shape[0], dtype=np.uint8)
    rc = _lib.oracle_collide_poses(robot.ctypes.data, robot.shape[0], env.ctypes.data, env.shape[0],
                                   poses.ctypes.data, poses.shape[0], poses.shape[1], int(prune),
                                   hit.ctypes.data)
    if rc != 0:
        raise ValueError("oracle_collide_poses: bad pose_dim")
    return hit


if __name__ == "__main__":
    print(build(force=True))
