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
