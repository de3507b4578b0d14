"""Exact (rational arithmetic) geometry decisions used to ANCHOR the collision oracle.

TEST INFRASTRUCTURE ONLY — see the header of ``oracle/minsnap_oracle.py`` for who may import
``oracle/``.

The reference's collision answers come from FCL, which is absent here (``collision_oracle.py``
header: parity unpinned).  What CAN be pinned is the predicate the restatement claims to
implement — "robot mesh at a pose and environment mesh share a point, closed triangles, touching
counts" (what fcl.collide with a default request decides at src/RigidBodyPlanners/
fcl_checker.py:93-100) — by deciding it in exact arithmetic with a method that shares nothing
with a separating-axis test: two triangles meet iff an edge of one meets the other (closed
segment against closed triangle; coplanar configurations handled in 2-D).
"""
from __future__ import annotations

from fractions import Fraction as F


def _sub(a, b):
    return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def _orient2(a, b, c):
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])


def _point_in_tri2(p, t):
    s = [_orient2(t[i], t[(i + 1) % 3], p) for i in range(3)]
    return all(x >= 0 for x in s) or all(x <= 0 for x in s)


def _seg_seg2(p, q, a, b):
    d1, d2 = _orient2(a, b, p), _orient2(a, b, q)
    d3, d4 = _orient2(p, q, a), _orient2(p, q, b)
    if d1 == 0 and d2 == 0 and d3 == 0 and d4 == 0:      # collinear: overlap of the ranges
        k = 0 if p[0] != q[0] or a[0] != b[0] else 1
        return max(min(p[k], q[k]), min(a[k], b[k])) <= min(max(p[k], q[k]), max(a[k], b[k]))
    return (d1 * d2 <= 0) and (d3 * d4 <= 0)


def _seg_tri(p, q, t):
    n = _cross(_sub(t[1], t[0]), _sub(t[2], t[0]))
    dp, dq = _dot(n, _sub(p, t[0])), _dot(n, _sub(q, t[0]))
    if (dp > 0 and dq > 0) or (dp < 0 and dq < 0):
        return False
    if dp == 0 and dq == 0:                               # in the plane: drop the dominant axis
        k = max(range(3), key=lambda i: abs(n[i]))
        keep = [i for i in range(3) if i != k]
        P, Q = [p[i] for i in keep], [q[i] for i in keep]
        T = [[v[i] for i in keep] for v in t]
        return (_point_in_tri2(P, T) or _point_in_tri2(Q, T) or
                any(_seg_seg2(P, Q, T[i], T[(i + 1) % 3]) for i in range(3)))
    s = dp / (dp - dq)
    x = [p[i] + s * (q[i] - p[i]) for i in range(3)]
    side = [_dot(_cross(_sub(t[(i + 1) % 3], t[i]), _sub(x, t[i])), n) for i in range(3)]
    return all(v >= 0 for v in side)


def rational_triangle(tri):
    """Corners as exact rationals (every double is a rational)."""
    return [[F(float(x)) for x in v] for v in tri]


def _hull(tri):
    """A "triangle" with collinear corners is the segment between its two extreme corners (or a
    point): the shipped robot-scene-triangle mesh has four such facets after the 2-decimal
    rounding of fcl_checker.py:20-23.  Returns ("tri", tri) or ("seg", (p, q))."""
    n = _cross(_sub(tri[1], tri[0]), _sub(tri[2], tri[0]))
    if n[0] != 0 or n[1] != 0 or n[2] != 0:
        return "tri", tri
    best, ends = -1, (tri[0], tri[0])
    for i in range(3):
        for j in range(i + 1, 3):
            d = _sub(tri[i], tri[j])
            dd = _dot(d, d)
            if dd > best:
                best, ends = dd, (tri[i], tri[j])
    return "seg", ends


def _seg_seg3(p, q, a, b):
    """Closed segments pq and ab in space (either may be a point)."""
    u, v, w = _sub(q, p), _sub(b, a), _sub(a, p)
    n = _cross(u, v)
    if _dot(n, w) != 0:
        return False                                      # not coplanar
    if n[0] != 0 or n[1] != 0 or n[2] != 0:
        k = max(range(3), key=lambda i: abs(n[i]))
        keep = [i for i in range(3) if i != k]
        return _seg_seg2([p[i] for i in keep], [q[i] for i in keep], [a[i] for i in keep], [b[i] for i in keep])
    # parallel (or a point involved): they meet only if all four points are collinear
    d = u if any(x != 0 for x in u) else v
    if not any(x != 0 for x in d):
        return p == a                                     # two points
    for x in (p, q, a, b):
        c = _cross(d, _sub(x, p if any(y != 0 for y in u) else a))
        if any(y != 0 for y in c):
            return False
    k = max(range(3), key=lambda i: abs(d[i]))
    return max(min(p[k], q[k]), min(a[k], b[k])) <= min(max(p[k], q[k]), max(a[k], b[k]))


def rational_triangles_meet(A, B):
    """A, B: triangles of ``Fraction`` corners, as closed point sets (degenerate ones are
    segments or points)."""
    ka, ha = _hull(A)
    kb, hb = _hull(B)
    if ka == "tri" and kb == "tri":
        return (any(_seg_tri(A[i], A[(i + 1) % 3], B) for i in range(3)) or
                any(_seg_tri(B[i], B[(i + 1) % 3], A) for i in range(3)))
    if ka == "seg" and kb == "tri":
        return _seg_tri(ha[0], ha[1], B)
    if ka == "tri" and kb == "seg":
        return _seg_tri(hb[0], hb[1], A)
    return _seg_seg3(ha[0], ha[1], hb[0], hb[1])


def triangles_meet(A, B):
    """Do two closed triangles (float corners) share a point?  Exact."""
    return rational_triangles_meet(rational_triangle(A), rational_triangle(B))


def _boxes_apart(A, B):
    for k in range(3):
        if max(v[k] for v in A) < min(v[k] for v in B) or max(v[k] for v in B) < min(v[k] for v in A):
            return True
    return False


def robot_meets_env(robot_tris, env_tris, R, T):
    """Exact decision of ``Fcl_checker.check_collision`` (fcl_checker.py:93-100): does the robot
    mesh placed at ``x -> R x + T`` share a point with the environment mesh?  ``R`` (3x3) and ``T``
    (3) are taken as the exact rationals their doubles are; the transform itself is exact."""
    Rq = [[F(float(R[i][j])) for j in range(3)] for i in range(3)]
    Tq = [F(float(T[i])) for i in range(3)]
    env = [rational_triangle(t) for t in env_tris]
    for tri in robot_tris:
        A = []
        for v in tri:
            vq = [F(float(x)) for x in v]
            A.append([Rq[i][0] * vq[0] + Rq[i][1] * vq[1] + Rq[i][2] * vq[2] + Tq[i] for i in range(3)])
        for B in env:
            if _boxes_apart(A, B):
                continue
            if rational_triangles_meet(A, B):
                return True
    return False
