// Triangle-triangle and mesh-mesh intersection arithmetic shared by the collision kernels;
// __host__ __device__ so tests/hostcheck can run the very same code on the CPU.
// Semantics and reference call sites: see collide.cu.
#pragma once
#include <math.h>

#include "mst_common.cuh"

namespace mst {

struct V3 { double x, y, z; };
__host__ __device__ __forceinline__ V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__host__ __device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// true when `ax` does NOT separate the triangles (p1 is the origin after translation)
__host__ __device__ __forceinline__ bool overlap_on(V3 ax, V3 p2, V3 p3, V3 q1, V3 q2, V3 q3) {
  const double a2 = dot(ax, p2), a3 = dot(ax, p3);
  const double b1 = dot(ax, q1), b2 = dot(ax, q2), b3 = dot(ax, q3);
  const double mx1 = fmax(fmax(0.0, a2), a3), mn1 = fmin(fmin(0.0, a2), a3);
  const double mx2 = fmax(fmax(b1, b2), b3), mn2 = fmin(fmin(b1, b2), b3);
  return !(mn1 > mx2 || mn2 > mx1);
}

// 17-axis SAT in the axis order of FCL's intersect_Triangle
__host__ __device__ inline bool triangles_intersect(V3 P1, V3 P2, V3 P3, V3 Q1, V3 Q2, V3 Q3) {
  const V3 p2 = sub(P2, P1), p3 = sub(P3, P1);
  const V3 q1 = sub(Q1, P1), q2 = sub(Q2, P1), q3 = sub(Q3, P1);
  const V3 e1 = p2, e2 = sub(p3, p2), e3 = {-p3.x, -p3.y, -p3.z};
  const V3 f1 = sub(q2, q1), f2 = sub(q3, q2), f3 = sub(q1, q3);
  const V3 n1 = cross(e1, e2);
  if (!overlap_on(n1, p2, p3, q1, q2, q3)) return false;
  const V3 m1 = cross(f1, f2);
  if (!overlap_on(m1, p2, p3, q1, q2, q3)) return false;
  if (!overlap_on(cross(e1, f1), p2, p3, q1, q2, q3)) return false;
  if (!overlap_on(cross(e1, f2), p2, p3, q1, q2, q3)) return false;
  if (!overlap_on(cross(e1, f3), p2, p3, q1, q2, q3)) return false;
  if (!overlap_on(cross(e2, f1), p2, p3, q1, q2, q3)) return false;
  if (!overlap_on(cross(e2, f2), p2, p3, q1, q2, q3)) return false;
  if (!overlap_on(cross(e2, f3), p2, p3, q1, q2, q3)) return false;
  if (!overlap_on(cross(e3, f1), p2, p3, q1, q2, q3)) return false;
  if (!overlap_on(cross(e3, f2), p2, p3, q1, q2, q3)) return false;
  if (!overlap_on(cross(e3, f3), p2, p3, q1, q2, q3)) return false;
  if (!overlap_on(cross(e1, n1), p2, p3, q1, q2, q3)) return false;
  if (!overlap_on(cross(e2, n1), p2, p3, q1, q2, q3)) return false;
  if (!overlap_on(cross(e3, n1), p2, p3, q1, q2, q3)) return false;
  if (!overlap_on(cross(f1, m1), p2, p3, q1, q2, q3)) return false;
  if (!overlap_on(cross(f2, m1), p2, p3, q1, q2, q3)) return false;
  if (!overlap_on(cross(f3, m1), p2, p3, q1, q2, q3)) return false;
  return true;
}

__host__ __device__ __forceinline__ V3 xform(const double* R, const double* T, const double* v) {
  return {R[0] * v[0] + R[1] * v[1] + R[2] * v[2] + T[0],
          R[3] * v[0] + R[4] * v[1] + R[5] * v[2] + T[1],
          R[6] * v[0] + R[7] * v[1] + R[8] * v[2] + T[2]};
}

// robot at (R, T) against the staged environment
__host__ __device__ inline bool robot_hits_env(const double* R, const double* T, const double* __restrict__ rtri, int Tr,
                               const double* __restrict__ etri, const double* __restrict__ ebox, int Te,
                               const double* root, double radius, bool rigid) {
  // bounding sphere of the robot about its origin vs the environment's root box (only
  // when R is a rotation by construction; a caller-supplied quaternion may not be unit)
  if (rigid && (T[0] + radius < root[0] || T[0] - radius > root[3] || T[1] + radius < root[1] ||
      T[1] - radius > root[4] || T[2] + radius < root[2] || T[2] - radius > root[5]))
    return false;
  for (int r = 0; r < Tr; ++r) {
    const V3 P1 = xform(R, T, rtri + 9 * r), P2 = xform(R, T, rtri + 9 * r + 3),
             P3 = xform(R, T, rtri + 9 * r + 6);
    const double lo0 = fmin(fmin(P1.x, P2.x), P3.x), hi0 = fmax(fmax(P1.x, P2.x), P3.x);
    const double lo1 = fmin(fmin(P1.y, P2.y), P3.y), hi1 = fmax(fmax(P1.y, P2.y), P3.y);
    const double lo2 = fmin(fmin(P1.z, P2.z), P3.z), hi2 = fmax(fmax(P1.z, P2.z), P3.z);
    if (hi0 < root[0] || lo0 > root[3] || hi1 < root[1] || lo1 > root[4] || hi2 < root[2] ||
        lo2 > root[5])
      continue;
    for (int e = 0; e < Te; ++e) {
      const double* bx = ebox + 6 * e;
      if (hi0 < bx[0] || lo0 > bx[3] || hi1 < bx[1] || lo1 > bx[4] || hi2 < bx[2] || lo2 > bx[5])
        continue;
      const double* q = etri + 9 * e;
      const V3 Q1 = {q[0], q[1], q[2]}, Q2 = {q[3], q[4], q[5]}, Q3 = {q[6], q[7], q[8]};
      if (triangles_intersect(P1, P2, P3, Q1, Q2, Q3)) return true;
    }
  }
  return false;
}

}  // namespace mst
