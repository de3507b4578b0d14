// Triangle-triangle and mesh-mesh intersection arithmetic shared by the collision kernels;
// __host__ __device__ so tests/hostcheck can run the very same code on the CPU.
// Semantics and reference call sites: see collide.cu.
#pragma once
#include <math.h>

#include "mst_common.cuh"

// hides a value's origin from the optimiser (both compilers take the empty asm)
#define MST_OPAQUE(x) asm volatile("" : "+r"(x))

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

// 17-axis SAT in the axis order of FCL's intersect_Triangle.  Kept out of line: the kernels reach it
// only for parallel / coplanar pairs, and inlined at every call site it tripled their code size
// (instruction-cache misses showed as 0.5 "no instruction" stall cycles per issue).
__host__ __device__ __noinline__ inline bool triangles_intersect(V3 P1, V3 P2, V3 P3, V3 Q1, V3 Q2, V3 Q3) {
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

// Interval form of the triangle-triangle test (Moller, "A fast triangle-triangle intersection
// test", 1997, division-free variant): both triangles must straddle (or touch) the other's
// plane; they then meet the line where the two planes cross in one interval each, and
// intersect iff the two intervals overlap.  In exact arithmetic this is the same predicate
// as the 17-axis SAT for non-coplanar triangles — closed triangles, touching counts — at a
// fraction of the operations (the nine edge-edge axes collapse into one interval
// comparison).  Parallel / coplanar / degenerate pairs (plane normals not independent) go
// to the 17-axis test, which is what decides them in FCL too.  The two predicates can only
// disagree inside the rounding band around touching configurations.
// Written with selects instead of one assignment block per case: the compiler turned the
// case blocks (and the dominant-axis choice below) into separate code paths, which a warp
// whose lanes hold different triangle pairs then walked one after the other with a third
// of its lanes each (profiles/r1_collision_history.md, generation r1g).
__host__ __device__ __forceinline__ bool interval_terms(double v0, double v1, double v2, double d0, double d1,
                                                        double d2, double& a, double& b, double& c, double& x0,
                                                        double& x1) {
  // the vertex that is alone on its side of the other plane (or the one off the plane)
  const bool s01 = d0 * d1 > 0.0, s02 = d0 * d2 > 0.0, s12 = d1 * d2 > 0.0;
  int lone = s01 ? 2 : (s02 ? 1 : ((s12 || d0 != 0.0) ? 0 : (d1 != 0.0 ? 1 : (d2 != 0.0 ? 2 : -1))));
  MST_OPAQUE(lone);               // or the cases are threaded into separate copies of what follows
  if (lone < 0) return false;     // the triangle lies in the other plane
  const double va = lone == 0 ? v0 : (lone == 1 ? v1 : v2), da = lone == 0 ? d0 : (lone == 1 ? d1 : d2);
  const double vb = lone == 0 ? v1 : v0, db = lone == 0 ? d1 : d0;
  const double vc = lone == 2 ? v1 : v2, dc = lone == 2 ? d1 : d2;
  a = va; b = (vb - va) * da; c = (vc - va) * da; x0 = da - db; x1 = da - dc;
  return true;
}

__host__ __device__ inline bool triangles_intersect_interval(V3 P1, V3 P2, V3 P3, V3 Q1, V3 Q2, V3 Q3) {
  // same translation as the SAT: everything relative to P1
  const V3 p2 = sub(P2, P1), p3 = sub(P3, P1);
  const V3 q1 = sub(Q1, P1), q2 = sub(Q2, P1), q3 = sub(Q3, P1);
  const V3 n2 = cross(sub(q2, q1), sub(q3, q2));
  const double c2 = dot(n2, q1);
  const double dp1 = -c2, dp2 = dot(n2, p2) - c2, dp3 = dot(n2, p3) - c2;
  if ((dp1 > 0.0 && dp2 > 0.0 && dp3 > 0.0) || (dp1 < 0.0 && dp2 < 0.0 && dp3 < 0.0)) return false;
  const V3 n1 = cross(p2, sub(p3, p2));
  const double dq1 = dot(n1, q1), dq2 = dot(n1, q2), dq3 = dot(n1, q3);
  if ((dq1 > 0.0 && dq2 > 0.0 && dq3 > 0.0) || (dq1 < 0.0 && dq2 < 0.0 && dq3 < 0.0)) return false;
  const V3 D = cross(n1, n2);
  const double ax = fabs(D.x), ay = fabs(D.y), az = fabs(D.z);
  const double dmax = fmax(ax, fmax(ay, az));
  // normals (numerically) dependent: parallel planes, coplanar or degenerate triangles
  const double scale = (fabs(n1.x) + fabs(n1.y) + fabs(n1.z)) * (fabs(n2.x) + fabs(n2.y) + fabs(n2.z));
  if (!(dmax > 1e-12 * scale)) return triangles_intersect(P1, P2, P3, Q1, Q2, Q3);
  // coordinate along the dominant axis of the line (selects, see interval_terms)
  int axis = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
  MST_OPAQUE(axis);
  const bool ux = axis == 0, uy = axis == 1;
  const double pv1 = 0.0;
  const double pv2 = ux ? p2.x : (uy ? p2.y : p2.z), pv3 = ux ? p3.x : (uy ? p3.y : p3.z);
  const double qv1 = ux ? q1.x : (uy ? q1.y : q1.z), qv2 = ux ? q2.x : (uy ? q2.y : q2.z);
  const double qv3 = ux ? q3.x : (uy ? q3.y : q3.z);
  double a, b, c, x0, x1, d, e, f, y0, y1;
  if (!interval_terms(pv1, pv2, pv3, dp1, dp2, dp3, a, b, c, x0, x1) ||
      !interval_terms(qv1, qv2, qv3, dq1, dq2, dq3, d, e, f, y0, y1))
    return triangles_intersect(P1, P2, P3, Q1, Q2, Q3);
  const double xx = x0 * x1, yy = y0 * y1, xxyy = xx * yy;
  double t = a * xxyy;
  const double i10 = t + b * x1 * yy, i11 = t + c * x0 * yy;
  t = d * xxyy;
  const double i20 = t + e * xx * y1, i21 = t + f * xx * y0;
  const double lo1 = fmin(i10, i11), hi1 = fmax(i10, i11), lo2 = fmin(i20, i21), hi2 = fmax(i20, i21);
  return !(hi1 < lo2 || hi2 < lo1);
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

// ---------------------------------------------------------------------------------------
// double -> float rounded towards -inf / +inf (conservative boxes)
__host__ __device__ __forceinline__ float round_down_f(double x) {
#ifdef __CUDA_ARCH__
  return __double2float_rd(x);
#else
  float f = (float)x;
  return (double)f > x ? nextafterf(f, -INFINITY) : f;
#endif
}
__host__ __device__ __forceinline__ float round_up_f(double x) {
#ifdef __CUDA_ARCH__
  return __double2float_ru(x);
#else
  float f = (float)x;
  return (double)f < x ? nextafterf(f, INFINITY) : f;
#endif
}

// Culled mesh-mesh test used by the kernels.  Same answer as robot_hits_env (brute-force
// SAT over all pairs) away from the touching boundary; every cull is a valid separating
// axis in exact arithmetic:
//   1. robot bounding sphere / box against the environment's root box;
//   2. per environment triangle: robot box against the triangle's box (single precision, every
//      bound rounded outward — conservative, the exact tests follow);
//   3. per environment triangle: the triangle's PLANE and its three EDGE PLANES against ALL
//      unique robot vertices at once (the planes are moved into the robot frame: 12 FMAs each,
//      then 3 FMAs per vertex) — SAT axis 2 (the env normal) and the in-plane edge normals,
//      shared by every robot triangle, recorded as "strictly above" / "strictly below" /
//      "beyond edge k" bit masks over the unique vertices;
//   4. per robot triangle: skipped when its three corner bits are all above, all below, or
//      all beyond the same edge;
//   5. the 17-axis SAT on the few surviving pairs.
// This is the host-checkable statement of the culls the warp engine below applies (the engine
// keeps them as sets of robot triangles and uses the interval pair test).
// ROT = false: translation only (R = identity), the K = 3 pipeline.
// e_first / e_step: the environment triangles e_first, e_first + e_step, ... only (the single-query
// kernel spreads them over the lanes of a warp); 0 / 1 = all of them.
template <bool ROT>
__host__ __device__ inline bool robot_hits_env_culled(const double* R, const double* T, const MeshView& rb,
                                                      const MeshBounds& rbb, const MeshView& ev,
                                                      const MeshBounds& evb, bool rigid, int e_first = 0,
                                                      int e_step = 1) {
  const double* root = evb.root;
  if (rigid && (T[0] + rbb.radius < root[0] || T[0] - rbb.radius > root[3] || T[1] + rbb.radius < root[1] ||
                T[1] - rbb.radius > root[4] || T[2] + rbb.radius < root[2] || T[2] - rbb.radius > root[5]))
    return false;
  // world box of the robot: exact for a translation, else the box of the rotated local box
  double lo0, lo1, lo2, hi0, hi1, hi2;
  if (!ROT) {
    lo0 = rbb.root[0] + T[0]; lo1 = rbb.root[1] + T[1]; lo2 = rbb.root[2] + T[2];
    hi0 = rbb.root[3] + T[0]; hi1 = rbb.root[4] + T[1]; hi2 = rbb.root[5] + T[2];
  } else {
    const double c0 = 0.5 * (rbb.root[0] + rbb.root[3]), c1 = 0.5 * (rbb.root[1] + rbb.root[4]),
                 c2 = 0.5 * (rbb.root[2] + rbb.root[5]);
    // half extents, padded so rounding in the products below cannot shrink the box
    const double h0 = 0.5 * (rbb.root[3] - rbb.root[0]), h1 = 0.5 * (rbb.root[4] - rbb.root[1]),
                 h2 = 0.5 * (rbb.root[5] - rbb.root[2]);
    const double pad = 1e-12 * (rbb.radius + fabs(T[0]) + fabs(T[1]) + fabs(T[2]));
    const double w0 = R[0] * c0 + R[1] * c1 + R[2] * c2 + T[0], w1 = R[3] * c0 + R[4] * c1 + R[5] * c2 + T[1],
                 w2 = R[6] * c0 + R[7] * c1 + R[8] * c2 + T[2];
    const double e0 = fabs(R[0]) * h0 + fabs(R[1]) * h1 + fabs(R[2]) * h2 + pad,
                 e1 = fabs(R[3]) * h0 + fabs(R[4]) * h1 + fabs(R[5]) * h2 + pad,
                 e2 = fabs(R[6]) * h0 + fabs(R[7]) * h1 + fabs(R[8]) * h2 + pad;
    lo0 = w0 - e0; hi0 = w0 + e0; lo1 = w1 - e1; hi1 = w1 + e1; lo2 = w2 - e2; hi2 = w2 + e2;
  }
  if (hi0 < root[0] || lo0 > root[3] || hi1 < root[1] || lo1 > root[4] || hi2 < root[2] || lo2 > root[5])
    return false;
  const float flo0 = round_down_f(lo0), flo1 = round_down_f(lo1), flo2 = round_down_f(lo2);
  const float fhi0 = round_up_f(hi0), fhi1 = round_up_f(hi1), fhi2 = round_up_f(hi2);
  for (int e = e_first; e < ev.T; e += e_step) {
    const double* bx = ev.box + 6 * e;
    const float* fb = ev.fbox + 8 * e;
    if (fhi0 < fb[0] || flo0 > fb[3] || fhi1 < fb[1] || flo1 > fb[4] || fhi2 < fb[2] || flo2 > fb[5]) continue;
    // planes of the env triangle in the robot frame: n.(R v + T) - d = (R^T n).v + (n.T - d)
    const double* pl = ev.plane + 4 * e;
    const double* ed = ev.edge + 12 * e;
    double m[4][3], off[4];
    for (int j = 0; j < 4; ++j) {
      const double* q = j == 0 ? pl : ed + 4 * (j - 1);
      if (ROT) {
        m[j][0] = R[0] * q[0] + R[3] * q[1] + R[6] * q[2];
        m[j][1] = R[1] * q[0] + R[4] * q[1] + R[7] * q[2];
        m[j][2] = R[2] * q[0] + R[5] * q[1] + R[8] * q[2];
      } else {
        m[j][0] = q[0]; m[j][1] = q[1]; m[j][2] = q[2];
      }
      off[j] = q[0] * T[0] + q[1] * T[1] + q[2] * T[2] - q[3];
    }
    unsigned long long above = 0ull, below = 0ull, out0 = 0ull, out1 = 0ull, out2 = 0ull;
    for (int v = 0; v < rb.V; ++v) {
      const double* p = rb.vert + 3 * v;
      const double dist = m[0][0] * p[0] + m[0][1] * p[1] + m[0][2] * p[2] + off[0];
      above |= (unsigned long long)(dist > 0.0) << v;
      below |= (unsigned long long)(dist < 0.0) << v;
      out0 |= (unsigned long long)(m[1][0] * p[0] + m[1][1] * p[1] + m[1][2] * p[2] + off[1] > 0.0) << v;
      out1 |= (unsigned long long)(m[2][0] * p[0] + m[2][1] * p[1] + m[2][2] * p[2] + off[2] > 0.0) << v;
      out2 |= (unsigned long long)(m[3][0] * p[0] + m[3][1] * p[1] + m[3][2] * p[2] + off[3] > 0.0) << v;
    }
    const unsigned long long all = rb.V >= 64 ? ~0ull : ((1ull << rb.V) - 1ull);
    if (above == all || below == all || out0 == all || out1 == all || out2 == all) continue;
    const double* q = ev.tri + 9 * e;
    const V3 Q1 = {q[0], q[1], q[2]}, Q2 = {q[3], q[4], q[5]}, Q3 = {q[6], q[7], q[8]};
    for (int r = 0; r < rb.T; ++r) {
      const unsigned long long mk = rb.mask[r];
      if ((above & mk) == mk || (below & mk) == mk || (out0 & mk) == mk || (out1 & mk) == mk || (out2 & mk) == mk)
        continue;
      const double* pr = rb.tri + 9 * r;
      V3 P1, P2, P3;
      if (ROT) {
        P1 = xform(R, T, pr); P2 = xform(R, T, pr + 3); P3 = xform(R, T, pr + 6);
      } else {
        P1 = {pr[0] + T[0], pr[1] + T[1], pr[2] + T[2]};
        P2 = {pr[3] + T[0], pr[4] + T[1], pr[5] + T[2]};
        P3 = {pr[6] + T[0], pr[7] + T[1], pr[8] + T[2]};
      }
      if (fmax(fmax(P1.x, P2.x), P3.x) < bx[0] || fmin(fmin(P1.x, P2.x), P3.x) > bx[3] ||
          fmax(fmax(P1.y, P2.y), P3.y) < bx[1] || fmin(fmin(P1.y, P2.y), P3.y) > bx[4] ||
          fmax(fmax(P1.z, P2.z), P3.z) < bx[2] || fmin(fmin(P1.z, P2.z), P3.z) > bx[5])
        continue;
      if (triangles_intersect(P1, P2, P3, Q1, Q2, Q3)) return true;
    }
  }
  return false;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------
// Warp-level collision engine used by the kernels: a ring of poses waiting for the test,
// drained 32 at a time, every lane walking ITS pose's candidate triangle pairs lazily and
// stopping at the first hit.
//
// How it got here (profiles/r1_collision_history.md): a per-lane loop over all pairs ran the
// pair test with 4.5/32 lanes active (neighbouring poses survive the culls with different
// pairs); compacting (pose, e, r) items across the warp fixed the lane utilisation but still
// tested every candidate pair of every pose — and on the benchmark's trajectories 99 % of the
// poses whose box reaches the obstacle's box do collide, with the first intersecting pair
// among the first 2 candidates for two thirds of them (23 candidates on average).  So the
// work to avoid is everything after the first hit:
//   * a pose enters the ring only if it passes the root-box culls (pose_near_environment);
//   * draining 32 poses, each lane advances a CURSOR (env triangle e, bit mask of the robot
//     triangles still to test against e, bit mask of the env triangles of e's block of 32
//     whose box meets the robot's) to its next candidate — the culls are exact-safe ones:
//     env-triangle boxes (one single-precision pass per block, bounds rounded outward), then
//     the plane of e and its three edge planes against all unique robot vertices, kept as
//     sets of robot triangles (a triangle is a candidate if it straddles the plane and is
//     not wholly beyond an edge) — and all lanes that have a candidate run the pair test
//     together (interval form);
//   * after ROUNDS candidates a pose that is still undecided goes back to the tail of the
//     ring WITH its cursor, so the next batch is dense again instead of 32 lanes waiting for
//     the slowest one.
// A pose is reported (hit or free) exactly once.  The answer equals the brute-force test over
// all pairs: stopping early only skips pairs after a hit.
//
// POSE: 0 translation (x,y,z); 1 yaw (x,y,z,sin(yaw/2),cos(yaw/2)); 2 quaternion
// (x,y,z,qx,qy,qz,qw).  nv (POSE 0 only; may be null): tables n_e . v and m_ek . v of the env planes
// and edge planes against the robot's unique vertices (build_plane_vertex_table) — a translation
// leaves them constant, so the signed distance of vertex v to a plane is one add.  Without the
// table (environments too large for shared memory) the general form runs with R = identity.
constexpr int COLLIDE_MAX_V = 32;    // unique robot vertices the bit masks can hold
constexpr int COLLIDE_MAX_TR = 32;   // robot triangles the cursor's bit mask can hold
constexpr int COLLIDE_RING = 64;     // poses per warp ring (a power of two, >= 2 * 32)
constexpr int COLLIDE_ROUNDS = 2;    // candidates tested per pose and drain (at most)
constexpr int COLLIDE_MIN_LANES = 12; // undecided poses a drain needs to run another round

template <int POSE> struct PoseDim { static constexpr int N = POSE == 0 ? 3 : (POSE == 1 ? 5 : 7); };

// yaw poses: R = Rz(yaw) = [[c, -s, 0], [s, c, 0], [0, 0, 1]] with c = 1 - 2 z^2, s = 2 z w of the half-angle pair
// (z, w) = (sin, cos)(yaw / 2) — the entries quat_to_matrix(0, 0, z, w) computes; the specialised forms below
// leave out the products with the structural zeros and ones (same values, fewer FP64 instructions).
struct Yaw { double c, s; };
__device__ __forceinline__ Yaw yaw_of(const double* pp) { return {1.0 - 2.0 * (pp[3] * pp[3]), 2.0 * (pp[3] * pp[4])}; }
// R v + T
__device__ __forceinline__ V3 xform_yaw(const Yaw& y, const double* T, const double* v) {
  return {y.c * v[0] - y.s * v[1] + T[0], y.s * v[0] + y.c * v[1] + T[1], v[2] + T[2]};
}

template <int POSE>
__device__ __forceinline__ void pose_rotation(const double* pp, double* R) {
  if (POSE == 1) {
    const Yaw y = yaw_of(pp);
    R[0] = y.c; R[1] = -y.s; R[2] = 0.0; R[3] = y.s; R[4] = y.c; R[5] = 0.0; R[6] = 0.0; R[7] = 0.0; R[8] = 1.0;
  }
  if (POSE == 2) quat_to_matrix(pp[3], pp[4], pp[5], pp[6], R);
}

__host__ __device__ __forceinline__ bool collide_engine_supports(const MeshView& rb, const MeshView& ev) {
  return rb.V <= COLLIDE_MAX_V && rb.T <= COLLIDE_MAX_TR && ev.T < (1 << 20);
}

// Translation-only poses: tables of the env planes against the robot's unique vertices, which a
// translation leaves constant (the signed distance of a moved vertex is then one add).  Four
// rows of V per env triangle: nv[e TS + 0 V + v] = n_e . vertex_v (the triangle's plane),
// nv[e TS + (1 + k) V + v] = m_ek . vertex_v (edge plane k).  TS = 4 V made odd: the lanes of a warp walk
// DIFFERENT triangles e at the same v, and with TS a multiple of 16 doubles (V = 8: the benchmark's robot) they
// all met on one bank pair — 1.8 wavefronts per load on the engine's most frequent shared-memory access.  Every
// thread of the CTA calls the builder, followed by a CTA barrier.
constexpr int COLLIDE_TABLE_ROWS = 4;
__host__ __device__ __forceinline__ int collide_table_stride(int robotV) { return (COLLIDE_TABLE_ROWS * robotV) | 1; }
__host__ __device__ __forceinline__ size_t collide_table_doubles(int envT, int robotV) {
  return ((size_t)envT * collide_table_stride(robotV) + 1) & ~(size_t)1;   // what follows stays 16-byte aligned
}
__device__ __forceinline__ void build_plane_vertex_table(const MeshView& rb, const MeshView& ev, double* nv) {
  const int TS = collide_table_stride(rb.V);
  for (int i = threadIdx.x; i < COLLIDE_TABLE_ROWS * ev.T * rb.V; i += blockDim.x) {
    const int row = i / rb.V, v = i - row * rb.V;
    const int e = row >> 2, j = row & 3;
    const double* pl = j == 0 ? ev.plane + 4 * e : ev.edge + 12 * e + 4 * (j - 1);
    const double* p = rb.vert + 3 * v;
    nv[e * TS + j * rb.V + v] = pl[0] * p[0] + pl[1] * p[1] + pl[2] * p[2];
  }
}

// world box of the robot at a pose: exact for a translation, else the box of the rotated local box
template <int POSE>
__device__ __forceinline__ void robot_world_box(const double* pp, const double* R, const MeshBounds& rbb,
                                                double* lo, double* hi) {
  const double* T = pp;
  if (POSE == 0) {
    lo[0] = rbb.root[0] + T[0]; lo[1] = rbb.root[1] + T[1]; lo[2] = rbb.root[2] + T[2];
    hi[0] = rbb.root[3] + T[0]; hi[1] = rbb.root[4] + T[1]; hi[2] = rbb.root[5] + T[2];
  } else {
    const double c0 = 0.5 * (rbb.root[0] + rbb.root[3]), c1 = 0.5 * (rbb.root[1] + rbb.root[4]),
                 c2 = 0.5 * (rbb.root[2] + rbb.root[5]);
    // half extents, padded so rounding in the products below cannot shrink the box
    const double h0 = 0.5 * (rbb.root[3] - rbb.root[0]), h1 = 0.5 * (rbb.root[4] - rbb.root[1]),
                 h2 = 0.5 * (rbb.root[5] - rbb.root[2]);
    const double pad = 1e-12 * (rbb.radius + fabs(T[0]) + fabs(T[1]) + fabs(T[2]));
    if (POSE == 1) {   // rotation about z: R = [[c, -s, 0], [s, c, 0], [0, 0, 1]] (c = R[0], s = R[3])
      const double c = R[0], sn = R[3];
      const double w0 = c * c0 - sn * c1 + T[0], w1 = sn * c0 + c * c1 + T[1], w2 = c2 + T[2];
      const double e0 = fabs(c) * h0 + fabs(sn) * h1 + pad, e1 = fabs(sn) * h0 + fabs(c) * h1 + pad, e2 = h2 + pad;
      lo[0] = w0 - e0; hi[0] = w0 + e0; lo[1] = w1 - e1; hi[1] = w1 + e1; lo[2] = w2 - e2; hi[2] = w2 + e2;
      return;
    }
    for (int a = 0; a < 3; ++a) {
      const double w = R[3 * a] * c0 + R[3 * a + 1] * c1 + R[3 * a + 2] * c2 + T[a];
      const double ext = fabs(R[3 * a]) * h0 + fabs(R[3 * a + 1]) * h1 + fabs(R[3 * a + 2]) * h2 + pad;
      lo[a] = w - ext;
      hi[a] = w + ext;
    }
  }
}

// lane-local, position only: can the robot's bounding sphere about its origin reach the environment's
// root box?  (The sampling kernels ask this before they spend a sincos on the yaw of the sample.)
// For a rotation about z (the yaw poses) the bound is a cylinder: planar radius on x / y, the mesh's own z range.
__device__ __forceinline__ bool sphere_near_environment(const double* T, const MeshBounds& rbb, const MeshBounds& evb) {
  const double* root = evb.root;
  return !(T[0] + rbb.rxy < root[0] || T[0] - rbb.rxy > root[3] || T[1] + rbb.rxy < root[1] ||
           T[1] - rbb.rxy > root[4] || T[2] + rbb.root[5] < root[2] || T[2] + rbb.root[2] > root[5]);
}

// lane-local: can the robot at this pose touch the environment's root box at all?
// (bounding sphere for rigid poses, then the robot's world box)
template <int POSE>
__device__ __forceinline__ bool pose_near_environment(const double* pp, const MeshBounds& rbb, const MeshBounds& evb) {
  const double* root = evb.root;
  const double* T = pp;
  // bounding sphere first for rotated poses (cheaper than rotating the box); a translated box
  // is exact and just as cheap, and POSE 2 takes whatever quaternion the caller gave (maybe not
  // unit), so neither uses the sphere
  if (POSE == 1 && !sphere_near_environment(T, rbb, evb)) return false;
  double R[9], lo[3], hi[3];
  pose_rotation<POSE>(pp, R);
  robot_world_box<POSE>(pp, R, rbb, lo, hi);
  return !(hi[0] < root[0] || lo[0] > root[3] || hi[1] < root[1] || lo[1] > root[4] || hi[2] < root[2] ||
           lo[2] > root[5]);
}

// per-warp ring of poses waiting for (more of) the collision test

template <int NP>
struct PoseRing {
  double pose[COLLIDE_RING][NP];
  int id0[COLLIDE_RING], id1[COLLIDE_RING];  // caller's identification of the pose
  int e[COLLIDE_RING];                       // cursor: env triangle whose `need` bits are current (-1: none yet)
  unsigned need[COLLIDE_RING];               // cursor: robot triangles still to test against triangle e
  unsigned emask[COLLIDE_RING];              // cursor: triangles after e in e's block of 32 whose box the robot's meets
};

// append the poses of the lanes with `push` set (ballot-compacted); all 32 lanes call it
template <int POSE>
__device__ __forceinline__ void ring_push(PoseRing<PoseDim<POSE>::N>& ring, unsigned& tail, bool push, const double* pp,
                                          int id0, int id1, int e, unsigned need, unsigned emask) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  __syncwarp();
  const unsigned vote = __ballot_sync(FULL, push);
  if (!vote) return;
  if (push) {
    const unsigned slot = (tail + __popc(vote & ((1u << lane) - 1u))) & (COLLIDE_RING - 1);
#pragma unroll
    for (int i = 0; i < PoseDim<POSE>::N; ++i) ring.pose[slot][i] = pp[i];
    ring.id0[slot] = id0; ring.id1[slot] = id1; ring.e[slot] = e; ring.need[slot] = need;
    ring.emask[slot] = emask;
  }
  tail += __popc(vote);
  __syncwarp();
}

// test up to COLLIDE_ROUNDS candidates for each of the first `count` waiting poses; decided
// poses are handed to report(id0, id1, hit), the others return to the tail with their cursor
// TABLE (POSE 0 only): nv holds the plane x vertex table; false = no table (mesh images read in place
// from device memory), the general form then runs with R = identity.
template <int POSE, bool TABLE = true, class Report>
__device__ __forceinline__ void ring_drain(PoseRing<PoseDim<POSE>::N>& ring, unsigned& head, unsigned& tail, int count,
                                           const MeshView& rb, const MeshBounds& rbb, const MeshView& ev,
                                           const double* nv, Report&& report) {
  constexpr int NP = PoseDim<POSE>::N;
  const int lane = threadIdx.x & 31;
  const bool valid = lane < count;
  const unsigned slot = (head + (valid ? lane : 0)) & (COLLIDE_RING - 1);
  double pp[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) pp[i] = ring.pose[slot][i];
  const int id0 = ring.id0[slot], id1 = ring.id1[slot];
  int e = ring.e[slot];
  unsigned need = ring.need[slot], emask = ring.emask[slot];
  __syncwarp();  // entries are in registers: the slots may be reused by the re-queue below
  head += (unsigned)count;

  double R[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0}, lo[3], hi[3];
  pose_rotation<POSE>(pp, R);
  robot_world_box<POSE>(pp, R, rbb, lo, hi);
  const float flo0 = round_down_f(lo[0]), flo1 = round_down_f(lo[1]), flo2 = round_down_f(lo[2]);
  const float fhi0 = round_up_f(hi[0]), fhi1 = round_up_f(hi[1]), fhi2 = round_up_f(hi[2]);
  bool hit = false, exhausted = !valid;
#pragma unroll 1
  for (int round = 0; round < COLLIDE_ROUNDS; ++round) {
    bool have = false;
    int r = 0;
    // later rounds only while enough poses of the batch are still undecided to fill the lanes:
    // most poses are decided by their first candidate, and a round for a handful of lanes
    // costs as much as one for 32 (they wait in the ring for a denser batch instead)
    bool go = true;
    if (round > 0) {
      __syncwarp();
      go = __popc(__ballot_sync(0xffffffffu, !hit && !exhausted)) >= COLLIDE_MIN_LANES;
    }
    if (go && !hit && !exhausted) {
      // Advance the cursor to the next candidate pair.  Keep this loop SINGLE-EXIT: a version
      // with `break` / `continue` compiled (nvcc 12.9, sm_100a) to BSSY.RELIABLE/BREAK control
      // flow that left the warp split at the ballot of the re-queue below — ptxas emits that
      // ballot as a bare VOTE and drops explicit warp barriers as redundant — so the fragments
      // advanced `tail` differently and ring entries were overwritten (poses never reported;
      // caught by tests/test_gpu_collision.py::test_every_pose_is_answered_once).
      while (need == 0u && !exhausted) {
        if (emask == 0u) {
          // next block of 32 env triangles: all their boxes against the robot's in one pass, so
          // that the plane tests below run on lanes that each HAVE a box-passing triangle
          // (walking e one by one, a quarter of the lanes passed the box test at any e)
          const int base = ((e >> 5) + 1) << 5;
          if (base >= ev.T) {
            exhausted = true;
          } else {
            // single precision with every bound rounded outward: a box pair that meets in double
            // precision meets here too (the plane and pair tests that follow are the exact ones)
            const int cnt = min(32, ev.T - base);
            // meshes of more than one block: the block's own box first (the level above the triangle
            // boxes; blocks are compact because the triangles are stored in Morton order)
            bool block_near = true;
            if (ev.T > 32) {
              const float4* bb = reinterpret_cast<const float4*>(ev.bbox) + 2 * (base >> 5);
              const float4 b0 = bb[0], b1 = bb[1];
              block_near = !(fhi0 < b0.x || flo0 > b0.w || fhi1 < b0.y || flo1 > b1.x || fhi2 < b0.z || flo2 > b1.y);
            }
            if (block_near) {
              const float4* fb = reinterpret_cast<const float4*>(ev.fbox) + 2 * base;
#pragma unroll 4
              for (int i = 0; i < cnt; ++i) {
                const float4 b0 = fb[2 * i], b1 = fb[2 * i + 1];  // min x y z, max x | max y z, pad
                emask |= (unsigned)(!(fhi0 < b0.x || flo0 > b0.w || fhi1 < b0.y || flo1 > b1.x || fhi2 < b0.z ||
                                      flo2 > b1.y)) << i;
              }
            }
            e = base + 31;  // nothing of this block consumed yet; (e >> 5) names the block
          }
        } else {
          e = (e & ~31) | (__ffs(emask) - 1);
          emask &= emask - 1u;
          const double* pl = ev.plane + 4 * e;
          const double off = pl[0] * pp[0] + pl[1] * pp[1] + pl[2] * pp[2] - pl[3];
          // robot triangles with a corner that is not strictly above / not strictly below the
          // plane of e (a triangle in both sets straddles or touches the plane), and with a
          // corner that is not beyond edge k of e (a triangle whose corners all are cannot touch
          // e; without this cull a robot crossing the plane of a face next to e, not inside it,
          // spent all its triangles on e: 45 % of the pair tests ended that way)
          unsigned not_above = 0u, not_below = 0u, in0 = 0u, in1 = 0u, in2 = 0u;
          const double* ed = ev.edge + 12 * e;
          const double o0 = ed[0] * pp[0] + ed[1] * pp[1] + ed[2] * pp[2] - ed[3];
          const double o1 = ed[4] * pp[0] + ed[5] * pp[1] + ed[6] * pp[2] - ed[7];
          const double o2 = ed[8] * pp[0] + ed[9] * pp[1] + ed[10] * pp[2] - ed[11];
          if (POSE == 0 && TABLE) {
            const double* row = nv + (size_t)e * collide_table_stride(rb.V);
            for (int v = 0; v < rb.V; ++v) {
              const double dist = row[v] + off;
              const unsigned tv = rb.vtri[v];
              not_above |= dist > 0.0 ? 0u : tv;
              not_below |= dist < 0.0 ? 0u : tv;
              in0 |= row[rb.V + v] + o0 > 0.0 ? 0u : tv;
              in1 |= row[2 * rb.V + v] + o1 > 0.0 ? 0u : tv;
              in2 |= row[3 * rb.V + v] + o2 > 0.0 ? 0u : tv;
            }
          } else {
            // planes in the robot frame: n.(R v + T) - d = (R^T n).v + (n.T - d)
            double m[4][3];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const double* q = j == 0 ? pl : ed + 4 * (j - 1);
              if (POSE == 1) {   // R^T q for a rotation about z
                m[j][0] = R[0] * q[0] + R[3] * q[1];
                m[j][1] = R[0] * q[1] - R[3] * q[0];
                m[j][2] = q[2];
              } else {
                m[j][0] = R[0] * q[0] + R[3] * q[1] + R[6] * q[2];
                m[j][1] = R[1] * q[0] + R[4] * q[1] + R[7] * q[2];
                m[j][2] = R[2] * q[0] + R[5] * q[1] + R[8] * q[2];
              }
            }
            for (int v = 0; v < rb.V; ++v) {
              const double* p = rb.vert + 3 * v;
              const double dist = m[0][0] * p[0] + m[0][1] * p[1] + m[0][2] * p[2] + off;
              const unsigned tv = rb.vtri[v];
              not_above |= dist > 0.0 ? 0u : tv;
              not_below |= dist < 0.0 ? 0u : tv;
              in0 |= m[1][0] * p[0] + m[1][1] * p[1] + m[1][2] * p[2] + o0 > 0.0 ? 0u : tv;
              in1 |= m[2][0] * p[0] + m[2][1] * p[1] + m[2][2] * p[2] + o1 > 0.0 ? 0u : tv;
              in2 |= m[3][0] * p[0] + m[3][1] * p[1] + m[3][2] * p[2] + o2 > 0.0 ? 0u : tv;
            }
          }
          need = not_above & not_below & in0 & in1 & in2;
        }
      }
      if (!exhausted) {
        r = __ffs(need) - 1;
        need &= need - 1u;
        have = true;
      }
    }
    if (have) {
      const double* pr = rb.tri + 9 * r;
      V3 P1, P2, P3;
      if (POSE == 0) {
        P1 = {pr[0] + pp[0], pr[1] + pp[1], pr[2] + pp[2]};
        P2 = {pr[3] + pp[0], pr[4] + pp[1], pr[5] + pp[2]};
        P3 = {pr[6] + pp[0], pr[7] + pp[1], pr[8] + pp[2]};
      } else if (POSE == 1) {
        const Yaw y = {R[0], R[3]};
        P1 = xform_yaw(y, pp, pr); P2 = xform_yaw(y, pp, pr + 3); P3 = xform_yaw(y, pp, pr + 6);
      } else {
        P1 = xform(R, pp, pr); P2 = xform(R, pp, pr + 3); P3 = xform(R, pp, pr + 6);
      }
      const double* qe = ev.tri + 9 * e;
      const V3 Q1 = {qe[0], qe[1], qe[2]}, Q2 = {qe[3], qe[4], qe[5]}, Q3 = {qe[6], qe[7], qe[8]};
      if (triangles_intersect_interval(P1, P2, P3, Q1, Q2, Q3)) hit = true;
    }
  }
  // a pose whose last candidate was just consumed and missed may still have later env triangles:
  // it is "exhausted" only when the cursor ran off the end
  if (valid && (hit || exhausted)) report(id0, id1, hit);
  ring_push<POSE>(ring, tail, valid && !hit && !exhausted, pp, id0, id1, e, need, emask);
}
#endif  // __CUDACC__

}  // namespace mst
