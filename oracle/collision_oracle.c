/*
 * C restatement of the collision oracle (TEST INFRASTRUCTURE ONLY; see the header of
 * oracle/collision_oracle.py for what it follows and why parity is UNPINNED: the
 * reference's arithmetic lives in FCL, which is not in the reference tree).
 *
 * Same semantics as collision_oracle.collide_poses: robot mesh at pose (R, T) against the
 * environment at identity, hit iff some triangle pair is not separated on any of the 17
 * axes of FCL's intersect_Triangle/project6 (src/RigidBodyPlanners/fcl_checker.py:93-100
 * is the call site).  Used (a) as an independent check of the numpy oracle and (b) as the
 * timed CPU arm of bench.py, where a brute-force numpy SAT would be unfairly slow: like
 * FCL's BVH it prunes with bounding boxes and leaves the SAT on the first separating axis.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/build_oracle.py).
 */
#include <math.h>
#include <stddef.h>

typedef struct { double x, y, z; } v3;

static v3 sub(v3 a, v3 b) { v3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static v3 cross(v3 a, v3 b) {
  v3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
  return r;
}
static double dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static double max3(double a, double b, double c) { double m = a > b ? a : b; return m > c ? m : c; }
static double min3(double a, double b, double c) { double m = a < b ? a : b; return m < c ? m : c; }

/* 1 when the axis does not separate the triangles (p1 is the origin after translation) */
static int project6(v3 ax, v3 p1, v3 p2, v3 p3, v3 q1, v3 q2, v3 q3) {
  double a1 = dot(ax, p1), a2 = dot(ax, p2), a3 = dot(ax, p3);
  double b1 = dot(ax, q1), b2 = dot(ax, q2), b3 = dot(ax, q3);
  double mx1 = max3(a1, a2, a3), mn1 = min3(a1, a2, a3);
  double mx2 = max3(b1, b2, b3), mn2 = min3(b1, b2, b3);
  if (mn1 > mx2) return 0;
  if (mn2 > mx1) return 0;
  return 1;
}

static int tri_tri(v3 P1, v3 P2, v3 P3, v3 Q1, v3 Q2, v3 Q3) {
  v3 p1 = sub(P1, P1), p2 = sub(P2, P1), p3 = sub(P3, P1);
  v3 q1 = sub(Q1, P1), q2 = sub(Q2, P1), q3 = sub(Q3, P1);
  v3 e[3], f[3], n1, m1;
  int i, j;
  e[0] = sub(p2, p1); e[1] = sub(p3, p2); e[2] = sub(p1, p3);
  f[0] = sub(q2, q1); f[1] = sub(q3, q2); f[2] = sub(q1, q3);
  n1 = cross(e[0], e[1]);
  m1 = cross(f[0], f[1]);
  if (!project6(n1, p1, p2, p3, q1, q2, q3)) return 0;
  if (!project6(m1, p1, p2, p3, q1, q2, q3)) return 0;
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j)
      if (!project6(cross(e[i], f[j]), p1, p2, p3, q1, q2, q3)) return 0;
  for (i = 0; i < 3; ++i)
    if (!project6(cross(e[i], n1), p1, p2, p3, q1, q2, q3)) return 0;
  for (i = 0; i < 3; ++i)
    if (!project6(cross(f[i], m1), p1, p2, p3, q1, q2, q3)) return 0;
  return 1;
}

static void quat_matrix(double x, double y, double z, double w, double* R) {
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}

/*
 * robot[Tr][9], env[Te][9] corner coordinates; pose[P][pose_dim], pose_dim 4 = (x,y,z,yaw)
 * with q = quaternion_from_euler(0,0,yaw), 7 = (x,y,z,qx,qy,qz,qw); hit[P] out.
 * prune != 0 skips pairs whose axis-aligned boxes are disjoint (same answer, faster).
 */
int oracle_collide_poses(const double* robot, int Tr, const double* env, int Te, const double* pose,
                         long P, int pose_dim, int prune, unsigned char* hit) {
  long p;
  if (pose_dim != 4 && pose_dim != 7) return -1;
  for (p = 0; p < P; ++p) {
    const double* ps = pose + p * pose_dim;
    double R[9];
    int r, e, c, found = 0;
    if (pose_dim == 4) quat_matrix(0.0, 0.0, sin(ps[3] / 2.0), cos(ps[3] / 2.0), R);
    else quat_matrix(ps[3], ps[4], ps[5], ps[6], R);
    for (r = 0; r < Tr && !found; ++r) {
      v3 w[3];
      double lo[3], hi[3];
      for (c = 0; c < 3; ++c) {
        const double* v = robot + 9 * r + 3 * c;
        w[c].x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2] + ps[0];
        w[c].y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2] + ps[1];
        w[c].z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2] + ps[2];
      }
      lo[0] = min3(w[0].x, w[1].x, w[2].x); hi[0] = max3(w[0].x, w[1].x, w[2].x);
      lo[1] = min3(w[0].y, w[1].y, w[2].y); hi[1] = max3(w[0].y, w[1].y, w[2].y);
      lo[2] = min3(w[0].z, w[1].z, w[2].z); hi[2] = max3(w[0].z, w[1].z, w[2].z);
      for (e = 0; e < Te; ++e) {
        const double* q = env + 9 * e;
        v3 Q1 = {q[0], q[1], q[2]}, Q2 = {q[3], q[4], q[5]}, Q3 = {q[6], q[7], q[8]};
        if (prune) {
          if (hi[0] < min3(Q1.x, Q2.x, Q3.x) || lo[0] > max3(Q1.x, Q2.x, Q3.x) ||
              hi[1] < min3(Q1.y, Q2.y, Q3.y) || lo[1] > max3(Q1.y, Q2.y, Q3.y) ||
              hi[2] < min3(Q1.z, Q2.z, Q3.z) || lo[2] > max3(Q1.z, Q2.z, Q3.z))
            continue;
        }
        if (tri_tri(w[0], w[1], w[2], Q1, Q2, Q3)) { found = 1; break; }
      }
    }
    hit[p] = (unsigned char)found;
  }
  return 0;
}
