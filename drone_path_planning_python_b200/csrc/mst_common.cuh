// Shared device helpers and launch plumbing for libmst (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mst.h"

#define MST_NCOEF 8
#define MST_SM_COUNT 148            // B200: 2 dies x 74 SMs
#define MST_MAX_SMEM (227 * 1024)   // opt-in dynamic shared memory per CTA

namespace mst {

// set by every failed CUDA call; read back through mst_last_cuda_error()
void note_cuda_error(cudaError_t e);
int check_launch();

// k!/(k-j)! for 0 <= j <= k <= 7 (exact in double)
__host__ __device__ __forceinline__ double falling_factorial(int k, int j) {
  double f = 1.0;
  for (int i = 0; i < j; ++i) f *= (double)(k - i);
  return f;
}

// x^e for small non-negative integer e by repeated multiplication; x^0 == 1 also for
// x == 0, as Python's float ** int gives (uav_trajectory.py:34)
__host__ __device__ __forceinline__ double ipow(double x, int e) {
  double r = 1.0;
  for (int i = 0; i < e; ++i) r *= x;
  return r;
}

// rotation matrix (row-major 3x3) of the quaternion (x,y,z,w) — the matrix FCL builds
// from fcl.Transform(q_wxyz, T) at fcl_checker.py:54-59
__device__ __forceinline__ void quat_to_matrix(double x, double y, double z, double w,
                                               double* R) {
  R[0] = 1.0 - 2.0 * (y * y + z * z);
  R[1] = 2.0 * (x * y - z * w);
  R[2] = 2.0 * (x * z + y * w);
  R[3] = 2.0 * (x * y + z * w);
  R[4] = 1.0 - 2.0 * (x * x + z * z);
  R[5] = 2.0 * (y * z - x * w);
  R[6] = 2.0 * (x * z - y * w);
  R[7] = 2.0 * (y * z + x * w);
  R[8] = 1.0 - 2.0 * (x * x + y * y);
}

// pose_dim 4: (x,y,z,yaw) with q = quaternion_from_euler(0,0,yaw)
// (RB_planning_sep_coll_check.py:212); pose_dim 7: (x,y,z,qx,qy,qz,qw)
__device__ __forceinline__ void pose_to_transform(const double* pose, int pose_dim,
                                                  double* R, double* T) {
  T[0] = pose[0];
  T[1] = pose[1];
  T[2] = pose[2];
  if (pose_dim == 4) {
    double s, c;
    sincos(pose[3] * 0.5, &s, &c);
    quat_to_matrix(0.0, 0.0, s, c, R);
  } else {
    quat_to_matrix(pose[3], pose[4], pose[5], pose[6], R);
  }
}

}  // namespace mst

// device-side mesh record behind the opaque mst_mesh_t handle
struct mst_mesh {
  int T;             // triangle count
  double* d_tri;     // [T][9]   corners, device
  double* d_box;     // [T][6]   per-triangle AABB (min xyz, max xyz), device
  double root[6];    // AABB of the whole mesh (host copy)
  double radius;     // max |vertex| (bound of the mesh under any rotation about its origin)
  double* h_tri;     // host copy of the corners (for packing into constant/shared memory)
};
