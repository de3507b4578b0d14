// Shared device helpers and launch plumbing for libmst (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mst.h"

#define MST_NCOEF 8
#define MST_SM_COUNT (mst::sm_count())  // SMs of the current device (148 on B200: 2 dies x 74)
#define MST_MAX_SMEM (227 * 1024)   // opt-in dynamic shared memory per CTA
// mesh images (+ plane x vertex table) up to this size are staged into every CTA's shared memory;
// larger ones are read in place from device memory (4 resident CTAs per SM still fit below it)
#define MST_STAGE_LIMIT (40 * 1024)

namespace mst {

// set by every failed CUDA call; read back through mst_last_cuda_error()
void note_cuda_error(cudaError_t e);
int check_launch();
// opt a kernel in to `dynamic_bytes` of dynamic shared memory (on top of its static usage);
// MST_ERR_TOO_LARGE when static + dynamic exceed what one CTA can have on sm_100a
int allow_dynamic_smem(const void* kernel, size_t dynamic_bytes);
// multiprocessor count of the current device (queried once per device)
int sm_count();

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

// A mesh as the kernels see it: one contiguous, 16-byte-aligned image that a single bulk
// (TMA) copy stages into shared memory.  Layout in doubles from `base`:
//   tri[T][9] | box[T][6] | plane[T][4] | vert[V][3] | pad to 16 B | idx[T][3] int32 | pad |
//   mask[T] uint64 (bit v set when unique vertex v is a corner of the triangle)
struct MeshView {
  int T, V;
  const double* tri;    // corners
  const double* box;    // per-triangle AABB: min xyz, max xyz
  const double* plane;  // nx, ny, nz, d  with n = (Q2-Q1) x (Q3-Q2), d = n . Q1
  const double* vert;   // unique vertices
  const int* idx;       // corner -> unique vertex
  const unsigned long long* mask;  // per triangle: bits of its (first 64) unique vertices
  const unsigned* vtri;            // per unique vertex: bits of the (first 32) triangles that use it
  const float* fbox;               // per triangle, 8 floats: the AABB rounded OUTWARD (min xyz, max xyz, 2 pad)
  const double* edge;              // per triangle, 3 x (mx, my, mz, c): in-plane outward normal of edge k and its
                                   // offset, so that m . x - c > 0 only for points beyond that edge's line
  const float* bbox;               // per BLOCK of 32 consecutive triangles, 8 floats: the block's box rounded outward
                                   // (the level above the per-triangle boxes; triangles are stored in Morton order
                                   // of their centroids, so a block is spatially compact)
};

struct MeshLayout {
  int T, V;
  size_t off_box, off_plane, off_vert, off_idx, off_mask, off_vtri, off_fbox, off_edge, off_bbox, bytes;  // byte offsets from base
};

__host__ __device__ __forceinline__ MeshLayout mesh_layout(int T, int V) {
  MeshLayout L;
  L.T = T; L.V = V;
  size_t o = sizeof(double) * 9 * (size_t)T;
  L.off_box = o;   o += sizeof(double) * 6 * (size_t)T;
  L.off_plane = o; o += sizeof(double) * 4 * (size_t)T;
  L.off_vert = o;  o += sizeof(double) * 3 * (size_t)V;
  o = (o + 15) & ~(size_t)15;
  L.off_idx = o;   o += sizeof(int) * 3 * (size_t)T;
  o = (o + 15) & ~(size_t)15;
  L.off_mask = o;  o += sizeof(unsigned long long) * (size_t)T;
  L.off_vtri = o;  o += sizeof(unsigned) * (size_t)V;
  o = (o + 15) & ~(size_t)15;
  L.off_fbox = o;  o += sizeof(float) * 8 * (size_t)T;
  L.off_edge = o;  o += sizeof(double) * 12 * (size_t)T;
  L.off_bbox = o;  o += sizeof(float) * 8 * (size_t)((T + 31) / 32);
  L.bytes = (o + 15) & ~(size_t)15;
  return L;
}

__host__ __device__ __forceinline__ MeshView mesh_view(const void* base, const MeshLayout& L) {
  const char* b = (const char*)base;
  MeshView v;
  v.T = L.T; v.V = L.V;
  v.tri = (const double*)b;
  v.box = (const double*)(b + L.off_box);
  v.plane = (const double*)(b + L.off_plane);
  v.vert = (const double*)(b + L.off_vert);
  v.idx = (const int*)(b + L.off_idx);
  v.mask = (const unsigned long long*)(b + L.off_mask);
  v.vtri = (const unsigned*)(b + L.off_vtri);
  v.fbox = (const float*)(b + L.off_fbox);
  v.edge = (const double*)(b + L.off_edge);
  v.bbox = (const float*)(b + L.off_bbox);
  return v;
}

// whole-mesh bounds, passed to kernels by value
struct MeshBounds {
  double root[6];   // AABB in the mesh's own frame
  double radius;    // max |vertex|, rounded up: bound of the mesh under any rotation about its origin
  double rxy;       // max sqrt(x^2 + y^2) over the vertices, rounded up: bound on x / y under any rotation about z
};

// host-side record behind the opaque mst_mesh_t handle
struct mst_mesh {
  int T, V;
  MeshLayout layout;
  void* d_image;       // device copy of the image
  void* h_image;       // host copy
  MeshBounds bounds;
  // views into the device image kept for the kernels that take raw arrays
  double* d_tri;
  double* d_box;
};
