// Mesh-mesh collision queries: Fcl_mesh / Fcl_checker.check_collision
// (src/RigidBodyPlanners/fcl_checker.py:13-59,93-103) as driven by isStateValid
// (src/RigidBodyPlanners/RB_planning_sep_coll_check.py:208-215).
//
// Semantics (the reference delegates to FCL, whose mesh-mesh leaf test this follows):
// robot mesh at pose (R, T), environment at identity, hit iff some triangle pair
// intersects; a pair is tested with the 17-axis separating-axis test, both triangles
// translated by -P1, and is separated on an axis only when min1 > max2 or min2 > max1
// (strict: touching counts as a hit).
//
// Layout: both meshes are staged once per CTA in shared memory (<= 56 triangles x 72 B
// for every shipped mesh) together with the environment's per-triangle AABBs; one thread
// walks one pose.  Culling is conservative and exact-safe: bounding sphere of the robot
// vs the environment's root box, then triangle AABB vs triangle AABB with non-strict
// comparisons, then the SAT with early exit on the first separating axis.
#include <stdlib.h>
#include <string.h>

#include "collide_core.cuh"

namespace mst {

struct RootBox { double v[6]; double radius; };

// pose_dim 3: (x,y,z) with identity rotation; 4: (x,y,z,yaw); 7: (x,y,z,qx,qy,qz,qw)
__global__ void __launch_bounds__(128)
collide_kernel(const double* __restrict__ rtri_g, int Tr, const double* __restrict__ etri_g,
               const double* __restrict__ ebox_g, int Te, RootBox root,
               const double* __restrict__ pose, long long P, int pose_dim,
               uint8_t* __restrict__ hit) {
  extern __shared__ double sm[];
  double* rtri = sm;
  double* etri = rtri + 9 * Tr;
  double* ebox = etri + 9 * Te;
  for (int i = threadIdx.x; i < 9 * Tr; i += blockDim.x) rtri[i] = rtri_g[i];
  for (int i = threadIdx.x; i < 9 * Te; i += blockDim.x) etri[i] = etri_g[i];
  for (int i = threadIdx.x; i < 6 * Te; i += blockDim.x) ebox[i] = ebox_g[i];
  __syncthreads();
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < P;
       idx += (long long)gridDim.x * blockDim.x) {
    const double* ps = pose + idx * pose_dim;
    double R[9], T[3];
    if (pose_dim == 3) {
      T[0] = ps[0]; T[1] = ps[1]; T[2] = ps[2];
      R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    } else {
      pose_to_transform(ps, pose_dim, R, T);
    }
    hit[idx] = robot_hits_env(R, T, rtri, Tr, etri, ebox, Te, root.v, root.radius, pose_dim != 7) ? 1 : 0;
  }
}

// any_hit[b] = OR_s hit[b][s]; one warp per trajectory
__global__ void any_hit_kernel(const uint8_t* __restrict__ hit, int B, int S, uint8_t* __restrict__ any_hit) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long b = warp; b < B; b += nwarps) {
    int any = 0;
    for (int s = lane; s < S; s += 32) any |= hit[b * S + s];
    any = __reduce_or_sync(0xffffffffu, any);
    if (lane == 0) any_hit[b] = any ? 1 : 0;
  }
}

int launch_collide(const mst_mesh* robot, const mst_mesh* env, const double* pose, long long P,
                   int pose_dim, uint8_t* hit, cudaStream_t stream) {
  if (P == 0) return MST_OK;
  const size_t smem = sizeof(double) * (9 * (size_t)robot->T + 15 * (size_t)env->T);
  if (smem > MST_MAX_SMEM) return MST_ERR_TOO_LARGE;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(collide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         MST_MAX_SMEM);
    if (e != cudaSuccess) { note_cuda_error(e); return MST_ERR_CUDA; }
  }
  RootBox root;
  for (int i = 0; i < 6; ++i) root.v[i] = env->root[i];
  root.radius = robot->radius;
  long long blocks = (P + 127) / 128;
  const long long cap = (long long)MST_SM_COUNT * 16;
  if (blocks > cap) blocks = cap;
  collide_kernel<<<(unsigned)blocks, 128, smem, stream>>>(robot->d_tri, robot->T, env->d_tri, env->d_box,
                                                          env->T, root, pose, P, pose_dim, hit);
  return check_launch();
}

int launch_any_hit(const uint8_t* hit, int B, int S, uint8_t* any_hit, cudaStream_t stream) {
  if (B == 0) return MST_OK;
  long long blocks = ((long long)B * 32 + 255) / 256;
  const long long cap = (long long)MST_SM_COUNT * 16;
  if (blocks > cap) blocks = cap;
  any_hit_kernel<<<(unsigned)blocks, 256, 0, stream>>>(hit, B, S, any_hit);
  return check_launch();
}

}  // namespace mst

// ------------------------------------------------------------------ mesh handles (C ABI)
extern "C" int mst_mesh_create(const double* tri, int T, mst_mesh_t* out) {
  if (!tri || !out || T < 0) return MST_ERR_INVALID;
  mst_mesh* m = (mst_mesh*)calloc(1, sizeof(mst_mesh));
  if (!m) return MST_ERR_NOMEM;
  m->T = T;
  const size_t nt = (size_t)(T > 0 ? T : 1);
  m->h_tri = (double*)malloc(sizeof(double) * 9 * nt);
  double* box = (double*)malloc(sizeof(double) * 6 * nt);
  if (!m->h_tri || !box) { free(m->h_tri); free(box); free(m); return MST_ERR_NOMEM; }
  memcpy(m->h_tri, tri, sizeof(double) * 9 * (size_t)T);
  for (int k = 0; k < 3; ++k) { m->root[k] = 1e300; m->root[3 + k] = -1e300; }
  m->radius = 0.0;
  for (int t = 0; t < T; ++t) {
    for (int k = 0; k < 3; ++k) { box[6 * t + k] = 1e300; box[6 * t + 3 + k] = -1e300; }
    for (int c = 0; c < 3; ++c) {
      double r2 = 0.0;
      for (int k = 0; k < 3; ++k) {
        const double v = tri[9 * t + 3 * c + k];
        if (v < box[6 * t + k]) box[6 * t + k] = v;
        if (v > box[6 * t + 3 + k]) box[6 * t + 3 + k] = v;
        r2 += v * v;
      }
      // rounded up a little: the cull must stay conservative
      const double r = sqrt(r2) * (1.0 + 1e-12) + 1e-300;
      if (r > m->radius) m->radius = r;
    }
    for (int k = 0; k < 3; ++k) {
      if (box[6 * t + k] < m->root[k]) m->root[k] = box[6 * t + k];
      if (box[6 * t + 3 + k] > m->root[3 + k]) m->root[3 + k] = box[6 * t + 3 + k];
    }
  }
  cudaError_t e = cudaMalloc(&m->d_tri, sizeof(double) * 9 * nt);
  if (e == cudaSuccess) e = cudaMalloc(&m->d_box, sizeof(double) * 6 * nt);
  if (e == cudaSuccess && T > 0) e = cudaMemcpy(m->d_tri, tri, sizeof(double) * 9 * (size_t)T, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && T > 0) e = cudaMemcpy(m->d_box, box, sizeof(double) * 6 * (size_t)T, cudaMemcpyHostToDevice);
  free(box);
  if (e != cudaSuccess) {
    mst::note_cuda_error(e);
    if (m->d_tri) cudaFree(m->d_tri);
    if (m->d_box) cudaFree(m->d_box);
    free(m->h_tri);
    free(m);
    return e == cudaErrorMemoryAllocation ? MST_ERR_NOMEM : MST_ERR_CUDA;
  }
  *out = m;
  return MST_OK;
}

extern "C" int mst_mesh_destroy(mst_mesh_t mesh) {
  if (!mesh) return MST_ERR_INVALID;
  cudaFree(mesh->d_tri);
  cudaFree(mesh->d_box);
  free(mesh->h_tri);
  free(mesh);
  return MST_OK;
}

extern "C" int mst_mesh_triangle_count(mst_mesh_t mesh) { return mesh ? mesh->T : MST_ERR_INVALID; }
