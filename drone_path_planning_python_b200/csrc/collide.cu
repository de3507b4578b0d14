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
// Layout: both meshes travel as one contiguous image each (mesh_image.cuh) that a CTA stages
// into shared memory with one bulk (TMA) copy; poses run through the warp-level collision
// engine of collide_core.cuh (root-box cull, pose ring, per-lane cursors with early exit).
#include <stdlib.h>
#include <string.h>

#include "collide_core.cuh"
#include "mesh_image.cuh"
#include "stage.cuh"

namespace mst {

// POSE 0: pose = (x,y,z), identity rotation; 1: (x,y,z,yaw); 2: (x,y,z,qx,qy,qz,qw)
// GLOBAL = true: the mesh images stay in device memory (environments / robots too large to stage);
// the cursor engine then walks the block boxes first, and no plane x vertex table exists.
template <int POSE, bool GLOBAL>
__global__ void __launch_bounds__(128, 4)
collide_kernel(const void* __restrict__ robot_img, MeshLayout rl, MeshBounds rbb,
               const void* __restrict__ env_img, MeshLayout el, MeshBounds evb,
               const double* __restrict__ pose, long long P, uint8_t* __restrict__ hit) {
  constexpr int NP = PoseDim<POSE>::N;
  constexpr int pose_dim = POSE == 0 ? 3 : (POSE == 1 ? 4 : 7);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ PoseRing<NP> rings[4];
  PoseRing<NP>& ring = rings[threadIdx.x >> 5];
  if (!GLOBAL) stage_meshes(smem_raw, robot_img, rl.bytes, env_img, el.bytes, &bar);
  const MeshView rb = GLOBAL ? mesh_view(robot_img, rl) : mesh_view(smem_raw, rl);
  const MeshView ev = GLOBAL ? mesh_view(env_img, el) : mesh_view(smem_raw + rl.bytes, el);
  const bool engine = collide_engine_supports(rb, ev);
  double* nv = GLOBAL ? nullptr : reinterpret_cast<double*>(smem_raw + rl.bytes + el.bytes);  // plane x vertex table
  if (!GLOBAL && engine && POSE == 0) build_plane_vertex_table(rb, ev, nv);
  __syncthreads();
  unsigned ring_head = 0u, ring_tail = 0u;  // warp-uniform
  auto report = [&](int hi32, int lo32, bool h) {
    hit[((long long)hi32 << 32) | (unsigned)lo32] = h ? 1 : 0;
  };
  // warp-uniform trip count (the ring operations vote across the warp)
  for (long long base = blockIdx.x * (long long)blockDim.x; base < P; base += (long long)gridDim.x * blockDim.x) {
    const long long idx = base + threadIdx.x;
    const bool active = idx < P;
    const double* ps = pose + (active ? idx : P - 1) * pose_dim;
    double pp[NP];
    pp[0] = ps[0]; pp[1] = ps[1]; pp[2] = ps[2];
    if (POSE == 1) sincos(ps[3] * 0.5, &pp[3], &pp[4]);
    if (POSE == 2) { pp[3] = ps[3]; pp[4] = ps[4]; pp[5] = ps[5]; pp[6] = ps[6]; }
    if (!engine) {  // meshes the bit-mask cursors cannot hold: plain per-lane test over all pairs
      double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      pose_rotation<POSE>(pp, R);
      if (active) hit[idx] = robot_hits_env(R, pp, rb.tri, rb.T, ev.tri, ev.box, ev.T, evb.root, rbb.radius, POSE != 2) ? 1 : 0;
      continue;
    }
    const bool near = active && pose_near_environment<POSE>(pp, rbb, evb);
    if (active && !near) hit[idx] = 0;
    ring_push<POSE>(ring, ring_tail, near, pp, (int)(idx >> 32), (int)(idx & 0xffffffffll), -1, 0u, 0u);
    while (ring_tail - ring_head >= 32u) ring_drain<POSE, !GLOBAL>(ring, ring_head, ring_tail, 32, rb, rbb, ev, nv, report);
  }
  while (ring_tail != ring_head)
    ring_drain<POSE, !GLOBAL>(ring, ring_head, ring_tail, (int)min(32u, ring_tail - ring_head), rb, rbb, ev, nv, report);
}

// Motion validation for a sampling planner: states (x, y, z, yaw) interpolated linearly between
// a[m] and b[m] at fractions j/steps, j = 1..steps (the end state included, the start state
// assumed valid — the discrete motion validation OMPL runs at the resolution set in
// RB_planning_sep_coll_check.py:79), every interpolated state collision-checked like
// isStateValid does (:208-215).  invalid[m] = 1 iff some state collides.  invalid[] must be
// zeroed by the caller (the launcher does it).
template <bool GLOBAL>
__global__ void __launch_bounds__(128, 4)
collide_motions_kernel(const void* __restrict__ robot_img, MeshLayout rl, MeshBounds rbb,
                       const void* __restrict__ env_img, MeshLayout el, MeshBounds evb,
                       const double* __restrict__ sa, const double* __restrict__ sb, long long M, int steps,
                       uint8_t* __restrict__ invalid) {
  constexpr int POSE = 1, NP = PoseDim<POSE>::N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ PoseRing<NP> rings[4];
  PoseRing<NP>& ring = rings[threadIdx.x >> 5];
  if (!GLOBAL) stage_meshes(smem_raw, robot_img, rl.bytes, env_img, el.bytes, &bar);
  const MeshView rb = GLOBAL ? mesh_view(robot_img, rl) : mesh_view(smem_raw, rl);
  const MeshView ev = GLOBAL ? mesh_view(env_img, el) : mesh_view(smem_raw + rl.bytes, el);
  const bool engine = collide_engine_supports(rb, ev);
  const double* nv = nullptr;
  unsigned ring_head = 0u, ring_tail = 0u;
  auto report = [&](int hi32, int lo32, bool h) {
    if (h) invalid[((long long)hi32 << 32) | (unsigned)lo32] = 1;
  };
  const long long total = M * steps;
  for (long long base = blockIdx.x * (long long)blockDim.x; base < total; base += (long long)gridDim.x * blockDim.x) {
    const long long idx = base + threadIdx.x;
    const bool active = idx < total;
    const long long m = (active ? idx : total - 1) / steps;
    const int j = (int)((active ? idx : total - 1) - m * steps) + 1;
    const double f = (double)j / (double)steps;
    const double* a = sa + 4 * m;
    const double* b = sb + 4 * m;
    double pp[NP];
    pp[0] = a[0] + (b[0] - a[0]) * f;
    pp[1] = a[1] + (b[1] - a[1]) * f;
    pp[2] = a[2] + (b[2] - a[2]) * f;
    const double yaw = a[3] + (b[3] - a[3]) * f;
    sincos(yaw * 0.5, &pp[3], &pp[4]);
    if (!engine) {
      double R[9];
      pose_rotation<POSE>(pp, R);
      if (active && robot_hits_env(R, pp, rb.tri, rb.T, ev.tri, ev.box, ev.T, evb.root, rbb.radius, true)) invalid[m] = 1;
      continue;
    }
    const bool near = active && pose_near_environment<POSE>(pp, rbb, evb);
    ring_push<POSE>(ring, ring_tail, near, pp, (int)(m >> 32), (int)(m & 0xffffffffll), -1, 0u, 0u);
    while (ring_tail - ring_head >= 32u) ring_drain<POSE, !GLOBAL>(ring, ring_head, ring_tail, 32, rb, rbb, ev, nv, report);
  }
  while (ring_tail != ring_head)
    ring_drain<POSE, !GLOBAL>(ring, ring_head, ring_tail, (int)min(32u, ring_tail - ring_head), rb, rbb, ev, nv, report);
}

// ---------------------------------------------------------------------------------------------
// Single query with HOST arguments: the latency path of Fcl_checker.check_collision as OMPL's
// isStateValid drives it, one state at a time (RB_planning_sep_coll_check.py:208-226).  One warp;
// the pose is read from, and the answer written to, mapped pinned host memory, so a query is one
// launch and one stream synchronisation — no staging copies, no allocation.  Meshes are read in
// place from device memory (no shared-memory staging: nothing to amortise it over); the lanes
// split the environment triangles and run the host-checkable culled routine, so meshes of any
// size are accepted.
__global__ void __launch_bounds__(32)
collide_single_kernel(const void* __restrict__ robot_img, MeshLayout rl, MeshBounds rbb,
                      const void* __restrict__ env_img, MeshLayout el, MeshBounds evb,
                      const double* __restrict__ pose, int pose_dim, int* __restrict__ out) {
  const MeshView rb = mesh_view(robot_img, rl);
  const MeshView ev = mesh_view(env_img, el);
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, T[3];
  if (pose_dim == 3) { T[0] = pose[0]; T[1] = pose[1]; T[2] = pose[2]; }
  else pose_to_transform(pose, pose_dim, R, T);
  const int lane = threadIdx.x;
  bool h;
  if (rb.V <= 64)   // vertex bit masks of the culled routine
    h = pose_dim == 3 ? robot_hits_env_culled<false>(R, T, rb, rbb, ev, evb, true, lane, 32)
                      : robot_hits_env_culled<true>(R, T, rb, rbb, ev, evb, pose_dim != 7, lane, 32);
  else {
    h = false;
    for (int e = lane; e < ev.T && !h; e += 32)
      h = robot_hits_env(R, T, rb.tri, rb.T, ev.tri + 9 * e, ev.box + 6 * e, 1, evb.root, rbb.radius, pose_dim != 7);
  }
  const unsigned any = __ballot_sync(0xffffffffu, h);
  if (lane == 0) *out = any ? 1 : 0;
}

int launch_collide_motions(const mst_mesh* robot, const mst_mesh* env, const double* a, const double* b,
                           long long M, int steps, uint8_t* invalid, cudaStream_t stream) {
  if (M == 0) return MST_OK;
  cudaError_t e = cudaMemsetAsync(invalid, 0, (size_t)M, stream);
  if (e != cudaSuccess) { note_cuda_error(e); return MST_ERR_CUDA; }
  size_t smem = robot->layout.bytes + env->layout.bytes;
  const bool global = smem > MST_STAGE_LIMIT;   // large meshes are read in place (block boxes cull)
  if (global) smem = 0;
  auto kern = global ? collide_motions_kernel<true> : collide_motions_kernel<false>;
  {
    const int rc = allow_dynamic_smem((const void*)kern, smem);
    if (rc != MST_OK) return rc;
  }
  long long blocks = (M * steps + 127) / 128;
  const long long cap = (long long)MST_SM_COUNT * 4;
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, 128, smem, stream>>>(robot->d_image, robot->layout, robot->bounds,
                                                                  env->d_image, env->layout, env->bounds, a, b, M,
                                                                  steps, invalid);
  return check_launch();
}

int launch_collide(const mst_mesh* robot, const mst_mesh* env, const double* pose, long long P,
                   int pose_dim, uint8_t* hit, cudaStream_t stream) {
  if (P == 0) return MST_OK;
  // the plane x vertex table serves translation-only poses alone
  size_t smem = robot->layout.bytes + env->layout.bytes +
                (pose_dim == 3 ? sizeof(double) * collide_table_doubles(env->T, robot->V) : 0);
  const bool global = smem > MST_STAGE_LIMIT;   // large meshes are read in place (block boxes cull)
  if (global) smem = 0;
  void (*kern)(const void*, MeshLayout, MeshBounds, const void*, MeshLayout, MeshBounds, const double*, long long,
               uint8_t*);
  if (global) kern = pose_dim == 3 ? collide_kernel<0, true> : (pose_dim == 4 ? collide_kernel<1, true> : collide_kernel<2, true>);
  else kern = pose_dim == 3 ? collide_kernel<0, false> : (pose_dim == 4 ? collide_kernel<1, false> : collide_kernel<2, false>);
  {
    const int rc = allow_dynamic_smem((const void*)kern, smem);
    if (rc != MST_OK) return rc;
  }
  long long blocks = (P + 127) / 128;
  const long long cap = (long long)MST_SM_COUNT * 4;
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, 128, smem, stream>>>(robot->d_image, robot->layout, robot->bounds, env->d_image,
                                                env->layout, env->bounds, pose, P, hit);
  return check_launch();
}

}  // namespace mst

// ------------------------------------------------------------------ mesh handles (C ABI)
extern "C" int mst_mesh_create(const double* tri, int T, mst_mesh_t* out) {
  if (!tri || !out || T < 0) return MST_ERR_INVALID;
  mst_mesh* m = (mst_mesh*)calloc(1, sizeof(mst_mesh));
  if (!m) return MST_ERR_NOMEM;
  m->h_image = mst::build_mesh_image(tri, T, &m->layout, &m->bounds);
  if (!m->h_image) { free(m); return MST_ERR_NOMEM; }
  m->T = T;
  m->V = m->layout.V;
  cudaError_t e = cudaMalloc(&m->d_image, m->layout.bytes > 0 ? m->layout.bytes : 16);
  if (e == cudaSuccess && m->layout.bytes > 0)
    e = cudaMemcpy(m->d_image, m->h_image, m->layout.bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    mst::note_cuda_error(e);
    if (m->d_image) cudaFree(m->d_image);
    free(m->h_image);
    free(m);
    return e == cudaErrorMemoryAllocation ? MST_ERR_NOMEM : MST_ERR_CUDA;
  }
  m->d_tri = (double*)m->d_image;
  m->d_box = (double*)((char*)m->d_image + m->layout.off_box);
  *out = m;
  return MST_OK;
}

extern "C" int mst_mesh_destroy(mst_mesh_t mesh) {
  if (!mesh) return MST_ERR_INVALID;
  cudaFree(mesh->d_image);
  free(mesh->h_image);
  free(mesh);
  return MST_OK;
}

extern "C" int mst_mesh_triangle_count(mst_mesh_t mesh) { return mesh ? mesh->T : MST_ERR_INVALID; }

namespace {
// per calling thread: a stream and one page of mapped pinned memory (pose in, flag out)
struct SyncSlot {
  cudaStream_t stream = nullptr;
  double* h_pose = nullptr;   // host view
  double* d_pose = nullptr;   // device view of the same memory
  int device = -1;
};
thread_local SyncSlot g_slot;

int sync_slot(SyncSlot** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { mst::note_cuda_error(e); return MST_ERR_CUDA; }
  if (g_slot.stream == nullptr || g_slot.device != dev) {
    if (g_slot.h_pose) cudaFreeHost(g_slot.h_pose);
    if (g_slot.stream) cudaStreamDestroy(g_slot.stream);
    g_slot = SyncSlot();
    e = cudaStreamCreateWithFlags(&g_slot.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&g_slot.h_pose, 256, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&g_slot.d_pose, g_slot.h_pose, 0);
    if (e != cudaSuccess) { mst::note_cuda_error(e); g_slot = SyncSlot(); return MST_ERR_CUDA; }
    g_slot.device = dev;
  }
  *out = &g_slot;
  return MST_OK;
}
}  // namespace

extern "C" int mst_collide_pose_sync(mst_mesh_t robot, mst_mesh_t env, const double* pose, int pose_dim, int* hit) {
  if (!robot || !env || !pose || !hit || (pose_dim != 3 && pose_dim != 4 && pose_dim != 7)) return MST_ERR_INVALID;
  SyncSlot* slot = nullptr;
  const int rc = sync_slot(&slot);
  if (rc != MST_OK) return rc;
  for (int i = 0; i < pose_dim; ++i) slot->h_pose[i] = pose[i];
  int* h_out = reinterpret_cast<int*>(slot->h_pose + 8);
  int* d_out = reinterpret_cast<int*>(slot->d_pose + 8);
  *h_out = -1;
  mst::collide_single_kernel<<<1, 32, 0, slot->stream>>>(robot->d_image, robot->layout, robot->bounds, env->d_image,
                                                        env->layout, env->bounds, slot->d_pose, pose_dim, d_out);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(slot->stream);
  if (e != cudaSuccess) { mst::note_cuda_error(e); return MST_ERR_CUDA; }
  *hit = *h_out;
  return MST_OK;
}

