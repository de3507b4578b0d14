// Fused sample -> pose -> collide -> any-hit kernel of the pipeline.
//
// What it replaces in the reference: PiecewisePolynomial.eval at S sample times
// (src/optimizations/uav_trajectory.py:154-169, sampled the way
// src/trajectory_visualising/visualization.py:53 samples), the robot pose at each sample
// (isStateValid's pos / yaw quaternion, src/RigidBodyPlanners/RB_planning_sep_coll_check.py:
// 208-215) and Fcl_checker.check_collision (src/RigidBodyPlanners/fcl_checker.py:93-100).
//
// Data flow: coefficients and durations come straight from the solver kernels' output
// (still L2 resident: the host wrapper walks the batch in L2-sized chunks), both meshes
// are staged once per CTA into shared memory with one bulk (TMA) copy each, sampled
// positions never leave registers; only hit[B][S] (1 byte per sample, coalesced) and
// any_hit[B] are written.
//
// Mapping: a CTA walks tiles of TB trajectories; inside a tile the (trajectory, sample)
// pairs are flattened over the threads, so a warp holds 32 CONSECUTIVE samples of one
// trajectory (neighbouring poses: the culling decisions of the lanes agree and the
// coefficient loads are warp-wide broadcasts).
//
// The evaluation is bit-identical to mst_sample_batch (same running-sum piece search, same
// non-fused Horner), so pipeline flags equal "sample, then collide" exactly.
#include "collide_core.cuh"
#include "stage.cuh"

namespace mst {

constexpr int FUSED_THREADS = 256;
constexpr int FUSED_TB = 32;  // trajectories per tile

template <int K>
__global__ void __launch_bounds__(FUSED_THREADS)
sample_collide_kernel(const double* __restrict__ coef, const double* __restrict__ dur, int B, int n, int S,
                      const void* __restrict__ robot_img, MeshLayout rl, MeshBounds rbb,
                      const void* __restrict__ env_img, MeshLayout el, MeshBounds evb,
                      uint8_t* __restrict__ hit, uint8_t* __restrict__ any_hit) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ int any_flag[FUSED_TB];
  stage_meshes(smem_raw, robot_img, rl.bytes, env_img, el.bytes, &bar);
  const MeshView rb = mesh_view(smem_raw, rl);
  const MeshView ev = mesh_view(smem_raw + rl.bytes, el);
  const bool culled = rb.V <= 64;
  // per-tile tables behind the meshes: knots[TB][n+1] (running sums), dt[TB]
  double* knots = reinterpret_cast<double*>(smem_raw + rl.bytes + el.bytes);
  double* dts = knots + FUSED_TB * (n + 1);

  const int tiles = (B + FUSED_TB - 1) / FUSED_TB;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int b0 = tile * FUSED_TB;
    const int nb = min(FUSED_TB, B - b0);
    __syncthreads();  // previous tile's tables are no longer read
    if (threadIdx.x < nb) {
      const double* T = dur + (size_t)(b0 + threadIdx.x) * n;
      double* kn = knots + threadIdx.x * (n + 1);
      double acc = 0.0;
      kn[0] = 0.0;
      for (int i = 0; i < n; ++i) { acc = __dadd_rn(acc, T[i]); kn[i + 1] = acc; }
      dts[threadIdx.x] = __ddiv_rn(acc, (double)S);
      any_flag[threadIdx.x] = 0;
    }
    __syncthreads();
    const int work = nb * S;
    for (int idx = threadIdx.x; idx < work; idx += FUSED_THREADS) {
      const int tl = idx / S;
      const int s = idx - tl * S;
      const double* kn = knots + tl * (n + 1);
      const double t = __dmul_rn((double)s, dts[tl]);
      // PiecewisePolynomial.eval: first piece with t < acc + T_i, else the last one at
      // t - sum(T[:-1])
      int piece = n - 1;
      for (int i = 0; i < n; ++i)
        if (t < kn[i + 1]) { piece = i; break; }
      const double local = __dsub_rn(t, kn[piece]);
      const double* cp = coef + (((size_t)(b0 + tl) * n + piece) * K) * MST_NCOEF;
      double pos[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const double2* src = reinterpret_cast<const double2*>(cp + k * MST_NCOEF);
        const double2 c01 = __ldg(src), c23 = __ldg(src + 1), c45 = __ldg(src + 2), c67 = __ldg(src + 3);
        double x = c67.y;  // 0*t + c7
        x = __dadd_rn(__dmul_rn(x, local), c67.x);
        x = __dadd_rn(__dmul_rn(x, local), c45.y);
        x = __dadd_rn(__dmul_rn(x, local), c45.x);
        x = __dadd_rn(__dmul_rn(x, local), c23.y);
        x = __dadd_rn(__dmul_rn(x, local), c23.x);
        x = __dadd_rn(__dmul_rn(x, local), c01.y);
        x = __dadd_rn(__dmul_rn(x, local), c01.x);
        pos[k] = x;
      }
      double R[9], T[3] = {pos[0], pos[1], pos[2]};
      bool h;
      if (K == 3) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        h = culled ? robot_hits_env_culled<false>(R, T, rb, rbb, ev, evb, true)
                   : robot_hits_env(R, T, rb.tri, rb.T, ev.tri, ev.box, ev.T, evb.root, rbb.radius, true);
      } else {
        double sn, cs;
        sincos(pos[K - 1] * 0.5, &sn, &cs);
        quat_to_matrix(0.0, 0.0, sn, cs, R);
        h = culled ? robot_hits_env_culled<true>(R, T, rb, rbb, ev, evb, true)
                   : robot_hits_env(R, T, rb.tri, rb.T, ev.tri, ev.box, ev.T, evb.root, rbb.radius, true);
      }
      hit[(size_t)b0 * S + idx] = h ? 1 : 0;
      if (h) any_flag[tl] = 1;  // benign race: every writer stores 1
    }
    __syncthreads();
    if (threadIdx.x < nb) any_hit[b0 + threadIdx.x] = any_flag[threadIdx.x] ? 1 : 0;
  }
}

int launch_sample_collide(const double* coef, const double* dur, int B, int n, int K, int S,
                          const mst_mesh* robot, const mst_mesh* env, uint8_t* hit, uint8_t* any_hit,
                          cudaStream_t stream) {
  if (B == 0) return MST_OK;
  const size_t smem = robot->layout.bytes + env->layout.bytes + sizeof(double) * FUSED_TB * (size_t)(n + 2);
  if (smem > MST_MAX_SMEM - 2048) return MST_ERR_TOO_LARGE;
  auto kern = K == 3 ? sample_collide_kernel<3> : sample_collide_kernel<4>;
  if (smem > 40 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MST_MAX_SMEM);
    if (e != cudaSuccess) { note_cuda_error(e); return MST_ERR_CUDA; }
  }
  const int tiles = (B + FUSED_TB - 1) / FUSED_TB;
  int blocks = tiles;
  const int cap = MST_SM_COUNT * 8;  // persistent CTAs; tiles are walked in a grid-stride loop
  if (blocks > cap) blocks = cap;
  kern<<<blocks, FUSED_THREADS, smem, stream>>>(coef, dur, B, n, S, robot->d_image, robot->layout,
                                                 robot->bounds, env->d_image, env->layout, env->bounds, hit,
                                                 any_hit);
  return check_launch();
}

}  // namespace mst
