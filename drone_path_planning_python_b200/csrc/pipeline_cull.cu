// sample -> pose -> collide -> any-hit kernel of the two-launch pipeline WITH far-piece culling.
//
// Same job, same flags as sample_collide_kernel (pipeline_fused.cu) — PiecewisePolynomial.eval at S
// sample times (src/optimizations/uav_trajectory.py:154-169), the robot at every sample,
// Fcl_checker.check_collision (src/RigidBodyPlanners/fcl_checker.py:93-100) — but the solver has already
// marked the pieces whose positions provably stay clear of the obstacles' root box (farcull.cuh, 73 % of
// the pieces on the benchmark).  A warp therefore
//   * zero-fills the flag rows of its tile with 16-byte stores,
//   * lists the samples of the remaining pieces only (exclusive scan of the pieces' sample counts, one
//     packed (trajectory, piece, sample) word per candidate sample),
//   * and runs the round-1 sampling loop over that list: 32 consecutive CANDIDATE samples per step, so
//     the Horner evaluation, the root-box test and the ring push are spent on 28 % of the samples, the
//     coefficients of far pieces are never read, and the collision batches stay as coherent as before
//     (candidates keep the order trajectory, sample).
// The evaluation is the bit-identical non-fused Horner of mst_sample_batch on the same running-sum piece
// assignment; a sample that is skipped is one pose_near_environment would have rejected.
#include "collide_core.cuh"
#include <stdlib.h>

#include "stage.cuh"

namespace mst {

constexpr int CULL_THREADS = 128;
constexpr int CULL_WARPS = CULL_THREADS / 32;

template <int K>
__global__ void __launch_bounds__(CULL_THREADS, 4)
sample_collide_cull_kernel(const double* __restrict__ coef, const double* __restrict__ dur,
                           const unsigned* __restrict__ far_mask, int B, int n, int S, int WT,
                           const void* __restrict__ robot_img, MeshLayout rl, MeshBounds rbb,
                           const void* __restrict__ env_img, MeshLayout el, MeshBounds evb,
                           uint8_t* __restrict__ hit, uint8_t* __restrict__ any_hit, int* __restrict__ next_tile) {
  constexpr int POSE = K == 3 ? 0 : 1;
  constexpr int NP = PoseDim<POSE>::N;
  const unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ PoseRing<NP> rings[CULL_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PoseRing<NP>& ring = rings[warp];
  stage_meshes(smem_raw, robot_img, rl.bytes, env_img, el.bytes, &bar);
  const MeshView rb = mesh_view(smem_raw, rl);
  const MeshView ev = mesh_view(smem_raw + rl.bytes, el);
  double* nv = reinterpret_cast<double*>(smem_raw + rl.bytes + el.bytes);
  if (POSE == 0) build_plane_vertex_table(rb, ev, nv);
  __syncthreads();
  // per-warp tables behind the meshes: knots[WT][n+1] | dt[WT] | thr[WT][n] | far[WT] | cand[WT * S]
  double* tables = nv + (POSE == 0 ? collide_table_doubles(ev.T, rb.V) : 0);
  const size_t per_warp_doubles = (size_t)WT * (n + 2);
  double* knots = tables + warp * per_warp_doubles;
  double* dts = knots + (size_t)WT * (n + 1);
  int* ibase = reinterpret_cast<int*>(tables + CULL_WARPS * per_warp_doubles);
  const size_t per_warp_ints = (size_t)WT * n + WT + (size_t)WT * S;
  int* thr = ibase + warp * per_warp_ints;
  unsigned* farm = reinterpret_cast<unsigned*>(thr + (size_t)WT * n);
  unsigned* cand = farm + WT;   // (trajectory of the tile << 24) | (piece << 16) | sample
  unsigned ring_head = 0u, ring_tail = 0u;  // warp-uniform

  auto report = [&](int b, int s, bool h) {
    if (h) { hit[(size_t)b * S + s] = 1; any_hit[b] = 1; }   // rows were zero-filled
  };

  // Tiles are handed out by a device counter (zeroed by the launcher): the work of a tile depends on how many of
  // its samples are near the obstacles, and with a fixed stride over the tiles the kernel waited for its unluckiest
  // warp.  The next ticket is drawn while the current tile is processed.
  const int tiles = (B + WT - 1) / WT;
  const int first = blockIdx.x * CULL_WARPS + warp, all_warps = gridDim.x * CULL_WARPS;
  int ticket = 0;
  if (next_tile != nullptr && lane == 0) ticket = atomicAdd(next_tile, 1);
  for (int tile = next_tile != nullptr ? __shfl_sync(FULL, ticket, 0) : first; tile < tiles;
       tile = next_tile != nullptr ? __shfl_sync(FULL, ticket, 0) : tile + all_warps) {
    if (next_tile != nullptr && lane == 0) ticket = atomicAdd(next_tile, 1);
    const int b0 = tile * WT;
    const int nb = min(WT, B - b0);
    __syncwarp();  // previous tile's tables are no longer read
    if (lane < nb) {
      const double* T = dur + (size_t)(b0 + lane) * n;
      double* kn = knots + lane * (n + 1);
      double acc = 0.0;
      kn[0] = 0.0;
#pragma unroll 1
      for (int i = 0; i < n; ++i) { acc = __dadd_rn(acc, T[i]); kn[i + 1] = acc; }
      dts[lane] = __ddiv_rn(acc, (double)S);
      any_hit[b0 + lane] = 0;
      const unsigned* fm = far_mask + (size_t)(b0 + lane) * 3;
      farm[lane] = fm[0] | fm[1] | fm[2];
    }
    // the tile's flag rows: all zero until a sample collides
    {
      uint8_t* rows = hit + (size_t)b0 * S;
      const size_t bytes = (size_t)nb * S;
      if ((reinterpret_cast<uintptr_t>(rows) & 15) == 0) {
        uint4* r4 = reinterpret_cast<uint4*>(rows);
        const size_t v = bytes >> 4;
        for (size_t i = lane; i < v; i += 32) r4[i] = make_uint4(0u, 0u, 0u, 0u);
        for (size_t i = (v << 4) + lane; i < bytes; i += 32) rows[i] = 0;
      } else {
        for (size_t i = lane; i < bytes; i += 32) rows[i] = 0;
      }
    }
    __syncwarp();
    // thresholds: first s with !(s * dt < knot), found from the quotient and corrected with the
    // very comparison PiecewisePolynomial.eval makes (t is non-decreasing in s)
#pragma unroll 1
    for (int item = lane; item < nb * n; item += 32) {
      const int q = item / n, i = item - q * n;
      int first = S;
      if (i < n - 1) {
        const double knot = knots[q * (n + 1) + i + 1], dt = dts[q];
        first = (int)fmin(fmax(ceil(__ddiv_rn(knot, dt)), 0.0), (double)S);
        while (first > 0 && !(__dmul_rn((double)(first - 1), dt) < knot)) --first;
        while (first < S && __dmul_rn((double)first, dt) < knot) ++first;
      }
      thr[item] = first;  // thr[q][n-1] = S closes the last piece
    }
    __syncwarp();
    // candidate samples: those of the pieces the solver did not mark far, in (trajectory, sample) order
    int work = 0;   // warp-uniform: candidates of the tile
#pragma unroll 1
    for (int i0 = 0; i0 < nb * n; i0 += 32) {
      const int item = i0 + lane;
      int q = 0, i = 0, from = 0, cnt = 0;
      if (item < nb * n) {
        q = item / n; i = item - q * n;
        from = i ? thr[item - 1] : 0;
        const int to = thr[item];
        if (!((farm[q] >> i) & 1u) && to > from) {
          cnt = to - from;
          // the piece's coefficients towards L2 now; the sampling front reaches them in a few hundred cycles
          const char* cb = reinterpret_cast<const char*>(coef + (((size_t)(b0 + q) * n + i) * K) * MST_NCOEF);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(cb));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(cb + K * MST_NCOEF * 8 - 8));
        }
      }
      int incl = cnt;   // inclusive scan of the counts over the lanes
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += up;
      }
      const int off = work + incl - cnt;
      const unsigned tag = ((unsigned)q << 24) | ((unsigned)i << 16);
#pragma unroll 1
      for (int x = 0; x < cnt; ++x) cand[off + x] = tag | (unsigned)(from + x);
      work += __shfl_sync(FULL, incl, 31);
    }
    __syncwarp();
    for (int base = 0; base < work; base += 32) {
      const int idx = base + lane;
      const bool active = idx < work;
      const unsigned c = cand[active ? idx : work - 1];   // inactive lanes ride on the last candidate (not reported)
      const int tl = (int)(c >> 24), piece = (int)((c >> 16) & 0xffu), s = (int)(c & 0xffffu);
      const double t = __dmul_rn((double)s, dts[tl]);
      const double local = __dsub_rn(t, knots[tl * (n + 1) + piece]);
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
      double pp[NP];
      pp[0] = pos[0]; pp[1] = pos[1]; pp[2] = pos[2];
      bool reach = true;
      if (POSE == 1) {
        reach = sphere_near_environment(pp, rbb, evb);
        pp[3] = 0.0; pp[4] = 1.0;
        if (reach) sincos(pos[K - 1] * 0.5, &pp[3], &pp[4]);
      }
      const bool near = active && reach && pose_near_environment<POSE>(pp, rbb, evb);
      ring_push<POSE>(ring, ring_tail, near, pp, b0 + tl, s, -1, 0u, 0u);
      while (ring_tail - ring_head >= 32u) ring_drain<POSE>(ring, ring_head, ring_tail, 32, rb, rbb, ev, nv, report);
    }
  }
  while (ring_tail != ring_head)
    ring_drain<POSE>(ring, ring_head, ring_tail, (int)min(32u, ring_tail - ring_head), rb, rbb, ev, nv, report);
}

static int cull_tile(int n, int S, size_t* per_traj_out) {
  int WT = 16;
  // per trajectory of a warp tile: knots[n+1], dt, thresholds[n], far word, one candidate word per sample
  const size_t per_traj = sizeof(double) * (size_t)(n + 2) + sizeof(int) * ((size_t)n + 1 + (size_t)S);
  while (WT > 1 && CULL_WARPS * WT * per_traj > 40 * 1024) WT /= 2;
  *per_traj_out = per_traj;
  return WT;
}

static size_t cull_mesh_bytes(int K, const mst_mesh* robot, const mst_mesh* env) {
  return robot->layout.bytes + env->layout.bytes + (K == 3 ? sizeof(double) * collide_table_doubles(env->T, robot->V) : 0);
}

// do the sizes suit the culling kernel?  (asked before the solver is told to compute the far bits)
bool sample_collide_cull_suits(int n, int K, int S, const mst_mesh* robot, const mst_mesh* env) {
  if ((K != 3 && K != 4) || n > 32 || S > 65535 || S < 1) return false;
  if (robot->V > COLLIDE_MAX_V || robot->T > COLLIDE_MAX_TR || env->T >= (1 << 20)) return false;
  if (cull_mesh_bytes(K, robot, env) > MST_STAGE_LIMIT) return false;
  size_t per_traj;
  const int WT = cull_tile(n, S, &per_traj);
  return cull_mesh_bytes(K, robot, env) + CULL_WARPS * WT * per_traj <= 56 * 1024;
}

// MST_ERR_TOO_LARGE when the sizes do not suit this kernel (the caller runs sample_collide_kernel).
// tile_counter (device int, may be null: fixed stride over the tiles): scratch for the tile tickets.
int launch_sample_collide_cull(const double* coef, const double* dur, const unsigned* far_mask, int B, int n, int K,
                               int S, const mst_mesh* robot, const mst_mesh* env, uint8_t* hit, uint8_t* any_hit,
                               int* tile_counter, cudaStream_t stream) {
  if (B == 0) return MST_OK;
  if (!sample_collide_cull_suits(n, K, S, robot, env)) return MST_ERR_TOO_LARGE;
  const size_t mesh_bytes = cull_mesh_bytes(K, robot, env);
  size_t per_traj;
  const int WT = cull_tile(n, S, &per_traj);
  const size_t smem = mesh_bytes + CULL_WARPS * WT * per_traj;
  auto kern = K == 3 ? sample_collide_cull_kernel<3> : sample_collide_cull_kernel<4>;
  {
    const int rc = allow_dynamic_smem((const void*)kern, smem);
    if (rc != MST_OK) return rc;
  }
  const int tiles = (B + WT - 1) / WT;
  int blocks = (tiles + CULL_WARPS - 1) / CULL_WARPS;
  static const int per_sm_env = getenv("MST_CULL_CTAS_PER_SM") ? atoi(getenv("MST_CULL_CTAS_PER_SM")) : 0;   // A/B
  const int cap = MST_SM_COUNT * (per_sm_env > 0 ? per_sm_env : 4);  // 4 resident per SM; warps stride over the tiles
  if (blocks > cap) blocks = cap;
  static const bool fixed_stride = getenv("MST_CULL_FIXED_STRIDE") != nullptr;   // A/B
  if (fixed_stride) tile_counter = nullptr;
  if (tile_counter != nullptr) {
    cudaError_t e = cudaMemsetAsync(tile_counter, 0, sizeof(int), stream);
    if (e != cudaSuccess) { note_cuda_error(e); return MST_ERR_CUDA; }
  }
  kern<<<blocks, CULL_THREADS, smem, stream>>>(coef, dur, far_mask, B, n, S, WT, robot->d_image, robot->layout,
                                               robot->bounds, env->d_image, env->layout, env->bounds, hit, any_hit,
                                               tile_counter);
  return check_launch();
}

}  // namespace mst
