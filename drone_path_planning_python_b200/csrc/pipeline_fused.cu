// Fused sample -> pose -> collide -> any-hit kernel of the pipeline.
//
// What it replaces in the reference: PiecewisePolynomial.eval at S sample times
// (src/optimizations/uav_trajectory.py:154-169, sampled the way
// src/trajectory_visualising/visualization.py:53 samples), the robot pose at each sample
// (isStateValid's pos / yaw quaternion, src/RigidBodyPlanners/RB_planning_sep_coll_check.py:
// 208-215) and Fcl_checker.check_collision (src/RigidBodyPlanners/fcl_checker.py:93-100).
//
// Data flow: coefficients and durations come straight from the solver kernels' output
// (pulled into L2 a few trajectories ahead of their use), both meshes
// are staged once per CTA into shared memory with one bulk (TMA) copy each, sampled
// positions never leave registers; only hit[B][S] (1 byte per sample, coalesced) and
// any_hit[B] are written.
//
// Mapping: a WARP walks tiles of (up to) 16 trajectories; inside a tile the (trajectory, sample)
// pairs are flattened over the lanes, so a warp holds 32 CONSECUTIVE samples of one
// trajectory (neighbouring poses: coefficient loads are warp-wide broadcasts, the broad
// phase decisions mostly agree) and the narrow phase is compacted across the warp.
//
// The evaluation is bit-identical to mst_sample_batch (same running-sum piece search, same
// non-fused Horner), so pipeline flags equal "sample, then collide" exactly.
#include "collide_core.cuh"
#include <stdlib.h>

#include "stage.cuh"

namespace mst {

constexpr int FUSED_THREADS = 128;
constexpr int FUSED_WARPS = FUSED_THREADS / 32;
constexpr int FUSED_PF = 4;      // trajectories between the L2 prefetch and the sampling front
constexpr int FUSED_WT_MAX = 32; // trajectories per warp tile (chosen by the launcher, <= 32)

// Every WARP walks its own tiles of FUSED_WT trajectories (no CTA-wide barrier after the
// meshes are staged): a warp that meets the obstacle takes several times longer over a
// tile than one that flies in free space, and a CTA barrier per tile made the fast warps
// wait (27 % of all warp-stall samples in profiles/r1_collision_history.md).
//
// Sampling and collision are decoupled inside the warp: every lane evaluates its sample and
// applies the root-box cull; samples that fail it get hit = 0 right away, the others are
// appended (ballot-compacted) to the warp's pose ring, and only when 32 of them are waiting
// does the warp run the collision engine (collide_core.cuh) — on a DENSE batch, whatever mix
// of near and far samples the trajectories produce.
// LIST = true: only the trajectories of the time groups in list[0 .. *list_count) are worked on
// (trajectory j of the launch is list[j / G] * G + j % G) — the groups the single-pass pipeline
// handed to the pivoted solver; the kernel exits at once when the list is empty.
// GLOBAL = true: mesh images read in place from device memory (too large to stage).
// With both false the code is the round-1 kernel unchanged (every LIST / GLOBAL branch folds away).
template <int K, bool TAB, bool LIST, bool GLOBAL>
__global__ void __launch_bounds__(FUSED_THREADS, 4)
sample_collide_kernel(const double* __restrict__ coef, const double* __restrict__ dur, int B, int n, int S, int FUSED_WT,
                      const void* __restrict__ robot_img, MeshLayout rl, MeshBounds rbb,
                      const void* __restrict__ env_img, MeshLayout el, MeshBounds evb,
                      uint8_t* __restrict__ hit, uint8_t* __restrict__ any_hit, const int* __restrict__ list,
                      const int* __restrict__ list_count, int G) {
  if (LIST) {
    const int listed = *list_count;
    if (listed == 0) return;
    B = listed * G;
  }
  // launch-local trajectory index -> trajectory of the batch
  auto bmap = [&](int j) -> size_t { return LIST ? (size_t)list[j / G] * G + (size_t)(j % G) : (size_t)j; };
  constexpr int POSE = K == 3 ? 0 : 1;
  constexpr int NP = PoseDim<POSE>::N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ PoseRing<NP> rings[FUSED_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PoseRing<NP>& ring = rings[warp];
  if (!GLOBAL) stage_meshes(smem_raw, robot_img, rl.bytes, env_img, el.bytes, &bar);
  const MeshView rb = GLOBAL ? mesh_view(robot_img, rl) : mesh_view(smem_raw, rl);
  const MeshView ev = GLOBAL ? mesh_view(env_img, el) : mesh_view(smem_raw + rl.bytes, el);
  const bool engine = collide_engine_supports(rb, ev);
  // behind the meshes: the plane x vertex table, then per-warp tables knots[WT][n+1]
  // (running sums of the durations) and dt[WT]
  double* nv = GLOBAL ? nullptr : reinterpret_cast<double*>(smem_raw + rl.bytes + el.bytes);
  if (!GLOBAL && engine && POSE == 0) build_plane_vertex_table(rb, ev, nv);
  __syncthreads();
  double* tables = GLOBAL ? reinterpret_cast<double*>(smem_raw)
                          : nv + ((engine && POSE == 0) ? collide_table_doubles(ev.T, rb.V) : 0);
  double* knots = tables + warp * FUSED_WT * (n + 2);
  double* dts = knots + FUSED_WT * (n + 1);
  // thr[WT][n]: first sample index of pieces 1 .. n-1, and (when the launcher found room: TAB)
  // piece_of[WT][S]: the piece of every sample as a byte, filled range by range from thr — the
  // per-sample search for the piece (a bisection over the knots with dependent shared-memory
  // loads, 12 % of the stall samples) becomes one byte load
  int* thr = reinterpret_cast<int*>(tables + FUSED_WARPS * FUSED_WT * (n + 2)) + warp * FUSED_WT * n;
  uint8_t* piece_of = reinterpret_cast<uint8_t*>(reinterpret_cast<int*>(tables + FUSED_WARPS * FUSED_WT * (n + 2)) +
                                                 FUSED_WARPS * FUSED_WT * n) + warp * FUSED_WT * S;
  const int traj_bytes = n * K * MST_NCOEF * (int)sizeof(double);
  unsigned ring_head = 0u, ring_tail = 0u;  // warp-uniform

  // zeroed at the start of the trajectory's tile; every writer of any_hit stores 1
  auto report = [&](int b, int s, bool h) {
    hit[(size_t)b * S + s] = h ? 1 : 0;
    if (h) any_hit[b] = 1;
  };

  const int tiles = (B + FUSED_WT - 1) / FUSED_WT;
  for (int tile = blockIdx.x * FUSED_WARPS + warp; tile < tiles; tile += gridDim.x * FUSED_WARPS) {
    const int b0 = tile * FUSED_WT;
    const int nb = min(FUSED_WT, B - b0);
    // Coefficients are pulled into L2 FUSED_PF trajectories ahead of the one being sampled (see
    // prefetch_trajectory below); the next tile's durations, which the table set-up of that tile
    // reads all at once, are pulled here.  The batch is far larger than L2, so all of it comes
    // from HBM, and the first Horner step of a new piece was the top stall of the kernel.
    const long long next_tile = (long long)tile + (long long)gridDim.x * FUSED_WARPS;
    const int next_b0 = next_tile < tiles ? (int)(next_tile * FUSED_WT) : B;
    const int next_nb = min(FUSED_WT, B - next_b0);
    {
      const char* dbase = reinterpret_cast<const char*>(dur + (size_t)next_b0 * n);
      const size_t dbytes = LIST ? 0 : (size_t)max(next_nb, 0) * n * sizeof(double);
#pragma unroll 1
      for (size_t off = (size_t)lane * 128; off < dbytes; off += 32 * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(dbase + off));
    }
    // q-th trajectory of this warp counted from the start of the tile; runs on into the next tile
    auto prefetch_trajectory = [&](int q) {
      const int b = q < nb ? b0 + q : (q - nb < next_nb ? next_b0 + (q - nb) : -1);
      if (b >= 0) {
        const char* cbase = reinterpret_cast<const char*>(coef) + bmap(b) * traj_bytes;
#pragma unroll 1
        for (int off = lane * 128; off < traj_bytes; off += 32 * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(cbase + off));
      }
    };
    __syncwarp();  // previous tile's tables are no longer read
    if (lane < nb) {
      const double* T = dur + bmap(b0 + lane) * n;
      double* kn = knots + lane * (n + 1);
      double acc = 0.0;
      kn[0] = 0.0;
#pragma unroll 1
      for (int i = 0; i < n; ++i) { acc = __dadd_rn(acc, T[i]); kn[i + 1] = acc; }
      dts[lane] = __ddiv_rn(acc, (double)S);
      any_hit[bmap(b0 + lane)] = 0;
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
    if (TAB) {
#pragma unroll 1
      for (int item = lane; item < nb * n; item += 32) {
        const int q = item / n, i = item - q * n;
        const int from = i ? thr[item - 1] : 0, to = thr[item];
#pragma unroll 1
        for (int x = from; x < to; ++x) piece_of[q * S + x] = (uint8_t)i;
      }
      __syncwarp();
    }
    const int work = nb * S;
    // warp-uniform trip count: every lane stays in the loop (ballots), lanes past the end of
    // the tile are simply inactive
    // (trajectory, sample) of this lane, advanced by 32 samples per iteration without dividing
    int tl = 0, s = lane;
    int pf_tl = 0, pf_base = 0;  // warp-uniform: trajectory whose first sample is at pf_base
    if (tile < (int)(gridDim.x * FUSED_WARPS))  // first tile of the warp: nothing was pulled ahead of it
#pragma unroll 1
      for (int q = 0; q < FUSED_PF; ++q) prefetch_trajectory(q);
    for (int base = 0; base < work; base += 32, s += 32) {
      const int idx = base + lane;
      const bool active = idx < work;
      while (base >= pf_base) { prefetch_trajectory(pf_tl + FUSED_PF); ++pf_tl; pf_base += S; }
      while (s >= S) { s -= S; ++tl; }
      if (!active) { tl = nb - 1; s = S - 1; }  // parked on the tile's last sample (not reported)
      const double* kn = knots + tl * (n + 1);
      const double t = __dmul_rn((double)s, dts[tl]);
      // PiecewisePolynomial.eval: first piece with t < acc + T_i, else the last one at
      // t - sum(T[:-1]) = the number of thresholds <= s.  (Decreasing stamps — reported in info —
      // can leave table entries unwritten: the clamp keeps the coefficient read inside the batch.)
      int piece = 0;
      if (TAB) {
        piece = min((int)piece_of[tl * S + s], n - 1);
      } else {
        const int* ti = thr + tl * n;
        int last = n - 1;
        while (piece < last) {
          const int mid = (piece + last) >> 1;
          if (s < ti[mid]) last = mid; else piece = mid + 1;
        }
      }
      const double local = __dsub_rn(t, kn[piece]);
      const size_t btl = bmap(b0 + tl);
      const double* cp = coef + ((btl * n + piece) * K) * MST_NCOEF;
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
      // yaw: only samples whose bounding sphere reaches the obstacle's box need the rotation (the others are
      // free whatever the heading); the decision is the one pose_near_environment makes first anyway
      bool reach = true;
      if (POSE == 1) {
        reach = sphere_near_environment(pp, rbb, evb);
        pp[3] = 0.0; pp[4] = 1.0;
        if (reach) sincos(pos[K - 1] * 0.5, &pp[3], &pp[4]);
      }
      if (!engine) {  // meshes the bit-mask cursors cannot hold: plain per-lane test over all pairs
        double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (POSE == 1) quat_to_matrix(0.0, 0.0, pp[3], pp[4], R);
        if (active) report((int)btl, s, robot_hits_env(R, pp, rb.tri, rb.T, ev.tri, ev.box, ev.T, evb.root, rbb.radius, true));
        continue;
      }
      const bool near = active && reach && pose_near_environment<POSE>(pp, rbb, evb);
      if (active && !near) {
        if (LIST) hit[btl * S + s] = 0; else hit[(size_t)b0 * S + idx] = 0;
      }
      ring_push<POSE>(ring, ring_tail, near, pp, LIST ? (int)btl : b0 + tl, s, -1, 0u, 0u);
      while (ring_tail - ring_head >= 32u) ring_drain<POSE, !GLOBAL>(ring, ring_head, ring_tail, 32, rb, rbb, ev, nv, report);
    }
  }
  while (ring_tail != ring_head)
    ring_drain<POSE, !GLOBAL>(ring, ring_head, ring_tail, (int)min(32u, ring_tail - ring_head), rb, rbb, ev, nv, report);
}

// list / list_count (device) non-null: list mode over the time groups of G trajectories named there;
// B is then the size of the whole batch (upper bound of the work)
int launch_sample_collide(const double* coef, const double* dur, int B, int n, int K, int S,
                          const mst_mesh* robot, const mst_mesh* env, uint8_t* hit, uint8_t* any_hit,
                          cudaStream_t stream, const int* list, const int* list_count, int G) {
  if (B == 0) return MST_OK;
  // trajectories per warp tile: large tiles amortise the per-tile set-up and leave fewer
  // half-empty last iterations (3.94 ms at 16 vs 4.13 ms at 4 per 1 M trajectories); bounded by
  // the knot tables' shared memory (about 48 kB per CTA keeps 4 CTAs per SM)
  static const int wt_env = getenv("MST_FUSED_WT") ? atoi(getenv("MST_FUSED_WT")) : 0;
  int FUSED_WT = wt_env > 0 ? wt_env : 16;
  if (FUSED_WT > FUSED_WT_MAX) FUSED_WT = FUSED_WT_MAX;
  // per trajectory: knots[n+1], dt, thresholds[n], and the piece-of-sample bytes when they are
  // small next to the rest
  const bool tab = n <= 255 && S <= 1024;
  const size_t per_traj = sizeof(double) * (size_t)(n + 2) + sizeof(int) * (size_t)n + (tab ? (size_t)S : 0);
  while (FUSED_WT > 1 && FUSED_WARPS * FUSED_WT * per_traj > 24 * 1024) FUSED_WT /= 2;
  // the plane x vertex table serves translation-only poses (K = 3) alone; it exists only when the
  // robot fits the cursor engine
  const bool engine = robot->V <= COLLIDE_MAX_V && robot->T <= COLLIDE_MAX_TR;
  size_t mesh_bytes = robot->layout.bytes + env->layout.bytes +
                      ((K == 3 && engine) ? sizeof(double) * collide_table_doubles(env->T, robot->V) : 0);
  const bool global = mesh_bytes > MST_STAGE_LIMIT;   // large meshes are read in place (block boxes cull)
  if (global) mesh_bytes = 0;
  const size_t smem = mesh_bytes + FUSED_WARPS * FUSED_WT * per_traj;
  void (*kern)(const double*, const double*, int, int, int, int, const void*, MeshLayout, MeshBounds, const void*,
               MeshLayout, MeshBounds, uint8_t*, uint8_t*, const int*, const int*, int);
#define MST_PICK(KK, TT, LL) (global ? sample_collide_kernel<KK, TT, LL, true> : sample_collide_kernel<KK, TT, LL, false>)
  if (list)
    kern = K == 3 ? (tab ? MST_PICK(3, true, true) : MST_PICK(3, false, true))
                  : (tab ? MST_PICK(4, true, true) : MST_PICK(4, false, true));
  else
    kern = K == 3 ? (tab ? MST_PICK(3, true, false) : MST_PICK(3, false, false))
                  : (tab ? MST_PICK(4, true, false) : MST_PICK(4, false, false));
#undef MST_PICK
  {
    const int rc = allow_dynamic_smem((const void*)kern, smem);
    if (rc != MST_OK) return rc;
  }
  const int tiles = (B + FUSED_WT - 1) / FUSED_WT;
  int blocks = (tiles + FUSED_WARPS - 1) / FUSED_WARPS;
  const int cap = MST_SM_COUNT * 4;  // persistent CTAs (4 resident per SM); warps stride over the tiles
  if (blocks > cap) blocks = cap;
  kern<<<blocks, FUSED_THREADS, smem, stream>>>(coef, dur, B, n, S, FUSED_WT, robot->d_image, robot->layout,
                                                 robot->bounds, env->d_image, env->layout, env->bounds, hit,
                                                 any_hit, list, list_count, G);
  return check_launch();
}

}  // namespace mst
