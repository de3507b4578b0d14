// Single-pass pipeline: solve -> sample -> pose -> collide -> any-hit in ONE persistent kernel.
//
// What it replaces in the reference: calculate_trajectory4D
// (src/optimizations/calculatingTrajectories.py:200-213) -> PiecewisePolynomial.eval at S sample
// times (src/optimizations/uav_trajectory.py:154-169) -> Fcl_checker.check_collision at every
// sampled pose (src/RigidBodyPlanners/fcl_checker.py:93-100).
//
// Why one kernel: in the two-launch pipeline (solve_condensed.cu + pipeline_fused.cu) the solver
// wrote 1,920 B of coefficients per trajectory and the second kernel read them back from HBM
// (1.82x the compulsory traffic of the step, and its first Horner step of every piece waited on
// that read).  Here a WARP owns a tile of 32/(G*K) time groups from the time stamps to the flags:
//
//   1. lane per right-hand-side column: classification, block LDL^T factors (group's first lane),
//      forward sweep, back sweep — the arithmetic of condensed_core.cuh, exactly as
//      mst_solve_batch runs it — but the back sweep keeps only the knot states (velocity,
//      acceleration, jerk: 3 doubles per knot and column) in shared memory, in place of the
//      forward values;
//   2. per tile: running duration sums, exact first-sample thresholds and the piece-of-sample byte
//      table (the sampling semantics of PiecewisePolynomial.eval, bit for bit);
//   3. trajectory by trajectory: the 8 coefficients of every (piece, axis) are formed ONCE by one
//      lane each from the knot states (piece_coefficients), stored to HBM with 16-byte coalesced
//      stores (their only trip through HBM) and staged in a two-trajectory shared-memory buffer,
//      from which the lanes — 32 consecutive samples per step, as in the two-launch kernel, so the
//      collision batches stay coherent — evaluate the non-fused Horner form;
//   4. root-box cull per sample, near samples through the warp's pose ring into the cursor engine
//      of collide_core.cuh; flags collect in shared memory and leave as whole rows.
//
// Optional wire targets (multi-GPU): the float32 polynomial matrix of path_to_pol
// (scripts/drones_pols_generator.py:63-77) and the flags are ALSO stored through up to 8 base
// pointers — this rank's gather buffer and its NVLink peer mappings — so the all-gather happens
// tile by tile from inside the kernel (peer stores), overlapped with the arithmetic.
//
// Groups the condensed solver must not take (duration spread > 4, t[0] != 0, bad stamps) are
// appended to the device-side list; the pivoted banded-LU kernel and the list mode of
// sample_collide_kernel finish them (both exit at once when the list is empty).
#include <stdlib.h>
#include <string.h>

#include "collide_core.cuh"
#include "condensed_core.cuh"
#include "onepass.cuh"
#include "stage.cuh"

namespace mst {

constexpr int ONEPASS_MAX_WARPS = 11;
constexpr int ONEPASS_TREGS = 4, ONEPASS_WREGS = 12;  // register tile of the input pipeline, doubles per lane

template <int K>
__global__ void __launch_bounds__(32 * ONEPASS_MAX_WARPS, 1)
onepass_kernel(const double* __restrict__ wp, const double* __restrict__ tstamps, int groups, int n, int G, int S,
               double* __restrict__ coef, double* __restrict__ dur, int* __restrict__ info,
               uint8_t* __restrict__ hit, uint8_t* __restrict__ any_hit, int* __restrict__ list,
               int* __restrict__ counters, const void* __restrict__ robot_img, MeshLayout rl, MeshBounds rbb,
               const void* __restrict__ env_img, MeshLayout el, MeshBounds evb, OnepassLayout L, WireTargets wire) {
  constexpr int POSE = K == 3 ? 0 : 1;
  constexpr int NP = PoseDim<POSE>::N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  stage_meshes(smem_raw, robot_img, rl.bytes, env_img, el.bytes, &bar);
  const MeshView rb = mesh_view(smem_raw, rl);
  const MeshView ev = mesh_view(smem_raw + rl.bytes, el);
  double* nv = reinterpret_cast<double*>(smem_raw + rl.bytes + el.bytes);
  if (POSE == 0) build_plane_vertex_table(rb, ev, nv);
  __syncthreads();

  const int R = G * K, GPW = L.GPW, TPT = L.TPT;
  const int gl = lane / R, col = lane - gl * R;
  const int d = col / K, k = col - d * K;
  const bool lane_used = gl < GPW;
  unsigned char* wbase = smem_raw + L.shared_bytes + (size_t)warp * L.bytes;
  double* wrho = reinterpret_cast<double*>(wbase + L.off_rho);    // [n][GPW]   1 / T
  double* wfac = reinterpret_cast<double*>(wbase + L.off_fac);    // [6(n-1)][GPW], later cbuf | dt | thr
  double* wy = reinterpret_cast<double*>(wbase + L.off_y);        // [3(n-1)][32] forward values, then knot states
  double* wt = reinterpret_cast<double*>(wbase + L.off_t);        // [GPW][n+1] stamps, later running duration sums
  double* ww = reinterpret_cast<double*>(wbase + L.off_w);        // [TPT][n+1][K] waypoints
  double* cbuf = wfac;                                            // [2][n][K][8] staged coefficients
  double* dts = wfac + L.cbuf_doubles;                            // [GPW]
  int* thr = reinterpret_cast<int*>(dts + GPW);                   // [GPW][n]
  uint8_t* hitb = wbase + L.off_hit;                              // [TPT][S] flags of the tile
  uint8_t* anyb = wbase + L.off_any;                              // [TPT]
  uint8_t* piece_of = wbase + L.off_piece;                        // [GPW][S]
  double* wT = reinterpret_cast<double*>(wbase + L.off_wire);     // wire mode: [GPW][n] durations T_i, then
  float* wstage = reinterpret_cast<float*>(wT + (size_t)GPW * n); //   [n][1 + 8K] one trajectory's float32 matrix
  PoseRing<NP>& ring = *reinterpret_cast<PoseRing<NP>*>(wbase + L.off_ring);
  unsigned ring_head = 0u, ring_tail = 0u;  // warp-uniform
  const bool wire_mat = wire.count > 0 && wire.mat[0] != nullptr;

  auto report = [&](int q, int s, bool h) {
    hitb[q * S + s] = h ? 1 : 0;
    if (h) anyb[q] = 1;
  };

  const long long sets = ((long long)groups + GPW - 1) / GPW;
  // sets are handed out by a ticket counter: a warp that meets the obstacle takes several times
  // longer over a tile than one in free space, and a static stride left the last warps running alone
  auto ticket = [&]() -> long long {
    int tk = 0;
    if (lane == 0) tk = atomicAdd(counters + 1, 1);
    return (long long)__shfl_sync(FULL, tk, 0);
  };
  // software pipeline of the input copies (as in condensed_cols_kernel): the NEXT set's stamps and
  // waypoints are loaded into registers before this set is worked on
  const bool piped = GPW * (n + 1) <= 32 * ONEPASS_TREGS && GPW * (n + 1) * R <= 32 * ONEPASS_WREGS;
  double tr[ONEPASS_TREGS], wr[ONEPASS_WREGS];
  auto load_set = [&](long long set2) {
    if (set2 >= sets) return;
    const long long h0 = set2 * GPW;
    const int c2 = (int)min((long long)GPW, groups - h0);
    const double* tb = tstamps + (size_t)h0 * (n + 1);
    const double* wb = wp + (size_t)h0 * (n + 1) * R;
#pragma unroll
    for (int j = 0; j < ONEPASS_TREGS; ++j) if (lane + 32 * j < c2 * (n + 1)) tr[j] = __ldg(tb + lane + 32 * j);
#pragma unroll
    for (int j = 0; j < ONEPASS_WREGS; ++j) if (lane + 32 * j < c2 * (n + 1) * R) wr[j] = __ldg(wb + lane + 32 * j);
  };

  long long set = ticket();
  if (piped) load_set(set);
  while (set < sets) {
    const long long next_set = ticket();
    const long long g0 = set * GPW;
    const int cnt = (int)min((long long)GPW, groups - g0);
    const int nbt = cnt * G;                 // trajectories of this tile
    const long long b0 = g0 * G;             // first trajectory of the tile
    __syncwarp();
    if (piped) {
#pragma unroll
      for (int j = 0; j < ONEPASS_TREGS; ++j) if (lane + 32 * j < cnt * (n + 1)) wt[lane + 32 * j] = tr[j];
#pragma unroll
      for (int j = 0; j < ONEPASS_WREGS; ++j) if (lane + 32 * j < cnt * (n + 1) * R) ww[lane + 32 * j] = wr[j];
    } else {
      for (int i = lane; i < cnt * (n + 1); i += 32) wt[i] = tstamps[(size_t)g0 * (n + 1) + i];
      for (int i = lane; i < cnt * (n + 1) * R; i += 32) ww[i] = wp[(size_t)g0 * (n + 1) * R + i];
    }
    __syncwarp();
    if (piped) load_set(next_set);

    // ---- 1. solve: classification and factorisation by the group's first column ----------------
    const bool mine = lane_used && gl < cnt;
    const double* tg = wt + (size_t)gl * (n + 1);
    int cls = 1;
    if (mine && col == 0) {
      double Tmin, Tmax;
      cls = classify_times(tg, n, &Tmin, &Tmax);
      if (cls != 0) {
        list[atomicAdd(counters, 1)] = (int)(g0 + gl);   // pivoted solver + list-mode sampling finish it
      } else {
        double* rho = wrho + gl;
        for (int i = 0; i < n; ++i) rho[(size_t)i * GPW] = tg[i + 1] - tg[i];
        condensed_factor(n, rho, wfac + gl, GPW);
      }
    }
    cls = __shfl_sync(FULL, cls, lane_used ? gl * R : 0);
    const unsigned okl = __ballot_sync(FULL, mine && col == 0 && cls == 0);  // bit gl * R per solvable group
    __syncwarp();
    if (mine && cls == 0) {
      const double* wcol = ww + ((size_t)gl * G + d) * (n + 1) * K + k;
      condensed_forward<1>(wcol, K, n, 1, wrho + gl, wfac + gl, GPW, wy + lane, 32);
      condensed_backward_states<1>(n, 1, wrho + gl, wfac + gl, GPW, wy + lane, 32);
    }
    __syncwarp();
    if (okl == 0u) { set = next_set; continue; }
    auto group_ok = [&](int g) -> bool { return (okl >> (g * R)) & 1u; };

    // ---- 2. per-tile tables ----------------------------------------------------------------------
    // durations and status out (every trajectory of a group carries its own copy of the durations)
    for (int item = lane; item < nbt * n; item += 32) {
      const int q = item / n, i = item - q * n, g = q / G;
      if (group_ok(g)) {
        const double Ti = wt[g * (n + 1) + i + 1] - wt[g * (n + 1) + i];
        dur[(size_t)(b0 + q) * n + i] = Ti;
        if (wire_mat && q == g * G) wT[g * n + i] = Ti;
      }
    }
    if (lane < nbt) {
      anyb[lane] = 0;
      if (group_ok(lane / G)) info[b0 + lane] = MST_INFO_OK;
    }
    __syncwarp();
    // running sums of the durations in place of the stamps (left-to-right, as
    // PiecewisePolynomial.eval accumulates t_counting) and the sample spacing
    if (lane < cnt && group_ok(lane)) {
      double* kn = wt + lane * (n + 1);
      double prev = kn[0], acc = 0.0;
      kn[0] = 0.0;
#pragma unroll 1
      for (int i = 0; i < n; ++i) {
        const double nx = kn[i + 1];
        acc = __dadd_rn(acc, nx - prev);
        prev = nx;
        kn[i + 1] = acc;
      }
      dts[lane] = __ddiv_rn(acc, (double)S);
    }
    __syncwarp();
    // thresholds: first s with !(s * dt < knot), found from the quotient and corrected with the very
    // comparison PiecewisePolynomial.eval makes (t is non-decreasing in s)
#pragma unroll 1
    for (int item = lane; item < cnt * n; item += 32) {
      const int g = item / n, i = item - g * n;
      int first = S;
      if (i < n - 1 && group_ok(g)) {
        const double knot = wt[g * (n + 1) + i + 1], dt = dts[g];
        first = (int)fmin(fmax(ceil(__ddiv_rn(knot, dt)), 0.0), (double)S);
        while (first > 0 && !(__dmul_rn((double)(first - 1), dt) < knot)) --first;
        while (first < S && __dmul_rn((double)first, dt) < knot) ++first;
      }
      thr[item] = first;  // thr[g][n-1] = S closes the last piece
    }
    __syncwarp();
#pragma unroll 1
    for (int item = lane; item < cnt * n; item += 32) {
      const int g = item / n, i = item - g * n;
      if (group_ok(g)) {
        const int from = i ? thr[item - 1] : 0, to = thr[item];
#pragma unroll 1
        for (int x = from; x < to; ++x) piece_of[g * S + x] = (uint8_t)i;
      }
    }
    __syncwarp();

    // ---- 3. coefficients of trajectory q: formed once, stored once, staged for its samples -----------
    auto stage_trajectory = [&](int q) {
      const int g = q / G;
      if (!group_ok(g)) return;
      double* cb = cbuf + (size_t)(q & 1) * n * K * MST_NCOEF;
      double* cd = coef + (size_t)(b0 + q) * n * K * MST_NCOEF;
#pragma unroll 1
      for (int item = lane; item < n * K; item += 32) {
        const int i = item / K, kk = item - i * K;
        const int c = q * K + kk;   // the lane that solved this column
        double v0 = 0.0, a0 = 0.0, j0 = 0.0, v1 = 0.0, a1 = 0.0, j1 = 0.0;
        if (i >= 1) { const double* x = wy + (size_t)(i - 1) * 3 * 32 + c; v0 = x[0]; a0 = x[32]; j0 = x[64]; }
        if (i + 1 < n) { const double* x = wy + (size_t)i * 3 * 32 + c; v1 = x[0]; a1 = x[32]; j1 = x[64]; }
        const double w0 = ww[((size_t)q * (n + 1) + i) * K + kk], w1 = ww[((size_t)q * (n + 1) + i + 1) * K + kk];
        double cc[MST_NCOEF];
        piece_coefficients(w0, w1 - w0, v0, a0, j0, v1, a1, j1, wrho[(size_t)i * GPW + g], cc);
        double2* s2 = reinterpret_cast<double2*>(cb + (size_t)item * MST_NCOEF);
        double2* g2 = reinterpret_cast<double2*>(cd + (size_t)item * MST_NCOEF);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const double2 v = make_double2(cc[2 * e], cc[2 * e + 1]);
          s2[e] = v;
          g2[e] = v;
        }
        if (wire_mat) {
          float* ws = wstage + i * (1 + MST_NCOEF * K) + 1 + MST_NCOEF * kk;
#pragma unroll
          for (int e = 0; e < MST_NCOEF; ++e) ws[e] = (float)cc[e];
          if (kk == 0) wstage[i * (1 + MST_NCOEF * K)] = (float)wT[g * n + i];
        }
      }
      if (wire_mat) {
        __syncwarp();
        const int words = n * (1 + MST_NCOEF * K);
        const size_t row = (size_t)(wire.row0 + b0 + q) * words;
        for (int t = 0; t < wire.count; ++t) {
          float* dst = wire.mat[t] + row;
          if (wire.mat[t] != nullptr)
            for (int w = lane; w < words; w += 32) dst[w] = wstage[w];
        }
      }
    };

    // ---- 4. sampling and collision -------------------------------------------------------------------
    const int work = nbt * S;
    int tl = 0, s = lane, dd = 0, gq = 0;   // (trajectory, sample) of this lane; drone and group of tl
    int next_stage = 0;
    for (int base = 0; base < work; base += 32, s += 32) {
      // trajectory q is staged when the sampling front has left trajectory q - 2 (whose buffer it takes)
      while (next_stage < nbt && (next_stage < 2 || base >= (next_stage - 1) * S)) {
        __syncwarp();
        stage_trajectory(next_stage);
        ++next_stage;
        __syncwarp();
      }
      while (s >= S) { s -= S; ++tl; if (++dd == G) { dd = 0; ++gq; } }
      const int idx = base + lane;
      bool active = idx < work;
      if (!active) { tl = nbt - 1; s = S - 1; gq = (nbt - 1) / G; }   // parked on the tile's last sample (not reported)
      active = active && group_ok(gq);
      const double t = __dmul_rn((double)s, dts[gq]);
      const int piece = min((int)piece_of[gq * S + s], n - 1);
      const double local = __dsub_rn(t, wt[gq * (n + 1) + piece]);
      const double* cp = cbuf + (((size_t)(tl & 1) * n + piece) * K) * MST_NCOEF;
      double pos[K];
#pragma unroll
      for (int a = 0; a < K; ++a) {
        const double2* src = reinterpret_cast<const double2*>(cp + a * MST_NCOEF);
        const double2 c01 = src[0], c23 = src[1], c45 = src[2], c67 = src[3];
        double x = c67.y;  // 0*t + c7
        x = __dadd_rn(__dmul_rn(x, local), c67.x);
        x = __dadd_rn(__dmul_rn(x, local), c45.y);
        x = __dadd_rn(__dmul_rn(x, local), c45.x);
        x = __dadd_rn(__dmul_rn(x, local), c23.y);
        x = __dadd_rn(__dmul_rn(x, local), c23.x);
        x = __dadd_rn(__dmul_rn(x, local), c01.y);
        x = __dadd_rn(__dmul_rn(x, local), c01.x);
        pos[a] = x;
      }
      double pp[NP];
      pp[0] = pos[0]; pp[1] = pos[1]; pp[2] = pos[2];
      if (POSE == 1) sincos(pos[K - 1] * 0.5, &pp[3], &pp[4]);
      const bool near = active && pose_near_environment<POSE>(pp, rbb, evb);
      if (active && !near) hitb[tl * S + s] = 0;
      ring_push<POSE>(ring, ring_tail, near, pp, tl, s, -1, 0u, 0u);
      while (ring_tail - ring_head >= 32u) ring_drain<POSE>(ring, ring_head, ring_tail, 32, rb, rbb, ev, nv, report);
    }
    // the tile's flags leave as whole rows: every queued pose is decided first
    while (ring_tail != ring_head)
      ring_drain<POSE>(ring, ring_head, ring_tail, (int)min(32u, ring_tail - ring_head), rb, rbb, ev, nv, report);
    __syncwarp();
    {
      const int targets = wire.count > 0 ? wire.count : 0;
      for (int t = -1; t < targets; ++t) {
        uint8_t* hdst = t < 0 ? hit : wire.hit[t];
        uint8_t* adst = t < 0 ? any_hit : wire.any[t];
        const size_t off = (size_t)(t < 0 ? 0 : wire.row0) + (size_t)b0;
        if (hdst == nullptr) continue;
        if ((S & 3) == 0 && ((reinterpret_cast<uintptr_t>(hdst) + off * S) & 3) == 0) {
          const unsigned* src = reinterpret_cast<const unsigned*>(hitb);
          unsigned* dst = reinterpret_cast<unsigned*>(hdst + off * S);
          const int wpr = S >> 2;   // words per row
          for (int w = lane; w < nbt * wpr; w += 32)
            if (group_ok((w / wpr) / G)) dst[w] = src[w];
        } else {
          for (int w = lane; w < nbt * S; w += 32)
            if (group_ok((w / S) / G)) hdst[off * S + w] = hitb[w];
        }
        if (lane < nbt && group_ok(lane / G)) adst[off + lane] = anyb[lane];
      }
    }
    set = next_set;
  }
}

// Wire outputs of the time groups the single-pass kernel handed to the pivoted solver: packed /
// copied from the local results once the list-mode kernels have produced them.  One CTA per
// trajectory; exits at once when the list is empty.
__global__ void __launch_bounds__(128)
wire_patch_kernel(const double* __restrict__ coef, const double* __restrict__ dur, const uint8_t* __restrict__ hit,
                  const uint8_t* __restrict__ any_hit, int n, int K, int G, int S, const int* __restrict__ list,
                  const int* __restrict__ list_count, WireTargets wire) {
  const int listed = *list_count;
  if (listed == 0) return;
  const int width = 1 + MST_NCOEF * K, words = n * width;
  for (long long j = blockIdx.x; j < (long long)listed * G; j += gridDim.x) {
    const size_t b = (size_t)list[j / G] * G + (size_t)(j % G);
    for (int t = 0; t < wire.count; ++t) {
      if (wire.mat[t] != nullptr) {
        float* dst = wire.mat[t] + (size_t)(wire.row0 + b) * words;
        for (int w = threadIdx.x; w < words; w += blockDim.x) {
          const int row = w / width, c = w - row * width;
          dst[w] = (float)(c == 0 ? dur[b * n + row] : coef[(b * n + row) * (width - 1) + c - 1]);
        }
      }
      if (wire.hit[t] != nullptr) {
        for (int s = threadIdx.x; s < S; s += blockDim.x) wire.hit[t][(size_t)(wire.row0 + b) * S + s] = hit[b * S + s];
        if (threadIdx.x == 0) wire.any[t][wire.row0 + b] = any_hit[b];
      }
    }
  }
}

int launch_wire_patch(const double* coef, const double* dur, const uint8_t* hit, const uint8_t* any_hit, int n, int K,
                      int G, int S, const int* list, const int* list_count, const WireTargets* wire,
                      cudaStream_t stream) {
  wire_patch_kernel<<<(unsigned)(sm_count() * 4), 128, 0, stream>>>(coef, dur, hit, any_hit, n, K, G, S, list,
                                                                   list_count, *wire);
  return check_launch();
}

// shared-memory plan of one warp (bytes from the warp's base) — host and device agree through the struct
OnepassLayout onepass_layout(int n, int K, int G, int S, bool wire, size_t shared_bytes) {
  OnepassLayout L;
  const int R = G * K;
  L.GPW = 32 / R;
  L.TPT = L.GPW * G;
  const int NP = K == 3 ? 3 : 5;
  const size_t ring_bytes = K == 3 ? sizeof(PoseRing<3>) : sizeof(PoseRing<5>);
  (void)NP;
  L.cbuf_doubles = 2 * (size_t)n * K * MST_NCOEF;
  const size_t fac_doubles = (size_t)L.GPW * 6 * (n - 1);
  const size_t alias_doubles = L.cbuf_doubles + L.GPW + ((size_t)L.GPW * n + 1) / 2;
  size_t o = 0;
  L.off_rho = o; o += sizeof(double) * (size_t)L.GPW * n;
  L.off_fac = o; o += sizeof(double) * (fac_doubles > alias_doubles ? fac_doubles : alias_doubles);
  L.off_y = o;   o += sizeof(double) * 32 * 3 * (size_t)(n - 1);
  L.off_t = o;   o += sizeof(double) * (size_t)L.GPW * (n + 1);
  L.off_w = o;   o += sizeof(double) * (size_t)L.TPT * (n + 1) * K;
  o = (o + 15) & ~(size_t)15;
  L.off_hit = o; o += ((size_t)L.TPT * S + 15) & ~(size_t)15;
  L.off_any = o; o += ((size_t)L.TPT + 15) & ~(size_t)15;
  L.off_piece = o; o += ((size_t)L.GPW * S + 15) & ~(size_t)15;
  L.off_wire = o;
  if (wire) o += (sizeof(double) * (size_t)L.GPW * n + sizeof(float) * (size_t)n * (1 + MST_NCOEF * K) + 15) & ~(size_t)15;
  L.off_ring = o; o += (ring_bytes + 15) & ~(size_t)15;
  L.bytes = o;
  L.shared_bytes = (shared_bytes + 15) & ~(size_t)15;
  return L;
}

// MST_ERR_TOO_LARGE: the sizes do not suit this kernel — the caller runs the two-launch pipeline
int launch_onepass(const double* wp, const double* t, int groups, int n, int K, int G, int S, double* coef, double* dur,
                   int* info, uint8_t* hit, uint8_t* any_hit, int* list, int* counters, const mst_mesh* robot,
                   const mst_mesh* env, const WireTargets* wire, cudaStream_t stream) {
  if ((K != 3 && K != 4) || G * K > 32 || n < 2 || n > 32 || S < 32 || S > 4096) return MST_ERR_TOO_LARGE;
  if (robot->V > COLLIDE_MAX_V || robot->T > COLLIDE_MAX_TR || env->T >= (1 << 20)) return MST_ERR_TOO_LARGE;
  WireTargets none;
  memset(&none, 0, sizeof(none));
  const WireTargets& wt = wire ? *wire : none;
  const size_t shared = robot->layout.bytes + env->layout.bytes +
                        (K == 3 ? sizeof(double) * collide_table_doubles(env->T, robot->V) : 0);
  const OnepassLayout L = onepass_layout(n, K, G, S, wt.count > 0, shared);
  const size_t budget = MST_MAX_SMEM - 1024;  // static shared memory of the kernel and alignment slack
  if (L.shared_bytes + 4 * L.bytes > budget) return MST_ERR_TOO_LARGE;   // fewer than 4 warps per SM: not worth it
  int warps = (int)((budget - L.shared_bytes) / L.bytes);
  if (warps > ONEPASS_MAX_WARPS) warps = ONEPASS_MAX_WARPS;
  static const int warps_env = getenv("MST_ONEPASS_WARPS") ? atoi(getenv("MST_ONEPASS_WARPS")) : 0;
  if (warps_env > 0 && warps_env < warps) warps = warps_env;
  const size_t smem = L.shared_bytes + (size_t)warps * L.bytes;
  auto kern = K == 3 ? onepass_kernel<3> : onepass_kernel<4>;
  {
    const int rc = allow_dynamic_smem((const void*)kern, smem);
    if (rc != MST_OK) return rc;
  }
  cudaError_t e = cudaMemsetAsync(counters, 0, 2 * sizeof(int), stream);
  if (e != cudaSuccess) { note_cuda_error(e); return MST_ERR_CUDA; }
  const long long sets = ((long long)groups + L.GPW - 1) / L.GPW;
  long long blocks = (sets + warps - 1) / warps;
  if (blocks > sm_count()) blocks = sm_count();   // persistent: one CTA per SM, warps draw tickets
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, 32 * warps, smem, stream>>>(wp, t, groups, n, G, S, coef, dur, info, hit, any_hit, list,
                                                       counters, robot->d_image, robot->layout, robot->bounds,
                                                       env->d_image, env->layout, env->bounds, L, wt);
  return check_launch();
}

}  // namespace mst
