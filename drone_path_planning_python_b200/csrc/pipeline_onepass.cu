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

// Guided schedule of the sampling phase: the CTA's trajectory slots [0, total) are cut into chunks
// of 4, then 2, then 1 trajectories; warps draw chunk numbers from a shared-memory counter.  Long
// chunks keep the lanes full (a warp's 32 consecutive samples run across the trajectories of its
// chunk), single trajectories at the end keep the warps' finishing times within one trajectory.
__device__ __forceinline__ void chunk_range(int k, int total, int& lo, int& hi) {
  const int n4 = (total * 5 / 8) / 4, a = 4 * n4;
  const int n2 = ((total - a) * 5 / 8) / 2, b = a + 2 * n2;
  if (k < n4) { lo = 4 * k; hi = lo + 4; }
  else if (k < n4 + n2) { lo = a + 2 * (k - n4); hi = lo + 2; }
  else { lo = b + (k - n4 - n2); hi = lo + 1; }
  if (hi > total) hi = total;
}

// The kernel runs in CTA-wide ROUNDS of three phases, one persistent CTA per SM:
//   phase 1  every warp solves its own tile (32/(G*K) time groups) and builds the tile's tables;
//   phase 2  the trajectories of all the CTA's tiles are sampled and collision-checked, handed out
//            to the warps chunk by chunk (chunk_range) — any warp reads any tile's knot states;
//   phase 3  every warp stores its tile's flag rows.
// Why phases: the first version let every warp walk solve -> sample -> collide on its own.  Its hot
// code (~5,000 instructions, 80 kB) then lived in the instruction cache (32 kB L1.5 per SM) all at
// once, warps being in different places: 5.9 of every 10 issue-slot cycles waited for instructions
// (profiles/r2_onepass_history.md) and the step took 8.0 ms.  With the warps of an SM in the same
// phase the working set is the phase's code: ~20 kB (solve) or ~24 kB (sample + collide).
// Poses still waiting in a warp's ring at the end of a round stay there: their flag bytes were stored
// as 0 with the tile's rows and are overwritten in HBM when the pose turns out to collide.
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
  __shared__ int s_set0[2];                      // first set of this round / of the next one
  __shared__ int s_chunk;                        // chunk counter of the sampling phase
  __shared__ unsigned s_okl[ONEPASS_MAX_WARPS];  // per tile: bit g * R set when group g was solved here
  // per trajectory slot of a round: tile | trajectory in the tile << 8 | group << 16 | drone << 24 (fixed for
  // the launch: no integer divisions in the sampling loop), and whether the slot was solved in this round
  __shared__ unsigned s_slot[ONEPASS_MAX_WARPS * 16];
  __shared__ unsigned char s_ok[ONEPASS_MAX_WARPS * 16];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  stage_meshes(smem_raw, robot_img, rl.bytes, env_img, el.bytes, &bar);
  const MeshView rb = mesh_view(smem_raw, rl);
  const MeshView ev = mesh_view(smem_raw + rl.bytes, el);
  double* nv = reinterpret_cast<double*>(smem_raw + rl.bytes + el.bytes);
  if (POSE == 0) build_plane_vertex_table(rb, ev, nv);
  const int R = G * K, GPW = L.GPW, TPT = L.TPT;
  const long long sets = ((long long)groups + GPW - 1) / GPW;
  if (threadIdx.x == 0) s_set0[0] = atomicAdd(counters + 1, W);
  for (int slot = threadIdx.x; slot < W * TPT; slot += blockDim.x) {
    const int tw = slot / TPT, q = slot - tw * TPT, g = q / G;
    s_slot[slot] = (unsigned)tw | ((unsigned)q << 8) | ((unsigned)g << 16) | ((unsigned)(q - g * G) << 24);
  }
  const unsigned magic_n = 0xffffffffu / (unsigned)n + 1u;   // item / n == __umulhi(item, magic_n) for the small items here
  __syncthreads();

  const int gl = lane / R, col = lane - gl * R;
  const int d = col / K, k = col - d * K;
  const bool lane_used = gl < GPW;
  unsigned char* const tiles = smem_raw + L.shared_bytes;          // tile w of the round lives at tiles + w * L.bytes
  unsigned char* const wbase = tiles + (size_t)warp * L.bytes;
  double* wrho = reinterpret_cast<double*>(wbase + L.off_rho);    // [n][GPW]   1 / T
  double* wfac = reinterpret_cast<double*>(wbase + L.off_fac);    // [6(n-1)][GPW], later cbuf | dt | thr
  double* wy = reinterpret_cast<double*>(wbase + L.off_y);        // [3(n-1)][32] forward values, then knot states
  double* wt = reinterpret_cast<double*>(wbase + L.off_t);        // [GPW][n+1] stamps, later running duration sums
  double* ww = reinterpret_cast<double*>(wbase + L.off_w);        // [TPT][n+1][K] waypoints
  double* cbuf = wfac;                                            // [2][n][K][8] coefficients staged by THIS warp
  const size_t off_dts = L.off_fac + sizeof(double) * L.cbuf_doubles;   // [GPW] sample spacing, then [GPW][n] thresholds
  double* dts = reinterpret_cast<double*>(wbase + off_dts);
  int* thr = reinterpret_cast<int*>(dts + GPW);
  uint8_t* piece_of = wbase + L.off_piece;                        // [GPW][S]
  double* wT = reinterpret_cast<double*>(wbase + L.off_wire);     // wire mode: [GPW][n] durations T_i, then
  float* wstage = reinterpret_cast<float*>(wT + (size_t)GPW * n); //   [n][1 + 8K] one trajectory's float32 matrix
  PoseRing<NP>& ring = *reinterpret_cast<PoseRing<NP>*>(wbase + L.off_ring);
  unsigned ring_head = 0u, ring_tail = 0u;  // warp-uniform
  const bool wire_mat = wire.count > 0 && wire.mat[0] != nullptr;
  const int width = 1 + MST_NCOEF * K;

  int round = 0;                 // low 15 bits travel with every queued pose
  long long round_b0 = 0;        // first trajectory of the CTA's current round
  bool sampling = false;         // true while this round's flag rows are still in shared memory
  // a decided pose: only collisions need recording (its flag byte was preset to 0)
  auto report = [&](int b, int id1, bool h) {
    if (!h) return;
    const int s = id1 & 0xffff;
    if (sampling && ((id1 >> 16) & 0x7fff) == (round & 0x7fff)) {
      const unsigned si = s_slot[(int)(b - round_b0)];
      unsigned char* tb = tiles + (size_t)(si & 0xff) * L.bytes;
      const int q = (si >> 8) & 0xff;
      tb[L.off_hit + (size_t)q * S + s] = 1;
      tb[L.off_any + q] = 1;
    } else {  // a pose of an earlier round: its rows are in HBM already
      hit[(size_t)b * S + s] = 1;
      any_hit[b] = 1;
      for (int t = 0; t < wire.count; ++t)
        if (wire.hit[t] != nullptr) {
          wire.hit[t][(size_t)(wire.row0 + b) * S + s] = 1;
          wire.any[t][wire.row0 + b] = 1;
        }
    }
  };

  int par = 0;
  for (;;) {
    const long long set0 = s_set0[par];
    // epilogue round (no sets left): phases 1 and 3 are skipped, phase 2 only decides the poses still
    // queued — through the SAME call site of the collision engine (one copy of its code in the kernel)
    const bool last = set0 >= sets;
    // ================= phase 1: this warp's tile =====================================================
    if (threadIdx.x == 0 && !last) { s_set0[par ^ 1] = atomicAdd(counters + 1, W); s_chunk = 0; }
    const long long set = set0 + warp;
    const long long g0 = set * GPW;
    const int cnt = last ? 0 : (int)max(0ll, min((long long)GPW, groups - g0));
    const int nbt = cnt * G;                 // trajectories of this tile
    const long long b0 = g0 * G;             // first trajectory of the tile
    round_b0 = set0 * GPW * G;
    unsigned okl = 0u;
    if (cnt > 0) {
      for (int i = lane; i < cnt * (n + 1); i += 32) wt[i] = __ldg(tstamps + (size_t)g0 * (n + 1) + i);
      for (int i = lane; i < cnt * (n + 1) * R; i += 32) ww[i] = __ldg(wp + (size_t)g0 * (n + 1) * R + i);
      __syncwarp();
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
      okl = __ballot_sync(FULL, mine && col == 0 && cls == 0);  // bit gl * R per solvable group
      __syncwarp();
      if (mine && cls == 0) {
        const double* wcol = ww + ((size_t)gl * G + d) * (n + 1) * K + k;
        condensed_forward<1>(wcol, K, n, 1, wrho + gl, wfac + gl, GPW, wy + lane, 32);
        condensed_backward_states<1>(n, 1, wrho + gl, wfac + gl, GPW, wy + lane, 32);
      }
      __syncwarp();
    }
    if (lane == 0) s_okl[warp] = okl;
    if (lane < TPT) s_ok[warp * TPT + lane] = (lane < nbt && ((okl >> ((lane / G) * R)) & 1u)) ? 1 : 0;
    if (okl != 0u) {
      auto group_ok = [&](int g) -> bool { return (okl >> (g * R)) & 1u; };
      // durations and status out (every trajectory of a group carries its own copy of the durations)
      for (int item = lane; item < nbt * n; item += 32) {
        const int q = (int)__umulhi((unsigned)item, magic_n), i = item - q * n, g = (int)((s_slot[q] >> 16) & 0xffu);
        if (group_ok(g)) {
          const double Ti = wt[g * (n + 1) + i + 1] - wt[g * (n + 1) + i];
          dur[(size_t)(b0 + q) * n + i] = Ti;
          if (wire_mat && q == g * G) wT[g * n + i] = Ti;
        }
      }
      if (lane < nbt) {
        wbase[L.off_any + lane] = 0;
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
        const int g = (int)__umulhi((unsigned)item, magic_n), i = item - g * n;
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
        const int g = (int)__umulhi((unsigned)item, magic_n), i = item - g * n;
        if (group_ok(g)) {
          const int from = i ? thr[item - 1] : 0, to = thr[item];
#pragma unroll 1
          for (int x = from; x < to; ++x) piece_of[g * S + x] = (uint8_t)i;
        }
      }
    }
    sampling = !last;
    // all tiles of the round are solved; s_set0[par ^ 1], s_chunk, s_okl are visible.  (Epilogue: every
    // warp has stored its last rows — a row store must not land on top of a late flag.)
    __syncthreads();

    // ================= phase 2: sample + collide, trajectories handed out chunk by chunk =============
    if (!last) {
      // next round's inputs towards L2 while this round's samples are worked on
      const long long nset = (long long)s_set0[par ^ 1] + warp;
      if (nset < sets) {
        const long long h0 = nset * GPW;
        const int c2 = (int)min((long long)GPW, groups - h0);
        const char* tb = reinterpret_cast<const char*>(tstamps + (size_t)h0 * (n + 1));
        const char* wb = reinterpret_cast<const char*>(wp + (size_t)h0 * (n + 1) * R);
        for (int off = lane * 128; off < c2 * (n + 1) * 8; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(tb + off));
        for (int off = lane * 128; off < c2 * (n + 1) * R * 8; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(wb + off));
      }
    }
    const int total = W * TPT;
    for (;;) {
      int lo = total, hi = total;
      if (!last) {
        int ck = 0;
        if (lane == 0) ck = atomicAdd(&s_chunk, 1);
        ck = __shfl_sync(FULL, ck, 0);
        chunk_range(ck, total, lo, hi);
        if (lo >= total) break;
      }
      const int nb = hi - lo, work = nb * S;
      // slot -> tile, trajectory within the tile, group, drone; advanced incrementally with tl
      int tw = 0, q = 0, g = 0, dd = 0;
      if (lo < total) { const unsigned si = s_slot[lo]; tw = si & 0xff; q = (si >> 8) & 0xff; g = (si >> 16) & 0xff; dd = si >> 24; }
      int tl = 0, s = lane, next_stage = 0;
      for (int base = 0;; base += 32, s += 32) {
        const bool more = base < work;
        if (more) {
        // trajectory j of the chunk is staged when the sampling front has left trajectory j - 2
        while (next_stage < nb && (next_stage < 2 || base >= (next_stage - 1) * S)) {
          __syncwarp();
          const int slot = lo + next_stage;
          const unsigned ssi = s_slot[slot];
          const int stw = ssi & 0xff, sq = (ssi >> 8) & 0xff, sg = (ssi >> 16) & 0xff;
          unsigned char* tb = tiles + (size_t)stw * L.bytes;
          if (s_ok[slot]) {
            const double* xy = reinterpret_cast<const double*>(tb + L.off_y);
            const double* xw = reinterpret_cast<const double*>(tb + L.off_w);
            const double* xrho = reinterpret_cast<const double*>(tb + L.off_rho);
            const double* xT = reinterpret_cast<const double*>(tb + L.off_wire);
            const long long bq = (round_b0 + slot);
            double* cb = cbuf + (size_t)(next_stage & 1) * n * K * MST_NCOEF;
            double* cd = coef + (size_t)bq * n * K * MST_NCOEF;
#pragma unroll 1
            for (int item = lane; item < n * K; item += 32) {
              const int i = item / K, kk = item - i * K;
              const int c = sq * K + kk;   // the lane that solved this column
              double v0 = 0.0, a0 = 0.0, j0 = 0.0, v1 = 0.0, a1 = 0.0, j1 = 0.0;
              if (i >= 1) { const double* x = xy + (size_t)(i - 1) * 3 * 32 + c; v0 = x[0]; a0 = x[32]; j0 = x[64]; }
              if (i + 1 < n) { const double* x = xy + (size_t)i * 3 * 32 + c; v1 = x[0]; a1 = x[32]; j1 = x[64]; }
              const double w0 = xw[((size_t)sq * (n + 1) + i) * K + kk], w1 = xw[((size_t)sq * (n + 1) + i + 1) * K + kk];
              double cc[MST_NCOEF];
              piece_coefficients(w0, w1 - w0, v0, a0, j0, v1, a1, j1, xrho[(size_t)i * GPW + sg], cc);
              double2* s2 = reinterpret_cast<double2*>(cb + (size_t)item * MST_NCOEF);
              double2* g2 = reinterpret_cast<double2*>(cd + (size_t)item * MST_NCOEF);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const double2 v = make_double2(cc[2 * e], cc[2 * e + 1]);
                s2[e] = v;
                g2[e] = v;
              }
              if (wire_mat) {
                float* ws = wstage + i * width + 1 + MST_NCOEF * kk;
#pragma unroll
                for (int e = 0; e < MST_NCOEF; ++e) ws[e] = (float)cc[e];
                if (kk == 0) wstage[i * width] = (float)xT[sg * n + i];
              }
            }
            if (wire_mat) {
              __syncwarp();
              const int words = n * width;
              const size_t row = (size_t)(wire.row0 + bq) * words;
              for (int t = 0; t < wire.count; ++t)
                if (wire.mat[t] != nullptr) {
                  float* dst = wire.mat[t] + row;
                  for (int w = lane; w < words; w += 32) dst[w] = wstage[w];
                }
            }
          }
          ++next_stage;
          __syncwarp();
        }
        while (s >= S) {
          s -= S; ++tl; ++q;
          if (++dd == G) { dd = 0; ++g; }
          if (q == TPT) { q = 0; g = 0; dd = 0; ++tw; }
        }
        bool active = base + lane < work;
        if (!active) {   // parked on the chunk's last sample (not recorded)
          const unsigned si = s_slot[hi - 1];
          tl = nb - 1; s = S - 1; tw = si & 0xff; q = (si >> 8) & 0xff; g = (si >> 16) & 0xff; dd = si >> 24;
        }
        const unsigned char* tb = tiles + (size_t)tw * L.bytes;
        active = active && s_ok[lo + tl];
        const double* kn = reinterpret_cast<const double*>(tb + L.off_t) + g * (n + 1);
        const double t = __dmul_rn((double)s, reinterpret_cast<const double*>(tb + off_dts)[g]);
        const int piece = min((int)tb[L.off_piece + (size_t)g * S + s], n - 1);
        const double local = __dsub_rn(t, kn[piece]);
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
        // yaw: only samples whose bounding sphere reaches the obstacle's box need the rotation (the others are
        // free whatever the heading); the decision is the one pose_near_environment makes first anyway
        bool reach = true;
        if (POSE == 1) {
          reach = sphere_near_environment(pp, rbb, evb);
          pp[3] = 0.0; pp[4] = 1.0;
          if (reach) sincos(pos[K - 1] * 0.5, &pp[3], &pp[4]);
        }
        const bool near = active && reach && pose_near_environment<POSE>(pp, rbb, evb);
        // every flag starts as 0; a queued pose that turns out to collide overwrites its byte
        if (active) const_cast<unsigned char*>(tb)[L.off_hit + (size_t)q * S + s] = 0;
        ring_push<POSE>(ring, ring_tail, near, pp, (int)(round_b0 + lo + tl), s | ((round & 0x7fff) << 16), -1, 0u, 0u);
        }
        // full batches of 32 while sampling; in the epilogue whatever is left
        const unsigned enough = (more || !last) ? 32u : 1u;
        while (ring_tail - ring_head >= enough)
          ring_drain<POSE>(ring, ring_head, ring_tail, (int)min(32u, ring_tail - ring_head), rb, rbb, ev, nv, report);
        if (!more) break;
      }
      if (last) break;
    }
    if (last) break;
    __syncthreads();   // every sample of the round is recorded (poses still queued: byte 0 for now)
    sampling = false;

    // ================= phase 3: this tile's flag rows =================================================
    if (okl != 0u) {
      auto group_ok = [&](int g) -> bool { return (okl >> (g * R)) & 1u; };
      const uint8_t* hitb = wbase + L.off_hit;
      const uint8_t* anyb = wbase + L.off_any;
      for (int t = -1; t < wire.count; ++t) {
        uint8_t* hdst = t < 0 ? hit : wire.hit[t];
        uint8_t* adst = t < 0 ? any_hit : wire.any[t];
        const size_t off = (size_t)(t < 0 ? 0 : wire.row0) + (size_t)b0;
        if (hdst == nullptr) continue;
        if ((S & 3) == 0 && ((reinterpret_cast<uintptr_t>(hdst) + off * S) & 3) == 0) {
          const unsigned* src = reinterpret_cast<const unsigned*>(hitb);
          unsigned* dst = reinterpret_cast<unsigned*>(hdst + off * S);
          const int wpr = S >> 2;   // words per row
          if (__popc(okl) == cnt) {   // the usual case: every group of the tile was solved here
            for (int w = lane; w < nbt * wpr; w += 32) dst[w] = src[w];
          } else {
            for (int w = lane; w < nbt * wpr; w += 32)
              if (group_ok((w / wpr) / G)) dst[w] = src[w];
          }
        } else {
          for (int w = lane; w < nbt * S; w += 32)
            if (group_ok((w / S) / G)) hdst[off * S + w] = hitb[w];
        }
        if (lane < nbt && group_ok(lane / G)) adst[off + lane] = anyb[lane];
      }
    }
    ++round;
    par ^= 1;
    // the next round's phase 1 only touches this warp's own solver arrays and tables; the flag rows
    // are not written again before the next phase-1 barrier
  }
  // A queued pose is at most a few hundred rounds old (the pipeline cuts batches into chunks of 2^22
  // trajectories), far from the 2^15 rounds its round tag can tell apart.
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
