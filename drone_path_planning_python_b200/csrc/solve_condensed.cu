// Fast solver: condensed block-tridiagonal LDL^T (condensed_core.cuh), one time group per
// thread, per-thread scratch in shared memory laid out slot-major ([slot][thread]) so a
// warp's accesses to one slot are 32 consecutive doubles (no bank conflicts).
//
// Replaces calculate_trajectory1D/4D (src/optimizations/calculatingTrajectories.py:37-213)
// for time groups whose duration spread allows it; every other group is appended to a
// device-side list that the banded pivoted-LU kernel consumes right after (no host sync).
#include "condensed_core.cuh"
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "farcull.cuh"
#include "stage.cuh"

namespace mst {

__host__ size_t condensed_workspace_bytes(int groups) {
  return sizeof(int) * (64 + (size_t)groups);
}

__device__ __forceinline__ void fill_failed(double* coef, int* info, size_t traj, int n, int K, int code) {
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  info[traj] = code;
  double* c = coef + traj * (size_t)n * K * MST_NCOEF;
  for (int e = 0; e < n * K * MST_NCOEF; ++e) c[e] = qnan;
}

// STAGED = true: the CTA's slice of the inputs (consecutive groups -> one contiguous byte range
// of `wp` and one of `tstamps`) is pulled into shared memory by two bulk (TMA) copies before
// the threads start, so the per-thread strided reads hit shared memory instead of L2: with the
// scratch taking most of the SM's shared memory the L1 is down to ~30 kB and the unstaged
// kernel spent 7 of every 10 issue-slot cycles on long-scoreboard stalls (profiles/
// r1_condensed_kernel_ncu.txt).  Used when G == 1 and the slices are 16-byte aligned.
template <int KC, bool STAGED>
__global__ void __launch_bounds__(128)
condensed_kernel(const double* __restrict__ wp, const double* __restrict__ tstamps, int groups, int n,
                 int K, int G, int force, double* __restrict__ coef, double* __restrict__ dur,
                 int* __restrict__ info, int* __restrict__ list, int* __restrict__ list_count) {
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) unsigned long long bar;
  const int stride = blockDim.x;
  double* scratch = sm + threadIdx.x;
  // staged tiles behind the scratch: wp[blockDim][n+1][K], t[blockDim][n+1]
  const int slots = condensed_slots(n, K);
  double* wp_tile = sm + (size_t)slots * blockDim.x;
  double* t_tile = wp_tile + (size_t)blockDim.x * (n + 1) * K;
  unsigned phase = 0;
  if (STAGED && threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (STAGED) __syncthreads();
  for (long long g0 = blockIdx.x * (long long)blockDim.x; g0 < groups; g0 += (long long)gridDim.x * blockDim.x) {
    const long long g = g0 + threadIdx.x;
    if (STAGED) {
      const int cnt = (int)min((long long)blockDim.x, groups - g0);
      __syncthreads();  // every thread is done with the previous tile
      if (cnt == (int)blockDim.x) {
        if (threadIdx.x == 0) {
          const unsigned wbytes = (unsigned)(cnt * (n + 1) * K * sizeof(double));
          const unsigned tbytes = (unsigned)(cnt * (n + 1) * sizeof(double));
          mbar_expect_tx(&bar, wbytes + tbytes);
          bulk_g2s(wp_tile, wp + (size_t)g0 * (n + 1) * K, wbytes, &bar);
          bulk_g2s(t_tile, tstamps + (size_t)g0 * (n + 1), tbytes, &bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1u;
      } else {  // last, partial tile: its byte count need not be a multiple of 16
        for (int i = threadIdx.x; i < cnt * (n + 1) * K; i += blockDim.x) wp_tile[i] = wp[(size_t)g0 * (n + 1) * K + i];
        for (int i = threadIdx.x; i < cnt * (n + 1); i += blockDim.x) t_tile[i] = tstamps[(size_t)g0 * (n + 1) + i];
        __syncthreads();
      }
    }
    if (g >= groups) continue;
    const double* tg = STAGED ? t_tile + (size_t)threadIdx.x * (n + 1) : tstamps + (size_t)g * (n + 1);
    double Tmin, Tmax;
    const int cls = classify_times(tg, n, &Tmin, &Tmax);
    // durations out (every trajectory of the group carries its own copy, as
    // PiecewisePolynomial.time_durations does)
    for (int d = 0; d < G; ++d) {
      double* dd = dur + ((size_t)g * G + d) * n;
      for (int i = 0; i < n; ++i) dd[i] = tg[i + 1] - tg[i];
    }
    if (cls >= 2) {
      for (int d = 0; d < G; ++d)
        fill_failed(coef, info, (size_t)g * G + d, n, K, cls == 2 ? MST_INFO_DECREASING : MST_INFO_NONFINITE);
      continue;
    }
    if (cls == 1 || force < 0) {
      if (force <= 0) {  // auto mode (or scratch does not fit): hand over to the pivoted solver
        list[atomicAdd(list_count, 1)] = (int)g;
        continue;
      }
      // forced condensed solve of something it cannot reproduce: a zero-length piece
      // (singular in the reference too) or the t[0] != 0 quirk
      if (!(Tmin > 0.0) || tg[0] != 0.0) {
        for (int d = 0; d < G; ++d)
          fill_failed(coef, info, (size_t)g * G + d, n, K, !(Tmin > 0.0) ? 1 : MST_INFO_DECLINED);
        continue;
      }
    }
    double* fac = scratch + (size_t)n * stride;
    double* ys = scratch + ((size_t)n + 6 * (size_t)(n - 1)) * stride;
    for (int i = 0; i < n; ++i) scratch[(size_t)i * stride] = tg[i + 1] - tg[i];
    condensed_factor(n, scratch, fac, stride);
    for (int d = 0; d < G; ++d) {
      const size_t traj = (size_t)g * G + d;
      const double* wpd = STAGED ? wp_tile + (size_t)threadIdx.x * (n + 1) * K : wp + traj * (size_t)(n + 1) * K;
      condensed_forward<KC>(wpd, K, n, K, scratch, fac, stride, ys, stride);
      double* cd = coef + traj * (size_t)n * K * MST_NCOEF;
      condensed_backward<KC>(wpd, K, n, K, scratch, fac, stride, ys, stride,
                             [&](int piece, int k, const double* c, double) {
                               double2* dst = reinterpret_cast<double2*>(cd + ((size_t)piece * K + k) * MST_NCOEF);
                               dst[0] = make_double2(c[0], c[1]);
                               dst[1] = make_double2(c[2], c[3]);
                               dst[2] = make_double2(c[4], c[5]);
                               dst[3] = make_double2(c[6], c[7]);
                             });
      info[traj] = MST_INFO_OK;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Lane-per-column variant.  One LANE per right-hand side column (trajectory d of the group,
// axis k): a warp holds GPW = 32 / (G*K) time groups, the LDL^T factors of a group live once in
// shared memory (written by the group's first lane, read by all its columns as broadcasts) and
// every lane keeps only its own column's forward values (3 per knot).  Shared memory per
// trajectory is the same as with one thread per group, but it now feeds G*K times as many
// threads: 14 warps per SM instead of 4 at n = 10, K = 3, which is what the latency-bound
// recurrences need (profiles/README.md).  The group's factorisation is not repeated per
// column; the price is that the other columns' lanes idle during it.  The lanes of one
// trajectory write adjacent 64-byte coefficient rows and read adjacent waypoint values.
// Inputs of the warp's groups are one contiguous range each: copied into shared memory with
// coalesced loads before use.
constexpr int COLS_MAX_WARPS = 7;
constexpr int COLS_TREGS = 4, COLS_WREGS = 12;  // register tile of the input pipeline, doubles per lane

// CULL = true (pipeline, n <= 32): every lane also bounds its pieces' positions while their coefficients are
// in registers and writes the far-piece bits of its axis (farcull.cuh) for the sampling kernel.
// MAT = true: the float32 polynomial matrix rows as a second output (FarCull::mat).
template <bool CULL, bool MAT>
__global__ void __launch_bounds__(COLS_MAX_WARPS * 32, 2)   // two CTAs of seven warps: 14 warps per SM need <= 146 registers
condensed_cols_kernel(const double* __restrict__ wp, const double* __restrict__ tstamps, int groups, int n,
                      int K, int G, int force, double* __restrict__ coef, double* __restrict__ dur,
                      int* __restrict__ info, int* __restrict__ list, int* __restrict__ list_count, FarCull cull) {
  extern __shared__ __align__(16) double sm[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int R = G * K;        // columns per group
  const int GPW = 32 / R;     // groups per warp
  const int gl = lane / R;    // group slot of this lane
  const int col = lane - gl * R;
  const int d = col / K, k = col - d * K;
  const bool lane_used = gl < GPW;
  // per-warp shared memory: rho[n][GPW] | fac[6(n-1)][GPW] | y[3(n-1)][32] | t[GPW][n+1] | wp[n+1][WS]
  // The waypoint tile is kept waypoint-major, one column of the warp per lane: lane c reads ww[i * WS + c], 32
  // consecutive doubles per access (the trajectory-major order of the input put three lanes on every bank pair:
  // half as many wavefronts again on the solver's most frequent shared-memory loads).  WS = 32 + K: consecutive
  // elements of the input (axis fastest, then waypoint) land K apart per waypoint — the transposing stores are
  // conflict free as well.
  const int WS = 32 + K;
  const size_t per_warp = (size_t)GPW * (n + 6 * (n - 1)) + 32 * (size_t)(3 * (n - 1)) + (size_t)GPW * (n + 1) +
                          (size_t)(n + 1) * WS;
  double* wrho = sm + warp * per_warp;
  double* wfac = wrho + (size_t)GPW * n;
  double* wy = wfac + (size_t)GPW * 6 * (n - 1);
  double* wt = wy + 32 * (size_t)(3 * (n - 1));
  double* ww = wt + (size_t)GPW * (n + 1);
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  const long long sets = ((long long)groups + GPW - 1) / GPW;
  // Software pipeline of the input copies: the NEXT set's stamps and waypoints are loaded into
  // registers before this set is solved and stored to shared memory afterwards, so their
  // latency (23 % of the kernel's stall samples when the copy was synchronous) is covered by
  // the solve.  Used when a set's inputs fit the register tile.
  const bool piped = GPW * (n + 1) <= 32 * COLS_TREGS && GPW * (n + 1) * R <= 32 * COLS_WREGS;
  double tr[COLS_TREGS], wr[COLS_WREGS];
  // where element e = [trajectory][waypoint][axis] of a set's waypoint slice goes in the tile (TileWalk,
  // condensed_core.cuh): a lane's elements are 32 apart
  const TileWalk walk0 = TileWalk::start(lane, n, K);
  int wdst[COLS_WREGS];
  {
    TileWalk w = walk0;
#pragma unroll
    for (int j = 0; j < COLS_WREGS; ++j) { wdst[j] = w.slot(WS, K); w.advance32(n, K); }
  }
  auto load_set = [&](long long set2) {
    if (set2 >= sets) return;
    const long long h0 = set2 * GPW;
    const int c2 = (int)min((long long)GPW, groups - h0);
    const double* tb = tstamps + (size_t)h0 * (n + 1);
    const double* wb = wp + (size_t)h0 * (n + 1) * R;
#pragma unroll
    for (int j = 0; j < COLS_TREGS; ++j) if (lane + 32 * j < c2 * (n + 1)) tr[j] = __ldg(tb + lane + 32 * j);
#pragma unroll
    for (int j = 0; j < COLS_WREGS; ++j) if (lane + 32 * j < c2 * (n + 1) * R) wr[j] = __ldg(wb + lane + 32 * j);
  };
  if (piped) load_set(blockIdx.x * (long long)warps + warp);
  for (long long set = blockIdx.x * (long long)warps + warp; set < sets; set += (long long)gridDim.x * warps) {
    const long long g0 = set * GPW;
    const int cnt = (int)min((long long)GPW, groups - g0);
    __syncwarp();  // the previous set's tiles are no longer read
    if (piped) {
      // this set's inputs were loaded into registers while the previous set was solved
#pragma unroll
      for (int j = 0; j < COLS_TREGS; ++j) if (lane + 32 * j < cnt * (n + 1)) wt[lane + 32 * j] = tr[j];
#pragma unroll
      for (int j = 0; j < COLS_WREGS; ++j) if (lane + 32 * j < cnt * (n + 1) * R) ww[wdst[j]] = wr[j];
    } else {
      for (int i = lane; i < cnt * (n + 1); i += 32) wt[i] = tstamps[(size_t)g0 * (n + 1) + i];
      TileWalk w = walk0;
      for (int i = lane; i < cnt * (n + 1) * R; i += 32) {
        ww[w.slot(WS, K)] = wp[(size_t)g0 * (n + 1) * R + i];
        w.advance32(n, K);
      }
    }
    __syncwarp();
    if (piped) load_set(set + (long long)gridDim.x * warps);
    const bool mine = lane_used && gl < cnt;
    const long long g = g0 + gl;
    const double* tg = wt + (size_t)gl * (n + 1);
    // classification and factorisation by the group's first column
    int cls = 0;
    if (mine && col == 0) {
      double Tmin, Tmax;
      cls = classify_times(tg, n, &Tmin, &Tmax);
      if (cls == 1 && force > 0 && Tmin > 0.0 && tg[0] == 0.0) cls = 0;            // forced: solve anyway
      else if (cls == 1 && force > 0) cls = !(Tmin > 0.0) ? 4 : 5;                 // forced but impossible
      if (cls == 1 || (cls == 0 && force < 0)) { cls = 1; list[atomicAdd(list_count, 1)] = (int)g; }
      if (cls == 0) {
        double* rho = wrho + gl;
        for (int i = 0; i < n; ++i) rho[(size_t)i * GPW] = tg[i + 1] - tg[i];
        condensed_factor(n, rho, wfac + gl, GPW);
      }
    }
    cls = __shfl_sync(FULL, cls, lane_used ? gl * R : 0);
    const unsigned solving = __ballot_sync(FULL, mine && cls == 0);   // the lanes that run the sweeps below
    __syncwarp();
    if (!mine) continue;
    const size_t traj = (size_t)g * G + d;
    if (k == 0) {
      double* dd = dur + traj * n;
      for (int i = 0; i < n; ++i) dd[i] = tg[i + 1] - tg[i];
    }
    if (cls == 1) continue;  // the pivoted solver will write coefficients and status
    double* cd = coef + traj * (size_t)n * K * MST_NCOEF + (size_t)k * MST_NCOEF;
    if (cls != 0) {
      for (int i = 0; i < n; ++i)
        for (int e = 0; e < MST_NCOEF; ++e) cd[(size_t)i * K * MST_NCOEF + e] = qnan;
      if (MAT) {   // the matrix rows of a failed trajectory: durations as they are, NaN coefficients
        for (int i = 0; i < n; ++i) {
          float* row = cull.mat + (traj * n + i) * (size_t)(1 + MST_NCOEF * K);
          if (k == 0) row[0] = (float)(tg[i + 1] - tg[i]);
          for (int e = 0; e < MST_NCOEF; ++e) row[1 + MST_NCOEF * k + e] = (float)qnan;
        }
      }
      if (k == 0) info[traj] = cls == 2 ? MST_INFO_DECREASING : (cls == 3 ? MST_INFO_NONFINITE : (cls == 4 ? 1 : MST_INFO_DECLINED));
      continue;
    }
    const double* wcol = ww + lane;   // lane = (gl G + d) K + k: its column of the tile
    condensed_forward<1>(wcol, WS, n, 1, wrho + gl, wfac + gl, GPW, wy + lane, 32);
    unsigned farbits = 0u;
    condensed_backward<1>(wcol, WS, n, 1, wrho + gl, wfac + gl, GPW, wy + lane, 32,
                          [&](int piece, int, const double* c, double) {
                            double2* dst = reinterpret_cast<double2*>(cd + (size_t)piece * K * MST_NCOEF);
                            dst[0] = make_double2(c[0], c[1]);
                            dst[1] = make_double2(c[2], c[3]);
                            dst[2] = make_double2(c[4], c[5]);
                            dst[3] = make_double2(c[6], c[7]);
                            if (MAT) {
                              float* row = cull.mat + (traj * n + piece) * (size_t)(1 + MST_NCOEF * K);
                              if (k == 0) row[0] = (float)(tg[piece + 1] - tg[piece]);
#pragma unroll
                              for (int e = 0; e < MST_NCOEF; ++e) row[1 + MST_NCOEF * k + e] = (float)c[e];
                            }
                          },
                          [&](int piece, int, double w0, double w1, double v0, double a0, double j0, double v1, double a1,
                              double j1, double) {
                            // far-piece bound from the end states (farcull.cuh)
                            if (!CULL) return;
                            const Hull8 h = hull_from_states(w0, w1, v0, a0, j0, v1, a1, j1, tg[piece + 1] - tg[piece]);
                            double lo = cull.lo[k < 3 ? k : 0], hi = cull.hi[k < 3 ? k : 0];
                            if (cull.yaw) {
                              // the robot turns by the sampled yaw (4th axis): its lane bounds |yaw| over the piece,
                              // the position lanes of the trajectory widen the obstacle box by the robot's reach
                              const double mine_abs = k == 3 ? hull_max_abs(w0, w1, h) : 0.0;
                              const double t = __shfl_sync(solving, mine_abs, lane - k + 3);
                              if (k < 3) axis_reach(cull, k, t, &lo, &hi);
                            }
                            if (k < 3 && hull_far(w0, w1, h, lo, hi)) farbits |= 1u << piece;
                          });
    if (CULL && k < 3) cull.mask[traj * 3 + k] = farbits;
    if (k == 0) info[traj] = MST_INFO_OK;
  }
}

static size_t cols_smem_per_warp(int n, int K, int G) {
  const int R = G * K, GPW = 32 / R;
  return sizeof(double) * ((size_t)GPW * (n + 6 * (n - 1)) + 32 * (size_t)(3 * (n - 1)) + (size_t)GPW * (n + 1) +
                           (size_t)(n + 1) * (32 + K));
}

// The occupancy search behind the launcher's choice of warps per CTA, remembered per (kernel, shared memory per
// warp): a dozen driver calls, too many for every launch of a single-trajectory solve.
static int cols_warps_per_cta(const void* kern, size_t per_warp, int* resident_ctas) {
  struct Entry { const void* kern; size_t per_warp; int wpb, ctas; };
  static Entry cache[32];
  static int used = 0;
  static std::mutex lock;
  std::lock_guard<std::mutex> guard(lock);
  for (int i = 0; i < used; ++i)
    if (cache[i].kern == kern && cache[i].per_warp == per_warp) { *resident_ctas = cache[i].ctas; return cache[i].wpb; }
  int wpb = 1, best = 0, best_ctas = 1;
  for (int w = 1; w <= COLS_MAX_WARPS && w * per_warp + 64 <= MST_MAX_SMEM; ++w) {
    if (allow_dynamic_smem(kern, w * per_warp) != MST_OK) break;
    int ctas = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, 32 * w, w * per_warp) != cudaSuccess) break;
    if (ctas * w > best) { best = ctas * w; wpb = w; best_ctas = ctas; }
  }
  if (used < 32) cache[used++] = Entry{kern, per_warp, wpb, best_ctas};
  *resident_ctas = best_ctas;
  return wpb;
}

// cull (may be null): far-piece bits for the pipeline's sampling kernel (cull->mask, honoured with n <= 32;
// the caller zeroes the mask beforehand, so groups solved any other way simply have no far pieces) and / or
// the float32 polynomial matrix (cull->mat).  Returns MST_ERR_TOO_LARGE when a matrix is asked for and the
// sizes go to the thread-per-group fallback kernel, which does not write it.
int launch_condensed(const double* wp, const double* t, int groups, int n, int K, int G, int force,
                     double* coef, double* dur, int* info, int* list, int* list_count,
                     cudaStream_t stream, const FarCull* cull) {
  if (K > 4) {
    // more axes than the per-thread register tile: everything goes to the pivoted solver
    if (force) return MST_ERR_INVALID;
  }
  cudaError_t e = cudaMemsetAsync(list_count, 0, sizeof(int), stream);
  if (e != cudaSuccess) { note_cuda_error(e); return MST_ERR_CUDA; }
  static const bool use_cols = getenv("MST_CONDENSED_THREAD_PER_GROUP") == nullptr;
  if (use_cols && G * K <= 32 && n >= 1) {
    const size_t per_warp = cols_smem_per_warp(n, K, G);
    if (per_warp + 64 <= MST_MAX_SMEM) {
      const int GPW = 32 / (G * K);
      const long long sets = ((long long)groups + GPW - 1) / GPW;
      const bool culling = cull != nullptr && cull->mask != nullptr && n <= 32;
      const bool packing = cull != nullptr && cull->mat != nullptr;
      auto kern = culling ? (packing ? condensed_cols_kernel<true, true> : condensed_cols_kernel<true, false>)
                          : (packing ? condensed_cols_kernel<false, true> : condensed_cols_kernel<false, false>);
      FarCull fc;
      if (cull) fc = *cull; else memset(&fc, 0, sizeof(fc));
      // warps per CTA (the warps are independent): whatever keeps most warps resident — every CTA pays 1 kB of
      // reserved shared memory, so few large CTAs pack an SM better than many small ones (n = 10, K = 3: two CTAs
      // of 7 warps = 14 warps, against 12 with CTAs of 2)
      int resident = 1;
      const int wpb = cols_warps_per_cta((const void*)kern, per_warp, &resident);
      const int rc = allow_dynamic_smem((const void*)kern, wpb * per_warp);
      if (rc != MST_OK) return rc;
      long long blocks = (sets + wpb - 1) / wpb;
      // many more CTAs than fit at once: the hardware scheduler then evens out the SMs (measured per 1 M
      // trajectories: 7 CTAs per SM = one wave 1.00 ms, 16: 1.00, 32: 0.92, 64: 0.90, 128: 0.885, one CTA per
      // 2 sets: 0.90; tools/gpu_cols_ab.sh)
      static const int per_sm_env = getenv("MST_COLS_CTAS_PER_SM") ? atoi(getenv("MST_COLS_CTAS_PER_SM")) : 0;   // A/B
      (void)resident;
      const long long cap = (long long)MST_SM_COUNT * (per_sm_env > 0 ? per_sm_env : 128);
      if (blocks > cap) blocks = cap;
      if (blocks < 1) blocks = 1;
      kern<<<(unsigned)blocks, 32 * wpb, wpb * per_warp, stream>>>(wp, t, groups, n, K, G, force, coef, dur, info, list,
                                                                   list_count, fc);
      return check_launch();
    }
  }
  if (cull != nullptr && cull->mat != nullptr) return MST_ERR_TOO_LARGE;
  const int Kc = K > 4 ? 4 : K;
  const size_t scratch_per_thread = sizeof(double) * (size_t)condensed_slots(n, Kc);
  // staged input tiles (G == 1, 16-byte aligned slices for any tile start)
  const size_t tile_per_thread = sizeof(double) * (size_t)(n + 1) * (Kc + 1);
  auto pick_threads = [](size_t per_thread) {
    int th = (int)((MST_MAX_SMEM - 1024) / per_thread);
    if (th >= 32) th -= th % 32;  // long trajectories: a partial warp per CTA still works
    if (th > 128) th = 128;
    if (th > 64 && per_thread * 64 * 2 <= MST_MAX_SMEM) th = 64;  // more CTAs per SM
    return th;
  };
  int threads = pick_threads(scratch_per_thread + tile_per_thread);
  bool staged = G == 1 && K <= 4 && threads >= 32 && ((uintptr_t)wp % 16 == 0) && ((uintptr_t)t % 16 == 0) &&
                ((size_t)threads * (n + 1) * K * sizeof(double)) % 16 == 0 &&
                ((size_t)threads * (n + 1) * sizeof(double)) % 16 == 0;
  if (!staged) threads = pick_threads(scratch_per_thread);
  const size_t per_thread = staged ? scratch_per_thread + tile_per_thread : scratch_per_thread;
  const bool fits = threads >= 8 && K <= 4;
  if (!fits && force) return MST_ERR_TOO_LARGE;
  if (!fits) threads = 32;
  const size_t smem = fits ? per_thread * threads : 0;
  void (*kern)(const double*, const double*, int, int, int, int, int, double*, double*, int*, int*, int*);
  if (K <= 3) kern = staged ? condensed_kernel<3, true> : condensed_kernel<3, false>;
  else kern = staged ? condensed_kernel<4, true> : condensed_kernel<4, false>;
  {
    const int rc = allow_dynamic_smem((const void*)kern, smem);
    if (rc != MST_OK) return rc;
  }
  long long blocks = ((long long)groups + threads - 1) / threads;
  const long long cap = (long long)MST_SM_COUNT * (staged ? 8 : 64);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  // when the scratch does not fit (very long trajectories) every group is declined
  kern<<<(unsigned)blocks, threads, smem, stream>>>(wp, t, groups, n, K, G, fits ? force : -1, coef, dur,
                                                     info, list, list_count);
  return check_launch();
}

}  // namespace mst
