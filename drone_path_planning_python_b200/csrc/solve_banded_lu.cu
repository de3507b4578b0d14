// Robust solver: the reference's own 8n x 8n system factorised by banded LU with partial pivoting,
// one warp per time group (one matrix, R = G*K right-hand sides).
//
// Replaces calculate_trajectory1D (src/optimizations/calculatingTrajectories.py:37-197)
// including its np.linalg.solve (:137).  The row layout follows :65-128 exactly
//  so pivoting sees the matrix LAPACK dgesv sees (entries t^k are formed by repeated
// multiplication here, by libm pow in the reference: equal to a few ulps, not always bitwise); the band is
// kl = 10 below / ku = 7 above the diagonal.  `A` never exists in HBM, and only a window of it exists at
// all: banded_core.cuh holds the arithmetic and the storage scheme (window ring in shared memory,
// finished columns of U in an L2-resident per-warp scratch, untouched columns recomputed from the
// durations when they enter the window).
#include <stdlib.h>

#include "banded_core.cuh"

namespace mst {

constexpr int BANDED_WARPS = 4;        // warps (time groups in flight) per CTA
constexpr int BANDED_MIN_CTAS = 6;     // resident CTAs the register bound keeps possible (80 registers)
constexpr int BANDED_WARPS_PER_SM = BANDED_WARPS * BANDED_MIN_CTAS;
constexpr int BANDED_VARIANT = 1;      // see banded_lu_kernel (MST_LU_VARIANT overrides, for A/B measurements)

__host__ size_t banded_lu_smem_per_warp(int n, int R) { return sizeof(double) * band_warp_doubles(n, R); }

struct BandedPlan {
  int warps;        // per CTA
  size_t smem;      // dynamic shared memory per CTA
  int ctas_per_sm;  // resident CTAs the shared memory and the 80-register bound allow
};

// 0 warps: one group does not fit an SM's shared memory
__host__ static BandedPlan banded_plan(int n, int R) {
  BandedPlan p{BANDED_WARPS, 0, 0};
  const size_t per_warp = banded_lu_smem_per_warp(n, R);
  const size_t table = sizeof(double) * BAND_TABLE_DOUBLES;
  while (p.warps > 1 && p.warps * per_warp + table > MST_MAX_SMEM) p.warps >>= 1;
  p.smem = p.warps * per_warp + table;
  if (p.smem > MST_MAX_SMEM) { p.warps = 0; return p; }
  const size_t sm_total = 228 * 1024, cta_reserve = 1024;
  p.ctas_per_sm = (int)(sm_total / (p.smem + cta_reserve));
  if (p.ctas_per_sm > BANDED_WARPS_PER_SM / p.warps) p.ctas_per_sm = BANDED_WARPS_PER_SM / p.warps;
  if (p.ctas_per_sm < 1) p.ctas_per_sm = 1;
  return p;
}

__host__ static long long banded_grid(const BandedPlan& p, int groups) {
  long long blocks = ((long long)groups + p.warps - 1) / p.warps;
  const long long cap = (long long)MST_SM_COUNT * p.ctas_per_sm;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : blocks;
}

// device scratch for the finished columns of U: one slab of 18 x 8n doubles per warp of the grid
__host__ size_t banded_lu_scratch_bytes(int groups, int n, int R) {
  const BandedPlan p = banded_plan(n, R);
  if (p.warps == 0 || groups < 1) return 0;
  return sizeof(double) * (size_t)banded_grid(p, groups) * p.warps * UROWS * MST_NCOEF * n;
}

// V: 1 = warp-wide pivot search (default), 0 = every lane runs the comparison tree (banded_core.cuh)
template <int V>
__global__ void __launch_bounds__(BANDED_WARPS * 32, BANDED_MIN_CTAS)
banded_lu_kernel(const double* __restrict__ wp, const double* __restrict__ tstamps,
                 int groups, int n, int K, int G, const int* __restrict__ list,
                 const int* __restrict__ list_count, double* __restrict__ coef,
                 double* __restrict__ dur, int* __restrict__ info, double* __restrict__ scratch,
                 int* __restrict__ ticket) {
  extern __shared__ double smem[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int N = MST_NCOEF * n;
  const int R = G * K;
  const int NS = band_rhs_stride(n);
  double* ff = smem;   // tables shared by the CTA: ff [8][8], cf [8][LD], pi [8][LD] bytes
  double* cf = ff + 64;
  unsigned char* pi = reinterpret_cast<unsigned char*>(cf + MST_NCOEF * LD);
  double* W = smem + BAND_TABLE_DOUBLES + warp * band_warp_doubles(n, R);   // window ring [WCOLS][LD]
  double* pw = W + WCOLS * LD;                               // [n+1][8]
  double* Bs = pw + MST_NCOEF * (n + 1);                     // [R][NS]
  double* Ug = scratch + ((size_t)blockIdx.x * warps_per_block + warp) * UROWS * N;   // [N][UROWS]
  const int todo = list ? *list_count : groups;
  if (blockIdx.x * warps_per_block >= todo) return;   // nothing for this CTA (the usual case behind the condensed solver)
  for (int e = threadIdx.x; e < 64 + MST_NCOEF * LD; e += blockDim.x) band_table_entry(e, ff, cf, pi);
  __syncthreads();
  const BandSystem sys{n, N, pw, ff, cf, pi};
  const bool mat_lane = lane < MAT_LANES;           // owns column j + 1 + lane of the window
  const int rhs0 = lane - MAT_LANES;                // first right-hand side of the lane (if >= 0)

  // groups are handed out by a device counter (null: fixed stride): equal work per group, but the SMs do not
  // run equally fast, and a fixed share per warp waits for the slowest
  int drawn = 0;
  if (ticket != nullptr && lane == 0) drawn = atomicAdd(ticket, 1);
  for (int item = ticket != nullptr ? __shfl_sync(FULL, drawn, 0) : blockIdx.x * warps_per_block + warp; item < todo;
       item = ticket != nullptr ? __shfl_sync(FULL, drawn, 0) : item + gridDim.x * warps_per_block) {
    if (ticket != nullptr && lane == 0) drawn = atomicAdd(ticket, 1);
    const int g = list ? list[item] : item;
    const double* tg = tstamps + (size_t)g * (n + 1);

    // ---- durations, their powers, input checks ------------------------------------------
    const double t0 = tg[0];
    int bad = 0;
    if (!(t0 >= 0.0)) bad = isfinite(t0) ? 1 : 2;
    for (int i = lane; i <= n; i += 32) {
      const double T = i < n ? tg[i + 1] - tg[i] : t0;
      band_powers(T, pw + MST_NCOEF * i);
      if (i < n && (!(T >= 0.0) || !isfinite(T))) bad = max(bad, isfinite(T) ? 1 : 2);   // +inf counts as non-finite input
    }
    bad = __reduce_max_sync(FULL, bad);
    for (int e = lane; e < G * n; e += 32)
      dur[((size_t)g * G + e / n) * n + e % n] = tg[e % n + 1] - tg[e % n];
    if (bad) {
      const double qnan = __longlong_as_double(0x7ff8000000000000LL);
      for (int d = 0; d < G; ++d) {
        const size_t traj = (size_t)g * G + d;
        if (lane == 0) info[traj] = bad == 1 ? MST_INFO_DECREASING : MST_INFO_NONFINITE;
        for (int e = lane; e < n * K * MST_NCOEF; e += 32) coef[traj * n * K * MST_NCOEF + e] = qnan;
      }
      __syncwarp();
      continue;
    }
    __syncwarp();   // pw is complete

    // ---- right-hand sides and the first kv + 1 columns ------------------------------------
    for (int e = lane; e < R * NS; e += 32) Bs[e] = 0.0;
    for (int e = lane; e < (KV + 1) * LD; e += 32) {
      const int c = e / LD, o = e - c * LD;
      W[e] = c < N ? band_entry_fast(sys, c, o) : 0.0;
    }
    __syncwarp();
    for (int e = lane; e < R * (n + 1); e += 32) {
      const int r = e / (n + 1), i = e - r * (n + 1);
      const int d = r / K, k = r - d * K;
      const double v = wp[(((size_t)g * G + d) * (n + 1) + i) * K + k];
      double* b = Bs + (size_t)r * NS;
      if (i == 0) b[0] = v;
      else if (i == n) b[N - 4] = v;
      else {
        b[4 + MST_NCOEF * (i - 1) + 6] = v;
        b[4 + MST_NCOEF * (i - 1) + 7] = v;
      }
    }
    __syncwarp();

    // ---- banded LU with partial pivoting, right-hand sides carried along --------------------
    // Lane l < kv owns column j + 1 + l of the window — its rows j..j+kl are contiguous in the band,
    // a stride of LD - 1 = 27 doubles apart between lanes (16 lanes = 16 bank pairs: one wavefront) — and
    // the lanes from kv on own one right-hand side each (more than 15 of them: the extra ones in further
    // passes).  EVERY lane reads column j (broadcast loads) and repeats the pivot search and the
    // multipliers in registers, so a step needs no reduction and no broadcast.  Column j itself is not
    // updated (the multipliers are not kept: the right-hand sides move along), which leaves ONE warp
    // barrier per step: what a step writes (columns j+1.., the right-hand sides, the slot column j-2 left)
    // is disjoint from what it reads before writing (column j, column j-1 on its way out).
    int singular_at = 0;
    double rinv_prev = 0.0;
    // three walking pointers, each one column slot further every step and back to the first slot after the
    // last (a countdown instead of a comparison against a pointer the compiler would recompute every step):
    //   colj    column j from its diagonal down;
    //   mine    row j of my column of the window (matrix lanes) / of my right-hand side (those just move one
    //           row down and never wrap);
    //   leaving my band position of the slot that is being refilled (column j-1 goes, j+kv+1 comes).
    // (element offsets from the start of the dynamic shared memory, made opaque to the compiler: carried in a
    // register instead of being re-derived from the kernel parameters every step, and still shared-memory
    // accesses — an opaque POINTER turns them into generic loads.)
    int colj = (int)(W - smem) + KV;
    int colj_left = WCOLS;
    int mine = (int)((mat_lane ? W + (lane + 1) * LD + KV - (lane + 1) : Bs + (size_t)rhs0 * NS) - smem);
    const int mine_step = mat_lane ? LD : 1;
    int mine_left = mat_lane ? WCOLS - (lane + 1) : 0x7fffffff;
    int live = mat_lane ? N - (lane + 1) : (rhs0 < R ? N : 0);   // steps this lane still takes part in
    int leaving = (int)(W - smem) + (WCOLS - 1) * LD + min(lane, LD - 1);    // column j-1 (none yet)
    int leaving_left = 1;
    int entering = leaving - LD;                                            // where column j+kv+1 goes
    int entering_left = 2;
    double* ucol = Ug + min(lane, KV) - UROWS;                          // column j-1 of the scratch
    asm volatile("" : "+r"(colj), "+r"(mine), "+r"(leaving), "+r"(entering));
    for (int j = 0; j < N; ++j) {
      double l[KL + 1];
      int jp;
      double rinv, leaves = 0.0, enters = 0.0;
      if (band_pivot<V != 0>(smem + colj, lane, l, jp, rinv)) {
        if (live > 0) band_update(smem + mine, jp, l);
        if (rhs0 >= 0)   // more right-hand sides than lanes: the same row of every 15th one behind mine
          for (int m = RHS_LANES; rhs0 + m < R; m += RHS_LANES) band_update(smem + mine + (size_t)m * NS, jp, l);
      } else if (singular_at == 0) {
        singular_at = j + 1;
      }
      band_retire_fetch(sys, smem + leaving, j + KV + 1, lane, &leaves, &enters);
      band_retire_store(sys, smem + entering, ucol, j == 0, j + KV + 1, lane, rinv_prev, leaves, enters);
      rinv_prev = rinv;
      --live;
      colj += LD;
      if (--colj_left == 0) { colj_left = WCOLS; colj -= WCOLS * LD; }
      mine += mine_step;
      if (--mine_left == 0) { mine_left = WCOLS; mine -= WCOLS * LD; }
      leaving += LD;
      if (--leaving_left == 0) { leaving_left = WCOLS; leaving -= WCOLS * LD; }
      entering += LD;
      if (--entering_left == 0) { entering_left = WCOLS; entering -= WCOLS * LD; }
      ucol += UROWS;
      __syncwarp();
    }
    if (lane <= KV) *ucol = lane == KV ? rinv_prev : smem[leaving];
    __syncwarp();

    // ---- back substitution with the banded upper factor -------------------------------------
    // lane d-1 holds U(j-d, j) for d = 1..kv (the first 16 of them: one half-warp), lane kv the reciprocal of
    // the diagonal; fetched from the scratch two groups of four columns ahead of their use.  x_j goes
    // straight to coef[traj][piece][axis][8].
    {
      auto fetch = [&](int j) -> double {
        return (lane <= KV && j >= 0) ? __ldcg(Ug + (size_t)j * UROWS + (lane < KV ? KV - 1 - lane : KV)) : 0.0;
      };
      double nx[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) nx[i] = fetch(N - 1 - i);
      // right-hand side r = d K + k is axis k of trajectory d of the group; lane r (mod 32) keeps where its
      // coefficients go and stores x_j itself
      double* const out0 = coef + (size_t)g * G * n * K * MST_NCOEF;
      int bs_at = (int)(Bs - smem) + N - 1, ns = NS;   // row j of the first right-hand side; carried like the pointers above
      asm volatile("" : "+r"(bs_at), "+r"(ns));
      const int next_traj = (n - 1) * K * MST_NCOEF + MST_NCOEF;   // from the last axis of one trajectory to the first of the next
      double* myout = out0 + ((size_t)(lane / K) * n * K + lane % K) * MST_NCOEF;
      for (int jc = N - 1; jc >= 0; jc -= 4) {
        double cu[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { cu[i] = nx[i]; nx[i] = fetch(jc - 4 - i); }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = jc - i;   // N is a multiple of 8: never negative
          const double rinv = __shfl_sync(FULL, cu[i], KV);
          const int d = (lane < KV && lane < j) ? lane + 1 : 0;   // my row of U exists
          double* bj = smem + bs_at;
          --bs_at;
          const int colofs = (j >> 3) * K * MST_NCOEF + (j & 7);
          if (R <= 32) {
            int r = 0;
            for (; r + 3 <= R; r += 3) {   // three independent right-hand sides in flight (the axes of one drone)
              const double x0 = bj[0] * rinv, x1 = bj[ns] * rinv, x2 = bj[2 * ns] * rinv;
              if (lane == r) myout[colofs] = x0;
              if (lane == r + 1) myout[colofs] = x1;
              if (lane == r + 2) myout[colofs] = x2;
              if (d > 0 && cu[i] != 0.0) {
                const double b0 = bj[-d], b1 = bj[ns - d], b2 = bj[2 * ns - d];
                bj[-d] = b0 - cu[i] * x0;
                bj[ns - d] = b1 - cu[i] * x1;
                bj[2 * ns - d] = b2 - cu[i] * x2;
              }
              bj += 3 * ns;
            }
            for (; r < R; ++r) {
              band_backsub(bj, d, lane, r, cu[i], rinv, myout + colofs);
              bj += ns;
            }
          } else {
            double* out = out0 + colofs;
            for (int r = 0, k = 0; r < R; ++r) {
              band_backsub(bj, d, lane, 0, cu[i], rinv, out);
              bj += ns;
              out += MST_NCOEF;
              if (++k == K) { k = 0; out += next_traj - MST_NCOEF; }
            }
          }
          __syncwarp();
        }
      }
    }
    if (lane < G) info[(size_t)g * G + lane] = singular_at;
    for (int d = 32 + lane; d < G; d += 32) info[(size_t)g * G + d] = singular_at;
    __syncwarp();
  }
}

// host launcher; list/list_count (device) restrict the work to listed groups when non-null;
// scratch: banded_lu_scratch_bytes(groups, n, G*K) bytes of device memory; ticket: one device int (may be null)
int launch_banded_lu(const double* wp, const double* t, int groups, int n, int K, int G,
                     const int* list, const int* list_count, double* coef, double* dur,
                     int* info, double* scratch, int* ticket, cudaStream_t stream) {
  const BandedPlan p = banded_plan(n, G * K);
  if (p.warps == 0) return MST_ERR_TOO_LARGE;
  if (!scratch) return MST_ERR_INVALID;
  static const int variant = getenv("MST_LU_VARIANT") ? atoi(getenv("MST_LU_VARIANT")) & 1 : BANDED_VARIANT;
  const void* kernels[2] = {(const void*)banded_lu_kernel<0>, (const void*)banded_lu_kernel<1>};
  {
    const int rc = allow_dynamic_smem(kernels[variant], p.smem);
    if (rc != MST_OK) return rc;
  }
  const unsigned grid = (unsigned)banded_grid(p, groups);
  static const bool fixed_stride = getenv("MST_LU_FIXED_STRIDE") != nullptr;   // A/B
  if (fixed_stride) ticket = nullptr;
  if (ticket != nullptr) {
    cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(int), stream);
    if (e != cudaSuccess) { note_cuda_error(e); return MST_ERR_CUDA; }
  }
  switch (variant) {
    case 0: banded_lu_kernel<0><<<grid, p.warps * 32, p.smem, stream>>>(wp, t, groups, n, K, G, list, list_count, coef, dur, info, scratch, ticket); break;
    default: banded_lu_kernel<1><<<grid, p.warps * 32, p.smem, stream>>>(wp, t, groups, n, K, G, list, list_count, coef, dur, info, scratch, ticket); break;
  }
  return check_launch();
}

}  // namespace mst
