// Robust solver: the reference's own 8n x 8n system, assembled in band storage in
// shared memory and factorised by banded LU with partial pivoting, one warp per time
// group (one matrix, R = G*K right-hand sides).
//
// Replaces calculate_trajectory1D (src/optimizations/calculatingTrajectories.py:37-197)
// including its np.linalg.solve (:137).  The row layout follows :65-128 exactly
//  so pivoting sees the matrix LAPACK dgesv sees (entries t^k are formed by repeated
// multiplication here, by libm pow in the reference: equal to a few ulps, not always bitwise); the band is
// kl = 10 below / ku = 7 above the diagonal (ku = 5 when t[0] == 0, the start rows then
// being diagonal).  `A` never exists in HBM: it is built from the n durations in shared
// memory, factorised there, and only the 8 coefficients per piece and axis leave.
#include "mst_common.cuh"

namespace mst {

constexpr int KL = 10;
constexpr int KU = 7;
constexpr int KV = KL + KU;          // upper bandwidth after fill-in
constexpr int LD = 2 * KL + KU + 1;  // 28 doubles per band column (LD-1 odd: row walks are bank-conflict free)

__host__ size_t banded_lu_smem_per_warp(int n, int R) {
  const size_t N = (size_t)MST_NCOEF * n;
  return sizeof(double) * ((size_t)LD * N + (size_t)R * N + (size_t)n);
}

// element (row, col) of the band lives at AB[col*LD + KV + row - col]
__device__ __forceinline__ void band_set(double* AB, int row, int col, double v) {
  AB[col * LD + KV + row - col] = v;
}

__global__ void __launch_bounds__(512)
banded_lu_kernel(const double* __restrict__ wp, const double* __restrict__ tstamps,
                 int groups, int n, int K, int G, const int* __restrict__ list,
                 const int* __restrict__ list_count, double* __restrict__ coef,
                 double* __restrict__ dur, int* __restrict__ info) {
  extern __shared__ double smem[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int N = MST_NCOEF * n;
  const int R = G * K;
  const size_t per_warp = (size_t)LD * N + (size_t)R * N + n;
  double* AB = smem + warp * per_warp;
  double* Bs = AB + (size_t)LD * N;  // [R][N]
  double* Ts = Bs + (size_t)R * N;   // [n]

  const int todo = list ? *list_count : groups;
  for (int item = blockIdx.x * warps_per_block + warp; item < todo;
       item += gridDim.x * warps_per_block) {
    const int g = list ? list[item] : item;
    const double* tg = tstamps + (size_t)g * (n + 1);

    // ---- durations and input checks -------------------------------------------------
    const double t0 = tg[0];
    int bad = 0;
    if (!(t0 >= 0.0)) bad = isfinite(t0) ? 1 : 2;
    for (int i = lane; i < n; i += 32) {
      const double T = tg[i + 1] - tg[i];
      Ts[i] = T;
      if (!(T >= 0.0) || !isfinite(T)) bad = max(bad, isfinite(T) ? 1 : 2);   // +inf counts as non-finite input
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

    // ---- assemble [A | b] in shared memory ---------------------------------------------
    for (int e = lane; e < LD * N + R * N; e += 32) AB[e] = 0.0;  // AB and Bs are contiguous
    __syncwarp();
    {
      // first waypoint: derivatives 0..3 of piece 0 at local time t0 (quirk: t0 itself)
      const int j = lane >> 3, k = lane & 7;
      if (k >= j) band_set(AB, j, k, falling_factorial(k, j) * ipow(t0, k - j));
      // last waypoint: derivatives 0..3 of piece n-1 at T_{n-1}
      const double Tl = Ts[n - 1];
      if (k >= j)
        band_set(AB, N - 4 + j, MST_NCOEF * (n - 1) + k, falling_factorial(k, j) * ipow(Tl, k - j));
    }
    for (int e = lane; e < 64 * (n - 1); e += 32) {
      const int i = (e >> 6) + 1;  // interior waypoint
      const int rr = (e >> 3) & 7, k = e & 7;
      const int s = 4 + MST_NCOEF * (i - 1);
      const int left = MST_NCOEF * (i - 1), right = MST_NCOEF * i;
      const double T = Ts[i - 1];
      if (rr < 6) {  // derivative j = 1..6 continuity
        const int j = rr + 1;
        if (k >= j) band_set(AB, s + rr, left + k, falling_factorial(k, j) * ipow(T, k - j));
        if (k == j) band_set(AB, s + rr, right + j, -falling_factorial(j, j));
      } else if (rr == 6) {  // piece i-1 ends on the waypoint
        band_set(AB, s + 6, left + k, ipow(T, k));
      } else if (k == 0) {   // piece i starts on the waypoint
        band_set(AB, s + 7, right, 1.0);
      }
    }
    for (int e = lane; e < R * (n + 1); e += 32) {
      const int r = e / (n + 1), i = e - r * (n + 1);
      const int d = r / K, k = r - d * K;
      const double v = wp[(((size_t)g * G + d) * (n + 1) + i) * K + k];
      double* b = Bs + (size_t)r * N;
      if (i == 0) b[0] = v;
      else if (i == n) b[N - 4] = v;
      else {
        b[4 + MST_NCOEF * (i - 1) + 6] = v;
        b[4 + MST_NCOEF * (i - 1) + 7] = v;
      }
    }
    __syncwarp();

    // ---- banded LU with partial pivoting, right-hand sides carried along --------------
    // Lane l owns column j + l of the active window (l <= KV) — its rows j..j+km are contiguous
    // in the band and a stride of LD - 1 = 27 doubles apart between lanes (conflict free) — and
    // lanes KV+1.. own one right-hand side each.  EVERY lane reads column j (broadcast loads) and
    // repeats the pivot search and the multipliers in registers, so a step needs no reduction,
    // no broadcast and a single warp barrier.  (First version: pivot search by shuffle
    // reduction, swap / scale / rank-1 update as separate phases with the 150 window elements
    // spread over the lanes: 6 barriers and ~3x the instructions per step.)
    int singular_at = 0;
    for (int j = 0; j < N; ++j) {
      const int km = min(KL, N - 1 - j);
      const double* colj = AB + (size_t)j * LD + KV;
      double a[KL + 1];
#pragma unroll
      for (int r = 0; r <= KL; ++r) a[r] = r <= km ? colj[r] : 0.0;
      // every lane holds column j in registers before lane 0 swaps and rewrites it below (the CUDA
      // model does not promise lockstep between the loads above and those stores)
      __syncwarp();
      int jp = 0;
      double best = fabs(a[0]);
#pragma unroll
      for (int r = 1; r <= KL; ++r) {
        const double v = fabs(a[r]);
        if (v > best) { best = v; jp = r; }  // first maximum, as LAPACK's idamax
      }
      if (best > 0.0) {
        double piv = a[0];
#pragma unroll
        for (int r = 1; r <= KL; ++r) if (r == jp) { piv = a[r]; a[r] = a[0]; }
        const double rinv = 1.0 / piv;
        // my column of the window, or my right-hand side
        double* ptr = nullptr;
        const int c = j + lane;
        if (lane <= KV && c < N) ptr = AB + (size_t)c * LD + KV - lane;
        for (int q = lane - (KV + 1); q < R; q += 32 - (KV + 1)) {
          if (q >= 0) ptr = Bs + (size_t)q * N + j;
          if (ptr) {
            const double x0 = ptr[0], xp = ptr[jp];
            ptr[jp] = x0;   // row swap (a no-op when jp == 0)
            ptr[0] = xp;
#pragma unroll
            for (int r = 1; r <= KL; ++r)
              if (r <= km) ptr[r] = ptr[r] - (a[r] * rinv) * xp;
          }
          if (q < 0) break;  // matrix lanes have exactly one column
        }
        // U(j, j) must read back as the pivot (lane 0 just wrote xp = piv there) and the
        // multipliers below it are not kept: the right-hand sides were updated in the same step
      } else if (singular_at == 0) {
        singular_at = j + 1;
      }
      __syncwarp();
    }

    // ---- back substitution with the banded upper factor -------------------------------
    for (int j = N - 1; j >= 0; --j) {
      const double* colj = AB + (size_t)j * LD;
      const double ujj = colj[KV];
      for (int r = lane; r < R; r += 32) Bs[(size_t)r * N + j] = Bs[(size_t)r * N + j] / ujj;
      __syncwarp();
      const int kd = min(KV, j);
      for (int e = lane; e < kd * R; e += 32) {
        const int rr = e / kd, d = e - rr * kd + 1;
        Bs[(size_t)rr * N + j - d] -= colj[KV - d] * Bs[(size_t)rr * N + j];
      }
      __syncwarp();
    }

    // ---- coefficients out: coef[traj][piece][axis][8] -----------------------------------
    for (int e = lane; e < R * N; e += 32) {
      const int r = e / N, pos = e - r * N;
      const int d = r / K, k = r - d * K;
      const int i = pos >> 3, kk = pos & 7;
      coef[((((size_t)g * G + d) * n + i) * K + k) * MST_NCOEF + kk] = Bs[e];
    }
    if (lane < G) info[(size_t)g * G + lane] = singular_at;
    for (int d = 32 + lane; d < G; d += 32) info[(size_t)g * G + d] = singular_at;
    __syncwarp();
  }
}

// host launcher; list/list_count (device) restrict the work to listed groups when non-null
int launch_banded_lu(const double* wp, const double* t, int groups, int n, int K, int G,
                     const int* list, const int* list_count, double* coef, double* dur,
                     int* info, cudaStream_t stream) {
  const size_t per_warp = banded_lu_smem_per_warp(n, G * K);
  int warps = (int)(MST_MAX_SMEM / per_warp);
  if (warps < 1) return MST_ERR_TOO_LARGE;
  if (warps > 16) warps = 16;
  const size_t smem = per_warp * warps;
  {
    const int rc = allow_dynamic_smem((const void*)banded_lu_kernel, smem);
    if (rc != MST_OK) return rc;
  }
  long long blocks = ((long long)groups + warps - 1) / warps;
  const long long cap = (long long)MST_SM_COUNT * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  banded_lu_kernel<<<(unsigned)blocks, warps * 32, smem, stream>>>(wp, t, groups, n, K, G, list,
                                                                    list_count, coef, dur, info);
  return check_launch();
}

}  // namespace mst
