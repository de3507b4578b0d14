// Arithmetic of the pivoted banded solver (solve_banded_lu.cu), one lane's share at a time, as
// __host__ __device__ functions: the kernel drives them with a warp, tests/hostcheck drives the
// same functions lane by lane on the host (CPU tier of the tests).
//
// The system is the reference's own 8n x 8n matrix (src/optimizations/calculatingTrajectories.py:65-128,
// SURVEY §8 a2) in LAPACK band storage: kl = 10 below / ku = 7 above the diagonal, kv = kl + ku above
// it after fill-in, element (row, col) at band position o = kv + row - col of its column.  Only a
// WINDOW of the band lives in shared memory: at elimination step j the columns j .. j+kv are being
// updated, column j-1 has just become a finished column of U and the columns from j+kv+1 on have not
// been touched yet.  So the window is a ring of kv + 3 column slots; a finished column goes out to a
// per-warp scratch in device memory (18 doubles, read back once by the back substitution — the
// scratch of all resident warps is ~50-100 MB, i.e. L2-resident), and the slot it leaves is refilled
// with the next untouched column, whose entries are recomputed from the table of duration powers
// (band_entry) instead of being stored.  Shared memory per warp drops from 28 x 8n doubles to
// 20 x 28, i.e. from 5 resident warps per SM at n = 20 to 20.  What bounds the kernel is the shared-memory
// data pipe (90 % of its wavefronts busy, profiles/): the layout rules below are about wavefronts.
#pragma once
#include <string.h>

#include "mst_common.cuh"

namespace mst {

constexpr int KL = 10;
constexpr int KU = 7;
constexpr int KV = KL + KU;          // upper bandwidth after fill-in
constexpr int LD = 2 * KL + KU + 1;  // 28 band positions per column
constexpr int WCOLS = KV + 3;        // column slots of the window ring: kv + 1 in use, one being refilled, and one
                                     // more so that WCOLS * LD is a multiple of 16 doubles — a column keeps its
                                     // shared-memory banks when the ring wraps
constexpr int MAT_LANES = KV;        // lane l < MAT_LANES owns column j + 1 + l of the window: lanes 0..15 (one
                                     // half-warp = one 128-byte wavefront per 64-bit access) are 27 doubles apart,
                                     // 16 different bank pairs; column j + kv and the right-hand sides share the
                                     // other half-warp
constexpr int RHS_LANES = 32 - MAT_LANES;
constexpr int UROWS = KV + 1;        // doubles per finished column of U (rows col-kv .. col)

struct BandSystem {
  int n, N;          // pieces, unknowns (8n)
  const double* pw;  // [n+1][8]: pw[i][m] = T_i^m by repeated multiplication (ipow); row n holds t[0]^m
  const double* ff;  // [8][8]:   ff[k][j] = k (k-1) ... (k-j+1)
  const double* cf;  // [8][LD]:  band_pattern() of an interior piece's columns, the factor ...
  const unsigned char* pi;  // [8][LD]: ... and the power of the piece's duration it multiplies
};

constexpr int BAND_TABLE_DOUBLES = 64 + MST_NCOEF * LD + MST_NCOEF * LD / 8;   // ff, cf, pi (bytes)

// pw row of one piece (or of t[0]): the same sequence of products as ipow()
__host__ __device__ __forceinline__ void band_powers(double T, double* row) {
  double p = 1.0;
  for (int m = 0; m < MST_NCOEF; ++m) { row[m] = p; p *= T; }
}

// Entry of the reference's matrix at band position o of column col (row = col - KV + o); col < N.
// Column col = 8p + k belongs to coefficient k of piece p.  It meets
//   the rows of the waypoint AFTER piece p  (:86-120, piece p is the left piece): derivative j = 1..6
//     continuity k!/(k-j)! T^(k-j), then "piece p ends on the waypoint" T^k;  for the last piece the
//     final-waypoint rows (:121-128), derivatives 0..3 at T;
//   the rows of the waypoint BEFORE piece p (piece p is the right piece): -j! on coefficient j of
//     continuity row j, 1 on coefficient 0 of "piece p starts on the waypoint";  for piece 0 the
//     first-waypoint rows (:65-85), derivatives 0..3 at local time t[0] (the reference's quirk).
__host__ __device__ __forceinline__ double band_entry(const BandSystem& s, int col, int o) {
  const int p = col >> 3, k = col & 7;
  // derivative order jd of the row this position belongs to, and the piece (or t[0]) whose powers it takes
  int jd = -1, prow = p;
  const int ra = o + k - 21;
  if (p < s.n - 1) {
    if (ra >= 0 && ra < 6) jd = ra + 1;
    else if (ra == 6) jd = 0;
  } else if (ra >= 0 && ra < 4) {
    jd = ra;
  }
  if (p == 0) {
    const int js = o + k - 17;
    if (js >= 0 && js < 4) { jd = js; prow = s.n; }
  }
  double v = 0.0;
  if (jd >= 0 && jd <= k) v = s.ff[k * 8 + jd] * s.pw[prow * 8 + k - jd];   // jd == 0: 1.0 * T^k, exact
  // right piece of the waypoint before: the -j! of continuity row j = k sits at band position 12, the 1 of
  // "piece p starts on the waypoint" at band position 20 of coefficient 0
  if (p > 0) {
    if (o == 12 && k >= 1 && k <= 6) v = -s.ff[k * 8 + k];
    if (o == 20 && k == 0) v = 1.0;
  }
  return v;
}

// Column 8p + k of a piece that is neither the first nor the last has the same pattern for every p: the
// entry at band position o is cf * T_p^pi (the constants -j! and 1 take T^0 = 1.0, exact; zeros take
// cf = 0).  Same rules as band_entry with 0 < p < n - 1.
__host__ __device__ __forceinline__ void band_pattern(int k, int o, double* cf, unsigned char* pi) {
  int jd = -1;
  const int ra = o + k - 21;
  if (ra >= 0 && ra < 6) jd = ra + 1;
  else if (ra == 6) jd = 0;
  double c = 0.0;
  int e = 0;
  if (jd >= 0 && jd <= k) { c = falling_factorial(k, jd); e = k - jd; }
  if (o == 12 && k >= 1 && k <= 6) { c = -falling_factorial(k, k); e = 0; }
  if (o == 20 && k == 0) { c = 1.0; e = 0; }
  *cf = c;
  *pi = (unsigned char)e;
}

// the three tables a CTA (or the host check) builds once: entry e of 64 + 8 LD
__host__ __device__ __forceinline__ void band_table_entry(int e, double* ff, double* cf, unsigned char* pi) {
  if (e < 64) ff[e] = falling_factorial(e >> 3, e & 7);
  else band_pattern((e - 64) / LD, (e - 64) % LD, cf + (e - 64), pi + (e - 64));
}

// band_entry through the pattern tables where they apply
__host__ __device__ __forceinline__ double band_entry_fast(const BandSystem& s, int col, int o) {
  const int p = col >> 3, k = col & 7;
  if (p == 0 || p >= s.n - 1) return band_entry(s, col, o);
  return s.cf[k * LD + o] * s.pw[p * 8 + s.pi[k * LD + o]];
}

// Step j, the part every lane repeats in registers: column j from its diagonal down (colj[0..kl]; the
// rows past the end of the matrix hold zeros), the pivot as LAPACK's idamax picks it (first maximum),
// the multipliers l_r = a_r / pivot formed as dgbtf2 does (reciprocal, then scale) — for the rows in
// their places BEFORE the swap: band_update puts the one row the swap moves right.  Returns false for
// a zero (or NaN) pivot column; rinv is then 1 / U(j,j), what a division by that diagonal multiplies with.
// one node of the pivot tournament: the left candidate (lower rows) stays unless the right one is strictly
// larger in magnitude — the first maximum wins, as in idamax
__host__ __device__ __forceinline__ void band_pick(double& v, int& i, double w, int iw) {
  if (fabs(w) > fabs(v)) { v = w; i = iw; }
}

// magnitude of a double as an integer key: |x| < |y|  <=>  key(x) < key(y) for everything but NaN
__host__ __device__ __forceinline__ unsigned long long band_key(double x) {
  unsigned long long u;
  memcpy(&u, &x, sizeof(u));
  return u & 0x7fffffffffffffffull;
}

// The pivot of step j: the first row of largest magnitude, as idamax picks it.  Two searches:
//   WARP = false  every lane runs the 10 comparisons itself, as a tree of depth 4 instead of a chain of 10
//                 (the search is on the step's critical path);
//   WARP = true   (device only; all 32 lanes must call) lane r offers row r and two integer warp reductions
//                 — high word of the magnitude bits, then low word among the lanes holding the largest high
//                 word — find the largest; the lowest such lane wins.  50 instructions fewer per step.
template <bool WARP>
__host__ __device__ __forceinline__ bool band_pivot(const double* colj, int lane, double (&l)[KL + 1], int& jp,
                                                    double& rinv) {
#pragma unroll
  for (int r = 0; r <= KL; ++r) l[r] = colj[r];
  double piv;
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 800
  if (WARP) {
    const double own = lane <= KL ? colj[lane] : 0.0;
    const unsigned hi = (unsigned)__double2hiint(own) & 0x7fffffffu;
    const unsigned top_hi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned lo = hi == top_hi ? (unsigned)__double2loint(own) : 0u;
    const unsigned top_lo = __reduce_max_sync(0xffffffffu, lo);
    jp = __ffs(__ballot_sync(0xffffffffu, hi == top_hi && lo == top_lo)) - 1;
    piv = __shfl_sync(0xffffffffu, own, jp);
  } else
#endif
  {
    (void)lane;
    double v0 = l[0], v2 = l[2], v4 = l[4], v6 = l[6], v8 = l[8];
    int i0 = 0, i2 = 2, i4 = 4, i6 = 6, i8 = 8;
    band_pick(v0, i0, l[1], 1);
    band_pick(v2, i2, l[3], 3);
    band_pick(v4, i4, l[5], 5);
    band_pick(v6, i6, l[7], 7);
    band_pick(v8, i8, l[9], 9);
    band_pick(v0, i0, v2, i2);
    band_pick(v4, i4, v6, i6);
    band_pick(v8, i8, l[10], 10);
    band_pick(v0, i0, v4, i4);
    band_pick(v0, i0, v8, i8);
    piv = v0;
    jp = i0;
  }
  rinv = 1.0 / piv;
  if (!(fabs(piv) > 0.0)) return false;
#pragma unroll
  for (int r = 0; r <= KL; ++r) l[r] *= rinv;
  return true;
}

// Step j, one lane's column of the window (or one right-hand side): ptr points at its row j, x0 / xp are
// its entries in row j and in the pivot row.  Rank-1 update of the rows below with the multipliers of the
// unswapped rows, then the row swap: the pivot row's entry goes up to row j, and row j's old entry lands in
// row j+jp updated with ITS multiplier (l[0]).  No row guard: past the end of the matrix the multipliers
// are zero and the storage (band slots, padded right-hand sides) is there.
__host__ __device__ __forceinline__ void band_update(double* ptr, int jp, const double (&l)[KL + 1]) {
  const double x0 = ptr[0], xp = ptr[jp];
#pragma unroll
  for (int r = 1; r <= KL; ++r) ptr[r] = ptr[r] - l[r] * xp;
  if (jp) ptr[jp] = x0 - l[0] * xp;
  ptr[0] = xp;
}

// Step j, lane `lane` < LD: the finished column j-1 leaves its slot for the scratch (ucol, null at
// j = 0; its diagonal is kept as the reciprocal the back substitution multiplies with) and column
// newcol = j + kv + 1 takes the slot.  Lane l reads band position l and then writes band position l:
// no other lane touches this slot during the step.
// (slot and ucol already point at this lane's band position; first: j == 0, nothing leaves yet.)  In two
// halves, loads then stores (issuing the loads at the top of the step instead changed nothing).
__host__ __device__ __forceinline__ void band_retire_fetch(const BandSystem& s, const double* slot, int newcol,
                                                           int lane, double* leaves, double* enters) {
  *leaves = 0.0;
  *enters = 0.0;
  if (lane >= LD) return;
  *leaves = *slot;
  if (newcol < s.N) *enters = band_entry_fast(s, newcol, lane);
}

// (slot: this lane's band position of the slot column newcol goes to — the one column j-2 left a step ago.)
__host__ __device__ __forceinline__ void band_retire_store(const BandSystem& s, double* slot, double* ucol,
                                                           bool first, int newcol, int lane, double rinv_prev,
                                                           double leaves, double enters) {
  if (lane >= LD) return;
  if (!first && lane <= KV) *ucol = lane == KV ? rinv_prev : leaves;
  if (newcol < s.N) *slot = enters;
}

// Back substitution, column j, one right-hand side b, lane d: every lane forms x_j = b_j / U(j,j)
// itself (rinv = the stored reciprocal), one lane sends it to its final place, lane d in 1..kv removes
// U(j-d, j) x_j from b[j-d].
// (bj points at b[j]; this lane holds u = U(j-d, j), d = 0 meaning none; lane `writer` is the one that stores
// x_j, through ITS out.  A zero u — the fill-in never reached that far — costs no shared-memory access.)
__host__ __device__ __forceinline__ void band_backsub(double* bj, int d, int lane, int writer, double u,
                                                      double rinv, double* out) {
  const double x = bj[0] * rinv;
  if (lane == writer) *out = x;
  if (d > 0 && u != 0.0) bj[-d] -= u * x;
}

// right-hand sides in shared memory: 8n entries and kl + 1 of slack for the unguarded updates of the last
// steps; an odd stride, so that the lanes (one right-hand side each, same row) fall on different banks
__host__ __device__ __forceinline__ int band_rhs_stride(int n) { return MST_NCOEF * n + KL + 1; }

// doubles of shared memory one warp needs: window ring + power table + right-hand sides
__host__ __device__ __forceinline__ size_t band_warp_doubles(int n, int R) {
  return (size_t)WCOLS * LD + (size_t)MST_NCOEF * (n + 1) + (size_t)R * band_rhs_stride(n);
}

}  // namespace mst
