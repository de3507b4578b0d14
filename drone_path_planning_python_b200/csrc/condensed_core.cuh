// Condensed minimum-snap solve for ONE time group, written as __host__ __device__ code so
// the same arithmetic can be unit-tested on the host (tests/hostcheck) and runs one
// group per thread on the GPU (solve_condensed.cu, pipeline_fused.cu).
//
// Formulation.  The reference's square system (interpolation + C^1..C^6 continuity +
// rest-to-rest ends, src/optimizations/calculatingTrajectories.py:48-131) is the
// first-order optimality system of the minimum-snap QP with the waypoint positions fixed
// and velocity / acceleration / jerk x_i = (v_i, a_i, j_i) free at the interior knots.
// In the Hermite basis of each piece (end-point position, velocity, acceleration, jerk)
// the snap cost  int_0^T (p'''')^2 dt  is  rho^7 * u^T Hbar u  with rho = 1/T,
// u = (w0, T v0, T^2 a0, T^3 j0, w1, T v1, T^2 a1, T^3 j1) and Hbar the constant
// integer matrix below (derived symbolically: Hbar = M^-T Qbar M^-1, Qbar the snap
// Hessian of the monomial basis on [0,1], M the Hermite collocation matrix).  Setting the
// gradient w.r.t. the interior x_i to zero gives a symmetric positive definite
// block-tridiagonal system with 3x3 blocks:
//     B_{i-1}^T x_{i-1} + A_i x_i + B_i x_{i+1} = r_i ,   x_0 = x_n = 0
//     A_i[j][k] = HEE[j][k] rho_{i-1}^(7-j-k) + HSS[j][k] rho_i^(7-j-k)      j,k in 1..3
//     B_i[j][k] = HSE[j][k] rho_i^(7-j-k)
//     r_i[j]    = gE[j] rho_{i-1}^(7-j) (w_i - w_{i-1}) + gS[j] rho_i^(7-j) (w_{i+1} - w_i)
// solved by block LDL^T (no pivoting needed: SPD), after which the 8 monomial
// coefficients of every piece follow from its end-point derivatives in closed form.
// Both the Hessian and the right-hand side only see waypoint DIFFERENCES, so a large
// common offset costs no accuracy.
//
// Accuracy.  Against the reference's pivoted dense solve the coefficients agree to
// <= 1e-10 (normwise, per axis) when max(T)/min(T) <= 4 inside the group and degrade
// roughly like (max T / min T)^6 beyond that (stiff short pieces next to long ones);
// the dispatcher therefore sends wider duration spreads to the banded pivoted-LU kernel,
// which reproduces the reference to ~1e-14 for any spread.  See DESIGN.md §Numerics.
#pragma once
#include "mst_common.cuh"

namespace mst {

// Which time groups the condensed solver takes: max T / min T <= 4 (anything else goes to the pivoted
// banded LU).  A wider rule (spread <= 12 with min T >= 0.25 s) was measured in round 2
// (profiles/r2_condensed_spread.md): normwise it stays below 7.1e-11 of the reference's coefficients, but
// the C^6 continuity residual of the SHORT pieces — small next to the trajectory's largest coefficients,
// so invisible normwise — grows past what the reference's own solution shows (tests/test_gpu_fullsize.py
// ::_check_system), so the rule was not adopted.
#define MST_CONDENSED_MAX_SPREAD 4.0

// scratch slots per group: rho[n] | factors 6*(n-1) | y 3*K*(n-1)
__host__ __device__ __forceinline__ int condensed_slots(int n, int K) {
  return n + (6 + 3 * K) * (n - 1);
}

struct RhoPow { double p1, p2, p3, p4, p5, p6; };
__host__ __device__ __forceinline__ RhoPow rho_powers(double rho) {
  RhoPow r;
  r.p1 = rho; r.p2 = rho * rho; r.p3 = r.p2 * rho; r.p4 = r.p2 * r.p2; r.p5 = r.p4 * rho;
  r.p6 = r.p3 * r.p3;
  return r;
}

struct Ldl3 { double l21, l31, l32, i1, i2, i3; };

__host__ __device__ __forceinline__ Ldl3 ldl3_factor(double s11, double s21, double s22, double s31,
                                                     double s32, double s33) {
  Ldl3 f;
  f.i1 = 1.0 / s11;
  f.l21 = s21 * f.i1;
  f.l31 = s31 * f.i1;
  const double d2 = s22 - f.l21 * s21;
  f.i2 = 1.0 / d2;
  const double t32 = s32 - f.l31 * s21;
  f.l32 = t32 * f.i2;
  const double d3 = s33 - f.l31 * s31 - f.l32 * t32;
  f.i3 = 1.0 / d3;
  return f;
}

__host__ __device__ __forceinline__ void ldl3_solve(const Ldl3& f, double b1, double b2, double b3,
                                                    double& x1, double& x2, double& x3) {
  const double f2 = b2 - f.l21 * b1;
  const double f3 = b3 - f.l31 * b1 - f.l32 * f2;
  x3 = f3 * f.i3;
  x2 = f2 * f.i2 - f.l32 * x3;
  x1 = b1 * f.i1 - f.l21 * x2 - f.l31 * x3;
}

// B = HSE o rho^(7-j-k), HSE = [[24480,-4680,360],[4680,-840,60],[360,-60,4]]
struct Mat3 { double m11, m12, m13, m21, m22, m23, m31, m32, m33; };
__host__ __device__ __forceinline__ Mat3 coupling_block(const RhoPow& r) {
  Mat3 b;
  b.m11 = 24480.0 * r.p5; b.m12 = -4680.0 * r.p4; b.m13 = 360.0 * r.p3;
  b.m21 = 4680.0 * r.p4;  b.m22 = -840.0 * r.p3;  b.m23 = 60.0 * r.p2;
  b.m31 = 360.0 * r.p3;   b.m32 = -60.0 * r.p2;   b.m33 = 4.0 * r.p1;
  return b;
}

// phase 1: block LDL^T factors.  rho[i * stride] (i < n) holds the duration T_i (> 0) on entry
// and rho_i = 1/T_i on exit; the six factor terms of knot i go to fac[((i-1)*6 + j) * stride].
__host__ __device__ inline void condensed_factor(int n, double* rho, double* fac, int stride) {
  double* scratch = rho;
  RhoPow pa = rho_powers(1.0 / scratch[0]);
  scratch[0] = pa.p1;
  // Schur correction C = B_{i-1}^T S_{i-1}^{-1} B_{i-1} carried between knots (symmetric)
  double c11 = 0, c21 = 0, c22 = 0, c31 = 0, c32 = 0, c33 = 0;
  for (int i = 1; i < n; ++i) {
    const RhoPow pb = rho_powers(1.0 / scratch[(size_t)i * stride]);
    scratch[(size_t)i * stride] = pb.p1;
    // S_i = A_i - C,  A_i = HEE o pa + HSS o pb,
    // HEE/HSS = [[25920,-/+5400,480],[-/+5400,1200,-/+120],[480,-/+120,16]]
    const Ldl3 f = ldl3_factor(25920.0 * (pa.p5 + pb.p5) - c11, 5400.0 * (pb.p4 - pa.p4) - c21,
                               1200.0 * (pa.p3 + pb.p3) - c22, 480.0 * (pa.p3 + pb.p3) - c31,
                               120.0 * (pb.p2 - pa.p2) - c32, 16.0 * (pa.p1 + pb.p1) - c33);
    double* fo = fac + (size_t)(i - 1) * 6 * stride;
    fo[0] = f.l21; fo[stride] = f.l31; fo[2 * (size_t)stride] = f.l32;
    fo[3 * (size_t)stride] = f.i1; fo[4 * (size_t)stride] = f.i2; fo[5 * (size_t)stride] = f.i3;
    if (i + 1 < n) {
      const Mat3 b = coupling_block(pb);  // B_i
      double w11, w21, w31, w12, w22, w32, w13, w23, w33;  // W = S_i^{-1} B_i, column by column
      ldl3_solve(f, b.m11, b.m21, b.m31, w11, w21, w31);
      ldl3_solve(f, b.m12, b.m22, b.m32, w12, w22, w32);
      ldl3_solve(f, b.m13, b.m23, b.m33, w13, w23, w33);
      // C = B_i^T W (symmetric; lower triangle)
      c11 = b.m11 * w11 + b.m21 * w21 + b.m31 * w31;
      c21 = b.m12 * w11 + b.m22 * w21 + b.m32 * w31;
      c22 = b.m12 * w12 + b.m22 * w22 + b.m32 * w32;
      c31 = b.m13 * w11 + b.m23 * w21 + b.m33 * w31;
      c32 = b.m13 * w12 + b.m23 * w22 + b.m33 * w32;
      c33 = b.m13 * w13 + b.m23 * w23 + b.m33 * w33;
    }
    pa = pb;
  }
}

__host__ __device__ __forceinline__ Ldl3 load_factor(const double* fac, int knot, int stride) {
  const double* fo = fac + (size_t)(knot - 1) * 6 * stride;
  Ldl3 f;
  f.l21 = fo[0]; f.l31 = fo[stride]; f.l32 = fo[2 * (size_t)stride];
  f.i1 = fo[3 * (size_t)stride]; f.i2 = fo[4 * (size_t)stride]; f.i3 = fo[5 * (size_t)stride];
  return f;
}

// the 8 ascending-power coefficients of one piece from its end states
//   (w0, v0, a0, j0) at local time 0 and (w0 + dw, v1, a1, j1) at local time T = 1/rho
__host__ __device__ __forceinline__ void piece_coefficients(double w0, double dw, double v0, double a0,
                                                            double j0, double v1, double a1, double j1,
                                                            double rho, double* c) {
  const double third2 = 2.0 / 3.0, sixth = 1.0 / 6.0;
  c[0] = w0;
  c[1] = v0;
  c[2] = 0.5 * a0;
  c[3] = sixth * j0;
  c[4] = rho * (rho * (rho * (rho * (35.0 * dw) - (20.0 * v0 + 15.0 * v1)) - (5.0 * a0 - 2.5 * a1)) -
                (third2 * j0 + sixth * j1));
  c[5] = rho * rho * (rho * (rho * (rho * (-84.0 * dw) + (45.0 * v0 + 39.0 * v1)) + (10.0 * a0 - 7.0 * a1)) +
                      (j0 + 0.5 * j1));
  const double r3 = rho * rho * rho;
  c[6] = r3 * (rho * (rho * (rho * (70.0 * dw) - (36.0 * v0 + 34.0 * v1)) - (7.5 * a0 - 6.5 * a1)) -
               (third2 * j0 + 0.5 * j1));
  c[7] = r3 * rho * (rho * (rho * (rho * (-20.0 * dw) + 10.0 * (v0 + v1)) + 2.0 * (a0 - a1)) +
                     sixth * (j0 + j1));
}

// phase 2 (per trajectory of the group): forward elimination of the K right-hand sides.
// wp points at this trajectory's [n+1][K] waypoints.  y is stored for the back sweep.
// wp[i * wstride + k]: waypoint i of column k (k < K columns handled by this thread);
// rho / fac as written by condensed_factor (stride fstride); y values go to
// ys[((i-1)*3*K + 3*k + j) * ystride].
template <int KC>
__host__ __device__ inline void condensed_forward(const double* __restrict__ wp, int wstride, int n, int K,
                                                  const double* rho, const double* fac, int fstride,
                                                  double* ys, int ystride) {
  const int stride = fstride;
  double wprev[KC], da[KC], z1[KC], z2[KC], z3[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    if (k < K) {
      const double w0 = wp[k], w1 = wp[wstride + k];
      da[k] = w1 - w0;
      wprev[k] = w1;
    }
    z1[k] = z2[k] = z3[k] = 0.0;
  }
  RhoPow pa = rho_powers(rho[0]);
  for (int i = 1; i < n; ++i) {
    const RhoPow pb = rho_powers(rho[(size_t)i * stride]);
    const Ldl3 f = load_factor(fac, i, stride);
    const Mat3 bp = coupling_block(pa);  // B_{i-1}
    const double ge1 = 50400.0 * pa.p6, ge2 = -10080.0 * pa.p5, ge3 = 840.0 * pa.p4;
    const double gs1 = 50400.0 * pb.p6, gs2 = 10080.0 * pb.p5, gs3 = 840.0 * pb.p4;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      if (k < K) {
        const double wn = wp[(size_t)(i + 1) * wstride + k];
        const double db = wn - wprev[k];
        wprev[k] = wn;
        // y_i = r_i - B_{i-1}^T z_{i-1}
        const double y1 = ge1 * da[k] + gs1 * db - (bp.m11 * z1[k] + bp.m21 * z2[k] + bp.m31 * z3[k]);
        const double y2 = ge2 * da[k] + gs2 * db - (bp.m12 * z1[k] + bp.m22 * z2[k] + bp.m32 * z3[k]);
        const double y3 = ge3 * da[k] + gs3 * db - (bp.m13 * z1[k] + bp.m23 * z2[k] + bp.m33 * z3[k]);
        double* yo = ys + ((size_t)(i - 1) * 3 * K + 3 * k) * ystride;
        yo[0] = y1; yo[ystride] = y2; yo[2 * (size_t)ystride] = y3;
        ldl3_solve(f, y1, y2, y3, z1[k], z2[k], z3[k]);
        da[k] = db;
      }
    }
    pa = pb;
  }
}

// phase 3: back sweep knot n-1 .. 1; after knot i is known piece i is complete and handed
// to `emit(piece, k, c[8], rho_i)`; piece 0 last.  (Pieces therefore arrive in DESCENDING order.)
// `ends` (optional): called as ends(piece, k, w0, w1, v0, a0, j0, v1, a1, j1, rho) with the piece's two end states
// right before its coefficients are formed (the pipeline's far-piece bound works from these).
struct NoEndStates {
  __host__ __device__ void operator()(int, int, double, double, double, double, double, double, double, double, double) const {}
};

template <int KC, class Emit, class Ends = NoEndStates>
__host__ __device__ inline void condensed_backward(const double* __restrict__ wp, int wstride, int n, int K,
                                                   const double* rho, const double* fac, int fstride,
                                                   const double* ys, int ystride, Emit&& emit, Ends&& ends = Ends()) {
  const int stride = fstride;
  double xv[KC], xa[KC], xj[KC];  // state at knot i+1
#pragma unroll
  for (int k = 0; k < KC; ++k) xv[k] = xa[k] = xj[k] = 0.0;
  for (int i = n - 1; i >= 0; --i) {
    const double rh = rho[(size_t)i * stride];
    double nv[KC], na[KC], nj[KC];  // state at knot i
    if (i >= 1) {
      const RhoPow pb = rho_powers(rh);
      const Mat3 b = coupling_block(pb);  // B_i
      const Ldl3 f = load_factor(fac, i, stride);
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        if (k < K) {
          const double* yo = ys + ((size_t)(i - 1) * 3 * K + 3 * k) * ystride;
          const double b1 = yo[0] - (b.m11 * xv[k] + b.m12 * xa[k] + b.m13 * xj[k]);
          const double b2 = yo[ystride] - (b.m21 * xv[k] + b.m22 * xa[k] + b.m23 * xj[k]);
          const double b3 = yo[2 * (size_t)ystride] - (b.m31 * xv[k] + b.m32 * xa[k] + b.m33 * xj[k]);
          ldl3_solve(f, b1, b2, b3, nv[k], na[k], nj[k]);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < KC; ++k) nv[k] = na[k] = nj[k] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      if (k < K) {
        const double w0 = wp[(size_t)i * wstride + k], w1 = wp[(size_t)(i + 1) * wstride + k];
        double c[MST_NCOEF];
        ends(i, k, w0, w1, nv[k], na[k], nj[k], xv[k], xa[k], xj[k], rh);
        piece_coefficients(w0, w1 - w0, nv[k], na[k], nj[k], xv[k], xa[k], xj[k], rh, c);
        emit(i, k, c, rh);
        xv[k] = nv[k]; xa[k] = na[k]; xj[k] = nj[k];
      }
    }
  }
}

// phase 3 without the coefficients: back sweep knot n-1 .. 1 that leaves the knot state
// (velocity, acceleration, jerk) of column k at knot i IN PLACE of y_i, i.e. at
// ys[((i-1)*3*K + 3*k + j) * ystride].  Same arithmetic as the first half of condensed_backward;
// the single-pass pipeline (pipeline_onepass.cu) keeps these 3 values per knot and column on chip
// — a third of the 8 coefficients per piece — and forms the coefficients of a trajectory with
// piece_coefficients() right before its samples are evaluated and its rows are stored.
template <int KC>
__host__ __device__ inline void condensed_backward_states(int n, int K, const double* rho, const double* fac,
                                                          int fstride, double* ys, int ystride) {
  const int stride = fstride;
  double xv[KC], xa[KC], xj[KC];  // state at knot i+1
#pragma unroll
  for (int k = 0; k < KC; ++k) xv[k] = xa[k] = xj[k] = 0.0;
  for (int i = n - 1; i >= 1; --i) {
    const double rh = rho[(size_t)i * stride];
    const RhoPow pb = rho_powers(rh);
    const Mat3 b = coupling_block(pb);  // B_i
    const Ldl3 f = load_factor(fac, i, stride);
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      if (k < K) {
        double* yo = ys + ((size_t)(i - 1) * 3 * K + 3 * k) * ystride;
        const double b1 = yo[0] - (b.m11 * xv[k] + b.m12 * xa[k] + b.m13 * xj[k]);
        const double b2 = yo[ystride] - (b.m21 * xv[k] + b.m22 * xa[k] + b.m23 * xj[k]);
        const double b3 = yo[2 * (size_t)ystride] - (b.m31 * xv[k] + b.m32 * xa[k] + b.m33 * xj[k]);
        double nv, na, nj;
        ldl3_solve(f, b1, b2, b3, nv, na, nj);
        yo[0] = nv; yo[ystride] = na; yo[2 * (size_t)ystride] = nj;
        xv[k] = nv; xa[k] = na; xj[k] = nj;
      }
    }
  }
}

// classification of a time group; returns 0 = condensed path may run,
// 1 = decline (needs the pivoted solver), 2 = decreasing times, 3 = non-finite times
__host__ __device__ inline int classify_times(const double* tg, int n, double* Tmin_out, double* Tmax_out) {
  const double t0 = tg[0];
  int bad = 0;
  if (!(t0 >= 0.0)) bad = (t0 == t0 && t0 - t0 == 0.0) ? 2 : 3;
  double Tmin = 1e300, Tmax = 0.0;
  double prev = t0;
  for (int i = 0; i < n; ++i) {
    const double nx = tg[i + 1];
    const double T = nx - prev;
    prev = nx;
    if (!(T >= 0.0)) { const int b = (T == T && T - T == 0.0) ? 2 : 3; if (b > bad) bad = b; }
    if (T < Tmin) Tmin = T;
    if (T > Tmax) Tmax = T;
  }
  *Tmin_out = Tmin;
  *Tmax_out = Tmax;
  if (bad) return bad;
  if (!(Tmax - Tmax == 0.0)) return 3;
  if (t0 != 0.0 || !(Tmin > 0.0) || Tmax > MST_CONDENSED_MAX_SPREAD * Tmin) return 1;
  return 0;
}

// Walk over the elements e, e + 32, e + 64, ... of a contiguous waypoint slice [trajectory][waypoint 0..n][axis
// 0..K-1] (what one lane of a warp copies): (trajectory, waypoint, axis) advance by carries instead of two
// divisions per element.  slot(): where the element goes in a waypoint-major tile of row stride WS whose columns
// are (trajectory, axis) pairs — the layout the lane-per-column solver reads without bank conflicts.
struct TileWalk {
  int tt, i, k;
  __host__ __device__ static TileWalk start(int e, int n, int K) {
    TileWalk w;
    w.tt = e / ((n + 1) * K);
    const int rem = e - w.tt * (n + 1) * K;
    w.i = rem / K;
    w.k = rem - w.i * K;
    return w;
  }
  __host__ __device__ int slot(int WS, int K) const { return i * WS + tt * K + k; }
  __host__ __device__ void advance32(int n, int K) {
    k += 32 % K;
    i += 32 / K;
    if (k >= K) { k -= K; ++i; }
    while (i > n) { i -= n + 1; ++tt; }
  }
};

}  // namespace mst
