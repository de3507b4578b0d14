// Far-piece culling of the pipeline: while the solver still holds a piece's coefficients in
// registers it bounds the piece's positions (Bernstein coefficients over [0, T]: a polynomial stays
// inside the hull of its Bernstein coefficients) and marks the piece FAR when that bound — widened by
// a margin far above every rounding on the path — cannot bring the robot's box (sphere, when the
// robot turns) to the obstacles' root box on some axis.  The sampling kernel then never evaluates the
// samples of far pieces (their flags are 0) nor reads their coefficients.  On the benchmark 73 % of
// the pieces are far.  The flags are unchanged by construction: a culled sample is one the exact path
// would have found "not near" (pose_near_environment) too.
#pragma once
#include <math.h>

namespace mst {

// robot-extended obstacle box: piece k-axis positions below lo[k] or above hi[k] are free.
// mask[traj * 3 + k]: bit i set when piece i is far on axis k (the sampler ORs the three words).
struct FarCull {
  double lo[3], hi[3];     // K = 3: obstacle root box widened by the robot's own box (translation only)
  // K = 4 (the robot turns about z by the sampled yaw): obstacle root box, the robot's local box and bounding
  // radius; the robot's reach on x / y then depends on a bound of |yaw| over the piece (axis_reach)
  double elo[3], ehi[3], rlo[3], rhi[3], radius;   // radius: planar (x, y) radius of the robot about its z axis
  int yaw;
  unsigned* mask;
  // optional second output of the solver (null: none): the float32 polynomial matrix of path_to_pol,
  // [trajectory][piece][1 + 8K] = [T | x0..x7 | y0..y7 | ...] (scripts/drones_pols_generator.py:63-77),
  // written while the coefficients are in registers instead of by a packing pass over HBM
  float* mat;
};

__host__ __device__ constexpr double bern_weight(int i, int j) {
  // C(i, j) / C(7, j)
  double num = 1.0, den = 1.0;
  for (int a = 0; a < j; ++a) { num *= (double)(i - a); den *= (double)(7 - a); }
  return num / den;
}

// c[8] ascending monomial coefficients of one axis on local time [0, T].  (An end-point pretest — the hull
// contains p(0) and p(T), so nothing is far unless both lie on the same far side — was measured slower:
// the lanes of a warp hold different trajectories and axes, some lane nearly always needs the full hull,
// and the warp then pays for both: 1.20 ms against 1.13 ms per 1 M trajectories for the solver.)
__device__ __forceinline__ bool axis_far(const double* c, double T, double lo, double hi) {
  double s[8], p = 1.0, sumabs = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = c[j] * p; p *= T; sumabs += fabs(s[j]); }
  double bmin = s[0], bmax = s[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    double b = s[0];
#pragma unroll
    for (int j = 1; j <= i; ++j) b = fma(bern_weight(i, j), s[j], b);
    bmin = fmin(bmin, b);
    bmax = fmax(bmax, b);
  }
  // rounding of the Bernstein sums, of the sampler's Horner evaluation and of its local times is below
  // 1e-14 * sumabs; the margin is five orders of magnitude above that.  NaN compares false: not far.
  const double m = 1e-9 * (1.0 + sumabs);
  return (bmax + m < lo) || (bmin - m > hi);
}

// The same hull from the piece's END STATES (position, velocity, acceleration, jerk at local time 0 and T),
// which the condensed solver holds right before it forms the coefficients: the Bernstein coefficients of a
// degree-7 polynomial on [0, T] are
//   b0 = w0, b1 = w0 + v0 T/7, b2 = w0 + 2 v0 T/7 + a0 T^2/42, b3 = w0 + 3 v0 T/7 + 3 a0 T^2/42 + j0 T^3/210,
//   b7 = w1, b6 = w1 - v1 T/7, b5 = w1 - 2 v1 T/7 + a1 T^2/42, b4 = w1 - 3 v1 T/7 + 3 a1 T^2/42 - j1 T^3/210
// — 14 FMAs instead of the 36 + 16 of the monomial route, and no cancellation (the terms are the size of the
// motion, not of the monomial coefficients).  The evaluated polynomial is the one with the ROUNDED monomial
// coefficients piece_coefficients() forms from these states; it differs from the exact interpolant by rounding
// relative to those coefficients' terms (up to ~200 |w1 - w0|), i.e. ~1e-13 of the motion — the margin below is
// 1e-9 of the piece's scale.  NaN compares false: not far.
struct Hull8 { double b1, b2, b3, b4, b5, b6, m; };   // inner control points and the safety margin

__device__ __forceinline__ Hull8 hull_from_states(double w0, double w1, double v0, double a0, double j0, double v1,
                                                  double a1, double j1, double T) {
  const double t1 = T * (1.0 / 7.0), t2 = T * T * (1.0 / 42.0), t3 = T * T * T * (1.0 / 210.0);
  const double p1 = v0 * t1, p2 = a0 * t2, p3 = j0 * t3, q1 = v1 * t1, q2 = a1 * t2, q3 = j1 * t3;
  Hull8 h;
  h.b1 = w0 + p1; h.b2 = fma(2.0, p1, w0) + p2; h.b3 = fma(3.0, p1, w0) + fma(3.0, p2, p3);
  h.b6 = w1 - q1; h.b5 = fma(-2.0, q1, w1) + q2; h.b4 = fma(-3.0, q1, w1) + fma(3.0, q2, -q3);
  const double scale = fabs(w0) + fabs(w1) + 300.0 * fabs(w1 - w0) +
                       3.0 * (fabs(p1) + fabs(q1) + fabs(p2) + fabs(q2)) + fabs(p3) + fabs(q3);
  h.m = 1e-9 * (1.0 + scale);
  return h;
}

// bound of |value| over the piece (used on the yaw axis); NaN propagates as +inf-like "unknown"
__device__ __forceinline__ double hull_max_abs(double w0, double w1, const Hull8& h) {
  const double a = fmax(fmax(fmax(fabs(w0), fabs(h.b1)), fmax(fabs(h.b2), fabs(h.b3))),
                        fmax(fmax(fabs(h.b4), fabs(h.b5)), fmax(fabs(h.b6), fabs(w1)))) + h.m;
  return a == a ? a : 1e300;
}

// Reach of the turning robot on axis k (0: x, 1: y, 2: z) when |yaw| <= t over the piece: the world box of the
// local box [rlo, rhi] rotated about z by an angle of at most t lies within [rlo_k - d, rhi_k + d] with
// d = |own|max * t^2 / 2 + |other|max * t on x / y (1 - cos t <= t^2 / 2, |sin t| <= t), d = 0 on z — and within
// the bounding radius in any case.  The widened obstacle box follows.
__device__ __forceinline__ void axis_reach(const FarCull& c, int k, double t, double* lo, double* hi) {
  double ext_lo = c.rlo[k], ext_hi = c.rhi[k];
  if (k < 2) {
    const double own = fmax(fabs(c.rlo[k]), fabs(c.rhi[k])), oth = fmax(fabs(c.rlo[1 - k]), fabs(c.rhi[1 - k]));
    const double d = own * (0.5 * t * t) + oth * t;
    ext_lo = fmax(-c.radius, ext_lo - d);
    ext_hi = fmin(c.radius, ext_hi + d);
  }
  *lo = c.elo[k] - ext_hi;   // positions below this keep the robot before the obstacles on this axis
  *hi = c.ehi[k] - ext_lo;
}

// all eight control points on one far side (chained predicate compares: cheaper than min / max in FP64)
__device__ __forceinline__ bool hull_far(double w0, double w1, const Hull8& h, double lo, double hi) {
  const double lo2 = lo - h.m, hi2 = hi + h.m;
  const bool below = w0 < lo2 && h.b1 < lo2 && h.b2 < lo2 && h.b3 < lo2 && h.b4 < lo2 && h.b5 < lo2 && h.b6 < lo2 && w1 < lo2;
  const bool above = w0 > hi2 && h.b1 > hi2 && h.b2 > hi2 && h.b3 > hi2 && h.b4 > hi2 && h.b5 > hi2 && h.b6 > hi2 && w1 > hi2;
  return below || above;
}

__device__ __forceinline__ bool axis_far_states(double w0, double w1, double v0, double a0, double j0, double v1,
                                                double a1, double j1, double T, double lo, double hi) {
  const Hull8 h = hull_from_states(w0, w1, v0, a0, j0, v1, a1, j1, T);
  return hull_far(w0, w1, h, lo, hi);
}

}  // namespace mst
