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
  double lo[3], hi[3];
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

}  // namespace mst
