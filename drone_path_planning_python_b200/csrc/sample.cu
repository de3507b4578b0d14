// Batched polynomial evaluation: Polynomial.eval / .derivative, PiecewisePolynomial.eval,
// Trajectory.eval and Polynomial4D.eval (src/optimizations/uav_trajectory.py:17-26,66-101,
// 119-127,154-169), plus the time-power rows of Polynomial.pol_coeffs_at_t (:28-36).
//
// One thread per (trajectory, sample).  The arithmetic mirrors the reference operation
// for operation — Horner with separately rounded multiply and add (Python floats have no
// FMA), derivative coefficients by the same chain of integer multiplications, the piece
// search with the same left-to-right running sum — so values are bit-identical to the
// reference for in-range samples.
#include "mst_common.cuh"

namespace mst {

// x^e (small e >= 0) rounded once: the power is carried as an unevaluated sum hi + lo
// (FMA error-free products), so the result is the correctly rounded power in all but
// astronomically rare cases — what Python's float ** int (libm pow) returns.
__device__ __forceinline__ double pow_rounded_once(double x, int e) {
  double hi = 1.0, lo = 0.0;
  for (int i = 0; i < e; ++i) {
    const double p = __dmul_rn(hi, x);
    const double err = __fma_rn(hi, x, -p);
    const double l = __fma_rn(lo, x, err);
    hi = __dadd_rn(p, l);
    lo = __dadd_rn(__dsub_rn(p, hi), l);
  }
  return hi;
}

// one Polynomial.derivative step on c[0..len-1]: c[i] = (i+1) * c[i+1]; returns len-1
__device__ __forceinline__ int derive_once(double* c, int len) {
  for (int i = 0; i + 1 < len; ++i) c[i] = __dmul_rn((double)(i + 1), c[i + 1]);
  return len > 0 ? len - 1 : 0;
}

// Polynomial.derivative applied `deriv` times to c[0..7]; returns remaining length
__device__ __forceinline__ int derive_in_place(double* c, int deriv) {
  int len = MST_NCOEF;
  for (int d = 0; d < deriv; ++d) len = derive_once(c, len);
  return len;
}

// Polynomial.eval: x = x*t + p[len-1-i], multiply and add rounded separately
__device__ __forceinline__ double horner_nofma(const double* c, int len, double t) {
  double x = 0.0;
  for (int i = len - 1; i >= 0; --i) x = __dadd_rn(__dmul_rn(x, t), c[i]);
  return x;
}

// piece index and local time; returns false when the reference would assert / fall through
__device__ __forceinline__ bool find_piece(const double* __restrict__ T, int n, double t, int mode,
                                           int* piece, double* local) {
  if (!(t >= 0.0)) return false;
  double acc = 0.0;
  if (mode == MST_SAMPLE_PIECEWISE) {
    for (int i = 0; i < n; ++i) {
      const double Ti = T[i];
      if (t < __dadd_rn(acc, Ti)) { *piece = i; *local = __dsub_rn(t, acc); return true; }
      acc = __dadd_rn(acc, Ti);
    }
    // past the end: last piece at t - sum(T[:-1])  (uav_trajectory.py:161-163)
    double head = 0.0;
    for (int i = 0; i + 1 < n; ++i) head = __dadd_rn(head, T[i]);
    *piece = n - 1;
    *local = __dsub_rn(t, head);
    return true;
  }
  for (int i = 0; i < n; ++i) {
    const double Ti = T[i];
    if (t <= __dadd_rn(acc, Ti)) { *piece = i; *local = __dsub_rn(t, acc); return true; }
    acc = __dadd_rn(acc, Ti);
  }
  return false;  // t > duration: the reference's assert (uav_trajectory.py:121)
}

__device__ __forceinline__ double sample_time(const double* __restrict__ ts, int ts_per_traj,
                                              const double* __restrict__ T, int n, int S, size_t b,
                                              int s) {
  if (ts) return ts_per_traj ? ts[b * S + s] : ts[s];
  double total = 0.0;
  for (int i = 0; i < n; ++i) total = __dadd_rn(total, T[i]);
  return __dmul_rn((double)s, __ddiv_rn(total, (double)S));
}

__global__ void __launch_bounds__(256)
sample_kernel(const double* __restrict__ coef, const double* __restrict__ dur, long long total,
              int n, int K, const double* __restrict__ ts, int ts_per_traj, int S, int mode,
              int deriv, double* __restrict__ out, uint8_t* __restrict__ status) {
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const size_t b = (size_t)(idx / S);
    const int s = (int)(idx - (long long)b * S);
    const double* T = dur + b * n;
    const double t = sample_time(ts, ts_per_traj, T, n, S, b, s);
    int piece = 0;
    double local = 0.0;
    const bool ok = find_piece(T, n, t, mode, &piece, &local);
    if (status) status[idx] = ok ? 0 : 1;
    double* o = out + (size_t)idx * K;
    if (!ok) {
      for (int k = 0; k < K; ++k) o[k] = qnan;
      continue;
    }
    const double* cp = coef + ((b * n + piece) * K) * MST_NCOEF;
    for (int k = 0; k < K; ++k) {
      double c[MST_NCOEF];
      const double2* src = reinterpret_cast<const double2*>(cp + k * MST_NCOEF);
      #pragma unroll
      for (int i = 0; i < 4; ++i) { const double2 v = src[i]; c[2 * i] = v.x; c[2 * i + 1] = v.y; }
      const int len = derive_in_place(c, deriv);
      o[k] = horner_nofma(c, len, local);
    }
  }
}

// Polynomial4D.eval (uav_trajectory.py:66-101): out[13] = pos vel acc omega yaw
__global__ void __launch_bounds__(128)
flat_kernel(const double* __restrict__ coef, const double* __restrict__ dur, long long total, int n,
            const double* __restrict__ ts, int ts_per_traj, int S, int mode,
            double* __restrict__ out, uint8_t* __restrict__ status) {
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const size_t b = (size_t)(idx / S);
    const int s = (int)(idx - (long long)b * S);
    const double* T = dur + b * n;
    const double t = sample_time(ts, ts_per_traj, T, n, S, b, s);
    int piece = 0;
    double local = 0.0;
    const bool ok = find_piece(T, n, t, mode, &piece, &local);
    if (status) status[idx] = ok ? 0 : 1;
    double* o = out + (size_t)idx * 13;
    if (!ok) {
      for (int k = 0; k < 13; ++k) o[k] = qnan;
      continue;
    }
    const double* cp = coef + ((b * n + piece) * 4) * MST_NCOEF;
    double val[4][4];  // [axis][derivative level]
    for (int k = 0; k < 4; ++k) {
      double c[MST_NCOEF];
      for (int i = 0; i < MST_NCOEF; ++i) c[i] = cp[k * MST_NCOEF + i];
      int len = MST_NCOEF;
      for (int level = 0; level < 4; ++level) {
        val[k][level] = horner_nofma(c, len, local);
        len = derive_once(c, len);
      }
    }
    // thrust direction and body axes, the numpy code's formulas; this file is built with the default
    // -fmad=true, so the dot / cross / norm products below may be FMA-contracted: omega agrees with
    // numpy to ~1e-11 relative, not bit for bit (pos / vel / acc / yaw above are the exact Horner)
    const double th0 = val[0][2], th1 = val[1][2], th2 = val[2][2] + 9.81;
    const double tn = sqrt(th0 * th0 + th1 * th1 + th2 * th2);
    const double zb0 = th0 / tn, zb1 = th1 / tn, zb2 = th2 / tn;
    const double yaw = val[3][0], dyaw = val[3][1];
    double sy, cy;
    sincos(yaw, &sy, &cy);
    // y_body = normalize(cross(z_body, x_world)), x_world = (cos, sin, 0)
    double y0 = zb1 * 0.0 - zb2 * sy, y1 = zb2 * cy - zb0 * 0.0, y2 = zb0 * sy - zb1 * cy;
    const double yn = sqrt(y0 * y0 + y1 * y1 + y2 * y2);
    y0 /= yn; y1 /= yn; y2 /= yn;
    const double x0 = y1 * zb2 - y2 * zb1, x1 = y2 * zb0 - y0 * zb2, x2 = y0 * zb1 - y1 * zb0;
    const double j0 = val[0][3], j1 = val[1][3], j2 = val[2][3];
    const double jz = j0 * zb0 + j1 * zb1 + j2 * zb2;
    const double h0 = (j0 - jz * zb0) / tn, h1 = (j1 - jz * zb1) / tn, h2 = (j2 - jz * zb2) / tn;
    o[0] = val[0][0]; o[1] = val[1][0]; o[2] = val[2][0];
    o[3] = val[0][1]; o[4] = val[1][1]; o[5] = val[2][1];
    o[6] = val[0][2]; o[7] = val[1][2]; o[8] = val[2][2];
    o[9] = -(h0 * y0 + h1 * y1 + h2 * y2);
    o[10] = h0 * x0 + h1 * x1 + h2 * x2;
    o[11] = zb2 * dyaw;
    o[12] = yaw;
  }
}

// rows[count][8][8]: derivative j, power k -> k!/(k-j)! * t^(k-j)
__global__ void time_power_kernel(const double* __restrict__ t, int count, double* __restrict__ rows) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)count * 64) return;
  const long long a = idx >> 6;
  const int j = (int)(idx >> 3) & 7, k = (int)idx & 7;
  rows[idx] = (k >= j) ? __dmul_rn(falling_factorial(k, j), pow_rounded_once(t[a], k - j)) : 0.0;
}

// Polynomial.derivative for `count` polynomials of `len` coefficients each:
// out[c][i] = (i+1) * p[c][i+1], i < len-1   (uav_trajectory.py:25-26)
__global__ void poly_derivative_kernel(const double* __restrict__ p, int count, int len, double* __restrict__ out) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)count * (len - 1)) return;
  const long long c = idx / (len - 1);
  const int i = (int)(idx - c * (len - 1));
  out[idx] = __dmul_rn((double)(i + 1), p[c * len + i + 1]);
}

// Polynomial.pol_coeffs_at_t: out[c][i] = p[c][i] * t[c]**i   (uav_trajectory.py:28-36)
__global__ void poly_terms_kernel(const double* __restrict__ p, const double* __restrict__ t, int count, int len,
                                  double* __restrict__ out) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)count * len) return;
  const long long c = idx / len;
  const int i = (int)(idx - c * len);
  out[idx] = __dmul_rn(p[idx], pow_rounded_once(t[c], i));
}

int launch_poly_derivative(const double* p, int count, int len, double* out, cudaStream_t stream) {
  const long long total = (long long)count * (len - 1);
  if (total <= 0) return MST_OK;
  poly_derivative_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, count, len, out);
  return check_launch();
}

int launch_poly_terms(const double* p, const double* t, int count, int len, double* out, cudaStream_t stream) {
  const long long total = (long long)count * len;
  if (total <= 0) return MST_OK;
  poly_terms_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, t, count, len, out);
  return check_launch();
}

// (n, 1 + 8K) float32 rows [T | x0..x7 | y0..y7 | ...] of path_to_pol
// (scripts/drones_pols_generator.py:63-77), one thread per output element
__global__ void pack_matrix_kernel(const double* __restrict__ coef, const double* __restrict__ dur, long long rows,
                                   int K, float* __restrict__ out) {
  const int width = 1 + MST_NCOEF * K;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < rows * width;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / width;
    const int col = (int)(idx - row * width);
    out[idx] = (float)(col == 0 ? dur[row] : coef[row * (width - 1) + col - 1]);
  }
}

int launch_pack_matrix(const double* coef, const double* dur, long long rows, int K, float* out, cudaStream_t stream) {
  if (rows <= 0) return MST_OK;
  const long long total = rows * (1 + MST_NCOEF * K);
  long long g = (total + 255) / 256;
  if (g > (long long)MST_SM_COUNT * 32) g = (long long)MST_SM_COUNT * 32;
  pack_matrix_kernel<<<(unsigned)g, 256, 0, stream>>>(coef, dur, rows, K, out);
  return check_launch();
}

// Snap cost J = sum over pieces and axes of int_0^T (d^4 p / dt^4)^2 dt = c^T Q(T) c, Q the snap
// Hessian of the monomial basis: Q[i][j] = i!/(i-4)! * j!/(j-4)! * T^(i+j-7) / (i+j-7), i,j >= 4.
// Not computed anywhere in the reference (its square system is the optimality system of exactly
// this cost); provided for time-allocation searches over re-solves (BASELINE config 3).
// One thread per trajectory.
__global__ void __launch_bounds__(128)
snap_cost_kernel(const double* __restrict__ coef, const double* __restrict__ dur, int B, int n, int K,
                 double* __restrict__ cost) {
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
    double total = 0.0;
    for (int i = 0; i < n; ++i) {
      const double T = dur[(size_t)b * n + i];
      const double T2 = T * T, T3 = T2 * T, T4 = T2 * T2, T5 = T4 * T, T6 = T3 * T3, T7 = T6 * T;
      for (int k = 0; k < K; ++k) {
        const double2* row = reinterpret_cast<const double2*>(coef + (((size_t)b * n + i) * K + k) * MST_NCOEF);
        const double2 c45 = __ldg(row + 2), c67 = __ldg(row + 3);
        // fourth derivative: a0 + a1 t + a2 t^2 + a3 t^3
        const double a0 = 24.0 * c45.x, a1 = 120.0 * c45.y, a2 = 360.0 * c67.x, a3 = 840.0 * c67.y;
        total += a0 * a0 * T + a0 * a1 * T2 + (2.0 * a0 * a2 + a1 * a1) * (T3 / 3.0) + (a0 * a3 + a1 * a2) * (T4 * 0.5) +
                 (2.0 * a1 * a3 + a2 * a2) * (T5 / 5.0) + a2 * a3 * (T6 / 3.0) + a3 * a3 * (T7 / 7.0);
      }
    }
    cost[b] = total;
  }
}

int launch_snap_cost(const double* coef, const double* dur, int B, int n, int K, double* cost, cudaStream_t stream) {
  if (B <= 0) return MST_OK;
  long long g = ((long long)B + 127) / 128;
  if (g > (long long)MST_SM_COUNT * 16) g = (long long)MST_SM_COUNT * 16;
  snap_cost_kernel<<<(unsigned)g, 128, 0, stream>>>(coef, dur, B, n, K, cost);
  return check_launch();
}

// d(cost)/d(T_i) of the OPTIMAL snap cost with the waypoints fixed and the knot derivatives free (they
// are what the solve optimises, so by the envelope theorem only the explicit dependence on T_i counts):
// minus the Hamiltonian of piece i, which is constant along an optimal piece and is evaluated at its
// start from the coefficients: H = x4^2 - 2 x5 x3 + 2 x6 x2 - 2 x7 x1 with xk the k-th derivative,
// summed over axes.  One thread per piece.  (Checked against central differences of re-solves to 1e-8.)
__global__ void __launch_bounds__(128)
time_gradient_kernel(const double* __restrict__ coef, long long pieces, int K, double* __restrict__ grad) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pieces; p += (long long)gridDim.x * blockDim.x) {
    double H = 0.0;
    for (int k = 0; k < K; ++k) {
      const double2* row = reinterpret_cast<const double2*>(coef + ((size_t)p * K + k) * MST_NCOEF);
      const double2 c01 = __ldg(row), c23 = __ldg(row + 1), c45 = __ldg(row + 2), c67 = __ldg(row + 3);
      const double x1 = c01.y, x2 = 2.0 * c23.x, x3 = 6.0 * c23.y, x4 = 24.0 * c45.x, x5 = 120.0 * c45.y,
                   x6 = 720.0 * c67.x, x7 = 5040.0 * c67.y;
      H += x4 * x4 - 2.0 * x5 * x3 + 2.0 * x6 * x2 - 2.0 * x7 * x1;
    }
    grad[p] = -H;
  }
}

int launch_time_gradient(const double* coef, int B, int n, int K, double* grad, cudaStream_t stream) {
  const long long pieces = (long long)B * n;
  if (pieces <= 0) return MST_OK;
  long long g = (pieces + 127) / 128;
  if (g > (long long)MST_SM_COUNT * 16) g = (long long)MST_SM_COUNT * 16;
  time_gradient_kernel<<<(unsigned)g, 128, 0, stream>>>(coef, pieces, K, grad);
  return check_launch();
}

static unsigned grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)MST_SM_COUNT * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

int launch_sample(const double* coef, const double* dur, int B, int n, int K, const double* ts,
                  int ts_per_traj, int S, int mode, int deriv, double* out, uint8_t* status,
                  cudaStream_t stream) {
  const long long total = (long long)B * S;
  if (total == 0) return MST_OK;
  sample_kernel<<<grid_for(total, 256), 256, 0, stream>>>(coef, dur, total, n, K, ts, ts_per_traj, S,
                                                          mode, deriv, out, status);
  return check_launch();
}

int launch_flat(const double* coef, const double* dur, int B, int n, const double* ts, int ts_per_traj,
                int S, int mode, double* out, uint8_t* status, cudaStream_t stream) {
  const long long total = (long long)B * S;
  if (total == 0) return MST_OK;
  flat_kernel<<<grid_for(total, 128), 128, 0, stream>>>(coef, dur, total, n, ts, ts_per_traj, S, mode,
                                                        out, status);
  return check_launch();
}

int launch_time_power(const double* t, int count, double* rows, cudaStream_t stream) {
  if (count == 0) return MST_OK;
  const long long total = (long long)count * 64;
  time_power_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(t, count, rows);
  return check_launch();
}

}  // namespace mst
