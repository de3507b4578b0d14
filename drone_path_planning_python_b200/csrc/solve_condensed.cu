// Fast solver: condensed block-tridiagonal LDL^T (condensed_core.cuh), one time group per
// thread, per-thread scratch in shared memory laid out slot-major ([slot][thread]) so a
// warp's accesses to one slot are 32 consecutive doubles (no bank conflicts).
//
// Replaces calculate_trajectory1D/4D (src/optimizations/calculatingTrajectories.py:37-213)
// for time groups whose duration spread allows it; every other group is appended to a
// device-side list that the banded pivoted-LU kernel consumes right after (no host sync).
#include "condensed_core.cuh"

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

template <int KC>
__global__ void __launch_bounds__(128)
condensed_kernel(const double* __restrict__ wp, const double* __restrict__ tstamps, int groups, int n,
                 int K, int G, int force, double* __restrict__ coef, double* __restrict__ dur,
                 int* __restrict__ info, int* __restrict__ list, int* __restrict__ list_count) {
  extern __shared__ double sm[];
  const int stride = blockDim.x;
  double* scratch = sm + threadIdx.x;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < groups;
       g += (long long)gridDim.x * blockDim.x) {
    const double* tg = tstamps + (size_t)g * (n + 1);
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
    for (int i = 0; i < n; ++i) scratch[(size_t)i * stride] = tg[i + 1] - tg[i];
    condensed_factor(n, scratch, stride);
    for (int d = 0; d < G; ++d) {
      const size_t traj = (size_t)g * G + d;
      const double* wpd = wp + traj * (size_t)(n + 1) * K;
      condensed_forward<KC>(wpd, n, K, scratch, stride);
      double* cd = coef + traj * (size_t)n * K * MST_NCOEF;
      condensed_backward<KC>(wpd, n, K, scratch, stride,
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

int launch_condensed(const double* wp, const double* t, int groups, int n, int K, int G, int force,
                     double* coef, double* dur, int* info, int* list, int* list_count,
                     cudaStream_t stream) {
  if (K > 4) {
    // more axes than the per-thread register tile: everything goes to the pivoted solver
    if (force) return MST_ERR_INVALID;
  }
  cudaError_t e = cudaMemsetAsync(list_count, 0, sizeof(int), stream);
  if (e != cudaSuccess) { note_cuda_error(e); return MST_ERR_CUDA; }
  const size_t per_thread = sizeof(double) * (size_t)condensed_slots(n, K > 4 ? 4 : K);
  int threads = (int)(MST_MAX_SMEM / per_thread);
  if (threads >= 32) threads -= threads % 32;  // long trajectories: a partial warp per CTA still works
  if (threads > 128) threads = 128;
  if (threads > 64 && per_thread * 64 * 3 <= MST_MAX_SMEM) threads = 64;  // more CTAs per SM
  const bool fits = threads >= 8 && K <= 4;
  if (!fits && force) return MST_ERR_TOO_LARGE;
  if (!fits) threads = 32;
  const size_t smem = fits ? per_thread * threads : 0;
  auto kern = condensed_kernel<4>;
  if (K <= 3) kern = condensed_kernel<3>;
  {  // per device and cheap, so set on every launch
    const cudaError_t a = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MST_MAX_SMEM);
    if (a != cudaSuccess) { note_cuda_error(a); return MST_ERR_CUDA; }
  }
  long long blocks = ((long long)groups + threads - 1) / threads;
  const long long cap = (long long)MST_SM_COUNT * 64;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  // when the scratch does not fit (very long trajectories) every group is declined
  kern<<<(unsigned)blocks, threads, smem, stream>>>(wp, t, groups, n, K, G, fits ? force : -1, coef, dur,
                                                     info, list, list_count);
  return check_launch();
}

}  // namespace mst
