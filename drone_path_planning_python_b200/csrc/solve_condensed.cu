// Fast solver: condensed block-tridiagonal LDL^T (condensed_core.cuh), one time group per
// thread, per-thread scratch in shared memory laid out slot-major ([slot][thread]) so a
// warp's accesses to one slot are 32 consecutive doubles (no bank conflicts).
//
// Replaces calculate_trajectory1D/4D (src/optimizations/calculatingTrajectories.py:37-213)
// for time groups whose duration spread allows it; every other group is appended to a
// device-side list that the banded pivoted-LU kernel consumes right after (no host sync).
#include "condensed_core.cuh"
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
    for (int i = 0; i < n; ++i) scratch[(size_t)i * stride] = tg[i + 1] - tg[i];
    condensed_factor(n, scratch, stride);
    for (int d = 0; d < G; ++d) {
      const size_t traj = (size_t)g * G + d;
      const double* wpd = STAGED ? wp_tile + (size_t)threadIdx.x * (n + 1) * K : wp + traj * (size_t)(n + 1) * K;
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
