// extern "C" entry points of libmst.so (declared in include/mst.h).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "farcull.cuh"
#include "mst_common.cuh"
#include "onepass.cuh"

namespace mst {

static thread_local char g_cuda_error[256] = "";

void note_cuda_error(cudaError_t e) {
  snprintf(g_cuda_error, sizeof(g_cuda_error), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
}

int check_launch() {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { note_cuda_error(e); return MST_ERR_CUDA; }
  return MST_OK;
}

int sm_count() {
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

int allow_dynamic_smem(const void* kernel, size_t dynamic_bytes) {
  cudaFuncAttributes attr;
  cudaError_t e = cudaFuncGetAttributes(&attr, kernel);
  if (e != cudaSuccess) { note_cuda_error(e); return MST_ERR_CUDA; }
  if (attr.sharedSizeBytes + dynamic_bytes > MST_MAX_SMEM) return MST_ERR_TOO_LARGE;
  if (attr.sharedSizeBytes + dynamic_bytes <= 48 * 1024) return MST_OK;  // within the default limit
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynamic_bytes);
  if (e != cudaSuccess) { note_cuda_error(e); return MST_ERR_CUDA; }
  return MST_OK;
}

// kernels' host launchers (one per .cu file)
size_t banded_lu_smem_per_warp(int n, int R);
int launch_banded_lu(const double* wp, const double* t, int groups, int n, int K, int G, const int* list,
                     const int* list_count, double* coef, double* dur, int* info, double* scratch, int* ticket,
                     cudaStream_t stream);
size_t banded_lu_scratch_bytes(int groups, int n, int R);
size_t condensed_workspace_bytes(int groups);
int launch_condensed(const double* wp, const double* t, int groups, int n, int K, int G, int force,
                     double* coef, double* dur, int* info, int* list, int* list_count,
                     cudaStream_t stream, const FarCull* cull = nullptr);
int launch_sample_collide_cull(const double* coef, const double* dur, const unsigned* far_mask, int B, int n, int K,
                               int S, const mst_mesh* robot, const mst_mesh* env, uint8_t* hit, uint8_t* any_hit,
                               int* tile_counter, cudaStream_t stream);
bool sample_collide_cull_suits(int n, int K, int S, const mst_mesh* robot, const mst_mesh* env);
int launch_sample(const double* coef, const double* dur, int B, int n, int K, const double* ts,
                  int ts_per_traj, int S, int mode, int deriv, double* out, uint8_t* status,
                  cudaStream_t stream);
int launch_flat(const double* coef, const double* dur, int B, int n, const double* ts, int ts_per_traj,
                int S, int mode, double* out, uint8_t* status, cudaStream_t stream);
int launch_time_power(const double* t, int count, double* rows, cudaStream_t stream);
int launch_pack_matrix(const double* coef, const double* dur, long long rows, int K, float* out, cudaStream_t stream);
int launch_snap_cost(const double* coef, const double* dur, int B, int n, int K, double* cost, cudaStream_t stream);
int launch_time_gradient(const double* coef, int B, int n, int K, double* grad, cudaStream_t stream);
int launch_poly_derivative(const double* p, int count, int len, double* out, cudaStream_t stream);
int launch_poly_terms(const double* p, const double* t, int count, int len, double* out, cudaStream_t stream);
int launch_collide(const mst_mesh* robot, const mst_mesh* env, const double* pose, long long P,
                   int pose_dim, uint8_t* hit, cudaStream_t stream);
int launch_collide_motions(const mst_mesh* robot, const mst_mesh* env, const double* a, const double* b,
                           long long M, int steps, uint8_t* invalid, cudaStream_t stream);
int launch_formation(const double* rb, int F, int m, int pose_dim, const double* off, int D, int K,
                     double* wp, cudaStream_t stream);

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

int launch_sample_collide(const double* coef, const double* dur, int B, int n, int K, int S,
                          const mst_mesh* robot, const mst_mesh* env, uint8_t* hit, uint8_t* any_hit,
                          cudaStream_t stream, const int* list = nullptr, const int* list_count = nullptr, int G = 1);
int launch_onepass(const double* wp, const double* t, int groups, int n, int K, int G, int S, double* coef, double* dur,
                   int* info, uint8_t* hit, uint8_t* any_hit, int* list, int* counters, const mst_mesh* robot,
                   const mst_mesh* env, const WireTargets* wire, cudaStream_t stream);

int launch_wire_patch(const double* coef, const double* dur, const uint8_t* hit, const uint8_t* any_hit, int n, int K,
                      int G, int S, const int* list, const int* list_count, const WireTargets* wire,
                      cudaStream_t stream);

// Which pipeline mst_pipeline(MST_SOLVER_AUTO) runs: the two-launch one by default — measured faster
// on B200 (3.4 ms against 4.2 ms per 1 M trajectories: the single-pass kernel keeps 10 warps per SM
// resident instead of 14-16, profiles/r2_onepass_history.md); MST_PIPELINE_ONE_PASS=1 switches the
// default.  MST_SOLVER_AUTO_ONE_PASS always asks for the single-pass kernel, mst_pipeline_wire needs it.
static bool one_pass_default() {
  static const bool on = getenv("MST_PIPELINE_ONE_PASS") != nullptr && atoi(getenv("MST_PIPELINE_ONE_PASS")) != 0;
  return on;
}

// Trajectories per pass of the pipeline (solver launch + fused sample/collide launch).
// Measured on B200 (profiles/r1_chunk_sweep.txt): passes small enough to keep a pass's
// coefficients L2-resident between the two kernels (~26 k trajectories) lose more to launch
// gaps and tail effects than the saved re-read is worth (18.6 ms vs 11.7 ms per 1 M), so a
// pass is as large as 32-bit indexing comfortably allows.  MST_PIPELINE_CHUNK overrides it.
static int pipeline_chunk(int B, int n, int K, int G) {
  long long c = 1ll << 22;
  (void)n; (void)K;
  const char* env = getenv("MST_PIPELINE_CHUNK");
  if (env && atoll(env) > 0) c = atoll(env);
  if (c < G) c = G;
  c -= c % G;
  if (c > B) c = B;
  return (int)c;
}

}  // namespace mst

using namespace mst;

extern "C" int mst_version(void) { return MST_VERSION; }

extern "C" const char* mst_strerror(int code) {
  switch (code) {
    case MST_OK: return "ok";
    case MST_ERR_INVALID: return "invalid argument";
    case MST_ERR_TOO_LARGE: return "problem does not fit the on-chip working set";
    case MST_ERR_CUDA: return "CUDA runtime error";
    case MST_ERR_NOMEM: return "out of memory";
    default: return "unknown error";
  }
}

extern "C" const char* mst_last_cuda_error(void) { return g_cuda_error; }

extern "C" int mst_time_power_rows(const double* t, int count, double* rows, void* stream) {
  if (count < 0 || (count > 0 && (!t || !rows))) return MST_ERR_INVALID;
  return launch_time_power(t, count, rows, (cudaStream_t)stream);
}

extern "C" int mst_poly_derivative(const double* p, int count, int len, double* out, void* stream) {
  if (count < 0 || len < 1 || len > 64) return MST_ERR_INVALID;
  if (count == 0 || len == 1) return MST_OK;
  if (!p || !out) return MST_ERR_INVALID;
  return launch_poly_derivative(p, count, len, out, (cudaStream_t)stream);
}

extern "C" int mst_poly_terms_at_t(const double* p, const double* t, int count, int len, double* out, void* stream) {
  if (count < 0 || len < 1 || len > 64) return MST_ERR_INVALID;
  if (count == 0) return MST_OK;
  if (!p || !t || !out) return MST_ERR_INVALID;
  return launch_poly_terms(p, t, count, len, out, (cudaStream_t)stream);
}

extern "C" int mst_snap_cost(const double* coef, const double* dur, int B, int n, int K, double* cost, void* stream) {
  if (B < 0 || n < 1 || K < 1) return MST_ERR_INVALID;
  if (B == 0) return MST_OK;
  if (!coef || !dur || !cost) return MST_ERR_INVALID;
  return launch_snap_cost(coef, dur, B, n, K, cost, (cudaStream_t)stream);
}

extern "C" int mst_time_gradient(const double* coef, int B, int n, int K, double* grad, void* stream) {
  if (B < 0 || n < 1 || K < 1) return MST_ERR_INVALID;
  if (B == 0) return MST_OK;
  if (!coef || !grad) return MST_ERR_INVALID;
  return launch_time_gradient(coef, B, n, K, grad, (cudaStream_t)stream);
}

extern "C" int mst_pack_pol_matrix(const double* coef, const double* dur, int B, int n, int K, float* out,
                                   void* stream) {
  if (B < 0 || n < 1 || K < 1) return MST_ERR_INVALID;
  if (B == 0) return MST_OK;
  if (!coef || !dur || !out) return MST_ERR_INVALID;
  return launch_pack_matrix(coef, dur, (long long)B * n, K, out, (cudaStream_t)stream);
}

// solver workspace: [counters (64 ints) | list of the groups handed to the pivoted solver] [its scratch
// for the finished columns of U]
static size_t solve_list_bytes(int groups) { return align256(condensed_workspace_bytes(groups + 1)); }
static double* lu_scratch(void* workspace, int groups) {
  return workspace ? reinterpret_cast<double*>(static_cast<char*>(workspace) + solve_list_bytes(groups)) : nullptr;
}

extern "C" size_t mst_solve_workspace_bytes(int B, int n, int K, int share_time_group) {
  if (B < 0 || n < 1 || K < 1 || share_time_group < 1) return 0;
  const int groups = B / share_time_group;
  return solve_list_bytes(groups) + align256(banded_lu_scratch_bytes(groups, n, share_time_group * K));
}

static int solve_impl(const double* wp, const double* t, int B, int n, int K, int share_time_group, int solver,
                      double* coef, double* dur, int* info, void* workspace, void* stream, const FarCull* cull);

extern "C" int mst_solve_batch(const double* wp, const double* t, int B, int n, int K,
                               int share_time_group, int solver, double* coef, double* dur,
                               int* info, void* workspace, void* stream) {
  return solve_impl(wp, t, B, n, K, share_time_group, solver, coef, dur, info, workspace, stream, nullptr);
}

// cull (pipeline only, may be null): the condensed solver also writes the far-piece bits (farcull.cuh)
static int solve_impl(const double* wp, const double* t, int B, int n, int K, int share_time_group, int solver,
                      double* coef, double* dur, int* info, void* workspace, void* stream, const FarCull* cull) {
  const int G = share_time_group;
  if (B < 0 || n < 1 || K < 1 || G < 1 || B % G != 0) return MST_ERR_INVALID;
  if (solver != MST_SOLVER_AUTO && solver != MST_SOLVER_BANDED_LU && solver != MST_SOLVER_CONDENSED)
    return MST_ERR_INVALID;
  if (B == 0) return MST_OK;
  if (!wp || !t || !coef || !dur || !info) return MST_ERR_INVALID;
  const int groups = B / G;
  cudaStream_t st = (cudaStream_t)stream;
  const bool banded_fits = banded_lu_smem_per_warp(n, G * K) <= MST_MAX_SMEM;
  if (!banded_fits && solver == MST_SOLVER_BANDED_LU) return MST_ERR_TOO_LARGE;
  // trajectories too long for the pivoted solver's on-chip band: AUTO degrades to the condensed
  // solver alone, and a group it has to decline is reported through info[] (MST_INFO_DECLINED /
  // singular) instead of being solved
  if (!banded_fits && solver == MST_SOLVER_AUTO) solver = MST_SOLVER_CONDENSED;
  if (!workspace) return MST_ERR_INVALID;
  if (solver == MST_SOLVER_BANDED_LU)
    return launch_banded_lu(wp, t, groups, n, K, G, nullptr, nullptr, coef, dur, info, lu_scratch(workspace, groups), (int*)workspace + 33, st);
  int* list_count = (int*)workspace;
  int* list = list_count + 64;
  int rc = launch_condensed(wp, t, groups, n, K, G, solver == MST_SOLVER_CONDENSED, coef, dur, info,
                            list, list_count, st, cull);
  if (rc != MST_OK || solver == MST_SOLVER_CONDENSED) return rc;
  // groups the condensed path declined (duration spread too wide, t[0] != 0, bad input)
  return launch_banded_lu(wp, t, groups, n, K, G, list, list_count, coef, dur, info, lu_scratch(workspace, groups), (int*)workspace + 33, st);
}

extern "C" int mst_sample_batch(const double* coef, const double* dur, int B, int n, int K,
                                const double* ts, int ts_per_traj, int S, int mode, int deriv,
                                double* out, uint8_t* status, void* stream) {
  if (B < 0 || n < 1 || K < 1 || S < 0 || deriv < 0 || deriv > 8) return MST_ERR_INVALID;
  if (mode != MST_SAMPLE_PIECEWISE && mode != MST_SAMPLE_TRAJECTORY) return MST_ERR_INVALID;
  if ((long long)B * S == 0) return MST_OK;
  if (!coef || !dur || !out) return MST_ERR_INVALID;
  return launch_sample(coef, dur, B, n, K, ts, ts_per_traj, S, mode, deriv, out, status,
                       (cudaStream_t)stream);
}

extern "C" int mst_flat_outputs(const double* coef, const double* dur, int B, int n, const double* ts,
                                int ts_per_traj, int S, int mode, double* out, uint8_t* status,
                                void* stream) {
  if (B < 0 || n < 1 || S < 0) return MST_ERR_INVALID;
  if (mode != MST_SAMPLE_PIECEWISE && mode != MST_SAMPLE_TRAJECTORY) return MST_ERR_INVALID;
  if ((long long)B * S == 0) return MST_OK;
  if (!coef || !dur || !out) return MST_ERR_INVALID;
  return launch_flat(coef, dur, B, n, ts, ts_per_traj, S, mode, out, status, (cudaStream_t)stream);
}

extern "C" int mst_formation_waypoints(const double* rb, int F, int m, int pose_dim, const double* off,
                                       int D, int K, double* wp, void* stream) {
  if (F < 0 || m < 0 || D < 0 || (pose_dim != 4 && pose_dim != 7) || (K != 3 && K != 4))
    return MST_ERR_INVALID;
  if ((long long)F * m * D == 0) return MST_OK;
  if (!rb || !off || !wp) return MST_ERR_INVALID;
  return launch_formation(rb, F, m, pose_dim, off, D, K, wp, (cudaStream_t)stream);
}

extern "C" int mst_collide_poses(mst_mesh_t robot, mst_mesh_t env, const double* pose, int P,
                                 int pose_dim, uint8_t* hit, void* stream) {
  if (!robot || !env || P < 0 || (pose_dim != 3 && pose_dim != 4 && pose_dim != 7)) return MST_ERR_INVALID;
  if (P == 0) return MST_OK;
  if (!pose || !hit) return MST_ERR_INVALID;
  return launch_collide(robot, env, pose, P, pose_dim, hit, (cudaStream_t)stream);
}

extern "C" int mst_collide_motions(mst_mesh_t robot, mst_mesh_t env, const double* state_a, const double* state_b,
                                   int M, int steps, uint8_t* invalid, void* stream) {
  if (!robot || !env || M < 0 || steps < 1) return MST_ERR_INVALID;
  if (M == 0) return MST_OK;
  if (!state_a || !state_b || !invalid) return MST_ERR_INVALID;
  return launch_collide_motions(robot, env, state_a, state_b, M, steps, invalid, (cudaStream_t)stream);
}

extern "C" int mst_collide_trajectories(const double* coef, const double* dur, int B, int n, int K, int S,
                                        mst_mesh_t robot, mst_mesh_t env, uint8_t* hit, uint8_t* any_hit,
                                        void* stream) {
  if (S < 1 || (K != 3 && K != 4) || !robot || !env || B < 0 || n < 1) return MST_ERR_INVALID;
  if (B == 0) return MST_OK;
  if (!coef || !dur || !hit || !any_hit) return MST_ERR_INVALID;
  return launch_sample_collide(coef, dur, B, n, K, S, robot, env, hit, any_hit, (cudaStream_t)stream);
}

// workspace of the pipeline: the solver's (declined-group list) and, behind it, three far-piece words per
// trajectory (farcull.cuh)
extern "C" size_t mst_pipeline_workspace_bytes(int B, int n, int K, int share_time_group, int S) {
  (void)S;
  if (B < 0 || K < 1 || n < 1 || share_time_group < 1) return 0;
  return mst_solve_workspace_bytes(B, n, K, share_time_group) + align256(sizeof(unsigned) * 3 * (size_t)B);
}

// MST_PIPELINE_NO_CULL=1 switches the far-piece culling of the two-launch pipeline off (A/B measurements)
static bool cull_enabled() {
  static const bool off = getenv("MST_PIPELINE_NO_CULL") != nullptr && atoi(getenv("MST_PIPELINE_NO_CULL")) != 0;
  return !off;
}

// does the single-pass kernel take these sizes?  (mirrors launch_onepass's own checks, minus the meshes)
static bool onepass_sizes(int n, int K, int G, int solver, int S) {
  const bool wanted = solver == MST_SOLVER_AUTO_ONE_PASS || (solver == MST_SOLVER_AUTO && one_pass_default());
  return wanted && (K == 3 || K == 4) && G * K <= 32 && n >= 2 && n <= 32 && S >= 32 && S <= 4096 &&
         banded_lu_smem_per_warp(n, G * K) <= MST_MAX_SMEM;
}

extern "C" int mst_pipeline_launch_count(int B, int n, int K, int share_time_group, int solver, int S) {
  if (B <= 0 || S < 1 || K < 1 || n < 1 || share_time_group < 1) return 0;
  const int chunk = pipeline_chunk(B, n, K, share_time_group);
  const int chunks = (B + chunk - 1) / chunk;
  // single pass: the fused kernel + the two list-mode kernels behind it (pivoted solver, sampling of
  // its groups; both exit at once when no group was handed over)
  if (onepass_sizes(n, K, share_time_group, solver, S)) return chunks * 3;
  const int solve = (solver == MST_SOLVER_AUTO || solver == MST_SOLVER_AUTO_ONE_PASS) ? 2 : 1;  // condensed (+ banded LU over the declined list)
  return chunks * (solve + 1);                           // + fused sample/collide/any-hit
}

// stages: 0 = the whole pipeline; 1 = its solver launch(es) only; 2 = its sampling / collision launch only
// (measurement hook: the second stage then works on what an earlier call left in coef / dur / workspace)
// mat (may be null): the float32 polynomial matrix as an extra output of the two-launch pipeline
static int pipeline_impl(const double* wp, const double* t, int B, int n, int K, int G, int solver, int S,
                         mst_mesh_t robot, mst_mesh_t env, double* coef, double* dur, int* info, uint8_t* hit,
                         uint8_t* any_hit, const WireTargets* wire, void* workspace, void* stream, int stage = 0,
                         float* mat = nullptr) {
  cudaStream_t st = (cudaStream_t)stream;
  const int chunk = pipeline_chunk(B, n, K, G);
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = (B - b0 < chunk) ? (B - b0) : chunk;
    const double* wc = wp + (size_t)b0 * (n + 1) * K;
    const double* tc = t + (size_t)(b0 / G) * (n + 1);
    double* cc = coef + (size_t)b0 * n * K * MST_NCOEF;
    double* dd = dur + (size_t)b0 * n;
    uint8_t* hh = hit ? hit + (size_t)b0 * S : nullptr;
    uint8_t* aa = any_hit ? any_hit + b0 : nullptr;
    int rc = MST_ERR_TOO_LARGE;
    if (mat == nullptr && onepass_sizes(n, K, G, solver, S)) {
      int* counters = (int*)workspace;
      int* list = counters + 64;
      WireTargets wchunk;
      if (wire) { wchunk = *wire; wchunk.row0 += b0; }
      rc = launch_onepass(wc, tc, nb / G, n, K, G, S, cc, dd, info + b0, hh, aa, list, counters, robot, env,
                          wire ? &wchunk : nullptr, st);
      if (rc == MST_OK) {
        // groups the condensed path must not take: pivoted solve, then their samples (list mode)
        rc = launch_banded_lu(wc, tc, nb / G, n, K, G, list, counters, cc, dd, info + b0, lu_scratch(workspace, nb / G), (int*)workspace + 33, st);
        if (rc != MST_OK) return rc;
        rc = launch_sample_collide(cc, dd, nb, n, K, S, robot, env, hh, aa, st, list, counters, G);
        if (rc != MST_OK) return rc;
        if (wire) {   // their wire rows, from the local results
          rc = launch_wire_patch(cc, dd, hh, aa, n, K, G, S, list, counters, &wchunk, st);
          if (rc != MST_OK) return rc;
        }
        continue;
      }
      if (rc != MST_ERR_TOO_LARGE) return rc;
    }
    if (wire) return MST_ERR_TOO_LARGE;   // the two-launch pipeline has no wire outputs
    const int solver2 = solver == MST_SOLVER_AUTO_ONE_PASS ? MST_SOLVER_AUTO : solver;
    if (cull_enabled() && solver2 != MST_SOLVER_BANDED_LU && sample_collide_cull_suits(n, K, S, robot, env)) {
      // far-piece culling: the solver bounds every piece it solves, the sampling kernel skips the far ones
      FarCull fc;
      memset(&fc, 0, sizeof(fc));
      for (int a = 0; a < 3; ++a) {
        fc.elo[a] = env->bounds.root[a];
        fc.ehi[a] = env->bounds.root[3 + a];
        fc.rlo[a] = robot->bounds.root[a];
        fc.rhi[a] = robot->bounds.root[3 + a];
        fc.lo[a] = fc.elo[a] - fc.rhi[a];       // K = 3: positions below this keep the robot before the obstacles
        fc.hi[a] = fc.ehi[a] - fc.rlo[a];
      }
      fc.radius = robot->bounds.rxy;
      fc.yaw = K == 4;
      fc.mask = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + mst_solve_workspace_bytes(B, n, K, G)) +
                (size_t)3 * b0;
      fc.mat = mat ? mat + (size_t)b0 * n * (1 + MST_NCOEF * K) : nullptr;
      if (stage != 2) {
        cudaError_t e = cudaMemsetAsync(fc.mask, 0, sizeof(unsigned) * 3 * (size_t)nb, st);   // groups solved elsewhere: nothing far
        if (e != cudaSuccess) { note_cuda_error(e); return MST_ERR_CUDA; }
        rc = solve_impl(wc, tc, nb, n, K, G, solver2, cc, dd, info + b0, workspace, stream, &fc);
        if (rc == MST_ERR_TOO_LARGE && fc.mat) {   // solver kernel without the matrix output: packing pass instead
          float* m = fc.mat;
          fc.mat = nullptr;
          rc = solve_impl(wc, tc, nb, n, K, G, solver2, cc, dd, info + b0, workspace, stream, &fc);
          if (rc == MST_OK) rc = launch_pack_matrix(cc, dd, (long long)nb * n, K, m, st);
        }
        if (rc != MST_OK) return rc;
        if (fc.mat && solver2 == MST_SOLVER_AUTO) {   // rows of the groups the pivoted solver finished
          WireTargets wt;
          memset(&wt, 0, sizeof(wt));
          wt.count = 1;
          wt.mat[0] = fc.mat;
          rc = launch_wire_patch(cc, dd, nullptr, nullptr, n, K, G, S, (int*)workspace + 64, (int*)workspace, &wt, st);
          if (rc != MST_OK) return rc;
        }
      }
      if (stage == 1) continue;
      rc = launch_sample_collide_cull(cc, dd, fc.mask, nb, n, K, S, robot, env, hh, aa, (int*)workspace + 32, st);   // counters[32]: tile tickets
      if (rc == MST_OK) continue;
      if (rc != MST_ERR_TOO_LARGE) return rc;
      rc = launch_sample_collide(cc, dd, nb, n, K, S, robot, env, hh, aa, st);
      if (rc != MST_OK) return rc;
      continue;
    }
    rc = mst_solve_batch(wc, tc, nb, n, K, G, solver2, cc, dd, info + b0, workspace, stream);
    if (rc != MST_OK) return rc;
    rc = launch_sample_collide(cc, dd, nb, n, K, S, robot, env, hh, aa, st);
    if (rc != MST_OK) return rc;
    if (mat) {   // no solver-side packing on this path: the packing pass
      rc = launch_pack_matrix(cc, dd, (long long)nb * n, K, mat + (size_t)b0 * n * (1 + MST_NCOEF * K), st);
      if (rc != MST_OK) return rc;
    }
  }
  return MST_OK;
}

extern "C" int mst_pipeline(const double* wp, const double* t, int B, int n, int K,
                            int share_time_group, int solver, int S, mst_mesh_t robot, mst_mesh_t env,
                            double* coef, double* dur, int* info, uint8_t* hit, uint8_t* any_hit,
                            void* workspace, void* stream) {
  const int G = share_time_group;
  if (S < 1 || (K != 3 && K != 4) || !robot || !env || G < 1 || B < 0 || n < 1 || B % G != 0)
    return MST_ERR_INVALID;
  if (solver != MST_SOLVER_AUTO && solver != MST_SOLVER_BANDED_LU && solver != MST_SOLVER_CONDENSED &&
      solver != MST_SOLVER_AUTO_ONE_PASS)
    return MST_ERR_INVALID;
  if (B == 0) return MST_OK;
  if (!hit || !any_hit || !workspace || !wp || !t || !coef || !dur || !info) return MST_ERR_INVALID;
  return pipeline_impl(wp, t, B, n, K, G, solver, S, robot, env, coef, dur, info, hit, any_hit, nullptr, workspace,
                       stream);
}

extern "C" int mst_pipeline_packed(const double* wp, const double* t, int B, int n, int K, int share_time_group,
                                   int solver, int S, mst_mesh_t robot, mst_mesh_t env, double* coef, double* dur,
                                   int* info, uint8_t* hit, uint8_t* any_hit, float* pol_matrix, void* workspace,
                                   void* stream) {
  const int G = share_time_group;
  if (S < 1 || (K != 3 && K != 4) || !robot || !env || G < 1 || B < 0 || n < 1 || B % G != 0) return MST_ERR_INVALID;
  if (solver != MST_SOLVER_AUTO && solver != MST_SOLVER_BANDED_LU && solver != MST_SOLVER_CONDENSED)
    return MST_ERR_INVALID;
  if (B == 0) return MST_OK;
  if (!hit || !any_hit || !workspace || !wp || !t || !coef || !dur || !info || !pol_matrix) return MST_ERR_INVALID;
  return pipeline_impl(wp, t, B, n, K, G, solver, S, robot, env, coef, dur, info, hit, any_hit, nullptr, workspace,
                       stream, 0, pol_matrix);
}

extern "C" int mst_pipeline_stage(int stage, const double* wp, const double* t, int B, int n, int K,
                                  int share_time_group, int solver, int S, mst_mesh_t robot, mst_mesh_t env,
                                  double* coef, double* dur, int* info, uint8_t* hit, uint8_t* any_hit,
                                  void* workspace, void* stream) {
  const int G = share_time_group;
  if (stage < 0 || stage > 2 || S < 1 || (K != 3 && K != 4) || !robot || !env || G < 1 || B < 0 || n < 1 || B % G != 0)
    return MST_ERR_INVALID;
  if (solver != MST_SOLVER_AUTO && solver != MST_SOLVER_CONDENSED) return MST_ERR_INVALID;
  if (B == 0) return MST_OK;
  if (!hit || !any_hit || !workspace || !wp || !t || !coef || !dur || !info) return MST_ERR_INVALID;
  if (!cull_enabled() || !sample_collide_cull_suits(n, K, S, robot, env)) return MST_ERR_TOO_LARGE;
  return pipeline_impl(wp, t, B, n, K, G, solver, S, robot, env, coef, dur, info, hit, any_hit, nullptr, workspace,
                       stream, stage);
}

extern "C" int mst_pipeline_wire(const double* wp, const double* t, int B, int n, int K, int share_time_group, int S,
                                 mst_mesh_t robot, mst_mesh_t env, double* coef, double* dur, int* info,
                                 uint8_t* hit, uint8_t* any_hit, const mst_wire_targets* wire, void* workspace,
                                 void* stream) {
  const int G = share_time_group;
  if (S < 1 || (K != 3 && K != 4) || !robot || !env || G < 1 || B < 0 || n < 1 || B % G != 0) return MST_ERR_INVALID;
  if (!wire || wire->count < 1 || wire->count > MST_WIRE_MAX_TARGETS || wire->row_offset < 0) return MST_ERR_INVALID;
  if (B == 0) return MST_OK;
  if (!workspace || !wp || !t || !coef || !dur || !info || !hit || !any_hit) return MST_ERR_INVALID;
  WireTargets w;
  memset(&w, 0, sizeof(w));
  w.count = wire->count;
  w.row0 = wire->row_offset;
  for (int i = 0; i < wire->count; ++i) {
    w.mat[i] = wire->pol_matrix ? wire->pol_matrix[i] : nullptr;
    w.hit[i] = wire->hit ? wire->hit[i] : nullptr;
    w.any[i] = wire->any_hit ? wire->any_hit[i] : nullptr;
    if ((w.hit[i] == nullptr) != (w.any[i] == nullptr)) return MST_ERR_INVALID;
  }
  return pipeline_impl(wp, t, B, n, K, G, MST_SOLVER_AUTO_ONE_PASS, S, robot, env, coef, dur, info, hit, any_hit, &w,
                       workspace, stream);
}
