/*
 * mst.h — C ABI of libmst.so: batched minimum-snap trajectory generation, sampling,
 * formation transform and mesh collision checking on NVIDIA B200 (sm_100a).
 *
 * The reference (mjmyt/drone_path_planning_python) is pure Python and has no FFI of
 * its own; the drop-in boundary is its Python module surface.  The
 * entry points below are what a binding for that surface calls — each one names the
 * reference interface it replaces (paths relative to the reference checkout).  The
 * Python side (drone_path_planning_python_b200/_abi.py) binds them with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - every data pointer is a CUDA DEVICE pointer owned by the caller (e.g.
 *    torch.Tensor.data_ptr()); `stream` is a cudaStream_t passed as void*
 *    (0 = default stream).  Calls are asynchronous on that stream and re-entrant for
 *    distinct streams / workspaces (rospy runs callback1/callback2 of
 *    scripts/drones_pols_generator.py:22-37 on different threads).
 *  - return value: MST_OK (0) or a negative MST_ERR_* code; nothing throws across
 *    the ABI; no allocation happens after mst_mesh_create.
 *  - all floating point is IEEE double.  Polynomial coefficients are in ASCENDING
 *    power order, 8 per piece (7th order), as in the reference.
 *  - layouts are row-major with the last index fastest.
 */
#ifndef MST_H_
#define MST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MST_VERSION 100 /* 0.1.0 */

enum {
  MST_OK = 0,
  MST_ERR_INVALID = -1,   /* bad argument (null pointer, negative size, bad mode)     */
  MST_ERR_TOO_LARGE = -2, /* problem does not fit the kernel's on-chip working set     */
  MST_ERR_CUDA = -3,      /* a CUDA runtime call failed; see mst_last_cuda_error()     */
  MST_ERR_NOMEM = -4      /* host or device allocation failed (mst_mesh_create only)   */
};

/* per-trajectory status written to info[] by the solvers */
enum {
  MST_INFO_OK = 0,            /* solved                                                  */
  MST_INFO_DECREASING = -1,   /* a duration t[i+1]-t[i] (or t[0]) is negative: the
                                 reference raises AssertionError
                                 (src/optimizations/uav_trajectory.py:30)              */
  MST_INFO_NONFINITE = -2,    /* NaN/Inf in times                                         */
  MST_INFO_DECLINED = -3      /* MST_SOLVER_CONDENSED was forced on a group it cannot
                                 reproduce (t[0] != 0)              */
  /* > 0: zero pivot met at column info (1-based): the reference raises
     numpy.linalg.LinAlgError("Singular matrix")
     (src/optimizations/calculatingTrajectories.py:137)                                  */
};

/* solver selection for mst_solve_batch / mst_pipeline */
enum {
  MST_SOLVER_AUTO = 0,   /* condensed LDL^T where the duration spread allows it (max T / min T
                            <= 4), banded LU with partial
                            pivoting otherwise (decided per time group).
                            Trajectories too long for the pivoted solver's on-chip band
                            (28 x 8n doubles + right-hand sides > 227 kB, n > ~120) are all
                            solved by the condensed solver, at its accuracy;
                            groups it cannot take (t[0] != 0, zero-length piece) are
                            reported through info[]                                        */
  MST_SOLVER_BANDED_LU = 1,
  MST_SOLVER_CONDENSED = 2,
  MST_SOLVER_AUTO_ONE_PASS = 3 /* mst_pipeline only: MST_SOLVER_AUTO's solver choice, run by the
                                  single-pass kernel (solve + sample + collide in one persistent
                                  launch; the coefficients reach HBM once and are never read back)
                                  when the sizes suit it: K = 3 / 4, 2 <= n <= 32, 32 <= S <= 4096,
                                  G*K <= 32.  Identical results; on B200 the two-launch pipeline
                                  is the faster one, so MST_SOLVER_AUTO runs that
                                  (MST_PIPELINE_ONE_PASS=1 in the environment switches it)          */
};

/* piece selection semantics for mst_sample_batch */
enum {
  MST_SAMPLE_PIECEWISE = 0, /* PiecewisePolynomial.eval: strict t < acc+T_i, extrapolates
                               the last piece (src/optimizations/uav_trajectory.py:154-169) */
  MST_SAMPLE_TRAJECTORY = 1 /* Trajectory.eval: inclusive t <= acc+T_i
                               (src/optimizations/uav_trajectory.py:119-127)              */
};

int mst_version(void);
const char* mst_strerror(int code);
/* text of the last CUDA error seen by the calling thread ("" if none) */
const char* mst_last_cuda_error(void);

/*
 * Time-power rows — Polynomial.pol_coeffs_at_t applied to the j-th derivative of the
 * all-ones polynomial and left-padded to 8 (src/optimizations/uav_trajectory.py:25-36,
 * src/optimizations/calculatingTrajectories.py:68-71,93-97,105-109).
 *   t[count]            -> rows[count][8 (derivative j)][8 (power k)] = k!/(k-j)! * t^(k-j)
 */
int mst_time_power_rows(const double* t, int count, double* rows, void* stream);

/*
 * Polynomial.derivative (src/optimizations/uav_trajectory.py:25-26) for `count` polynomials
 * of `len` ascending coefficients:  p[count][len] -> out[count][len-1], out[i] = (i+1)*p[i+1].
 */
int mst_poly_derivative(const double* p, int count, int len, double* out, void* stream);

/*
 * Polynomial.pol_coeffs_at_t (src/optimizations/uav_trajectory.py:28-36):
 *   p[count][len], t[count] -> out[count][len], out[c][i] = p[c][i] * t[c]**i
 */
int mst_poly_terms_at_t(const double* p, const double* t, int count, int len, double* out, void* stream);

/*
 * Batched minimum-snap solve — calculate_trajectory1D / calculate_trajectory4D
 * (src/optimizations/calculatingTrajectories.py:37-213): for every trajectory the
 * 8n x 8n interpolation + C^1..C^6 continuity + rest-to-rest system is solved for
 * all K axes at once.
 *   wp   [B][n+1][K]   waypoint values per axis (x, y, z[, yaw])
 *   t    [B/G][n+1]    time stamps; G = share_time_group consecutive trajectories use
 *                      the same stamps (the D drones of a formation inherit the rigid
 *                      body path's stamps, scripts/drones_pols_generator.py:44-56);
 *                      G = 1 gives every trajectory its own stamps.  B % G must be 0.
 *   coef [B][n][K][8]  piece-major: one row of the reference's Pol_matrix per piece
 *   dur  [B][n]        durations T_i = t[i+1]-t[i]  (PiecewisePolynomial.time_durations)
 *   info [B]           MST_INFO_* per trajectory
 *   workspace          mst_solve_workspace_bytes(B, n, K, G) bytes of device scratch, required by every
 *                      solver: the list of the groups handed to the pivoted solver, and that solver's
 *                      finished columns of U (18 x 8n doubles per resident warp, at most ~110 MB)
 */
size_t mst_solve_workspace_bytes(int B, int n, int K, int share_time_group);
int mst_solve_batch(const double* wp, const double* t, int B, int n, int K,
                    int share_time_group, int solver, double* coef, double* dur,
                    int* info, void* workspace, void* stream);

/*
 * Snap cost of solved trajectories: cost[b] = sum over pieces and axes of the integral of the
 * squared 4th derivative over the piece = c^T Q(T) c.  An extension: the reference never
 * evaluates the cost whose optimality system it solves
 * (src/optimizations/calculatingTrajectories.py:13-35); it is the objective of time-allocation
 * searches over re-solves (BASELINE config 3).
 *   coef [B][n][K][8], dur [B][n]  ->  cost [B]
 */
int mst_snap_cost(const double* coef, const double* dur, int B, int n, int K, double* cost, void* stream);

/*
 * Derivative of the optimal snap cost with respect to every piece duration, waypoints fixed:
 * grad[b][i] = -(Hamiltonian of piece i summed over axes), evaluated from the coefficients
 * (H = x4^2 - 2 x5 x3 + 2 x6 x2 - 2 x7 x1, xk the k-th derivative at the start of the piece).  An
 * extension like mst_snap_cost: the gradient of time-allocation searches, from ONE solve instead
 * of n finite-difference re-solves.
 *   coef [B][n][K][8]  ->  grad [B][n]
 */
int mst_time_gradient(const double* coef, int B, int n, int K, double* grad, void* stream);

/*
 * Polynomial-piece matrix — the wire format path_to_pol writes to CSV and publishes
 * (scripts/drones_pols_generator.py:63-87): per piece one float32 row
 * [T | x0..x7 | y0..y7 | z0..z7 | yaw0..yaw7].
 *   coef [B][n][K][8], dur [B][n]  ->  out [B][n][1 + 8K] float32
 */
int mst_pack_pol_matrix(const double* coef, const double* dur, int B, int n, int K, float* out,
                        void* stream);

/*
 * CSV text of polynomial matrices — the file np.savetxt(name, matrix, delimiter=",") writes in
 * path_to_pol (scripts/drones_pols_generator.py:79-81): one line per piece, `width` = 1 + 8K fields
 * in numpy's default '%.18e' format.  Byte-identical to numpy (digits by exact integer arithmetic).
 *   mat [B][n][width] float32  ->  text [B][stride] bytes (trajectory b's file starts at
 *   text + b * stride and is length[b] bytes long); stride >= mst_csv_stride(n, width)
 */
size_t mst_csv_stride(int n, int width);
int mst_format_pol_matrix_csv(const float* mat, int B, int n, int width, char* text, long long stride,
                              int* length, void* stream);

/*
 * Batched evaluation — Polynomial.eval / Polynomial.derivative /
 * PiecewisePolynomial.eval / Trajectory.eval
 * (src/optimizations/uav_trajectory.py:17-26,119-127,154-169).
 *   coef [B][n][K][8], dur [B][n]
 *   ts   [S] (ts_per_traj = 0), [B][S] (ts_per_traj = 1), or NULL: uniform
 *        t_s = s * (sum(dur)/S), the np.arange(0, duration, step) pattern of
 *        src/trajectory_visualising/visualization.py:53
 *   mode MST_SAMPLE_*; deriv 0..7 (Polynomial.derivative applied `deriv` times)
 *   out  [B][S][K]; status [B][S] (may be NULL): 0 ok, 1 = t violates the
 *        reference's assert (t < 0, or t > duration in TRAJECTORY mode) -> out = NaN
 */
int mst_sample_batch(const double* coef, const double* dur, int B, int n, int K,
                     const double* ts, int ts_per_traj, int S, int mode, int deriv,
                     double* out, uint8_t* status, void* stream);

/*
 * Differential-flatness outputs — Polynomial4D.eval through Trajectory.eval
 * (src/optimizations/uav_trajectory.py:66-101,119-127); K must be 4.
 *   out [B][S][13] = pos(3) vel(3) acc(3) omega(3) yaw(1)
 */
int mst_flat_outputs(const double* coef, const double* dur, int B, int n,
                     const double* ts, int ts_per_traj, int S, int mode,
                     double* out, uint8_t* status, void* stream);

/*
 * Formation rigid-body transform — transform(path) of
 * scripts/drones_traj_generator.py:56-89: p_d = R(q_rb) * offset_d + t_rb, the rigid
 * body's heading carried over to every drone.
 *   rb   [F][m][pose_dim]  pose_dim 4: (x,y,z,yaw); 7: (x,y,z,qx,qy,qz,qw)
 *   off  [D][3]
 *   wp   [F*D][m][K]       K = 3 (position) or 4 (position + yaw); trajectory index
 *                          f*D + d, ready for mst_solve_batch(share_time_group = D)
 */
int mst_formation_waypoints(const double* rb, int F, int m, int pose_dim,
                            const double* off, int D, int K, double* wp, void* stream);

/*
 * Triangle meshes — Fcl_mesh (src/RigidBodyPlanners/fcl_checker.py:13-59).  `tri` is a
 * HOST pointer to [T][3][3] doubles (already rounded to 2 decimals by the caller as
 * load_stl does, :20-23); the mesh (triangles, per-triangle AABBs, root AABB) is
 * copied to the current CUDA device.
 */
typedef struct mst_mesh* mst_mesh_t;
int mst_mesh_create(const double* tri, int T, mst_mesh_t* out);
int mst_mesh_destroy(mst_mesh_t mesh);
int mst_mesh_triangle_count(mst_mesh_t mesh);

/*
 * Batched collision query — Fcl_checker.set_robot_transform + check_collision
 * (src/RigidBodyPlanners/fcl_checker.py:93-103) as called by isStateValid
 * (src/RigidBodyPlanners/RB_planning_sep_coll_check.py:208-215): robot mesh at each
 * pose against the environment mesh at identity; hit = 1 iff some triangle pair
 * intersects (touching counts).
 *   pose [P][pose_dim]  pose_dim 4: (x,y,z,yaw) or 7: (x,y,z,qx,qy,qz,qw)
 *   hit  [P]
 */
int mst_collide_poses(mst_mesh_t robot, mst_mesh_t env, const double* pose, int P,
                      int pose_dim, uint8_t* hit, void* stream);

/*
 * One collision query with HOST arguments, synchronous — the latency path of
 * Fcl_checker.check_collision when OMPL calls isStateValid one state at a time
 * (src/RigidBodyPlanners/RB_planning_sep_coll_check.py:208-226).  `pose` (pose_dim doubles, as in
 * mst_collide_poses) and `hit` are HOST pointers; the call returns after the answer (0 / 1) is in
 * *hit.  One launch + one stream synchronisation on a per-thread stream, pose and answer travel
 * through mapped pinned memory (allocated on a thread's first call: the one exception to "no
 * allocation after mst_mesh_create"); safe to call from several threads.  Meshes of any size.
 */
int mst_collide_pose_sync(mst_mesh_t robot, mst_mesh_t env, const double* pose, int pose_dim, int* hit);

/*
 * Batched motion validation for the sampling planner: M candidate motions
 * between states (x, y, z, yaw); each is checked at `steps` states interpolated linearly at
 * fractions j/steps, j = 1..steps (end state included, start state assumed valid) — what OMPL's
 * discrete motion validator does one isStateValid callback at a time at the resolution set in
 * src/RigidBodyPlanners/RB_planning_sep_coll_check.py:79.
 *   state_a, state_b [M][4]  ->  invalid [M]  (1 iff some interpolated state collides)
 */
int mst_collide_motions(mst_mesh_t robot, mst_mesh_t env, const double* state_a, const double* state_b,
                        int M, int steps, uint8_t* invalid, void* stream);

/*
 * Collision-check already solved trajectories (the second kernel of the pipeline on its own):
 * sample S uniform times per trajectory, robot mesh at each sampled position (yaw = 4th axis
 * when K = 4, else 0), flags as in mst_pipeline.  K must be 3 or 4.
 *   coef [B][n][K][8], dur [B][n]  ->  hit [B][S], any_hit [B]
 */
int mst_collide_trajectories(const double* coef, const double* dur, int B, int n, int K, int S,
                             mst_mesh_t robot, mst_mesh_t env, uint8_t* hit, uint8_t* any_hit,
                             void* stream);

/*
 * Fused pipeline: solve -> sample S uniform times -> place the robot mesh at every
 * sampled position (yaw = sampled 4th axis when K = 4, else 0) -> collide.
 * Two launches (solver, then sample + collide) or, with MST_SOLVER_AUTO_ONE_PASS, one persistent
 * kernel (see the solver enum).  In the two-launch form the solver also bounds every piece it solves
 * (Bernstein hull of its positions) and the sampling kernel skips the pieces that provably stay clear
 * of the obstacles' root box — their samples are flagged 0 without being evaluated.  Results are
 * identical in all forms.
 *   inputs / coef / dur / info as mst_solve_batch
 *   hit     [B][S]  per-sample collision flag
 *   any_hit [B]     1 iff any sample of the trajectory collides
 *   workspace       mst_pipeline_workspace_bytes(B, n, K, G, S) bytes
 */
size_t mst_pipeline_workspace_bytes(int B, int n, int K, int share_time_group, int S);
/* number of kernel launches one mst_pipeline call with these sizes issues (for accounting) */
int mst_pipeline_launch_count(int B, int n, int K, int share_time_group, int solver, int S);
int mst_pipeline(const double* wp, const double* t, int B, int n, int K,
                 int share_time_group, int solver, int S, mst_mesh_t robot,
                 mst_mesh_t env, double* coef, double* dur, int* info, uint8_t* hit,
                 uint8_t* any_hit, void* workspace, void* stream);

/*
 * mst_pipeline with the float32 polynomial matrix as an additional output — what path_to_pol emits per
 * trajectory (scripts/drones_pols_generator.py:63-77): pol_matrix [B][n][1 + 8K] rows
 * [T | x0..x7 | y0..y7 | z0..z7 (| yaw0..yaw7)].  The solver kernel writes the rows while the coefficients
 * are in registers (no packing pass over HBM); identical to mst_pack_pol_matrix of the results.
 */
int mst_pipeline_packed(const double* wp, const double* t, int B, int n, int K, int share_time_group, int solver,
                        int S, mst_mesh_t robot, mst_mesh_t env, double* coef, double* dur, int* info,
                        uint8_t* hit, uint8_t* any_hit, float* pol_matrix, void* workspace, void* stream);

/*
 * Measurement hook: one stage of the default (two-launch, far-piece culling) pipeline on its own, so that
 * a benchmark can time its kernels separately with CUDA events.  stage 1 = the solver launches (they also
 * leave the far-piece words in `workspace`), stage 2 = the sampling / collision launch on what a stage-1
 * call with the same arguments left behind, stage 0 = mst_pipeline.  MST_ERR_TOO_LARGE when the sizes do
 * not take the culling pipeline (the stages are then mst_solve_batch and mst_collide_trajectories).
 */
int mst_pipeline_stage(int stage, const double* wp, const double* t, int B, int n, int K, int share_time_group,
                       int solver, int S, mst_mesh_t robot, mst_mesh_t env, double* coef, double* dur, int* info,
                       uint8_t* hit, uint8_t* any_hit, void* workspace, void* stream);

/*
 * Fused pipeline with WIRE outputs for the multi-GPU gather ("all-gather ... only for
 * the final coefficients and collision flags").  Besides the local results of mst_pipeline, the
 * kernel stores, tile by tile while it computes, what the reference's path_to_pol emits — the
 * float32 polynomial matrix [T | x0..x7 | y0..y7 | z0..z7 (| yaw0..yaw7)] per piece
 * (scripts/drones_pols_generator.py:63-77) — and / or the collision flags through `count` base
 * pointers: this rank's own gather buffer and the NVLink peer mappings of the other ranks'
 * buffers (e.g. torch.distributed._symmetric_memory).  Rows [row_offset, row_offset + B) of every
 * buffer are this rank's.  The stores to peer pointers ARE the all-gather; the caller closes the
 * step with a cross-rank barrier.  Solver selection is MST_SOLVER_AUTO.
 *   pol_matrix[i] -> [rows][n][1 + 8K] float32, hit[i] -> [rows][S], any_hit[i] -> [rows]
 *   (pol_matrix == NULL or hit == NULL / any_hit == NULL: that output is not gathered)
 * The three pointer arrays live in HOST memory; the pointers in them are device pointers.
 */
#define MST_WIRE_MAX_TARGETS 8
typedef struct mst_wire_targets {
  int count;
  float* const* pol_matrix;
  uint8_t* const* hit;
  uint8_t* const* any_hit;
  long long row_offset;
} mst_wire_targets;
int mst_pipeline_wire(const double* wp, const double* t, int B, int n, int K, int share_time_group, int S,
                      mst_mesh_t robot, mst_mesh_t env, double* coef, double* dur, int* info, uint8_t* hit,
                      uint8_t* any_hit, const mst_wire_targets* wire, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MST_H_ */
