// Formation rigid-body transform: transform(path) of scripts/drones_traj_generator.py:56-89.
// For every rigid-body pose and drone offset (identity orientation, :27-38):
//   p_d = R(q_rb) * offset_d + t_rb        (tf2 do_transform_pose, :77-82)
// and, when a 4th axis is requested, the heading path_to_pol later extracts from the
// drone pose's quaternion (scripts/drones_pols_generator.py:51-53) — q_rb itself, since
// q_rb (x) identity = q_rb.
#include "mst_common.cuh"

namespace mst {

// euler_from_quaternion(q)[2] of tf.transformations ('sxyz'): atan2(M10, M00) unless the
// pitch is at the gimbal singularity, where the convention returns 0
__device__ __forceinline__ double yaw_of_quaternion(double x, double y, double z, double w) {
  const double nq = x * x + y * y + z * z + w * w;
  if (nq < 4.0 * 2.220446049250313e-16) return 0.0;  // quaternion_matrix returns identity
  const double s = sqrt(2.0 / nq);
  x *= s; y *= s; z *= s; w *= s;
  const double m00 = 1.0 - (y * y + z * z), m10 = x * y + z * w, m20 = x * z - y * w;
  (void)m20;
  const double cy = sqrt(m00 * m00 + m10 * m10);
  if (cy > 4.0 * 2.220446049250313e-16) return atan2(m10, m00);
  return 0.0;
}

__global__ void __launch_bounds__(256)
formation_kernel(const double* __restrict__ rb, long long poses, int m, int pose_dim,
                 const double* __restrict__ off, int D, int K, double* __restrict__ wp) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < poses * D;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long pi = idx / D;      // pose index f*m + i
    const int d = (int)(idx - pi * D);
    const long long f = pi / m;
    const int i = (int)(pi - f * m);
    const double* ps = rb + pi * pose_dim;
    double R[9], T[3];
    pose_to_transform(ps, pose_dim, R, T);
    const double* o = off + 3 * d;
    double* out = wp + (((f * D + d) * m) + i) * K;
    out[0] = R[0] * o[0] + R[1] * o[1] + R[2] * o[2] + T[0];
    out[1] = R[3] * o[0] + R[4] * o[1] + R[5] * o[2] + T[1];
    out[2] = R[6] * o[0] + R[7] * o[1] + R[8] * o[2] + T[2];
    if (K == 4) out[3] = pose_dim == 4 ? ps[3] : yaw_of_quaternion(ps[3], ps[4], ps[5], ps[6]);
  }
}

int launch_formation(const double* rb, int F, int m, int pose_dim, const double* off, int D, int K,
                     double* wp, cudaStream_t stream) {
  const long long poses = (long long)F * m;
  if (poses == 0 || D == 0) return MST_OK;
  long long blocks = (poses * D + 255) / 256;
  const long long cap = (long long)MST_SM_COUNT * 16;
  if (blocks > cap) blocks = cap;
  formation_kernel<<<(unsigned)blocks, 256, 0, stream>>>(rb, poses, m, pose_dim, off, D, K, wp);
  return check_launch();
}

}  // namespace mst
