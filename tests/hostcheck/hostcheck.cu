// TEST INFRASTRUCTURE ONLY — never loaded by the product.
// Compiles the kernels' __host__ __device__ arithmetic (condensed_core.cuh, collide_core.cuh)
// for the HOST so `pytest -m "not gpu"` can check the algorithms against the oracle without a
// GPU.  The GPU tests exercise the same functions through libmst.so on the device.
#include <stdlib.h>
#include <vector>

#include "../../drone_path_planning_python_b200/csrc/collide_core.cuh"
#include "../../drone_path_planning_python_b200/csrc/condensed_core.cuh"
#include "../../drone_path_planning_python_b200/csrc/mesh_image.cuh"

using namespace mst;

extern "C" int hostcheck_condensed(const double* wp, const double* t, int groups, int n, int K, int G,
                                   int force, double* coef, int* cls_out) {
  if (K < 1 || K > 4) return -1;
  std::vector<double> scratch(condensed_slots(n, K) + 1);
  for (int g = 0; g < groups; ++g) {
    const double* tg = t + (size_t)g * (n + 1);
    double Tmin, Tmax;
    const int cls = classify_times(tg, n, &Tmin, &Tmax);
    cls_out[g] = cls;
    if (cls >= 2 || (cls == 1 && (!force || !(Tmin > 0.0) || tg[0] != 0.0))) continue;
    for (int i = 0; i < n; ++i) scratch[i] = tg[i + 1] - tg[i];
    condensed_factor(n, scratch.data(), 1);
    for (int d = 0; d < G; ++d) {
      const size_t traj = (size_t)g * G + d;
      const double* wpd = wp + traj * (size_t)(n + 1) * K;
      double* cd = coef + traj * (size_t)n * K * MST_NCOEF;
      auto emit = [&](int piece, int k, const double* c, double) {
        for (int e = 0; e < MST_NCOEF; ++e) cd[((size_t)piece * K + k) * MST_NCOEF + e] = c[e];
      };
      if (K <= 3) {
        condensed_forward<3>(wpd, n, K, scratch.data(), 1);
        condensed_backward<3>(wpd, n, K, scratch.data(), 1, emit);
      } else {
        condensed_forward<4>(wpd, n, K, scratch.data(), 1);
        condensed_backward<4>(wpd, n, K, scratch.data(), 1, emit);
      }
    }
  }
  return 0;
}

// robot_tri[Tr][9], env_tri[Te][9], env_box[Te][6], root[6]; R[P][9], T[P][3]
extern "C" int hostcheck_collide(const double* robot_tri, int Tr, const double* env_tri,
                                 const double* env_box, int Te, const double* root, double radius,
                                 const double* R, const double* T, int P, int rigid, unsigned char* hit) {
  for (int p = 0; p < P; ++p)
    hit[p] = robot_hits_env(R + 9 * (size_t)p, T + 3 * (size_t)p, robot_tri, Tr, env_tri, env_box, Te, root,
                            radius, rigid != 0) ? 1 : 0;
  return 0;
}

// the culled mesh-mesh routine the kernels use, on meshes given as triangle soups
extern "C" int hostcheck_collide_culled(const double* robot_tri, int Tr, const double* env_tri, int Te,
                                        const double* R, const double* T, int P, int rot, int rigid,
                                        unsigned char* hit) {
  MeshLayout rl, el;
  MeshBounds rbb, evb;
  void* ri = build_mesh_image(robot_tri, Tr, &rl, &rbb);
  void* ei = build_mesh_image(env_tri, Te, &el, &evb);
  if (!ri || !ei) return -1;
  const MeshView rb = mesh_view(ri, rl), ev = mesh_view(ei, el);
  for (int p = 0; p < P; ++p) {
    const double* Rp = R + 9 * (size_t)p;
    const double* Tp = T + 3 * (size_t)p;
    hit[p] = (rot ? robot_hits_env_culled<true>(Rp, Tp, rb, rbb, ev, evb, rigid != 0)
                  : robot_hits_env_culled<false>(Rp, Tp, rb, rbb, ev, evb, rigid != 0)) ? 1 : 0;
  }
  free(ri);
  free(ei);
  return rb.V;
}
