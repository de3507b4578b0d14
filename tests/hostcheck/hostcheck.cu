// TEST INFRASTRUCTURE ONLY — never loaded by the product.
// Compiles the kernels' __host__ __device__ arithmetic (condensed_core.cuh, collide_core.cuh)
// for the HOST so `pytest -m "not gpu"` can check the algorithms against the oracle without a
// GPU.  The GPU tests exercise the same functions through libmst.so on the device.
#include <stdlib.h>
#include <vector>

#include "../../drone_path_planning_python_b200/csrc/banded_core.cuh"
#include "../../drone_path_planning_python_b200/csrc/collide_core.cuh"
#include "../../drone_path_planning_python_b200/csrc/condensed_core.cuh"
#include "../../drone_path_planning_python_b200/csrc/csv_core.cuh"
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
    double* rho = scratch.data();
    double* fac = rho + n;
    double* ys = fac + 6 * (n - 1);
    for (int i = 0; i < n; ++i) rho[i] = tg[i + 1] - tg[i];
    condensed_factor(n, rho, fac, 1);
    for (int d = 0; d < G; ++d) {
      const size_t traj = (size_t)g * G + d;
      const double* wpd = wp + traj * (size_t)(n + 1) * K;
      double* cd = coef + traj * (size_t)n * K * MST_NCOEF;
      auto emit = [&](int piece, int k, const double* c, double) {
        for (int e = 0; e < MST_NCOEF; ++e) cd[((size_t)piece * K + k) * MST_NCOEF + e] = c[e];
      };
      if (K <= 3) {
        condensed_forward<3>(wpd, K, n, K, rho, fac, 1, ys, 1);
        condensed_backward<3>(wpd, K, n, K, rho, fac, 1, ys, 1, emit);
      } else {
        condensed_forward<4>(wpd, K, n, K, rho, fac, 1, ys, 1);
        condensed_backward<4>(wpd, K, n, K, rho, fac, 1, ys, 1, emit);
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

// pairwise triangle tests: tri[N][2][3][3]; out17 = 17-axis SAT, outi = interval form
extern "C" int hostcheck_tri_pairs(const double* tri, int N, unsigned char* out17, unsigned char* outi) {
  for (int i = 0; i < N; ++i) {
    const double* t = tri + 18 * (size_t)i;
    const V3 P1 = {t[0], t[1], t[2]}, P2 = {t[3], t[4], t[5]}, P3 = {t[6], t[7], t[8]};
    const V3 Q1 = {t[9], t[10], t[11]}, Q2 = {t[12], t[13], t[14]}, Q3 = {t[15], t[16], t[17]};
    out17[i] = triangles_intersect(P1, P2, P3, Q1, Q2, Q3) ? 1 : 0;
    outi[i] = triangles_intersect_interval(P1, P2, P3, Q1, Q2, Q3) ? 1 : 0;
  }
  return 0;
}

// "%.18e" of float32 values: out[N][32] (zero padded), lengths[N]
extern "C" int hostcheck_format_e18(const float* v, int N, char* out, int* lengths) {
  for (int i = 0; i < N; ++i) {
    char buf[32] = {0};
    lengths[i] = format_e18(v[i], buf);
    for (int j = 0; j < 32; ++j) out[32 * (size_t)i + j] = buf[j];
  }
  return 0;
}

// The pivoted banded solver (banded_core.cuh) with the kernel's schedule replayed on the host: every
// phase between two warp barriers of banded_lu_kernel is a loop over the 32 lanes here.  Inputs are
// assumed valid (finite, non-decreasing stamps).  info[g] = 0 or the first zero pivot column + 1.
extern "C" int hostcheck_banded_lu(const double* wp, const double* t, int groups, int n, int K, int G,
                                   double* coef, int* info) {
  const int N = MST_NCOEF * n, R = G * K, NS = band_rhs_stride(n);
  double ff[64], cf[MST_NCOEF * LD];
  unsigned char pi[MST_NCOEF * LD];
  for (int e = 0; e < 64 + MST_NCOEF * LD; ++e) band_table_entry(e, ff, cf, pi);
  std::vector<double> W(WCOLS * LD), pw(MST_NCOEF * (n + 1)), Bs((size_t)R * NS), Ug((size_t)UROWS * N);
  const BandSystem sys{n, N, pw.data(), ff, cf, pi};
  for (int g = 0; g < groups; ++g) {
    const double* tg = t + (size_t)g * (n + 1);
    for (int i = 0; i <= n; ++i) band_powers(i < n ? tg[i + 1] - tg[i] : tg[0], pw.data() + MST_NCOEF * i);
    std::fill(Bs.begin(), Bs.end(), 0.0);
    std::fill(W.begin(), W.end(), -77.0);   // the last slot is garbage until a column is assembled into it
    std::fill(Ug.begin(), Ug.end(), -99.0);
    for (int e = 0; e < (KV + 1) * LD; ++e) {
      const int c = e / LD, o = e - c * LD;
      W[e] = c < N ? band_entry_fast(sys, c, o) : 0.0;
      if (c < N && band_entry_fast(sys, c, o) != band_entry(sys, c, o)) return -2;
    }
    for (int r = 0; r < R; ++r) {
      const int d = r / K, k = r - d * K;
      double* b = Bs.data() + (size_t)r * NS;
      for (int i = 0; i <= n; ++i) {
        const double v = wp[(((size_t)g * G + d) * (n + 1) + i) * K + k];
        if (i == 0) b[0] = v;
        else if (i == n) b[N - 4] = v;
        else b[4 + MST_NCOEF * (i - 1) + 6] = b[4 + MST_NCOEF * (i - 1) + 7] = v;
      }
    }
    int singular_at = 0;
    double rinv_prev = 0.0;
    auto slot_of = [&](int col) { return W.data() + ((col + WCOLS) % WCOLS) * LD; };   // column -> its ring slot
    for (int j = 0; j < N; ++j) {
      double l[KL + 1];
      int jp;
      double rinv;
      const bool ok = band_pivot<false>(slot_of(j) + KV, 0, l, jp, rinv);
      if (!ok && singular_at == 0) singular_at = j + 1;
      for (int lane = 0; lane < 32; ++lane) {
        double leaves, enters;
        band_retire_fetch(sys, slot_of(j - 1) + (lane < LD ? lane : LD - 1), j + KV + 1, lane, &leaves, &enters);
        if (ok) {
          if (lane < MAT_LANES) {
            const int c = j + 1 + lane;
            if (c < N) band_update(slot_of(c) + KV - (lane + 1), jp, l);
          } else {
            for (int q = lane - MAT_LANES; q < R; q += RHS_LANES) band_update(Bs.data() + (size_t)q * NS + j, jp, l);
          }
        }
        if (lane < LD && j + KV + 1 < N && band_entry_fast(sys, j + KV + 1, lane) != band_entry(sys, j + KV + 1, lane))
          return -2;   // the pattern tables must reproduce the rules they were derived from
        band_retire_store(sys, slot_of(j + KV + 1) + (lane < LD ? lane : LD - 1),
                          Ug.data() + (size_t)(j > 0 ? j - 1 : 0) * UROWS + (lane < KV ? lane : KV), j == 0, j + KV + 1,
                          lane, rinv_prev, leaves, enters);
      }
      rinv_prev = rinv;
    }
    for (int lane = 0; lane <= KV; ++lane)
      Ug[(size_t)(N - 1) * UROWS + lane] = lane == KV ? rinv_prev : slot_of(N - 1)[lane];
    for (int j = N - 1; j >= 0; --j) {
      const double rinv = Ug[(size_t)j * UROWS + KV];
      for (int r = 0; r < R; ++r) {
        const int dd = r / K, k = r - dd * K;
        // lanes in DESCENDING order: a lane must not see what a lower lane wrote in this phase
        for (int lane = 31; lane >= 0; --lane) {
          const int d = (lane < KV && lane < j) ? lane + 1 : 0;
          const double u = d > 0 ? Ug[(size_t)j * UROWS + KV - d] : 0.0;
          band_backsub(Bs.data() + (size_t)r * NS + j, d, lane, r % 32, u, rinv,
                       coef + ((((size_t)g * G + dd) * n + (j >> 3)) * K + k) * MST_NCOEF + (j & 7));
        }
      }
    }
    for (int d = 0; d < G; ++d) info[(size_t)g * G + d] = singular_at;
  }
  return 0;
}

// TileWalk (condensed_core.cuh) against the direct index arithmetic: slots[e] for e = lane, lane + 32, ...
extern "C" int hostcheck_tile_walk(int n, int K, int WS, int total, int* slots) {
  for (int lane = 0; lane < 32; ++lane) {
    TileWalk w = TileWalk::start(lane, n, K);
    for (int e = lane; e < total; e += 32) {
      slots[e] = w.slot(WS, K);
      w.advance32(n, K);
    }
  }
  return 0;
}
