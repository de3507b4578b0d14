// Host-side construction of the mesh image (layout in mst_common.cuh): unique vertices,
// corner indices, per-triangle boxes / planes / vertex masks, whole-mesh bounds.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mst_common.cuh"

namespace mst {

// 30-bit Morton code of a point in the unit cube (10 bits per axis)
inline unsigned morton3(double x, double y, double z) {
  auto spread = [](double v) -> unsigned {
    double s = v * 1024.0;
    unsigned q = s <= 0.0 ? 0u : (s >= 1023.0 ? 1023u : (unsigned)s);
    q = (q | (q << 16)) & 0x030000FFu;
    q = (q | (q << 8)) & 0x0300F00Fu;
    q = (q | (q << 4)) & 0x030C30C3u;
    q = (q | (q << 2)) & 0x09249249u;
    return q;
  };
  return spread(x) | (spread(y) << 1) | (spread(z) << 2);
}

// returns a malloc'ed image of layout->bytes (>= 16) bytes, or NULL when out of memory
inline void* build_mesh_image(const double* tri_in, int T, MeshLayout* layout, MeshBounds* bounds) {
  const size_t nt = (size_t)(T > 0 ? T : 1);
  // Triangles are stored in Morton order of their centroids when there is more than one block of 32:
  // the blocks the collision cursor walks are then spatially compact and their boxes (bbox) cull
  // well.  A collision answer does not depend on the order (any intersecting pair is a hit).
  double* sorted = (double*)malloc(sizeof(double) * 9 * nt);
  if (!sorted) return NULL;
  memcpy(sorted, tri_in, sizeof(double) * 9 * (size_t)T);
  if (T > 32) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int c = 0; c < 3 * T; ++c)
      for (int k = 0; k < 3; ++k) {
        const double v = tri_in[3 * c + k];
        if (v < lo[k]) lo[k] = v;
        if (v > hi[k]) hi[k] = v;
      }
    struct Key { unsigned code; int t; };
    Key* keys = (Key*)malloc(sizeof(Key) * nt);
    if (!keys) { free(sorted); return NULL; }
    for (int t = 0; t < T; ++t) {
      double c[3];
      for (int k = 0; k < 3; ++k) {
        const double m = (tri_in[9 * t + k] + tri_in[9 * t + 3 + k] + tri_in[9 * t + 6 + k]) / 3.0;
        c[k] = hi[k] > lo[k] ? (m - lo[k]) / (hi[k] - lo[k]) : 0.0;
      }
      keys[t].code = morton3(c[0], c[1], c[2]);
      keys[t].t = t;
    }
    qsort(keys, (size_t)T, sizeof(Key), [](const void* a, const void* b) -> int {
      const Key* x = (const Key*)a; const Key* y = (const Key*)b;
      if (x->code != y->code) return x->code < y->code ? -1 : 1;
      return x->t < y->t ? -1 : (x->t > y->t ? 1 : 0);
    });
    for (int t = 0; t < T; ++t) memcpy(sorted + 9 * (size_t)t, tri_in + 9 * (size_t)keys[t].t, sizeof(double) * 9);
    free(keys);
  }
  const double* tri = sorted;
  int* idx = (int*)malloc(sizeof(int) * 3 * nt);
  double* uv = (double*)malloc(sizeof(double) * 9 * nt);
  if (!idx || !uv) { free(idx); free(uv); free(sorted); return NULL; }
  // unique vertices by exact equality (what the reference's indexed triangles are built on,
  // src/RigidBodyPlanners/fcl_checker.py:28-40)
  int V = 0;
  for (int c = 0; c < 3 * T; ++c) {
    const double* p = tri + 3 * c;
    int found = -1;
    for (int v = 0; v < V && found < 0; ++v)
      if (uv[3 * v] == p[0] && uv[3 * v + 1] == p[1] && uv[3 * v + 2] == p[2]) found = v;
    if (found < 0) { found = V; uv[3 * V] = p[0]; uv[3 * V + 1] = p[1]; uv[3 * V + 2] = p[2]; ++V; }
    idx[c] = found;
  }
  *layout = mesh_layout(T, V);
  char* base = (char*)calloc(1, layout->bytes > 0 ? layout->bytes : 16);
  if (!base) { free(idx); free(uv); free(sorted); return NULL; }
  double* itri = (double*)base;
  double* ibox = (double*)(base + layout->off_box);
  double* ipl = (double*)(base + layout->off_plane);
  double* ivert = (double*)(base + layout->off_vert);
  int* iidx = (int*)(base + layout->off_idx);
  unsigned long long* imask = (unsigned long long*)(base + layout->off_mask);
  unsigned* ivtri = (unsigned*)(base + layout->off_vtri);
  float* ifbox = (float*)(base + layout->off_fbox);
  double* iedge = (double*)(base + layout->off_edge);
  float* ibbox = (float*)(base + layout->off_bbox);
  for (int b = 0; b < (T + 31) / 32; ++b)
    for (int k = 0; k < 3; ++k) { ibbox[8 * b + k] = INFINITY; ibbox[8 * b + 3 + k] = -INFINITY; }
  for (int t = 0; t < T && t < 32; ++t)
    for (int c = 0; c < 3; ++c) ivtri[idx[3 * t + c]] |= 1u << t;
  memcpy(itri, tri, sizeof(double) * 9 * (size_t)T);
  memcpy(ivert, uv, sizeof(double) * 3 * (size_t)V);
  memcpy(iidx, idx, sizeof(int) * 3 * (size_t)T);
  for (int k = 0; k < 3; ++k) { bounds->root[k] = 1e300; bounds->root[3 + k] = -1e300; }
  bounds->radius = 0.0;
  bounds->rxy = 0.0;
  for (int t = 0; t < T; ++t) {
    double* box = ibox + 6 * t;
    for (int k = 0; k < 3; ++k) { box[k] = 1e300; box[3 + k] = -1e300; }
    unsigned long long mk = 0ull;
    for (int c = 0; c < 3; ++c) {
      double r2 = 0.0;
      for (int k = 0; k < 3; ++k) {
        const double v = tri[9 * t + 3 * c + k];
        if (v < box[k]) box[k] = v;
        if (v > box[3 + k]) box[3 + k] = v;
        r2 += v * v;
      }
      // rounded up a little: the sphere cull must stay conservative
      const double r = sqrt(r2) * (1.0 + 1e-12) + 1e-300;
      if (r > bounds->radius) bounds->radius = r;
      const double vx = tri[9 * t + 3 * c], vy = tri[9 * t + 3 * c + 1];
      const double rp = sqrt(vx * vx + vy * vy) * (1.0 + 1e-12) + 1e-300;
      if (rp > bounds->rxy) bounds->rxy = rp;
      if (idx[3 * t + c] < 64) mk |= 1ull << idx[3 * t + c];
    }
    imask[t] = mk;
    for (int k = 0; k < 3; ++k) {
      float lo = (float)box[k], hi = (float)box[3 + k];
      if ((double)lo > box[k]) lo = nextafterf(lo, -INFINITY);
      if ((double)hi < box[3 + k]) hi = nextafterf(hi, INFINITY);
      ifbox[8 * t + k] = lo;
      ifbox[8 * t + 3 + k] = hi;
      float* bb = ibbox + 8 * (t / 32);
      if (lo < bb[k]) bb[k] = lo;
      if (hi > bb[3 + k]) bb[3 + k] = hi;
    }
    for (int k = 0; k < 3; ++k) {
      if (box[k] < bounds->root[k]) bounds->root[k] = box[k];
      if (box[3 + k] > bounds->root[3 + k]) bounds->root[3 + k] = box[3 + k];
    }
    // plane through the triangle: n = (Q2-Q1) x (Q3-Q2), d = n . Q1
    const double* q = tri + 9 * t;
    const double f1[3] = {q[3] - q[0], q[4] - q[1], q[5] - q[2]};
    const double f2[3] = {q[6] - q[3], q[7] - q[4], q[8] - q[5]};
    double* pl = ipl + 4 * t;
    pl[0] = f1[1] * f2[2] - f1[2] * f2[1];
    pl[1] = f1[2] * f2[0] - f1[0] * f2[2];
    pl[2] = f1[0] * f2[1] - f1[1] * f2[0];
    pl[3] = pl[0] * q[0] + pl[1] * q[1] + pl[2] * q[2];
    // the three planes through the edges, perpendicular to the triangle: m_k = +-n x (Q_{k+1} - Q_k)
    // pointing away from the opposite corner.  A point set with m_k . x > m_k . Q_k throughout is
    // separated from the (closed) triangle.  The offset carries a slack far above the rounding of
    // the dot products and far below any length that matters, so the cull errs towards testing.
    double scale = 1.0;
    for (int c = 0; c < 9; ++c) if (fabs(q[c]) > scale) scale = fabs(q[c]);
    for (int k = 0; k < 3; ++k) {
      const double* a = q + 3 * k;
      const double* b = q + 3 * ((k + 1) % 3);
      const double* o = q + 3 * ((k + 2) % 3);
      const double ed[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
      double m[3] = {pl[1] * ed[2] - pl[2] * ed[1], pl[2] * ed[0] - pl[0] * ed[2], pl[0] * ed[1] - pl[1] * ed[0]};
      if (m[0] * (o[0] - a[0]) + m[1] * (o[1] - a[1]) + m[2] * (o[2] - a[2]) > 0.0) {
        m[0] = -m[0]; m[1] = -m[1]; m[2] = -m[2];
      }
      double* out = iedge + 12 * t + 4 * k;
      out[0] = m[0]; out[1] = m[1]; out[2] = m[2];
      out[3] = (m[0] * a[0] + m[1] * a[1] + m[2] * a[2]) + 1e-10 * (fabs(m[0]) + fabs(m[1]) + fabs(m[2])) * 2.0 * scale;
    }
  }
  free(idx);
  free(uv);
  free(sorted);
  return base;
}

}  // namespace mst
