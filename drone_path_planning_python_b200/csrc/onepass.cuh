// Parameter blocks of the single-pass pipeline kernel (pipeline_onepass.cu), shared with api.cu.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "mst_common.cuh"

namespace mst {

// Extra destinations of the per-trajectory wire outputs (multi-GPU gather by peer stores): up to 8
// base pointers each — this rank's own gather buffer and its NVLink peer mappings of the other
// ranks' buffers; rows [row0, row0 + B) of every buffer belong to this rank.
struct WireTargets {
  int count;
  float* mat[8];        // [rows][n][1 + 8K] float32 polynomial matrix (may be null entries: skipped)
  uint8_t* hit[8];      // [rows][S]
  uint8_t* any[8];      // [rows]
  long long row0;
};

// shared-memory plan of one warp of the kernel (byte offsets from the warp's base)
struct OnepassLayout {
  int GPW, TPT;                 // time groups / trajectories per warp tile
  size_t cbuf_doubles;          // staged coefficients of two trajectories
  size_t off_rho, off_fac, off_y, off_t, off_w, off_hit, off_any, off_piece, off_wire, off_ring, bytes;
  size_t shared_bytes;          // CTA-wide part in front of the warps' regions (mesh images, plane x vertex table)
};

}  // namespace mst
