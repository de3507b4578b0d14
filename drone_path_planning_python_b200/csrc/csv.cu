// Batched CSV text of polynomial matrices — what np.savetxt(file, matrix, delimiter=",") writes in
// path_to_pol (scripts/drones_pols_generator.py:79-81): per piece one line of 1 + 8K fields in
// numpy's default '%.18e' format, comma separated, '\n' terminated.  One CTA per trajectory; the
// digits come from exact integer arithmetic (csv_core.cuh), so the bytes equal numpy's.
#include "csv_core.cuh"
#include "mst_common.cuh"

namespace mst {

constexpr int CSV_FIELD_MAX = 25;   // "-d.dddddddddddddddddde+XX"

__global__ void __launch_bounds__(256)
csv_kernel(const float* __restrict__ mat, int n, int width, char* __restrict__ text, long long stride,
           int* __restrict__ length) {
  extern __shared__ int row_start[];   // [n + 1] byte offset of every line of this trajectory
  const long long b = blockIdx.x;
  const float* m = mat + b * (long long)n * width;
  char* out = text + b * stride;
  auto field_len = [](float v) -> int {
    union { float f; unsigned u; } x;
    x.f = v;
    const bool neg = x.u >> 31;
    if (((x.u >> 23) & 0xffu) == 0xffu) return (x.u & 0x7fffffu) ? 3 : (neg ? 4 : 3);
    return neg ? 25 : 24;
  };
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    int len = width;   // width - 1 commas and the newline
    for (int c = 0; c < width; ++c) len += field_len(m[r * width + c]);
    row_start[r + 1] = len;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    row_start[0] = 0;
    for (int r = 0; r < n; ++r) row_start[r + 1] += row_start[r];
    length[b] = row_start[n];
  }
  __syncthreads();
  for (int f = threadIdx.x; f < n * width; f += blockDim.x) {
    const int r = f / width, c = f - r * width;
    int off = row_start[r] + c;
    for (int j = 0; j < c; ++j) off += field_len(m[r * width + j]);
    char buf[32];
    const int len = format_e18(m[f], buf);
    for (int i = 0; i < len; ++i) out[off + i] = buf[i];
    out[off + len] = c + 1 < width ? ',' : '\n';
  }
}

int launch_csv(const float* mat, int B, int n, int width, char* text, long long stride, int* length,
               cudaStream_t stream) {
  if (B == 0) return MST_OK;
  csv_kernel<<<B, 256, sizeof(int) * (size_t)(n + 1), stream>>>(mat, n, width, text, stride, length);
  return check_launch();
}

}  // namespace mst

extern "C" size_t mst_csv_stride(int n, int width) {
  if (n < 0 || width < 0) return 0;
  return (size_t)n * (size_t)width * (mst::CSV_FIELD_MAX + 1);
}

extern "C" int mst_format_pol_matrix_csv(const float* mat, int B, int n, int width, char* text, long long stride,
                                         int* length, void* stream) {
  if (B < 0 || n < 1 || width < 1 || stride < (long long)mst_csv_stride(n, width)) return MST_ERR_INVALID;
  if (B == 0) return MST_OK;
  if (!mat || !text || !length) return MST_ERR_INVALID;
  return mst::launch_csv(mat, B, n, width, text, stride, length, (cudaStream_t)stream);
}
