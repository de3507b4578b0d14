// FP64 FMA peak microbenchmark (measurement tool, not part of libmst): every thread runs CHAINS
// independent dependent-FMA chains from registers, so the FP64 pipe is the only limiter.
// flops = 2 * blocks * threads * CHAINS * iters.  Timed from Python with CUDA events
// (tools/fp64_peak.py); SURVEY §8d asks for a MEASURED FP64 peak as the roofline denominator.
#include <cuda_runtime.h>

constexpr int CHAINS = 8;

__global__ void __launch_bounds__(256) fp64_fma_kernel(int iters, double a, double b, double* out) {
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = 1.0 + 1e-9 * (threadIdx.x + c);
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = __fma_rn(x[c], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c];
  if (s == 123.456) out[0] = s;   // never true: keeps the chains alive
}

// non-fused form: one multiply and one add per step, as the bit-exact Horner evaluation issues them
__global__ void __launch_bounds__(256) fp64_muladd_kernel(int iters, double a, double b, double* out) {
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = 1.0 + 1e-9 * (threadIdx.x + c);
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = __dadd_rn(__dmul_rn(x[c], a), b);
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c];
  if (s == 123.456) out[0] = s;
}

extern "C" int probe_chains(void) { return CHAINS; }

extern "C" int probe_fp64(int kind, int blocks, int threads, int iters, double* out, void* stream) {
  if (kind == 0)
    fp64_fma_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, 0.999999, 1e-7, out);
  else
    fp64_muladd_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, 0.999999, 1e-7, out);
  return (int)cudaGetLastError();
}
