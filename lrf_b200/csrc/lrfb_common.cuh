// Shared plumbing for the lrf_b200 CUDA kernels (sm_100a).
//
// The same sources also compile with g++ -DLRFB_SIM against tests/cpu_sim/cuda_sim.h, a test-only
// SIMT-on-CPU shim used to check index math and arithmetic against the oracle where no GPU is
// present.  The product library is built by nvcc only and contains no CPU path.
#pragma once

#ifdef LRFB_SIM
#include "cuda_sim.h"
#define LRFB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  sim::launch(kernel, grid, block, smem, __VA_ARGS__)
#define LRFB_DYN_SMEM(name) unsigned char* name = (unsigned char*)sim::dyn_smem()
#else
#include <cuda_runtime.h>
#define LRFB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define LRFB_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

#include <stdint.h>

namespace lrfb {

constexpr float kEps = 1e-16f;  // CoordinateDescent eps (factorization/qmf.py:83), rounded to f32

// at::bmm's small-problem scalar path (no FMA, k ascending) is taken when K*rows*cols < 400;
// see oracle/qmf_exact.c header.  The kernels reproduce that choice where it changes the bits.
__host__ __device__ inline bool bmm_native(long long K, long long rows, long long cols) {
  return K * rows * cols < 400;
}

// ---- async copy (LDGSTS) -----------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
#ifdef LRFB_SIM
  memcpy(smem_dst, gmem_src, 16);
#else
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
#endif
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
#ifdef LRFB_SIM
  memcpy(smem_dst, gmem_src, 4);
#else
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src));
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef LRFB_SIM
  asm volatile("cp.async.commit_group;\n" ::);
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#ifndef LRFB_SIM
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
#endif
}

// FP32 FFMA throughput probe: 16 independent chains per thread, `iters` x 16 x 2 FMAs each.
// flops = gridDim.x * blockDim.x * iters * 64;  bench.py times it to get the FP32 roofline denominator.
__global__ void __launch_bounds__(256) ffma_probe_kernel(float* out, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f;
  const float m = 1.0000001f, c = 1e-7f, m2 = 0.9999999f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], m, c);
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], m2, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace lrfb
