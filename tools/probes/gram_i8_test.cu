// Dev probe: gram64_i8_kernel (tcgen05 int8 exact Gram) vs an exact host computation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I lrf_b200/csrc -o tools/probes/gram_i8_test tools/probes/gram_i8_test.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#define LRFB_I8_DEBUG 1
#include "gram_i8.cuh"
using namespace lrfb;
int main(int argc, char** argv) {
  int M = argc > 1 ? atoi(argv[1]) : 6144, n_mat = argc > 2 ? atoi(argv[2]) : 4, split = argc > 3 ? atoi(argv[3]) : 1;
  std::vector<float> hx((size_t)n_mat * M * 64);
  unsigned s = 12345;
  for (auto& v : hx) { s = s * 1664525u + 1013904223u; v = (float)(s >> 8) * (255.5f / 16777216.0f); }
  float* dx; double* dg;
  cudaMalloc(&dx, hx.size() * 4); cudaMalloc(&dg, (size_t)n_mat * split * 4096 * 8);
  cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
  size_t smem = sizeof(GramI8Smem) + 1024;
  cudaFuncSetAttribute(gram64_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int mode = argc > 4 ? atoi(argv[4]) : 0;
  cudaMemcpyToSymbol(g_i8_mode, &mode, sizeof(int));
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    gram64_i8_kernel<<<dim3(split, n_mat), kI8Threads, smem>>>(dx, (long long)M * 64, M, dg, split);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("launch %d: %s, %.3f ms\n", rep, cudaGetErrorString(err), ms);
    if (err != cudaSuccess) return 1;
  }
  std::vector<double> hg((size_t)n_mat * split * 4096);
  cudaMemcpy(hg.data(), dg, hg.size() * 8, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int m = 0; m < (n_mat < 2 ? n_mat : 2); ++m) {
    for (int i = 0; i < 64; i += 7) for (int j = 0; j < 64; j += 5) {
      // exact: products of the Q8.24 truncations in long double / __int128
      __int128 acc = 0;
      for (int r = 0; r < M; ++r) {
        unsigned long long a = (unsigned long long)(hx[((size_t)m * M + r) * 64 + i] * 16777216.0f);
        unsigned long long b = (unsigned long long)(hx[((size_t)m * M + r) * 64 + j] * 16777216.0f);
        acc += (__int128)a * b;
      }
      double ref = (double)((long double)acc / 281474976710656.0L);  // 2^48
      double got = 0; for (int sp = 0; sp < split; ++sp) got += hg[((size_t)m * split + sp) * 4096 + i * 64 + j];
      double rel = fabs(got - ref) / fabs(ref);
      if (rel > worst) worst = rel;
      if (rel > 1e-12) { printf("MISMATCH m=%d (%d,%d): got %.17g ref %.17g rel %.3g\n", m, i, j, got, ref, rel); }
    }
  }
  printf("worst relative error %.3g\n", worst);
  return 0;
}
