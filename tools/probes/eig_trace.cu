// Dev probe: cycle timeline of eig64_topr_kernel (block 0) on random SPD 64x64 Gram matrices; grid sized like a
// 2048-image batch so the SM is as loaded as in the encode.
#define LRFB_EIG_TRACE 1
#include "../../lrf_b200/csrc/eig.cuh"
#include <cstdio>
#include <vector>
#include <random>
using namespace lrfb;
int main() {
  const int B = 2048, N = 64, R = 4;
  std::vector<double> g((size_t)B * N * N);
  std::mt19937 rng(1);
  std::normal_distribution<double> d(0.0, 1.0);
  std::vector<double> x(256 * N);
  for (int b = 0; b < B; ++b) {
    if (b < 8) {
      for (auto& v : x) v = d(rng) + 3.0;
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
          double s = 0;
          for (int m = 0; m < 256; ++m) s += x[m * N + i] * x[m * N + j];
          g[(size_t)b * N * N + i * N + j] = s;
        }
    } else {
      std::copy(g.begin() + (size_t)(b % 8) * N * N, g.begin() + (size_t)(b % 8 + 1) * N * N, g.begin() + (size_t)b * N * N);
    }
  }
  double *G, *evec, *sigma;
  float *v0, *s0;
  cudaMalloc(&G, g.size() * 8), cudaMalloc(&evec, (size_t)B * N * R * 8), cudaMalloc(&sigma, (size_t)B * R * 8);
  cudaMalloc(&v0, (size_t)B * N * R * 4), cudaMalloc(&s0, (size_t)B * R * 4);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemcpy(G, g.data(), g.size() * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    eig64_topr_kernel<<<B, 64>>>(G, R, evec, sigma, nullptr, 6144, v0, s0, nullptr, 0);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("eig64 x %d: %.3f ms (%s)\n", B, ms, cudaGetErrorString(e));
  }
  long long h[8];
  cudaMemcpyFromSymbol(h, g_eig_trace, sizeof(h));
  printf("block 0 cycles: tridiagonalisation %lld | multisection %lld | inverse iteration %lld | MGS %lld | back-transform %lld | output %lld | total %lld\n",
         h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[6] - h[0]);
  return 0;
}
