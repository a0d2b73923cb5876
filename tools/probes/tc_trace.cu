// Dev probe: cycle timeline of one CTA of bcd_tc_kernel<4,768,384> over the sweeps of its first matrix
// (clock64 at 9 points per sweep).  nvcc -DLRFB_TC_TRACE; synthetic X in [0,256), 64 matrices of 6144 x 64.
#define LRFB_TC_TRACE 1
#include "../../lrf_b200/csrc/bcd_tc.cuh"
#include <cstdio>
#include <vector>
#include <random>
using namespace lrfb;
int main(int argc, char** argv) {
  const bool chroma = argc > 1 && atoi(argv[1]) == 2;  // 2: chroma shape (1536 rows, R = 2, clusters of 2)
  const int B = chroma ? 592 : 240, M = chroma ? 1536 : 6144, N = 64, R = chroma ? 2 : 4;
  std::vector<float> hx((size_t)B * M * N), hv((size_t)B * N * R), hu((size_t)B * M * R);
  std::mt19937 rng(1);
  std::uniform_real_distribution<float> d(0.0f, 255.0f), dv(-1.0f, 1.0f);
  for (auto& x : hx) x = d(rng);
  for (auto& x : hv) x = dv(rng) * 4.0f;
  for (auto& x : hu) x = dv(rng) * 4.0f;
  float *X, *U, *V;
  cudaMalloc(&X, hx.size() * 4), cudaMalloc(&U, hu.size() * 4), cudaMalloc(&V, hv.size() * 4);
  cudaMemcpy(X, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(U, hu.data(), hu.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(V, hv.data(), hv.size() * 4, cudaMemcpyHostToDevice);
  BcdBatch b = {};
  b.X = X, b.x_stride = (long long)M * N, b.U = U, b.V = V, b.M = M, b.n_mat = B, b.num_iters = 10, b.lo = -16, b.hi = 15;
  b.x_u8_range = 1;
  int* counter;
  cudaMalloc(&counter, 4);
  b.work_counter = counter;
  const bool small = argc > 1 && atoi(argv[1]) == 1;  // 1: 384 rows x 192 threads, clusters of 16, 2 CTAs per SM
  auto kern = chroma ? bcd_tc_kernel<2, 768, 384> : small ? bcd_tc_kernel<4, 384, 192> : bcd_tc_kernel<4, 768, 384>;
  const size_t smem = chroma ? sizeof(TcSmem<2, 768, 384>) : small ? sizeof(TcSmem<4, 384, 192>) : sizeof(TcSmem<4, 768, 384>);
  const int csize = chroma ? 2 : small ? 16 : 8, nt = small ? 192 : 384, rows = small ? 384 : 768;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(nt), cfg.dynamicSmemBytes = smem, cfg.attrs = attr, cfg.numAttrs = 1, cfg.gridDim = dim3(chroma ? 148 : small ? 224 : 120);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(counter, 0, 4);
    CUtensorMap x_map = {};
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, b, csize, rows, x_map, 0);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("error %s %s\n", cudaGetErrorString(e), cudaGetErrorString(e2)); return 1; }
  }
  long long h[16 * 12];
  cudaMemcpyFromSymbol(h, g_tc_trace, sizeof(h));
  const char* names[14] = {"sweep start", "A-phase done", "GS done", "GS done + MMAs issued", "mma_done seen", "pushes issued", "full seen", "V gathered + gram", "after CTA barrier",
                           "proxy fence done", "tmem ld done", "combined", "slots summed", "vfull seen"};
  const int order[11] = {1, 3, 4, 10, 11, 5, 6, 12, 13, 7, 8};
  printf("cycles relative to sweep start (CTA 0, thread 0); last column = sweep length\n");
  for (int it = 0; it < 10; ++it) {
    printf("sweep %d:", it);
    for (int q = 0; q < 11; ++q) printf(" %s %lld |", names[order[q]], h[order[q] * 12 + it] - h[it]);
    if (it < 9) printf(" next start %lld", h[it + 1] - h[it]);
    printf("\n");
  }
  const long long* t = h + 14 * 12;
  printf("second matrix of CTA 0: V + barrier %lld | loads issued %lld | gram_small %lld | X landed %lld | byte planes in TMEM %lld | register row + barrier %lld | 10 sweeps %lld\n",
         t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5], t[7] - t[6]);
  return 0;
}
