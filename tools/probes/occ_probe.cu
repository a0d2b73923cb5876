// Dev probe: do two CTAs that both tcgen05.alloc (256 columns each) really share one SM?  The occupancy
// calculator says 1 block/SM for any kernel with tcgen05.alloc; this measures the actual placement.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <map>
extern __shared__ float dyn[];
template <int COLS>
__global__ void __launch_bounds__(128, 2) k(unsigned long long* rec, long long spin_ns) {
  __shared__ unsigned base;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (COLS > 0 && threadIdx.x < 32) {
    unsigned a = (unsigned)__cvta_generic_to_shared(&base);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while ((long long)(t1 - t0) < spin_ns);
  __syncthreads();
  if (COLS > 0 && threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(COLS));
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    rec[blockIdx.x * 3 + 0] = t0, rec[blockIdx.x * 3 + 1] = t1, rec[blockIdx.x * 3 + 2] = smid;
    dyn[0] = 0;
  }
}
template <int COLS>
void probe(int smem) {
  auto kern = k<COLS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int q = -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, 128, smem);
  const int n = 296;
  unsigned long long* d;
  cudaMalloc(&d, n * 3 * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  kern<<<n, 128, smem>>>(d, 1000);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  kern<<<n, 128, smem>>>(d, 200000);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<unsigned long long> h(n * 3);
  cudaMemcpy(h.data(), d, n * 3 * 8, cudaMemcpyDeviceToHost);
  // max overlap per SM
  int max_conc = 0;
  for (int i = 0; i < n; ++i) {
    int c = 0;
    for (int j = 0; j < n; ++j)
      if (h[j * 3 + 2] == h[i * 3 + 2] && h[j * 3] <= h[i * 3] && h[j * 3 + 1] > h[i * 3]) ++c;
    if (c > max_conc) max_conc = c;
  }
  printf("tmem cols=%d smem=%d: calculator blocks/SM=%d, 296 CTAs x 0.2 ms spin took %.3f ms, max concurrent CTAs on one SM=%d (%s)\n",
         COLS, smem, q, ms, max_conc, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  probe<0>(112296);
  probe<256>(112296);
  probe<256>(40000);
  probe<128>(40000);
  return 0;
}
