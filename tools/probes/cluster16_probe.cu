// Dev probe: how many 16-CTA (non-portable) clusters of large CTAs (1 per SM: 200 KB shared memory, 512 threads, 512
// TMEM columns) does a B200 keep resident, and how many 8-CTA ones?  Decides whether a two-matrices-per-cluster layout
// of the sweeps kernel (16 x 384 rows) can use more SMs than the 15 x 8 of the current one.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <set>
extern __shared__ float dyn[];
__global__ void __launch_bounds__(512, 1) k(unsigned long long* rec, long long spin_ns) {
  __shared__ unsigned base;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (threadIdx.x < 32) {
    unsigned a = (unsigned)__cvta_generic_to_shared(&base);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while ((long long)(t1 - t0) < spin_ns);
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512));
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    rec[blockIdx.x * 3 + 0] = t0, rec[blockIdx.x * 3 + 1] = t1, rec[blockIdx.x * 3 + 2] = smid;
    dyn[0] = 0;
  }
}
void probe(int csize, int smem) {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (csize > 8) cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(512), cfg.dynamicSmemBytes = smem, cfg.attrs = attr, cfg.numAttrs = 1, cfg.gridDim = dim3(csize);
  int maxc = -1;
  cudaError_t eq = cudaOccupancyMaxActiveClusters(&maxc, k, &cfg);
  const int n = 160 / csize * csize;  // more CTAs than SMs: the second wave shows up in the timing
  unsigned long long* d;
  cudaMalloc(&d, n * 3 * 8);
  cfg.gridDim = dim3(n);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaLaunchKernelEx(&cfg, k, d, 1000LL);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  cudaError_t el = cudaLaunchKernelEx(&cfg, k, d, 300000LL);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<unsigned long long> h(n * 3);
  cudaMemcpy(h.data(), d, n * 3 * 8, cudaMemcpyDeviceToHost);
  unsigned long long tmin = ~0ull;
  for (int i = 0; i < n; ++i) tmin = h[i * 3] < tmin ? h[i * 3] : tmin;
  std::set<unsigned long long> first_wave_sms;
  int first_wave = 0;
  for (int i = 0; i < n; ++i)
    if (h[i * 3] - tmin < 100000) ++first_wave, first_wave_sms.insert(h[i * 3 + 2]);
  printf("cluster %2d smem %d: occupancy query max active clusters = %d (%s); %d CTAs x 0.3 ms took %.3f ms; first wave: %d CTAs on %zu SMs (launch %s, run %s)\n",
         csize, smem, maxc, cudaGetErrorString(eq), n, ms, first_wave, first_wave_sms.size(), cudaGetErrorString(el), cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  probe(8, 210000);
  probe(16, 210000);
  probe(4, 210000);
  probe(2, 210000);
  return 0;
}
