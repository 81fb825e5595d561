// Probe: which (cluster size, threads, registers, dynamic smem) combinations cudaOccupancyMaxActiveClusters accepts.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __maxnreg__(192) k192(float* out) {
  extern __shared__ float s[];
  unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  s[threadIdx.x] = r;
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (out) out[blockIdx.x * blockDim.x + threadIdx.x] = s[threadIdx.x];
}
__global__ void __launch_bounds__(320) klb(float* out) {
  extern __shared__ float s[];
  unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  s[threadIdx.x] = r;
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (out) out[blockIdx.x * blockDim.x + threadIdx.x] = s[threadIdx.x];
}
template <typename K>
void probe(const char* name, K kern, int csize, int threads, size_t smem) {
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(threads); cfg.gridDim = dim3(csize * 8); cfg.dynamicSmemBytes = smem; cfg.attrs = attr; cfg.numAttrs = 1;
  int n = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
  float* out = nullptr;
  cudaMalloc(&out, sizeof(float) * csize * 8 * threads);
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, out);
  cudaError_t se = cudaDeviceSynchronize();
  printf("%-6s csize=%d threads=%d smem=%zu: query=%s n=%d launch=%s sync=%s\n", name, csize, threads, smem, cudaGetErrorString(e), n,
         cudaGetErrorString(le), cudaGetErrorString(se));
  cudaFree(out);
  cudaGetLastError();
}
int main() {
  for (int cs : {1, 2, 4, 8})
    for (size_t sm : {(size_t)4096, (size_t)100 * 1024, (size_t)170608, (size_t)200 * 1024}) {
      probe("k192", k192, cs, 320, sm);
      probe("klb", klb, cs, 320, sm);
    }
  probe("k192", k192, 4, 256, 170608);
  probe("k192", k192, 4, 128, 170608);
  return 0;
}
