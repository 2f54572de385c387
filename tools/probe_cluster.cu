// Bring-up probe: how many thread-block clusters of a given size (with ~200 KB smem, 1 CTA/SM) can be co-resident,
// and what a cooperative launch of 148 such CTAs allows.  Also times a software grid barrier and a cluster barrier.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
__global__ void k_dummy(int* p) { extern __shared__ char s[]; if (p && threadIdx.x == 9999) p[0] = s[0]; }
__global__ void k_gridbar(unsigned* ctr, int iters, long long* out) {
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(ctr, 1u);
      unsigned target = (unsigned)(i + 1) * gridDim.x;
      while (*((volatile unsigned*)ctr) < target) {}
      __threadfence();
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}
__global__ void k_clusterbar(int iters, long long* out) {
  cg::cluster_group c = cg::this_cluster();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) c.sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("%s SMs=%d smemoptin=%zu\n", pr.name, pr.multiProcessorCount, pr.sharedMemPerBlockOptin);
  int smem = 200 * 1024;
  cudaFuncSetAttribute(k_dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 10, 12, 14, 15, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 16); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {(unsigned)cs, 1, 1};
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k_dummy, &cfg);
    printf("cluster %2d: max active clusters %d (%d SMs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaGetLastError();
  }
  int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gridbar, 256, 0);
  unsigned* ctr; long long* out; cudaMalloc(&ctr, 4); cudaMalloc(&out, 8);
  for (int grid : {30, 60, 120, 148}) {
    cudaMemset(ctr, 0, 4);
    int iters = 1000; void* args[] = {&ctr, &iters, &out};
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k_gridbar, dim3(grid), dim3(256), args, 0, 0);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("grid barrier, %3d CTAs: %.3f us per barrier (%s)\n", grid, ms * 1e3 / iters, cudaGetErrorString(e));
  }
  for (int cs : {2, 4, 8, 16}) {
    cudaFuncSetAttribute(k_clusterbar, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs); cfg.blockDim = dim3(256);
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {(unsigned)cs, 1, 1};
    cfg.attrs = at; cfg.numAttrs = 1;
    int iters = 1000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_clusterbar, iters, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("cluster barrier, size %2d: %.3f us per barrier (%s)\n", cs, ms * 1e3 / iters, cudaGetErrorString(e));
  }
  return 0;
}
