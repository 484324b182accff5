// Probe: how many thread-block clusters of a given size (one CTA per SM, ~200 KB of shared memory each) can be
// co-resident on this GPU (cudaOccupancyMaxActiveClusters)?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 6, 7, 8, 14, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = cs; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %2d: max active clusters %3d (%3d SMs)  %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
