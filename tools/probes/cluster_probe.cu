// How many 2-CTA clusters of a 1-CTA-per-SM kernel (230 KB smem) can be resident on this GPU?
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void __launch_bounds__(256, 1) k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  int smem = 225 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 / cs * cs); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension;
    a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    cfg.attrs = a; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %d: max active clusters %d (%s) -> %d CTAs\n", cs, n, cudaGetErrorString(e), n * cs);
  }
  return 0;
}
