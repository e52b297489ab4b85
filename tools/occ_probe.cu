// cudaOccupancyMaxActiveClusters for a 704-thread, 230 KB CTA at cluster sizes 1 / 2 / 4 (B200: 148 / 74 / 33): nvcc -gencode arch=compute_100a,code=sm_100a -o occ_probe tools/occ_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(704, 1) k(float* p) { extern __shared__ char s[]; if (p) p[0] = s[threadIdx.x]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 230000);
  for (int cl = 1; cl <= 4; cl *= 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(704); cfg.dynamicSmemBytes = 230000;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster %d: max active clusters %d (%s)\n", cl, n, cudaGetErrorString(e));
  }
  return 0;
}
