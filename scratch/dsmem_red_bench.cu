// microbenchmark: RED.OR into the PEER CTA's shared memory (thread-block cluster of 2, distributed shared memory) mixed
// with local RED.OR -- would a replicate split over a CTA pair (half the rows each, two CTAs resident per SM) afford the
// quarter of its draws that land in the other CTA's rows?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_red_bench dsmem_red_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int kWords = 27 * 1024;   // 108 KB per CTA: two CTAs per SM
template <int REMOTE>               // of 8 REDs per iteration, REMOTE go to the peer
__global__ void __launch_bounds__(512, 2) k(uint32_t* out, int iters) {
  extern __shared__ uint32_t sm[];
  for (int i = threadIdx.x; i < kWords; i += blockDim.x) sm[i] = 0;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
  const uint32_t base = smem_u32(sm) + 4 * (threadIdx.x & 511);   // bank == lane; rows of 512 words, stride 2048 B
  uint32_t peer_base;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_base) : "r"(base), "r"(rank ^ 1u));
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
  uint32_t a8[8];
  for (int q = 0; q < 8; ++q) { x = x * 1664525u + 1013904223u; a8[q] = ((x >> 8) % 52) * 2048u; }
  for (int it = 0; it < iters; ++it) {
    const uint32_t m = 3u << (it & 15);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q < REMOTE) asm volatile("red.shared::cluster.or.b32 [%0], %1;" ::"r"(peer_base + a8[q]), "r"(m) : "memory");
      else asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(base + a8[q]), "r"(m) : "memory");
    }
  }
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
  out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
}
template <int REMOTE>
void run(uint32_t* out) {
  const int iters = 20000, threads = 512, blocks = 296;
  const size_t smem = kWords * 4;
  cudaFuncSetAttribute(k<REMOTE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n_clusters = 0;
  cudaOccupancyMaxActiveClusters(&n_clusters, k<REMOTE>, &cfg);
  cudaLaunchKernelEx(&cfg, k<REMOTE>, out, 100);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, k<REMOTE>, out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // per SM: 2 CTAs x 16 warps x iters x 8 warp-level REDs
  const double warp_ops = 2.0 * (threads / 32) * iters * 8;
  printf("%d of 8 REDs remote: %.3f ms, %.2f cycles per warp-RED per SM (at 1.965 GHz); co-resident clusters %d (148 = two CTAs per SM); %s\n",
         REMOTE, ms, ms * 1e-3 * 1.965e9 / warp_ops, n_clusters, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  uint32_t* out; cudaMalloc(&out, 296 * 512 * 4);
  run<0>(out); run<1>(out); run<2>(out); run<4>(out); run<8>(out);
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
