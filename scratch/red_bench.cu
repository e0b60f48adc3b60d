// microbenchmark: shared-memory read-modify-write throughput, LDS+LOP+STS vs RED.OR vs ATOMS.OR
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void k(uint32_t* out, int iters) {
  extern __shared__ uint32_t sm[];
  for (int i = threadIdx.x; i < 40 * 1024; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  const uint32_t base = smem_u32(sm) + 4 * threadIdx.x;     // bank == lane, conflict free; row of 40 words, stride 4096 B
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
  uint32_t acc = 0;
  uint32_t a8[8];
  for (int q = 0; q < 8; ++q) { x = x * 1664525u + 1013904223u; a8[q] = base + ((x >> 8) % 38) * 4096u; }
  for (int it = 0; it < iters; ++it) {
   const uint32_t m = 3u << (it & 15);
#pragma unroll
   for (int q = 0; q < 8; ++q) {
    const uint32_t addr = a8[q];
    if (MODE == 0) {
      uint32_t v;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v | m));
    } else if (MODE == 1) {
      asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(m) : "memory");
    } else {
      uint32_t v;
      asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(v) : "r"(addr), "r"(m) : "memory");
      acc ^= v;
    }
   }
  }
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x] ^ acc;
}
template <int MODE>
void run(const char* name, uint32_t* out) {
  const int iters = 20000, threads = 512, blocks = 148;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024 * 4);
  k<MODE><<<blocks, threads, 40 * 1024 * 4>>>(out, 100);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads, 40 * 1024 * 4>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // per SM: threads/32 warps * iters warp-level RMWs
  double warp_ops = (double)(threads / 32) * iters * 8;
  printf("%s: %.3f ms, %.2f cycles per warp-RMW per SM (at 1.9 GHz)\n", name, ms, ms * 1e-3 * 1.9e9 / warp_ops);
}
int main() {
  uint32_t* out; cudaMalloc(&out, 148 * 512 * 4);
  run<0>("LDS+LOP+STS", out); run<1>("RED.OR", out); run<2>("ATOMS.OR(ret)", out);
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
