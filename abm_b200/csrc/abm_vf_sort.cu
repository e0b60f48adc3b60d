// Spatial (Morton) re-ordering of the agents inside every replicate, and the permutation
// plumbing around it.  The step kernels give one thread to each focal agent and let all lanes of
// a warp visit the same neighbour; when the 32 focal agents of a warp are spatially close, the
// rare per-lane paths (wide intervals of near neighbours, zero-width far neighbours, distance
// culling) become warp-coherent.  The engine therefore keeps its state in an internal order
// (sorted by the Morton code of the position, refreshed every few steps -- agents move a few
// pixels per step) and a permutation `perm[slot] = API index`; every entry point of the C ABI
// speaks API order.  The sort itself is library plumbing (CUB segmented radix sort).
#include <cub/device/device_segmented_radix_sort.cuh>

#include "abm_common.cuh"

namespace abm {

__device__ __forceinline__ uint32_t spread16(uint32_t a) {
  a &= 0xffffu;
  a = (a | (a << 8)) & 0x00FF00FFu;
  a = (a | (a << 4)) & 0x0F0F0F0Fu;
  a = (a | (a << 2)) & 0x33333333u;
  a = (a | (a << 1)) & 0x55555555u;
  return a;
}

__global__ void morton_keys_kernel(const float4* rec, int N, long long n, float x0, float y0, float inv_cell,
                                   uint32_t* keys, int* vals) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const float4 v = rec[g];
  const float fx = fminf(fmaxf((v.x - x0) * inv_cell, 0.0f), 65535.0f);
  const float fy = fminf(fmaxf((v.y - y0) * inv_cell, 0.0f), 65535.0f);
  keys[g] = spread16((uint32_t)fx) | (spread16((uint32_t)fy) << 1);
  vals[g] = (int)(g % N);
}

__global__ void segment_offsets_kernel(int* off, int B, int N) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= B) off[b] = b * N;
}

// out[b, s] = in[b, order[b, s]]
template <typename T>
__global__ void gather_kernel(const T* in, const int* order, T* out, int N, long long n) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const long long base = g - (g % N);
  out[g] = in[base + order[g]];
}
// out[b, perm[b, s]] = in[b, s]
template <typename T>
__global__ void scatter_kernel(const T* in, const int* perm, T* out, int N, long long n) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const long long base = g - (g % N);
  out[base + perm[g]] = in[g];
}
__global__ void iota_kernel(int* p, int N, long long n) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) p[g] = (int)(g % N);
}

static inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

size_t vf_sort_temp_bytes(int B, int N) {
  size_t bytes = 0;
  cub::DeviceSegmentedRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                           (const int*)nullptr, (int*)nullptr, (long long)B * N, B, (const int*)nullptr,
                                           (const int*)nullptr, 0, 32, 0);
  return bytes;
}

// order[b, s] = current slot of the agent that belongs at sorted position s
cudaError_t vf_sort_order(const float4* rec, int B, int N, float x0, float y0, float extent, void* temp, size_t temp_bytes,
                          uint32_t* keys_in, uint32_t* keys_out, int* vals_in, int* order, int* offsets,
                          cudaStream_t stream) {
  const long long n = (long long)B * N;
  const float inv_cell = 65535.0f / fmaxf(extent, 1.0f);
  morton_keys_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(rec, N, n, x0, y0, inv_cell, keys_in, vals_in);
  segment_offsets_kernel<<<blocks_for(B + 1, 256), 256, 0, stream>>>(offsets, B, N);
  return cub::DeviceSegmentedRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, order, n, B, offsets,
                                                  offsets + 1, 0, 32, stream);
}

void launch_gather_f4(const float4* in, const int* order, float4* out, int N, long long n, cudaStream_t s) {
  gather_kernel<float4><<<blocks_for(n, 256), 256, 0, s>>>(in, order, out, N, n);
}
void launch_gather_f32(const float* in, const int* order, float* out, int N, long long n, cudaStream_t s) {
  gather_kernel<float><<<blocks_for(n, 256), 256, 0, s>>>(in, order, out, N, n);
}
void launch_gather_i32(const int* in, const int* order, int* out, int N, long long n, cudaStream_t s) {
  gather_kernel<int><<<blocks_for(n, 256), 256, 0, s>>>(in, order, out, N, n);
}
void launch_scatter_f32(const float* in, const int* perm, float* out, int N, long long n, cudaStream_t s) {
  scatter_kernel<float><<<blocks_for(n, 256), 256, 0, s>>>(in, perm, out, N, n);
}
void launch_iota(int* p, int N, long long n, cudaStream_t s) { iota_kernel<<<blocks_for(n, 256), 256, 0, s>>>(p, N, n); }

}  // namespace abm
