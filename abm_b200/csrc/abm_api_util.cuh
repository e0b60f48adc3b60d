// Helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <string>

#include "../../include/abm_b200.h"

namespace abm {

int api_fail(int code, const std::string& msg);   // records the thread-local error message, returns code
const char* api_last_error();

// Owning device buffer: freed on release() or when it goes out of scope (an early return of ABM_CUDA in an entry
// point that holds several of them leaks nothing).  Not copyable.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    return cudaMalloc(reinterpret_cast<void**>(&p), sizeof(T) * (count ? count : 1));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
  }
};

}  // namespace abm

#define ABM_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) {                                                                             \
      char _b[512];                                                                                      \
      snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return abm::api_fail(_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver ? ABM_E_NO_DEVICE \
                                                                                        : ABM_E_CUDA, _b); \
    }                                                                                                    \
  } while (0)
