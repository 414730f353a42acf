// Shared error handling for the C-ABI library.
#pragma once
#include <cuda_runtime.h>

#include "../../include/scilmm_b200.h"

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace slmm {

void set_last_error(const std::string& msg);
extern int64_t g_launch_count;   // kernels launched by this library since the last reset

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
  if (e != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e), file, line, what);
    throw CudaError(buf);
  }
}
#define CUDA_OK(x) ::slmm::cuda_check((x), #x, __FILE__, __LINE__)

template <typename T>
T* dev_alloc(size_t count) {
  T* p = nullptr;
  if (count == 0) count = 1;
  CUDA_OK(cudaMalloc((void**)&p, count * sizeof(T)));
  return p;
}

template <typename T>
T* dev_upload(const T* host, size_t count, cudaStream_t st = 0) {
  T* p = dev_alloc<T>(count);
  if (count) CUDA_OK(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, st));
  return p;
}

inline void dev_free(void* p) {
  if (p) cudaFree(p);
}

}  // namespace slmm

#define SLMM_TRY try {
#define SLMM_CATCH                                                       \
  }                                                                      \
  catch (const ::slmm::CudaError& e) {                                   \
    ::slmm::set_last_error(e.what());                                    \
    return SLMM_ERR_CUDA;                                                \
  }                                                                      \
  catch (const std::invalid_argument& e) {                               \
    ::slmm::set_last_error(e.what());                                    \
    return SLMM_ERR_INVALID;                                             \
  }                                                                      \
  catch (const std::exception& e) {                                      \
    ::slmm::set_last_error(e.what());                                    \
    return SLMM_ERR_INTERNAL;                                            \
  }
