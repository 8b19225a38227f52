// common.hpp -- error plumbing and small helpers shared by the library's translation units.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <string>

namespace hpccg {

// Error codes returned through the C-ABI (CUDA errors are returned as their own positive values).
enum : int {
  HPCCG_OK = 0,
  HPCCG_ERR_ARG = -1,      // bad argument
  HPCCG_ERR_STATE = -2,    // call made in the wrong state (no communicator, no mirror, ...)
  HPCCG_ERR_NCCL = -3,     // NCCL failure or NCCL not loadable
  HPCCG_ERR_ALLOC = -4,    // host allocation failure
  HPCCG_ERR_COMM = -5,     // set-up collective failed
};

void set_error(const std::string &msg);
const char *last_error();
int fail(int code, const char *fmt, ...);
int fail_cuda(cudaError_t e, const char *what, const char *file, int line);

extern std::atomic<long long> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

}  // namespace hpccg

#define HPCCG_CUDA(call)                                                        \
  do {                                                                          \
    cudaError_t hpccg_e_ = (call);                                              \
    if (hpccg_e_ != cudaSuccess) return ::hpccg::fail_cuda(hpccg_e_, #call, __FILE__, __LINE__); \
  } while (0)

#define HPCCG_TRY(call)                \
  do {                                 \
    int hpccg_rc_ = (call);            \
    if (hpccg_rc_ != 0) return hpccg_rc_; \
  } while (0)

#define HPCCG_LAUNCH_CHECK()                                                                      \
  do {                                                                                            \
    cudaError_t hpccg_e_ = cudaPeekAtLastError();                                                 \
    if (hpccg_e_ != cudaSuccess) return ::hpccg::fail_cuda(hpccg_e_, "kernel launch", __FILE__, __LINE__); \
  } while (0)
