// Shared host-side helpers for the C-ABI entry points.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/aki_mma.h"

namespace aki {

void set_last_cuda_error(const char* msg);
// aki_mma_set_timing_events: one-shot pair of events recorded around the next tcgen05 attention kernel
void timing_hook_begin(cudaStream_t st);
void timing_hook_end(cudaStream_t st);

void count_launch();   // aki_mma_launch_count(): one tick per kernel this library enqueues

inline int check_launch() {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_cuda_error(cudaGetErrorString(e));
    return AKI_ERR_CUDA;
  }
  return AKI_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define AKI_REQUIRE(cond, code) \
  do {                          \
    if (!(cond)) return (code); \
  } while (0)

}  // namespace aki
