// error.hpp -- the reference's error convention (src/error.hpp:22-99): library failures print
// to stderr and exit(1); argument errors throw std::runtime_error at the call site.
#pragma once

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "../../include/dj_b200.h"

#define CUDA_RT_CALL(call)                                                                  \
  do {                                                                                      \
    cudaError_t dj_status_ = (call);                                                        \
    if (dj_status_ != cudaSuccess) {                                                        \
      std::fprintf(stderr, "ERROR: CUDA RT call \"%s\" in line %d of file %s failed with %s (%d).\n", \
                   #call, __LINE__, __FILE__, cudaGetErrorString(dj_status_), (int)dj_status_); \
      std::exit(1);                                                                         \
    }                                                                                       \
  } while (0)

// every libdj_b200 entry point returns 0 or an error code + message
#define DJ_CALL(call)                                                                       \
  do {                                                                                      \
    int dj_rc_ = (call);                                                                    \
    if (dj_rc_ != DJ_OK) {                                                                  \
      std::fprintf(stderr, "ERROR: \"%s\" in line %d of file %s failed with %d: %s\n", #call, \
                   __LINE__, __FILE__, dj_rc_, dj_last_error());                            \
      std::exit(1);                                                                         \
    }                                                                                       \
  } while (0)

// NCCL is only reached through libdj_b200, so NCCL_CALL wraps the same convention
#define NCCL_CALL(call) DJ_CALL(call)

#define CHECK_ERROR(rtv, expected_value, msg)                                               \
  do {                                                                                      \
    if ((rtv) != (expected_value)) {                                                        \
      std::fprintf(stderr, "ERROR on line %d of file %s: %s returned %d\n", __LINE__, __FILE__, \
                   msg, (int)(rtv));                                                        \
      std::exit(1);                                                                         \
    }                                                                                       \
  } while (0)
