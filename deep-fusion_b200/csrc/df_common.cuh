// df_common.cuh -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/dfcuda.h"

namespace df {

// thread-local text behind df_last_error()
char* last_error_buf();
int fail(int code, const char* fmt, ...);

#define DF_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t df_e_ = (expr);                                                        \
    if (df_e_ != cudaSuccess)                                                          \
      return ::df::fail((int)df_e_, "%s failed: %s (%s:%d)", #expr,                    \
                        cudaGetErrorString(df_e_), __FILE__, __LINE__);                \
  } while (0)

inline int dtype_size(int dt) { return (dt == DF_F32 || dt == DF_S32) ? 4 : ((dt == DF_S8 || dt == DF_U8) ? 1 : 0); }

}  // namespace df
