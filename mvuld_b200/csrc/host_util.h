// Host-side helpers shared by the C-ABI translation units: error string, TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mv {

// Records a message retrievable through mvuld_last_error(); returns `code` for convenience.
int fail(int code, const char* fmt, ...);

#define MV_CHECK_ARG(cond, ...)                      \
  do {                                               \
    if (!(cond)) return ::mv::fail(-1, __VA_ARGS__); \
  } while (0)

#define MV_CUDA_OK(expr)                                                                              \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) return ::mv::fail((int)_e, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                             __FILE__, __LINE__);                                     \
  } while (0)

#define MV_LAUNCH_OK()                                                                                          \
  do {                                                                                                          \
    cudaError_t _e = cudaGetLastError();                                                                        \
    if (_e != cudaSuccess)                                                                                      \
      return ::mv::fail((int)_e, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// 16-bit element tensor map (bf16 or fp16 -- same size, no arithmetic in TMA).  dims/box innermost first;
// strides_bytes[i] is the byte stride of dim i+1 (rank-1 entries).  swizzle_bytes in {0, 32, 64, 128}.
int make_tmap_16b(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int swizzle_bytes);

// SM count of the CURRENT device (cached per device ordinal).
int num_sms();

// Per-device once-flag for function attributes (cudaFuncSetAttribute is per device): `mask` is a function-local static;
// returns true the first time it is asked on the current device.  Thread safe.
bool first_use_on_current_device(unsigned long long* mask);

}  // namespace mv
