// Shared helpers for the fnssl_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fnssl_b200.h"

namespace fnssl {

// thread-local error text behind fnssl_last_error()
void set_error(const char* fmt, ...);

#define FNSSL_FAIL(...)          \
  do {                           \
    ::fnssl::set_error(__VA_ARGS__); \
    return 1;                    \
  } while (0)

#define FNSSL_REQUIRE(cond, ...)       \
  do {                                 \
    if (!(cond)) FNSSL_FAIL(__VA_ARGS__); \
  } while (0)

#define FNSSL_CUDA(expr)                                                                  \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) FNSSL_FAIL("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define FNSSL_LAUNCH_CHECK(name)                                                          \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) FNSSL_FAIL("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

// element load/store with conversion to/from fp32
template <typename T> __device__ __forceinline__ float ld_act(const T* p);
template <> __device__ __forceinline__ float ld_act<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_act<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void st_act(T* p, float v);
template <> __device__ __forceinline__ void st_act<float>(float* p, float v) { *p = v; }
// fp16 stores saturate at +-65504 instead of overflowing to inf (an inf activation becomes NaN in a recurrent state); NaN stays NaN
__device__ __forceinline__ float sat_f16(float x) { return x != x ? x : fminf(fmaxf(x, -65504.0f), 65504.0f); }
template <> __device__ __forceinline__ void st_act<__half>(__half* p, float v) { *p = __float2half_rn(sat_f16(v)); }

__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Gate non-linearities.  ex2.approx has 2 ulp error, the divisions are IEEE (no fast-math):
// the whole path is checked against the fp32 CPU oracle at 1e-3 (tensor engine) / 2e-5 (SIMT engine).
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) {
  // tanh(x) = 1 - 2/(1+e^{2x}); saturates cleanly for |x| large (e^{2x} -> inf gives 1, -> 0 gives -1)
  return 1.0f - 2.0f / (1.0f + __expf(2.0f * x));
}

}  // namespace fnssl
