// HOST EMULATION SHIM (test infrastructure, never part of libfnssl_b200.so): stands in for csrc/common.cuh so that a few .cu
// files -- kernels AND their host launch code -- compile with plain g++ and run on the CPU, one OS thread per CUDA thread, one
// thread block at a time (`__shared__` = static storage, `__syncthreads()` = a pthread barrier).  tools/host_emu/build.py rewrites
// `kernel<<<grid, block, smem, stream>>>(args)` into emu_launch(...) and gives `extern __shared__` arrays a fixed size.  Used by
// tests/test_host_emulation.py to check index arithmetic / launch geometry of the CUDA-core training kernels without a GPU.
#pragma once
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/fnssl_b200.h"

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) float4 { float x, y, z, w; };
inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
struct __half { unsigned short v; };
struct __half2 { __half x, y; };
struct float2 { float x, y; };
struct uint4 { unsigned x, y, z, w; };
inline float2 __half22float2(__half2) { float2 r; r.x = 0.0f; r.y = 0.0f; return r; }     // fp16 grids are not emulated

inline thread_local dim3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;
inline pthread_barrier_t g_emu_barrier;
inline std::mutex g_emu_atomic;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __syncthreads() pthread_barrier_wait(&g_emu_barrier)

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline const char* cudaGetErrorString(int) { return "host emulation"; }
inline int cudaGetLastError() { return 0; }
template <class F> int cudaFuncSetAttribute(F, int, int) { return 0; }
inline int cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline float __expf(float x) { return expf(x); }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline float tanhf_emu_unused(float x) { return tanhf(x); }
inline float atomicAdd(float* p, float v) { std::lock_guard<std::mutex> g(g_emu_atomic); float o = *p; *p = o + v; return o; }

// warp shuffles: the 32 threads of a warp meet on a per-warp barrier around a per-warp exchange buffer (converged code only)
inline float g_emu_shfl[32][32];
inline pthread_barrier_t g_emu_warp_barrier[32];
inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  const unsigned t = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const unsigned w = t / 32, l = t % 32;
  g_emu_shfl[w][l] = v;
  pthread_barrier_wait(&g_emu_warp_barrier[w]);
  const float r = g_emu_shfl[w][l ^ (unsigned)lane_mask];
  pthread_barrier_wait(&g_emu_warp_barrier[w]);
  return r;
}

inline char g_emu_log[4096];                      // names of the kernels launched since the last emu_launch_log_reset()
inline void emu_launch(const char* name, const std::function<void()>& body, dim3 grid, dim3 block) {
  if (strlen(g_emu_log) + strlen(name) + 40 < sizeof(g_emu_log))          // "name[gx,gy,gz];"
    snprintf(g_emu_log + strlen(g_emu_log), 40 + strlen(name), "%s[%u,%u,%u];", name, grid.x, grid.y, grid.z);
  const unsigned nthreads = block.x * block.y * block.z;
  blockDim = block;
  gridDim = grid;
  pthread_barrier_init(&g_emu_barrier, nullptr, nthreads);
  for (unsigned w = 0; w < (nthreads + 31) / 32 && w < 32; ++w)
    pthread_barrier_init(&g_emu_warp_barrier[w], nullptr, nthreads - 32 * w < 32 ? nthreads - 32 * w : 32);
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nthreads; ++t)
    pool.emplace_back([&, t] {
      threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            blockIdx = dim3(bx, by, bz);
            body();
            pthread_barrier_wait(&g_emu_barrier);   // a block's shared (static) storage is free before the next block starts
          }
    });
  for (auto& th : pool) th.join();
  pthread_barrier_destroy(&g_emu_barrier);
  for (unsigned w = 0; w < (nthreads + 31) / 32 && w < 32; ++w) pthread_barrier_destroy(&g_emu_warp_barrier[w]);
}

namespace fnssl {

inline char g_emu_error[512];
inline void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_emu_error, sizeof(g_emu_error), fmt, ap);
  va_end(ap);
}

#define FNSSL_FAIL(...)              \
  do {                               \
    ::fnssl::set_error(__VA_ARGS__); \
    return 1;                        \
  } while (0)
#define FNSSL_REQUIRE(cond, ...)          \
  do {                                    \
    if (!(cond)) FNSSL_FAIL(__VA_ARGS__); \
  } while (0)
#define FNSSL_CUDA(expr) \
  do {                   \
    (void)(expr);        \
  } while (0)
#define FNSSL_LAUNCH_CHECK(name) \
  do {                           \
  } while (0)

template <typename T> inline float ld_act(const T* p);
template <> inline float ld_act<float>(const float* p) { return *p; }
template <> inline float ld_act<__half>(const __half*) { return 0.0f; }      // fp16 grids are not emulated
template <typename T> inline void st_act(T* p, float v);
template <> inline void st_act<float>(float* p, float v) { *p = v; }
template <> inline void st_act<__half>(__half*, float) {}

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
inline float tanh_f(float x) { return 1.0f - 2.0f / (1.0f + expf(2.0f * x)); }

}  // namespace fnssl

extern "C" __attribute__((weak)) const char* emu_last_error(void) { return fnssl::g_emu_error; }
extern "C" __attribute__((weak)) const char* emu_launch_log(void) { return g_emu_log; }
extern "C" __attribute__((weak)) void emu_launch_log_reset(void) { g_emu_log[0] = 0; }
