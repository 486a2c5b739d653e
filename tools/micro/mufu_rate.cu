// MUFU throughput probe (sm_100a): instructions / clk / SM for tanh.approx.f32, tanh.approx.f16 (scalar halves of an f16x2
// pair, which is how ptxas lowers tanh.approx.f16x2), ex2.approx.f32 and rcp.approx.f32, with 16 resident warps per SM
// (the LSTM kernel's epilogue) and 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate.bin mufu_rate.cu && ./mufu_rate.bin
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i + 1);
  unsigned h[8];
  for (int i = 0; i < 8; ++i) { __half2 t = __floats2half2_rn(v[i], -v[i]); h[i] = *reinterpret_cast<unsigned*>(&t); }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (MODE == 1) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (MODE == 3) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (MODE == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i] + __low2float(*reinterpret_cast<__half2*>(&h[i]));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}

int main() {
  float* d;
  cudaMalloc(&d, 148 * 512 * sizeof(float));
  const int iters = 2000;
  const char* names[5] = {"tanh.approx.f32", "tanh.approx.f16x2 (2 results / instr)", "ex2.approx.f32", "rcp.approx.f32", "fma.f32"};
  for (int m = 0; m < 5; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      if (m == 0) k<0><<<148, 512>>>(d, iters);
      if (m == 1) k<1><<<148, 512>>>(d, iters);
      if (m == 2) k<2><<<148, 512>>>(d, iters);
      if (m == 3) k<3><<<148, 512>>>(d, iters);
      if (m == 4) k<4><<<148, 512>>>(d, iters);
      cudaDeviceSynchronize();
    }
    float cyc;
    cudaMemcpy(&cyc, d, sizeof(float), cudaMemcpyDeviceToHost);
    const double thread_instr = 512.0 * iters * 8;
    printf("%-40s %10.0f cycles  -> %.2f thread-instructions / clk / SM\n", names[m], cyc, thread_instr / cyc);
  }
  return 0;
}
