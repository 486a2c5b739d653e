// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (SS mode, K-major 128B-swizzled operands) on sm_100a.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/mma_rate.bin tools/micro/mma_rate.cu
// Prints SM cycles per K=16 MMA for M=128 with N in {64,128,256}, cta_group::1, and for cta_group::2 (M=256) N in {128,256}.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

template <int CG, int MODE>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int iters, int commit_every, long long* out) {
  const int ab_same_k = 0;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ __align__(8) unsigned long long bar2;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_addr(smem_dyn) + 1023u) & ~1023u;
  const uint32_t a_base = base;               // 4 stages x [128 x 64] fp16 = 64 KB
  const uint32_t b_base = base + 4 * 16384;   // 4 stages x [256 x 64] fp16 = 128 KB (N/CG rows used)
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_dyn + (base - smem_addr(smem_dyn)))[i] = 0x3c003c00u;   // fp16 1.0
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar2)), "r"(1 << 20));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t M = (CG == 2) ? 256 : 128;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  long long t0 = 0, t1 = 0, t2 = 0;
  if (threadIdx.x < 32 && rank == 0) {
    bool leader;
    if (MODE == 0) leader = threadIdx.x == 0;
    else {
      uint32_t pred;
      asm volatile("{\n .reg .pred q;\n elect.sync _|q, 0xffffffff;\n selp.u32 %0, 1, 0, q;\n}" : "=r"(pred));
      leader = pred != 0;
    }
    if (MODE == 0) {
      if (leader) {
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
          const int st = it & 3;
          const uint64_t ad = make_sw128_desc(a_base + st * 16384);
          const uint64_t bd = make_sw128_desc(b_base + st * 32768);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t kk = 2u * k;
            const uint32_t acc = (it | k) ? 1u : 0u;
            asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
                         "l"(ad + kk), "l"(bd + kk), "r"(idesc), "r"(acc) : "memory");
          }
          if (commit_every == 4)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bar2)) : "memory");
        }
        t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bar)) : "memory");
      }
    } else {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const int st = it & 3;
        const uint64_t ad = make_sw128_desc(a_base + st * 16384);
        const uint64_t bd = make_sw128_desc(b_base + st * 32768);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t kk = 2u * k;
          const uint32_t acc = (it | k) ? 1u : 0u;
          if (MODE == 1) {
            if (leader)
              asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
                           "l"(ad + kk), "l"(bd + kk), "r"(idesc), "r"(acc) : "memory");
          } else {
            asm volatile("{\n .reg .pred p, q;\n elect.sync _|q, 0xffffffff;\n setp.ne.b32 p, %4, 0;\n @q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
                         "l"(ad + kk), "l"(bd + kk), "r"(idesc), "r"(acc) : "memory");
          }
        }
        if (commit_every == 4) {
          if (MODE == 1) {
            if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bar2)) : "memory");
          } else {
            asm volatile("{\n .reg .pred q;\n elect.sync _|q, 0xffffffff;\n @q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}" ::"r"(smem_addr(&bar2)) : "memory");
          }
        }
        __syncwarp();
      }
      t1 = clock64();
      if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bar)) : "memory");
    }
  }
  if (threadIdx.x == 0) {
    while (!try_wait(smem_addr(&bar), 0)) {}
    t2 = clock64();
    if (rank == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int CG, int MODE>
static void run(int N, int grid, int same_k, long long* dev, long long* host) {
  const int iters = 256;
  const size_t smem = 4 * 16384 + 4 * 32768 + 1024;
  auto kern = mma_rate_kernel<CG, MODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(128, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, N, iters, same_k, dev);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cta_group::%d N=%d: %s\n", CG, N, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(host, dev, 16, cudaMemcpyDeviceToHost);
  const double n = iters * 4.0;
  printf("mode %d cta_group::%d M=%d N=%3d grid=%3d commit_every=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (%.0f MAC/clk/SM)\n", MODE, CG, CG * 128, N, grid,
         same_k, host[0] / n, host[1] / n, 128.0 * N * 16 / (host[1] / n));
}

int main() {
  long long *dev, host[2];
  cudaMalloc(&dev, 16);
  const int grid = 148;
  for (int ce : {0, 4}) {
    for (int N : {64, 128}) {
      run<1, 0>(N, grid, ce, dev, host);
      run<1, 1>(N, grid, ce, dev, host);
      run<1, 2>(N, grid, ce, dev, host);
    }
  }
  return 0;
}
