// Output heads of the FN-SSL path.
//   fnssl_ipd_head_forward : AvgPool2d((12,1)) -> Linear(C,2) -> tanh -> [ch0 over f | ch1 over f]
//                            (FN-SSL/Lightning/Model.py:79-87)
//   fnssl_linear_forward   : the DOA classifier Linear(512,180) (Model.py:71,88-89)
// Both are HBM-bound single-pass kernels; the head reads the last narrow-band output exactly once.
#include "common.cuh"

namespace fnssl {

// one warp per (b, t2, f): 12 frames x C channels -> 2 outputs
template <typename T>
__global__ void __launch_bounds__(256)
ipd_head_kernel(const T* __restrict__ x, int ld, int nb, int nt, int nf, int C, const float* __restrict__ w,
                const float* __restrict__ bias, float* __restrict__ out) {
  const int nt2 = nt / 12;
  const int64_t total = (int64_t)nb * nt2 * nf;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= total) return;
  const int f = (int)(warp % nf);
  const int t2 = (int)((warp / nf) % nt2);
  const int b = (int)(warp / ((int64_t)nf * nt2));
  float a0 = 0.0f, a1 = 0.0f;
  for (int c = lane; c < C; c += 32) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 12; ++k) s += ld_act<T>(x + (((int64_t)b * nt + t2 * 12 + k) * nf + f) * ld + c);
    s *= (1.0f / 12.0f);
    a0 = fmaf(s, w[c], a0);
    a1 = fmaf(s, w[C + c], a1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if (lane == 0) {
    float* o = out + ((int64_t)b * nt2 + t2) * (2 * nf);
    o[f] = tanhf(a0 + bias[0]);
    o[nf + f] = tanhf(a1 + bias[1]);
  }
}

// fp16 grids with C % 8 == 0: every lane reads 16 bytes (8 channels) per frame, so a 256-channel position is ONE
// coalesced 512-byte warp access and the 12 frames of the pooling window are 12 independent loads in flight.
__global__ void __launch_bounds__(256)
ipd_head_h8_kernel(const __half* __restrict__ x, int ld, int nb, int nt, int nf, int C, const float* __restrict__ w,
                   const float* __restrict__ bias, float* __restrict__ out) {
  const int nt2 = nt / 12;
  const int64_t total = (int64_t)nb * nt2 * nf;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= total) return;
  const int f = (int)(warp % nf);
  const int t2 = (int)((warp / nf) % nt2);
  const int b = (int)(warp / ((int64_t)nf * nt2));
  float a0 = 0.0f, a1 = 0.0f;
  for (int c8 = lane; c8 * 8 < C; c8 += 32) {
    uint4 v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k)
      v[k] = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)b * nt + t2 * 12 + k) * nf + f) * ld + c8 * 8));
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.0f;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const __half2* h = reinterpret_cast<const __half2*>(&v[k]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 p = __half22float2(h[i]);
        s[2 * i] += p.x; s[2 * i + 1] += p.y;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float m = s[i] * (1.0f / 12.0f);
      a0 = fmaf(m, __ldg(w + c8 * 8 + i), a0);
      a1 = fmaf(m, __ldg(w + C + c8 * 8 + i), a1);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if (lane == 0) {
    float* o = out + ((int64_t)b * nt2 + t2) * (2 * nf);
    o[f] = tanhf(a0 + bias[0]);
    o[nf + f] = tanhf(a1 + bias[1]);
  }
}

// Backward of ipd_head_kernel on fp32 grids (training side; autograd of Model.py:79-87): one warp per (b, t2, f), grid-stride.
//   g_o = dy_o (1 - y_o^2);  dx[b, 12 t2 + k, f, c] = (g_0 w[0][c] + g_1 w[1][c]) / 12;  dw[o][c] += g_o mean_k x;  db[o] += g_o
// Per-lane register partial sums of dw over the warp's items, combined per CTA in shared memory, one global atomic per CTA and
// channel.  dx frames beyond 12 * (nt / 12) are not written (the caller zero-fills them).
constexpr int kHeadBwdMaxC = 512;
__global__ void __launch_bounds__(256)
ipd_head_bwd_kernel(const float* __restrict__ x, int ld, int nb, int nt, int nf, int C, const float* __restrict__ w,
                    const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, int dld,
                    float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float s_dw[2 * kHeadBwdMaxC + 2];
  for (int i = threadIdx.x; i < 2 * C + 2; i += blockDim.x) s_dw[i] = 0.0f;
  __syncthreads();
  const int nt2 = nt / 12;
  const int64_t total = (int64_t)nb * nt2 * nf;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float a0[kHeadBwdMaxC / 32], a1[kHeadBwdMaxC / 32];
#pragma unroll
  for (int i = 0; i < kHeadBwdMaxC / 32; ++i) { a0[i] = 0.0f; a1[i] = 0.0f; }
  float b0 = 0.0f, b1 = 0.0f;
  for (int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < total; item += nwarps) {
    const int f = (int)(item % nf);
    const int t2 = (int)((item / nf) % nt2);
    const int b = (int)(item / ((int64_t)nf * nt2));
    const float* yo = y + ((int64_t)b * nt2 + t2) * (2 * nf);
    const float* go = dy + ((int64_t)b * nt2 + t2) * (2 * nf);
    const float g0 = go[f] * (1.0f - yo[f] * yo[f]);
    const float g1 = go[nf + f] * (1.0f - yo[nf + f] * yo[nf + f]);
    b0 += g0; b1 += g1;
#pragma unroll
    for (int i = 0; i < kHeadBwdMaxC / 32; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < 12; ++k) s += x[(((int64_t)b * nt + t2 * 12 + k) * nf + f) * ld + c];
        s *= (1.0f / 12.0f);
        a0[i] = fmaf(g0, s, a0[i]);
        a1[i] = fmaf(g1, s, a1[i]);
        const float d = (g0 * __ldg(w + c) + g1 * __ldg(w + C + c)) * (1.0f / 12.0f);
#pragma unroll
        for (int k = 0; k < 12; ++k) dx[(((int64_t)b * nt + t2 * 12 + k) * nf + f) * dld + c] = d;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kHeadBwdMaxC / 32; ++i) {
    const int c = lane + 32 * i;
    if (c < C) { atomicAdd(&s_dw[c], a0[i]); atomicAdd(&s_dw[C + c], a1[i]); }
  }
  if (lane == 0) { atomicAdd(&s_dw[2 * C], b0); atomicAdd(&s_dw[2 * C + 1], b1); }   // every lane holds the same b0 / b1
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(dw + i, s_dw[i]);
  if (threadIdx.x < 2) atomicAdd(db + threadIdx.x, s_dw[2 * C + threadIdx.x]);
}

// y[r][o] = b[o] + sum_k x[r][k] w[o][k]; one CTA per row, x row staged in shared memory
__global__ void __launch_bounds__(256)
linear_rows_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, int in_f,
                   int out_f, float* __restrict__ y) {
  extern __shared__ float xs[];
  const int r = blockIdx.x;
  for (int k = threadIdx.x; k < in_f; k += blockDim.x) xs[k] = x[(size_t)r * in_f + k];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int o = warp; o < out_f; o += nwarp) {
    float a = 0.0f;
    for (int k = lane; k < in_f; k += 32) a = fmaf(xs[k], w[(size_t)o * in_f + k], a);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
    if (lane == 0) y[(size_t)r * out_f + o] = a + b[o];
  }
}

// Backward of linear_rows_kernel (training side: the DOA classifier Linear(512,180), Model.py:71,88-89).
//   dx[r][k] = sum_o dy[r][o] w[o][k]         one CTA per row, threads over k (w rows read coalesced)
//   dw[o][k] = sum_r dy[r][o] x[r][k],  db[o] = sum_r dy[r][o]      one CTA per output o, threads over k: no atomics
__global__ void __launch_bounds__(256)
linear_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ w, int in_f, int out_f, float* __restrict__ dx) {
  extern __shared__ float gs[];                       // dy row
  const int r = blockIdx.x;
  for (int o = threadIdx.x; o < out_f; o += blockDim.x) gs[o] = dy[(size_t)r * out_f + o];
  __syncthreads();
  for (int k = threadIdx.x; k < in_f; k += blockDim.x) {
    float a = 0.0f;
    for (int o = 0; o < out_f; ++o) a = fmaf(gs[o], __ldg(w + (size_t)o * in_f + k), a);
    dx[(size_t)r * in_f + k] = a;
  }
}

__global__ void __launch_bounds__(256)
linear_bwd_dw_kernel(const float* __restrict__ x, const float* __restrict__ dy, int rows, int in_f, int out_f, float* __restrict__ dw,
                     float* __restrict__ db) {
  const int o = blockIdx.x;
  for (int k = threadIdx.x; k < in_f; k += blockDim.x) {
    float a = 0.0f;
    for (int r = 0; r < rows; ++r) a = fmaf(__ldg(dy + (size_t)r * out_f + o), __ldg(x + (size_t)r * in_f + k), a);
    dw[(size_t)o * in_f + k] = a;
  }
  if (threadIdx.x == 0) {
    float a = 0.0f;
    for (int r = 0; r < rows; ++r) a += dy[(size_t)r * out_f + o];
    db[o] = a;
  }
}

}  // namespace fnssl

using namespace fnssl;

extern "C" {

int fnssl_ipd_head_forward(const void* x, int dtype, int ld, int nb, int nt, int nf, int C, const float* w, const float* b,
                           float* out, void* stream) {
  FNSSL_REQUIRE(x && w && b && out, "ipd_head: null pointer");
  FNSSL_REQUIRE(nb > 0 && nf > 0 && C > 0 && ld >= C, "ipd_head: bad shape");
  if (nt / 12 == 0) return 0;
  const int64_t warps = (int64_t)nb * (nt / 12) * nf;
  const int64_t blocks = (warps * 32 + 255) / 256;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FNSSL_F32)
    ipd_head_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)x, ld, nb, nt, nf, C, w, b, out);
  else if (dtype == FNSSL_F16 && C % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0)
    ipd_head_h8_kernel<<<(unsigned)blocks, 256, 0, st>>>((const __half*)x, ld, nb, nt, nf, C, w, b, out);
  else if (dtype == FNSSL_F16)
    ipd_head_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>((const __half*)x, ld, nb, nt, nf, C, w, b, out);
  else
    FNSSL_FAIL("ipd_head: bad dtype %d", dtype);
  FNSSL_LAUNCH_CHECK("ipd_head_kernel");
  return 0;
}

int fnssl_ipd_head_backward(const float* x, int ld, int nb, int nt, int nf, int C, const float* w, const float* y, const float* dy,
                            float* dx, int dld, float* dw, float* db, void* stream) {
  FNSSL_REQUIRE(x && w && y && dy && dx && dw && db, "ipd_head_backward: null pointer");
  FNSSL_REQUIRE(nb > 0 && nf > 0 && C > 0 && C <= kHeadBwdMaxC && ld >= C && dld >= C, "ipd_head_backward: bad shape (C <= %d)", kHeadBwdMaxC);
  cudaStream_t st = (cudaStream_t)stream;
  FNSSL_CUDA(cudaMemsetAsync(dw, 0, (size_t)2 * C * sizeof(float), st));
  FNSSL_CUDA(cudaMemsetAsync(db, 0, 2 * sizeof(float), st));
  if (nt / 12 == 0) return 0;
  const int64_t warps = (int64_t)nb * (nt / 12) * nf;
  int64_t blocks = (warps * 32 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ipd_head_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, ld, nb, nt, nf, C, w, y, dy, dx, dld, dw, db);
  FNSSL_LAUNCH_CHECK("ipd_head_bwd_kernel");
  return 0;
}

int fnssl_linear_forward(const float* x, const float* w, const float* b, int rows, int in_features, int out_features,
                         float* y, void* stream) {
  FNSSL_REQUIRE(x && w && b && y, "linear: null pointer");
  FNSSL_REQUIRE(rows >= 0 && in_features > 0 && out_features > 0 && in_features <= 12288, "linear: bad shape");
  if (rows == 0) return 0;
  linear_rows_kernel<<<rows, 256, in_features * sizeof(float), (cudaStream_t)stream>>>(x, w, b, in_features, out_features, y);
  FNSSL_LAUNCH_CHECK("linear_rows_kernel");
  return 0;
}

int fnssl_linear_backward(const float* x, const float* w, const float* dy, int rows, int in_features, int out_features, float* dx,
                          float* dw, float* db, void* stream) {
  FNSSL_REQUIRE(x && w && dy && dw && db, "linear_backward: null pointer");
  FNSSL_REQUIRE(rows >= 0 && in_features > 0 && out_features > 0 && out_features <= 12288, "linear_backward: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (dx && rows > 0) {
    linear_bwd_dx_kernel<<<rows, 256, out_features * sizeof(float), st>>>(dy, w, in_features, out_features, dx);
    FNSSL_LAUNCH_CHECK("linear_bwd_dx_kernel");
  }
  linear_bwd_dw_kernel<<<out_features, 256, 0, st>>>(x, dy, rows, in_features, out_features, dw, db);
  FNSSL_LAUNCH_CHECK("linear_bwd_dw_kernel");
  return 0;
}

}  // extern "C"
