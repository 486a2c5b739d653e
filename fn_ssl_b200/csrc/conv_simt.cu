// CausCnnBlock of IPDnet (IPDnet/FixedAarryIPDnet.py:42-73) on channels-last grids, fp32 CUDA cores.
//
//   conv1 3x3 (pad (1,2), no bias) -> ReLU -> crop last 2 frames -> AvgPool(1,3)
//   conv2 3x3                     -> ReLU -> crop              -> AvgPool(1,4)
//   conv3 3x3                     -> crop -> tanh
// "pad 2 then crop 2" along time makes every conv causal: out[f,t] = sum_{kf,kt} W[o,c,kf,kt] * in[f+kf-1, t+kt-2].
// ReLU and the average pooling are fused into the conv epilogue (pooling window = POOL consecutive frames
// computed by the same thread), so each intermediate is written once, already pooled.
#include <stdlib.h>

#include "common.cuh"

namespace fnssl {

// (O, C, 3, 3) -> [kf][kt][c][o]
__global__ void repack_conv_weight_kernel(const float* __restrict__ w, int O, int C, float* __restrict__ wr) {
  const int total = O * C * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int o = i % O;
    const int c = (i / O) % C;
    const int tap = i / (O * C);
    wr[i] = w[((size_t)o * C + c) * 9 + tap];
  }
}

constexpr int kConvTF = 8;

// thread = output channel o (blockDim.x == O, O multiple of 32); CTA = (b, t_out, 8 bins).
// input = concat(src0[c0], src1[c1]); output grid (nb, nt_out, nf, O) f32, nt_out = nt_in / POOL.
template <typename T, int POOL>
__global__ void __launch_bounds__(128)
conv3x3_pool_relu_kernel(const T* __restrict__ src0, int c0, int ld0, const T* __restrict__ src1, int c1, int ld1,
                         int nt_in, int nf, const float* __restrict__ wr, int O, float* __restrict__ out) {
  extern __shared__ __align__(16) float patch[];  // [(POOL+2)][(kConvTF+2)][Cp]
  const int C = c0 + c1;
  const int Cp = (C + 3) & ~3;
  const int nt_out = nt_in / POOL;
  const int f0 = blockIdx.x * kConvTF;
  const int to = blockIdx.y;
  const int b = blockIdx.z;
  const int o = threadIdx.x;
  // stage the input patch: frames [to*POOL-2, to*POOL+POOL-1], bins [f0-1, f0+kConvTF]
  const int NTP = POOL + 2, NFP = kConvTF + 2;
  for (int idx = threadIdx.x; idx < NTP * NFP * Cp; idx += blockDim.x) {
    const int c = idx % Cp;
    const int fl = (idx / Cp) % NFP;
    const int tl = idx / (Cp * NFP);
    const int t = to * POOL - 2 + tl, f = f0 - 1 + fl;
    float v = 0.0f;
    if (c < C && t >= 0 && t < nt_in && f >= 0 && f < nf) {
      const int64_t pos = ((int64_t)b * nt_in + t) * nf + f;
      v = (c < c0) ? ld_act<T>(src0 + pos * ld0 + c) : ld_act<T>(src1 + pos * ld1 + (c - c0));
    }
    patch[idx] = v;
  }
  __syncthreads();
  float acc[POOL][kConvTF];
#pragma unroll
  for (int i = 0; i < POOL; ++i)
#pragma unroll
    for (int j = 0; j < kConvTF; ++j) acc[i][j] = 0.0f;
  for (int kf = 0; kf < 3; ++kf) {
    for (int kt = 0; kt < 3; ++kt) {
      const float* wt = wr + (size_t)(kf * 3 + kt) * C * O + o;
      for (int c = 0; c < C; c += 4) {
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = (c + q < C) ? __ldg(wt + (size_t)(c + q) * O) : 0.0f;
#pragma unroll
        for (int i = 0; i < POOL; ++i) {
#pragma unroll
          for (int j = 0; j < kConvTF; ++j) {
            const float4 a = *reinterpret_cast<const float4*>(patch + ((size_t)(i + kt) * NFP + (j + kf)) * Cp + c);
            acc[i][j] = fmaf(a.x, w[0], acc[i][j]);
            acc[i][j] = fmaf(a.y, w[1], acc[i][j]);
            acc[i][j] = fmaf(a.z, w[2], acc[i][j]);
            acc[i][j] = fmaf(a.w, w[3], acc[i][j]);
          }
        }
      }
    }
  }
  if (to < nt_out) {
#pragma unroll
    for (int j = 0; j < kConvTF; ++j) {
      const int f = f0 + j;
      if (f >= nf) break;
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < POOL; ++i) s += fmaxf(acc[i][j], 0.0f);
      out[(((int64_t)b * nt_out + to) * nf + f) * O + o] = s * (1.0f / POOL);
    }
  }
}

// last conv: tiny cout, one thread per output element, output in the reference layout (nb, cout, nf, nt)
template <typename T>
__global__ void conv3x3_tanh_kernel(const T* __restrict__ in, int C, int nb, int nt, int nf,
                                    const float* __restrict__ wr, int O, float* __restrict__ out) {
  const int64_t total = (int64_t)nb * nt * nf * O;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % O);
    const int f = (int)((i / O) % nf);
    const int t = (int)((i / ((int64_t)O * nf)) % nt);
    const int b = (int)(i / ((int64_t)O * nf * nt));
    float acc = 0.0f;
    for (int kf = 0; kf < 3; ++kf) {
      const int ff = f + kf - 1;
      if (ff < 0 || ff >= nf) continue;
      for (int kt = 0; kt < 3; ++kt) {
        const int tt = t + kt - 2;
        if (tt < 0) continue;
        const T* a = in + (((int64_t)b * nt + tt) * nf + ff) * C;
        const float* w = wr + (size_t)(kf * 3 + kt) * C * O + o;
        for (int c = 0; c < C; ++c) acc = fmaf(ld_act<T>(a + c), __ldg(w + (size_t)c * O), acc);
      }
    }
    out[(((int64_t)b * O + o) * nf + f) * nt + t] = tanhf(acc);
  }
}

// tcgen05 implicit-GEMM path (conv_tc.cu)
bool causcnn_tc_supports(int c0, int c1, int hid, int dtype);
size_t causcnn_tc_workspace_bytes(int nb, int nt, int nf, int c0, int c1, int cout);
int causcnn_forward_tc(const void* src0, int c0, int c0_real, int ld0, const void* src1, int c1, int c1_real, int ld1, int nb, int nt,
                       int nf, const float* w1, const float* w2, void* work, __half** y2_out, float** w3r_out, cudaStream_t st);

}  // namespace fnssl

using namespace fnssl;

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static int pad16(int c) { return (c + 15) / 16 * 16; }

extern "C" {

size_t fnssl_causcnn_workspace_bytes(int nb, int nt, int nf, int cin, int hid, int cout) {
  const size_t nt1 = nt / 3, nt2 = nt1 / 4;
  const size_t simt = align256((size_t)nb * nt1 * nf * hid * 4) + align256((size_t)nb * nt2 * nf * hid * 4) +
                      align256((size_t)9 * cin * hid * 4) + align256((size_t)9 * hid * hid * 4) + align256((size_t)9 * hid * cout * 4);
  const size_t tc = causcnn_tc_workspace_bytes(nb, nt, nf, pad16(cin), 64, cout);   // upper bound on the slab count
  return simt > tc ? simt : tc;
}

int fnssl_causcnn_forward(const void* src0, int c0, int ld0, const void* src1, int c1, int ld1, int dtype, int nb, int nt,
                          int nf, const float* w1, const float* w2, const float* w3, int hid, int cout, void* work,
                          float* out, void* stream) {
  FNSSL_REQUIRE(src0 && w1 && w2 && w3 && work && out, "causcnn: null pointer");
  FNSSL_REQUIRE(hid == 128 || hid == 64 || hid == 32 || hid == 96, "causcnn: cnn_hidden_dim %d not supported (32/64/96/128)", hid);
  FNSSL_REQUIRE(c1 == 0 || src1, "causcnn: src1 missing");
  FNSSL_REQUIRE(dtype == FNSSL_F32 || dtype == FNSSL_F16, "causcnn: bad dtype");
  const int cin = c0 + c1;
  const int nt1 = nt / 3, nt2 = nt1 / 4;
  if (nt2 == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  {
    // tensor-core path: fp16 grids, 128 hidden channels, sources zero-padded to multiples of 16 channels.
    // FNSSL_CONV_ENGINE=simt forces the CUDA-core path below (tests).
    const int c0p = pad16(c0), c1p = c1 > 0 ? pad16(c1) : 0;
    const char* e = getenv("FNSSL_CONV_ENGINE");
    const bool want_tc = !(e && e[0] == 's');
    if (want_tc && ld0 >= c0p && (c1 == 0 || ld1 >= c1p) && causcnn_tc_supports(c0p, c1p, hid, dtype)) {
      __half* y2h = nullptr;
      float* w3r = nullptr;
      if (causcnn_forward_tc(src0, c0p, c0, ld0, src1, c1p, c1, ld1, nb, nt, nf, w1, w2, work, &y2h, &w3r, st)) return 1;
      repack_conv_weight_kernel<<<16, 256, 0, st>>>(w3, cout, hid, w3r);
      const int64_t total = (int64_t)nb * nt2 * nf * cout;
      const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
      conv3x3_tanh_kernel<__half><<<blocks, 256, 0, st>>>(y2h, hid, nb, nt2, nf, w3r, cout, out);
      FNSSL_LAUNCH_CHECK("conv3x3_tanh_kernel");
      return 0;
    }
  }
  char* wsp = (char*)work;
  float* y1 = (float*)wsp; wsp += align256((size_t)nb * nt1 * nf * hid * 4);
  float* y2 = (float*)wsp; wsp += align256((size_t)nb * nt2 * nf * hid * 4);
  float* w1r = (float*)wsp; wsp += align256((size_t)9 * cin * hid * 4);
  float* w2r = (float*)wsp; wsp += align256((size_t)9 * hid * hid * 4);
  float* w3r = (float*)wsp;
  repack_conv_weight_kernel<<<64, 256, 0, st>>>(w1, hid, cin, w1r);
  repack_conv_weight_kernel<<<64, 256, 0, st>>>(w2, hid, hid, w2r);
  repack_conv_weight_kernel<<<16, 256, 0, st>>>(w3, cout, hid, w3r);
  FNSSL_LAUNCH_CHECK("repack_conv_weight_kernel");
  {
    const int Cp = (cin + 3) & ~3;
    const size_t smem = (size_t)5 * (kConvTF + 2) * Cp * 4;
    FNSSL_REQUIRE(smem <= 220 * 1024, "causcnn: too many input channels (%d)", cin);
    dim3 grid((nf + kConvTF - 1) / kConvTF, nt1, nb);
    if (dtype == FNSSL_F32) {
      FNSSL_CUDA(cudaFuncSetAttribute(conv3x3_pool_relu_kernel<float, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      conv3x3_pool_relu_kernel<float, 3><<<grid, hid, smem, st>>>((const float*)src0, c0, ld0, (const float*)src1, c1, ld1, nt,
                                                                   nf, w1r, hid, y1);
    } else {
      FNSSL_CUDA(cudaFuncSetAttribute(conv3x3_pool_relu_kernel<__half, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      conv3x3_pool_relu_kernel<__half, 3><<<grid, hid, smem, st>>>((const __half*)src0, c0, ld0, (const __half*)src1, c1, ld1,
                                                                    nt, nf, w1r, hid, y1);
    }
    FNSSL_LAUNCH_CHECK("conv3x3_pool_relu_kernel<3>");
  }
  {
    const size_t smem = (size_t)6 * (kConvTF + 2) * hid * 4;
    dim3 grid((nf + kConvTF - 1) / kConvTF, nt2, nb);
    FNSSL_CUDA(cudaFuncSetAttribute(conv3x3_pool_relu_kernel<float, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3x3_pool_relu_kernel<float, 4><<<grid, hid, smem, st>>>(y1, hid, hid, nullptr, 0, 0, nt1, nf, w2r, hid, y2);
    FNSSL_LAUNCH_CHECK("conv3x3_pool_relu_kernel<4>");
  }
  {
    const int64_t total = (int64_t)nb * nt2 * nf * cout;
    const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    conv3x3_tanh_kernel<float><<<blocks, 256, 0, st>>>(y2, hid, nb, nt2, nf, w3r, cout, out);
    FNSSL_LAUNCH_CHECK("conv3x3_tanh_kernel");
  }
  return 0;
}

}  // extern "C"
