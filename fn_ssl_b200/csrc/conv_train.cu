// Training side of IPDnet's causal conv block: the three products autograd runs behind nn.Conv2d in
// CausCnnBlock (IPDnet/FixedAarryIPDnet.py:42-73; kernel 3x3, padding (1,2), crop of the last two frames, no bias) on fp32
// channels-last grids (nb, nt, nf, C):
//
//   forward        y[b,t,f,o]  = sum_{c,kf,kt} w[o,c,kf,kt] x[b, t+kt-2, f+kf-1, c]            (zero outside the grid: causal in t)
//   backward data  dx[b,t,f,c] = sum_{o,kf,kt} w[o,c,kf,kt] dy[b, t-kt+2, f-kf+1, o]
//   backward weight dw[o,c,kf,kt] = sum_{b,t,f} dy[b,t,f,o] x[b, t+kt-2, f+kf-1, c]
//
// Forward and backward-data are ONE gathered product (conv3x3_gemm_kernel: [positions x 9 Cin] x [9 Cin x Cout], the operand
// row of a position is gathered tap by tap with the shift sign as a parameter; backward-data swaps the channel roles through a
// second weight packing); backward-weight reduces over positions (split over CTAs, fp32 atomics).  128 x 128 x 16 register tiles
// as in lstm_train.cu.  ReLU, average pooling and tanh between the convs are elementwise host-side ops (fn_ssl_b200/training.py).
#include "common.cuh"

namespace fnssl {
namespace {

constexpr int kThreads = 256;
constexpr int TM = 128, TN = 128, TK = 16, TLD = TM + 4;

struct Tile {
  float v[8][8];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int w = 0; w < 8; ++w) v[u][w] = 0.0f;
  }
  __device__ __forceinline__ void mac(const float (*As)[TLD], const float (*Bs)[TLD], int ty, int tx) {
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int w = 0; w < 8; ++w) v[u][w] = fmaf(a[u], b[w], v[u][w]);
    }
  }
  __device__ __forceinline__ static int idx(int q, int u) { return (u < 4 ? 0 : 60) + q * 4 + u; }
};

// (O, C, 3, 3) -> fwd: wr[(tap*C + c)*O + o];  swapped (for backward-data): wr[(tap*O + o)*C + c]
__global__ void conv_train_repack_kernel(const float* __restrict__ w, int O, int C, int swapped, float* __restrict__ wr) {
  const int total = O * C * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 9;
    const int c = (i / 9) % C;
    const int o = i / (9 * C);
    wr[swapped ? ((size_t)tap * O + o) * C + c : ((size_t)tap * C + c) * O + o] = w[i];
  }
}

// dwr[(tap*C + c)*O + o] -> dw (O, C, 3, 3)
__global__ void conv_train_unpack_kernel(const float* __restrict__ dwr, int O, int C, float* __restrict__ dw) {
  const int total = O * C * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 9;
    const int c = (i / 9) % C;
    const int o = i / (9 * C);
    dw[i] = dwr[((size_t)tap * C + c) * O + o];
  }
}

struct ConvGemmParams {
  const float* in0; int c0; int ld0;     // input = concat(in0[c0], in1[c1]) -- C = c0 + c1 channels
  const float* in1; int c1; int ld1;
  int nb, nt, nf;
  int sign;                              // +1: row of position (t,f), tap (kf,kt) = in[t + kt - 2, f + kf - 1];  -1: in[t - kt + 2, f - kf + 1]
  const float* wr;                       // [tap][C][O]
  int O;
  float* out0; int o0; int old0;         // result columns [0, o0) -> out0, [o0, O) -> out1 (either may be NULL: not needed)
  float* out1; int old1;
};

__global__ void __launch_bounds__(kThreads)
conv3x3_gemm_kernel(const ConvGemmParams p) {
  __shared__ __align__(16) float As[TK][TLD];
  __shared__ __align__(16) float Bs[TK][TLD];
  const int t = threadIdx.x, tx = t % 16, ty = t / 16;
  const int C = p.c0 + p.c1;
  const int Kt = 9 * C;
  const int npos = p.nb * p.nt * p.nf;
  const int m0 = blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const int lq = (t % 4) * 4;
  // the two positions whose operand rows this thread gathers, decomposed once
  int pb[2], pt[2], pf[2];
  bool pv[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int pos = m0 + t / 4 + 64 * r;
    pv[r] = pos < npos;
    pf[r] = pos % p.nf;
    pt[r] = (pos / p.nf) % p.nt;
    pb[r] = pos / (p.nf * p.nt);
  }
  Tile acc;
  acc.clear();
  for (int kk0 = 0; kk0 < Kt; kk0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kidx = kk0 + lq + i;
      const int tap = kidx / C, c = kidx - tap * C;
      const int kf = tap / 3, kt = tap - kf * 3;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float v = 0.0f;
        if (kidx < Kt && pv[r]) {
          const int tt = pt[r] + p.sign * (kt - 2), ff = pf[r] + p.sign * (kf - 1);
          if (tt >= 0 && tt < p.nt && ff >= 0 && ff < p.nf) {
            const int64_t q = ((int64_t)pb[r] * p.nt + tt) * p.nf + ff;
            v = (c < p.c0) ? p.in0[q * p.ld0 + c] : p.in1[q * p.ld1 + (c - p.c0)];
          }
        }
        As[lq + i][t / 4 + 64 * r] = v;
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {           // B slab: Bs[kk][n] = wr[(kk0 + kk)*O + n0 + n], coalesced along n
      const int e = t + kThreads * r;
      const int kk = e / TN, n = e % TN;
      const int kidx = kk0 + kk;
      Bs[kk][n] = (kidx < Kt && n0 + n < p.O) ? __ldg(p.wr + (size_t)kidx * p.O + n0 + n) : 0.0f;
    }
    __syncthreads();
    acc.mac(As, Bs, ty, tx);
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int pos = m0 + Tile::idx(ty, u);
    if (pos >= npos) continue;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int n = n0 + Tile::idx(tx, w);
      if (n >= p.O) continue;
      if (n < p.o0) { if (p.out0) p.out0[(int64_t)pos * p.old0 + n] = acc.v[u][w]; }
      else if (p.out1) p.out1[(int64_t)pos * p.old1 + (n - p.o0)] = acc.v[u][w];
    }
  }
}

struct ConvWgradParams {
  const float* in0; int c0; int ld0;
  const float* in1; int c1; int ld1;
  const float* dy; int O; int dld;
  int nb, nt, nf;
  float* dwr;                            // [tap][C][O], zero on entry
  int chunk;                             // positions per CTA along the reduction (multiple of TK)
};

// C[(tap, c)][o] = sum_pos in[pos + shift(tap)][c] dy[pos][o];  grid = (ceil(9C / 128), ceil(O / 128), splits)
__global__ void __launch_bounds__(kThreads)
conv3x3_wgrad_kernel(const ConvWgradParams p) {
  __shared__ __align__(16) float As[TK][TLD];
  __shared__ __align__(16) float Bs[TK][TLD];
  const int t = threadIdx.x, tx = t % 16, ty = t / 16;
  const int C = p.c0 + p.c1;
  const int Kt = 9 * C;
  const int npos = p.nb * p.nt * p.nf;
  const int k0 = blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const int mb = blockIdx.z * p.chunk;
  const int me = (mb + p.chunk < npos) ? mb + p.chunk : npos;
  const int pk = t / 16, kq = t % 16;
  // this thread's 8 result rows (tap, c) of the A slab, decomposed once
  int rc[8], rkf[8], rkt[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int kidx = k0 + kq + 16 * r;
    const int tap = kidx / C;
    rc[r] = kidx < Kt ? kidx - tap * C : -1;
    rkf[r] = tap / 3;
    rkt[r] = tap - rkf[r] * 3;
  }
  Tile acc;
  acc.clear();
  for (int m0 = mb; m0 < me; m0 += TK) {
    const int pos = m0 + pk;
    const bool valid = pos < me;
    const int f = pos % p.nf, tt0 = (pos / p.nf) % p.nt, b = pos / (p.nf * p.nt);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float v = 0.0f;
      if (valid && rc[r] >= 0) {
        const int tt = tt0 + rkt[r] - 2, ff = f + rkf[r] - 1;
        if (tt >= 0 && tt < p.nt && ff >= 0 && ff < p.nf) {
          const int64_t q = ((int64_t)b * p.nt + tt) * p.nf + ff;
          v = (rc[r] < p.c0) ? p.in0[q * p.ld0 + rc[r]] : p.in1[q * p.ld1 + (rc[r] - p.c0)];
        }
      }
      As[pk][kq + 16 * r] = v;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {           // B slab: Bs[pk][n] = dy[pos][n0 + n]
      const int n = kq + 16 * r;
      Bs[pk][n] = (valid && n0 + n < p.O) ? p.dy[(int64_t)pos * p.dld + n0 + n] : 0.0f;
    }
    __syncthreads();
    acc.mac(As, Bs, ty, tx);
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int kidx = k0 + Tile::idx(ty, u);
    if (kidx >= Kt) continue;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int n = n0 + Tile::idx(tx, w);
      if (n < p.O) atomicAdd(p.dwr + (size_t)kidx * p.O + n, acc.v[u][w]);
    }
  }
}

int check_geom(const char* who, int nb, int nt, int nf) {
  FNSSL_REQUIRE(nb > 0 && nt > 0 && nf > 0 && (int64_t)nb * nt * nf < (int64_t)1 << 31, "%s: bad grid %d x %d x %d", who, nb, nt, nf);
  return 0;
}

}  // namespace
}  // namespace fnssl

using namespace fnssl;

extern "C" {

size_t fnssl_conv3x3_train_workspace_bytes(int cin, int cout) { return (size_t)9 * cin * cout * sizeof(float); }

int fnssl_conv3x3_forward(const float* in0, int c0, int ld0, const float* in1, int c1, int ld1, int nb, int nt, int nf, const float* w,
                          int cout, float* work, float* out, int out_ld, void* stream) {
  FNSSL_REQUIRE(in0 && w && work && out, "conv3x3_forward: null pointer");
  FNSSL_REQUIRE(c0 > 0 && ld0 >= c0 && c1 >= 0 && (c1 == 0 || (in1 && ld1 >= c1)) && cout > 0 && out_ld >= cout, "conv3x3_forward: bad channels");
  if (int rc = check_geom("conv3x3_forward", nb, nt, nf)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int C = c0 + c1;
  conv_train_repack_kernel<<<64, 256, 0, st>>>(w, cout, C, 0, work);
  ConvGemmParams p;
  p.in0 = in0; p.c0 = c0; p.ld0 = ld0; p.in1 = in1; p.c1 = c1; p.ld1 = ld1;
  p.nb = nb; p.nt = nt; p.nf = nf; p.sign = 1; p.wr = work; p.O = cout;
  p.out0 = out; p.o0 = cout; p.old0 = out_ld; p.out1 = nullptr; p.old1 = 0;
  const int64_t npos = (int64_t)nb * nt * nf;
  dim3 grid((unsigned)ceil_div64(npos, TM), (unsigned)((cout + TN - 1) / TN));
  conv3x3_gemm_kernel<<<grid, kThreads, 0, st>>>(p);
  FNSSL_LAUNCH_CHECK("conv3x3_gemm_kernel");
  return 0;
}

int fnssl_conv3x3_backward_data(const float* dy, int cout, int dy_ld, int nb, int nt, int nf, const float* w, int c0, int c1, float* work,
                                float* din0, int din0_ld, float* din1, int din1_ld, void* stream) {
  FNSSL_REQUIRE(dy && w && work && (din0 || din1), "conv3x3_backward_data: null pointer");
  FNSSL_REQUIRE(cout > 0 && dy_ld >= cout && c0 > 0 && c1 >= 0 && (!din0 || din0_ld >= c0) && (!din1 || (c1 > 0 && din1_ld >= c1)),
                "conv3x3_backward_data: bad channels");
  if (int rc = check_geom("conv3x3_backward_data", nb, nt, nf)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int C = c0 + c1;
  conv_train_repack_kernel<<<64, 256, 0, st>>>(w, cout, C, 1, work);       // [tap][o][c]: the "input" channels are dy's
  ConvGemmParams p;
  p.in0 = dy; p.c0 = cout; p.ld0 = dy_ld; p.in1 = nullptr; p.c1 = 0; p.ld1 = 0;
  p.nb = nb; p.nt = nt; p.nf = nf; p.sign = -1; p.wr = work; p.O = C;
  p.out0 = din0; p.o0 = c0; p.old0 = din0_ld; p.out1 = din1; p.old1 = din1_ld;
  const int64_t npos = (int64_t)nb * nt * nf;
  dim3 grid((unsigned)ceil_div64(npos, TM), (unsigned)((C + TN - 1) / TN));
  conv3x3_gemm_kernel<<<grid, kThreads, 0, st>>>(p);
  FNSSL_LAUNCH_CHECK("conv3x3_gemm_kernel");
  return 0;
}

int fnssl_conv3x3_backward_weight(const float* in0, int c0, int ld0, const float* in1, int c1, int ld1, const float* dy, int cout, int dy_ld,
                                  int nb, int nt, int nf, float* work, float* dw, void* stream) {
  FNSSL_REQUIRE(in0 && dy && work && dw, "conv3x3_backward_weight: null pointer");
  FNSSL_REQUIRE(c0 > 0 && ld0 >= c0 && c1 >= 0 && (c1 == 0 || (in1 && ld1 >= c1)) && cout > 0 && dy_ld >= cout, "conv3x3_backward_weight: bad channels");
  if (int rc = check_geom("conv3x3_backward_weight", nb, nt, nf)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int C = c0 + c1;
  FNSSL_CUDA(cudaMemsetAsync(work, 0, (size_t)9 * C * cout * sizeof(float), st));
  ConvWgradParams p;
  p.in0 = in0; p.c0 = c0; p.ld0 = ld0; p.in1 = in1; p.c1 = c1; p.ld1 = ld1;
  p.dy = dy; p.O = cout; p.dld = dy_ld; p.nb = nb; p.nt = nt; p.nf = nf; p.dwr = work;
  const int64_t npos = (int64_t)nb * nt * nf;
  const int64_t tiles = (int64_t)((9 * C + TM - 1) / TM) * ((cout + TN - 1) / TN);
  int64_t want = (148 * 4 + tiles - 1) / tiles;
  if (want < 1) want = 1;
  int64_t chunk = ceil_div64(npos, want);
  if (chunk < 1024) chunk = 1024;
  chunk = ceil_div64(chunk, TK) * TK;
  p.chunk = (int)chunk;
  dim3 grid((unsigned)((9 * C + TM - 1) / TM), (unsigned)((cout + TN - 1) / TN), (unsigned)ceil_div64(npos, chunk));
  conv3x3_wgrad_kernel<<<grid, kThreads, 0, st>>>(p);
  FNSSL_LAUNCH_CHECK("conv3x3_wgrad_kernel");
  conv_train_unpack_kernel<<<64, 256, 0, st>>>(work, cout, C, dw);
  FNSSL_LAUNCH_CHECK("conv_train_unpack_kernel");
  return 0;
}

}  // extern "C"
