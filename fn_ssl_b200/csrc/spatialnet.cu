// IPDnet2 (OnlineSpatialNet with Mamba time modules) forward on sm_100a -- SURVEY.md section 8, row a11.
//
// Reference call sites replaced (Audio-WestlakeU/FN-SSL, IPDnet2/):
//   STFT.forward (center=True)              IPDnet2/Module.py:46-64          -> fnssl_reflect_pad + fnssl_stft_forward(hop 320)
//   CausalConv1d encoder                    IPDnet2/IPDnet2.py:66-76,335     \
//   SpatialNetLayer._fconv / _full          IPDnet2/IPDnet2.py:222-253        > fnssl_sn_freq_forward (one launch per layer)
//   AvgPool2d (1,2) / (1,8) over frequency  IPDnet2/IPDnet2.py:134-135,148,153/
//   SpatialNetLayer._mamba x 2 (mamba_ssm.Mamba.forward, third party) + AvgPool2d((5,1)) over time   :166-181,347
//                                                                            -> fnssl_sn_time_forward (one launch per layer)
//   FreqInverse + decoder + output reshape  IPDnet2/IPDnet2.py:37-43,357-364 -> fnssl_sn_head_forward
//
// Design.  The reference materialises (B, 256, T, 96) fp32 after the encoder (1.9 GB at cfg5) and makes ~25 passes over
// it in layer 0 alone.  Here the whole frequency-axis half of a layer (encoder, LN, grouped conv along F, PReLU, the
// squeeze / full-band linear / unsqueeze path, both frequency pools) runs inside ONE CTA per (utterance, frame) with the
// frame's [F x 96] tile resident in shared memory: HBM sees the 2M-channel feature grid once and the pooled [16 x 96]
// result once.  The time-axis half (both Mamba blocks and the 5-frame pooling) is one CTA per (utterance, band)
// sequence, one thread per inner channel, with the selective-scan state in registers and the sequence processed in
// chunks of 15 frames.  All arithmetic is fp32 on the CUDA cores (these stages are FP32-FMA bound, not GEMM-shaped enough
// to pay for fp16 staging: K = 60..192, and parity is held at 2e-5 instead of 1e-3).
//
// Activation layout between the launches: (B, T, F, 96) fp32, channels last (the reference's (B, F, T, H) permuted).
#include "common.cuh"

namespace fnssl {
namespace sn {

constexpr int kH = 96;          // dim_hidden
constexpr int kHP = 97;         // padded row stride of the [rows x 96] tiles (conflict-free row-per-thread access)
constexpr int kG = 8;           // conv groups
constexpr int kGC = 12;         // channels per group
constexpr int kFK = 5;          // kernel size along F
constexpr int kSQ = 8;          // dim_squeeze
constexpr int kRows = 256;      // rows per CTA tile = threads per CTA
constexpr int kWFloats = 7680;  // weight / scratch region (30 KB): encoder 5*16*96, conv 5760 + 4*96, full stage S + L + 2*768
constexpr int kEncK = 5;

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

// ----------------------------------------------------------------------------------------------------------------
// frequency stage
// ----------------------------------------------------------------------------------------------------------------

// LayerNorm over the 96 channels of row r (biased variance, eps 1e-5), written to Y for channels [h0, h1).
__device__ __forceinline__ void ln_row(const float* __restrict__ X, float* __restrict__ Y, int r, const float* __restrict__ lw,
                                       const float* __restrict__ lb, int h0, int h1) {
  const float* x = X + r * kHP;
  float s = 0.0f;
#pragma unroll 8
  for (int h = 0; h < kH; ++h) s += x[h];
  const float mean = s * (1.0f / kH);
  float v = 0.0f;
#pragma unroll 8
  for (int h = 0; h < kH; ++h) { const float d = x[h] - mean; v = fmaf(d, d, v); }
  const float rstd = rsqrtf(v * (1.0f / kH) + 1e-5f);
  float* y = Y + r * kHP;
#pragma unroll 8
  for (int h = h0; h < h1; ++h) y[h] = (x[h] - mean) * rstd * lw[h] + lb[h];
}

// x += PReLU(Conv1d_grouped(LN(x))) for a tile of nrows = 256 / NP rows; row r = tl * F + f, the conv runs along f
// inside each segment of F rows with zero padding.  W region: conv weights [g][k][i][o] (5760) | bias | prelu | ln_w | ln_b.
template <int NP>
__device__ __forceinline__ void fconv(float* __restrict__ X, float* __restrict__ Y, float* __restrict__ W,
                                      const fnssl_sn_fconv_weights& w, int F, int tid) {
  constexpr int nrows = kRows / NP;
  const int r = tid % nrows, part = tid / nrows;
  float* cw = W;
  float* cb = W + kG * kFK * kGC * kGC;
  float* pr = cb + kH;
  float* lw = pr + kH;
  float* lb = lw + kH;
  for (int i = tid; i < kG * kFK * kGC * kGC; i += kRows) cw[i] = __ldg(w.conv_wp + i);
  if (tid < kH) {
    cb[tid] = __ldg(w.conv_b + tid); pr[tid] = __ldg(w.prelu + tid);
    lw[tid] = __ldg(w.ln_w + tid);   lb[tid] = __ldg(w.ln_b + tid);
  }
  __syncthreads();
  ln_row(X, Y, r, lw, lb, part * (kH / NP), (part + 1) * (kH / NP));
  __syncthreads();
  const int f = r % F;
  for (int g = part; g < kG; g += NP) {
    float acc[kGC];
#pragma unroll
    for (int o = 0; o < kGC; ++o) acc[o] = cb[g * kGC + o];
#pragma unroll
    for (int k = 0; k < kFK; ++k) {
      const int ff = f + k - kFK / 2;
      if (ff < 0 || ff >= F) continue;
      const float* y = Y + (r + k - kFK / 2) * kHP + g * kGC;
      const float4* wv = reinterpret_cast<const float4*>(cw + (g * kFK + k) * kGC * kGC);
#pragma unroll
      for (int i = 0; i < kGC; ++i) {
        const float yi = y[i];
        const float4 a = wv[i * 3 + 0], b = wv[i * 3 + 1], c = wv[i * 3 + 2];
        acc[0] = fmaf(yi, a.x, acc[0]); acc[1] = fmaf(yi, a.y, acc[1]); acc[2] = fmaf(yi, a.z, acc[2]); acc[3] = fmaf(yi, a.w, acc[3]);
        acc[4] = fmaf(yi, b.x, acc[4]); acc[5] = fmaf(yi, b.y, acc[5]); acc[6] = fmaf(yi, b.z, acc[6]); acc[7] = fmaf(yi, b.w, acc[7]);
        acc[8] = fmaf(yi, c.x, acc[8]); acc[9] = fmaf(yi, c.y, acc[9]); acc[10] = fmaf(yi, c.z, acc[10]); acc[11] = fmaf(yi, c.w, acc[11]);
      }
    }
    float* x = X + r * kHP + g * kGC;
#pragma unroll
    for (int o = 0; o < kGC; ++o) {
      const float v = acc[o];
      x[o] += v > 0.0f ? v : pr[g * kGC + o] * v;
    }
  }
  __syncthreads();
}

// x += SiLU(unsqueeze(Linear_F(SiLU(squeeze(LN(x))))))  (IPDnet2.py:235-253).  W region: S [8][nrows] | L [nrows][9] |
// squeeze weights [h][8] | unsqueeze weights [h][8] | sq_b 8 | usq_b 96 | ln_w 96 | ln_b 96.
template <int NP>
__device__ __forceinline__ void full_band(float* __restrict__ X, float* __restrict__ Y, float* __restrict__ W,
                                          const fnssl_sn_freq_args& a, int F, int tid) {
  constexpr int nrows = kRows / NP;
  constexpr int JP = kSQ / NP;           // squeezed channels per thread
  const int r = tid % nrows, part = tid / nrows;
  float* S = W;
  float* L = S + kSQ * nrows;
  float* sqw = L + nrows * 9;
  float* usw = sqw + kH * kSQ;
  float* sqb = usw + kH * kSQ;
  float* usb = sqb + kSQ;
  float* lw = usb + kH;
  float* lb = lw + kH;
  for (int i = tid; i < kH * kSQ; i += kRows) { sqw[i] = __ldg(a.sq_wt + i); usw[i] = __ldg(a.usq_w + i); }
  if (tid < kH) { usb[tid] = __ldg(a.usq_b + tid); lw[tid] = __ldg(a.lnf_w + tid); lb[tid] = __ldg(a.lnf_b + tid); }
  if (tid < kSQ) sqb[tid] = __ldg(a.sq_b + tid);
  __syncthreads();
  ln_row(X, Y, r, lw, lb, part * (kH / NP), (part + 1) * (kH / NP));
  __syncthreads();
  {
    float acc[JP];
#pragma unroll
    for (int j = 0; j < JP; ++j) acc[j] = sqb[part * JP + j];
    const float* y = Y + r * kHP;
#pragma unroll 4
    for (int h = 0; h < kH; ++h) {
      const float yh = y[h];
#pragma unroll
      for (int j = 0; j < JP; ++j) acc[j] = fmaf(yh, sqw[h * kSQ + part * JP + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < JP; ++j) S[(part * JP + j) * nrows + r] = silu_f(acc[j]);
  }
  __syncthreads();
  {
    const int f = r % F, seg = r - f;
    float acc[JP];
    const float fb = __ldg(a.full_b + f);
#pragma unroll
    for (int j = 0; j < JP; ++j) acc[j] = fb;
    const float* wt = a.full_wt + f;                     // [f'][f]: coalesced across the threads of a segment
#pragma unroll 4
    for (int fp = 0; fp < F; ++fp) {
      const float wv = __ldg(wt + (size_t)fp * F);
#pragma unroll
      for (int j = 0; j < JP; ++j) acc[j] = fmaf(wv, S[(part * JP + j) * nrows + seg + fp], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < JP; ++j) L[r * 9 + part * JP + j] = acc[j];
  }
  __syncthreads();
  {
    float l[kSQ];
#pragma unroll
    for (int j = 0; j < kSQ; ++j) l[j] = L[r * 9 + j];
    float* x = X + r * kHP;
#pragma unroll 4
    for (int h = part * (kH / NP); h < (part + 1) * (kH / NP); ++h) {
      float acc = usb[h];
#pragma unroll
      for (int j = 0; j < kSQ; ++j) acc = fmaf(usw[h * kSQ + j], l[j], acc);
      x[h] += silu_f(acc);
    }
  }
  __syncthreads();
}

// FIRST: one CTA = one (b, t): encoder over all 256 bins -> fconv1 -> pool 2 -> full -> fconv2 -> pool 8 -> out (16 x 96).
// else : one CTA = 256 / nf frames of one utterance at nf bins: fconv1 -> full -> fconv2 -> out.
template <bool FIRST>
__global__ void __launch_bounds__(kRows, 1)
sn_freq_kernel(const fnssl_sn_freq_args a) {
  extern __shared__ __align__(16) float smem[];
  float* X = smem;
  float* Y = smem + kRows * kHP;
  float* W = smem + 2 * kRows * kHP;
  const int tid = threadIdx.x;
  if constexpr (FIRST) {
    const int b = blockIdx.y, t = blockIdx.x;
    // ---- encoder: causal Conv1d along t, cin -> 96, kernel 5 (IPDnet2.py:66-76): thread = bin f, 96 accumulators
    const int ld = a.x_ld;
    for (int i = tid; i < kEncK * ld * kH; i += kRows) W[i] = __ldg(a.enc_wp + i);      // [k][c < ld][h], zero rows beyond cin
    __syncthreads();
    {
      const int f = tid;
      float acc[kH];
#pragma unroll
      for (int h = 0; h < kH; ++h) acc[h] = __ldg(a.enc_b + h);
      for (int k = 0; k < kEncK; ++k) {
        const int tt = t - (kEncK - 1) + k;
        if (tt < 0) continue;
        const float4* src = reinterpret_cast<const float4*>(a.x + (((size_t)b * a.nt + tt) * a.nf + f) * ld);
        for (int c4 = 0; c4 < ld / 4; ++c4) {
          const float4 v = __ldg(src + c4);
          const float in[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4* wv = reinterpret_cast<const float4*>(W + ((k * ld) + c4 * 4 + q) * kH);
#pragma unroll
            for (int h4 = 0; h4 < kH / 4; ++h4) {
              const float4 ww = wv[h4];
              acc[h4 * 4 + 0] = fmaf(in[q], ww.x, acc[h4 * 4 + 0]);
              acc[h4 * 4 + 1] = fmaf(in[q], ww.y, acc[h4 * 4 + 1]);
              acc[h4 * 4 + 2] = fmaf(in[q], ww.z, acc[h4 * 4 + 2]);
              acc[h4 * 4 + 3] = fmaf(in[q], ww.w, acc[h4 * 4 + 3]);
            }
          }
        }
      }
      float* x = X + f * kHP;
#pragma unroll
      for (int h = 0; h < kH; ++h) x[h] = acc[h];
    }
    __syncthreads();
    fconv<1>(X, Y, W, a.fconv1, 256, tid);
    // ---- AvgPool over pairs of bins (:148): 256 rows of X -> 128 rows in the Y region, then swap roles
    for (int i = tid; i < 128 * kH; i += kRows) {
      const int r2 = i / kH, h = i - r2 * kH;
      Y[r2 * kHP + h] = 0.5f * (X[(2 * r2) * kHP + h] + X[(2 * r2 + 1) * kHP + h]);
    }
    __syncthreads();
    float* X2 = Y; float* Y2 = X;
    full_band<2>(X2, Y2, W, a, 128, tid);
    fconv<2>(X2, Y2, W, a.fconv2, 128, tid);
    // ---- AvgPool over 8 bins (:153) and store (b, t, 16, 96)
    float* out = a.out + ((size_t)b * a.nt + t) * 16 * kH;
    for (int i = tid; i < 16 * kH; i += kRows) {
      const int fc = i / kH, h = i - fc * kH;
      float s = 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += X2[(fc * 8 + j) * kHP + h];
      out[i] = s * 0.125f;
    }
  } else {
    const int F = a.nf;                       // 16
    const int tpb = kRows / F;                // frames per CTA
    const int b = blockIdx.y, t0 = blockIdx.x * tpb;
    const int nvalid = min(tpb, a.nt - t0) * F * kH;
    const float* src = a.x + ((size_t)b * a.nt + t0) * F * kH;
    for (int i = tid; i < kRows * kH; i += kRows) {
      const int r = i / kH, h = i - r * kH;
      X[r * kHP + h] = i < nvalid ? __ldg(src + i) : 0.0f;
    }
    __syncthreads();
    fconv<1>(X, Y, W, a.fconv1, F, tid);
    full_band<1>(X, Y, W, a, F, tid);
    fconv<1>(X, Y, W, a.fconv2, F, tid);
    float* out = a.out + ((size_t)b * a.nt + t0) * F * kH;
    for (int i = tid; i < nvalid; i += kRows) {
      const int r = i / kH, h = i - r * kH;
      out[i] = X[r * kHP + h];
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------
// time stage: x += Mamba_1(LN(x)); x += Mamba_2(LN(x)); optional AvgPool over 5 frames
// ----------------------------------------------------------------------------------------------------------------

constexpr int kDI = 192;        // d_inner = 2 * d_model
constexpr int kNS = 16;         // d_state
constexpr int kDR = 6;          // dt_rank = ceil(96 / 16)
constexpr int kDK = 4;          // d_conv
constexpr int kXP = kDR + 2 * kNS;   // 38 rows of x_proj
constexpr int kTT = 15;         // frames per chunk
constexpr int kUS = 20;         // row stride of the transposed LN tile [h][t]
constexpr int kDB = 40;         // row stride of the (B | C | dt) tile

struct MambaState {
  float h[kNS];                 // selective-scan state of this thread's channel
  float xprev[kDK - 1];         // causal conv history (pre-activation inner channel)
};

__device__ __forceinline__ void mamba_chunk(const fnssl_mamba_weights& w, MambaState& st, int nval,
                                            float* __restrict__ xs, float* __restrict__ uT, float* __restrict__ xc,
                                            float* __restrict__ zs, float* __restrict__ dbc, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  // ---- LayerNorm per frame (one warp per frame), stored transposed uT[h][t]
  for (int t = warp; t < kTT; t += kDI / 32) {
    const float* x = xs + t * kH;
    const float v0 = x[lane], v1 = x[lane + 32], v2 = x[lane + 64];
    float s = v0 + v1 + v2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / kH);
    const float d0 = v0 - mean, d1 = v1 - mean, d2 = v2 - mean;
    float q = d0 * d0 + d1 * d1 + d2 * d2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / kH) + 1e-5f);
    uT[lane * kUS + t] = d0 * rstd * __ldg(w.ln_w + lane) + __ldg(w.ln_b + lane);
    uT[(lane + 32) * kUS + t] = d1 * rstd * __ldg(w.ln_w + lane + 32) + __ldg(w.ln_b + lane + 32);
    uT[(lane + 64) * kUS + t] = d2 * rstd * __ldg(w.ln_w + lane + 64) + __ldg(w.ln_b + lane + 64);
  }
  __syncthreads();
  const int d = tid;
  // ---- in_proj (96 -> 2 x 192, no bias): thread = inner channel d, outer loop over the input channel
  {
    float ax[16], az[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) { ax[t] = 0.0f; az[t] = 0.0f; }
#pragma unroll 2
    for (int h = 0; h < kH; ++h) {
      const float w1 = __ldg(w.in_proj_wt + h * 2 * kDI + d);
      const float w2 = __ldg(w.in_proj_wt + h * 2 * kDI + kDI + d);
      const float4* u4 = reinterpret_cast<const float4*>(uT + h * kUS);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 u = u4[q];
        ax[q * 4 + 0] = fmaf(w1, u.x, ax[q * 4 + 0]); az[q * 4 + 0] = fmaf(w2, u.x, az[q * 4 + 0]);
        ax[q * 4 + 1] = fmaf(w1, u.y, ax[q * 4 + 1]); az[q * 4 + 1] = fmaf(w2, u.y, az[q * 4 + 1]);
        ax[q * 4 + 2] = fmaf(w1, u.z, ax[q * 4 + 2]); az[q * 4 + 2] = fmaf(w2, u.z, az[q * 4 + 2]);
        ax[q * 4 + 3] = fmaf(w1, u.w, ax[q * 4 + 3]); az[q * 4 + 3] = fmaf(w2, u.w, az[q * 4 + 3]);
      }
    }
    // ---- causal depthwise conv (k = 4) + bias + SiLU; the last three raw inputs carry over to the next chunk
    const float c0 = __ldg(w.conv_w + d * kDK + 0), c1 = __ldg(w.conv_w + d * kDK + 1), c2 = __ldg(w.conv_w + d * kDK + 2),
                c3 = __ldg(w.conv_w + d * kDK + 3), cb = __ldg(w.conv_b + d);
    float p0 = st.xprev[0], p1 = st.xprev[1], p2 = st.xprev[2];
#pragma unroll
    for (int t = 0; t < kTT; ++t) {
      const float v = ax[t];
      xc[t * kDI + d] = silu_f(fmaf(c0, p0, fmaf(c1, p1, fmaf(c2, p2, fmaf(c3, v, cb)))));
      zs[t * kDI + d] = az[t];
      if (t < nval) { p0 = p1; p1 = p2; p2 = v; }
    }
    st.xprev[0] = p0; st.xprev[1] = p1; st.xprev[2] = p2;
  }
  __syncthreads();
  // ---- x_proj (192 -> 6 + 16 + 16, no bias): thread = (output j, group of 3 frames)
  if (tid < kXP * 5) {
    const int j = tid % kXP, tg = tid / kXP;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    const float* x0 = xc + (tg * 3) * kDI;
#pragma unroll 2
    for (int dd = 0; dd < kDI; dd += 4) {
      const float w0 = __ldg(w.x_proj_wt + (dd + 0) * kXP + j), w1 = __ldg(w.x_proj_wt + (dd + 1) * kXP + j),
                  w2 = __ldg(w.x_proj_wt + (dd + 2) * kXP + j), w3 = __ldg(w.x_proj_wt + (dd + 3) * kXP + j);
      const float4 u0 = *reinterpret_cast<const float4*>(x0 + dd);
      const float4 u1 = *reinterpret_cast<const float4*>(x0 + kDI + dd);
      const float4 u2 = *reinterpret_cast<const float4*>(x0 + 2 * kDI + dd);
      a0 = fmaf(w0, u0.x, fmaf(w1, u0.y, fmaf(w2, u0.z, fmaf(w3, u0.w, a0))));
      a1 = fmaf(w0, u1.x, fmaf(w1, u1.y, fmaf(w2, u1.z, fmaf(w3, u1.w, a1))));
      a2 = fmaf(w0, u2.x, fmaf(w1, u2.y, fmaf(w2, u2.z, fmaf(w3, u2.w, a2))));
    }
    const int pos = j < kDR ? 2 * kNS + j : j - kDR;       // tile row = [B 16 | C 16 | dt 6 | pad 2]
    dbc[(tg * 3 + 0) * kDB + pos] = a0;
    dbc[(tg * 3 + 1) * kDB + pos] = a1;
    dbc[(tg * 3 + 2) * kDB + pos] = a2;
  }
  __syncthreads();
  // ---- dt_proj + softplus, selective scan, D skip, gate (y overwrites this thread's column of xc)
  {
    float wd[kDR], A[kNS];
#pragma unroll
    for (int r = 0; r < kDR; ++r) wd[r] = __ldg(w.dt_proj_w + d * kDR + r);
#pragma unroll
    for (int n = 0; n < kNS; ++n) A[n] = -__expf(__ldg(w.A_log + d * kNS + n));    // A = -exp(A_log), re-derived per chunk (registers)
    const float bd = __ldg(w.dt_proj_b + d), Dd = __ldg(w.D + d);
#pragma unroll 1
    for (int t = 0; t < nval; ++t) {
      const float* row = dbc + t * kDB;
      float dtv = bd;
#pragma unroll
      for (int r = 0; r < kDR; ++r) dtv = fmaf(wd[r], row[2 * kNS + r], dtv);
      const float delta = dtv > 20.0f ? dtv : log1pf(__expf(dtv));
      const float xv = xc[t * kDI + d];
      const float dx = delta * xv;
      float y = 0.0f;
#pragma unroll
      for (int n4 = 0; n4 < kNS / 4; ++n4) {
        const float4 Bv = *reinterpret_cast<const float4*>(row + n4 * 4);
        const float4 Cv = *reinterpret_cast<const float4*>(row + kNS + n4 * 4);
        const float bb[4] = {Bv.x, Bv.y, Bv.z, Bv.w}, cc[4] = {Cv.x, Cv.y, Cv.z, Cv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = n4 * 4 + q;
          st.h[n] = fmaf(__expf(delta * A[n]), st.h[n], dx * bb[q]);
          y = fmaf(cc[q], st.h[n], y);
        }
      }
      y = fmaf(Dd, xv, y);
      xc[t * kDI + d] = y * silu_f(zs[t * kDI + d]);
    }
    for (int t = nval; t < kTT; ++t) xc[t * kDI + d] = 0.0f;
  }
  __syncthreads();
  // ---- out_proj (192 -> 96, no bias) + residual: thread = (output m, half of the frames)
  {
    const int m = tid % kH, tg = tid / kH;
    const int tb = tg * 8;                     // frames [0, 8) or [8, 15) (+ one padded slot)
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
#pragma unroll 2
    for (int dd = 0; dd < kDI; dd += 4) {
      const float w0 = __ldg(w.out_proj_wt + (dd + 0) * kH + m), w1 = __ldg(w.out_proj_wt + (dd + 1) * kH + m),
                  w2 = __ldg(w.out_proj_wt + (dd + 2) * kH + m), w3 = __ldg(w.out_proj_wt + (dd + 3) * kH + m);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 yv = *reinterpret_cast<const float4*>(xc + (tb + i) * kDI + dd);     // row 15 is the zeroed pad row
        acc[i] = fmaf(w0, yv.x, fmaf(w1, yv.y, fmaf(w2, yv.z, fmaf(w3, yv.w, acc[i]))));
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (tb + i < nval) xs[(tb + i) * kH + m] += acc[i];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kDI, 3)
sn_time_kernel(const fnssl_sn_time_args a) {
  __shared__ __align__(16) float xs[kTT * kH];
  __shared__ __align__(16) float uT[kH * kUS];
  __shared__ __align__(16) float xc[(kTT + 1) * kDI];
  __shared__ __align__(16) float zs[kTT * kDI];
  __shared__ __align__(16) float dbc[kTT * kDB];
  const int tid = threadIdx.x;
  const int f = blockIdx.x, b = blockIdx.y;
  const int nf = a.nf, nt = a.nt;
  MambaState st[2];
#pragma unroll
  for (int m = 0; m < 2; ++m) {
#pragma unroll
    for (int n = 0; n < kNS; ++n) st[m].h[n] = 0.0f;
#pragma unroll
    for (int k = 0; k < kDK - 1; ++k) st[m].xprev[k] = 0.0f;
  }
  for (int i = tid; i < kDI; i += kDI) xc[kTT * kDI + i] = 0.0f;          // pad row read by the second out_proj group
  const int pool = a.pool;
  const int nt_out = nt / pool;
  for (int t0 = 0; t0 < nt; t0 += kTT) {
    const int nval = min(kTT, nt - t0);
    for (int i = tid; i < kTT * kH; i += kDI) {
      const int t = i / kH, h = i - t * kH;
      xs[i] = t < nval ? __ldg(a.x + (((size_t)b * nt + t0 + t) * nf + f) * kH + h) : 0.0f;
    }
    __syncthreads();
    mamba_chunk(a.m[0], st[0], nval, xs, uT, xc, zs, dbc, tid);
    mamba_chunk(a.m[1], st[1], nval, xs, uT, xc, zs, dbc, tid);
    if (pool == 1) {
      for (int i = tid; i < nval * kH; i += kDI) {
        const int t = i / kH, h = i - t * kH;
        a.out[(((size_t)b * nt + t0 + t) * nf + f) * kH + h] = xs[i];
      }
    } else {                                   // kTT is a multiple of the pooling width (5)
      for (int i = tid; i < (kTT / 5) * kH; i += kDI) {
        const int tp = i / kH, h = i - tp * kH;
        const int to = t0 / 5 + tp;
        if (to < nt_out) {
          float s = 0.0f;
#pragma unroll
          for (int j = 0; j < 5; ++j) s += xs[(tp * 5 + j) * kH + h];
          a.out[(((size_t)b * nt_out + to) * nf + f) * kH + h] = s * 0.2f;
        }
      }
    }
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------------------------------------------
// head: FreqInverse (per-band 1x1 conv 96 -> ratio*out, tanh) + decoder Linear + the reference's output reshape
// ----------------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
sn_head_kernel(const float* __restrict__ x, int nt, int nfc, const float* __restrict__ trans_wt, const float* __restrict__ trans_b,
               const float* __restrict__ dec_w, const float* __restrict__ dec_b, int dim_out, int ratio, int n_src,
               float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                               // [nfc][96]
  float* ys = sm + nfc * kH;                    // [nfc][ratio * dim_out]  (channel = o * ratio + j)
  const int tid = threadIdx.x;
  const int t = blockIdx.x, b = blockIdx.y;
  const int C = ratio * dim_out;
  const float* src = x + ((size_t)b * nt + t) * nfc * kH;
  for (int i = tid; i < nfc * kH; i += 256) xs[i] = __ldg(src + i);
  __syncthreads();
  for (int i = tid; i < nfc * C; i += 256) {
    const int fc = i / C, c = i - fc * C;
    float acc = __ldg(trans_b + c);
    const float* xr = xs + fc * kH;
#pragma unroll 4
    for (int h = 0; h < kH; ++h) acc = fmaf(__ldg(trans_wt + h * C + c), xr[h], acc);
    ys[i] = tanhf(acc);
  }
  __syncthreads();
  const int nF = nfc * ratio;
  const int K = dim_out / n_src, K2 = K / 2;
  float* dst = out + ((size_t)b * nt + t) * (size_t)(2 * nF) * K2 * n_src;
  for (int i = tid; i < nF * dim_out; i += 256) {
    const int f = i / dim_out, op = i - f * dim_out;
    const int fc = f / ratio, j = f - fc * ratio;
    float acc = __ldg(dec_b + op);
    const float* yr = ys + fc * C + j;
    for (int o = 0; o < dim_out; ++o) acc = fmaf(__ldg(dec_w + op * dim_out + o), yr[o * ratio], acc);
    const int s = op / K, k = op - s * K;
    dst[((size_t)(2 * f + k / K2) * K2 + (k % K2)) * n_src + s] = acc;
  }
}

__global__ void reflect_pad_kernel(const float* __restrict__ src, int n, int nch, int pad, float* __restrict__ dst, size_t total) {
  const int np = n + 2 * pad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % nch);
    const size_t q = i / nch;
    const int p = (int)(q % np);
    const size_t b = q / np;
    int j = p - pad;
    if (j < 0) j = -j;
    if (j >= n) j = 2 * (n - 1) - j;
    dst[i] = src[(b * n + j) * nch + ch];
  }
}

}  // namespace sn
}  // namespace fnssl

using namespace fnssl;

extern "C" {

int fnssl_reflect_pad(const float* signal, int nb, int nsample, int nch, int pad, float* out, void* stream) {
  FNSSL_REQUIRE(signal && out && nb > 0 && nch > 0 && pad >= 0, "reflect_pad: bad arguments");
  FNSSL_REQUIRE(nsample > pad, "reflect_pad: needs nsample (%d) > pad (%d), as torch.stft(center=True) does", nsample, pad);
  const size_t total = (size_t)nb * (nsample + 2 * pad) * nch;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  sn::reflect_pad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(signal, nsample, nch, pad, out, total);
  FNSSL_LAUNCH_CHECK("reflect_pad_kernel");
  return 0;
}

static int sn_check_fconv(const fnssl_sn_fconv_weights& w) {
  return w.ln_w && w.ln_b && w.conv_wp && w.conv_b && w.prelu;
}

int fnssl_sn_freq_forward(const fnssl_sn_freq_args* a, void* stream) {
  FNSSL_REQUIRE(a && a->x && a->out, "sn_freq: null argument");
  FNSSL_REQUIRE(a->hidden == sn::kH && a->squeeze == sn::kSQ && a->groups == sn::kG && a->fkernel == sn::kFK,
                "sn_freq: built for dim_hidden 96, dim_squeeze 8, 8 groups, kernel 5 (got %d/%d/%d/%d)", a->hidden, a->squeeze,
                a->groups, a->fkernel);
  FNSSL_REQUIRE(a->nb > 0 && a->nb <= 65535 && a->nt > 0, "sn_freq: bad nb/nt (%d/%d)", a->nb, a->nt);
  FNSSL_REQUIRE(sn_check_fconv(a->fconv1) && sn_check_fconv(a->fconv2) && a->lnf_w && a->lnf_b && a->sq_wt && a->sq_b &&
                a->full_wt && a->full_b && a->usq_w && a->usq_b, "sn_freq: null weight pointer");
  const size_t smem = (size_t)(2 * sn::kRows * sn::kHP + sn::kWFloats) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->is_first) {
    FNSSL_REQUIRE(a->nf == 256, "sn_freq(first): the first layer runs on 256 bins (got %d)", a->nf);
    FNSSL_REQUIRE(a->enc_wp && a->enc_b && a->enc_kernel == sn::kEncK, "sn_freq(first): encoder weights / kernel size 5 required");
    FNSSL_REQUIRE(a->cin > 0 && a->x_ld >= a->cin && a->x_ld % 4 == 0 && a->x_ld <= 16,
                  "sn_freq(first): cin %d / ld %d (ld must be a multiple of 4, <= 16)", a->cin, a->x_ld);
    FNSSL_REQUIRE((reinterpret_cast<uintptr_t>(a->x) & 15) == 0, "sn_freq(first): x must be 16-byte aligned");
    FNSSL_CUDA(cudaFuncSetAttribute(sn::sn_freq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(a->nt, a->nb);
    sn::sn_freq_kernel<true><<<grid, sn::kRows, smem, st>>>(*a);
  } else {
    FNSSL_REQUIRE(a->nf == 16, "sn_freq: later layers run on 16 bands (got %d)", a->nf);
    FNSSL_CUDA(cudaFuncSetAttribute(sn::sn_freq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tpb = sn::kRows / a->nf;
    dim3 grid((a->nt + tpb - 1) / tpb, a->nb);
    sn::sn_freq_kernel<false><<<grid, sn::kRows, smem, st>>>(*a);
  }
  FNSSL_LAUNCH_CHECK("sn_freq_kernel");
  return 0;
}

int fnssl_sn_time_forward(const fnssl_sn_time_args* a, void* stream) {
  FNSSL_REQUIRE(a && a->x && a->out, "sn_time: null argument");
  FNSSL_REQUIRE(a->hidden == sn::kH && a->d_inner == sn::kDI && a->d_state == sn::kNS && a->dt_rank == sn::kDR && a->d_conv == sn::kDK,
                "sn_time: built for Mamba(d_model 96, d_inner 192, d_state 16, dt_rank 6, d_conv 4)");
  FNSSL_REQUIRE(a->nb > 0 && a->nb <= 65535 && a->nt > 0 && a->nf > 0, "sn_time: bad nb/nt/nf");
  FNSSL_REQUIRE(a->pool == 1 || a->pool == 5, "sn_time: pool must be 1 or 5 (got %d)", a->pool);
  for (int m = 0; m < 2; ++m) {
    const fnssl_mamba_weights& w = a->m[m];
    FNSSL_REQUIRE(w.ln_w && w.ln_b && w.in_proj_wt && w.conv_w && w.conv_b && w.x_proj_wt && w.dt_proj_w && w.dt_proj_b &&
                  w.A_log && w.D && w.out_proj_wt, "sn_time: null weight pointer (mamba %d)", m);
  }
  dim3 grid(a->nf, a->nb);
  sn::sn_time_kernel<<<grid, sn::kDI, 0, (cudaStream_t)stream>>>(*a);
  FNSSL_LAUNCH_CHECK("sn_time_kernel");
  return 0;
}

int fnssl_sn_head_forward(const float* x, int nb, int nt, int nfc, int hidden, const float* trans_wt, const float* trans_b,
                          const float* dec_w, const float* dec_b, int dim_out, int ratio, int n_src, float* out, void* stream) {
  FNSSL_REQUIRE(x && trans_wt && trans_b && dec_w && dec_b && out, "sn_head: null pointer");
  FNSSL_REQUIRE(hidden == sn::kH, "sn_head: built for dim_hidden 96 (got %d)", hidden);
  FNSSL_REQUIRE(nb > 0 && nb <= 65535 && nt > 0 && nfc > 0 && ratio > 0, "sn_head: bad sizes");
  FNSSL_REQUIRE(n_src > 0 && dim_out > 0 && dim_out % (2 * n_src) == 0, "sn_head: dim_output %d is not 2 * n_src (%d) * pairs", dim_out,
                n_src);
  const size_t smem = ((size_t)nfc * sn::kH + (size_t)nfc * ratio * dim_out) * sizeof(float);
  FNSSL_REQUIRE(smem <= 200 * 1024, "sn_head: dim_output too large (%d)", dim_out);
  FNSSL_CUDA(cudaFuncSetAttribute(sn::sn_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(nt, nb);
  sn::sn_head_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, nt, nfc, trans_wt, trans_b, dec_w, dec_b, dim_out, ratio, n_src, out);
  FNSSL_LAUNCH_CHECK("sn_head_kernel");
  return 0;
}

}  // extern "C"
