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
constexpr int kG = 8;           // conv groups
constexpr int kGC = 12;         // channels per group
constexpr int kFK = 5;          // kernel size along F
constexpr int kSQ = 8;          // dim_squeeze
constexpr int kRows = 256;      // rows of the resident tile
constexpr int kQ = 4;           // row quarters
constexpr int kFT = kH * kQ;    // 384 threads: thread = (channel, row quarter)
constexpr int kEncK = 5;

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ----------------------------------------------------------------------------------------------------------------
// frequency stage.  The tile [rows x 96] lives in shared memory twice: X (the residual stream) and Y (the normalised
// operand of the current module).  The three heavy modules -- encoder, fconv1, fconv2: 93 % of the MACs -- run with
// thread = (output channel, row quarter): the thread's weights (80 / 60 / 60 floats) sit in REGISTERS for the whole
// tile, and the operand rows are read with warp-broadcast LDS.128 (the 32 lanes of a warp are 32 channels of the same
// row), so the inner loops are 3..20 shared-memory instructions per 60..80 FMAs.
// ----------------------------------------------------------------------------------------------------------------

// LayerNorm over the 96 channels (biased variance, eps 1e-5) of rows [0, nrows): one warp per row, Y = LN(X).
__device__ __forceinline__ void ln_rows(const float* __restrict__ X, float* __restrict__ Y, int nrows,
                                        const float* __restrict__ lw, const float* __restrict__ lb, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  const float w0 = __ldg(lw + lane), w1 = __ldg(lw + lane + 32), w2 = __ldg(lw + lane + 64);
  const float b0 = __ldg(lb + lane), b1 = __ldg(lb + lane + 32), b2 = __ldg(lb + lane + 64);
  constexpr int kNW = kFT / 32;
  for (int r = warp; r < nrows; r += 2 * kNW) {       // two independent rows per iteration: the shuffle chains overlap
    const int r2 = r + kNW;
    const bool has2 = r2 < nrows;
    const float* xa = X + r * kH;
    const float* xb = X + (has2 ? r2 : r) * kH;
    const float a0 = xa[lane], a1 = xa[lane + 32], a2 = xa[lane + 64];
    const float c0 = xb[lane], c1 = xb[lane + 32], c2 = xb[lane + 64];
    float sa = a0 + a1 + a2, sb = c0 + c1 + c2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
    const float ma = sa * (1.0f / kH), mb = sb * (1.0f / kH);
    const float da0 = a0 - ma, da1 = a1 - ma, da2 = a2 - ma;
    const float db0 = c0 - mb, db1 = c1 - mb, db2 = c2 - mb;
    float qa = da0 * da0 + da1 * da1 + da2 * da2, qb = db0 * db0 + db1 * db1 + db2 * db2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { qa += __shfl_xor_sync(0xffffffffu, qa, o); qb += __shfl_xor_sync(0xffffffffu, qb, o); }
    const float ra = rsqrtf(qa * (1.0f / kH) + 1e-5f), rb = rsqrtf(qb * (1.0f / kH) + 1e-5f);
    float* ya = Y + r * kH;
    ya[lane] = da0 * ra * w0 + b0; ya[lane + 32] = da1 * ra * w1 + b1; ya[lane + 64] = da2 * ra * w2 + b2;
    if (has2) {
      float* yb = Y + r2 * kH;
      yb[lane] = db0 * rb * w0 + b0; yb[lane + 32] = db1 * rb * w1 + b1; yb[lane + 64] = db2 * rb * w2 + b2;
    }
  }
}

__device__ __forceinline__ void load_row12(float (&dst)[kGC], const float* __restrict__ Y, int row, int lo, int hi, int g) {
  if (row >= lo && row < hi) {
    const float4* p = reinterpret_cast<const float4*>(Y + row * kH + g * kGC);
    const float4 a = p[0], b = p[1], c = p[2];
    dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w; dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
    dst[8] = c.x; dst[9] = c.y; dst[10] = c.z; dst[11] = c.w;
  } else {
#pragma unroll
    for (int i = 0; i < kGC; ++i) dst[i] = 0.0f;
  }
}

// x += PReLU(Conv1d_grouped(LN(x))) on nrows rows made of segments of F rows (the conv runs along the rows of a segment,
// zero padded).  Thread = (channel c = (g, o), quarter q): 60 weights in registers, a 5-row window of the group's 12
// normalised channels slides down the quarter's rows (3 LDS.128 + 60 FMA per output).
__device__ __forceinline__ void fconv(float* __restrict__ X, float* __restrict__ Y, const fnssl_sn_fconv_weights& w,
                                      int nrows, int F, int tid) {
  ln_rows(X, Y, nrows, w.ln_w, w.ln_b, tid);
  const int c = tid % kH, q = tid / kH;
  const int g = c / kGC, o = c - g * kGC;
  float wr[kFK][kGC];
#pragma unroll
  for (int k = 0; k < kFK; ++k)
#pragma unroll
    for (int i = 0; i < kGC; ++i) wr[k][i] = __ldg(w.conv_wp + ((g * kFK + k) * kGC + i) * kGC + o);
  const float bias = __ldg(w.conv_b + c), slope = __ldg(w.prelu + c);
  __syncthreads();
  const int per = nrows / kQ;
  for (int s0 = q * per; s0 < (q + 1) * per; s0 += F) {      // one pass per segment (or per quarter when F > per)
    const int lo = (s0 / F) * F, hi = lo + F;                // segment bounds of this run of rows
    const int r1 = min(s0 + F, (q + 1) * per);
    float win[kFK][kGC];
    load_row12(win[0], Y, s0 - 2, lo, hi, g);
    load_row12(win[1], Y, s0 - 1, lo, hi, g);
    load_row12(win[2], Y, s0, lo, hi, g);
    load_row12(win[3], Y, s0 + 1, lo, hi, g);
    for (int r = s0; r < r1; r += kFK) {
#pragma unroll
      for (int u = 0; u < kFK; ++u) {
        if (r + u < r1) {
          load_row12(win[(u + 4) % kFK], Y, r + u + 2, lo, hi, g);
          float a0 = bias, a1 = 0.0f, a2 = 0.0f;
#pragma unroll
          for (int k = 0; k < kFK; ++k) {
#pragma unroll
            for (int i = 0; i < kGC; i += 3) {
              a0 = fmaf(wr[k][i], win[(u + k) % kFK][i], a0);
              a1 = fmaf(wr[k][i + 1], win[(u + k) % kFK][i + 1], a1);
              a2 = fmaf(wr[k][i + 2], win[(u + k) % kFK][i + 2], a2);
            }
          }
          const float v = a0 + (a1 + a2);
          X[(r + u) * kH + c] += v > 0.0f ? v : slope * v;
        }
      }
    }
  }
  __syncthreads();
}

// x += SiLU(unsqueeze(Linear_F(SiLU(squeeze(LN(x))))))  (IPDnet2.py:235-253) on nrows rows in segments of F rows.
// S [8][nrows] and L [nrows][8] live in the scratch region W.
__device__ __forceinline__ void full_band(float* __restrict__ X, float* __restrict__ Y, float* __restrict__ W,
                                          const fnssl_sn_freq_args& a, int nrows, int F, int tid) {
  float* S = W;
  float* L = W + kSQ * kRows;
  ln_rows(X, Y, nrows, a.lnf_w, a.lnf_b, tid);
  __syncthreads();
  {   // squeeze 96 -> 8 (+ SiLU): thread = (j, row slot), the 96 weights of output j in registers
    const int j = tid % kSQ, slot = tid / kSQ;
    float wq[kH];
#pragma unroll
    for (int h = 0; h < kH; ++h) wq[h] = __ldg(a.sq_wt + h * kSQ + j);
    const float bj = __ldg(a.sq_b + j);
    for (int r = slot; r < nrows; r += kFT / kSQ) {
      const float4* y = reinterpret_cast<const float4*>(Y + r * kH);
      float a0 = bj, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
      for (int h4 = 0; h4 < kH / 4; ++h4) {
        const float4 v = y[h4];
        a0 = fmaf(wq[h4 * 4 + 0], v.x, a0); a1 = fmaf(wq[h4 * 4 + 1], v.y, a1);
        a2 = fmaf(wq[h4 * 4 + 2], v.z, a2); a3 = fmaf(wq[h4 * 4 + 3], v.w, a3);
      }
      S[j * nrows + r] = silu_f((a0 + a1) + (a2 + a3));
    }
  }
  __syncthreads();
  if (tid < kRows) {   // Linear along the rows of a segment: thread = (row, group of squeezed channels)
    const int r = tid % nrows, jg = tid / nrows, njg = kRows / nrows, JP = kSQ / njg;
    const int f = r % F, seg = r - f;
    const float fb = __ldg(a.full_b + f);
    float acc[kSQ];
#pragma unroll
    for (int j = 0; j < kSQ; ++j) acc[j] = fb;
    const float* wt = a.full_wt + f;                     // [f'][f]: coalesced across the threads of a segment
    const float* Sj = S + (jg * JP) * nrows + seg;
    if (JP == 4) {
#pragma unroll 4
      for (int fp = 0; fp < F; ++fp) {
        const float wv = __ldg(wt + (size_t)fp * F);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(wv, Sj[j * nrows + fp], acc[j]);
      }
    } else {
#pragma unroll 4
      for (int fp = 0; fp < F; ++fp) {
        const float wv = __ldg(wt + (size_t)fp * F);
#pragma unroll
        for (int j = 0; j < kSQ; ++j) acc[j] = fmaf(wv, Sj[j * nrows + fp], acc[j]);
      }
    }
    for (int j = 0; j < JP; ++j) L[r * kSQ + jg * JP + j] = acc[j];
  }
  __syncthreads();
  {   // unsqueeze 8 -> 96 (+ SiLU) and residual: thread = (channel, quarter)
    const int c = tid % kH, q = tid / kH;
    float wu[kSQ];
#pragma unroll
    for (int j = 0; j < kSQ; ++j) wu[j] = __ldg(a.usq_w + c * kSQ + j);
    const float bu = __ldg(a.usq_b + c);
    const int per = nrows / kQ;
    for (int r = q * per; r < (q + 1) * per; ++r) {
      const float4 l0 = *reinterpret_cast<const float4*>(L + r * kSQ), l1 = *reinterpret_cast<const float4*>(L + r * kSQ + 4);
      float acc = bu;
      acc = fmaf(wu[0], l0.x, acc); acc = fmaf(wu[1], l0.y, acc); acc = fmaf(wu[2], l0.z, acc); acc = fmaf(wu[3], l0.w, acc);
      acc = fmaf(wu[4], l1.x, acc); acc = fmaf(wu[5], l1.y, acc); acc = fmaf(wu[6], l1.z, acc); acc = fmaf(wu[7], l1.w, acc);
      X[r * kH + c] += silu_f(acc);
    }
  }
  __syncthreads();
}

// Encoder: causal Conv1d along t, cin -> 96, kernel 5 (IPDnet2.py:66-76,335) for the 256 bins of frame t.  The 5 input
// frames [k][f][LD] are staged in `in`; thread = (output channel, bin quarter) with its 5 * LD weights in registers.
template <int LD>
__device__ __forceinline__ void encoder(const fnssl_sn_freq_args& a, int b, int t, float* __restrict__ X, float* __restrict__ in,
                                        int tid) {
  for (int k = 0; k < kEncK; ++k) {
    const int tt = t - (kEncK - 1) + k;
    float4* dst = reinterpret_cast<float4*>(in + k * 256 * LD);
    if (tt >= 0) {
      const float4* src = reinterpret_cast<const float4*>(a.x + ((size_t)b * a.nt + tt) * 256 * LD);
      for (int i = tid; i < 256 * LD / 4; i += kFT) dst[i] = __ldg(src + i);
    } else {
      for (int i = tid; i < 256 * LD / 4; i += kFT) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const int h = tid % kH, q = tid / kH;
  float wr[kEncK][LD];
#pragma unroll
  for (int k = 0; k < kEncK; ++k)
#pragma unroll
    for (int c = 0; c < LD; ++c) wr[k][c] = __ldg(a.enc_wp + (k * LD + c) * kH + h);
  const float bias = __ldg(a.enc_b + h);
  __syncthreads();
#pragma unroll 4
  for (int f = q * 64; f < (q + 1) * 64; ++f) {
    float a0 = bias, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
    for (int k = 0; k < kEncK; ++k) {
      const float4* p = reinterpret_cast<const float4*>(in + (k * 256 + f) * LD);
#pragma unroll
      for (int c4 = 0; c4 < LD / 4; ++c4) {
        const float4 v = p[c4];
        a0 = fmaf(wr[k][c4 * 4 + 0], v.x, a0); a1 = fmaf(wr[k][c4 * 4 + 1], v.y, a1);
        a2 = fmaf(wr[k][c4 * 4 + 2], v.z, a2); a3 = fmaf(wr[k][c4 * 4 + 3], v.w, a3);
      }
    }
    X[f * kH + h] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
}

// LD > 0: first layer, one CTA = one (b, t): encoder over 256 bins -> fconv1 -> pool 2 -> full -> fconv2 -> pool 8 -> (16 x 96).
// LD = 0: later layers, one CTA = 16 frames x 16 bands of one utterance: fconv1 -> full -> fconv2.
template <int LD>
__global__ void __launch_bounds__(kFT, 1)
sn_freq_kernel(const fnssl_sn_freq_args a) {
  extern __shared__ __align__(16) float smem[];
  float* X = smem;
  float* Y = smem + kRows * kH;
  float* W = smem + 2 * kRows * kH;
  const int tid = threadIdx.x;
  if constexpr (LD > 0) {
    const int b = blockIdx.y, t = blockIdx.x + a.t_begin;     // frames [0, t_begin) are history for the causal encoder only
    encoder<LD>(a, b, t, X, Y, tid);
    fconv(X, Y, a.fconv1, 256, 256, tid);
    // ---- AvgPool over pairs of bins (:148): 256 rows of X -> 128 rows in the Y region, then swap roles
    for (int i = tid; i < 128 * kH; i += kFT) {
      const int r2 = i / kH, h = i - r2 * kH;
      Y[i] = 0.5f * (X[(2 * r2) * kH + h] + X[(2 * r2 + 1) * kH + h]);
    }
    __syncthreads();
    float* X2 = Y; float* Y2 = X;
    full_band(X2, Y2, W, a, 128, 128, tid);
    fconv(X2, Y2, a.fconv2, 128, 128, tid);
    // ---- AvgPool over 8 bins (:153) and store (b, t, 16, 96)
    float* out = a.out + ((size_t)b * (a.nt - a.t_begin) + (t - a.t_begin)) * 16 * kH;
    for (int i = tid; i < 16 * kH; i += kFT) {
      const int fc = i / kH, h = i - fc * kH;
      float s = 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += X2[(fc * 8 + j) * kH + h];
      out[i] = s * 0.125f;
    }
  } else {
    const int F = a.nf;                       // 16
    const int tpb = kRows / F;                // frames per CTA
    const int b = blockIdx.y, t0 = blockIdx.x * tpb;
    const int nvalid = min(tpb, a.nt - t0) * F * kH;
    const float* src = a.x + ((size_t)b * a.nt + t0) * F * kH;
    for (int i = tid; i < kRows * kH; i += kFT) X[i] = i < nvalid ? __ldg(src + i) : 0.0f;
    __syncthreads();
    fconv(X, Y, a.fconv1, kRows, F, tid);
    full_band(X, Y, W, a, kRows, F, tid);
    fconv(X, Y, a.fconv2, kRows, F, tid);
    float* out = a.out + ((size_t)b * a.nt + t0) * F * kH;
    for (int i = tid; i < nvalid; i += kFT) out[i] = X[i];
  }
}

// ----------------------------------------------------------------------------------------------------------------
// time stage: x += Mamba(LN(x)) along t for every (utterance, band) sequence; one launch per Mamba block, the second
// one also applies the 5-frame AvgPool.  One CTA = one sequence, 192 threads = the inner channels; the sequence is
// processed in chunks of 15 frames with the selective-scan state and the conv history carried in registers.
// The three projections run with the thread's weights in registers (48 at a time) against warp-broadcast LDS.128 reads
// of the chunk's activations: 12 shared-memory instructions per 48 FMAs, no global load inside the frame loops.
// ----------------------------------------------------------------------------------------------------------------

constexpr int kDI = 192;        // d_inner = 2 * d_model
constexpr int kNS = 16;         // d_state
constexpr int kDR = 6;          // dt_rank = ceil(96 / 16)
constexpr int kDK = 4;          // d_conv
constexpr int kXP = kDR + 2 * kNS;   // 38 rows of x_proj
constexpr int kTT = 15;         // frames per chunk
constexpr int kDB = 40;         // row stride of the (B | C | dt) tile
constexpr int kKW = 32;         // weights held in registers per pass (in_proj / out_proj)
constexpr int kKX = 24;         // ... per pass of x_proj (K quarter of 48 = 2 passes)

// acc[t] += sum_{i < KW} w[i] * src[t * stride + i]   (src: shared memory, 16-byte aligned rows)
template <int STRIDE, int KW>
__device__ __forceinline__ void mv(float (&acc)[kTT], const float (&w)[KW], const float* __restrict__ src) {
#pragma unroll
  for (int t = 0; t < kTT; ++t) {
    const float4* p = reinterpret_cast<const float4*>(src + t * STRIDE);
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
    for (int i4 = 0; i4 < KW / 4; ++i4) {
      const float4 v = p[i4];
      a0 = fmaf(w[i4 * 4 + 0], v.x, a0); a1 = fmaf(w[i4 * 4 + 1], v.y, a1);
      a2 = fmaf(w[i4 * 4 + 2], v.z, a2); a3 = fmaf(w[i4 * 4 + 3], v.w, a3);
    }
    acc[t] += (a0 + a1) + (a2 + a3);
  }
}

__global__ void __launch_bounds__(kDI, 3)
sn_time_kernel(const float* __restrict__ x, float* __restrict__ out, int nt, int nf, int pool, const fnssl_mamba_weights w,
               float* __restrict__ state, int state_flags) {
  __shared__ __align__(16) float xs[kTT * kH];            // residual stream of the chunk
  __shared__ __align__(16) float us[kTT * kH];            // LN(x)
  __shared__ __align__(16) float xc[kTT * kDI];           // conv + SiLU output; overwritten with the gated scan output
  __shared__ __align__(16) float zs[kTT * kDI];           // gate branch
  __shared__ __align__(16) float dbp[kTT * 4 * kDB];      // x_proj partial sums [t][K quarter][pos]
  __shared__ __align__(16) float dbc[kTT * kDB];          // [t][B 16 | C 16 | dt 6 | pad 2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.x, b = blockIdx.y;
  const int d = tid;
  float hst[kNS];
#pragma unroll
  for (int n = 0; n < kNS; ++n) hst[n] = 0.0f;
  float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;                  // causal conv history (raw inner channel of the last 3 frames)
  float* sp = state ? state + ((size_t)b * nf + f) * (kNS + kDK - 1) * kDI + d : nullptr;   // [seq][16 + 3][192]
  if (sp && (state_flags & 1)) {                          // a chunk of a longer stream resumes the scan
#pragma unroll
    for (int n = 0; n < kNS; ++n) hst[n] = sp[n * kDI];
    p0 = sp[kNS * kDI]; p1 = sp[(kNS + 1) * kDI]; p2 = sp[(kNS + 2) * kDI];
  }
  const int nt_out = nt / pool;
  const float lw0 = __ldg(w.ln_w + lane), lw1 = __ldg(w.ln_w + lane + 32), lw2 = __ldg(w.ln_w + lane + 64);
  const float lb0 = __ldg(w.ln_b + lane), lb1 = __ldg(w.ln_b + lane + 32), lb2 = __ldg(w.ln_b + lane + 64);

  for (int t0 = 0; t0 < nt; t0 += kTT) {
    const int nval = min(kTT, nt - t0);
    for (int i = tid; i < kTT * kH; i += kDI) {
      const int t = i / kH, h = i - t * kH;
      xs[i] = t < nval ? __ldg(x + (((size_t)b * nt + t0 + t) * nf + f) * kH + h) : 0.0f;
    }
    __syncthreads();
    // ---- LayerNorm per frame (one warp per frame)
    for (int t = warp; t < kTT; t += kDI / 32) {
      const float* xr = xs + t * kH;
      const float v0 = xr[lane], v1 = xr[lane + 32], v2 = xr[lane + 64];
      float s = v0 + v1 + v2;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * (1.0f / kH);
      const float d0 = v0 - mean, d1 = v1 - mean, d2 = v2 - mean;
      float q = d0 * d0 + d1 * d1 + d2 * d2;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q * (1.0f / kH) + 1e-5f);
      us[t * kH + lane] = d0 * rstd * lw0 + lb0;
      us[t * kH + lane + 32] = d1 * rstd * lw1 + lb1;
      us[t * kH + lane + 64] = d2 * rstd * lw2 + lb2;
    }
    __syncthreads();
    // ---- in_proj (96 -> 2 x 192, no bias): the x and the gate output of channel d share every LDS.128 of LN(x)
    //      (16 + 16 weights in registers per pass, 4 shared-memory reads per 32 FMAs) + causal depthwise conv (k = 4) + SiLU
    {
      float ax[kTT], az[kTT];
#pragma unroll
      for (int t = 0; t < kTT; ++t) { ax[t] = 0.0f; az[t] = 0.0f; }
#pragma unroll 1
      for (int ps = 0; ps < kH / 16; ++ps) {
        float wx[16], wz[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          wx[i] = __ldg(w.in_proj_wt + (ps * 16 + i) * 2 * kDI + d);
          wz[i] = __ldg(w.in_proj_wt + (ps * 16 + i) * 2 * kDI + kDI + d);
        }
#pragma unroll
        for (int t = 0; t < kTT; ++t) {
          const float4* p4 = reinterpret_cast<const float4*>(us + t * kH + ps * 16);
          float x0 = 0.0f, x1 = 0.0f, z0 = 0.0f, z1 = 0.0f;
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 v = p4[i4];
            x0 = fmaf(wx[i4 * 4 + 0], v.x, x0); x1 = fmaf(wx[i4 * 4 + 1], v.y, x1);
            x0 = fmaf(wx[i4 * 4 + 2], v.z, x0); x1 = fmaf(wx[i4 * 4 + 3], v.w, x1);
            z0 = fmaf(wz[i4 * 4 + 0], v.x, z0); z1 = fmaf(wz[i4 * 4 + 1], v.y, z1);
            z0 = fmaf(wz[i4 * 4 + 2], v.z, z0); z1 = fmaf(wz[i4 * 4 + 3], v.w, z1);
          }
          ax[t] += x0 + x1; az[t] += z0 + z1;
        }
      }
      const float c0 = __ldg(w.conv_w + d * kDK + 0), c1 = __ldg(w.conv_w + d * kDK + 1),
                  c2 = __ldg(w.conv_w + d * kDK + 2), c3 = __ldg(w.conv_w + d * kDK + 3), cb = __ldg(w.conv_b + d);
#pragma unroll
      for (int t = 0; t < kTT; ++t) {
        const float v = ax[t];
        xc[t * kDI + d] = silu_f(fmaf(c0, p0, fmaf(c1, p1, fmaf(c2, p2, fmaf(c3, v, cb)))));
        zs[t * kDI + d] = az[t];
        if (t < nval) { p0 = p1; p1 = p2; p2 = v; }
      }
    }
    __syncthreads();
    // ---- x_proj (192 -> 6 + 16 + 16, no bias): thread = (output j, K quarter), partial sums reduced through smem
    if (tid < kXP * 4) {
      const int j = tid % kXP, part = tid / kXP;
      float acc[kTT], wr[kKX];
#pragma unroll
      for (int t = 0; t < kTT; ++t) acc[t] = 0.0f;
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
#pragma unroll
        for (int i = 0; i < kKX; ++i) wr[i] = __ldg(w.x_proj_wt + (part * 48 + ps * kKX + i) * kXP + j);
        mv<kDI, kKX>(acc, wr, xc + part * 48 + ps * kKX);
      }
      const int pos = j < kDR ? 2 * kNS + j : j - kDR;      // tile row = [B 16 | C 16 | dt 6 | pad 2]
#pragma unroll
      for (int t = 0; t < kTT; ++t) dbp[(t * 4 + part) * kDB + pos] = acc[t];
    }
    __syncthreads();
    for (int i = tid; i < kTT * kDB; i += kDI) {
      const int t = i / kDB, pos = i - t * kDB;
      if (pos < kXP) {
        const float* q = dbp + t * 4 * kDB + pos;
        dbc[i] = (q[0] + q[kDB]) + (q[2 * kDB] + q[3 * kDB]);
      }
    }
    __syncthreads();
    // ---- dt_proj + softplus, selective scan, D skip, gate (y overwrites this thread's column of xc)
    {
      float wd[kDR], A[kNS];
#pragma unroll
      for (int r = 0; r < kDR; ++r) wd[r] = __ldg(w.dt_proj_w + d * kDR + r);
#pragma unroll
      for (int n = 0; n < kNS; ++n) A[n] = -1.4426950408889634f * __expf(__ldg(w.A_log + d * kNS + n));   // A log2(e), A = -exp(A_log)
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
        float y0 = 0.0f, y1 = 0.0f;
#pragma unroll
        for (int n4 = 0; n4 < kNS / 4; ++n4) {
          const float4 Bv = *reinterpret_cast<const float4*>(row + n4 * 4);
          const float4 Cv = *reinterpret_cast<const float4*>(row + kNS + n4 * 4);
          const float bb[4] = {Bv.x, Bv.y, Bv.z, Bv.w}, cc[4] = {Cv.x, Cv.y, Cv.z, Cv.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int n = n4 * 4 + q;
            hst[n] = fmaf(ex2_fast(delta * A[n]), hst[n], dx * bb[q]);      // exp(delta A) = 2^(delta A log2 e)
            if (q & 1) y1 = fmaf(cc[q], hst[n], y1); else y0 = fmaf(cc[q], hst[n], y0);
          }
        }
        const float y = fmaf(Dd, xv, y0 + y1);
        xc[t * kDI + d] = y * silu_f(zs[t * kDI + d]);
      }
      for (int t = nval; t < kTT; ++t) xc[t * kDI + d] = 0.0f;
    }
    __syncthreads();
    // ---- out_proj (192 -> 96, no bias) + residual: thread = (output m, K half); the halves add one after the other
    {
      const int m = tid % kH, half = tid / kH;
      float acc[kTT], wr[kKW];
#pragma unroll
      for (int t = 0; t < kTT; ++t) acc[t] = 0.0f;
#pragma unroll
      for (int ps = 0; ps < kH / kKW; ++ps) {
#pragma unroll
        for (int i = 0; i < kKW; ++i) wr[i] = __ldg(w.out_proj_wt + (half * kH + ps * kKW + i) * kH + m);
        mv<kDI, kKW>(acc, wr, xc + half * kH + ps * kKW);
      }
      if (half == 0) {
#pragma unroll
        for (int t = 0; t < kTT; ++t) xs[t * kH + m] += acc[t];
      }
      __syncthreads();
      if (half == 1) {
#pragma unroll
        for (int t = 0; t < kTT; ++t) xs[t * kH + m] += acc[t];
      }
    }
    __syncthreads();
    if (pool == 1) {
      for (int i = tid; i < nval * kH; i += kDI) {
        const int t = i / kH, h = i - t * kH;
        out[(((size_t)b * nt + t0 + t) * nf + f) * kH + h] = xs[i];
      }
    } else {                                   // kTT is a multiple of the pooling width (5)
      for (int i = tid; i < (kTT / 5) * kH; i += kDI) {
        const int tp = i / kH, h = i - tp * kH;
        const int to = t0 / 5 + tp;
        if (to < nt_out) {
          float sum = 0.0f;
#pragma unroll
          for (int j = 0; j < 5; ++j) sum += xs[(tp * 5 + j) * kH + h];
          out[(((size_t)b * nt_out + to) * nf + f) * kH + h] = sum * 0.2f;
        }
      }
    }
    __syncthreads();
  }
  if (sp && (state_flags & 2)) {
#pragma unroll
    for (int n = 0; n < kNS; ++n) sp[n * kDI] = hst[n];
    sp[kNS * kDI] = p0; sp[(kNS + 1) * kDI] = p1; sp[(kNS + 2) * kDI] = p2;
  }
}

// ----------------------------------------------------------------------------------------------------------------
// head: FreqInverse (per-band 1x1 conv 96 -> ratio*out, tanh) + decoder Linear + the reference's output reshape
// ----------------------------------------------------------------------------------------------------------------

constexpr int kMaxOut = 32;     // dim_output <= 32 (8 mics x 2 sources = 28)

__global__ void __launch_bounds__(256)
sn_head_kernel(const float* __restrict__ x, int nt, int nfc, const float* __restrict__ trans_wt, const float* __restrict__ trans_b,
               const float* __restrict__ dec_w, const float* __restrict__ dec_b, int dim_out, int ratio, int n_src,
               float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                               // [nfc][96]
  float* dw = sm + nfc * kH;                    // [dim_out][kMaxOut] decoder weights, zero padded
  float* ys = dw + kMaxOut * kMaxOut;           // [nfc][ratio * dim_out]  (channel = o * ratio + j)
  const int tid = threadIdx.x;
  const int t = blockIdx.x, b = blockIdx.y;
  const int C = ratio * dim_out;
  const float* src = x + ((size_t)b * nt + t) * nfc * kH;
  for (int i = tid; i < nfc * kH; i += 256) xs[i] = __ldg(src + i);
  for (int i = tid; i < kMaxOut * kMaxOut; i += 256) {
    const int op = i / kMaxOut, o = i - op * kMaxOut;
    dw[i] = (op < dim_out && o < dim_out) ? __ldg(dec_w + op * dim_out + o) : 0.0f;
  }
  __syncthreads();
  // FreqInverse: thread = output channel c with its 96 weights in registers; the band rows are warp-broadcast reads
  for (int c = tid; c < C; c += 256) {
    float wr[kH];
#pragma unroll
    for (int h = 0; h < kH; ++h) wr[h] = __ldg(trans_wt + h * C + c);
    const float bc = __ldg(trans_b + c);
    for (int fc = 0; fc < nfc; ++fc) {
      const float4* xr = reinterpret_cast<const float4*>(xs + fc * kH);
      float a0 = bc, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
      for (int h4 = 0; h4 < kH / 4; ++h4) {
        const float4 v = xr[h4];
        a0 = fmaf(wr[h4 * 4 + 0], v.x, a0); a1 = fmaf(wr[h4 * 4 + 1], v.y, a1);
        a2 = fmaf(wr[h4 * 4 + 2], v.z, a2); a3 = fmaf(wr[h4 * 4 + 3], v.w, a3);
      }
      ys[fc * C + c] = tanhf((a0 + a1) + (a2 + a3));
    }
  }
  __syncthreads();
  // decoder Linear + the reference's reshape: thread = frequency f, its dim_out tanh values in registers
  const int nF = nfc * ratio;
  const int K = dim_out / n_src, K2 = K / 2;
  float* dst = out + ((size_t)b * nt + t) * (size_t)(2 * nF) * K2 * n_src;
  for (int f = tid; f < nF; f += 256) {
    const int fc = f / ratio, j = f - fc * ratio;
    float y[kMaxOut];
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) y[o] = o < dim_out ? ys[fc * C + o * ratio + j] : 0.0f;
    for (int op = 0; op < dim_out; ++op) {
      const float4* wrow = reinterpret_cast<const float4*>(dw + op * kMaxOut);
      float a0 = __ldg(dec_b + op), a1 = 0.0f;
#pragma unroll
      for (int o4 = 0; o4 < kMaxOut / 4; ++o4) {
        const float4 wv = wrow[o4];
        a0 = fmaf(wv.x, y[o4 * 4 + 0], a0); a1 = fmaf(wv.y, y[o4 * 4 + 1], a1);
        a0 = fmaf(wv.z, y[o4 * 4 + 2], a0); a1 = fmaf(wv.w, y[o4 * 4 + 3], a1);
      }
      const int s = op / K, k = op - s * K;
      dst[((size_t)(2 * f + k / K2) * K2 + (k % K2)) * n_src + s] = a0 + a1;
    }
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
  const size_t smem = (size_t)(2 * sn::kRows * sn::kH + 2 * sn::kSQ * sn::kRows) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->is_first) {
    FNSSL_REQUIRE(a->nf == 256, "sn_freq(first): the first layer runs on 256 bins (got %d)", a->nf);
    FNSSL_REQUIRE(a->enc_wp && a->enc_b && a->enc_kernel == sn::kEncK, "sn_freq(first): encoder weights / kernel size 5 required");
    FNSSL_REQUIRE(a->cin > 0 && a->x_ld >= a->cin && a->x_ld % 4 == 0 && a->x_ld <= 16,
                  "sn_freq(first): cin %d / ld %d (ld must be a multiple of 4, <= 16)", a->cin, a->x_ld);
    FNSSL_REQUIRE((reinterpret_cast<uintptr_t>(a->x) & 15) == 0, "sn_freq(first): x must be 16-byte aligned");
    FNSSL_REQUIRE(a->t_begin >= 0 && a->t_begin < a->nt, "sn_freq(first): t_begin %d outside [0, %d)", a->t_begin, a->nt);
    dim3 grid(a->nt - a->t_begin, a->nb);
#define FNSSL_SN_FIRST(LD)                                                                                              \
  do {                                                                                                                  \
    FNSSL_CUDA(cudaFuncSetAttribute(sn::sn_freq_kernel<LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    sn::sn_freq_kernel<LD><<<grid, sn::kFT, smem, st>>>(*a);                                                            \
  } while (0)
    if (a->x_ld == 4) FNSSL_SN_FIRST(4);
    else if (a->x_ld == 8) FNSSL_SN_FIRST(8);
    else if (a->x_ld == 12) FNSSL_SN_FIRST(12);
    else FNSSL_SN_FIRST(16);
#undef FNSSL_SN_FIRST
  } else {
    FNSSL_REQUIRE(a->nf == 16, "sn_freq: later layers run on 16 bands (got %d)", a->nf);
    FNSSL_CUDA(cudaFuncSetAttribute(sn::sn_freq_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tpb = sn::kRows / a->nf;
    dim3 grid((a->nt + tpb - 1) / tpb, a->nb);
    sn::sn_freq_kernel<0><<<grid, sn::kFT, smem, st>>>(*a);
  }
  FNSSL_LAUNCH_CHECK("sn_freq_kernel");
  return 0;
}

int fnssl_sn_time_forward(const fnssl_sn_time_args* a, void* stream) {
  FNSSL_REQUIRE(a && a->x && a->out && a->work, "sn_time: null argument (x, out and work are required)");
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
  FNSSL_REQUIRE((a->state_flags & ~3) == 0 && (!a->state_flags || (a->state[0] && a->state[1])),
                "sn_time: bad state_flags %d / null state", a->state_flags);
  sn::sn_time_kernel<<<grid, sn::kDI, 0, (cudaStream_t)stream>>>(a->x, a->work, a->nt, a->nf, 1, a->m[0], a->state[0], a->state_flags);
  FNSSL_LAUNCH_CHECK("sn_time_kernel");
  sn::sn_time_kernel<<<grid, sn::kDI, 0, (cudaStream_t)stream>>>(a->work, a->out, a->nt, a->nf, a->pool, a->m[1], a->state[1],
                                                              a->state_flags);
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
  FNSSL_REQUIRE(dim_out <= sn::kMaxOut, "sn_head: dim_output %d > %d", dim_out, sn::kMaxOut);
  const size_t smem = ((size_t)nfc * sn::kH + sn::kMaxOut * sn::kMaxOut + (size_t)nfc * ratio * dim_out) * sizeof(float);
  FNSSL_REQUIRE(smem <= 200 * 1024, "sn_head: too many bands (%d)", nfc);
  FNSSL_CUDA(cudaFuncSetAttribute(sn::sn_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(nt, nb);
  sn::sn_head_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, nt, nfc, trans_wt, trans_b, dec_w, dec_b, dim_out, ratio, n_src, out);
  FNSSL_LAUNCH_CHECK("sn_head_kernel");
  return 0;
}

}  // extern "C"
