// Training side of the LSTM layer: the forward that keeps what back-propagation needs, and the
// backward pass (back-propagation through time) of one uni- / bi-directional layer over a grid -- what autograd does behind
// nn.LSTM at FN-SSL/Lightning/Model.py:38,46 (training_step, FN-SSL/Lightning/main.py:95-109) and IPDnet/FixedAarryIPDnet.py:32,36.
// fp32 CUDA cores (the exact engine, lstm_simt.cu): gradients are held to 1e-4 of torch's autograd on the oracle.
//
//   fnssl_lstm_forward_train : lstm_simt_kernel with two extra outputs per (direction, row, position, unit): the ACTIVATED gates
//                              (i, f, g, o) as one float4 and the cell state c_t.
//   fnssl_lstm_backward      : three kernels
//     1. lstm_bwd_seq_kernel  the sequential part, one CTA per 16..64 sequences and direction, walking the steps backwards:
//                             dh_t = dout_t + W_hh^T dG_{t+1};  dG_t (gradient w.r.t. the gate pre-activations) from the saved
//                             gates / cells, written over the saved gates; the recurrent product reads dG_t of the CTA's rows
//                             from shared memory and W_hh (transposed copy, coalesced float4 per unit pair) from L2.
//     2. lstm_bwd_dx_kernel   dx = dG . W_ih for every position at once (a [positions x dirs*4H] x [dirs*4H x I] product)
//                             scattered to the gradient grids of the two input sources.
//     3. lstm_bwd_dw_kernel   dW_ih | dW_hh | db = [x | h_{t-1} | 1]^T . dG  reduced over all positions (split over CTAs,
//                             fp32 atomics), in the packed layout of the forward weights.
//     Both products are 128 x 128 x 16 register-tiled fp32 kernels (8 x 8 results per thread) whose operands are read in place:
//     the saved buffer is laid out by grid position, so [x | h_prev] rows and dG rows of a position are plain strided reads.
#include <stdlib.h>

#include "common.cuh"

namespace fnssl {

int lstm_forward_simt_save(const fnssl_lstm_args* a, float* gates, float* cells, cudaStream_t st);   // lstm_simt.cu

namespace {

constexpr int kThreads = 256;

// sequence addressing of a (nb, nt, nf, C) grid, as in lstm_simt.cu
struct SeqGeom {
  int64_t rows; int steps; int nf; int nt; int axis;
  __device__ __forceinline__ int64_t base(int64_t row) const {
    return axis == FNSSL_ALONG_FREQ ? row * nf : (row / nf) * (int64_t)nt * nf + (row % nf);
  }
  __device__ __forceinline__ int64_t stride() const { return axis == FNSSL_ALONG_FREQ ? 1 : nf; }
};

// ---- 1. sequential pass ---------------------------------------------------------------------------------------------------
// `saved` is indexed by GRID POSITION: element (dir, pos, unit) with pos = (b*nt + t)*nf + f, so that the two products below
// walk positions linearly (no per-element division back to (sequence, step)).
struct BwdSeqParams {
  SeqGeom g;
  const float* dout; int dout_ld;   // grid, channels [dir*H, dir*H + H)
  float4* gates;                    // in: activated (i,f,g,o);  out: gradient w.r.t. the gate pre-activations
  const float* cells;               // c_t
  const float4* whh_t;              // [dirs][H (unit j)][H (input k)] float4 over the gates: W_hh[gate*H + j][k]
};

template <int H, int RPT>
__global__ void __launch_bounds__(kThreads)
lstm_bwd_seq_kernel(const BwdSeqParams p) {
  constexpr int kRowsPerThread = RPT;          // 16, or 8 / 4 on small batches (grid coverage), as the training forward
  constexpr int G = kThreads / H;
  constexpr int R = kRowsPerThread * G;
  extern __shared__ __align__(16) float4 sm_dg[];   // [R][H]
  __shared__ int64_t s_base[R];
  const int tid = threadIdx.x;
  const int j = tid % H;
  const int rg = tid / H;
  const int dir = blockIdx.y;
  const int64_t row0 = (int64_t)blockIdx.x * R;
  const int steps = p.g.steps;
  const int64_t ss = p.g.stride();
  const int64_t plane = (int64_t)dir * p.g.rows * steps;   // positions of the directions before this one
  for (int lr = tid; lr < R; lr += kThreads) s_base[lr] = (row0 + lr < p.g.rows) ? p.g.base(row0 + lr) : -1;
  __syncthreads();
  const float4* wt = p.whh_t + (size_t)dir * H * H + j;
  float dh_rec[kRowsPerThread], dc_next[kRowsPerThread];
#pragma unroll
  for (int i = 0; i < kRowsPerThread; ++i) { dh_rec[i] = 0.0f; dc_next[i] = 0.0f; }

  for (int step = steps - 1; step >= 0; --step) {       // step = position in the order the forward processed the sequence
    const int s = dir ? (steps - 1 - step) : step;
    const int64_t dprev = dir ? ss : -ss;               // where the forward came from (valid when step > 0)
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const int lr = rg * kRowsPerThread + i;
      const int64_t base = s_base[lr];
      float4 d = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (base >= 0) {
        const int64_t pos = base + (int64_t)s * ss;
        const int64_t idx = (plane + pos) * H + j;
        const float4 a = p.gates[idx];
        const float ct = p.cells[idx];
        const float cp = step > 0 ? p.cells[idx + dprev * H] : 0.0f;
        const float dh = p.dout[pos * p.dout_ld + dir * H + j] + dh_rec[i];
        const float tc = tanh_f(ct);
        const float dc = dc_next[i] + dh * a.w * (1.0f - tc * tc);
        d.x = dc * a.z * a.x * (1.0f - a.x);            // i
        d.y = dc * cp * a.y * (1.0f - a.y);             // f
        d.z = dc * a.x * (1.0f - a.z * a.z);            // g
        d.w = dh * tc * a.w * (1.0f - a.w);             // o
        dc_next[i] = dc * a.y;
        p.gates[idx] = d;
      }
      sm_dg[lr * H + j] = d;
    }
    __syncthreads();
    if (step > 0) {                                     // dh_{t-1} += W_hh^T dG_t
      float acc[kRowsPerThread];
#pragma unroll
      for (int i = 0; i < kRowsPerThread; ++i) acc[i] = 0.0f;
      const float4* drow = sm_dg + (size_t)(rg * kRowsPerThread) * H;
#pragma unroll 2
      for (int jj = 0; jj < H; ++jj) {
        const float4 w = __ldg(wt + (size_t)jj * H);
#pragma unroll
        for (int i = 0; i < kRowsPerThread; ++i) {
          const float4 a = drow[(size_t)i * H + jj];
          acc[i] = fmaf(a.x, w.x, acc[i]); acc[i] = fmaf(a.y, w.y, acc[i]);
          acc[i] = fmaf(a.z, w.z, acc[i]); acc[i] = fmaf(a.w, w.w, acc[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < kRowsPerThread; ++i) dh_rec[i] = acc[i];
    }
    __syncthreads();
  }
}

// ---- 128 x 128 x 16 fp32 tile product: 256 threads, 8 x 8 results per thread (two 4-wide groups per dimension) --------------
constexpr int TM = 128, TN = 128, TK = 16, TLD = TM + 4;     // +4 floats: rows stay 16-byte aligned

struct TileAcc {
  float v[8][8];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int w = 0; w < 8; ++w) v[u][w] = 0.0f;
  }
  // As[kk][m], Bs[kk][n]: one K = 16 slab.  Thread (ty, tx) owns rows {4 ty .. +3, 64 + 4 ty .. +3}, columns likewise with tx.
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
  __device__ __forceinline__ static int row_of(int ty, int u) { return (u < 4 ? 0 : 60) + ty * 4 + u; }   // u >= 4 -> 64 + 4 ty + (u - 4)
};

// ---- 2. input gradient: dx[pos][n] = sum_{dir, cc} dG[dir][pos][cc] W[dir][n][cc] -------------------------------------------
struct BwdDxParams {
  int64_t npos;                      // nb * nt * nf
  int H, dirs, I, c0, Kp;
  int n_begin;                       // first input column that needs a gradient (c0 when dsrc0 == NULL)
  const float* dgates;               // [dirs][npos][4H]  (column = unit*4 + gate)
  const float* w;                    // packed forward weights: [dirs][Kp][4H] in the same column order
  float* dsrc0; int dld0;
  float* dsrc1; int dld1;
};

__global__ void __launch_bounds__(kThreads)
lstm_bwd_dx_kernel(const BwdDxParams p) {
  __shared__ __align__(16) float As[TK][TLD];
  __shared__ __align__(16) float Bs[TK][TLD];
  const int t = threadIdx.x, tx = t % 16, ty = t / 16;
  const int G4 = 4 * p.H;
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int n0 = p.n_begin + blockIdx.y * TN;
  const int lq = (t % 4) * 4;                          // this thread's 4 consecutive k of the slab
  TileAcc acc;
  acc.clear();
  for (int dir = 0; dir < p.dirs; ++dir) {
    const float* dg = p.dgates + (int64_t)dir * p.npos * G4;
    const float* wd = p.w + (int64_t)dir * p.Kp * G4;
    for (int cc0 = 0; cc0 < G4; cc0 += TK) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int mm = t / 4 + 64 * r;
        const int64_t pos = m0 + mm;
        float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (pos < p.npos) a = *reinterpret_cast<const float4*>(dg + pos * G4 + cc0 + lq);
        As[lq + 0][mm] = a.x; As[lq + 1][mm] = a.y; As[lq + 2][mm] = a.z; As[lq + 3][mm] = a.w;
        const int n = n0 + mm;
        float4 b = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (n < p.I) b = __ldg(reinterpret_cast<const float4*>(wd + (int64_t)n * G4 + cc0 + lq));
        Bs[lq + 0][mm] = b.x; Bs[lq + 1][mm] = b.y; Bs[lq + 2][mm] = b.z; Bs[lq + 3][mm] = b.w;
      }
      __syncthreads();
      acc.mac(As, Bs, ty, tx);
      __syncthreads();
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int64_t pos = m0 + TileAcc::row_of(ty, u);
    if (pos >= p.npos) continue;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int n = n0 + TileAcc::row_of(tx, w);
      if (n >= p.I) continue;
      if (n < p.c0) { if (p.dsrc0) p.dsrc0[pos * p.dld0 + n] = acc.v[u][w]; }
      else if (p.dsrc1) p.dsrc1[pos * p.dld1 + (n - p.c0)] = acc.v[u][w];
    }
  }
}

// ---- 3. weight gradient: dW[dir][k][cc] = sum_pos [x | h_prev][pos][k] dG[dir][pos][cc];  db[dir][cc] = sum_pos dG[dir][pos][cc] -----------------------------------
struct BwdDwParams {
  SeqGeom g;
  int64_t npos;
  int H, dirs, I, c0, Kp;
  const float* src0; int ld0;
  const float* src1; int ld1;
  const float* hout; int hld; int hoff;   // the forward's h grid (h_{t-1} operand of W_hh)
  const float* dgates;
  float* dw;                              // packed: [dirs][Kp][4H] weights ++ [dirs][4H] bias, zero on entry
  int64_t chunk;                          // positions per CTA along the reduction (multiple of TK)
};

__global__ void __launch_bounds__(kThreads)
lstm_bwd_dw_kernel(const BwdDwParams p) {
  __shared__ __align__(16) float As[TK][TLD];
  __shared__ __align__(16) float Bs[TK][TLD];
  const int t = threadIdx.x, tx = t % 16, ty = t / 16;
  const int G4 = 4 * p.H;
  const int K = p.I + p.H;
  const int dir = blockIdx.z % p.dirs;
  const int64_t split = blockIdx.z / p.dirs;
  const int64_t mb = split * p.chunk;
  const int64_t me = (mb + p.chunk < p.npos) ? mb + p.chunk : p.npos;
  const int k0 = blockIdx.x * TM;          // rows of the result: columns of [x | h_prev | 1]
  const int n0 = blockIdx.y * TN;          // gate columns
  const int64_t ss = p.g.stride();
  const float* dg = p.dgates + (int64_t)dir * p.npos * G4;
  const int pk = t / 16;                   // this thread's position within a slab of 16 (operand A: 8 columns k of ONE position)
  const int kq = t % 16;
  TileAcc acc;
  acc.clear();
  const bool bias_cta = blockIdx.x == 0 && ty == 0;   // d bias = column sums of dG: the first row of tiles adds them up on the side
  float bsum[8];
#pragma unroll
  for (int w = 0; w < 8; ++w) bsum[w] = 0.0f;
  for (int64_t m0 = mb; m0 < me; m0 += TK) {
    {   // A slab: As[pk][k - k0] = [x | h_prev](pos, k)
      const int64_t pos = m0 + pk;
      const bool valid = pos < me;
      bool first = true;                   // is `pos` the first step of its sequence in this direction (h_prev = 0)?
      if (valid) {
        const int s = p.g.axis == FNSSL_ALONG_FREQ ? (int)(pos % p.g.nf) : (int)((pos / p.g.nf) % p.g.nt);
        first = dir ? (s == p.g.steps - 1) : (s == 0);
      }
      const float* hp = p.hout + (pos + (dir ? ss : -ss)) * p.hld + p.hoff + dir * p.H;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int kk = kq + 16 * r;
        const int k = k0 + kk;
        float v = 0.0f;
        if (valid) {
          if (k < p.c0) v = p.src0[pos * p.ld0 + k];
          else if (k < p.I) v = p.src1[pos * p.ld1 + (k - p.c0)];
          else if (k < K) v = first ? 0.0f : hp[k - p.I];
        }
        As[pk][kk] = v;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {          // B slab: Bs[pk][cc - n0] = dG(pos, cc), float4 per thread
      const int64_t pos = m0 + pk;
      const int nn = kq * 4 + 64 * r;
      float4 b = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (pos < me) b = *reinterpret_cast<const float4*>(dg + pos * G4 + n0 + nn);
      *reinterpret_cast<float4*>(&Bs[pk][nn]) = b;
    }
    __syncthreads();
    acc.mac(As, Bs, ty, tx);
    if (bias_cta) {
#pragma unroll
      for (int kk = 0; kk < TK; ++kk)
#pragma unroll
        for (int w = 0; w < 8; ++w) bsum[w] += Bs[kk][TileAcc::row_of(tx, w)];
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int k = k0 + TileAcc::row_of(ty, u);
    if (k >= K) continue;
    float* dst = p.dw + ((int64_t)dir * p.Kp + k) * G4;
#pragma unroll
    for (int w = 0; w < 8; ++w) atomicAdd(dst + n0 + TileAcc::row_of(tx, w), acc.v[u][w]);
  }
  if (bias_cta) {
    float* db = p.dw + (size_t)p.dirs * p.Kp * G4 + (int64_t)dir * G4;
#pragma unroll
    for (int w = 0; w < 8; ++w) atomicAdd(db + n0 + TileAcc::row_of(tx, w), bsum[w]);
  }
}

// Second version of the weight-gradient product for layers whose channel counts and strides are multiples of 4 (every layer of
// the networks; the first version above stays for odd shapes): the [x | h_prev] operand is gathered as float4 -- a group of 4
// columns never straddles the src0 / src1 / h boundaries --, the position -> (step-of-sequence) bookkeeping is incremental
// instead of a 64-bit division per slab, and the global loads of slab i+1 are issued before the FMAs of slab i (register
// double buffering).
__global__ void __launch_bounds__(kThreads, 2)
lstm_bwd_dw2_kernel(const BwdDwParams p) {
  __shared__ __align__(16) float As[TK][TLD];
  __shared__ __align__(16) float Bs[TK][TLD];
  const int t = threadIdx.x, tx = t % 16, ty = t / 16;
  const int G4 = 4 * p.H;
  const int K = p.I + p.H;
  const int dir = blockIdx.z % p.dirs;
  const int64_t split = blockIdx.z / p.dirs;
  const int64_t mb = split * p.chunk;
  const int64_t me = (mb + p.chunk < p.npos) ? mb + p.chunk : p.npos;
  const int k0 = blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const int64_t ss = p.g.stride();
  const float* dg = p.dgates + (int64_t)dir * p.npos * G4;
  const int pk = t / 16;                   // this thread's position within a slab of 16
  const int kq = t % 16;                   // its float4 column groups: 4 (kq + 16 r), r = 0, 1
  // step-of-sequence bookkeeping of position mb + pk, advanced by 16 positions per slab
  int f_idx, s_idx;                        // ALONG_FREQ: s_idx = f;  ALONG_TIME: (f_idx, s_idx = t)
  {
    const int64_t pos = mb + pk;
    f_idx = (int)(pos % p.g.nf);
    s_idx = p.g.axis == FNSSL_ALONG_FREQ ? f_idx : (int)((pos / p.g.nf) % p.g.nt);
  }
  const int s_first = dir ? p.g.steps - 1 : 0;
  const int64_t dprev = dir ? ss : -ss;

  float4 ra[2], rb[2];
  auto fetch = [&](int64_t m0) {
    const int64_t pos = m0 + pk;
    const bool valid = pos < me;
    const bool first = s_idx == s_first;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int k = k0 + 4 * (kq + 16 * r);
      float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (valid) {
        if (k < p.c0) v = *reinterpret_cast<const float4*>(p.src0 + pos * p.ld0 + k);
        else if (k < p.I) v = *reinterpret_cast<const float4*>(p.src1 + pos * p.ld1 + (k - p.c0));
        else if (k < K && !first) v = *reinterpret_cast<const float4*>(p.hout + (pos + dprev) * p.hld + p.hoff + dir * p.H + (k - p.I));
      }
      ra[r] = v;
      float4 b = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (valid) b = *reinterpret_cast<const float4*>(dg + pos * G4 + n0 + kq * 4 + 64 * r);
      rb[r] = b;
    }
    // advance the bookkeeping to the next slab
    f_idx += TK;
    if (p.g.axis == FNSSL_ALONG_FREQ) {
      while (f_idx >= p.g.nf) f_idx -= p.g.nf;
      s_idx = f_idx;
    } else {
      while (f_idx >= p.g.nf) { f_idx -= p.g.nf; if (++s_idx == p.g.nt) s_idx = 0; }
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      *reinterpret_cast<float4*>(&As[pk][4 * (kq + 16 * r)]) = ra[r];
      *reinterpret_cast<float4*>(&Bs[pk][kq * 4 + 64 * r]) = rb[r];
    }
  };

  TileAcc acc;
  acc.clear();
  const bool bias_cta = blockIdx.x == 0 && ty == 0;
  float bsum[8];
#pragma unroll
  for (int w = 0; w < 8; ++w) bsum[w] = 0.0f;
  if (mb < me) {
    fetch(mb);
    stash();
    __syncthreads();
    for (int64_t m0 = mb; m0 < me; m0 += TK) {
      const bool more = m0 + TK < me;
      if (more) fetch(m0 + TK);
      acc.mac(As, Bs, ty, tx);
      if (bias_cta) {
#pragma unroll
        for (int kk = 0; kk < TK; ++kk)
#pragma unroll
          for (int w = 0; w < 8; ++w) bsum[w] += Bs[kk][TileAcc::row_of(tx, w)];
      }
      __syncthreads();
      if (more) stash();
      __syncthreads();
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int k = k0 + TileAcc::row_of(ty, u);
    if (k >= K) continue;
    float* dst = p.dw + ((int64_t)dir * p.Kp + k) * G4;
#pragma unroll
    for (int w = 0; w < 8; ++w) atomicAdd(dst + n0 + TileAcc::row_of(tx, w), acc.v[u][w]);
  }
  if (bias_cta) {
    float* db = p.dw + (size_t)p.dirs * p.Kp * G4 + (int64_t)dir * G4;
#pragma unroll
    for (int w = 0; w < 8; ++w) atomicAdd(db + n0 + TileAcc::row_of(tx, w), bsum[w]);
  }
}

template <int H, int RPT>
int launch_bwd_seq_rpt(const BwdSeqParams& p, int dirs, cudaStream_t st) {
  constexpr int R = RPT * (kThreads / H);
  const size_t smem = (size_t)R * H * sizeof(float4);
  FNSSL_CUDA(cudaFuncSetAttribute(lstm_bwd_seq_kernel<H, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div64(p.g.rows, R), dirs);
  lstm_bwd_seq_kernel<H, RPT><<<grid, kThreads, smem, st>>>(p);
  FNSSL_LAUNCH_CHECK("lstm_bwd_seq_kernel");
  return 0;
}

template <int H>
int launch_bwd_seq(const BwdSeqParams& p, int dirs, cudaStream_t st) {
  constexpr int G = kThreads / H;
  int rpt = ceil_div64(p.g.rows, 16 * G) * dirs >= 148 ? 16 : ceil_div64(p.g.rows, 8 * G) * dirs >= 148 ? 8 : 4;
  if (const char* e = getenv("FNSSL_TRAIN_RPT")) rpt = atoi(e);          // tests: force a variant on a small grid
  if (rpt == 16) return launch_bwd_seq_rpt<H, 16>(p, dirs, st);
  if (rpt == 8) return launch_bwd_seq_rpt<H, 8>(p, dirs, st);
  return launch_bwd_seq_rpt<H, 4>(p, dirs, st);
}

int check_train_args(const fnssl_lstm_args* a, const char* who) {
  FNSSL_REQUIRE(a != nullptr, "%s: null args", who);
  FNSSL_REQUIRE(a->engine == FNSSL_ENGINE_SIMT && a->dtype == FNSSL_F32, "%s: the training path runs the fp32 engine on fp32 grids", who);
  FNSSL_REQUIRE(a->axis == FNSSL_ALONG_FREQ || a->axis == FNSSL_ALONG_TIME, "%s: bad axis %d", who, a->axis);
  FNSSL_REQUIRE(a->nb > 0 && a->nt > 0 && a->nf > 0, "%s: bad grid %d x %d x %d", who, a->nb, a->nt, a->nf);
  FNSSL_REQUIRE(a->num_dirs == 1 || a->num_dirs == 2, "%s: num_dirs must be 1 or 2 (got %d)", who, a->num_dirs);
  FNSSL_REQUIRE(a->hidden == 32 || a->hidden == 64 || a->hidden == 128 || a->hidden == 256, "%s: hidden size %d not supported (32, 64, 128, 256)",
                who, a->hidden);
  FNSSL_REQUIRE(a->src0 && a->c0 > 0 && a->ld0 >= a->c0, "%s: bad src0 (c0=%d ld0=%d)", who, a->c0, a->ld0);
  FNSSL_REQUIRE(a->c1 >= 0 && (a->c1 == 0 || (a->src1 && a->ld1 >= a->c1)), "%s: bad src1 (c1=%d ld1=%d)", who, a->c1, a->ld1);
  FNSSL_REQUIRE(a->weights && a->out0, "%s: null weights / h grid", who);
  FNSSL_REQUIRE(a->out0_off >= 0 && a->out0_ld >= a->out0_off + a->num_dirs * a->hidden, "%s: out0 window exceeds its stride", who);
  FNSSL_REQUIRE(a->state_flags == 0, "%s: a carried recurrent state is an inference (streaming) feature", who);
  return 0;
}

int64_t positions_of(const fnssl_lstm_args* a) { return (int64_t)a->nb * a->nt * a->nf; }   // rows * steps on either axis

}  // namespace
}  // namespace fnssl

using namespace fnssl;

extern "C" {

int64_t fnssl_lstm_train_saved_bytes(int nb, int nt, int nf, int hidden, int num_dirs) {
  if (nb <= 0 || nt <= 0 || nf <= 0 || hidden <= 0 || num_dirs <= 0) return 0;
  return (int64_t)num_dirs * nb * nt * nf * hidden * 5 * (int64_t)sizeof(float);   // float4 gates + float cell per unit
}

int fnssl_lstm_forward_train(const fnssl_lstm_args* a, void* saved, int64_t saved_bytes, void* stream) {
  if (int rc = check_train_args(a, "lstm_forward_train")) return rc;
  const int64_t units = (int64_t)a->num_dirs * positions_of(a) * a->hidden;
  FNSSL_REQUIRE(saved && saved_bytes == units * 5 * (int64_t)sizeof(float) && (reinterpret_cast<uintptr_t>(saved) & 15) == 0,
                "lstm_forward_train: `saved` must be a 16-byte aligned buffer of fnssl_lstm_train_saved_bytes() = %lld bytes (got %lld)",
                (long long)(units * 5 * (int64_t)sizeof(float)), (long long)saved_bytes);
  float* gates = reinterpret_cast<float*>(saved);
  return lstm_forward_simt_save(a, gates, gates + units * 4, (cudaStream_t)stream);
}

int fnssl_lstm_backward(const fnssl_lstm_args* a, void* saved, int64_t saved_bytes, const float* whh_t, const float* dout, int dout_ld,
                        float* dsrc0, int dsrc0_ld, float* dsrc1, int dsrc1_ld, float* dweights, void* stream) {
  if (int rc = check_train_args(a, "lstm_backward")) return rc;
  const int H = a->hidden, dirs = a->num_dirs;
  const int64_t units = (int64_t)dirs * positions_of(a) * H;
  FNSSL_REQUIRE(saved && saved_bytes == units * 5 * (int64_t)sizeof(float) && (reinterpret_cast<uintptr_t>(saved) & 15) == 0,
                "lstm_backward: `saved` is not the buffer fnssl_lstm_forward_train filled for this layer");
  FNSSL_REQUIRE(whh_t && (reinterpret_cast<uintptr_t>(whh_t) & 15) == 0, "lstm_backward: null / unaligned whh_t");
  FNSSL_REQUIRE(dout && dout_ld >= dirs * H, "lstm_backward: bad dout (ld %d)", dout_ld);
  FNSSL_REQUIRE(!dsrc0 || dsrc0_ld >= a->c0, "lstm_backward: bad dsrc0 stride %d", dsrc0_ld);
  FNSSL_REQUIRE(!dsrc1 || (a->c1 > 0 && dsrc1_ld >= a->c1), "lstm_backward: bad dsrc1 (c1 %d, stride %d)", a->c1, dsrc1_ld);
  FNSSL_REQUIRE(dweights != nullptr, "lstm_backward: null dweights");
  const int I = a->c0 + a->c1, K = I + H, Kp = (K + 3) & ~3;
  const int64_t need = ((int64_t)dirs * Kp * H + (int64_t)dirs * H) * 16;
  FNSSL_REQUIRE(a->weights_bytes == need, "lstm_backward: packed weight buffer is %lld bytes, expected %lld", (long long)a->weights_bytes,
                (long long)need);
  cudaStream_t st = (cudaStream_t)stream;
  SeqGeom g;
  g.nf = a->nf; g.nt = a->nt; g.axis = a->axis;
  if (a->axis == FNSSL_ALONG_FREQ) { g.rows = (int64_t)a->nb * a->nt; g.steps = a->nf; }
  else { g.rows = (int64_t)a->nb * a->nf; g.steps = a->nt; }
  float* gates = reinterpret_cast<float*>(saved);
  const float* cells = gates + units * 4;

  BwdSeqParams sp;
  sp.g = g; sp.dout = dout; sp.dout_ld = dout_ld; sp.gates = reinterpret_cast<float4*>(gates); sp.cells = cells;
  sp.whh_t = reinterpret_cast<const float4*>(whh_t);
  int rc = 0;
  switch (H) {
    case 32: rc = launch_bwd_seq<32>(sp, dirs, st); break;
    case 64: rc = launch_bwd_seq<64>(sp, dirs, st); break;
    case 128: rc = launch_bwd_seq<128>(sp, dirs, st); break;
    default: rc = launch_bwd_seq<256>(sp, dirs, st); break;
  }
  if (rc) return rc;

  const int64_t M = g.rows * g.steps;                       // = nb * nt * nf grid positions
  if (dsrc0 || dsrc1) {
    BwdDxParams xp;
    xp.npos = M; xp.H = H; xp.dirs = dirs; xp.I = I; xp.c0 = a->c0; xp.Kp = Kp;
    xp.n_begin = dsrc0 ? 0 : a->c0;
    const int n_end = dsrc1 ? I : a->c0;                    // columns outside [n_begin, n_end) need no gradient
    xp.dgates = gates; xp.w = reinterpret_cast<const float*>(a->weights);
    xp.dsrc0 = dsrc0; xp.dld0 = dsrc0_ld; xp.dsrc1 = dsrc1; xp.dld1 = dsrc1_ld;
    dim3 grid((unsigned)ceil_div64(M, TM), (unsigned)((n_end - xp.n_begin + TN - 1) / TN));
    lstm_bwd_dx_kernel<<<grid, kThreads, 0, st>>>(xp);
    FNSSL_LAUNCH_CHECK("lstm_bwd_dx_kernel");
  }

  FNSSL_CUDA(cudaMemsetAsync(dweights, 0, (size_t)need, st));
  BwdDwParams wp;
  wp.g = g; wp.npos = M; wp.H = H; wp.dirs = dirs; wp.I = I; wp.c0 = a->c0; wp.Kp = Kp;
  wp.src0 = reinterpret_cast<const float*>(a->src0); wp.ld0 = a->ld0;
  wp.src1 = reinterpret_cast<const float*>(a->src1); wp.ld1 = a->ld1;
  wp.hout = reinterpret_cast<const float*>(a->out0); wp.hld = a->out0_ld; wp.hoff = a->out0_off;
  wp.dgates = gates; wp.dw = dweights;
  const int64_t tiles = (int64_t)((K + TM - 1) / TM) * (4 * H / TN) * dirs;
  int64_t want = (148 * 4 + tiles - 1) / tiles;             // splits of the reduction: about four CTAs per SM in total
  if (want < 1) want = 1;
  int64_t chunk = ceil_div64(M, want);
  if (chunk < 1024) chunk = 1024;
  chunk = ceil_div64(chunk, TK) * TK;
  wp.chunk = chunk;
  const int64_t splits = ceil_div64(M, chunk);
  dim3 wgrid((unsigned)((K + TM - 1) / TM), (unsigned)(4 * H / TN), (unsigned)(dirs * splits));
  // float4 operand gathers need every channel count / stride / pointer to be a multiple of 4 floats
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool vec_ok = a->c0 % 4 == 0 && a->c1 % 4 == 0 && a->ld0 % 4 == 0 && (a->c1 == 0 || a->ld1 % 4 == 0) && a->out0_ld % 4 == 0 &&
                      a->out0_off % 4 == 0 && al16(a->src0) && (a->c1 == 0 || al16(a->src1)) && al16(a->out0);
  const char* dw_env = getenv("FNSSL_TRAIN_DW");            // 2 (default): lstm_bwd_dw2_kernel when vec_ok; 1: lstm_bwd_dw_kernel always
  const int dw_version = dw_env ? atoi(dw_env) : 2;         // (B = 16: 514 -> 416 ms per training step, profiles/r2_train_bench_v55.jsonl)
  if (vec_ok && dw_version == 2) {
    lstm_bwd_dw2_kernel<<<wgrid, kThreads, 0, st>>>(wp);
    FNSSL_LAUNCH_CHECK("lstm_bwd_dw2_kernel");
  } else {
    lstm_bwd_dw_kernel<<<wgrid, kThreads, 0, st>>>(wp);
    FNSSL_LAUNCH_CHECK("lstm_bwd_dw_kernel");
  }
  return 0;
}

}  // extern "C"
