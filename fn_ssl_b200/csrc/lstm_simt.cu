// fp32 CUDA-core LSTM engine (FNSSL_ENGINE_SIMT): exact-precision, shape-generic persistent kernel.
//
// Replaces nn.LSTM as used at FN-SSL/Lightning/Model.py:38,46 and IPDnet/FixedAarryIPDnet.py:32,36
// (gate order i,f,g,o; b_ih + b_hh; zero initial state; direction 1 = reversed sequence) together with
// the reshape/permute/cat/add glue around it (Model.py:35-37,41-45,49).
//
// One CTA owns R = 16*(256/H) sequences ("rows") of one direction for ALL steps: the cell state lives in
// registers, [x_t | h_{t-1}] in shared memory, and the (I+H) x 4H weight matrix is streamed from L2 every
// step as float4 (i,f,g,o) per (k, unit).  It is the reference engine of the package: any I, H in
// {32,64,128,256}, fp32 or fp16 grids.  The tensor-core engine (lstm_tc.cu) is the fast one.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace fnssl {

constexpr int kSimtThreads = 256;

struct SimtParams {
  const void* src0; int c0; int ld0;
  const void* src1; int c1; int ld1;
  const float4* w4;     // [dirs][Kp][H]
  const float4* bias4;  // [dirs][H]
  void* out0; int out0_ld; int out0_off;
  const void* addend; int addend_ld;
  void* out1; int out1_ld;
  int64_t rows; int steps; int nf; int nt; int axis;
  int I, K, Kp;
  float* h_state; float* c_state; int state_flags;   // optional carried state, fp32 (rows, H)
  float4* save_gates; float* save_cells;             // training forward (lstm_train.cu): activated (i,f,g,o) and c_t per
};                                                   // (dir, grid position, unit), NULL for inference

// RPT = rows per thread: 16 for inference (the weights streamed from L2 are reused by 16 rows); the training forward picks 8 / 4
// on small batches so that the grid still covers the GPU (its CTAs are latency-bound otherwise).
template <typename T, int H, int RPT = 16>
__global__ void __launch_bounds__(kSimtThreads)
lstm_simt_kernel(const SimtParams p) {
  constexpr int kRowsPerThread = RPT;
  constexpr int G = kSimtThreads / H;          // row groups
  constexpr int R = kRowsPerThread * G;        // rows per CTA
  extern __shared__ __align__(16) float smem_a[];  // [R][Kp]
  __shared__ int64_t s_base[R];

  const int tid = threadIdx.x;
  const int j = tid % H;
  const int rg = tid / H;
  const int dir = blockIdx.y;
  const int64_t row0 = (int64_t)blockIdx.x * R;
  const int Kp = p.Kp, I = p.I;
  const int64_t sstride = (p.axis == FNSSL_ALONG_FREQ) ? 1 : p.nf;

  for (int lr = tid; lr < R; lr += kSimtThreads) {
    const int64_t row = row0 + lr;
    int64_t base = -1;
    if (row < p.rows) {
      if (p.axis == FNSSL_ALONG_FREQ) base = row * p.nf;
      else base = (row / p.nf) * (int64_t)p.nt * p.nf + (row % p.nf);
    }
    s_base[lr] = base;
  }
  for (int idx = tid; idx < R * Kp; idx += kSimtThreads) smem_a[idx] = 0.0f;  // h_{-1} = 0, K padding = 0
  __syncthreads();

  const float4* w4 = p.w4 + (size_t)dir * Kp * H + j;
  const float4 bias = p.bias4[dir * H + j];
  const T* src0 = reinterpret_cast<const T*>(p.src0);
  const T* src1 = reinterpret_cast<const T*>(p.src1);

  float c[kRowsPerThread];
#pragma unroll
  for (int i = 0; i < kRowsPerThread; ++i) c[i] = 0.0f;
  if (p.state_flags & 1) {   // resume from a carried (h, c)
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const int lr = rg * kRowsPerThread + i;
      if (row0 + lr < p.rows) {
        c[i] = p.c_state[(row0 + lr) * H + j];
        smem_a[lr * Kp + I + j] = p.h_state[(row0 + lr) * H + j];
      }
    }
    __syncthreads();
  }

  for (int step = 0; step < p.steps; ++step) {
    const int s = dir ? (p.steps - 1 - step) : step;
    // 1. x_t -> shared
    for (int idx = tid; idx < R * I; idx += kSimtThreads) {
      const int lr = idx / I, k = idx - lr * I;
      const int64_t base = s_base[lr];
      float v = 0.0f;
      if (base >= 0) {
        const int64_t pos = base + (int64_t)s * sstride;
        v = (k < p.c0) ? ld_act<T>(src0 + pos * p.ld0 + k) : ld_act<T>(src1 + pos * p.ld1 + (k - p.c0));
      }
      smem_a[lr * Kp + k] = v;
    }
    __syncthreads();
    // 2. gates = [x_t | h_{t-1}] @ W + b
    float acc[kRowsPerThread][4];
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) { acc[i][0] = bias.x; acc[i][1] = bias.y; acc[i][2] = bias.z; acc[i][3] = bias.w; }
    const float* arow = smem_a + (size_t)(rg * kRowsPerThread) * Kp;
#pragma unroll 1
    for (int k = 0; k < Kp; k += 4) {
      const float4 w0 = __ldg(w4 + (size_t)(k + 0) * H);
      const float4 w1 = __ldg(w4 + (size_t)(k + 1) * H);
      const float4 w2 = __ldg(w4 + (size_t)(k + 2) * H);
      const float4 w3 = __ldg(w4 + (size_t)(k + 3) * H);
#pragma unroll
      for (int i = 0; i < kRowsPerThread; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(arow + (size_t)i * Kp + k);
        acc[i][0] = fmaf(a.x, w0.x, acc[i][0]); acc[i][1] = fmaf(a.x, w0.y, acc[i][1]);
        acc[i][2] = fmaf(a.x, w0.z, acc[i][2]); acc[i][3] = fmaf(a.x, w0.w, acc[i][3]);
        acc[i][0] = fmaf(a.y, w1.x, acc[i][0]); acc[i][1] = fmaf(a.y, w1.y, acc[i][1]);
        acc[i][2] = fmaf(a.y, w1.z, acc[i][2]); acc[i][3] = fmaf(a.y, w1.w, acc[i][3]);
        acc[i][0] = fmaf(a.z, w2.x, acc[i][0]); acc[i][1] = fmaf(a.z, w2.y, acc[i][1]);
        acc[i][2] = fmaf(a.z, w2.z, acc[i][2]); acc[i][3] = fmaf(a.z, w2.w, acc[i][3]);
        acc[i][0] = fmaf(a.w, w3.x, acc[i][0]); acc[i][1] = fmaf(a.w, w3.y, acc[i][1]);
        acc[i][2] = fmaf(a.w, w3.z, acc[i][2]); acc[i][3] = fmaf(a.w, w3.w, acc[i][3]);
      }
    }
    __syncthreads();  // every thread is done reading h_{t-1}
    // 3. cell update, h_t -> shared (next step's operand) and global
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const int lr = rg * kRowsPerThread + i;
      const float ig = sigmoid_f(acc[i][0]);
      const float fg = sigmoid_f(acc[i][1]);
      const float gg = tanh_f(acc[i][2]);
      const float og = sigmoid_f(acc[i][3]);
      c[i] = fg * c[i] + ig * gg;
      const float h = og * tanh_f(c[i]);
      smem_a[lr * Kp + I + j] = h;
      const int64_t base = s_base[lr];
      if (base >= 0) {
        const int64_t pos = base + (int64_t)s * sstride;
        if (p.save_gates) {      // grid-position order: [dir][(b, t, f)][unit]
          const int64_t sidx = ((int64_t)dir * p.rows * p.steps + pos) * H + j;
          p.save_gates[sidx] = make_float4(ig, fg, gg, og);
          p.save_cells[sidx] = c[i];
        }
        const int ch = dir * H + j;
        if (p.out0) st_act<T>(reinterpret_cast<T*>(p.out0) + pos * p.out0_ld + p.out0_off + ch, h);
        if (p.out1) {
          const float a = p.addend ? ld_act<T>(reinterpret_cast<const T*>(p.addend) + pos * p.addend_ld + ch) : 0.0f;
          st_act<T>(reinterpret_cast<T*>(p.out1) + pos * p.out1_ld + ch, h + a);
        }
      }
    }
    // (the x_t load of the next step touches columns [0, I) only; the barrier after it also orders these h writes)
  }
  if (p.state_flags & 2) {
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const int lr = rg * kRowsPerThread + i;
      if (row0 + lr < p.rows) {
        p.c_state[(row0 + lr) * H + j] = c[i];
        p.h_state[(row0 + lr) * H + j] = smem_a[lr * Kp + I + j];   // written by this thread in the last step
      }
    }
  }
}

template <typename T, int H, int RPT>
static int launch_simt_rpt(const SimtParams& p, int dirs, cudaStream_t st) {
  constexpr int R = RPT * (kSimtThreads / H);
  const size_t smem = (size_t)R * p.Kp * sizeof(float);
  FNSSL_REQUIRE(smem <= 220 * 1024, "lstm(simt): input size %d too large for H=%d", p.I, H);
  FNSSL_CUDA(cudaFuncSetAttribute(lstm_simt_kernel<T, H, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div64(p.rows, R), dirs);
  lstm_simt_kernel<T, H, RPT><<<grid, kSimtThreads, smem, st>>>(p);
  FNSSL_LAUNCH_CHECK("lstm_simt_kernel");
  return 0;
}

template <typename T, int H>
static int launch_simt(const SimtParams& p, int dirs, cudaStream_t st) {
  if (p.save_gates) {   // training forward (fp32 grids): fewer rows per CTA while the 16-row grid would leave SMs idle
    if constexpr (std::is_same<T, float>::value) {
      constexpr int G = kSimtThreads / H;
      int rpt = ceil_div64(p.rows, 16 * G) * dirs >= 148 ? 16 : ceil_div64(p.rows, 8 * G) * dirs >= 148 ? 8 : 4;
      if (const char* e = getenv("FNSSL_TRAIN_RPT")) rpt = atoi(e);      // tests: force a variant on a small grid
      if (rpt == 8) return launch_simt_rpt<T, H, 8>(p, dirs, st);
      if (rpt == 4) return launch_simt_rpt<T, H, 4>(p, dirs, st);
    }
  }
  return launch_simt_rpt<T, H, 16>(p, dirs, st);
}

// gates / cells: NULL (inference) or the buffers of the training forward (fnssl_lstm_forward_train)
int lstm_forward_simt_save(const fnssl_lstm_args* a, float* gates, float* cells, cudaStream_t st) {
  SimtParams p;
  p.save_gates = reinterpret_cast<float4*>(gates); p.save_cells = cells;
  p.src0 = a->src0; p.c0 = a->c0; p.ld0 = a->ld0;
  p.src1 = a->src1; p.c1 = a->c1; p.ld1 = a->ld1;
  p.I = a->c0 + a->c1;
  p.K = p.I + a->hidden;
  p.Kp = (p.K + 3) & ~3;
  const int64_t need = ((int64_t)a->num_dirs * p.Kp * a->hidden + (int64_t)a->num_dirs * a->hidden) * 16;
  FNSSL_REQUIRE(a->weights_bytes == need, "lstm(simt): packed weight buffer is %lld bytes, expected %lld",
                (long long)a->weights_bytes, (long long)need);
  p.w4 = reinterpret_cast<const float4*>(a->weights);
  p.bias4 = p.w4 + (size_t)a->num_dirs * p.Kp * a->hidden;
  p.out0 = a->out0; p.out0_ld = a->out0_ld; p.out0_off = a->out0_off;
  p.addend = a->addend; p.addend_ld = a->addend_ld; p.out1 = a->out1; p.out1_ld = a->out1_ld;
  p.nf = a->nf; p.nt = a->nt; p.axis = a->axis;
  p.h_state = a->h_state; p.c_state = a->c_state; p.state_flags = a->state_flags;
  if (a->axis == FNSSL_ALONG_FREQ) { p.rows = (int64_t)a->nb * a->nt; p.steps = a->nf; }
  else { p.rows = (int64_t)a->nb * a->nf; p.steps = a->nt; }
#define FNSSL_SIMT_CASE(HH)                                                         \
  case HH:                                                                          \
    return a->dtype == FNSSL_F32 ? launch_simt<float, HH>(p, a->num_dirs, st)       \
                                 : launch_simt<__half, HH>(p, a->num_dirs, st);
  switch (a->hidden) {
    FNSSL_SIMT_CASE(32)
    FNSSL_SIMT_CASE(64)
    FNSSL_SIMT_CASE(128)
    FNSSL_SIMT_CASE(256)
    default:
      FNSSL_FAIL("lstm(simt): hidden size %d not supported (32, 64, 128, 256)", a->hidden);
  }
#undef FNSSL_SIMT_CASE
}

int lstm_forward_simt(const fnssl_lstm_args* a, cudaStream_t st) { return lstm_forward_simt_save(a, nullptr, nullptr, st); }

}  // namespace fnssl
