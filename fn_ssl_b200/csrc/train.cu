// Training-side forward pieces: DP-IPD regression targets and the losses, on the GPU.
//
//   dpipd_targets_kernel   DPIPD.forward(source_doa) + the ground-truth branch of data_preprocess:
//                          FN-SSL/Lightning/Module.py:464-497 and main.py:227-265 (targets summed over sources, VAD-gated);
//                          IPDnet/runIPDnetOn.py:256-290 (per-source targets, silent sources -> the non-source target)
//   ipd_mse_*              cal_loss, FN-SSL/Lightning/main.py:191-198 (RemoveChFromBatch + permute + mse_loss)
//   ipd_pit_*              frame-level PIT loss, IPDnet/runIPDnetOn.py:188-206
//
// The reference builds the targets in float64 numpy on the host (a Python loop over mic pairs, then a host->device copy per
// batch); here one thread computes one (frame, source, pair, bin) phase in double precision, so the float32 result equals the
// reference's rounding.  Reductions are two-stage with a fixed order: results are deterministic run to run.
#include <math.h>

#include "common.cuh"

namespace fnssl {

// out layout: SUM : (rows, 2*nbins, P)          rows = nb*nt
//             PER : (rows, 2*nbins, P, ns)
__global__ void __launch_bounds__(256)
dpipd_targets_kernel(const float* __restrict__ doa, const float* __restrict__ vad, const float* __restrict__ mic,
                     const int* __restrict__ pairs, int rows, int ns, int P, int nf, double fre_max, double speed, int bin_lo,
                     int nbins, float vad_th, int per_source, const float* __restrict__ nonsrc, float* __restrict__ out) {
  const long long total = (long long)rows * nbins * P;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int p = (int)(idx % P);
  const int k = (int)((idx / P) % nbins);
  const long long r = idx / ((long long)P * nbins);
  const int m1 = pairs[2 * p], m2 = pairs[2 * p + 1];
  const double dx = (double)mic[3 * m1] - (double)mic[3 * m2];
  const double dy = (double)mic[3 * m1 + 1] - (double)mic[3 * m2 + 1];
  const double dz = (double)mic[3 * m1 + 2] - (double)mic[3 * m2 + 2];
  const double fre = fre_max * (double)(bin_lo + k) / (double)(nf - 1);       // np.linspace(0, fre_max, nf)[bin]
  float acc_re = 0.0f, acc_im = 0.0f;
  for (int s = 0; s < ns; ++s) {
    const double ele = (double)doa[(r * 2 + 0) * ns + s], azi = (double)doa[(r * 2 + 1) * ns + s];
    double se, ce, sa, ca;
    sincos(ele, &se, &ce);
    sincos(azi, &sa, &ca);
    const double itd = (se * ca * dx + se * sa * dy + ce * dz) / speed;          // Module.py:481-485 (t2 - t1)
    const double ph = 2.0 * M_PI * fre * itd;                                     // :486-487: -2 pi f ITD * (-1)
    double sp, cp;
    sincos(ph, &sp, &cp);
    float re = (float)cp, im = (float)sp;                                         // complex64 -> float32 (main.py:240-241)
    float gate = 1.0f;
    if (vad) gate = vad[r * ns + s] > vad_th ? 1.0f : 0.0f;                       // main.py:253-255 / runIPDnetOn.py:270-272
    if (per_source) {
      re *= gate; im *= gate;
      if (gate == 0.0f && nonsrc) {                                               // runIPDnetOn.py:277-281
        re = nonsrc[(size_t)k * P + p];
        im = nonsrc[(size_t)(nbins + k) * P + p];
      }
      out[((r * 2 * nbins + k) * P + p) * ns + s] = re;
      out[((r * 2 * nbins + nbins + k) * P + p) * ns + s] = im;
    } else {
      acc_re += re * gate; acc_im += im * gate;                                   // main.py:259 sum over sources
    }
  }
  if (!per_source) {
    out[(r * 2 * nbins + k) * P + p] = acc_re;
    out[(r * 2 * nbins + nbins + k) * P + p] = acc_im;
  }
}

// ---- deterministic two-stage sum -------------------------------------------------------------------------------
template <int kThreads>
__device__ __forceinline__ float block_sum(float v, float* sm) {
  const int tid = threadIdx.x;
  sm[tid] = v;
  __syncthreads();
#pragma unroll
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (tid < o) sm[tid] += sm[tid + o];
    __syncthreads();
  }
  return sm[0];
}

__global__ void __launch_bounds__(256) final_sum_kernel(const float* __restrict__ part, int n, float scale, float* __restrict__ out) {
  __shared__ float sm[256];
  float v = 0.0f;
  for (int i = threadIdx.x; i < n; i += 256) v += part[i];
  const float s = block_sum<256>(v, sm);
  if (threadIdx.x == 0) *out = s * scale;
}

// pred (nb*P, nt, nf2) [row b*P + p], gt (nb, nt, nf2, P): partial sums of (pred - gt)^2, one block per (b, t)
__global__ void __launch_bounds__(256)
ipd_mse_partial_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int P, int nt, int nf2, float* __restrict__ part) {
  __shared__ float sm[256];
  const int bt = blockIdx.x, b = bt / nt, t = bt % nt;
  float v = 0.0f;
  for (int i = threadIdx.x; i < nf2 * P; i += 256) {
    const int p = i % P, k = i / P;
    const float d = pred[((size_t)(b * P + p) * nt + t) * nf2 + k] - gt[((size_t)bt * nf2 + k) * P + p];
    v = fmaf(d, d, v);
  }
  const float s = block_sum<256>(v, sm);
  if (threadIdx.x == 0) part[bt] = s;
}

// Frame-level PIT.  pred, gt: (rows, K, ns) (the reference reshapes (nb,nt,2nf,nmic-1,ns) to exactly this, runIPDnetOn.py:199-200).
// cost[i][j] = sum_k (pred[k, i] - gt[k, j])^2; best permutation perm (target j <- prediction perm[j]) minimises sum_j cost[perm[j]][j];
// permutations are enumerated in itertools.permutations order and the first minimum wins (torchmetrics' exhaustive search).
constexpr int kMaxSrc = 4;
__global__ void __launch_bounds__(256)
ipd_pit_partial_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int K, int ns, float* __restrict__ part,
                       int* __restrict__ best_perm) {
  __shared__ float sm[256];
  __shared__ float cost[kMaxSrc * kMaxSrc];
  const int r = blockIdx.x;
  const float* pr = pred + (size_t)r * K * ns;
  const float* gr = gt + (size_t)r * K * ns;
  for (int ij = 0; ij < ns * ns; ++ij) {
    const int i = ij / ns, j = ij % ns;
    float v = 0.0f;
    for (int k = threadIdx.x; k < K; k += 256) {
      const float d = pr[(size_t)k * ns + i] - gr[(size_t)k * ns + j];
      v = fmaf(d, d, v);
    }
    const float s = block_sum<256>(v, sm);
    if (threadIdx.x == 0) cost[ij] = s;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int perm[kMaxSrc], best[kMaxSrc];
    for (int i = 0; i < ns; ++i) perm[i] = best[i] = i;
    float bestc = INFINITY;
    // lexicographic enumeration of the permutations of 0..ns-1 (== itertools.permutations order)
    while (true) {
      float c = 0.0f;
      for (int j = 0; j < ns; ++j) c += cost[perm[j] * ns + j];
      if (c < bestc) { bestc = c; for (int j = 0; j < ns; ++j) best[j] = perm[j]; }
      int i = ns - 2;
      while (i >= 0 && perm[i] > perm[i + 1]) --i;
      if (i < 0) break;
      int j = ns - 1;
      while (perm[j] < perm[i]) --j;
      int tmp = perm[i]; perm[i] = perm[j]; perm[j] = tmp;
      for (int a = i + 1, b = ns - 1; a < b; ++a, --b) { tmp = perm[a]; perm[a] = perm[b]; perm[b] = tmp; }
    }
    part[r] = bestc;
    if (best_perm) for (int j = 0; j < ns; ++j) best_perm[(size_t)r * ns + j] = best[j];
  }
}

}  // namespace fnssl

using namespace fnssl;

extern "C" int fnssl_dpipd_targets(const float* source_doa, const float* vad, const float* mic_pos, const int* pairs, int nb, int nt,
                                   int ns, int nmic, int P, int nf, float fre_max, float speed, int bin_lo, int nbins,
                                   float vad_threshold, int per_source, const float* nonsrc, float* out, void* stream) {
  FNSSL_REQUIRE(source_doa && mic_pos && pairs && out, "dpipd_targets: null pointer");
  FNSSL_REQUIRE(nb > 0 && nt > 0 && ns > 0 && nmic >= 2 && P > 0 && nf >= 2, "dpipd_targets: bad shape");
  FNSSL_REQUIRE(bin_lo >= 0 && nbins > 0 && bin_lo + nbins <= nf, "dpipd_targets: bins [%d, %d) outside 0..%d", bin_lo, bin_lo + nbins, nf);
  FNSSL_REQUIRE(speed > 0.0f, "dpipd_targets: bad speed of sound");
  const long long total = (long long)nb * nt * nbins * P;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  dpipd_targets_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(source_doa, vad, mic_pos, pairs, nb * nt, ns, P, nf, (double)fre_max,
                                                                (double)speed, bin_lo, nbins, vad_threshold, per_source, nonsrc, out);
  FNSSL_LAUNCH_CHECK("dpipd_targets_kernel");
  return 0;
}

extern "C" int fnssl_ipd_mse_loss(const float* pred, const float* gt, int nb, int P, int nt, int nf2, float* workspace, float* loss,
                                  void* stream) {
  FNSSL_REQUIRE(pred && gt && workspace && loss, "ipd_mse_loss: null pointer");
  FNSSL_REQUIRE(nb > 0 && P > 0 && nt > 0 && nf2 > 0, "ipd_mse_loss: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  ipd_mse_partial_kernel<<<nb * nt, 256, 0, st>>>(pred, gt, P, nt, nf2, workspace);
  FNSSL_LAUNCH_CHECK("ipd_mse_partial_kernel");
  final_sum_kernel<<<1, 256, 0, st>>>(workspace, nb * nt, 1.0f / ((float)nb * P * nt * nf2), loss);
  FNSSL_LAUNCH_CHECK("final_sum_kernel");
  return 0;
}

extern "C" int fnssl_ipd_pit_mse_loss(const float* pred, const float* gt, int rows, int K, int ns, float* workspace, float* loss,
                                      int* best_perm, void* stream) {
  FNSSL_REQUIRE(pred && gt && workspace && loss, "ipd_pit_mse_loss: null pointer");
  FNSSL_REQUIRE(rows > 0 && K > 0 && ns > 0 && ns <= kMaxSrc, "ipd_pit_mse_loss: bad shape (1..%d sources)", kMaxSrc);
  cudaStream_t st = (cudaStream_t)stream;
  ipd_pit_partial_kernel<<<rows, 256, 0, st>>>(pred, gt, K, ns, workspace, best_perm);
  FNSSL_LAUNCH_CHECK("ipd_pit_partial_kernel");
  final_sum_kernel<<<1, 256, 0, st>>>(workspace, rows, 1.0f / ((float)rows * K * ns), loss);
  FNSSL_LAUNCH_CHECK("final_sum_kernel");
  return 0;
}
