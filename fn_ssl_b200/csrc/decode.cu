// IPD -> DOA decoding (SURVEY.md section 8f "next" #1): spatial spectrum against DP-IPD templates and iterative
// source detection / localisation.  Replaces SourceDetectLocalize.forward, meth_mode 'IDL'
// (FN-SSL/Lightning/Module.py:525-581), whose Python double loop over (batch, frame) with per-element device syncs
// dominates the reference's post-processing.
//
//   map[r][c]  = (1/(K/2)) * sum_k cur[r][k] * T[c][k]            r = (b, t), c = (ele, azi) candidate      (:553-556)
//   c*         = argmax_c map[r][c]  (first maximum)                                                          (:558)
//   ratio      = <T[c*], cur[r]> / <T[c*], T[c*]>;   cur[r] -= ratio * T[c*]                                  (:571-580)
// repeated max_sources times; fp32 throughout.
#include "common.cuh"

namespace fnssl {

constexpr int kSpecRows = 8;

// grid (ceil(ncand/128), ceil(R/8)); thread = candidate; templates transposed (K, ncand) for coalesced reads
__global__ void __launch_bounds__(128)
doa_spectrum_kernel(const float* __restrict__ cur, const float* __restrict__ templ_t, int R, int K, int ncand, float inv_scale,
                    float* __restrict__ map, float* __restrict__ ss_copy) {
  __shared__ float rows[kSpecRows][512];
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * kSpecRows;
  float acc[kSpecRows];
#pragma unroll
  for (int i = 0; i < kSpecRows; ++i) acc[i] = 0.0f;
  for (int k0 = 0; k0 < K; k0 += 512) {
    const int kn = min(512, K - k0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < kSpecRows * 512; idx += 128) {
      const int i = idx >> 9, k = idx & 511;
      rows[i][k] = (r0 + i < R && k < kn) ? cur[(size_t)(r0 + i) * K + k0 + k] : 0.0f;
    }
    __syncthreads();
    if (c < ncand) {
      for (int k = 0; k < kn; ++k) {
        const float t = __ldg(templ_t + (size_t)(k0 + k) * ncand + c);
#pragma unroll
        for (int i = 0; i < kSpecRows; ++i) acc[i] = fmaf(rows[i][k], t, acc[i]);
      }
    }
  }
  if (c < ncand) {
#pragma unroll
    for (int i = 0; i < kSpecRows; ++i) {
      if (r0 + i < R) {
        const float v = acc[i] * inv_scale;
        map[(size_t)(r0 + i) * ncand + c] = v;
        if (ss_copy) ss_copy[(size_t)(r0 + i) * ncand + c] = v;
      }
    }
  }
}

// one CTA per row: first-maximum argmax, projection ratio, residual update
__global__ void __launch_bounds__(256)
doa_idl_step_kernel(const float* __restrict__ map, const float* __restrict__ templ, int K, int ncand, int src, int nsrc,
                    int vad_mode, float* __restrict__ cur, int* __restrict__ idx_out, float* __restrict__ vad_out) {
  __shared__ float s_val[256];
  __shared__ int s_idx[256];
  __shared__ float s_num[256], s_den[256];
  const int r = blockIdx.x, tid = threadIdx.x;
  // torch.argmax semantics (Module.py:558): first maximum, and a NaN counts as the maximum (first NaN wins) -- so a row of
  // NaN / -inf still yields a valid candidate index instead of an out-of-range one
  auto better = [](float v, int i, float bv, int bi) {
    if (bi == 0x7fffffff) return true;                 // nothing held yet
    if (i == 0x7fffffff) return false;
    const bool vn = v != v, bn = bv != bv;
    if (vn || bn) return vn && (!bn || i < bi);
    return v > bv || (v == bv && i < bi);
  };
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = tid; c < ncand; c += 256) {
    const float v = map[(size_t)r * ncand + c];
    if (better(v, c, best, bi)) { best = v; bi = c; }
  }
  s_val[tid] = best; s_idx[tid] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      const float v = s_val[tid + o];
      const int i = s_idx[tid + o];
      if (better(v, i, s_val[tid], s_idx[tid])) { s_val[tid] = v; s_idx[tid] = i; }
    }
    __syncthreads();
  }
  const int cs = min(max(s_idx[0], 0), ncand - 1);
  const float* t = templ + (size_t)cs * K;
  float* x = cur + (size_t)r * K;
  float num = 0.0f, den = 0.0f;
  for (int k = tid; k < K; k += 256) { const float tv = t[k]; num = fmaf(tv, x[k], num); den = fmaf(tv, tv, den); }
  s_num[tid] = num; s_den[tid] = den;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { s_num[tid] += s_num[tid + o]; s_den[tid] += s_den[tid + o]; }
    __syncthreads();
  }
  const float ratio = s_num[0] / s_den[0];
  for (int k = tid; k < K; k += 256) x[k] -= ratio * t[k];
  if (tid == 0) {
    idx_out[(size_t)r * nsrc + src] = cs;
    vad_out[(size_t)r * nsrc + src] = vad_mode == 1 ? 1.0f : (vad_mode == 2 ? ratio : 0.0f);
  }
}

}  // namespace fnssl

using namespace fnssl;

extern "C" int fnssl_doa_decode_idl(const float* pred_ipd, const float* templ, const float* templ_t, int R, int K, int ncand,
                                    int max_sources, int vad_mode, float* cur, float* map, float* ss, int* idx_out,
                                    float* vad_out, void* stream) {
  FNSSL_REQUIRE(pred_ipd && templ && templ_t && cur && map && ss && idx_out && vad_out, "doa_decode: null pointer");
  FNSSL_REQUIRE(R > 0 && K > 0 && ncand > 0 && max_sources > 0, "doa_decode: bad shape");
  FNSSL_REQUIRE(vad_mode >= 0 && vad_mode <= 2, "doa_decode: bad vad_mode %d", vad_mode);
  cudaStream_t st = (cudaStream_t)stream;
  FNSSL_CUDA(cudaMemcpyAsync(cur, pred_ipd, (size_t)R * K * sizeof(float), cudaMemcpyDeviceToDevice, st));
  dim3 grid((ncand + 127) / 128, (R + kSpecRows - 1) / kSpecRows);
  const float inv_scale = 1.0f / ((float)K / 2.0f);
  for (int s = 0; s < max_sources; ++s) {
    doa_spectrum_kernel<<<grid, 128, 0, st>>>(cur, templ_t, R, K, ncand, inv_scale, map, s == 0 ? ss : nullptr);
    FNSSL_LAUNCH_CHECK("doa_spectrum_kernel");
    doa_idl_step_kernel<<<R, 256, 0, st>>>(map, templ, K, ncand, s, max_sources, vad_mode, cur, idx_out, vad_out);
    FNSSL_LAUNCH_CHECK("doa_idl_step_kernel");
  }
  return 0;
}
