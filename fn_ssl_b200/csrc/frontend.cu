// Front end of the FN-SSL / IPDnet hot path: batched multi-channel STFT (512/256, periodic Hann,
// center=False), magnitude normaliser (forgetting_norm / global mean) and feature assembly.
//
// Reference call sites replaced (Audio-WestlakeU/FN-SSL):
//   STFT.forward            FN-SSL/Lightning/Module.py:48-68, IPDnet/Module.py:45-63
//   AddChToBatch.forward    FN-SSL/Lightning/Module.py:384-405
//   forgetting_norm         FN-SSL/Lightning/utils_.py:9-55
//   data_preprocess         FN-SSL/Lightning/main.py:206-225, IPDnet/runIPDnetOn.py:240-254, runIPDnetOff.py:248-251
//
// All three kernels are HBM-bound (about 4 FLOP/B); algorithmic bytes per TF-frame: 3080*M.
#include <stdarg.h>

#include <type_traits>

#include "common.cuh"

namespace fnssl {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// STFT
// ------------------------------------------------------------------------------------------------

constexpr int kWin = 512;
constexpr int kBins = 257;
constexpr int kStftThreads = 256;
constexpr int kStftWarps = kStftThreads / 32;
constexpr int kFftPad = 512 + 64;      // padded length of one real / imaginary work array (fft_pad)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 8-point DFT (forward, e^{-2 pi i mk/8}) in registers, natural order in and out: one decimation-in-frequency split
// into two 4-point DFTs.
__device__ __forceinline__ void dft8(float (&xr)[8], float (&xi)[8]) {
  const float kS = 0.70710678118654752440f;
  float ar[8], ai[8];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    ar[m] = xr[m] + xr[m + 4]; ai[m] = xi[m] + xi[m + 4];
    ar[m + 4] = xr[m] - xr[m + 4]; ai[m + 4] = xi[m] - xi[m + 4];
  }
  {   // odd half: multiply by W8^1, W8^2 = -i, W8^3
    const float r5 = (ar[5] + ai[5]) * kS, i5 = (ai[5] - ar[5]) * kS;
    const float r6 = ai[6], i6 = -ar[6];
    const float r7 = (ai[7] - ar[7]) * kS, i7 = -(ar[7] + ai[7]) * kS;
    ar[5] = r5; ai[5] = i5; ar[6] = r6; ai[6] = i6; ar[7] = r7; ai[7] = i7;
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {   // 4-point DFT of (b0..b3) = a[4h..4h+3] -> X[2q + h], q = 0..3
    const float c0r = ar[4 * h] + ar[4 * h + 2], c0i = ai[4 * h] + ai[4 * h + 2];
    const float c1r = ar[4 * h] - ar[4 * h + 2], c1i = ai[4 * h] - ai[4 * h + 2];
    const float c2r = ar[4 * h + 1] + ar[4 * h + 3], c2i = ai[4 * h + 1] + ai[4 * h + 3];
    const float dr = ar[4 * h + 1] - ar[4 * h + 3], di = ai[4 * h + 1] - ai[4 * h + 3];
    const float c3r = di, c3i = -dr;                                   // (b1 - b3) * (-i)
    xr[h] = c0r + c2r;     xi[h] = c0i + c2i;
    xr[2 + h] = c1r + c3r; xi[2 + h] = c1i + c3i;
    xr[4 + h] = c0r - c2r; xi[4 + h] = c0i - c2i;
    xr[6 + h] = c1r - c3r; xi[6 + h] = c1i - c3i;
  }
}

__device__ __forceinline__ int fft_pad(int i) { return i + (i >> 3); }   // one pad word per 8: conflict-free radix-8 strides

// One CTA = FR consecutive frames of one utterance, all channels.
//  1. the sample span [(t0*hop)*nch, ...) is staged in shared memory with ONE 1-D TMA bulk copy
//     (cp.async.bulk, completion on an mbarrier) when 16-byte aligned, else with plain loads;
//  2. each warp runs 512-point complex FFTs out of shared memory, one (frame, channel PAIR) job at a time: two real
//     channels ride in the real and imaginary part of one transform and are separated afterwards
//     (X0[k] = (Z[k] + conj Z[N-k]) / 2, X1[k] = (Z[k] - conj Z[N-k]) / 2i).  The transform is three radix-8 passes
//     (512 = 8 x 8 x 8, decimation in frequency, digit-reversed result) with the 8-point DFTs in registers: two
//     butterflies per lane per pass, two exchanges through a padded per-warp buffer instead of nine radix-2 passes;
//  3. the (FR*nch) x 257 results are transposed through shared memory so that every bin row of the
//     (nb, 257, nt, nch) output is written as one contiguous FR*nch*8-byte run.
//  4. FUSED FEATURE MODE (fnssl_stft_features_forward, T != void): instead of the spectrum the CTA writes the normalised
//     network features of its frames straight into the channels-last grid (R, nt, 256, ld) -- re/(mu+eps), im/(mu+eps) of the
//     row's channels, bins 1..256 (main.py:206-225) -- so the complex spectrum never exists in HBM.  mu comes from a first
//     pass of this same kernel that only produces the per-frame magnitude sums (spec == feat == nullptr).
struct StftFeat {
  void* feat;          // grid (R, nt, 256, ld) of T; nullptr = not in feature mode
  const float* mu;     // (R, nt) normaliser (nullptr with norm == NONE)
  int ld, pairing, norm;
  float eps;
};

template <typename T> __device__ __forceinline__ void st_chunk(T* dst, const float (&v)[8]);
__device__ __forceinline__ void row_channels(int r, int nch, int pairing, int& b, int& ci, int& cj);

template <typename T>
__global__ void __launch_bounds__(kStftThreads)
stft512_kernel(const float* __restrict__ signal, int nsample, int nch, int hop, int nt, int FR, int use_bulk,
               float2* __restrict__ spec, float* __restrict__ magsum, const StftFeat ff) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FR;
  const int nfr = min(FR, nt - t0);
  const int span = ((nfr - 1) * hop + kWin) * nch;  // floats

  float2* tw = reinterpret_cast<float2*>(smem_raw);                         // 512: exp(-2 pi i q / 512)
  float* hann = reinterpret_cast<float*>(tw + kWin);                         // 512
  float* fftbuf = hann + kWin;                                               // kStftWarps * 2 * kFftPad
  float2* stage = reinterpret_cast<float2*>(fftbuf + kStftWarps * 2 * kFftPad);   // FR*nch*257
  float* samples = reinterpret_cast<float*>(stage + (((size_t)FR * nch * kBins + 1) & ~(size_t)1));  // 16B aligned
  // more than 2 channels: the interleaved span is re-laid channel-major (odd row pitch) so that the FFT's stride-64 sample
  // reads are conflict-free instead of nch-way conflicted
  const bool deint = nch > 2;
  const int span_n = (FR - 1) * hop + kWin;
  const int pitch = span_n | 1;
  float* samples_t = samples + (((size_t)span_n * nch + 3) & ~(size_t)3);
  __shared__ __align__(8) unsigned long long mbar;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* src = signal + ((size_t)b * nsample + (size_t)t0 * hop) * nch;

  if (use_bulk) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)span * 4u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(samples)),
          "l"(src), "r"(bytes), "r"(smem_u32(&mbar))
          : "memory");
    }
  } else {
    for (int i = tid; i < span; i += kStftThreads) samples[i] = src[i];
  }
  // twiddles exp(-2*pi*i*q/512) and the periodic Hann window, computed with exact-argument sinpi/cospi
  for (int n = tid; n < kWin; n += kStftThreads) {
    float sn, cs;
    sincospif((float)n * (1.0f / 256.0f), &sn, &cs);
    tw[n] = make_float2(cs, -sn);
    hann[n] = 0.5f - 0.5f * cs;
  }
  if (use_bulk) {
    // wait for the bulk copy (phase 0); bounded spin, then trap instead of hanging the GPU
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; ++it) {
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
          : "=r"(done)
          : "r"(smem_u32(&mbar)), "r"(0u)
          : "memory");
    }
    if (!done) __trap();
  }
  __syncthreads();

  if (deint) {
    for (int i = tid; i < span; i += kStftThreads) {
      const int n = i / nch, c = i - n * nch;
      samples_t[c * pitch + n] = samples[i];
    }
    __syncthreads();
  }
  float* wre = fftbuf + warp * 2 * kFftPad;
  float* wim = wre + kFftPad;
  const int npair = (nch + 1) >> 1;
  const int njobs = nfr * npair;
  for (int job = warp; job < njobs; job += kStftWarps) {
    const int tl = job / npair, pj = job - tl * npair;
    const int c0 = 2 * pj;
    const bool has1 = c0 + 1 < nch;
    const float* x0 = deint ? samples_t + (size_t)c0 * pitch + tl * hop : samples + (size_t)tl * hop * nch + c0;
    const int sn = deint ? 1 : nch, sc = deint ? pitch : 1;      // sample / channel strides of x0
    // ---- pass 1 (span 512): butterfly j takes z[j + 64 m]; windowed samples straight from the staged span
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h;
      float xr[8], xi[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int n = j + 64 * m;
        const float w = hann[n];
        xr[m] = x0[n * sn] * w;
        xi[m] = has1 ? x0[n * sn + sc] * w : 0.0f;
      }
      dft8(xr, xi);
      wre[fft_pad(j)] = xr[0]; wim[fft_pad(j)] = xi[0];
#pragma unroll
      for (int r = 1; r < 8; ++r) {
        const float2 t = tw[(j * r) & 511];
        wre[fft_pad(j + 64 * r)] = xr[r] * t.x - xi[r] * t.y;
        wim[fft_pad(j + 64 * r)] = xr[r] * t.y + xi[r] * t.x;
      }
    }
    __syncwarp();
    // ---- pass 2 (span 64 inside each block of 64)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = lane + 32 * h, blk = t >> 3, j = t & 7;
      float xr[8], xi[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) { xr[m] = wre[fft_pad(64 * blk + j + 8 * m)]; xi[m] = wim[fft_pad(64 * blk + j + 8 * m)]; }
      dft8(xr, xi);
      wre[fft_pad(64 * blk + j)] = xr[0]; wim[fft_pad(64 * blk + j)] = xi[0];
#pragma unroll
      for (int r = 1; r < 8; ++r) {
        const float2 tq = tw[(8 * j * r) & 511];
        wre[fft_pad(64 * blk + j + 8 * r)] = xr[r] * tq.x - xi[r] * tq.y;
        wim[fft_pad(64 * blk + j + 8 * r)] = xr[r] * tq.y + xi[r] * tq.x;
      }
    }
    __syncwarp();
    // ---- pass 3 (span 8): position 64 k0 + 8 k1 + k2 ends up holding Z[k0 + 8 k1 + 64 k2]
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = lane + 32 * h;
      float xr[8], xi[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) { xr[m] = wre[fft_pad(8 * c + m)]; xi[m] = wim[fft_pad(8 * c + m)]; }
      dft8(xr, xi);
#pragma unroll
      for (int r = 0; r < 8; ++r) { wre[fft_pad(8 * c + r)] = xr[r]; wim[fft_pad(8 * c + r)] = xi[r]; }
    }
    __syncwarp();
    // ---- separate the two real channels, bins 0..256 -> staging; magnitude sums for the normaliser
    float2* st0 = stage + (size_t)(tl * nch + c0) * kBins;
    float ms0 = 0.0f, ms1 = 0.0f;
    for (int k = lane; k < kBins; k += 32) {
      const int kn = (kWin - k) & (kWin - 1);
      const int pk = fft_pad(64 * (k & 7) + 8 * ((k >> 3) & 7) + (k >> 6));
      const int pn = fft_pad(64 * (kn & 7) + 8 * ((kn >> 3) & 7) + (kn >> 6));
      const float a = wre[pk], bq = wim[pk];
      if (has1) {
        const float c = wre[pn], d = wim[pn];
        const float2 v0 = make_float2(0.5f * (a + c), 0.5f * (bq - d));
        const float2 v1 = make_float2(0.5f * (bq + d), 0.5f * (c - a));
        st0[k] = v0; st0[kBins + k] = v1;
        ms0 += sqrtf(v0.x * v0.x + v0.y * v0.y);
        ms1 += sqrtf(v1.x * v1.x + v1.y * v1.y);
      } else {
        st0[k] = make_float2(a, bq);
        ms0 += sqrtf(a * a + bq * bq);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ms0 += __shfl_xor_sync(0xffffffffu, ms0, o);
      ms1 += __shfl_xor_sync(0xffffffffu, ms1, o);
    }
    if (lane == 0 && magsum) {
      magsum[((size_t)b * nch + c0) * nt + t0 + tl] = ms0;
      if (has1) magsum[((size_t)b * nch + c0 + 1) * nt + t0 + tl] = ms1;
    }
    __syncwarp();
  }
  __syncthreads();
  if (spec) {
    // transposed, coalesced write: spec[b][f][t0 + tl][ch], (tl, ch) fastest
    const int run = nfr * nch;
    for (int idx = tid; idx < kBins * run; idx += kStftThreads) {
      const int f = idx / run, j = idx - f * run;
      spec[((size_t)b * kBins + f) * nt * nch + (size_t)t0 * nch + j] = stage[(size_t)j * kBins + f];
    }
  }
  if constexpr (!std::is_same<T, void>::value) {
    if (ff.feat) {
      // one thread = one (row, frame, bin) position: its ld channels as 16-byte chunks; consecutive threads = consecutive bins,
      // so a warp writes one contiguous run of 32 * ld elements and a frame of one row is one contiguous 256 * ld run
      constexpr int kChunk = sizeof(T) == 2 ? 8 : 4;
      const bool all = (ff.pairing == FNSSL_PAIRS_ALL);
      const int P = all ? 1 : (ff.pairing == FNSSL_PAIRS_M ? nch - 1 : nch * (nch - 1) / 2);
      const int C = all ? 2 * nch : 4;
      const int half = C / 2;
      T* feat = reinterpret_cast<T*>(ff.feat);
      for (int idx = tid; idx < P * nfr * 256; idx += kStftThreads) {
        const int fl = idx & 255, tl = (idx >> 8) % nfr, pp = (idx >> 8) / nfr;
        const int r = b * P + pp;
        int bb, ci, cj;
        row_channels(r, nch, ff.pairing, bb, ci, cj);
        const float den = ff.norm != FNSSL_NORM_NONE ? ff.mu[(size_t)r * nt + t0 + tl] + ff.eps : 1.0f;
        const float2* src = stage + (size_t)tl * nch * kBins + 1 + fl;          // + ch * kBins: bin 1 + fl of channel ch
        T* dst = feat + (((size_t)r * nt + t0 + tl) * 256 + fl) * ff.ld;
        for (int c0 = 0; c0 < ff.ld; c0 += kChunk) {
          float v[8];
#pragma unroll
          for (int i = 0; i < kChunk; ++i) {
            const int c = c0 + i;
            float x = 0.0f;
            if (c < C) {
              const int k = (c < half) ? c : c - half;
              const int ch = all ? k : (k == 0 ? ci : cj);
              const float2 z = src[(size_t)ch * kBins];
              x = (c < half) ? z.x : z.y;
              if (ff.norm != FNSSL_NORM_NONE) x = x / den;
            }
            v[i] = x;
          }
          st_chunk(dst + c0, v);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// normaliser: one warp per feature row (utils_.py:27-44)
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void row_channels(int r, int nch, int pairing, int& b, int& ci, int& cj) {
  if (pairing == FNSSL_PAIRS_ALL) { b = r; ci = 0; cj = 0; return; }
  const int P = (pairing == FNSSL_PAIRS_M) ? (nch - 1) : nch * (nch - 1) / 2;
  b = r / P;
  int p = r - b * P;
  if (pairing == FNSSL_PAIRS_M) { ci = 0; cj = p + 1; return; }
  int i = 0;
  while (p >= nch - 1 - i) { p -= nch - 1 - i; ++i; }   // lexicographic (i<j), Module.py:398-404
  ci = i; cj = i + 1 + p;
}

// One warp per feature row: the per-frame means are gathered in parallel into shared memory, lane 0 runs the
// T-sequential recursion out of shared memory (no dependent global loads), the result is written back coalesced.
__global__ void __launch_bounds__(32)
norm_scan_kernel(const float* __restrict__ magsum, int R, int nt, int nch, int nbins, int pairing, int norm,
                 int sample_length, float* __restrict__ mu_out, long long t0, float* __restrict__ mu_state) {
  extern __shared__ float fm[];   // [nt] frame means, overwritten by mu | [nt] a_t | [nt] 1 - a_t
  float* ca = fm + nt;
  float* com = ca + nt;
  const int r = blockIdx.x;
  const int lane = threadIdx.x;
  if (r >= R) return;
  int b, ci, cj;
  row_channels(r, nch, pairing, b, ci, cj);
  const int C = (pairing == FNSSL_PAIRS_ALL) ? nch : 2;
  const float cnt = (float)(C * nbins);
  for (int t = lane; t < nt; t += 32) {
    float s;
    if (pairing == FNSSL_PAIRS_ALL) {
      s = 0.0f;
      for (int c = 0; c < nch; ++c) s += magsum[((size_t)b * nch + c) * nt + t];
    } else {
      s = magsum[((size_t)b * nch + ci) * nt + t] + magsum[((size_t)b * nch + cj) * nt + t];
    }
    fm[t] = s / cnt;
    // recursion coefficients (fp32 tensor arithmetic in the reference, utils_.py:31), one frame per lane
    const double alpha = (double)(sample_length - 1) / (double)(sample_length + 1);
    const long long tt = t0 + t;
    if (tt < sample_length) {
      const float a = (float)fmin((double)(tt - 1) / (double)(tt + 1), alpha);
      ca[t] = a; com[t] = 1.0f - a;
    } else {
      ca[t] = (float)alpha; com[t] = (float)(1.0 - alpha);
    }
  }
  __syncwarp();
  if (lane == 0) {
    if (norm == FNSSL_NORM_GLOBAL) {
      float s = 0.0f;
      for (int t = 0; t < nt; ++t) s += fm[t];
      const float m = s / (float)nt;
      for (int t = 0; t < nt; ++t) fm[t] = m;
    } else {
      float mu = (t0 > 0 && mu_state) ? mu_state[r] : 0.0f;   // a chunk of a longer stream resumes the recursion
      for (int t = 0; t < nt; ++t) {
        mu = ca[t] * mu + com[t] * fm[t];
        fm[t] = mu;
      }
      if (mu_state) mu_state[r] = mu;
    }
  }
  __syncwarp();
  for (int t = lane; t < nt; t += 32) mu_out[(size_t)r * nt + t] = fm[t];
}

// ------------------------------------------------------------------------------------------------
// feature assembly: (nb, 257, nt, nch) complex -> grid (R, nt, 256, ld) [+ (R, C, 256, nt) f32]
// ------------------------------------------------------------------------------------------------

constexpr int kAsmTF = 16;  // bins per tile
constexpr int kAsmTT = 32;  // frames per tile

// store `n` consecutive channels (n = 8 halves or 4 floats = 16 bytes) of one grid position
// fp16 grids saturate at +-65504: a normalised feature re/(mu+eps) can exceed the fp16 range (worst case C*257/(1-alpha), e.g.
// a tonal onset after digital silence), and an inf would turn into NaN in the along-time LSTM's carried state (sat_f16, common.cuh)
template <> __device__ __forceinline__ void st_chunk<__half>(__half* dst, const float (&v)[8]) {
  const __half2 h0 = __floats2half2_rn(sat_f16(v[0]), sat_f16(v[1])), h1 = __floats2half2_rn(sat_f16(v[2]), sat_f16(v[3]));
  const __half2 h2 = __floats2half2_rn(sat_f16(v[4]), sat_f16(v[5])), h3 = __floats2half2_rn(sat_f16(v[6]), sat_f16(v[7]));
  uint4 pk;
  pk.x = *reinterpret_cast<const uint32_t*>(&h0); pk.y = *reinterpret_cast<const uint32_t*>(&h1);
  pk.z = *reinterpret_cast<const uint32_t*>(&h2); pk.w = *reinterpret_cast<const uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(dst) = pk;
}
template <> __device__ __forceinline__ void st_chunk<float>(float* dst, const float (&v)[8]) {
  *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
}

// VEC: ld is a multiple of the 16-byte chunk (8 halves / 4 floats) and the grid is 16-byte aligned: one thread owns one
// (frame, bin) position and writes its ld channels as 16-byte stores, consecutive threads = consecutive bins, so a warp
// writes one contiguous run of 32 * ld elements.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256)
assemble_kernel(const float2* __restrict__ spec, const float* __restrict__ mu, int nb, int nt, int nch, int pairing,
                int norm, float eps, T* __restrict__ feat, int ld, float* __restrict__ feat_cfirst) {
  extern __shared__ float2 tile[];  // [kAsmTT][kAsmTF][nch]
  constexpr int kChunk = sizeof(T) == 2 ? 8 : 4;
  const int b = blockIdx.z;
  const int f0 = 1 + blockIdx.y * kAsmTF;  // spectrum bin of the first feature bin in the tile
  const int t0 = blockIdx.x * kAsmTT;
  const int tn = min(kAsmTT, nt - t0);
  const int tid = threadIdx.x;
  // coalesced read: for each bin a contiguous run of tn*nch complex values
  const int runlen = tn * nch;
  for (int idx = tid; idx < kAsmTF * runlen; idx += blockDim.x) {
    const int fl = idx / runlen, j = idx - fl * runlen;
    const int tl = j / nch, ch = j - tl * nch;
    tile[((size_t)tl * kAsmTF + fl) * nch + ch] = spec[((size_t)b * kBins + f0 + fl) * nt * nch + (size_t)t0 * nch + j];
  }
  __syncthreads();
  const bool all = (pairing == FNSSL_PAIRS_ALL);
  const int P = all ? 1 : (pairing == FNSSL_PAIRS_M ? nch - 1 : nch * (nch - 1) / 2);
  const int C = all ? 2 * nch : 4;
  const int half = C / 2;
  for (int p = 0; p < P; ++p) {
    const int r = b * P + p;
    int bb, ci, cj;
    row_channels(r, nch, pairing, bb, ci, cj);
    if (VEC) {
      for (int idx = tid; idx < tn * kAsmTF; idx += blockDim.x) {
        const int fl = idx % kAsmTF, tl = idx / kAsmTF;
        const float2* src = tile + ((size_t)tl * kAsmTF + fl) * nch;
        const float den = norm != FNSSL_NORM_NONE ? mu[(size_t)r * nt + t0 + tl] + eps : 1.0f;
        T* dst = feat + (((size_t)r * nt + t0 + tl) * 256 + (f0 - 1 + fl)) * ld;
        for (int c0 = 0; c0 < ld; c0 += kChunk) {
          float v[8];
#pragma unroll
          for (int i = 0; i < kChunk; ++i) {
            const int c = c0 + i;
            float x = 0.0f;
            if (c < C) {
              const int k = (c < half) ? c : c - half;
              const int ch = all ? k : (k == 0 ? ci : cj);
              const float2 z = src[ch];
              x = (c < half) ? z.x : z.y;
              if (norm != FNSSL_NORM_NONE) x = x / den;
            }
            v[i] = x;
          }
          st_chunk(dst + c0, v);
        }
      }
    } else {
      // grid write: (t, f, c), c fastest
      for (int idx = tid; idx < tn * kAsmTF * ld; idx += blockDim.x) {
        const int c = idx % ld;
        const int fl = (idx / ld) % kAsmTF;
        const int tl = idx / (ld * kAsmTF);
        float v = 0.0f;
        if (c < C) {
          const int k = (c < half) ? c : c - half;
          const int ch = all ? k : (k == 0 ? ci : cj);
          const float2 x = tile[((size_t)tl * kAsmTF + fl) * nch + ch];
          v = (c < half) ? x.x : x.y;
          if (norm != FNSSL_NORM_NONE) v = v / (mu[(size_t)r * nt + t0 + tl] + eps);
        }
        st_act<T>(feat + (((size_t)r * nt + t0 + tl) * 256 + (f0 - 1 + fl)) * ld + c, v);
      }
    }
    if (feat_cfirst) {
      for (int idx = tid; idx < C * kAsmTF * tn; idx += blockDim.x) {
        const int tl = idx % tn;
        const int fl = (idx / tn) % kAsmTF;
        const int c = idx / (tn * kAsmTF);
        const int k = (c < half) ? c : c - half;
        const int ch = all ? k : (k == 0 ? ci : cj);
        const float2 x = tile[((size_t)tl * kAsmTF + fl) * nch + ch];
        float v = (c < half) ? x.x : x.y;
        if (norm != FNSSL_NORM_NONE) v = v / (mu[(size_t)r * nt + t0 + tl] + eps);
        feat_cfirst[(((size_t)r * C + c) * 256 + (f0 - 1 + fl)) * nt + t0 + tl] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// layout converters
// ------------------------------------------------------------------------------------------------

// (nb, C, nf, nt) f32 -> grid (nb, nt, nf, ld)[ch_off + c].  CTA = (b, f, 32 frames), all channels.
template <typename T>
__global__ void cfirst_to_grid_kernel(const float* __restrict__ src, int C, int nf, int nt, T* __restrict__ dst, int ld,
                                      int ch_off) {
  extern __shared__ float ctile[];  // [C][33]
  const int b = blockIdx.z, f = blockIdx.y, t0 = blockIdx.x * 32;
  const int tn = min(32, nt - t0);
  for (int idx = threadIdx.x; idx < C * 32; idx += blockDim.x) {
    const int c = idx >> 5, tl = idx & 31;
    if (tl < tn) ctile[c * 33 + tl] = src[(((size_t)b * C + c) * nf + f) * nt + t0 + tl];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < tn * C; idx += blockDim.x) {
    const int tl = idx / C, c = idx - tl * C;
    st_act<T>(dst + (((size_t)b * nt + t0 + tl) * nf + f) * ld + ch_off + c, ctile[c * 33 + tl]);
  }
}

template <typename T>
__global__ void grid_to_cfirst_kernel(const T* __restrict__ src, int ld, int ch_off, int C, int nf, int nt,
                                      float* __restrict__ dst) {
  extern __shared__ float ctile[];  // [C][33]
  const int b = blockIdx.z, f = blockIdx.y, t0 = blockIdx.x * 32;
  const int tn = min(32, nt - t0);
  for (int idx = threadIdx.x; idx < tn * C; idx += blockDim.x) {
    const int tl = idx / C, c = idx - tl * C;
    ctile[c * 33 + tl] = ld_act<T>(src + (((size_t)b * nt + t0 + tl) * nf + f) * ld + ch_off + c);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < C * 32; idx += blockDim.x) {
    const int c = idx >> 5, tl = idx & 31;
    if (tl < tn) dst[(((size_t)b * C + c) * nf + f) * nt + t0 + tl] = ctile[c * 33 + tl];
  }
}

template <typename TS, typename TD>
__global__ void grid_copy_kernel(const TS* __restrict__ src, int src_ld, int src_off, TD* __restrict__ dst, int dst_ld,
                                 int dst_off, int64_t npos, int C) {
  const int64_t total = npos * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / C;
    const int c = (int)(i - p * C);
    st_act<TD>(dst + p * dst_ld + dst_off + c, ld_act<TS>(src + p * src_ld + src_off + c));
  }
}

// same dtype, same channel stride, no offsets: a plain 16-byte-vector copy at HBM rate (4 vectors in flight per thread)
__global__ void __launch_bounds__(256) grid_copy_vec_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t nvec) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    const uint4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
    dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
  }
  for (; i < nvec; i += stride) dst[i] = __ldcs(src + i);
}

template <typename T>
__global__ void grid_add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    st_act<T>(dst + i, ld_act<T>(a + i) + ld_act<T>(b + i));
}

}  // namespace fnssl

using namespace fnssl;

extern "C" {

int fnssl_abi_version(void) { return FNSSL_ABI_VERSION; }
const char* fnssl_last_error(void) { return fnssl::g_err; }

int fnssl_stft_num_frames(int nsample, int win_len, int hop) {
  if (nsample < win_len || hop <= 0) return 0;
  return (nsample - win_len) / hop + 1;
}

}  // extern "C"

// shared launcher of the three modes of stft512_kernel: spectrum (+ magnitude sums), magnitude sums only, fused features
template <typename T>
static int launch_stft(const float* signal, int nb, int nsample, int nch, int hop, int nt, float* spec, float* magsum,
                       const fnssl::StftFeat& ff, cudaStream_t st) {
  int FR = 16 / nch;
  if (FR < 1) FR = 1;
  if (FR > nt) FR = nt;
  const size_t span = ((size_t)(FR - 1) * hop + kWin) * nch;
  const size_t smem = kWin * sizeof(float2) + kWin * sizeof(float) + (size_t)kStftWarps * 2 * kFftPad * sizeof(float) +
                      (((size_t)FR * nch * kBins + 1) & ~(size_t)1) * sizeof(float2) + ((span + 3) & ~(size_t)3) * sizeof(float) +
                      (nch > 2 ? (size_t)nch * ((((size_t)(FR - 1) * hop + kWin)) | 1) * sizeof(float) : 0);
  FNSSL_REQUIRE(smem <= 200 * 1024, "stft: too many channels for one CTA (%d)", nch);
  const int use_bulk = ((reinterpret_cast<uintptr_t>(signal) & 15) == 0) && (((size_t)hop * nch * 4) % 16 == 0) &&
                       (((size_t)nsample * nch * 4) % 16 == 0) && (((size_t)kWin * nch * 4) % 16 == 0);
  FNSSL_CUDA(cudaFuncSetAttribute(stft512_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((nt + FR - 1) / FR, nb);
  stft512_kernel<T><<<grid, kStftThreads, smem, st>>>(signal, nsample, nch, hop, nt, FR, use_bulk, reinterpret_cast<float2*>(spec),
                                                    magsum, ff);
  FNSSL_LAUNCH_CHECK("stft512_kernel");
  return 0;
}

extern "C" {

int fnssl_stft_forward(const float* signal, int nb, int nsample, int nch, int win_len, int hop, int nfft, float* spec,
                       float* magsum, void* stream) {
  FNSSL_REQUIRE(win_len == 512 && nfft == 512, "stft: only win_len = nfft = 512 is implemented (got %d/%d)", win_len, nfft);
  FNSSL_REQUIRE(hop > 0 && hop <= 512, "stft: hop must be in (0, 512] (got %d)", hop);
  FNSSL_REQUIRE(nb > 0 && nch > 0 && nch <= 64, "stft: bad nb/nch (%d/%d)", nb, nch);
  const int nt = fnssl_stft_num_frames(nsample, win_len, hop);
  FNSSL_REQUIRE(nt > 0, "stft: signal shorter than one window (nsample=%d)", nsample);
  FNSSL_REQUIRE(signal && spec, "stft: null pointer");
  return launch_stft<void>(signal, nb, nsample, nch, hop, nt, spec, magsum, fnssl::StftFeat{nullptr, nullptr, 0, 0, 0, 0.0f},
                           (cudaStream_t)stream);
}

int fnssl_norm_forward(const float* magsum, int nb, int nch, int nt, int nbins, int pairing, int norm, int sample_length,
                       float* mu, void* stream);

int fnssl_stft_features_forward(const float* signal, int nb, int nsample, int nch, int win_len, int hop, int nfft, int pairing,
                                int norm, int sample_length, float eps, float* magsum, float* mu, void* feat, int dtype, int ld,
                                void* stream) {
  FNSSL_REQUIRE(win_len == 512 && nfft == 512, "stft_features: only win_len = nfft = 512 is implemented (got %d/%d)", win_len, nfft);
  FNSSL_REQUIRE(hop > 0 && hop <= 512, "stft_features: hop must be in (0, 512] (got %d)", hop);
  FNSSL_REQUIRE(nb > 0 && nch > 0 && nch <= 64, "stft_features: bad nb/nch (%d/%d)", nb, nch);
  FNSSL_REQUIRE(pairing >= 0 && pairing <= 2, "stft_features: bad pairing %d", pairing);
  FNSSL_REQUIRE(pairing == FNSSL_PAIRS_ALL || nch >= 2, "stft_features: pair modes need >= 2 channels");
  FNSSL_REQUIRE(norm >= 0 && norm <= 3, "stft_features: bad norm %d", norm);
  FNSSL_REQUIRE(dtype == FNSSL_F32 || dtype == FNSSL_F16, "stft_features: bad dtype %d", dtype);
  const int nt = fnssl_stft_num_frames(nsample, win_len, hop);
  FNSSL_REQUIRE(nt > 0, "stft_features: signal shorter than one window (nsample=%d)", nsample);
  const int C = fnssl_feature_channels(nch, pairing);
  FNSSL_REQUIRE(ld >= C && ld % (dtype == FNSSL_F16 ? 8 : 4) == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0,
                "stft_features: the grid must be 16-byte aligned with ld (%d) >= %d channels and a multiple of 16 bytes", ld, C);
  FNSSL_REQUIRE(signal && feat && (norm == FNSSL_NORM_NONE || mu) && (norm == FNSSL_NORM_NONE || norm == FNSSL_NORM_GIVEN || magsum),
                "stft_features: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (norm == FNSSL_NORM_FORGETTING || norm == FNSSL_NORM_GLOBAL) {
    // pass 1: FFT -> per-(utterance, channel, frame) sums of |X| only; then the T-sequential recursion (utils_.py:27-44)
    if (launch_stft<void>(signal, nb, nsample, nch, hop, nt, nullptr, magsum, fnssl::StftFeat{nullptr, nullptr, 0, 0, 0, 0.0f}, st)) return 1;
    if (fnssl_norm_forward(magsum, nb, nch, nt, kBins, pairing, norm, sample_length, mu, stream)) return 1;
  }
  // pass 2: FFT again (the signal is L2-resident), features written directly
  const fnssl::StftFeat ff{feat, norm == FNSSL_NORM_NONE ? nullptr : mu, ld, pairing, norm, eps};
  if (dtype == FNSSL_F16) return launch_stft<__half>(signal, nb, nsample, nch, hop, nt, nullptr, nullptr, ff, st);
  return launch_stft<float>(signal, nb, nsample, nch, hop, nt, nullptr, nullptr, ff, st);
}

int fnssl_feature_rows(int nb, int nch, int pairing) {
  if (pairing == FNSSL_PAIRS_M) return nb * (nch - 1);
  if (pairing == FNSSL_PAIRS_MM) return nb * (nch * (nch - 1) / 2);
  return nb;
}
int fnssl_feature_channels(int nch, int pairing) { return pairing == FNSSL_PAIRS_ALL ? 2 * nch : 4; }

int fnssl_norm_forward(const float* magsum, int nb, int nch, int nt, int nbins, int pairing, int norm, int sample_length,
                       float* mu, void* stream) {
  FNSSL_REQUIRE(magsum && mu && nb > 0 && nch > 0 && nt > 0 && nbins > 0, "norm: bad arguments");
  FNSSL_REQUIRE(pairing >= 0 && pairing <= 2, "norm: bad pairing %d", pairing);
  FNSSL_REQUIRE(norm == FNSSL_NORM_FORGETTING || norm == FNSSL_NORM_GLOBAL, "norm: bad norm %d", norm);
  FNSSL_REQUIRE(pairing == FNSSL_PAIRS_ALL || nch >= 2, "norm: pair modes need >= 2 channels");
  const int R = fnssl_feature_rows(nb, nch, pairing);
  FNSSL_REQUIRE((size_t)nt * 12 <= 200 * 1024, "norm: too many frames (%d)", nt);
  FNSSL_CUDA(cudaFuncSetAttribute(norm_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, nt * 12));
  norm_scan_kernel<<<R, 32, (size_t)nt * 12, (cudaStream_t)stream>>>(magsum, R, nt, nch, nbins, pairing, norm, sample_length, mu,
                                                                    0, nullptr);
  FNSSL_LAUNCH_CHECK("norm_scan_kernel");
  return 0;
}

int fnssl_norm_stream_forward(const float* magsum, int nb, int nch, int nt, int nbins, int pairing, int sample_length,
                              long long t0, float* mu_state, float* mu, void* stream) {
  FNSSL_REQUIRE(magsum && mu && mu_state && nb > 0 && nch > 0 && nt > 0 && nbins > 0 && t0 >= 0, "norm(stream): bad arguments");
  FNSSL_REQUIRE(pairing >= 0 && pairing <= 2, "norm(stream): bad pairing %d", pairing);
  FNSSL_REQUIRE(pairing == FNSSL_PAIRS_ALL || nch >= 2, "norm(stream): pair modes need >= 2 channels");
  const int R = fnssl_feature_rows(nb, nch, pairing);
  FNSSL_REQUIRE((size_t)nt * 12 <= 200 * 1024, "norm(stream): too many frames (%d)", nt);
  FNSSL_CUDA(cudaFuncSetAttribute(norm_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, nt * 12));
  norm_scan_kernel<<<R, 32, (size_t)nt * 12, (cudaStream_t)stream>>>(magsum, R, nt, nch, nbins, pairing, FNSSL_NORM_FORGETTING,
                                                                    sample_length, mu, t0, mu_state);
  FNSSL_LAUNCH_CHECK("norm_scan_kernel");
  return 0;
}

int fnssl_features_forward(const float* spec, const float* magsum, int nb, int nt, int nch, int pairing, int norm,
                           int sample_length, float eps, float* mu, void* feat, int dtype, int ld, float* feat_cfirst,
                           void* stream) {
  FNSSL_REQUIRE(pairing >= 0 && pairing <= 2, "features: bad pairing %d", pairing);
  FNSSL_REQUIRE(norm >= 0 && norm <= 3, "features: bad norm %d", norm);
  FNSSL_REQUIRE(pairing == FNSSL_PAIRS_ALL || nch >= 2, "features: pair modes need >= 2 channels");
  FNSSL_REQUIRE(dtype == FNSSL_F32 || dtype == FNSSL_F16, "features: bad dtype %d", dtype);
  const int C = fnssl_feature_channels(nch, pairing);
  const int R = fnssl_feature_rows(nb, nch, pairing);
  FNSSL_REQUIRE(ld >= C, "features: ld (%d) < channels (%d)", ld, C);
  FNSSL_REQUIRE(spec && feat && (norm == FNSSL_NORM_NONE || (norm == FNSSL_NORM_GIVEN && mu) || (magsum && mu)), "features: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (norm == FNSSL_NORM_FORGETTING || norm == FNSSL_NORM_GLOBAL) {
    if (fnssl_norm_forward(magsum, nb, nch, nt, kBins, pairing, norm, sample_length, mu, stream)) return 1;
  }
  dim3 grid((nt + kAsmTT - 1) / kAsmTT, 256 / kAsmTF, nb);
  const size_t smem = (size_t)kAsmTF * kAsmTT * nch * sizeof(float2);
#define FNSSL_ASM(T, VEC)                                                                                              \
  do {                                                                                                                   \
    FNSSL_CUDA(cudaFuncSetAttribute(assemble_kernel<T, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    assemble_kernel<T, VEC><<<grid, 256, smem, st>>>(reinterpret_cast<const float2*>(spec), mu, nb, nt, nch, pairing, norm, \
                                                     eps, (T*)feat, ld, feat_cfirst);                                    \
  } while (0)
  const bool al16 = (reinterpret_cast<uintptr_t>(feat) & 15) == 0;
  if (dtype == FNSSL_F32) {
    if (al16 && ld % 4 == 0) FNSSL_ASM(float, true); else FNSSL_ASM(float, false);
  } else {
    if (al16 && ld % 8 == 0) FNSSL_ASM(__half, true); else FNSSL_ASM(__half, false);
  }
#undef FNSSL_ASM
  FNSSL_LAUNCH_CHECK("assemble_kernel");
  return 0;
}

int fnssl_cfirst_to_grid(const float* src, int nb, int C, int nf, int nt, void* dst, int dtype, int ld, int ch_off,
                         void* stream) {
  FNSSL_REQUIRE(src && dst && nb > 0 && C > 0 && nf > 0 && nt > 0, "cfirst_to_grid: bad arguments");
  FNSSL_REQUIRE(ld >= ch_off + C && ch_off >= 0, "cfirst_to_grid: ld %d < ch_off %d + C %d", ld, ch_off, C);
  FNSSL_REQUIRE(nf <= 65535 && nb <= 65535, "cfirst_to_grid: grid too large");
  dim3 grid((nt + 31) / 32, nf, nb);
  const size_t smem = (size_t)C * 33 * sizeof(float);
  FNSSL_REQUIRE(smem <= 200 * 1024, "cfirst_to_grid: too many channels (%d)", C);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FNSSL_F32) {
    FNSSL_CUDA(cudaFuncSetAttribute(cfirst_to_grid_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfirst_to_grid_kernel<float><<<grid, 128, smem, st>>>(src, C, nf, nt, (float*)dst, ld, ch_off);
  } else if (dtype == FNSSL_F16) {
    FNSSL_CUDA(cudaFuncSetAttribute(cfirst_to_grid_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfirst_to_grid_kernel<__half><<<grid, 128, smem, st>>>(src, C, nf, nt, (__half*)dst, ld, ch_off);
  } else {
    FNSSL_FAIL("cfirst_to_grid: bad dtype %d", dtype);
  }
  FNSSL_LAUNCH_CHECK("cfirst_to_grid_kernel");
  return 0;
}

int fnssl_grid_to_cfirst(const void* src, int dtype, int ld, int ch_off, int nb, int C, int nf, int nt, float* dst,
                         void* stream) {
  FNSSL_REQUIRE(src && dst && nb > 0 && C > 0 && nf > 0 && nt > 0, "grid_to_cfirst: bad arguments");
  FNSSL_REQUIRE(ld >= ch_off + C && ch_off >= 0, "grid_to_cfirst: ld %d < ch_off %d + C %d", ld, ch_off, C);
  FNSSL_REQUIRE(nf <= 65535 && nb <= 65535, "grid_to_cfirst: grid too large");
  dim3 grid((nt + 31) / 32, nf, nb);
  const size_t smem = (size_t)C * 33 * sizeof(float);
  FNSSL_REQUIRE(smem <= 200 * 1024, "grid_to_cfirst: too many channels (%d)", C);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FNSSL_F32) {
    FNSSL_CUDA(cudaFuncSetAttribute(grid_to_cfirst_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    grid_to_cfirst_kernel<float><<<grid, 128, smem, st>>>((const float*)src, ld, ch_off, C, nf, nt, dst);
  } else if (dtype == FNSSL_F16) {
    FNSSL_CUDA(cudaFuncSetAttribute(grid_to_cfirst_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    grid_to_cfirst_kernel<__half><<<grid, 128, smem, st>>>((const __half*)src, ld, ch_off, C, nf, nt, dst);
  } else {
    FNSSL_FAIL("grid_to_cfirst: bad dtype %d", dtype);
  }
  FNSSL_LAUNCH_CHECK("grid_to_cfirst_kernel");
  return 0;
}

int fnssl_grid_copy(const void* src, int src_dtype, int src_ld, int src_off, void* dst, int dst_dtype, int dst_ld,
                    int dst_off, int64_t npos, int C, void* stream) {
  FNSSL_REQUIRE(src && dst && npos > 0 && C > 0, "grid_copy: bad arguments");
  FNSSL_REQUIRE(src_ld >= src_off + C && dst_ld >= dst_off + C, "grid_copy: channel window out of range");
  const int64_t total = npos * C;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  cudaStream_t st = (cudaStream_t)stream;
  {
    const size_t esz = src_dtype == FNSSL_F16 ? 2 : 4;
    const size_t bytes = (size_t)npos * C * esz;
    if (src_dtype == dst_dtype && src_ld == C && dst_ld == C && src_off == 0 && dst_off == 0 && bytes % 16 == 0 &&
        ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
      const int64_t nvec = (int64_t)(bytes / 16);
      const int vblocks = (int)((nvec + 1023) / 1024 < 148 * 8 ? (nvec + 1023) / 1024 : 148 * 8);
      grid_copy_vec_kernel<<<vblocks > 0 ? vblocks : 1, 256, 0, st>>>((const uint4*)src, (uint4*)dst, nvec);
      FNSSL_LAUNCH_CHECK("grid_copy_vec_kernel");
      return 0;
    }
  }
#define FNSSL_GC(TS, TD) \
  grid_copy_kernel<TS, TD><<<blocks, 256, 0, st>>>((const TS*)src, src_ld, src_off, (TD*)dst, dst_ld, dst_off, npos, C)
  if (src_dtype == FNSSL_F32 && dst_dtype == FNSSL_F32) FNSSL_GC(float, float);
  else if (src_dtype == FNSSL_F32 && dst_dtype == FNSSL_F16) FNSSL_GC(float, __half);
  else if (src_dtype == FNSSL_F16 && dst_dtype == FNSSL_F32) FNSSL_GC(__half, float);
  else if (src_dtype == FNSSL_F16 && dst_dtype == FNSSL_F16) FNSSL_GC(__half, __half);
  else FNSSL_FAIL("grid_copy: bad dtype %d/%d", src_dtype, dst_dtype);
#undef FNSSL_GC
  FNSSL_LAUNCH_CHECK("grid_copy_kernel");
  return 0;
}

int fnssl_grid_add(const void* a, const void* b, void* dst, int dtype, int64_t n, void* stream) {
  FNSSL_REQUIRE(a && b && dst && n > 0, "grid_add: bad arguments");
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  if (dtype == FNSSL_F32) grid_add_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)a, (const float*)b, (float*)dst, n);
  else if (dtype == FNSSL_F16) grid_add_kernel<__half><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)a, (const __half*)b, (__half*)dst, n);
  else FNSSL_FAIL("grid_add: bad dtype %d", dtype);
  FNSSL_LAUNCH_CHECK("grid_add_kernel");
  return 0;
}

}  // extern "C"
