// tcgen05 implicit-GEMM version of IPDnet's CausCnnBlock (IPDnet/FixedAarryIPDnet.py:42-73) for fp16 grids.
//
//   conv1 3x3 (pad (1,2), no bias) -> ReLU -> crop 2 -> AvgPool(1,3)       (264 -> 128 channels, 86 % of the block's MACs)
//   conv2 3x3                      -> ReLU -> crop 2 -> AvgPool(1,4)       (128 -> 128)
//   conv3 3x3 -> crop -> tanh                                               (128 -> 2(M-1)S, tiny: CUDA cores, conv_simt.cu)
//
// "pad 2 then crop 2" along time makes the conv causal: out[t,f,o] = sum_{kf,kt,c} W[o,c,kf,kt] in[t+kt-2, f+kf-1, c].
// One CTA computes, for one (utterance, pooled frame, 128-bin tile), the POOL consecutive conv frames of the pooling
// window as POOL accumulators [128 bins x 128 channels] in TMEM:
//   * A operand tiles [128 bins x 64 channels] are plain TMA boxes of the channels-last input grid at the shifted
//     coordinates (f0 + kf - 1, t + kt - 2): out-of-range coordinates are zero-filled by TMA, which IS the zero padding;
//   * for each (kf, channel slab) the POOL+2 input frames and the 3 kt weight slabs are streamed once through a
//     12-slot TMA ring and reused by all (output frame, kt) pairs: 3*POOL MMAs groups per 3+POOL+2 tiles;
//   * epilogue: ReLU, average over the window, fp16 store -- the intermediate is written once, already pooled.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace fnssl {
namespace convtc {

constexpr int kThreads = 320;          // producer warp, MMA warp, 8 epilogue warps
constexpr int kSlab = 128 * 128;       // [128 rows x 64] fp16 tile (A or W)
constexpr int kSlots = 12;
constexpr int kMaxSlabs = 6;
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // M = N = 128

struct Params {
  int nslab;                        // channel slabs of the input
  int src[kMaxSlabs], k0[kMaxSlabs], nk16[kMaxSlabs];
  int nt_in, nt_out, nf, ftiles;
  __half* out;                      // grid (nb, nt_out, nf, 128) fp16
  int* error_flag;
};

// (O=128, C, 3, 3) f32 -> [tap = kf*3+kt][slab][o][64] fp16, zero beyond the real channels of a slab
__global__ void repack_weights_kernel(const float* __restrict__ w, int C, int c0_real, int c0_slabs, int c1_real, int nslab,
                                      __half* __restrict__ wp) {
  const int total = 9 * nslab * 128 * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kk = i & 63;
    const int o = (i >> 6) & 127;
    const int s = (i >> 13) % nslab;
    const int tap = i / (nslab * 8192);
    int c = -1;
    if (s < c0_slabs) { const int cc = s * 64 + kk; if (cc < c0_real) c = cc; }
    else { const int cc = (s - c0_slabs) * 64 + kk; if (cc < c1_real) c = c0_real + cc; }
    wp[i] = __float2half_rn(c >= 0 ? w[((size_t)o * C + c) * 9 + tap] : 0.0f);
  }
}

template <int POOL>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_src0, const __grid_constant__ CUtensorMap map_src1,
                  const __grid_constant__ CUtensorMap map_w, const Params p) {
  constexpr int NA = POOL + 2;           // input frames feeding one pooling window
  constexpr int NT = 3 + NA;             // tiles per (kf, slab) iteration: 3 weight slabs + NA input tiles
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) unsigned long long bars[2 * kSlots + 1];
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ft = blockIdx.x % p.ftiles;
  const int to = (blockIdx.x / p.ftiles) % p.nt_out;
  const int b = blockIdx.x / (p.ftiles * p.nt_out);
  const int f0 = ft * 128;
  const int nslab = p.nslab;
  const int niter = 3 * nslab;

  const uint32_t ring = (smem_addr(smem_dyn) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_addr(bars);
  auto FULL = [&](int i) { return bar0 + 8u * i; };
  auto EMPTY = [&](int i) { return bar0 + 8u * (kSlots + i); };
  const uint32_t ACC_FULL = bar0 + 8u * (2 * kSlots);

  if (tid == 0) {
    for (int i = 0; i < kSlots; ++i) { mbar_init(FULL(i), 1); mbar_init(EMPTY(i), 1); }
    mbar_init(ACC_FULL, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&map_src0); prefetch_tmap(&map_src1); prefetch_tmap(&map_w); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_base_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;   // accumulator i (conv frame POOL*to + i): columns [128 i, 128 i + 128)

  if (warp == 0) {
    if (lane == 0) {
      int n = 0;
      for (int it = 0; it < niter; ++it) {
        const int kf = it / nslab, s = it - kf * nslab;
        for (int j = 0; j < NT; ++j, ++n) {
          const int slot = n % kSlots, use = n / kSlots;
          if (use > 0) mbar_wait(EMPTY(slot), (uint32_t)((use - 1) & 1), p.error_flag, 400 + slot);
          mbar_expect_tx(FULL(slot), kSlab);
          const uint32_t dst = ring + (uint32_t)slot * kSlab;
          if (j < 3) {   // weight slab of tap (kf, kt = j)
            tma_load_2d(dst, &map_w, FULL(slot), 0, ((kf * 3 + j) * nslab + s) * 128);
          } else {       // input frame a = j - 3 of the window, bins shifted by kf - 1 (TMA zero-fills the padding)
            const CUtensorMap* m = p.src[s] ? &map_src1 : &map_src0;
            tma_load_4d(dst, m, FULL(slot), p.k0[s], f0 + kf - 1, POOL * to - 2 + (j - 3), b);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int n = 0;
      uint32_t started = 0;     // accumulators that already hold a partial sum
      for (int it = 0; it < niter; ++it) {
        const int s = it % nslab;
        const int nk16 = p.nk16[s];
        uint32_t wdesc_lo[3];
        int wslot[3];
        for (int j = 0; j < 3; ++j, ++n) {
          const int slot = n % kSlots;
          mbar_wait(FULL(slot), (uint32_t)((n / kSlots) & 1), p.error_flag, 420 + slot);
          wslot[j] = slot;
          wdesc_lo[j] = ring + (uint32_t)slot * kSlab;
        }
        for (int a = 0; a < NA; ++a, ++n) {
          const int slot = n % kSlots;
          mbar_wait(FULL(slot), (uint32_t)((n / kSlots) & 1), p.error_flag, 440 + slot);
          tc_fence_after();
          const uint64_t a_desc = make_sw128_desc(ring + (uint32_t)slot * kSlab);
#pragma unroll
          for (int kt = 0; kt < 3; ++kt) {
            const int i = a - kt;            // conv frame of the window fed by (input frame a, tap kt)
            if (i < 0 || i >= POOL) continue;
            const uint64_t b_desc = make_sw128_desc(wdesc_lo[kt]);
            for (int k = 0; k < nk16; ++k) {
              umma_f16(tmem + (uint32_t)i * 128u, a_desc + 2u * k, b_desc + 2u * k, kIdesc, (started >> i) & 1u);
              started |= 1u << i;
            }
          }
          umma_commit(EMPTY(slot));
        }
        for (int j = 0; j < 3; ++j) umma_commit(EMPTY(wslot[j]));
      }
      umma_commit(ACC_FULL);
    }
    __syncwarp();
  } else {
    // epilogue: ReLU, mean over the POOL conv frames, fp16 store (thread = bin, 64 of the 128 output channels)
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int f = f0 + q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    mbar_wait(ACC_FULL, 0, p.error_flag, 460);
    tc_fence_after();
    __half* dst = p.out + (((size_t)b * p.nt_out + to) * p.nf + f) * 128 + half * 64;
#pragma unroll 1
    for (int cb = 0; cb < 8; ++cb) {
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
#pragma unroll
      for (int i = 0; i < POOL; ++i) {
        float v[8];
        tmem_ld8(tmem + lane_off + (uint32_t)(i * 128 + half * 64 + cb * 8), v);
        tmem_wait_ld();
        tmem_ld_dep(v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += fmaxf(v[e], 0.0f);
      }
      if (f < p.nf) {
        const float sc = 1.0f / POOL;
        __half2 h0 = __floats2half2_rn(acc[0] * sc, acc[1] * sc), h1 = __floats2half2_rn(acc[2] * sc, acc[3] * sc);
        __half2 h2 = __floats2half2_rn(acc[4] * sc, acc[5] * sc), h3 = __floats2half2_rn(acc[6] * sc, acc[7] * sc);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst + cb * 8) = pk;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

static int make_w_map(CUtensorMap* m, const void* wp, int rows) {
  // [rows][64] fp16, box {64, 128}
  void* enc_ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  FNSSL_REQUIRE(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &enc_ptr, cudaEnableDefault, &q) == cudaSuccess &&
                    q == cudaDriverEntryPointSuccess,
                "causcnn(tcgen05): cuTensorMapEncodeTiled unavailable");
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(enc_ptr);
  const uint64_t dims[2] = {64, (uint64_t)rows};
  const uint64_t str[1] = {128};
  const uint32_t box[2] = {64, 128};
  const uint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(wp), dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FNSSL_REQUIRE(r == CUDA_SUCCESS, "causcnn(tcgen05): weight tensor map failed (%d)", (int)r);
  return 0;
}

template <int POOL>
static int launch_conv(const void* src0, int c0, int ld0, const void* src1, int c1, int ld1, int nb, int nt_in, int nf,
                       const __half* wp, __half* out, cudaStream_t st) {
  Params p{};
  int ns = 0;
  for (int src = 0; src < 2; ++src) {
    const int c = src ? c1 : c0;
    for (int k0 = 0; k0 < c; k0 += 64) {
      p.src[ns] = src; p.k0[ns] = k0;
      p.nk16[ns] = (((c - k0) < 64 ? (c - k0) : 64) + 15) / 16;
      ++ns;
    }
  }
  p.nslab = ns;
  p.nt_in = nt_in; p.nt_out = nt_in / POOL; p.nf = nf; p.ftiles = (nf + 127) / 128;
  p.out = out;
  p.error_flag = tc_wait_timeout_enabled() ? tc_error_flag() : nullptr;
  CUtensorMap m0, m1, mw;
  if (make_grid_map(&m0, src0, c0, ld0, nb, nt_in, nf, FNSSL_ALONG_TIME, 128)) return 1;
  if (c1 > 0) { if (make_grid_map(&m1, src1, c1, ld1, nb, nt_in, nf, FNSSL_ALONG_TIME, 128)) return 1; }
  else m1 = m0;
  if (make_w_map(&mw, wp, 9 * ns * 128)) return 1;
  const size_t smem = (size_t)kSlots * kSlab + 1024;
  FNSSL_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (long long)nb * p.nt_out * p.ftiles;
  FNSSL_REQUIRE(blocks > 0 && blocks < (1ll << 31), "causcnn(tcgen05): bad grid");
  conv3x3_tc_kernel<POOL><<<(unsigned)blocks, kThreads, smem, st>>>(m0, m1, mw, p);
  FNSSL_LAUNCH_CHECK("conv3x3_tc_kernel");
  return 0;
}

}  // namespace convtc

static size_t align256c(size_t x) { return (x + 255) & ~(size_t)255; }

bool causcnn_tc_supports(int c0, int c1, int hid, int dtype) {
  const int ns = (c0 + 63) / 64 + (c1 + 63) / 64;
  return dtype == FNSSL_F16 && hid == 128 && c0 > 0 && c0 % 16 == 0 && c1 % 16 == 0 && ns <= convtc::kMaxSlabs;
}

size_t causcnn_tc_workspace_bytes(int nb, int nt, int nf, int c0, int c1, int cout) {
  const size_t nt1 = nt / 3, nt2 = nt1 / 4;
  const int ns1 = (c0 + 63) / 64 + (c1 + 63) / 64;
  return align256c((size_t)nb * nt1 * nf * 128 * 2) + align256c((size_t)nb * nt2 * nf * 128 * 2) +
         align256c((size_t)9 * ns1 * 128 * 64 * 2) + align256c((size_t)9 * 2 * 128 * 64 * 2) + align256c((size_t)9 * 128 * cout * 4);
}

// c0/c1: PADDED channel counts of the two fp16 sources (multiples of 16); c0_real/c1_real: channels the weights cover
int causcnn_forward_tc(const void* src0, int c0, int c0_real, int ld0, const void* src1, int c1, int c1_real, int ld1, int nb, int nt,
                       int nf, const float* w1, const float* w2, void* work, __half** y2_out, float** w3r_out, cudaStream_t st) {
  using namespace convtc;
  const int nt1 = nt / 3, nt2 = nt1 / 4;
  const int ns1 = (c0 + 63) / 64 + (c1 + 63) / 64;
  char* wsp = (char*)work;
  __half* y1 = (__half*)wsp; wsp += align256c((size_t)nb * nt1 * nf * 128 * 2);
  __half* y2 = (__half*)wsp; wsp += align256c((size_t)nb * nt2 * nf * 128 * 2);
  __half* wp1 = (__half*)wsp; wsp += align256c((size_t)9 * ns1 * 128 * 64 * 2);
  __half* wp2 = (__half*)wsp; wsp += align256c((size_t)9 * 2 * 128 * 64 * 2);
  *w3r_out = (float*)wsp;
  *y2_out = y2;
  repack_weights_kernel<<<128, 256, 0, st>>>(w1, c0_real + c1_real, c0_real, (c0 + 63) / 64, c1_real, ns1, wp1);
  repack_weights_kernel<<<64, 256, 0, st>>>(w2, 128, 128, 2, 0, 2, wp2);
  FNSSL_LAUNCH_CHECK("repack_weights_kernel");
  if (launch_conv<3>(src0, c0, ld0, src1, c1, ld1, nb, nt, nf, wp1, y1, st)) return 1;
  if (launch_conv<4>(y1, 128, 128, nullptr, 0, 0, nb, nt1, nf, wp2, y2, st)) return 1;
  return 0;
}

}  // namespace fnssl
