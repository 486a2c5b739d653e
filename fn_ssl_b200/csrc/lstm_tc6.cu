// CTA-pair (tcgen05 cta_group::2, M = 128) cluster LSTM kernel for H = 256 single-source layers -- the narrow-band uni-LSTM of
// the ONLINE models (FN_SSL(is_online=True) blocks 2 and 3, FN-SSL/Lightning/Model.py:22-29,46).
//
// Why.  lstm_tc4.cu runs H = 256 as clusters of 8 CTAs x 32 units; the h state of all 256 units only fits next to the weights
// for 64-row tiles, so every MMA there is an M = 64 instruction -- half the tensor rate -- and the tensor pipe is what that
// layer waits for (ncu: 76 % active, profiles/r2_ncu_full_v1_lstm_tc4_online_h256_summary.txt).  Here the cluster's 8 CTAs
// are 4 PAIRS; a pair (ranks 2p, 2p+1) issues tcgen05.mma.cta_group::2 with M = 128: each CTA still holds 64 rows, but the pair's
// MMA runs at the full rate (max(M,128) N / (256 cta_group) cycles: 32 per K = 16 step instead of 64).  A pair owns 64 hidden
// units (two 32-unit chunks = unit halves uh), each CTA keeps HALF of the B operand (64 of a chunk's 128 gate columns) --
// the same 128 KB of weights per CTA as before -- and its own rows' h (64 rows x 256 units, two chains: 64 KB).  A CTA exchanges
// its [64 x 32] tile with the 3 CTAs that hold the same rows in the other pairs (12 KB per half-slot instead of 28 KB per slot).
//
// TMEM layout of an M = 128 cta_group::2 accumulator (per CTA, N = 128): row r (0..63), column n -> lane r + 64 (n / 64), TMEM
// column n % 64.  The B rows are therefore ordered [half][gate][16 units] so that lanes 0-63 hold all four gates of units 0-15 of
// the chunk and lanes 64-127 those of units 16-31 (a thread needs i, f, g, o of its elements): CTA j of the pair supplies B rows
// 64 j .. 64 j + 63 = gates x units 16 j .. 16 j + 15, loaded as four 16-row TMA boxes per K slab out of the ordinary packed buffer.
//
// Half-slot n = 4 t + 2 chain + uh as in lstm_tc5.cu: one accumulator (64 TMEM columns), the epilogue of 64 rows x 32 units (thread =
// row x 4 units), one [64 x 32] h tile.  Roles as lstm_tc5.cu: warp 0 TMA producer (both CTAs), warp 1 leader: h-part issuer /
// other CTA: relay of "weights loaded" and "h tiles complete", warps 2..17 epilogue, warp 18 leader: x-part issuer / other CTA:
// relay of "accumulator drained", warp 19 publisher.
// A second source of <= 16 channels (the raw-feature skip of IPDnet / of FN-SSL's first narrow-band layer) is one K = 16 step with
// a 2 KB weight half per unit half and a two-entry ring of [64 x 16] slabs of its own (32B swizzle), as lstm_tc4.cu's NARROW mode.
// Not supported here (-> lstm_tc4.cu): a wider second source, carried (h, c) state, outputs that cannot go through TMA.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

#ifndef TC6_PUSH_TILE
#define TC6_PUSH_TILE 0      // 1: one DSMEM bulk copy of the whole [64 x 32] tile per peer instead of one per row half (measured: a draw, off)
#endif

namespace fnssl {
namespace tc6 {

constexpr int kThreads = 640;
constexpr int kXWarp = 18, kPubWarp = 19;
constexpr int kEpiWarps = 16;
constexpr int kSlabK = 64;
constexpr int kRows = 64;                  // rows per CTA per chain
constexpr int kChainRows = 128;            // rows per chain (pair: 2 x 64)
constexpr int kWHalf = 64 * 128;           // [64 B rows x 64 K] fp16: this CTA's half of a chunk's [128 x 64] weight slab
constexpr int kXSlab = kRows * 128;        // [64 rows x 64] fp16
constexpr int kHTile = kRows * 64;         // [64 rows x 32 units] fp16 (64B swizzle)
constexpr int kXSmall = kRows * 32;        // [64 rows x 16 channels] fp16 (32B swizzle): a slab of the narrow second source
constexpr int kWSmall = 64 * 32;           // [64 B rows x 16 K] fp16 (32B swizzle): this CTA's half of a chunk's narrow weight slab
constexpr int kChunkN = 128;
constexpr int kMaxXSlabs = 4, kMaxXStages = 6, kAccBufs = 4;      // accumulators: one PAIR (both unit halves) per chain
constexpr int kAccCols = 64;               // TMEM columns of one accumulator (N / 2)
constexpr int kSmemLimit = 232448;
constexpr int B_WFULL = 0, B_WMATE = 1, B_XFULL = 2, B_XEMPTY = B_XFULL + kMaxXStages, B_ACCFULL = B_XEMPTY + kMaxXStages,
              B_ACCEMPTY = B_ACCFULL + kAccBufs, B_XPDONE = B_ACCEMPTY + kAccBufs, B_HFULL = B_XPDONE + kAccBufs,
              B_HMATE = B_HFULL + 2, B_HFREE = B_HMATE + 2, B_HREADY = B_HFREE + 2, B_ACCDRAIN = B_HREADY + 8, B_X2FULL = B_ACCDRAIN + kAccBufs,
              B_X2EMPTY = B_X2FULL + 2, kNumBars = B_X2EMPTY + 2;
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kChunkN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32, M = 128, N = 128

struct Params {
  int nxs;                         // 64-channel slabs of the first source
  int small;                       // 1: a second source of <= 16 channels (one K = 16 step, own two-entry ring, 32B swizzle)
  uint32_t xs_nkpack;              // nibble j: K=16 steps of slab j (1..4)
  int xstages;
  int steps, axis, nf, nt;
  long long rows;
  int chains_per_b, nchains;
  const float* bias;               // [dirs][4H], accumulator column order [chunk][gate][unit]
  int out0_off;
  int tma_out;                     // bit 0: out0 tile stores; bit 1: in-place reduce-add onto out1 == addend; bit 2: out1 = second copy of h
  int* error_flag;
  int debug;
};

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lstm_cell_tanh(float gi, float gf, float gg, float go, float& c) {
  const float si = fmaf(0.5f, tanh_approx(0.5f * gi), 0.5f);
  const float sf = fmaf(0.5f, tanh_approx(0.5f * gf), 0.5f);
  const float so = fmaf(0.5f, tanh_approx(0.5f * go), 0.5f);
  const float cn = fmaf(sf, c, si * tanh_approx(gg));
  c = cn;
  return so * tanh_approx(cn);
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void st_shared_v2(uint32_t saddr, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(saddr), "r"(x), "r"(y) : "memory");
}

template <int H>      // 256: cluster of 8 = 4 pairs; 128: cluster of 4 = 2 pairs (mid-size layers, see lstm_tc6_wants)
__global__ void __launch_bounds__(kThreads, 1)
lstm_tc6_kernel(const __grid_constant__ CUtensorMap map_src0, const __grid_constant__ CUtensorMap map_src1, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_out0, const __grid_constant__ CUtensorMap map_out1,
                const Params p) {
  constexpr int kCluster = H / 32;           // CTAs per cluster
  constexpr int kPairs = kCluster / 2;
  constexpr int kNH = H / 32;                // chunk tiles of h per chain
  constexpr int kNHS = H / kSlabK;           // K slabs of W_h
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) unsigned long long bars[kNumBars];
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)(rank >> 1), jh = (int)(rank & 1);      // pair -> hidden units [64 pair, +64); jh -> row half of a chain
  const bool leader = jh == 0;
  const uint32_t leader_rank = rank & ~1u;
  const uint16_t pair_mask = (uint16_t)(3u << (2 * pair));
  const int cluster_tile = blockIdx.x / kCluster;
  const int nxs = p.nxs, XS = p.xstages, L = p.steps;
  const int nhalf = 4 * L;

  const uint32_t dyn0 = (smem_addr(smem_dyn) + 1023u) & ~1023u;
  const int nslabs = nxs + kNHS;                                          // RESIDENT 64-wide slabs (the packed buffer has p.small more)
  const bool small1 = p.small != 0;
  const uint32_t w_base = dyn0;                                           // [uh][slab: nxs x, then 4 h] halves of 8 KB
  const uint32_t w2_base = w_base + (uint32_t)(2 * nslabs) * kWHalf;      // [uh] halves of the narrow weight slab (2 KB each)
  const uint32_t hs_base = w2_base + (small1 ? 2u * kWSmall : 0u);        // h operand: [chain][chunk 0..7] tiles of 4 KB
  const uint32_t xr_base = hs_base + (uint32_t)(2 * kNH) * kHTile;        // x ring
  const uint32_t x2_base = xr_base + (uint32_t)XS * kXSlab;               // two-entry ring of the narrow source
  const uint32_t bias_base = x2_base + (small1 ? 2u * kXSmall : 0u);      // 256 floats: [uh][gate][unit]
  float* bias_s = reinterpret_cast<float*>(smem_dyn + (bias_base - smem_addr(smem_dyn)));

  const uint32_t bar0 = smem_addr(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  if (tid == 0) {
    mbar_init(BAR(B_WFULL), 1);
    mbar_init(BAR(B_WMATE), 1);
    for (int i = 0; i < kMaxXStages; ++i) { mbar_init(BAR(B_XFULL + i), 1); mbar_init(BAR(B_XEMPTY + i), 1); }
    for (int i = 0; i < kAccBufs; ++i) {
      mbar_init(BAR(B_ACCFULL + i), 1);
      mbar_init(BAR(B_ACCEMPTY + i), kEpiWarps + 1);     // the leader's epilogue warps + the other CTA's relayed "drained"
      mbar_init(BAR(B_ACCDRAIN + i), kEpiWarps);
      mbar_init(BAR(B_XPDONE + i), 1);
    }
    for (int c = 0; c < 2; ++c) {
      mbar_init(BAR(B_HFULL + c), 5);     // expect_tx arrive + 2 unit halves x 2 row halves of the CTA's own tiles (+ 24 KB of tx from 3 peers)
      mbar_init(BAR(B_HMATE + c), 1);
      mbar_init(BAR(B_HFREE + c), kPairs + 1);     // the pair leaders' commits + the publisher ("stores drained")
    }
    for (int i = 0; i < 8; ++i) mbar_init(BAR(B_HREADY + i), 8);     // [(chain, uh)][row half]: that row half's 8 epilogue warps
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_X2FULL + i), 1); mbar_init(BAR(B_X2EMPTY + i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&map_src0); prefetch_tmap(&map_w); if (small1) { prefetch_tmap(&map_src1); prefetch_tmap(&map_w2); } }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_addr(&tmem_base_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 2 * kChunkN; i += kThreads) bias_s[i] = p.bias[dir * 4 * H + (2 * pair) * kChunkN + i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  // (columns [0,64) are unused: the cell state lives in the epilogue threads' registers)
  const uint32_t tmem_acc = tmem;           // gate accumulators: kAccBufs buffers x 64 columns (all 256 allocated columns)

  const bool along_f = p.axis == FNSSL_ALONG_FREQ;
  int cb_[2], cr0_[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int g = 2 * cluster_tile + c;                    // chain tile index
    if (along_f) {
      cb_[c] = 0;
      cr0_[c] = g * kChainRows + jh * kRows;
    } else {
      cb_[c] = g / p.chains_per_b;
      cr0_[c] = (g % p.chains_per_b) * kChainRows + jh * kRows;
    }
    if (g >= p.nchains) { cb_[c] = 0; cr0_[c] = 1 << 28; }    // an absent chain: every TMA access is out of range (zero fill / clipped)
  }
  const int cbA = cb_[0], cbB = cb_[1], crA = cr0_[0], crB = cr0_[1];
#define CB(c) ((c) ? cbB : cbA)
#define CR0(c) ((c) ? crB : crA)

  if (warp == 0) {
    // ============================== TMA producer (both CTAs) ==============================
    if (elect_one()) {
      // resident weights: for each unit half, gate g, the 16 rows of units 16 jh .. 16 jh + 15 of chunk 2 pair + uh, every K slab
      // (packed column blocks: nxs slabs of the first source, the narrow slab if any, 4 h slabs)
      mbar_expect_tx(BAR(B_WFULL), (uint32_t)(2 * nslabs) * kWHalf + (small1 ? 2u * kWSmall : 0u));
      for (int uh = 0; uh < 2; ++uh) {
        for (int j = 0; j < nslabs; ++j)
#pragma unroll
          for (int g = 0; g < 4; ++g)
            tma_load_2d(w_base + (uint32_t)(uh * nslabs + j) * kWHalf + (uint32_t)g * 2048u, &map_w, BAR(B_WFULL),
                        (j < nxs ? j : j + p.small) * kSlabK, (dir * kNH + 2 * pair + uh) * kChunkN + g * 32 + 16 * jh);
        if (small1)
#pragma unroll
          for (int g = 0; g < 4; ++g)
            tma_load_2d(w2_base + (uint32_t)uh * kWSmall + (uint32_t)g * 512u, &map_w2, BAR(B_WFULL), nxs * kSlabK,
                        (dir * kNH + 2 * pair + uh) * kChunkN + g * 32 + 16 * jh);
      }
      const uint32_t lead_xfull0 = mapa_shared(BAR(B_XFULL), leader_rank);
      const uint32_t lead_x2full0 = mapa_shared(BAR(B_X2FULL), leader_rank);
      int stage = 0;
      uint32_t phase = 0;
      bool wrapped = false;
      for (int n2 = 0; n2 < 2 * L; ++n2) {        // one pass per (step, chain): its slabs serve both unit halves
        const int t = n2 >> 1, c = n2 & 1;
        const int s = dir ? (L - 1 - t) : t;
        for (int j = 0; j < nxs; ++j) {
          if (wrapped) mbar_wait(BAR(B_XEMPTY + stage), phase, p.error_flag, 100 + stage);
          if (leader) mbar_expect_tx(BAR(B_XFULL + stage), 2u * kXSlab);
          const uint32_t dst = xr_base + (uint32_t)stage * kXSlab;
          const uint32_t fb = lead_xfull0 + 8u * (uint32_t)stage;
          if (along_f) tma2_load_4d(dst, &map_src0, fb, j * kSlabK, s, CR0(c), 0);
          else tma2_load_4d(dst, &map_src0, fb, j * kSlabK, CR0(c), s, CB(c));
          if (++stage == XS) { stage = 0; phase ^= wrapped ? 1u : 0u; wrapped = true; }
        }
        if (small1) {      // the narrow source's slab of this pass: entry n2 & 1 of its own ring, use number n2 >> 1
          const int s2 = n2 & 1;
          if (n2 >= 2) mbar_wait(BAR(B_X2EMPTY + s2), (uint32_t)(((n2 >> 1) - 1) & 1), p.error_flag, 110 + s2);
          if (leader) mbar_expect_tx(BAR(B_X2FULL + s2), 2u * kXSmall);
          const uint32_t dst = x2_base + (uint32_t)s2 * kXSmall;
          const uint32_t fb = lead_x2full0 + 8u * (uint32_t)s2;
          if (along_f) tma2_load_4d(dst, &map_src1, fb, 0, s, CR0(c), 0);
          else tma2_load_4d(dst, &map_src1, fb, 0, CR0(c), s, CB(c));
        }
      }
    }
    __syncwarp();
  } else if (warp == kXWarp) {
    if (leader && elect_one()) {
      // ============================== x-part MMA issuer (pair leader) ==============================
      mbar_wait(BAR(B_WFULL), 0, p.error_flag, 200);
      mbar_wait(BAR(B_WMATE), 0, p.error_flag, 201);
      const uint64_t a_desc0 = make_sw128_desc(xr_base);
      const uint64_t b_desc0 = make_sw128_desc(w_base);
      int xstage = 0, a0 = 0;
      uint32_t xphase = 0, empty_par = 0;
      for (int n2 = 0; n2 < 2 * L; ++n2) {
        const int a1 = (a0 == kAccBufs - 1) ? 0 : a0 + 1;
        if (2 * n2 >= kAccBufs) {
          mbar_wait(BAR(B_ACCEMPTY + a0), (empty_par >> a0) & 1u, p.error_flag, 202 + a0);
          empty_par ^= 1u << a0;
        }
        if (2 * n2 + 1 >= kAccBufs) {
          mbar_wait(BAR(B_ACCEMPTY + a1), (empty_par >> a1) & 1u, p.error_flag, 202 + a1);
          empty_par ^= 1u << a1;
        }
        tc_fence_after();
        const uint32_t d0 = tmem_acc + (uint32_t)a0 * kAccCols, d1 = tmem_acc + (uint32_t)a1 * kAccCols;
        uint32_t nkp = p.xs_nkpack;
        for (int j = 0; j < nxs; ++j, nkp >>= 4) {
          mbar_wait(BAR(B_XFULL + xstage), xphase, p.error_flag, 210 + xstage);
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + (uint64_t)(xstage * (kXSlab >> 4));
          const uint64_t b0 = b_desc0 + (uint64_t)(j * (kWHalf >> 4));
          const uint64_t b1 = b_desc0 + (uint64_t)((nslabs + j) * (kWHalf >> 4));
          const uint32_t nk = nkp & 15u;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if ((uint32_t)k < nk) umma2_f16(d0, a_desc + 2u * k, b0 + 2u * k, kIdesc, (uint32_t)(j | k));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if ((uint32_t)k < nk) umma2_f16(d1, a_desc + 2u * k, b1 + 2u * k, kIdesc, (uint32_t)(j | k));
          umma2_commit_mc(BAR(B_XEMPTY + xstage), pair_mask);
          if (++xstage == XS) { xstage = 0; xphase ^= 1u; }
        }
        if (small1) {      // one K = 16 step per unit half against the 2 KB weight halves, both operands in the 32B-swizzled layout
          const int s2 = n2 & 1;
          mbar_wait(BAR(B_X2FULL + s2), (uint32_t)((n2 >> 1) & 1), p.error_flag, 216 + s2);
          tc_fence_after();
          const uint64_t a2 = make_sw32_desc(x2_base + (uint32_t)s2 * kXSmall);
          umma2_f16(d0, a2, make_sw32_desc(w2_base), kIdesc, 1u);
          umma2_f16(d1, a2, make_sw32_desc(w2_base + kWSmall), kIdesc, 1u);
          umma2_commit_mc(BAR(B_X2EMPTY + s2), pair_mask);
        }
        umma2_commit_mc(BAR(B_XPDONE + a0), (uint16_t)(1u << leader_rank));
        umma2_commit_mc(BAR(B_XPDONE + a1), (uint16_t)(1u << leader_rank));
        a0 = (a1 == kAccBufs - 1) ? 0 : a1 + 1;
      }
    } else if (!leader && elect_one()) {
      // the pair's other CTA: relay "all my epilogue warps have drained accumulator buffer a" to the leader's ACC_EMPTY barrier
      const uint32_t lead_empty0 = mapa_shared(BAR(B_ACCEMPTY), leader_rank);
      int a = 0;
      uint32_t par = 0;
      for (int n = 0; n < nhalf; ++n) {
        mbar_wait(BAR(B_ACCDRAIN + a), (par >> a) & 1u, p.error_flag, 260 + a);
        par ^= 1u << a;
        mbar_arrive_remote(lead_empty0 + 8u * (uint32_t)a);
        a = (a == kAccBufs - 1) ? 0 : a + 1;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader) {
      // ============================== h-part MMA issuer (pair leader) ==============================
      if (elect_one()) {
        mbar_wait(BAR(B_WFULL), 0, p.error_flag, 230);
        mbar_wait(BAR(B_WMATE), 0, p.error_flag, 231);
        const uint64_t h_desc0 = make_sw64_desc(hs_base);
        const uint64_t wh_desc0 = make_sw128_desc(w_base + (uint32_t)nxs * kWHalf);      // h slabs follow the x slabs (uh = 0)
        int a = 0;
        uint32_t xp_par = 0;
        for (int n = 0; n < nhalf; ++n) {
          const int t = n >> 2, c = (n >> 1) & 1, uh = n & 1;
          mbar_wait(BAR(B_XPDONE + a), (xp_par >> a) & 1u, p.error_flag, 240 + a);
          xp_par ^= 1u << a;
          if (t > 0) {
            if (uh == 0) {
              mbar_expect_tx(BAR(B_HFULL + c), (uint32_t)(2 * (kPairs - 1)) * kHTile);      // the peers' two tiles each
              mbar_wait_cluster(BAR(B_HFULL + c), (uint32_t)((t - 1) & 1), p.error_flag, 220 + c);
              mbar_wait_cluster(BAR(B_HMATE + c), (uint32_t)((t - 1) & 1), p.error_flag, 222 + c);
            }
            tc_fence_after();
            const uint32_t d_tmem = tmem_acc + (uint32_t)a * kAccCols;
            const uint64_t a_chain = h_desc0 + (uint64_t)(c * kNH * (kHTile >> 4));
            const uint64_t b_uh = wh_desc0 + (uint64_t)(uh * nslabs * (kWHalf >> 4));
#pragma unroll
            for (int kc = 0; kc < kNH; ++kc) {   // K = 32 units of chunk kc: two K=16 steps; W columns inside 128B-swizzled slab kc/2
              const uint64_t a_desc = a_chain + (uint64_t)(kc * (kHTile >> 4));
              const uint64_t b_desc = b_uh + (uint64_t)((kc >> 1) * (kWHalf >> 4) + 4 * (kc & 1));
#pragma unroll
              for (int k = 0; k < 2; ++k) umma2_f16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, kIdesc, 1u);
            }
            if (uh == 1) umma2_commit_mc(BAR(B_HFREE + c), (uint16_t)((1u << kCluster) - 1u));
          } else {
            tc_fence_after();
          }
          umma2_commit_mc(BAR(B_ACCFULL + a), pair_mask);
          a = (a == kAccBufs - 1) ? 0 : a + 1;
        }
      }
    } else {
      // ============================== relay (the pair's other CTA) ==============================
      if (elect_one()) {
        mbar_wait(BAR(B_WFULL), 0, p.error_flag, 250);
        mbar_arrive_remote(mapa_shared(BAR(B_WMATE), leader_rank));
        const uint32_t lead_hmate0 = mapa_shared(BAR(B_HMATE), leader_rank);
        for (int t = 1; t < L; ++t) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            mbar_expect_tx(BAR(B_HFULL + c), (uint32_t)(2 * (kPairs - 1)) * kHTile);
            mbar_wait_cluster(BAR(B_HFULL + c), (uint32_t)((t - 1) & 1), p.error_flag, 252 + c);
            mbar_arrive_remote(lead_hmate0 + 8u * (uint32_t)c);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kPubWarp) {
    // ============================== publisher ==============================
    if (elect_one()) {
      uint32_t peer_hs[kPairs - 1], peer_hfull0[kPairs - 1];
#pragma unroll
      for (int d = 0; d < kPairs - 1; ++d) {
        const uint32_t pr = (uint32_t)(2 * ((pair + 1 + d) % kPairs) + jh);      // the CTAs that hold the same rows in the other pairs
        peer_hs[d] = mapa_shared(hs_base, pr);
        peer_hfull0[d] = mapa_shared(BAR(B_HFULL), pr);
      }
      const bool tma_any = p.tma_out != 0;
      for (int n = 0; n < nhalf; ++n) {
        const int t = n >> 2, c = (n >> 1) & 1, uh = n & 1;
        const int s = dir ? (L - 1 - t) : t;
        const bool push = t + 1 < L;
        const int kc = 2 * pair + uh;
        const uint32_t tile_off = (uint32_t)(c * kNH + kc) * kHTile;
        if (push || tma_any) {
#if TC6_PUSH_TILE
#pragma unroll
          for (int rh = 0; rh < 2; ++rh)
            mbar_wait(BAR(B_HREADY + (c * 2 + uh) * 2 + rh), (uint32_t)(t & 1), p.error_flag, 400 + (c * 2 + uh) * 2 + rh);
          if (push) {
#pragma unroll
            for (int d = 0; d < kPairs - 1; ++d)
              bulk_copy_s2c(peer_hs[d] + tile_off, hs_base + tile_off, (uint32_t)kHTile, peer_hfull0[d] + 8u * (uint32_t)c);
            mbar_arrive(BAR(B_HFULL + c));
            mbar_arrive(BAR(B_HFULL + c));
          }
#else
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) {       // rows 0-31 / 32-63 of the tile: 2 KB each
            mbar_wait(BAR(B_HREADY + (c * 2 + uh) * 2 + rh), (uint32_t)(t & 1), p.error_flag, 400 + (c * 2 + uh) * 2 + rh);
            if (push) {
              const uint32_t off = tile_off + (uint32_t)rh * 2048u;
#pragma unroll
              for (int d = 0; d < kPairs - 1; ++d) bulk_copy_s2c(peer_hs[d] + off, hs_base + off, 2048u, peer_hfull0[d] + 8u * (uint32_t)c);
              mbar_arrive(BAR(B_HFULL + c));
            }
          }
#endif
          if (tma_any) {
            const int out_c = dir * H + kc * 32;
            {      // one [64 rows x 32 channels] box per destination (a TMA store costs the issuing lane ~170 cycles)
              const uint32_t off = tile_off;
              const int r0 = CR0(c);
              if (p.tma_out & 1) {
                if (along_f) tma_store_4d(&map_out0, hs_base + off, p.out0_off + out_c, s, r0, 0);
                else tma_store_4d(&map_out0, hs_base + off, p.out0_off + out_c, r0, s, CB(c));
              }
              if (p.tma_out & 2) {
                if (along_f) tma_reduce_add_4d(&map_out1, hs_base + off, out_c, s, r0, 0);
                else tma_reduce_add_4d(&map_out1, hs_base + off, out_c, r0, s, CB(c));
              }
              if (p.tma_out & 4) {
                if (along_f) tma_store_4d(&map_out1, hs_base + off, out_c, s, r0, 0);
                else tma_store_4d(&map_out1, hs_base + off, out_c, r0, s, CB(c));
              }
            }
            bulk_commit_group();
          }
        }
        if (uh == 1 && n >= 3) {     // the stores of the OTHER chain's two half-slots (groups n-3, n-2) have read their tiles
          if (tma_any) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
          mbar_arrive(BAR(B_HFREE + (c ^ 1)));
        }
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue warps ==============================
    // lanes 0-63 of the accumulator hold units 0-15 of the chunk (all four gates), lanes 64-127 units 16-31: warp w serves TMEM
    // lane quadrant q = w & 3 -> rows 32 (q & 1) + lane, unit half-of-chunk q >> 1; the quadrant's four warps split its 16 units.
    // ONE pass per (step, chain) covers BOTH unit halves (two accumulator buffers): a half-slot of this kernel has only 4 elements
    // per thread (20 MUFU ops), so the fixed costs of a pass -- barrier wake-up, TMEM load round trip, fence, hand-off -- would
    // otherwise dominate it (measured: 2.3 k cycles per half-slot with one pass per half-slot, of which 0.64 k MUFU).
    const int q = warp & 3;
    const int widx = (warp - 2) >> 2;                      // 0..3: units 4 widx .. 4 widx + 3 of the quadrant's 16
    const int rh = q & 1, hi = q >> 1;
    const int r = rh * 32 + lane;                          // row of this thread inside the CTA's 64 rows
    const int ul = 16 * hi + 4 * widx;                     // first of the thread's 4 units inside a 32-unit chunk
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t colg = (uint32_t)(4 * widx);            // column of the thread's units inside a gate's 16 columns
    // 8-byte piece of the thread inside a [64 x 32] tile (64B swizzle: 16-byte chunk ^= (row >> 1) & 3)
    const uint32_t hpiece = (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u +
                            (uint32_t)((((uint32_t)ul >> 3) ^ (((uint32_t)r >> 1) & 3u)) << 4) + (uint32_t)(ul & 7) * 2u;
    const bool tma_any = p.tma_out != 0;
    float creg[2][2][4];                                   // cell state of the thread's (chain, unit half) x 4 units
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int uh = 0; uh < 2; ++uh)
#pragma unroll
        for (int e = 0; e < 4; ++e) creg[c][uh][e] = 0.0f;
    int a = 0;
    uint32_t full_par = 0;
#pragma unroll 1
    for (int t = 0; t < L; ++t) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int a0 = a, a1 = (a == kAccBufs - 1) ? 0 : a + 1;
        float gi[2][4], gf[2][4], gg[2][4], go[2][4];
        mbar_wait(BAR(B_ACCFULL + a0), (full_par >> a0) & 1u, p.error_flag, 300 + a0);
        full_par ^= 1u << a0;
        tc_fence_after();
        {
          const uint32_t acc = tmem_acc + (uint32_t)a0 * kAccCols + lane_off + colg;
          tmem_ld4(acc + 0 * 16, gi[0]); tmem_ld4(acc + 1 * 16, gf[0]); tmem_ld4(acc + 2 * 16, gg[0]); tmem_ld4(acc + 3 * 16, go[0]);
        }
        mbar_wait(BAR(B_ACCFULL + a1), (full_par >> a1) & 1u, p.error_flag, 300 + a1);      // committed right behind the first one
        full_par ^= 1u << a1;
        tc_fence_after();
        {
          const uint32_t acc = tmem_acc + (uint32_t)a1 * kAccCols + lane_off + colg;
          tmem_ld4(acc + 0 * 16, gi[1]); tmem_ld4(acc + 1 * 16, gf[1]); tmem_ld4(acc + 2 * 16, gg[1]); tmem_ld4(acc + 3 * 16, go[1]);
        }
        const bool will_publish = t + 1 < L || tma_any;
        const bool hfree_ok = !(will_publish && t > 0) || mbar_test_wait(BAR(B_HFREE + c), (uint32_t)((t - 1) & 1));
        tmem_wait_ld();
#pragma unroll
        for (int uh = 0; uh < 2; ++uh) { tmem_ld_dep4(gi[uh]); tmem_ld_dep4(gf[uh]); tmem_ld_dep4(gg[uh]); tmem_ld_dep4(go[uh]); }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {       // both accumulators drained (this warp)
          mbar_arrive(BAR((leader ? B_ACCEMPTY : B_ACCDRAIN) + a0));
          mbar_arrive(BAR((leader ? B_ACCEMPTY : B_ACCDRAIN) + a1));
        }
        a = (a1 == kAccBufs - 1) ? 0 : a1 + 1;
        uint32_t pk[2][2];
#pragma unroll
        for (int uh = 0; uh < 2; ++uh) {
          const float* bsp = bias_s + uh * kChunkN + ul;
          float hv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            hv[e] = lstm_cell_tanh(gi[uh][e] + bsp[e], gf[uh][e] + bsp[32 + e], gg[uh][e] + bsp[64 + e], go[uh][e] + bsp[96 + e], creg[c][uh][e]);
          __half2 h01 = __floats2half2_rn(hv[0], hv[1]), h23 = __floats2half2_rn(hv[2], hv[3]);
          pk[uh][0] = *reinterpret_cast<uint32_t*>(&h01);
          pk[uh][1] = *reinterpret_cast<uint32_t*>(&h23);
        }
        if (will_publish) {
          if (!hfree_ok) mbar_wait(BAR(B_HFREE + c), (uint32_t)((t - 1) & 1), p.error_flag, 320 + c);
          const uint32_t tile0 = hs_base + (uint32_t)(c * kNH + 2 * pair) * kHTile + hpiece;
          st_shared_v2(tile0, pk[0][0], pk[0][1]);
          st_shared_v2(tile0 + kHTile, pk[1][0], pk[1][1]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(BAR(B_HREADY + (c * 2 + 0) * 2 + rh));
            mbar_arrive(BAR(B_HREADY + (c * 2 + 1) * 2 + rh));
          }
        }
      }
    }
  }
#undef CB
#undef CR0

  bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
  }
}

struct Plan { bool ok; int xstages; int nxs; int small; size_t smem; };

static Plan make_plan(int H, int c0, int c1) {
  const int kNH = H / 32, kNHS = H / kSlabK;
  Plan pl{false, 0, 0, 0, 0};
  if (c0 % 16 || c0 <= 0 || c1 < 0 || c1 > 16 || c1 % 16) return pl;      // second source: none, or one K = 16 step
  const int nxs = (c0 + 63) / 64;
  if (nxs > kMaxXSlabs) return pl;
  const int small = c1 > 0 ? 1 : 0;
  const long fixed = 2L * (nxs + kNHS) * kWHalf + 2L * kNH * kHTile + 2 * kChunkN * 4 + 1024 + small * 2L * (kWSmall + kXSmall);
  long xs = (kSmemLimit - 1024 - fixed) / kXSlab;
  if (xs > kMaxXStages) xs = kMaxXStages;
  if (xs < 2) return pl;
  pl.ok = true; pl.xstages = (int)xs; pl.nxs = nxs; pl.small = small; pl.smem = (size_t)fixed + (size_t)xs * kXSlab;
  return pl;
}

// 2-D fp16 map over the packed weights, box = [16 rows x 64 K]: one gate's 16 units of one CTA's half of a slab
static int make_gate_weight_map(CUtensorMap* m, const void* weights, int nslabs, int nchunks_total, bool small = false) {
  static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  if (!enc) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  FNSSL_REQUIRE(enc, "lstm(tcgen05): cuTensorMapEncodeTiled is unavailable in this driver");
  const uint64_t dims[2] = {(uint64_t)nslabs * kSlabK, (uint64_t)nchunks_total * kChunkN};
  const uint64_t str[1] = {(uint64_t)nslabs * kSlabK * 2};
  const uint32_t box[2] = {small ? 16u : (uint32_t)kSlabK, 16};      // narrow slab: [16 rows x 16 K], 32B swizzle
  const uint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(weights), dims, str, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, small ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FNSSL_REQUIRE(r == CUDA_SUCCESS, "lstm(tcgen05): gate-slab weight tensor map failed (%d)", (int)r);
  return 0;
}

template <int H>
static int launch(const fnssl_lstm_args* a, const Plan& pl, cudaStream_t st) {
  constexpr int kCluster = H / 32, kNH = H / 32, kNHS = H / kSlabK;
  Params p{};
  int nxs = 0;
  for (int k0 = 0; k0 < a->c0; k0 += kSlabK) {
    p.xs_nkpack |= (uint32_t)((((a->c0 - k0) < kSlabK ? (a->c0 - k0) : kSlabK) + 15) / 16) << (4 * nxs);
    ++nxs;
  }
  p.nxs = nxs;
  p.small = pl.small;
  p.xstages = pl.xstages;
  const int nslabs = nxs + pl.small + kNHS;      // 64-wide column blocks of the packed buffer
  const int64_t wbytes = (int64_t)a->num_dirs * kNH * kChunkN * nslabs * kSlabK * 2;
  const int64_t need = wbytes + (int64_t)a->num_dirs * 4 * H * 4;
  FNSSL_REQUIRE(a->weights_bytes == need, "lstm(tcgen05): packed weight buffer is %lld bytes, expected %lld",
                (long long)a->weights_bytes, (long long)need);
  FNSSL_REQUIRE((reinterpret_cast<uintptr_t>(a->weights) & 15) == 0, "lstm(tcgen05): weights not 16-byte aligned");
  p.axis = a->axis; p.nf = a->nf; p.nt = a->nt;
  if (a->axis == FNSSL_ALONG_FREQ) {
    p.rows = (long long)a->nb * a->nt; p.steps = a->nf; p.chains_per_b = 0;
    p.nchains = (int)((p.rows + kChainRows - 1) / kChainRows);
  } else {
    p.rows = (long long)a->nb * a->nf; p.steps = a->nt; p.chains_per_b = (a->nf + kChainRows - 1) / kChainRows;
    p.nchains = a->nb * p.chains_per_b;
  }
  const int clusters = (p.nchains + 1) / 2;
  p.bias = reinterpret_cast<const float*>(reinterpret_cast<const char*>(a->weights) + wbytes);
  p.out0_off = a->out0_off;
  p.error_flag = tc_wait_timeout_enabled() ? tc_error_flag() : nullptr;
  p.debug = tc_debug_bits(0);

  CUtensorMap m0, mw;
  if (make_grid_map(&m0, a->src0, a->c0, a->ld0, a->nb, a->nt, a->nf, a->axis, kRows)) return 1;
  if (make_gate_weight_map(&mw, a->weights, nslabs, a->num_dirs * kNH)) return 1;
  CUtensorMap m1 = m0, mw2 = mw;
  if (pl.small) {
    if (make_small_grid_map(&m1, a->src1, a->c1, a->ld1, a->nb, a->nt, a->nf, a->axis, kRows)) return 1;
    if (make_gate_weight_map(&mw2, a->weights, nslabs, a->num_dirs * kNH, true)) return 1;
  }
  CUtensorMap mo0 = m0, mo1 = m0;
  if (a->out0) {
    if (make_out_map(&mo0, a->out0, a->out0_ld, a->nb, a->nt, a->nf, a->axis, kRows)) return 1;
    p.tma_out |= 1;
  }
  if (a->out1) {
    if (make_out_map(&mo1, a->out1, a->out1_ld, a->nb, a->nt, a->nf, a->axis, kRows)) return 1;
    p.tma_out |= a->addend ? 2 : 4;      // in-place reduce-add onto the residual operand / plain second copy of h
  }
  FNSSL_CUDA(cudaFuncSetAttribute(lstm_tc6_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)clusters * kCluster, (unsigned)a->num_dirs, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  FNSSL_CUDA(cudaLaunchKernelEx(&cfg, lstm_tc6_kernel<H>, m0, m1, mw, mw2, mo0, mo1, p));
  FNSSL_LAUNCH_CHECK("lstm_tc6_kernel");
  return 0;
}

}  // namespace tc6

// lstm_tc6.cu serves
//  * H = 256 layers without carried state whose outputs all go through TMA (FNSSL_TC_PAIR256=0 switches it off), by wave count:
//    both H = 256 kernels run clusters of 8 CTAs with one CTA per SM, kResident of them fit the GPU at a time (measured:
//    cudaOccupancyMaxActiveClusters = 15 on a 148-SM B200), so a launch takes ceil(clusters / kResident) waves; a wave of this
//    kernel covers twice the rows of one of lstm_tc4.cu (256 vs 128 per cluster) in ~1.35x its time (0.93-1.0 vs 0.71 ms at
//    c0 = 256, 249 steps: profiles/r2_lstm_variants.txt) -- it wins unless the wave counts quantise against it;
//  * H = 128 TWO-SOURCE layers of the same kind that are too small for lstm_tc5.cu (< 30 clusters of 512 rows; the dispatcher asks
//    that kernel first) and too large for lstm_tc4.cu's 64-row small-grid plan: clusters of 4 = 2 pairs x 2 chains of 128 rows --
//    the same CTA count as lstm_tc4.cu, a third of its DSMEM traffic, six x-ring stages instead of three (FNSSL_TC_PAIR=0
//    switches it off).
bool lstm_tc6_wants(const fnssl_lstm_args* a) {
  if (a->hidden != 256 && a->hidden != 128) return false;
  if (const char* e = getenv(a->hidden == 256 ? "FNSSL_TC_PAIR256" : "FNSSL_TC_PAIR")) { if (atoi(e) == 0) return false; }
  if (a->state_flags) return false;
  if (!tc6::make_plan(a->hidden, a->c0, a->c1).ok) return false;
  if (a->out0 && a->out0_off % 8) return false;
  if (a->out1 && a->addend && !(a->out1 == a->addend && a->out1_ld == a->addend_ld)) return false;
  const long long chains = a->axis == FNSSL_ALONG_FREQ ? ((long long)a->nb * a->nt + 127) / 128 : (long long)a->nb * ((a->nf + 127) / 128);
  const long long clusters = (chains + 1) / 2 * a->num_dirs;
  if (a->hidden == 128) {
    if (const char* e = getenv("FNSSL_TC_PAIR128_MIN")) return clusters >= atoi(e);
    // lstm_tc4.cu's small-grid plan (64-row tiles while they fit one wave: twice the CTAs) keeps the smallest layers
    const long long tiles64 = a->axis == FNSSL_ALONG_FREQ ? ((long long)a->nb * a->nt + 63) / 64 : (long long)a->nb * ((a->nf + 63) / 64);
    if (tiles64 * a->num_dirs * 4 <= 132) return false;
    // Measured at cfg2 size (16 utterances, ms): in16 0.73 vs 0.68 for lstm_tc4.cu, in256 0.82-0.84 vs 0.81-0.83 -- a draw, the step is
    // the same serial epilogue chain in both -- but 0.86 vs 1.07 for the two-source layer (in 256 + 16), where lstm_tc4.cu's x ring
    // is one stage short: only that shape is taken.
    return a->c1 > 0 && clusters >= 8;
  }
  if (const char* e = getenv("FNSSL_TC_PAIR256_MIN")) return clusters >= atoi(e);
  constexpr long long kResident = 15;
  const long long waves6 = (clusters + kResident - 1) / kResident;
  const long long waves4 = (chains * a->num_dirs + kResident - 1) / kResident;
  return waves6 * 27 < waves4 * 20;
}

int lstm_forward_tc6(const fnssl_lstm_args* a, cudaStream_t st) {
  using namespace tc6;
  const Plan pl = make_plan(a->hidden, a->c0, a->c1);
  FNSSL_REQUIRE(pl.ok && (a->hidden == 256 || a->hidden == 128), "lstm(tcgen05 pair kernel, M = 128): unsupported layer (H=%d c0=%d c1=%d)",
                a->hidden, a->c0, a->c1);
  return a->hidden == 256 ? launch<256>(a, pl, st) : launch<128>(a, pl, st);
}

}  // namespace fnssl
