// CTA-pair (tcgen05 cta_group::2) cluster LSTM kernel for H = 128: the kernel of FNSSL_ENGINE_TCGEN05 for LARGE layers (at
// least one wave of clusters, see lstm_tc5_wants below; FNSSL_TC_PAIR=0 disables it).  Smaller layers run lstm_tc4.cu, whose
// 128-row tiles fill the GPU at small batch.
//
// Why.  lstm_tc4.cu splits the 4H gate columns over a cluster of 4 CTAs that all hold the SAME rows, so every step each CTA
// pushes its [128 x 32] h tile to 3 peers and receives 3 tiles: 48 KB per slot through DSMEM, which measures at ~17-20 B/clk per
// SM -- >= 2.4 k cycles per slot, the bound of that kernel (profiles/r2_lstm_variants.txt).  Here the
// cluster's 4 CTAs are 2 PAIRS.  A pair (cluster ranks 2p, 2p+1) executes M = 256 MMAs with tcgen05.mma.cta_group::2: each CTA
// keeps its own 128 rows of the A operand (x_t slab / h tiles) and HALF of the B operand (64 of the 128 gate columns of a
// chunk), the hardware shares B across the pair.  So a pair owns 64 hidden units (two 32-unit chunks, "unit halves" uh = 0, 1)
// with the same 96 KB of resident weights per CTA as before, a CTA only needs h for ITS 128 rows, and it exchanges its
// [128 x 32] tile with ONE peer (the CTA with the same row half in the other pair): 16 KB per half-slot instead of 48 KB.
//
// Work decomposition.  Cluster tile = 2 chains x 256 rows.  Chain c, CTA (pair p, row half j): rows c*256 + j*128 .. +128 of the
// chain's tile, hidden units 64p .. 64p+64.  A *half-slot* n = (step t, chain c, unit half uh), n = 4t + 2c + uh, is what a slot
// is in lstm_tc4.cu: one [128 rows x 128 gate columns] accumulator per CTA, the epilogue of 128 rows x 32 units, one h tile.
// The two unit halves of a (t, c) share the x_t slabs (loaded once, released after the second use) and the h_{t-1} operand.
//
//   warp 0        TMA producer (both CTAs load their own rows; the "full" barrier lives on the pair's leader, rank 2p)
//   warp 1        leader: h-part MMA issuer;  other CTA: relay ("my weights / my h tiles are in place" -> leader's barriers)
//   warps 2..17   epilogue: all 16 warps take the half-slots one after the other (thread = row x 8 hidden units)
//   warp 18       leader: x-part MMA issuer
//   warp 19       publisher: DSMEM push of the CTA's h tile to its ONE peer, TMA output stores
//
// Single-buffered h as in lstm_tc4.cu: H_FREE(chain) collects one multicast commit per pair leader (after the h-part of BOTH
// unit halves of the step) plus the publisher's "stores drained" arrive; the epilogue of either unit half waits for it before
// it overwrites its tile.  The h-parts of uh = 0 and uh = 1 are issued back to back, so that wait is over long before the gate
// math of uh = 0 finishes.
//
// TMEM (per CTA, allocated with cta_group::2): four accumulators of 128 columns, one pair (both unit halves) per chain; the cell
// state lives in the epilogue threads' registers.
// Replaces nn.LSTM at FN-SSL/Lightning/Model.py:38,46 and IPDnet/FixedAarryIPDnet.py:32,36 (+ glue :35-37,41-45,49).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace fnssl {
namespace tc5 {

// TC5_SPLIT = 1: the epilogue warps form two groups that take the two unit halves of a (step, chain) concurrently; 0: all 16
// warps take the half-slots one after the other.  (A/B on the B200: profiles/r2_lstm_variants.txt)
#ifndef TC5_SPLIT
#define TC5_SPLIT 0
#endif
constexpr int H = 128;
constexpr int kThreads = 640;
constexpr int kXWarp = 18, kPubWarp = 19;
constexpr int kEpiWarps = 16;
constexpr int kSlabK = 64;
constexpr int kRows = 128;                 // rows per CTA per chain
constexpr int kChainRows = 256;            // rows per chain (pair: 2 x 128)
constexpr int kWHalf = 64 * 128;           // [64 gate columns x 64] fp16: this CTA's half of a [128 x 64] weight slab
constexpr int kXSlab = kRows * 128;        // [128 rows x 64] fp16
constexpr int kHTile = kRows * 64;         // [128 rows x 32 units] fp16 (64B swizzle)
constexpr int kChunkN = 128;
#ifndef TC5_PUSH_TILE
#define TC5_PUSH_TILE 0      // 1: one DSMEM bulk copy of the whole [128 x 32] tile to the peer instead of four quadrant copies (measured: a draw, off)
#endif
#ifndef TC5_STORE_QUADRANTS
#define TC5_STORE_QUADRANTS 0      // 1: four [32 x 32] output boxes per tile and destination (A/B; 0 = one [128 x 32] box)
#endif
#ifndef TC5_ACCBUFS
#define TC5_ACCBUFS 4      // 4: one accumulator PAIR per chain (the x-part of one chain overlaps the other chain's h-part / epilogue)
#endif
constexpr int kMaxXSlabs = 6, kMaxXStages = 6, kAccBufs = TC5_ACCBUFS;
constexpr int kSmemLimit = 232448;
// barriers
constexpr int B_WFULL = 0, B_WMATE = 1, B_XFULL = 2, B_XEMPTY = B_XFULL + kMaxXStages, B_ACCFULL = B_XEMPTY + kMaxXStages,
              B_ACCEMPTY = B_ACCFULL + kAccBufs, B_XPDONE = B_ACCEMPTY + kAccBufs, B_HFULL = B_XPDONE + kAccBufs,
              B_HMATE = B_HFULL + 2, B_HFREE = B_HMATE + 2, B_HREADY = B_HFREE + 2, B_ACCDRAIN = B_HREADY + 16, kNumBars = B_ACCDRAIN + kAccBufs;
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kChunkN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // f16 x f16 -> f32, M = 256, N = 128

struct Params {
  int nxs;
  uint32_t xs_srcmask;             // bit j: slab j comes from src1
  unsigned long long xs_k0pack;    // byte j: first channel of slab j / 16
  uint32_t xs_nkpack;              // nibble j: K=16 steps of slab j (1..4)
  int xstages;
  int steps, axis, nf, nt;
  long long rows;                  // sequences of the layer
  int chains_per_b;                // ALONG_TIME: 256-row chain tiles per utterance
  int nchains;                     // chain tiles of the layer (the last cluster may own an empty second chain)
  const float* bias;               // [dirs][4H], accumulator column order [chunk][gate][unit]
  int out0_off;
  int tma_out;                     // bit 0: out0 tile stores; bit 1: in-place reduce-add onto out1 == addend; bit 2: out1 = second copy of h
  int* error_flag;
  long long* trace;                // FNSSL_TC_TRACE: clock64 stamps of cluster 0 / CTA rank 0 (pair leader): [half-slot 32..47][16 events]
  int debug;                       // timing experiments (FNSSL_TC_DEBUG; wrong results): 1 = no gate math, 2 = no fence.proxy.async, 4 = no TMA stores
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

// ---- cta_group::2 forms -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (when every MMA issued so far by this thread has completed) on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// TMA tile load into THIS CTA's shared memory whose transaction bytes complete on a barrier of the pair's leader CTA
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
lstm_tc5_kernel(const __grid_constant__ CUtensorMap map_src0, const __grid_constant__ CUtensorMap map_src1,
                const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out0,
                const __grid_constant__ CUtensorMap map_out1, const Params p) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) unsigned long long bars[kNumBars];
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)(rank >> 1), jh = (int)(rank & 1);      // pair -> hidden units [64 pair, +64); jh -> row half of a chain
  const bool leader = jh == 0;
  const uint32_t leader_rank = rank & ~1u;
  const uint32_t peer_rank = (uint32_t)(2 * (1 - pair) + jh);   // the CTA that holds the same rows in the other pair
  const uint16_t pair_mask = (uint16_t)(3u << (2 * pair));
  const int cluster_tile = blockIdx.x >> 2;
  const int nxs = p.nxs, XS = p.xstages, L = p.steps;
  const int nhalf = 4 * L;                                       // half-slots

  const uint32_t dyn0 = (smem_addr(smem_dyn) + 1023u) & ~1023u;
  const uint32_t w_base = dyn0;                                           // [uh][slab: nxs x, then 2 h] halves of 8 KB
  const int nslabs = nxs + 2;
  const uint32_t hs_base = w_base + (uint32_t)(2 * nslabs) * kWHalf;      // h operand: [chain][chunk 0..3] tiles
  const uint32_t xr_base = hs_base + 8u * kHTile;                         // x ring
  const uint32_t bias_base = xr_base + (uint32_t)XS * kXSlab;             // 256 floats: [uh][gate][unit]
  float* bias_s = reinterpret_cast<float*>(smem_dyn + (bias_base - smem_addr(smem_dyn)));

  const uint32_t bar0 = smem_addr(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const bool tr_cta = p.trace != nullptr && blockIdx.x == ((p.debug & 16) ? 1u : 0u) && blockIdx.y == 0;
#define TP(n, ev) do { if (tr_cta && (n) >= 32 && (n) < 48) p.trace[((n) - 32) * 16 + (ev)] = clock64(); } while (0)

  if (tid == 0) {
    mbar_init(BAR(B_WFULL), 1);
    mbar_init(BAR(B_WMATE), 1);
    for (int i = 0; i < kMaxXStages; ++i) { mbar_init(BAR(B_XFULL + i), 1); mbar_init(BAR(B_XEMPTY + i), 1); }
    for (int i = 0; i < kAccBufs; ++i) {
      mbar_init(BAR(B_ACCFULL + i), 1);
      // leader's barrier: its own epilogue warps of the half-slot + ONE arrive relayed from the other CTA (whose warps arrive on
      // their local ACCDRAIN barrier: 16 remote release-arrives per half-slot slowed that CTA's epilogue down measurably)
      mbar_init(BAR(B_ACCEMPTY + i), (TC5_SPLIT ? kEpiWarps / 2 : kEpiWarps) + ((p.debug & 8) ? 0 : 1));
      mbar_init(BAR(B_ACCDRAIN + i), TC5_SPLIT ? kEpiWarps / 2 : kEpiWarps);
      mbar_init(BAR(B_XPDONE + i), 1);
    }
    for (int c = 0; c < 2; ++c) {
      mbar_init(BAR(B_HFULL + c), 9);     // expect_tx arrive + 2 unit halves x 4 quadrants of the CTA's own tiles (+ 16 KB of tx from the peer)
      mbar_init(BAR(B_HMATE + c), 1);     // relay of the pair's other CTA
      mbar_init(BAR(B_HFREE + c), 3);     // the two pair leaders' commits + the publisher ("stores drained")
    }
    for (int i = 0; i < 16; ++i) mbar_init(BAR(B_HREADY + i), TC5_SPLIT ? 2 : 4);     // [(chain, unit half)][quadrant]: that quadrant's warps of the half-slot
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&map_src0); prefetch_tmap(&map_src1); prefetch_tmap(&map_w); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_base_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 2 * kChunkN; i += kThreads) bias_s[i] = p.bias[dir * 4 * H + (2 * pair) * kChunkN + i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // every CTA's barriers are initialised before any multicast / remote traffic can reach them
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  [[maybe_unused]] const uint32_t tmem_c = tmem;             // (cell state lives in registers; columns [0,128) are free)
  const uint32_t tmem_acc = tmem + (uint32_t)(512 - kAccBufs * 128);     // gate accumulators: kAccBufs buffers x 128 columns

  // per-chain coordinates of this CTA's 128 rows
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

  // (selects instead of run-time indexed arrays: the single-lane roles must not touch local memory)
  const int cbA = cb_[0], cbB = cb_[1], crA = cr0_[0], crB = cr0_[1];
#define CB(c) ((c) ? cbB : cbA)
#define CR0(c) ((c) ? crB : crA)
  if (warp == 0) {
    // ============================== TMA producer (both CTAs) ==============================
    if (elect_one()) {
      // resident weights: for each unit half the 64 gate columns [64 jh, +64) of chunk 2 pair + uh, every K slab
      mbar_expect_tx(BAR(B_WFULL), (uint32_t)(2 * nslabs) * kWHalf);
      for (int uh = 0; uh < 2; ++uh)
        for (int j = 0; j < nslabs; ++j)
          tma_load_2d(w_base + (uint32_t)(uh * nslabs + j) * kWHalf, &map_w, BAR(B_WFULL), j * kSlabK,
                      (dir * 4 + 2 * pair + uh) * kChunkN + 64 * jh);
      const uint32_t lead_xfull0 = mapa_shared(BAR(B_XFULL), leader_rank);
      int stage = 0;
      uint32_t phase = 0;
      bool wrapped = false;
      for (int n2 = 0; n2 < 2 * L; ++n2) {        // one pass per (step, chain): its slabs serve both unit halves
        const int t = n2 >> 1, c = n2 & 1;
        const int s = dir ? (L - 1 - t) : t;
        for (int j = 0; j < nxs; ++j) {
          if (wrapped) mbar_wait(BAR(B_XEMPTY + stage), phase, p.error_flag, 100 + stage);
          if (leader) mbar_expect_tx(BAR(B_XFULL + stage), 2u * kXSlab);     // this CTA's slab + the other CTA's
          const CUtensorMap* m = ((p.xs_srcmask >> j) & 1) ? &map_src1 : &map_src0;
          const uint32_t dst = xr_base + (uint32_t)stage * kXSlab;
          const int k0 = (int)((p.xs_k0pack >> (8 * j)) & 0xff) * 16;
          const uint32_t fb = lead_xfull0 + 8u * (uint32_t)stage;
          if (along_f) tma2_load_4d(dst, m, fb, k0, s, CR0(c), 0);
          else tma2_load_4d(dst, m, fb, k0, CR0(c), s, CB(c));
          if (++stage == XS) { stage = 0; phase ^= wrapped ? 1u : 0u; wrapped = true; }
        }
      }
    }
    __syncwarp();
  } else if (warp == kXWarp) {
    // ============================== x-part MMA issuer (pair leader only) ==============================
    // One pass per (step, chain): every x slab feeds BOTH unit halves back to back (two accumulator buffers) and is released
    // right after, so the ring never has to hold a whole pass (it may have fewer stages than the pass has slabs).
    if (leader && elect_one()) {
      mbar_wait(BAR(B_WFULL), 0, p.error_flag, 200);
      mbar_wait(BAR(B_WMATE), 0, p.error_flag, 201);
      const uint64_t a_desc0 = make_sw128_desc(xr_base);
      const uint64_t b_desc0 = make_sw128_desc(w_base);
      int xstage = 0, a0 = 0;
      uint32_t xphase = 0, empty_par = 0;
      for (int n2 = 0; n2 < 2 * L; ++n2) {
        const int a1 = (a0 == kAccBufs - 1) ? 0 : a0 + 1;
        TP(2 * n2, 9);
        if (2 * n2 >= kAccBufs) {
          mbar_wait(BAR(B_ACCEMPTY + a0), (empty_par >> a0) & 1u, p.error_flag, 202 + a0);
          empty_par ^= 1u << a0;
        }
        if (2 * n2 + 1 >= kAccBufs) {
          mbar_wait(BAR(B_ACCEMPTY + a1), (empty_par >> a1) & 1u, p.error_flag, 202 + a1);
          empty_par ^= 1u << a1;
        }
        tc_fence_after();
        TP(2 * n2, 10);
        const uint32_t d0 = tmem_acc + (uint32_t)a0 * kChunkN, d1 = tmem_acc + (uint32_t)a1 * kChunkN;
        uint32_t nkp = p.xs_nkpack;
        for (int j = 0; j < nxs; ++j, nkp >>= 4) {
          mbar_wait(BAR(B_XFULL + xstage), xphase, p.error_flag, 210 + xstage);    // both CTAs' copies of the slab have landed
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
          umma2_commit_mc(BAR(B_XEMPTY + xstage), pair_mask);      // both unit halves have read the slab, in both CTAs
          if (++xstage == XS) { xstage = 0; xphase ^= 1u; }
        }
        umma2_commit_mc(BAR(B_XPDONE + a0), (uint16_t)(1u << leader_rank));
        umma2_commit_mc(BAR(B_XPDONE + a1), (uint16_t)(1u << leader_rank));
        TP(2 * n2, 11);
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
        if (!(p.debug & 8)) mbar_arrive_remote(lead_empty0 + 8u * (uint32_t)a);      // (debug 8: timing experiment, unsafe)
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
          TP(n, 0);
          mbar_wait(BAR(B_XPDONE + a), (xp_par >> a) & 1u, p.error_flag, 240 + a);     // G_x of this half-slot is complete
          xp_par ^= 1u << a;
          TP(n, 1);
          if (t > 0) {
            if (uh == 0) {
              // h_{t-1} of the chain: this CTA's two tiles by local arrives, the peer's two tiles as DSMEM bulk copies (tx bytes);
              // the other CTA of the pair reports the same for its rows through the relay
              mbar_expect_tx(BAR(B_HFULL + c), 2u * kHTile);
              mbar_wait_cluster(BAR(B_HFULL + c), (uint32_t)((t - 1) & 1), p.error_flag, 220 + c);
              TP(n, 2);
              mbar_wait_cluster(BAR(B_HMATE + c), (uint32_t)((t - 1) & 1), p.error_flag, 222 + c);
              TP(n, 3);
            }
            tc_fence_after();
            const uint32_t d_tmem = tmem_acc + (uint32_t)a * kChunkN;
            const uint64_t a_chain = h_desc0 + (uint64_t)(c * 4 * (kHTile >> 4));
            const uint64_t b_uh = wh_desc0 + (uint64_t)(uh * nslabs * (kWHalf >> 4));
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {   // K = 32 units of chunk kc: two K=16 steps; W columns inside 128B-swizzled slab kc/2
              const uint64_t a_desc = a_chain + (uint64_t)(kc * (kHTile >> 4));
              const uint64_t b_desc = b_uh + (uint64_t)((kc >> 1) * (kWHalf >> 4) + 4 * (kc & 1));
#pragma unroll
              for (int k = 0; k < 2; ++k) umma2_f16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, kIdesc, 1u);
            }
            // after the second unit half h_{t-1} of this chain has been read by this pair: every CTA of the cluster learns it
            if (uh == 1) umma2_commit_mc(BAR(B_HFREE + c), (uint16_t)0xF);
          } else {
            tc_fence_after();
          }
          umma2_commit_mc(BAR(B_ACCFULL + a), pair_mask);     // both CTAs' epilogues
          TP(n, 4);
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
            mbar_expect_tx(BAR(B_HFULL + c), 2u * kHTile);
            TP(4 * t + 2 * c, 1);
            mbar_wait_cluster(BAR(B_HFULL + c), (uint32_t)((t - 1) & 1), p.error_flag, 252 + c);
            TP(4 * t + 2 * c, 2);
            mbar_arrive_remote(lead_hmate0 + 8u * (uint32_t)c);
            TP(4 * t + 2 * c, 3);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kPubWarp) {
    // ============================== publisher ==============================
    if (elect_one()) {
      const uint32_t peer_hs = mapa_shared(hs_base, peer_rank);
      const uint32_t peer_hfull0 = mapa_shared(BAR(B_HFULL), peer_rank);
      const bool tma_any = p.tma_out != 0 && !(p.debug & 4);
      for (int n = 0; n < nhalf; ++n) {
        const int t = n >> 2, c = (n >> 1) & 1, uh = n & 1;
        const int s = dir ? (L - 1 - t) : t;
        const bool push = t + 1 < L;
        const int kc = 2 * pair + uh;
        const uint32_t tile_off = (uint32_t)(c * 4 + kc) * kHTile;
        TP(n, 12);
        if (push || tma_any) {
#if TC5_PUSH_TILE
#pragma unroll
          for (int q = 0; q < 4; ++q)
            mbar_wait(BAR(B_HREADY + (c * 2 + uh) * 4 + q), (uint32_t)(t & 1), p.error_flag, 400 + (c * 2 + uh) * 4 + q);
          if (push) {
            bulk_copy_s2c(peer_hs + tile_off, hs_base + tile_off, (uint32_t)kHTile, peer_hfull0 + 8u * (uint32_t)c);
#pragma unroll
            for (int q = 0; q < 4; ++q) mbar_arrive(BAR(B_HFULL + c));     // the local copy of every quadrant is in place
          }
#else
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            mbar_wait(BAR(B_HREADY + (c * 2 + uh) * 4 + q), (uint32_t)(t & 1), p.error_flag, 400 + (c * 2 + uh) * 4 + q);
            if (push) {
              const uint32_t off = tile_off + (uint32_t)q * 2048u;
              bulk_copy_s2c(peer_hs + off, hs_base + off, 2048u, peer_hfull0 + 8u * (uint32_t)c);
              mbar_arrive(BAR(B_HFULL + c));     // the local copy of this quadrant is in place
            }
          }
#endif
          TP(n, 13);
          if (tma_any) {
            // ONE [128 rows x 32 channels] box per destination: a TMA store costs the issuing lane ~170 cycles (timeline: 0.7 k for
            // four per-quadrant stores), and with two destinations per-quadrant stores made the publisher the slowest role
            const int out_c = dir * H + kc * 32;
#if TC5_STORE_QUADRANTS
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t off = tile_off + (uint32_t)q * 2048u;
              const int r0 = CR0(c) + q * 32;
#else
            {
              const uint32_t off = tile_off;
              const int r0 = CR0(c);
#endif
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
        TP(n, 14);
        // the stores of the OTHER chain's two half-slots (groups n-3, n-2) have read their tiles
        if (uh == 1 && n >= 3) {
          if (tma_any) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
          mbar_arrive(BAR(B_HFREE + (c ^ 1)));
        }
      }
    }
    __syncwarp();
  } else {
#if TC5_SPLIT
    // ============================== epilogue warps ==============================
    // Two GROUPS of 8 warps (2 per TMEM lane quadrant = 2 per SM sub-partition); a thread covers 16 units of a half-slot in two
    // batches of 8.
    //   TC5_SPLIT = 1: group g takes the unit half uh = g of every (step, chain).  The two unit halves become ready 0.5 k cycles
    //                  apart, so the groups run IN phase: no overlap is won and every chain's step gets longer (measured -7..-13 %).
    //   TC5_SPLIT = 2: group g takes CHAIN g (both unit halves, one after the other).  The two chains run half a step apart by
    //                  construction, so one group's barrier wake-ups, TMEM loads, tile stores and its wait for the chain's next
    //                  h-part overlap the other group's gate math on the same MUFU pipe.
    const int q = warp & 3;                    // TMEM lane quadrant of this warp
    const int widx = (warp - 2) >> 2;          // 0..3
    const int g = widx & 1;                    // group
    const int ubase = (widx >> 1) * 16;        // this warp's 16 of the half-slot's 32 units (two batches of 8)
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int r = q * 32 + lane;               // row of this thread inside the CTA's 128 rows
    const uint32_t hrow = (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u;     // row offset inside a [128 x 32] tile (64B swizzle)
    const uint32_t hsw = (uint32_t)((r >> 1) & 3);                                  // 16-byte chunk ^= (row >> 1) & 3
    const bool tma_any = p.tma_out != 0;
    static_assert(kAccBufs == 4, "the split epilogues assume one accumulator pair per chain (buffer = 2 chain + unit half)");

    // the cell state of this thread's two half-slots x two batches x 8 units lives in registers
    float creg[2][2][8];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int i = 0; i < 8; ++i) creg[c][b][i] = 0.0f;
#pragma unroll 1
    for (int t = 0; t < L; ++t) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {            // the group's two half-slots of this step
        const int c = (TC5_SPLIT == 2) ? g : k, uh = (TC5_SPLIT == 2) ? k : g;
        const int cu = 2 * c + uh, a = cu;
        const bool trw = warp == 2 + 4 * g && lane == 0 && g == 0;
        if (trw) TP(4 * t + cu, 5);
        mbar_wait(BAR(B_ACCFULL + a), (uint32_t)(t & 1), p.error_flag, 300 + a);
        if (trw) TP(4 * t + cu, 6);
        tc_fence_after();
        const bool will_publish = t + 1 < L || tma_any;
        const uint32_t tile = hs_base + (uint32_t)(c * 4 + 2 * pair + uh) * kHTile + hrow;
        const float* bias_g = bias_s + uh * kChunkN;
        bool hfree_ok = true;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int u0 = ubase + 8 * b;
          const uint32_t acc = tmem_acc + (uint32_t)a * kChunkN + lane_off + u0;
          float (&cs)[8] = creg[k][b];
          float gti[8], gtf[8], gtg[8], gto[8];
          tmem_ld8(acc + 0 * 32, gti);
          tmem_ld8(acc + 1 * 32, gtf);
          tmem_ld8(acc + 2 * 32, gtg);
          tmem_ld8(acc + 3 * 32, gto);
          // poll "h_{t-1} of this chain has been read everywhere" now; the answer is only needed after the gate math
          if (b == 0) hfree_ok = !(will_publish && t > 0) || mbar_test_wait(BAR(B_HFREE + c), (uint32_t)((t - 1) & 1));
          tmem_wait_ld();
          tmem_ld_dep(gti); tmem_ld_dep(gtf); tmem_ld_dep(gtg); tmem_ld_dep(gto);
          if (b == 1) {       // accumulator drained (this warp): the leader's x-part issuer may refill it
            if (trw) TP(4 * t + cu, 15);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(BAR((leader ? B_ACCEMPTY : B_ACCDRAIN) + a));
            }
          }
          const float* bsp = bias_g + u0;
          float hv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            hv[e] = lstm_cell_tanh(gti[e] + bsp[e], gtf[e] + bsp[32 + e], gtg[e] + bsp[64 + e], gto[e] + bsp[96 + e], cs[e]);
          __half2 h01 = __floats2half2_rn(hv[0], hv[1]), h23 = __floats2half2_rn(hv[2], hv[3]);
          __half2 h45 = __floats2half2_rn(hv[4], hv[5]), h67 = __floats2half2_rn(hv[6], hv[7]);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h01); pk.y = *reinterpret_cast<uint32_t*>(&h23);
          pk.z = *reinterpret_cast<uint32_t*>(&h45); pk.w = *reinterpret_cast<uint32_t*>(&h67);
          if (will_publish) {
            // both pairs' h-parts of step t have finished reading h_{t-1} of this chain, the peer's pushes of h_{t-1} have
            // landed and the stores that read these tiles have drained
            if (b == 0 && !hfree_ok) mbar_wait(BAR(B_HFREE + c), (uint32_t)((t - 1) & 1), p.error_flag, 320 + c);
            st_shared_v4(tile + ((((uint32_t)(u0 >> 3)) ^ hsw) << 4), pk);
          }
        }
        if (trw) TP(4 * t + cu, 7);
        if (will_publish) {
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_HREADY + cu * 4 + q));
        }
        if (trw) TP(4 * t + cu, 8);
      }
    }
  }

#else
    // ============================== epilogue warps (TC5_SPLIT = 0) ==============================
    // all 16 warps take the half-slots one after the other (as lstm_tc4.cu's 128-row path): thread = (row, 8 hidden units)
    const int q = warp & 3;                    // TMEM lane quadrant of this warp
    const int sg = (warp - 2) >> 2;            // 8-unit group of the half-slot's 32 units
    const int u0 = sg * 8;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int r = q * 32 + lane;               // row of this thread inside the CTA's 128 rows
    const uint32_t hpiece = (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u + (uint32_t)((sg ^ ((r >> 1) & 3)) << 4);
    const bool tma_any = p.tma_out != 0;
    float creg[4][8];                          // cell state per (chain, unit half)
#pragma unroll
    for (int cu = 0; cu < 4; ++cu)
#pragma unroll
      for (int i = 0; i < 8; ++i) creg[cu][i] = 0.0f;
    int a = 0;
    uint32_t full_par = 0;
#pragma unroll 1
    for (int t = 0; t < L; ++t) {
#pragma unroll
      for (int cu = 0; cu < 4; ++cu) {
        const int c = cu >> 1, uh = cu & 1;
        const bool trw = warp == 2 && lane == 0;
        if (trw) TP(4 * t + cu, 5);
        mbar_wait(BAR(B_ACCFULL + a), (full_par >> a) & 1u, p.error_flag, 300 + a);
        full_par ^= 1u << a;
        if (trw) TP(4 * t + cu, 6);
        tc_fence_after();
        const uint32_t acc = tmem_acc + (uint32_t)a * kChunkN + lane_off + u0;
        float (&cs)[8] = creg[cu];
        float gti[8], gtf[8], gtg[8], gto[8];
        tmem_ld8(acc + 0 * 32, gti);
        tmem_ld8(acc + 1 * 32, gtf);
        tmem_ld8(acc + 2 * 32, gtg);
        tmem_ld8(acc + 3 * 32, gto);
        const bool will_publish = t + 1 < L || tma_any;
        // (H_FREE only orders this thread's OVERWRITE of the tile after the MMAs' / stores' reads of it -- no data is acquired
        // through it, so the CTA-scope forms are sufficient; debug 32 uses the cluster-scope acquire forms for comparison)
        const bool hfree_ok = !(will_publish && t > 0) ||
                              ((p.debug & 32) ? mbar_test_wait_cluster(BAR(B_HFREE + c), (uint32_t)((t - 1) & 1))
                                              : mbar_test_wait(BAR(B_HFREE + c), (uint32_t)((t - 1) & 1)));
        tmem_wait_ld();
        tmem_ld_dep(gti); tmem_ld_dep(gtf); tmem_ld_dep(gtg); tmem_ld_dep(gto);
        if (trw) TP(4 * t + cu, 15);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {      // accumulator drained (this warp): the leader's x-part issuer may refill it
          mbar_arrive(BAR((leader ? B_ACCEMPTY : B_ACCDRAIN) + a));
        }
        a = (a == kAccBufs - 1) ? 0 : a + 1;
        const float* bsp = bias_s + uh * kChunkN + u0;
        float hv[8];
        if (p.debug & 1) {
#pragma unroll
          for (int e = 0; e < 8; ++e) hv[e] = gti[e] + gtf[e] + gtg[e] + gto[e] + cs[e];
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            hv[e] = lstm_cell_tanh(gti[e] + bsp[e], gtf[e] + bsp[32 + e], gtg[e] + bsp[64 + e], gto[e] + bsp[96 + e], cs[e]);
        }
        __half2 h01 = __floats2half2_rn(hv[0], hv[1]), h23 = __floats2half2_rn(hv[2], hv[3]);
        __half2 h45 = __floats2half2_rn(hv[4], hv[5]), h67 = __floats2half2_rn(hv[6], hv[7]);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h01); pk.y = *reinterpret_cast<uint32_t*>(&h23);
        pk.z = *reinterpret_cast<uint32_t*>(&h45); pk.w = *reinterpret_cast<uint32_t*>(&h67);
        if (trw) TP(4 * t + cu, 7);
        if (will_publish) {
          if (!hfree_ok) {
            if (p.debug & 32) mbar_wait_cluster(BAR(B_HFREE + c), (uint32_t)((t - 1) & 1), p.error_flag, 320 + c);
            else mbar_wait(BAR(B_HFREE + c), (uint32_t)((t - 1) & 1), p.error_flag, 320 + c);
          }
          st_shared_v4(hs_base + (uint32_t)(c * 4 + 2 * pair + uh) * kHTile + hpiece, pk);
          if (!(p.debug & 2)) fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_HREADY + cu * 4 + q));
        }
        if (trw) TP(4 * t + cu, 8);
      }
    }
  }

#endif
#undef TP
#undef CB
#undef CR0
  bulk_wait_all();         // TMA tile stores of this thread (if any) are complete before the CTA may exit
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // no CTA leaves while a peer may still write into its shared memory / while the pair's MMAs read it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------

struct Plan { bool ok; int xstages; int nxs; size_t smem; };

static Plan make_plan(int c0, int c1) {
  Plan pl{false, 0, 0, 0};
  if (c0 % 16 || c1 % 16 || c0 <= 0) return pl;
  const int nxs = (c0 + 63) / 64 + (c1 + 63) / 64;
  if (nxs > kMaxXSlabs) return pl;
  const long fixed = 2L * (nxs + 2) * kWHalf + 8L * kHTile + 2 * kChunkN * 4 + 1024;
  long xs = (kSmemLimit - 1024 - fixed) / kXSlab;
  if (xs > kMaxXStages) xs = kMaxXStages;
  if (xs < 2) return pl;
  pl.ok = true; pl.xstages = (int)xs; pl.nxs = nxs; pl.smem = (size_t)fixed + (size_t)xs * kXSlab;
  return pl;
}

// 2-D fp16 map over the packed weights, box = [64 rows x 64 K]: one CTA's half of a [128 x 64] slab
static int make_half_weight_map(CUtensorMap* m, const void* weights, int nslabs, int nchunks_total) {
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
  const uint32_t box[2] = {kSlabK, 64};
  const uint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(weights), dims, str, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FNSSL_REQUIRE(r == CUDA_SUCCESS, "lstm(tcgen05): half-slab weight tensor map failed (%d)", (int)r);
  return 0;
}

}  // namespace tc5

// EXPERIMENTAL, off by default (FNSSL_TC_PAIR=1 enables it for H = 128 layers without carried state whose outputs all go through
// TMA and that have at least FNSSL_TC_PAIR_MIN [30] 512-row cluster tiles).  Measured on the B200 (profiles/r2_lstm_variants.txt):
// correct and bit-identical to the round-1 kernel's accumulation order, a third of lstm_tc4.cu's DSMEM bytes per row, but only
// 0-6 % faster per row: with the exchange out of the way a half-slot is bound by the epilogue warps' serial work (barrier wake-up
// + TMEM loads + MUFU-bound gate math ~1.4 k + hand-off = 2.2-2.8 k cycles), which both kernels share.  Kept as the starting point
// for the H = 256 version (M = 128 per pair turns that layer's half-rate M = 64 MMAs into full-rate ones) and tested.
bool lstm_tc5_wants(const fnssl_lstm_args* a) {
  if (const char* on = getenv("FNSSL_TC_PAIR")) { if (atoi(on) == 0) return false; }      // FNSSL_TC_PAIR=0: lstm_tc4.cu only
  if (a->hidden != 128 || a->state_flags) return false;
  if (!tc5::make_plan(a->c0, a->c1).ok) return false;
  if (a->out0 && a->out0_off % 8) return false;
  if (a->out1 && a->addend && !(a->out1 == a->addend && a->out1_ld == a->addend_ld)) return false;      // residual output: in place only
  const long long chains = a->axis == FNSSL_ALONG_FREQ ? ((long long)a->nb * a->nt + 255) / 256 : (long long)a->nb * ((a->nf + 255) / 256);
  const long long clusters = (chains + 1) / 2 * a->num_dirs;
  int min_clusters = 30;       // 33 clusters of 4 CTAs are co-resident; below ~one wave lstm_tc4.cu's smaller tiles win
  if (const char* e = getenv("FNSSL_TC_PAIR_MIN")) min_clusters = atoi(e);
  return clusters >= min_clusters;
}

static long long* g_trace5 = nullptr;     // FNSSL_TC_TRACE (diagnostic, single device)
long long* lstm_tc5_trace_buffer() { return g_trace5; }

int lstm_forward_tc5(const fnssl_lstm_args* a, cudaStream_t st) {
  using namespace tc5;
  const Plan pl = make_plan(a->c0, a->c1);
  FNSSL_REQUIRE(pl.ok && a->hidden == 128, "lstm(tcgen05 pair kernel): unsupported layer (H=%d c0=%d c1=%d)", a->hidden, a->c0, a->c1);
  Params p{};
  int nxs = 0;
  for (int src = 0; src < 2; ++src) {
    const int c = src ? a->c1 : a->c0;
    for (int k0 = 0; k0 < c; k0 += kSlabK) {
      p.xs_srcmask |= (uint32_t)src << nxs;
      p.xs_k0pack |= (unsigned long long)(k0 / 16) << (8 * nxs);
      p.xs_nkpack |= (uint32_t)((((c - k0) < kSlabK ? (c - k0) : kSlabK) + 15) / 16) << (4 * nxs);
      ++nxs;
    }
  }
  p.nxs = nxs;
  p.xstages = pl.xstages;
  const int nslabs = nxs + 2;
  const int64_t wbytes = (int64_t)a->num_dirs * 4 * kChunkN * nslabs * kSlabK * 2;
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
  p.debug = tc_debug_bits(16 | 32);
  if (getenv("FNSSL_TC_TRACE")) {
    if (!g_trace5) { if (cudaMalloc(&g_trace5, 256 * sizeof(long long)) != cudaSuccess) g_trace5 = nullptr; }
    if (g_trace5) cudaMemsetAsync(g_trace5, 0, 256 * sizeof(long long), st);
    p.trace = g_trace5;
  }

  CUtensorMap m0, m1, mw;
  if (make_grid_map(&m0, a->src0, a->c0, a->ld0, a->nb, a->nt, a->nf, a->axis, kRows)) return 1;
  if (a->c1 > 0) { if (make_grid_map(&m1, a->src1, a->c1, a->ld1, a->nb, a->nt, a->nf, a->axis, kRows)) return 1; }
  else m1 = m0;
  if (make_half_weight_map(&mw, a->weights, nslabs, a->num_dirs * 4)) return 1;
  CUtensorMap mo0 = m0, mo1 = m0;
  if (a->out0) {
    if (make_out_map(&mo0, a->out0, a->out0_ld, a->nb, a->nt, a->nf, a->axis, TC5_STORE_QUADRANTS ? 32 : kRows)) return 1;
    p.tma_out |= 1;
  }
  if (a->out1) {
    if (make_out_map(&mo1, a->out1, a->out1_ld, a->nb, a->nt, a->nf, a->axis, TC5_STORE_QUADRANTS ? 32 : kRows)) return 1;
    p.tma_out |= a->addend ? 2 : 4;      // in-place reduce-add onto the residual operand / plain second copy of h
  }
  FNSSL_CUDA(cudaFuncSetAttribute(lstm_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)clusters * 4, (unsigned)a->num_dirs, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  FNSSL_CUDA(cudaLaunchKernelEx(&cfg, lstm_tc5_kernel, m0, m1, mw, mo0, mo1, p));
  FNSSL_LAUNCH_CHECK("lstm_tc5_kernel");
  return 0;
}

}  // namespace fnssl
