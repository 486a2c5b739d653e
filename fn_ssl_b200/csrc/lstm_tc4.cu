// Two-chain cluster-resident tcgen05 LSTM kernel (fourth generation of FNSSL_ENGINE_TCGEN05), H in {64, 128, 256}.
//
// Decomposition -- the 4H gate columns are split over a cluster of C = H/32 CTAs, each keeping its
// weight slice resident in shared memory, x_t slabs TMA-multicast to the cluster, h_t exchanged by DSMEM bulk copies --
// but a cluster now owns TWO full row tiles (sub-tiles A and B, SUB = 128 rows each; 64 for H = 256) whose recurrences
// are independent and run half a step apart.  Generation 2's step is a serial chain (h-part MMA -> gate math -> DSMEM
// exchange -> barrier hand-offs, ~6 k cycles measured) that leaves the MUFU pipe, the tensor pipe and the DSMEM fabric
// idle most of the time; here the 16 epilogue warps alternate A, B, A, B ... so that one chain's gate math covers the
// other chain's exchange + MMA.  Unlike generation 3 (two 64-row halves of ONE tile) every MMA is a full-rate M = 128
// instruction, the same resident weights serve twice the rows (half the clusters, one wave on cfg2), and:
//
//   * h is SINGLE-buffered (that is what makes two 128-row tiles fit beside the weights).  The write-after-read hazard
//     on h_{t-1} is closed by an explicit hand-shake: the MMA thread of every CTA multicasts a tcgen05.commit arrive
//     ("my h-part of step t has finished reading h_{t-1}") to the H_FREE barrier of all CTAs; the epilogue waits for
//     that phase -- one whole gate-math time old by then -- before it overwrites / pushes h_t.
//   * three accumulator buffers rotate over the slot sequence (A,0) (B,0) (A,1) (B,1) ...; the x-part of slot n+2 is
//     issued right behind the h-part of slot n, so the input projection never sits on a recurrence's critical path.
//
// TMEM: columns [0,64) cell state (32 per sub-tile), [128,512) three gate accumulators of 128 columns.
// Replaces nn.LSTM at FN-SSL/Lightning/Model.py:38,46 and IPDnet/FixedAarryIPDnet.py:32,36 (+ glue :35-37,41-45,49).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "tc_common.cuh"

namespace fnssl {
namespace tc4 {

// Build-time variants (A/B measured on the B200, profiles/r2_lstm_variants.txt):
//   TC4_PROD2 : a second single-lane TMA producer (warp 19) -- the x ring was producer-ISSUE bound (~300 instructions per slot)
//   TC4_PUB   : 1 or 2 publisher warps take the h exchange (DSMEM pushes) and the TMA output stores off the epilogue warps,
//               which then never meet on a named barrier nor wait for a bulk store to drain (0 = the epilogue publishes itself)
//   (more than 20 warps per CTA cap the kernel at 80 registers per thread: 6 warps on one SM sub-partition)
#ifndef TC4_PROD2
#define TC4_PROD2 0
#endif
#ifndef TC4_PUB
#define TC4_PUB 1
#endif
//   TC4_CREG  : the cell state lives in the epilogue threads' registers (8 or 4 values per sub-tile) instead of TMEM: one
//               tcgen05.ld less per slot and no tcgen05.st / wait::st on the epilogue's serial path
//   TC4_HFREE_EARLY : the "h_{t-1} has been read everywhere" barrier is polled (non-blocking test_wait) BEFORE the gate math so
//               that its ~100-cycle latency hides under the MUFU work; the blocking wait only runs if that poll failed
//   (both measured neutral-to-slightly-negative on the B200 -- the epilogue is MUFU-pipe bound, not latency bound -- and CREG
//   costs registers, so they are off by default)
//   TC4_XORDER : the x-part of slot n+2 is issued only after the h-part of slot n HAS BEEN ISSUED, so the tensor pipe runs
//               h(n) x(n+2) h(n+1) x(n+3) ...  Without it the x-part is issued as soon as its accumulator buffer drains, i.e.
//               shortly BEFORE h(n) becomes ready, and the latency-critical h-part queues behind up to 1 k cycles of x-part MMAs
#ifndef TC4_CREG
#define TC4_CREG 0
#endif
#ifndef TC4_HFREE_EARLY
#define TC4_HFREE_EARLY 0
#endif
//   TC4_HTILE : one "h_{t-1} tile has arrived" barrier per SOURCE CTA instead of one per sub-tile: the h-part MMAs of a chunk
//               are issued as soon as that chunk's tile is in place (own tile first, then the peers in push order), so the ~0.5 k
//               cycles of h-part issue overlap the tail of the DSMEM exchange instead of following it
#ifndef TC4_XORDER
#define TC4_XORDER 0
#endif
#ifndef TC4_PUSH_TILE
#define TC4_PUSH_TILE 0      // 1 (single publisher warp): ONE DSMEM bulk copy of the whole h tile per peer instead of one per quadrant (measured: -8 %, off)
#endif
#ifndef TC4_HTILE
#define TC4_HTILE 1
#endif
// timing experiments only (wrong results): FNSSL_TC_DEBUG bit 2 = one x-part MMA per slab instead of four, bit 4 = no TMA output
// stores, bit 16 = one h-part MMA per chunk instead of two
#ifndef TC4_EXPERIMENT
#define TC4_EXPERIMENT 0
#endif
constexpr int kXWarp = 18;
constexpr int kProd2Warp = TC4_PROD2 ? 19 : -1;          // second TMA producer lane (odd ring stages + the L2 prefetch)
constexpr int kPubWarp0 = 19 + (TC4_PROD2 ? 1 : 0);      // publisher warps (kNumPub of them)
constexpr int kNumPub = TC4_PUB;
constexpr int kQPerPub = kNumPub ? 4 / kNumPub : 4;     // TMEM lane quadrants served by one publisher warp
static_assert(kNumPub == 0 || kNumPub == 1 || kNumPub == 2, "TC4_PUB in {0, 1, 2}");
// producer warp + h-part MMA warp + 16 epilogue warps + x-part MMA warp [+ 2nd producer warp] [+ 2 publisher warps]
constexpr int kThreads = 32 * (19 + (TC4_PROD2 ? 1 : 0) + kNumPub);
constexpr int kEpiThreads = 512;
constexpr int kSlabK = 64;
constexpr int kWSlab = 128 * 128;      // [128 gate columns x 64] fp16
constexpr int kChunkUnits = 32;
constexpr int kChunkN = 128;
constexpr int kMaxXSlabs = 6;
constexpr int kMaxXStages = 6;
constexpr int kAccBufs = 3;
constexpr int kSmemLimit = 232448;
constexpr int kNumBarsBase = 1 + 2 * kMaxXStages + 3 * kAccBufs + 4;
constexpr int kMaxC = 8;                        // largest cluster (H = 256)
constexpr int kNumBars = kNumBarsBase + 4 + 8 + kAccBufs + 2 * kMaxC;   // + X2_FULL / X2_EMPTY of the narrow second source's ring, + H_READY[sub][quadrant], + HP_ISSUED[acc buffer],
                                                                    // + H_TILE[sub][source CTA] (TC4_HTILE)
constexpr int kWSmall = kChunkN * 32;          // [128 gate columns x 16] fp16 weight slab of a narrow source (32B swizzle)

struct Params {
  int nxs;
  uint32_t xs_srcmask;             // bit j: slab j comes from src1
  unsigned long long xs_k0pack;    // byte j: first channel of slab j / 16
  uint32_t xs_nkpack;              // nibble j: K=16 steps of slab j (1..4)
  int xstages;
  int steps, axis, nf, nt;
  long long rows;
  int tiles_per_b;
  const float* bias;               // [dirs][4H], accumulator column order [chunk][gate][unit]
  __half* out0; int out0_ld; int out0_off;
  const __half* addend; int addend_ld;
  __half* out1; int out1_ld;
  float* h_state; float* c_state; int state_flags;   // optional carried state, fp32 (rows, H); dirs == 1 (see lstm_tc2.cu)
  int store_tile;                  // 1: one [SUB x 32] output box per tile and destination instead of one per quadrant (see the publisher)
  int tma_out;                     // bit 0: out0 is written by TMA tile stores out of the h exchange tiles; bit 1: out1 aliases
                                   // addend and is produced in place by TMA reduce-add (global += h)
  int* error_flag;
  long long* trace;                // FNSSL_TC_TRACE: clock64 stamps of CTA (0,0): [slot 16..31][16 events]
  int debug;                       // experiments only (FNSSL_TC_DEBUG): 1 = skip gate math, 2 = skip MMA issue, 32 = ex2/rcp gates instead of tanh.approx, 64 = no L2 prefetch
};

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One (row, unit): c' = sigmoid(f) c + sigmoid(i) tanh(g), h = sigmoid(o) tanh(c').  sigmoid(x) = 1/(1+2^(-x log2 e)),
// tanh as (1-E)/(1+E); the cell update runs over ONE reciprocal (5 ex2 + 2 rcp per element, see lstm_tc2.cu).
__device__ __forceinline__ float lstm_cell(float gi, float gf, float gg, float go, float& c) {
  const float kL2E = 1.4426950408889634f;
  const float xg = fminf(fmaxf(gg, -15.f), 15.f);
  const float ei = ex2_approx(-kL2E * fmaxf(gi, -20.f));
  const float ef = ex2_approx(-kL2E * fmaxf(gf, -20.f));
  const float eg = ex2_approx(-2.0f * kL2E * xg);
  const float eo = ex2_approx(-kL2E * go);
  const float ab = (1.0f + ei) * (1.0f + eg);
  const float ff = 1.0f + ef;
  const float cn = fmaf(c, ab, (1.0f - eg) * ff) * rcp_approx(ff * ab);
  c = cn;
  const float ec = ex2_approx(-2.0f * kL2E * fminf(fmaxf(cn, -15.f), 15.f));
  return (1.0f - ec) * rcp_approx((1.0f + eo) * (1.0f + ec));
}
// experiment (FNSSL_TC_DEBUG & 32): 5 MUFU.TANH per element instead of 5 EX2 + 2 RCP (2^-11 relative error per tanh)
__device__ __forceinline__ float lstm_cell_tanh(float gi, float gf, float gg, float go, float& c) {
  const float si = fmaf(0.5f, tanh_approx(0.5f * gi), 0.5f);
  const float sf = fmaf(0.5f, tanh_approx(0.5f * gf), 0.5f);
  const float so = fmaf(0.5f, tanh_approx(0.5f * go), 0.5f);
  const float cn = fmaf(sf, c, si * tanh_approx(gg));
  c = cn;
  return so * tanh_approx(cn);
}

// NARROW: the second source is <= 16 channels wide -- its x slabs travel through a two-entry ring of their own (32B swizzle,
// 4 KB weight slab) and every CTA fetches its own x slabs (no multicast, local "slot empty" hand-shake); see make_plan.
template <int H, int SUB, bool TRACE, bool NARROW>
__global__ void __launch_bounds__(kThreads, 1)
lstm_tc4_kernel(const __grid_constant__ CUtensorMap map_src0, const __grid_constant__ CUtensorMap map_src1,
                const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w2,
                const __grid_constant__ CUtensorMap map_out0, const __grid_constant__ CUtensorMap map_out1, const Params p) {
  constexpr int C = H / kChunkUnits;      // cluster size == number of 32-unit chunks
  constexpr int NHS = H / kSlabK;         // K slabs of h in the weight layout
  constexpr int kXSlab = SUB * 128;       // one [SUB x 64] fp16 x slab (128B swizzle)
  constexpr int kHTile = SUB * 64;        // one [SUB x 32] fp16 h tile (one chunk, 64B swizzle)
  constexpr int kXSmall = SUB * 32;       // one [SUB x 16] fp16 x slab of a narrow second source (32B swizzle)
  constexpr int kTileRows = 2 * SUB;      // rows of a cluster tile (two sub-tiles)
  constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kChunkN >> 3) << 17) | ((uint32_t)(SUB >> 4) << 24);
  static_assert(H == 64 || H == 128 || H == 256, "H in {64,128,256}");
  static_assert(SUB == 64 || SUB == 128, "SUB in {64,128}");

  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) unsigned long long bars[kNumBars];
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;
  const uint32_t rank = cluster_ctarank();          // == chunk owned by this CTA
  const int tile = blockIdx.x / C;
  const uint16_t mask = (uint16_t)((1u << C) - 1u);
  const int nxs = p.nxs, XS = p.xstages, L = p.steps;
  constexpr bool small1 = NARROW, nomc = NARROW;
  const int nxb = small1 ? nxs - 1 : nxs;                           // x slabs that travel through the big ring
  const int nslots = 2 * L;

  const uint32_t dyn0 = (smem_addr(smem_dyn) + 1023u) & ~1023u;
  const uint32_t w_base = dyn0;                                     // resident weights: nxb x slabs, then NHS h slabs
  const uint32_t w2_base = w_base + (uint32_t)(nxb + NHS) * kWSlab; // [128 x 16] slab of the narrow second source
  const uint32_t hs_base = w2_base + (small1 ? (uint32_t)kWSmall : 0u);   // h operand: [sub][chunk] tiles (single buffer)
  const uint32_t xr_base = hs_base + 2u * C * kHTile;               // x ring: XS slabs
  const uint32_t x2_base = xr_base + (uint32_t)XS * kXSlab;         // narrow-source ring: 2 slabs of kXSmall
  const uint32_t bias_base = x2_base + (small1 ? 2u * kXSmall : 0u);   // 128 floats
  float* bias_s = reinterpret_cast<float*>(smem_dyn + (bias_base - smem_addr(smem_dyn)));

  const uint32_t bar0 = smem_addr(bars);
  const uint32_t W_FULL = bar0;
  auto X_FULL = [&](int i) { return bar0 + 8u * (1 + i); };
  auto X_EMPTY = [&](int i) { return bar0 + 8u * (1 + kMaxXStages + i); };
  auto ACC_FULL = [&](int i) { return bar0 + 8u * (1 + 2 * kMaxXStages + i); };
  auto ACC_EMPTY = [&](int i) { return bar0 + 8u * (1 + 2 * kMaxXStages + kAccBufs + i); };
  auto H_FULL = [&](int sub) { return bar0 + 8u * (1 + 2 * kMaxXStages + 2 * kAccBufs + sub); };
  auto H_FREE = [&](int sub) { return bar0 + 8u * (1 + 2 * kMaxXStages + 2 * kAccBufs + 2 + sub); };
  auto XP_DONE = [&](int i) { return bar0 + 8u * (1 + 2 * kMaxXStages + 2 * kAccBufs + 4 + i); };
  auto X2_FULL = [&](int i) { return bar0 + 8u * (kNumBarsBase + i); };
  auto X2_EMPTY = [&](int i) { return bar0 + 8u * (kNumBarsBase + 2 + i); };
  auto H_READY = [&](int sub, int q) { return bar0 + 8u * (kNumBarsBase + 4 + sub * 4 + q); };   // quadrant q of h_t(sub) is in shared memory
  auto HP_ISSUED = [&](int i) { return bar0 + 8u * (kNumBarsBase + 12 + i); };
  // TC4_HTILE: chunk `src` of h_{t-1}(sub) is in this CTA's shared memory (own chunk: 4 quadrant arrives; a peer's: tx bytes)
  auto H_TILE = [&](int sub, int src) { return bar0 + 8u * (kNumBarsBase + 12 + kAccBufs + sub * kMaxC + src); };
  // the barrier the quadrant pushes of CTA `src` complete on / the local quadrant arrives go to
  auto H_IN = [&](int sub, int src) { return TC4_HTILE ? H_TILE(sub, src) : H_FULL(sub); };                    // the h-part of the slot using buffer i has been issued

  if (tid == 0) {
    mbar_init(W_FULL, 1);
    const int xe = nomc ? 1 : C;
    for (int i = 0; i < kMaxXStages; ++i) { mbar_init(X_FULL(i), 1); mbar_init(X_EMPTY(i), xe); }
    for (int i = 0; i < kAccBufs; ++i) {
      mbar_init(ACC_FULL(i), 1); mbar_init(ACC_EMPTY(i), kEpiThreads / 32); mbar_init(XP_DONE(i), 1); mbar_init(HP_ISSUED(i), 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(X2_FULL(i), 1); mbar_init(X2_EMPTY(i), xe); }
    for (int sub = 0; sub < 2; ++sub) {
      mbar_init(H_FULL(sub), 5);    // MMA thread's expect_tx + 4 local quadrants (+ tx bytes of the C-1 remote tiles)
      for (int src = 0; src < C; ++src) mbar_init(H_TILE(sub, src), (uint32_t)src == rank ? 4 : 1);
      mbar_init(H_FREE(sub), C + kNumPub);    // one multicast commit per CTA of the cluster (+ the publishers: stores drained)
      for (int q = 0; q < 4; ++q) mbar_init(H_READY(sub, q), 4);   // the four epilogue warps of a quadrant
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&map_src0); prefetch_tmap(&map_src1); prefetch_tmap(&map_w); prefetch_tmap(&map_w2); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_base_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < kChunkN; i += kThreads) bias_s[i] = p.bias[dir * 4 * H + rank * kChunkN + i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // every CTA's barriers are initialised before any multicast / remote traffic can reach them
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t tmem_c = tmem;             // cell state: columns [32 sub, 32 sub + 32)
  const uint32_t tmem_acc = tmem + 128;     // gate accumulators: kAccBufs buffers x 128 columns

  int coord_b = 0, coord_r0 = 0;
  long long row0;
  int valid_rows;
  if (p.axis == FNSSL_ALONG_FREQ) {
    row0 = (long long)tile * kTileRows;
    coord_r0 = (int)row0;
    valid_rows = (int)min((long long)kTileRows, p.rows - row0);
  } else {
    coord_b = tile / p.tiles_per_b;
    coord_r0 = (tile % p.tiles_per_b) * kTileRows;
    row0 = (long long)coord_b * p.nf + coord_r0;
    valid_rows = min(kTileRows, p.nf - coord_r0);
  }
  const bool tr_cta = TRACE && p.trace && blockIdx.x == 0 && blockIdx.y == 0;
  const int hoff = (p.state_flags & 1) ? 1 : 0;    // a carried state adds an h-part (and an H_FREE phase) at step 0

  if (hoff) {
    // Resume from a carried state: h_{-1} takes the place step -1 would have written, i.e. all C tiles of both sub-tiles
    // (every CTA needs the whole h vector of its rows); c_{-1} is loaded by the epilogue warps below.
    constexpr int PPR = H / 8;                 // 16-byte pieces (8 units) per row
    for (int idx = tid; idx < kTileRows * PPR; idx += kThreads) {
      const int R = idx / PPR, pc = idx % PPR, kc = pc >> 2, sb = pc & 3;
      const int sub = R / SUB, r = R % SUB;
      uint4 pk = make_uint4(0, 0, 0, 0);
      if (R < valid_rows) {
        const float4* g = reinterpret_cast<const float4*>(p.h_state + (row0 + R) * H + pc * 8);
        const float4 a = __ldg(g), b = __ldg(g + 1);
        __half2 h01 = __floats2half2_rn(a.x, a.y), h23 = __floats2half2_rn(a.z, a.w);
        __half2 h45 = __floats2half2_rn(b.x, b.y), h67 = __floats2half2_rn(b.z, b.w);
        pk.x = *reinterpret_cast<uint32_t*>(&h01); pk.y = *reinterpret_cast<uint32_t*>(&h23);
        pk.z = *reinterpret_cast<uint32_t*>(&h45); pk.w = *reinterpret_cast<uint32_t*>(&h67);
      }
      st_shared_v4(hs_base + (uint32_t)(sub * C + kc) * kHTile + (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u +
                       (uint32_t)((sb ^ ((r >> 1) & 3)) << 4), pk);
    }
    fence_async_smem();
    __syncthreads();
  }

  // The three single-lane roles below are instruction-latency bound (one warp executing a serial program at ~8 cycles
  // per instruction; ncu: 560 warp instructions per slot in the first version of this loop = 5 k cycles, more than the
  // tensor pipe's 1.5 k), so they are kept lean: the elected lane runs the whole loop (ptxas emits tcgen05 / TMA
  // instructions straight-line behind an elect.sync predicate), descriptors are base + offset adds, and the x-part and
  // the h-part are issued by two different warps.
  if (warp == 0 || warp == kProd2Warp) {
    // ============================== TMA producers ==============================
    // Two single-lane producers walk the same slab sequence (stage / phase / fetcher counters advance for every slab) but
    // producer `pid` only ACTS on the ring stages with (stage & 1) == pid, so the per-slab issue work is split in two.  The
    // split is by STAGE, not by slab: all uses of one stage must be refilled by the same lane in program order, otherwise a
    // lane running ahead of the other could pass a parity wait one phase early (mbarrier parity only disambiguates
    // consecutive phases).
    // Producer 0 also loads the weights and feeds the narrow-source ring; producer 1 also issues the L2 prefetches.
    const int pid = TC4_PROD2 ? (warp == 0 ? 0 : 1) : 2;     // 2 = the only producer: acts on every stage
    if (elect_one()) {
      if (pid != 1) {
      mbar_expect_tx(W_FULL, (uint32_t)(nxb + NHS) * kWSlab + (small1 ? (uint32_t)kWSmall : 0u));
      for (int j = 0; j < nxb; ++j)
        tma_load_2d(w_base + j * kWSlab, &map_w, W_FULL, j * kSlabK, (dir * C + (int)rank) * kChunkN);
      for (int j = 0; j < NHS; ++j)     // the h slabs follow ALL x slabs in the packed buffer
        tma_load_2d(w_base + (nxb + j) * kWSlab, &map_w, W_FULL, (nxs + j) * kSlabK, (dir * C + (int)rank) * kChunkN);
      if (small1) tma_load_2d(w2_base, &map_w2, W_FULL, nxb * kSlabK, (dir * C + (int)rank) * kChunkN);
      }
      int stage = 0, fetcher = 0;
      uint32_t phase = 0;                 // parity of the X_EMPTY wait; the first pass over the ring does not wait
      bool wrapped = false;
      const bool along_f = p.axis == FNSSL_ALONG_FREQ;
      // x_t comes from HBM (~3.4 k cycles per TMA round trip, the whole ring holds one slot): the CTA that will fetch a
      // slab pulls it into L2 kAhead slots earlier, so the ring's real loads are L2 hits
      const int ahead = (p.debug & 64) ? 0 : ((p.debug >> 8) & 15 ? (p.debug >> 8) & 15 : 4);
      auto prefetch_slot = [&](int nn) {
        const int tt = nn >> 1;
        const int ss = dir ? (L - 1 - tt) : tt;
        const int rr0 = coord_r0 + (nn & 1) * SUB;
        // slab j of slot nn is fetched by CTA (nn * nxs + j) % C: visit only this CTA's slabs
        for (int j = (int)((rank + (uint32_t)C - (uint32_t)((nn * nxs) % C)) % (uint32_t)C); j < nxs; j += C) {
          const CUtensorMap* m = ((p.xs_srcmask >> j) & 1) ? &map_src1 : &map_src0;
          const int k0 = (int)((p.xs_k0pack >> (8 * j)) & 0xff) * 16;
          if (along_f) tma_prefetch_l2_4d(m, k0, ss, rr0, 0);
          else tma_prefetch_l2_4d(m, k0, rr0, ss, coord_b);
        }
      };
      if (pid != 0) for (int nn = 1; nn < ahead && nn < nslots; ++nn) prefetch_slot(nn);
      for (int n = 0; n < nslots; ++n) {
        const int t = n >> 1, sub = n & 1;
        const int s = dir ? (L - 1 - t) : t;
        const int r0 = coord_r0 + sub * SUB;
        if (pid != 0 && ahead && n + ahead < nslots) prefetch_slot(n + ahead);
        for (int j = 0; j < nxb; ++j) {
          if (pid == 2 || (stage & 1) == pid) {
          if (wrapped) mbar_wait(X_EMPTY(stage), phase, p.error_flag, 100 + stage);
          mbar_expect_tx(X_FULL(stage), kXSlab);
          if (nomc) {
            const CUtensorMap* m = ((p.xs_srcmask >> j) & 1) ? &map_src1 : &map_src0;
            const uint32_t dst = xr_base + (uint32_t)stage * kXSlab;
            const int k0 = (int)((p.xs_k0pack >> (8 * j)) & 0xff) * 16;
            if (along_f) tma_load_4d(dst, m, X_FULL(stage), k0, s, r0, 0);
            else tma_load_4d(dst, m, X_FULL(stage), k0, r0, s, coord_b);
          } else if ((uint32_t)fetcher == rank) {   // one CTA fetches the slab for the whole cluster
            const CUtensorMap* m = ((p.xs_srcmask >> j) & 1) ? &map_src1 : &map_src0;
            const uint32_t dst = xr_base + (uint32_t)stage * kXSlab;
            const int k0 = (int)((p.xs_k0pack >> (8 * j)) & 0xff) * 16;
            if (along_f) tma_load_4d_mc(dst, m, X_FULL(stage), k0, s, r0, 0, mask);
            else tma_load_4d_mc(dst, m, X_FULL(stage), k0, r0, s, coord_b, mask);
          }
          }
          if (++fetcher == C) fetcher = 0;
          if (++stage == XS) { stage = 0; phase ^= wrapped ? 1u : 0u; wrapped = true; }
        }
        if (small1 && pid == 1) { if (++fetcher == C) fetcher = 0; }
        if (small1 && pid != 1) {      // the narrow source's slab of this slot: entry n & 1 of its own ring, use number n >> 1
          const int s2 = n & 1;
          if (n >= 2) mbar_wait(X2_EMPTY(s2), (uint32_t)(((n >> 1) - 1) & 1), p.error_flag, 110 + s2);
          mbar_expect_tx(X2_FULL(s2), kXSmall);
          if (nomc) {
            const uint32_t dst = x2_base + (uint32_t)s2 * kXSmall;
            if (along_f) tma_load_4d(dst, &map_src1, X2_FULL(s2), 0, s, r0, 0);
            else tma_load_4d(dst, &map_src1, X2_FULL(s2), 0, r0, s, coord_b);
          } else if ((uint32_t)fetcher == rank) {
            const uint32_t dst = x2_base + (uint32_t)s2 * kXSmall;
            if (along_f) tma_load_4d_mc(dst, &map_src1, X2_FULL(s2), 0, s, r0, 0, mask);
            else tma_load_4d_mc(dst, &map_src1, X2_FULL(s2), 0, r0, s, coord_b, mask);
          }
          if (++fetcher == C) fetcher = 0;
        }
      }
    }
    __syncwarp();
  } else if (warp == kXWarp) {
    // ============================== x-part MMA issuer ==============================
    // G_x of slot n -> accumulator buffer n % 3, up to three slots ahead of the recurrences (bounded by ACC_EMPTY and
    // the x ring); XP_DONE tells the h-part issuer that the buffer holds the complete input projection.
    if (elect_one()) {
      mbar_wait(W_FULL, 0, p.error_flag, 200);
      const uint64_t a_desc0 = make_sw128_desc(xr_base);
      const uint64_t b_desc0 = make_sw128_desc(w_base);
      int xstage = 0, a = 0;
      uint32_t xphase = 0, empty_par = 0;
      [[maybe_unused]] uint32_t hp_par = 0;
      for (int n = 0; n < nslots; ++n) {
        long long* tp = (TRACE && tr_cta && n >= 16 && n < 32) ? p.trace + (n - 16) * 16 : nullptr;
        long long w_acc = 0, e_acc = 0;
        if (TRACE && tp) { tp[8] = clock64(); e_acc = clock64(); }
        if (n >= kAccBufs) {
          mbar_wait(ACC_EMPTY(a), (empty_par >> a) & 1u, p.error_flag, 201 + a);
          empty_par ^= 1u << a;
        }
        if (TC4_XORDER && n >= 2) {     // queue behind the h-part of slot n-2 (buffer (n-2) % 3 == (a+1) % 3), never ahead of it
          const int b = (a == kAccBufs - 1) ? 0 : a + 1;
          mbar_wait(HP_ISSUED(b), (hp_par >> b) & 1u, p.error_flag, 205 + b);
          hp_par ^= 1u << b;
        }
        if (TRACE && tp) e_acc = clock64() - e_acc;
        tc_fence_after();
        const uint32_t d_tmem = tmem_acc + (uint32_t)a * kChunkN;
        uint32_t nkp = p.xs_nkpack;
        for (int j = 0; j < nxb; ++j, nkp >>= 4) {
          const long long c1 = (TRACE && tp) ? clock64() : 0;
          mbar_wait(X_FULL(xstage), xphase, p.error_flag, 210 + xstage);
          if (TRACE && tp) w_acc += clock64() - c1;
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + (uint64_t)(xstage * (kXSlab >> 4));
          const uint64_t b_desc = b_desc0 + (uint64_t)(j * (kWSlab >> 4));
          const uint32_t nk = nkp & 15u;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if ((uint32_t)k < nk && !(TC4_EXPERIMENT && (p.debug & 2) && k > 0))
              umma_f16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, kIdesc, (uint32_t)(j | k));
          if (nomc) umma_commit(X_EMPTY(xstage));
          else umma_commit_mc(X_EMPTY(xstage), mask);    // this CTA is done with the slab: tell every CTA's ring
          if (++xstage == XS) { xstage = 0; xphase ^= 1u; }
        }
        if (small1) {      // one K = 16 step against the 4 KB weight slab, both operands in the 32B-swizzled layout
          const int s2 = n & 1;
          mbar_wait(X2_FULL(s2), (uint32_t)((n >> 1) & 1), p.error_flag, 216 + s2);
          tc_fence_after();
          umma_f16(d_tmem, make_sw32_desc(x2_base + (uint32_t)s2 * kXSmall), make_sw32_desc(w2_base), kIdesc, 1u);
          if (nomc) umma_commit(X2_EMPTY(s2));
          else umma_commit_mc(X2_EMPTY(s2), mask);
        }
        umma_commit(XP_DONE(a));
        if (TRACE && tp) { tp[9] = clock64(); tp[11] = e_acc; tp[12] = w_acc; }
        a = (a == kAccBufs - 1) ? 0 : a + 1;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================== h-part MMA issuer (the recurrences' critical path) ==============================
    if (elect_one()) {
      mbar_wait(W_FULL, 0, p.error_flag, 230);
      const uint64_t h_desc0 = make_sw64_desc(hs_base);
      const uint64_t wh_desc0 = make_sw128_desc(w_base + (uint32_t)nxb * kWSlab);
      int a = 0;
      uint32_t xp_par = 0;
      // per-chunk operand descriptor offsets and arrival barriers in the order the chunks are consumed (constant per CTA)
      uint32_t a_off[C], b_off[C], tbar0[C], tbar1[C];
#pragma unroll
      for (int i = 0; i < C; ++i) {
        const int kc = TC4_HTILE ? (int)((rank + (uint32_t)(C - i)) % C) : i;   // own chunk first, then rank-1, rank-2, ... (push order)
        a_off[i] = (uint32_t)(kc * (kHTile >> 4));
        b_off[i] = (uint32_t)((kc >> 1) * (kWSlab >> 4) + 4 * (kc & 1));         // W columns inside 128B-swizzled slab kc/2
        tbar0[i] = H_TILE(0, kc);
        tbar1[i] = H_TILE(1, kc);
      }
      for (int n = 0; n < nslots; ++n) {
        const int t = n >> 1, sub = n & 1;
        long long* tp = (TRACE && tr_cta && n >= 16 && n < 32) ? p.trace + (n - 16) * 16 : nullptr;
        if (TRACE && tp) tp[0] = clock64();
        mbar_wait(XP_DONE(a), (xp_par >> a) & 1u, p.error_flag, 240 + a);     // G_x of this slot is complete
        xp_par ^= 1u << a;
        if (t > 0 || hoff) {
          if (t > 0 && !TC4_HTILE) {
            // h_{t-1} of this sub-tile: C-1 remote tiles arrive as DSMEM bulk copies (tx bytes), the local one by arrives
            mbar_expect_tx(H_FULL(sub), (uint32_t)((C - 1) * kHTile));
            mbar_wait_cluster(H_FULL(sub), (uint32_t)((t - 1) & 1), p.error_flag, 220 + sub);
          }   // t == 0 with a carried state: h_{-1} was placed in the tiles before the roles split
          if (t > 0 && TC4_HTILE) {     // arm every remote tile's barrier first, then take the tiles as they arrive
#pragma unroll
            for (int i = 1; i < C; ++i) mbar_expect_tx(sub ? tbar1[i] : tbar0[i], (uint32_t)kHTile);
          }
          if (TRACE && tp) tp[1] = clock64();
          tc_fence_after();
          const uint32_t d_tmem = tmem_acc + (uint32_t)a * kChunkN;
          const uint64_t a_sub = h_desc0 + (uint64_t)(sub * C * (kHTile >> 4));
#pragma unroll
          for (int i = 0; i < C; ++i) {   // K = 32 units of one chunk: two K=16 steps
            if (TC4_HTILE && t > 0) {
              mbar_wait_cluster(sub ? tbar1[i] : tbar0[i], (uint32_t)((t - 1) & 1), p.error_flag, 220 + sub);
              tc_fence_after();
            }
            const uint64_t a_desc = a_sub + (uint64_t)a_off[i];
            const uint64_t b_desc = wh_desc0 + (uint64_t)b_off[i];
#pragma unroll
            for (int k = 0; k < 2; ++k)
              if (!(TC4_EXPERIMENT && (p.debug & 16) && k > 0)) umma_f16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, kIdesc, 1u);
          }
          // h_{t-1} has been read once these MMAs retire: every CTA may then overwrite its copy with h_t
          umma_commit_mc(H_FREE(sub), mask);
        } else {
          tc_fence_after();
        }
        umma_commit(ACC_FULL(a));            // fires when the h-part (and G_x, complete since XP_DONE) is done
        if (TC4_XORDER) mbar_arrive(HP_ISSUED(a));   // the x-part issuer may now queue G_x of slot n+2 behind these MMAs
        if (TRACE && tp) tp[2] = clock64();
        a = (a == kAccBufs - 1) ? 0 : a + 1;
      }
    }
    __syncwarp();
  } else if (TC4_PUB && warp >= kPubWarp0) {
    // ============================== publisher warps ==============================
    // One elected lane per warp serves kQPerPub TMEM lane quadrants: as soon as a quadrant's four epilogue warps have written
    // their part of h_t into the CTA's own exchange tile (H_READY), it pushes that 1-2 KB region to every peer (DSMEM bulk
    // copies completing tx bytes on the peers' H_FULL), arrives locally, and issues the TMA tile store / reduce-add of the
    // same region to HBM.  The epilogue warps therefore never wait for each other or for a bulk operation.
    // Write-after-read on the tile: the stores of slot n (committed as one bulk group) must have READ the tile before the
    // epilogue overwrites it at slot n + 2; the publisher checks that after issuing slot n + 1 (wait_group.read 1) and then
    // arrives on H_FREE(sub(n)) -- the barrier the epilogue already waits on for "every CTA's h-part has read h_{t-1}".
    const int pw = warp - kPubWarp0;
    if (elect_one()) {
      constexpr uint32_t kQuadBytes = (SUB == 128) ? 2048u : 1024u;
      constexpr int kQuadRows = SUB / 4;
      const bool along_f = p.axis == FNSSL_ALONG_FREQ;
      const int out_c = p.out0_off + dir * H + (int)rank * kChunkUnits;
      const int out1_c = dir * H + (int)rank * kChunkUnits;
      uint32_t peer_hs[C > 1 ? C - 1 : 1], peer_bar0[C > 1 ? C - 1 : 1], peer_bar1[C > 1 ? C - 1 : 1];
#pragma unroll
      for (int dd = 1; dd < C; ++dd) {
        const uint32_t d = (rank + (uint32_t)dd) % C;
        peer_hs[dd - 1] = mapa_shared(hs_base, d);
        peer_bar0[dd - 1] = mapa_shared(H_IN(0, (int)rank), d);
        peer_bar1[dd - 1] = mapa_shared(H_IN(1, (int)rank), d);
      }
      if (hoff) { mbar_arrive(H_FREE(0)); mbar_arrive(H_FREE(1)); }   // the first phase has no earlier stores to wait for
      const bool tma_any = p.tma_out != 0 && !(TC4_EXPERIMENT && (p.debug & 4));
      for (int n = 0; n < nslots; ++n) {
        const int t = n >> 1, sub = n & 1;
        const int s = dir ? (L - 1 - t) : t;
        const bool push = t + 1 < L;
        if (push || tma_any) {
          // the exchange first (it sits on the recurrence's critical path), the HBM stores of all quadrants afterwards: issuing a
          // TMA store / reduce-add costs this single lane ~50 cycles each, which used to delay the next quadrant's pushes
          // (measured: 0.82 -> 0.73 ms per 256-channel layer without the stores, profiles/r2_lstm_variants.txt)
          if (TC4_PUSH_TILE && kQPerPub == 4) {
            // the four quadrants finish within a few hundred cycles of each other, and every bulk copy costs this lane ~100-170
            // cycles to issue: 3 copies of the whole tile instead of 12 quadrant copies per slot
#pragma unroll
            for (int q = 0; q < 4; ++q) mbar_wait(H_READY(sub, q), (uint32_t)(t & 1), p.error_flag, 400 + sub * 4 + q);
            if (push) {
              const uint32_t off = (uint32_t)(sub * C + (int)rank) * kHTile;
#pragma unroll
              for (int dd = 1; dd < C; ++dd)
                bulk_copy_s2c(peer_hs[dd - 1] + off, hs_base + off, (uint32_t)kHTile, sub ? peer_bar1[dd - 1] : peer_bar0[dd - 1]);
#pragma unroll
              for (int q = 0; q < 4; ++q) mbar_arrive(H_IN(sub, (int)rank));     // the local copy of every quadrant is in place
            }
          } else {
#pragma unroll
          for (int qq = 0; qq < kQPerPub; ++qq) {
            const int q = kQPerPub * pw + qq;
            mbar_wait(H_READY(sub, q), (uint32_t)(t & 1), p.error_flag, 400 + sub * 4 + q);
            if (push) {
              const uint32_t off = (uint32_t)(sub * C + (int)rank) * kHTile + (uint32_t)q * kQuadBytes;
#pragma unroll
              for (int dd = 1; dd < C; ++dd)
                bulk_copy_s2c(peer_hs[dd - 1] + off, hs_base + off, kQuadBytes, sub ? peer_bar1[dd - 1] : peer_bar0[dd - 1]);
              mbar_arrive(H_IN(sub, (int)rank));     // the local copy of this quadrant is in place
            }
          }
          }
          if (tma_any) {
            // Whole-tile boxes ([SUB x 32], one per destination; Params::store_tile) for small grids (latency-bound: fewer publisher
            // operations per slot, one utterance 1.25 -> 1.16 ms per forward) and when the layer has TWO plain destinations (out0 + the second copy of h):
            // the lane needs ~170 cycles per TMA store and eight of them per slot made the publisher the slowest role (16-channel layer:
            // 0.56 -> 0.47 ms for one utterance, 0.75 -> 0.72 at B = 16).  With one destination the quadrant boxes stay: whole-tile boxes
            // measured +3..6 % on the 256-channel layers at cfg2 size (profiles/r2_lstm_variants.txt, calls 45-47).
            const int kStores = (p.store_tile && kQPerPub == 4) ? 1 : kQPerPub;
#pragma unroll 1
            for (int qq = 0; qq < kStores; ++qq) {
              const int q = kQPerPub * pw + qq;
              const uint32_t off = (uint32_t)(sub * C + (int)rank) * kHTile + (uint32_t)q * kQuadBytes;
              const int r0 = coord_r0 + sub * SUB + q * kQuadRows;
              if (p.tma_out & 1) {
                if (along_f) tma_store_4d(&map_out0, hs_base + off, out_c, s, r0, 0);
                else tma_store_4d(&map_out0, hs_base + off, out_c, r0, s, coord_b);
              }
              if (p.tma_out & 2) {
                if (along_f) tma_reduce_add_4d(&map_out1, hs_base + off, out1_c, s, r0, 0);
                else tma_reduce_add_4d(&map_out1, hs_base + off, out1_c, r0, s, coord_b);
              }
              if (p.tma_out & 4) {      // out1 = a second copy of h (no residual operand)
                if (along_f) tma_store_4d(&map_out1, hs_base + off, out1_c, s, r0, 0);
                else tma_store_4d(&map_out1, hs_base + off, out1_c, r0, s, coord_b);
              }
            }
            bulk_commit_group();
          }
        }
        if (n >= 1) {     // slot n-1's stores (the other sub-tile's exchange tile) have read their source
          if (tma_any) bulk_wait_read_1();
          mbar_arrive(H_FREE(sub ^ 1));
        }
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue warps ==============================
    const int q = warp & 3;                    // TMEM lane quadrant of this warp
    const int sg = (warp - 2) >> 2;            // 8-unit group of the CTA's 32 units
    const int u0 = sg * 8;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float* bsp = bias_s + u0;            // bias of (gate, unit) at bsp[gate * 32 + e] (broadcast shared-memory reads)
    const long long sstride = (p.axis == FNSSL_ALONG_FREQ) ? 1 : p.nf;
    const bool fast = !(p.debug & 32);
    const bool tr = tr_cta && warp == 2 && lane == 0;
    constexpr uint32_t kQuadBytes = (SUB == 128) ? 2048u : 1024u;   // a quadrant's rows x 64 B of the CTA's own h tile
    [[maybe_unused]] const uint32_t hquad = (uint32_t)rank * kHTile + (uint32_t)q * kQuadBytes;

    // publish the quadrant's region of the CTA's own h tile: 4 warps meet, one thread pushes it to every peer
    // The same shared-memory tile feeds the HBM outputs: one TMA tile store per quadrant writes h_t (out0), one TMA
    // reduce-add accumulates it onto the residual operand in place (out1 == addend).  Per-thread global accesses of this
    // epilogue are 16-byte pieces of 32 different rows per warp instruction (32 LSU wavefronts each, ~1.5 k cycles of LSU
    // time per slot for addend + out0 + out1); the bulk copies cost no issue slots at all.
    constexpr int kQuadRows = SUB / 4;
    const bool pusher = elect_one() && sg == 0;      // one fixed lane per quadrant: issues the pushes and owns the bulk groups
    const int out_c = p.out0_off + dir * H + (int)rank * kChunkUnits;    // first channel of this CTA's tile in out0
    const int out1_c = dir * H + (int)rank * kChunkUnits;
    const bool along_f = p.axis == FNSSL_ALONG_FREQ;
    auto publish_quadrant = [&](uint32_t buf, int sub, int s, bool push) {
      fence_async_smem();
#if TC4_PUB
      // hand the quadrant to the publisher warp: one arrive per warp, no rendezvous of the epilogue warps
      __syncwarp();
      if (lane == 0) mbar_arrive(H_READY(sub, q));
#else
      named_bar_sync(1 + q, 128);
      {
        if (pusher) {
          if (push) {
            const uint32_t hb = H_IN(sub, (int)rank);
#pragma unroll
            for (int dd = 1; dd < C; ++dd) {
              const uint32_t d = (rank + (uint32_t)dd) % C;
              bulk_copy_s2c(mapa_shared(buf + hquad, d), buf + hquad, kQuadBytes, mapa_shared(hb, d));
            }
            mbar_arrive(hb);     // the local copy of this quadrant is in place
          }
          if (p.tma_out) {
            const int r0 = coord_r0 + sub * SUB + q * kQuadRows;
            if (p.tma_out & 1) {
              if (along_f) tma_store_4d(&map_out0, buf + hquad, out_c, s, r0, 0);
              else tma_store_4d(&map_out0, buf + hquad, out_c, r0, s, coord_b);
            }
            if (p.tma_out & 2) {
              if (along_f) tma_reduce_add_4d(&map_out1, buf + hquad, out1_c, s, r0, 0);
              else tma_reduce_add_4d(&map_out1, buf + hquad, out1_c, r0, s, coord_b);
            }
            if (p.tma_out & 4) {
              if (along_f) tma_store_4d(&map_out1, buf + hquad, out1_c, s, r0, 0);
              else tma_store_4d(&map_out1, buf + hquad, out1_c, r0, s, coord_b);
            }
            bulk_commit_group();
          }
        }
      }
#endif
    };
    const bool thr_out0 = p.out0 && !(p.tma_out & 1);     // per-thread stores (fallback paths)
    const bool thr_out1 = p.out1 && !(p.tma_out & 6);
    const bool thr_add = thr_out1 && p.addend != nullptr;      // (out1 without a residual operand is a second copy of h)
    const bool tma_any = p.tma_out != 0;

    int a = 0;
    uint32_t full_par = 0;

    if constexpr (SUB == 128) {
      // ---- M = 128: TMEM lane == row of the sub-tile; thread = (row, 8 hidden units)
      const int r = q * 32 + lane;
      const int ua = (int)rank * kChunkUnits + u0;      // absolute hidden unit of this thread's first element
      bool valid[2];
      long long base[2];
      uint4 addv_next[2];
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const int R = sub * SUB + r;
        valid[sub] = R < valid_rows;
        base[sub] = (p.axis == FNSSL_ALONG_FREQ) ? (row0 + R) * p.nf : ((long long)coord_b * p.nt * p.nf + coord_r0 + R);
        addv_next[sub] = make_uint4(0, 0, 0, 0);
        if (thr_add && valid[sub]) {
          const long long pos0 = base[sub] + (long long)(dir ? (L - 1) : 0) * sstride;
          addv_next[sub] = __ldg(reinterpret_cast<const uint4*>(p.addend + pos0 * p.addend_ld + dir * H + ua));
        }
      }
      // this thread's 16-byte h piece inside the CTA's own [128 x 32] tile (64B swizzle: chunk ^= (row >> 1) & 3)
      const uint32_t hpiece = (uint32_t)rank * kHTile + (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u +
                              (uint32_t)((sg ^ ((r >> 1) & 3)) << 4);
      [[maybe_unused]] float creg[2][8];      // TC4_CREG: this thread's cell state per sub-tile (the sub loop is unrolled)
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        float z[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = 0.0f;
        if (hoff && valid[sub]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) z[i] = __ldg(p.c_state + (row0 + sub * SUB + r) * H + ua + i);
        }
#if TC4_CREG
#pragma unroll
        for (int i = 0; i < 8; ++i) creg[sub][i] = z[i];
#else
        tmem_st8(tmem_c + (uint32_t)(sub * 32) + lane_off + u0, z);
#endif
      }
      if (!TC4_CREG) tmem_wait_st();
#pragma unroll 1
      for (int t = 0; t < L; ++t) {
        const int s = dir ? (L - 1 - t) : t;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          long long* tp = (tr && t >= 8 && t < 16) ? p.trace + ((t - 8) * 2 + sub) * 16 : nullptr;
          const long long pos = base[sub] + (long long)s * sstride;
          const uint4 addv = addv_next[sub];
          if (thr_add && valid[sub] && t + 1 < L) {     // residual operand of the next layer, fetched one step ahead
            const long long posn = base[sub] + (long long)(dir ? (L - 2 - t) : (t + 1)) * sstride;
            addv_next[sub] = __ldg(reinterpret_cast<const uint4*>(p.addend + posn * p.addend_ld + dir * H + ua));
          }
          if (!TC4_PUB && tma_any && pusher) bulk_wait_read_all();   // the tile stores of the previous slot have read their source
          if (TRACE && tp) tp[4] = clock64();
          mbar_wait(ACC_FULL(a), (full_par >> a) & 1u, p.error_flag, 300 + a);
          full_par ^= 1u << a;
          if (TRACE && tp) tp[5] = clock64();
          tc_fence_after();
          const uint32_t acc = tmem_acc + (uint32_t)a * kChunkN + lane_off + u0;
          [[maybe_unused]] const uint32_t cad = tmem_c + (uint32_t)(sub * 32) + lane_off + u0;
          float gti[8], gtf[8], gtg[8], gto[8];
#if TC4_CREG
          float (&cs)[8] = creg[sub];
#else
          float cs[8];
#endif
          tmem_ld8(acc + 0 * kChunkUnits, gti);
          tmem_ld8(acc + 1 * kChunkUnits, gtf);
          tmem_ld8(acc + 2 * kChunkUnits, gtg);
          tmem_ld8(acc + 3 * kChunkUnits, gto);
          if (!TC4_CREG) tmem_ld8(cad, cs);
          // poll "h_{t-1} has been read by every CTA" now: the answer is only needed after the gate math
          const bool will_publish = t + 1 < L || tma_any;
          bool hfree_ok = !(will_publish && t + hoff > 0);
          if (TC4_HFREE_EARLY && !hfree_ok) hfree_ok = mbar_test_wait_cluster(H_FREE(sub), (uint32_t)((t - 1 + hoff) & 1));
          tmem_wait_ld();
          tmem_ld_dep(gti); tmem_ld_dep(gtf); tmem_ld_dep(gtg); tmem_ld_dep(gto);
          if (!TC4_CREG) tmem_ld_dep(cs);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(ACC_EMPTY(a));      // accumulator drained (this warp): the MMA thread may produce G_x of slot n+3 into it
          a = (a == kAccBufs - 1) ? 0 : a + 1;
          float hv[8];
          if (p.debug & 1) {
#pragma unroll
            for (int e = 0; e < 8; ++e) hv[e] = gti[e] + gtf[e] + gtg[e] + gto[e] + cs[e];
          } else if (fast) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              hv[e] = lstm_cell_tanh(gti[e] + bsp[e], gtf[e] + bsp[kChunkUnits + e], gtg[e] + bsp[2 * kChunkUnits + e],
                                     gto[e] + bsp[3 * kChunkUnits + e], cs[e]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              hv[e] = lstm_cell(gti[e] + bsp[e], gtf[e] + bsp[kChunkUnits + e], gtg[e] + bsp[2 * kChunkUnits + e],
                                gto[e] + bsp[3 * kChunkUnits + e], cs[e]);
          }
          __half2 h01 = __floats2half2_rn(hv[0], hv[1]), h23 = __floats2half2_rn(hv[2], hv[3]);
          __half2 h45 = __floats2half2_rn(hv[4], hv[5]), h67 = __floats2half2_rn(hv[6], hv[7]);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h01); pk.y = *reinterpret_cast<uint32_t*>(&h23);
          pk.z = *reinterpret_cast<uint32_t*>(&h45); pk.w = *reinterpret_cast<uint32_t*>(&h67);
          if (TRACE && tp) tp[6] = clock64();
          if (will_publish) {
            // every CTA's h-part of step t has finished reading h_{t-1} (and with it all pushes of h_{t-1} have landed)
            const long long c2 = (TRACE && tp) ? clock64() : 0;
            if (!hfree_ok) mbar_wait_cluster(H_FREE(sub), (uint32_t)((t - 1 + hoff) & 1), p.error_flag, 320 + sub);
            if (TRACE && tp) tp[10] = clock64() - c2;
            const uint32_t buf = hs_base + (uint32_t)(sub * C) * kHTile;
            st_shared_v4(buf + hpiece, pk);
            publish_quadrant(buf, sub, s, t + 1 < L);
          }
          if (TRACE && tp) tp[7] = clock64();
          if (!TC4_CREG) tmem_st8(cad, cs);
          if ((p.state_flags & 2) && t + 1 == L && valid[sub]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              p.c_state[(row0 + sub * SUB + r) * H + ua + i] = cs[i];
              p.h_state[(row0 + sub * SUB + r) * H + ua + i] = __half2float(__float2half_rn(hv[i]));   // the fp16 value the next step would read
            }
          }
          if (valid[sub]) {
            if (thr_out0) *reinterpret_cast<uint4*>(p.out0 + pos * p.out0_ld + p.out0_off + dir * H + ua) = pk;
            if (thr_out1) {
              const __half2* av = reinterpret_cast<const __half2*>(&addv);
              __half2 o0 = __floats2half2_rn(hv[0] + __low2float(av[0]), hv[1] + __high2float(av[0]));
              __half2 o1 = __floats2half2_rn(hv[2] + __low2float(av[1]), hv[3] + __high2float(av[1]));
              __half2 o2 = __floats2half2_rn(hv[4] + __low2float(av[2]), hv[5] + __high2float(av[2]));
              __half2 o3 = __floats2half2_rn(hv[6] + __low2float(av[3]), hv[7] + __high2float(av[3]));
              uint4 ok;
              ok.x = *reinterpret_cast<uint32_t*>(&o0); ok.y = *reinterpret_cast<uint32_t*>(&o1);
              ok.z = *reinterpret_cast<uint32_t*>(&o2); ok.w = *reinterpret_cast<uint32_t*>(&o3);
              *reinterpret_cast<uint4*>(p.out1 + pos * p.out1_ld + dir * H + ua) = ok;
            }
          }
          if (!TC4_CREG) tmem_wait_st();
        }
      }
    } else {
      // ---- M = 64: the accumulator occupies lanes 0-15 of each quadrant.  16x256b TMEM accesses keep all 32 threads
      // busy: thread T owns rows {T/4, T/4+8} of the quadrant's 16 rows and units {2(T%4), 2(T%4)+1} of the 8-unit group.
      const int uo = 2 * (lane & 3);                       // unit offset inside the 8-unit group
      const int ua = (int)rank * kChunkUnits + u0 + uo;    // absolute hidden unit of the thread's first column
      const float bias_i[2] = {bsp[uo], bsp[uo + 1]};
      const float bias_f[2] = {bsp[kChunkUnits + uo], bsp[kChunkUnits + uo + 1]};
      const float bias_g[2] = {bsp[2 * kChunkUnits + uo], bsp[2 * kChunkUnits + uo + 1]};
      const float bias_o[2] = {bsp[3 * kChunkUnits + uo], bsp[3 * kChunkUnits + uo + 1]};
      long long base[2][2];
      bool valid[2][2];
      uint32_t hpiece[2];
      uint32_t addn[2][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int rs = q * 16 + (lane >> 2) + 8 * i;       // row inside the sub-tile
        hpiece[i] = (uint32_t)rank * kHTile + (uint32_t)(rs >> 3) * 512u + (uint32_t)(rs & 7) * 64u +
                    (uint32_t)((sg ^ ((rs >> 1) & 3)) << 4) + (uint32_t)uo * 2u;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int R = sub * SUB + rs;                    // row inside the cluster tile
          valid[sub][i] = R < valid_rows;
          base[sub][i] = (p.axis == FNSSL_ALONG_FREQ) ? (row0 + R) * p.nf : ((long long)coord_b * p.nt * p.nf + coord_r0 + R);
          addn[sub][i] = 0u;
          if (thr_add && valid[sub][i]) {
            const long long pos0 = base[sub][i] + (long long)(dir ? (L - 1) : 0) * sstride;
            addn[sub][i] = __ldg(reinterpret_cast<const unsigned int*>(p.addend + pos0 * p.addend_ld + dir * H + ua));
          }
        }
      }
      [[maybe_unused]] float creg[2][4];
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (hoff) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (valid[sub][e >> 1])
              z[e] = __ldg(p.c_state + (row0 + sub * SUB + q * 16 + (lane >> 2) + 8 * (e >> 1)) * H + ua + (e & 1));
        }
#if TC4_CREG
#pragma unroll
        for (int e = 0; e < 4; ++e) creg[sub][e] = z[e];
#else
        tmem_st4_16x256(tmem_c + (uint32_t)(sub * 32) + lane_off + u0, z);
#endif
      }
      if (!TC4_CREG) tmem_wait_st();
#pragma unroll 1
      for (int t = 0; t < L; ++t) {
        const int s = dir ? (L - 1 - t) : t;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          long long* tp = (tr && t >= 8 && t < 16) ? p.trace + ((t - 8) * 2 + sub) * 16 : nullptr;
          const uint32_t addc[2] = {addn[sub][0], addn[sub][1]};
          if (thr_add && t + 1 < L) {     // residual operand of the next layer, fetched one step ahead
#pragma unroll
            for (int i = 0; i < 2; ++i)
              if (valid[sub][i]) {
                const long long posn = base[sub][i] + (long long)(dir ? (L - 2 - t) : (t + 1)) * sstride;
                addn[sub][i] = __ldg(reinterpret_cast<const unsigned int*>(p.addend + posn * p.addend_ld + dir * H + ua));
              }
          }
          if (!TC4_PUB && tma_any && pusher) bulk_wait_read_all();
          if (TRACE && tp) tp[4] = clock64();
          mbar_wait(ACC_FULL(a), (full_par >> a) & 1u, p.error_flag, 300 + a);
          full_par ^= 1u << a;
          if (TRACE && tp) tp[5] = clock64();
          tc_fence_after();
          const uint32_t acc = tmem_acc + (uint32_t)a * kChunkN + lane_off + u0;
          [[maybe_unused]] const uint32_t cad = tmem_c + (uint32_t)(sub * 32) + lane_off + u0;
          float gi[4], gf[4], gg[4], go[4];
#if TC4_CREG
          float (&cs)[4] = creg[sub];
#else
          float cs[4];
#endif
          tmem_ld4_16x256(acc + 0 * kChunkUnits, gi);
          tmem_ld4_16x256(acc + 1 * kChunkUnits, gf);
          tmem_ld4_16x256(acc + 2 * kChunkUnits, gg);
          tmem_ld4_16x256(acc + 3 * kChunkUnits, go);
          if (!TC4_CREG) tmem_ld4_16x256(cad, cs);
          const bool will_publish = t + 1 < L || tma_any;
          bool hfree_ok = !(will_publish && t + hoff > 0);
          if (TC4_HFREE_EARLY && !hfree_ok) hfree_ok = mbar_test_wait_cluster(H_FREE(sub), (uint32_t)((t - 1 + hoff) & 1));
          tmem_wait_ld();
          tmem_ld_dep4(gi); tmem_ld_dep4(gf); tmem_ld_dep4(gg); tmem_ld_dep4(go);
          if (!TC4_CREG) tmem_ld_dep4(cs);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(ACC_EMPTY(a));
          a = (a == kAccBufs - 1) ? 0 : a + 1;
          float hv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int w = e & 1;
            if (p.debug & 1) hv[e] = gi[e] + gf[e] + gg[e] + go[e] + cs[e];
            else if (fast) hv[e] = lstm_cell_tanh(gi[e] + bias_i[w], gf[e] + bias_f[w], gg[e] + bias_g[w], go[e] + bias_o[w], cs[e]);
            else hv[e] = lstm_cell(gi[e] + bias_i[w], gf[e] + bias_f[w], gg[e] + bias_g[w], go[e] + bias_o[w], cs[e]);
          }
          __half2 hp[2] = {__floats2half2_rn(hv[0], hv[1]), __floats2half2_rn(hv[2], hv[3])};
          if (TRACE && tp) tp[6] = clock64();
          if (will_publish) {
            if (!hfree_ok) mbar_wait_cluster(H_FREE(sub), (uint32_t)((t - 1 + hoff) & 1), p.error_flag, 320 + sub);
            const uint32_t buf = hs_base + (uint32_t)(sub * C) * kHTile;
            st_shared_b32(buf + hpiece[0], *reinterpret_cast<uint32_t*>(&hp[0]));
            st_shared_b32(buf + hpiece[1], *reinterpret_cast<uint32_t*>(&hp[1]));
            publish_quadrant(buf, sub, s, t + 1 < L);
          }
          if (TRACE && tp) tp[7] = clock64();
          if (!TC4_CREG) tmem_st4_16x256(cad, cs);
          if ((p.state_flags & 2) && t + 1 == L) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (valid[sub][e >> 1]) {
                const long long so = (row0 + sub * SUB + q * 16 + (lane >> 2) + 8 * (e >> 1)) * H + ua + (e & 1);
                p.c_state[so] = cs[e];
                p.h_state[so] = __half2float(__float2half_rn(hv[e]));
              }
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (!valid[sub][i]) continue;
            const long long pos = base[sub][i] + (long long)s * sstride;
            if (thr_out0) *reinterpret_cast<__half2*>(p.out0 + pos * p.out0_ld + p.out0_off + dir * H + ua) = hp[i];
            if (thr_out1) {
              const __half2 av = *reinterpret_cast<const __half2*>(&addc[i]);
              *reinterpret_cast<__half2*>(p.out1 + pos * p.out1_ld + dir * H + ua) =
                  __floats2half2_rn(hv[2 * i] + __low2float(av), hv[2 * i + 1] + __high2float(av));
            }
          }
          if (!TC4_CREG) tmem_wait_st();
        }
      }
    }
  }

  bulk_wait_all();         // TMA tile stores of this thread (if any) are complete before the CTA may exit
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // no CTA leaves while a peer may still write into its shared memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------

struct Plan { bool ok; int sub; int xstages; int nxs; size_t smem; bool small1; };

static Plan make_plan(int H, int c0, int c1, int axis = -1, bool small_grid = false) {
  Plan pl{false, 0, 0, 0, 0, false};
  if (H != 64 && H != 128 && H != 256) return pl;
  if (c0 % 16 || c1 % 16 || c0 <= 0) return pl;
  const int nxs = (c0 + 63) / 64 + (c1 + 63) / 64;
  if (nxs > kMaxXSlabs) return pl;
  const int C = H / kChunkUnits, NHS = H / 64;
  // A second source of <= 16 channels (the raw-feature skip of the first FN-SSL block and of every IPDnet layer) is one
  // K = 16 step.  As a full [SUB x 64] slab it would cost a 16 KB weight slab and a whole ring stage per slot -- which leaves
  // the 272-channel layers 2-3 stages for 5 slabs per slot.  It gets a 4 KB weight slab and a two-entry ring of [SUB x 16]
  // slabs instead (32B swizzle), so the big ring carries 4 slabs per slot again.
  // Measured (B200, profiles/r1_lstm_narrow_source_ring.txt): H = 256 layers 4.65 -> 3.2 ms (3 big stages instead of 2), and
  // 3.0 ms when every CTA also fetches its own x slabs (no multicast: the cluster-wide "slot empty" hand-shake is what a
  // short ring cannot hide); H = 128: along-frequency layers 2.38 -> 2.08 ms (IPDnet cfg3), along-time layers not faster
  // (1.23 -> 1.27 ms), so the mode (template flag NARROW) is used for H = 256 and for H = 128 full-band layers.
  // FNSSL_TC_SMALL1 = 0 / 1 forces it off / on for every H (tests).
  bool small1 = c1 > 0 && c1 <= 16 && (H == 256 || (H == 128 && axis == FNSSL_ALONG_FREQ));
  if (const char* e = getenv("FNSSL_TC_SMALL1")) small1 = c1 > 0 && c1 <= 16 && atoi(e) != 0;
  const int nslabs = (small1 ? nxs - 1 : nxs) + NHS;
  // Small grids (few utterances): with 64-row sub-tiles the layer has twice the clusters, each slot half the gate math and half
  // the exchange -- the recurrence is latency bound there, so the half-rate M = 64 MMAs do not matter.  Measured (B200, one
  // utterance, profiles/r2_lstm_variants.txt): full in16 0.71 -> 0.59 ms, narrow in272 0.92 -> 0.66 ms; 2 utterances in256
  // 0.79 -> 0.66 ms.  Used whenever the 64-row plan still fits the GPU in one wave (lstm_forward_tc4).
  int sub_first = small_grid ? 64 : 128;
  if (const char* e = getenv("FNSSL_TC_ROWS")) { sub_first = atoi(e) == 64 ? 64 : 128; }   // tests / profiling
  for (int sub = sub_first; sub >= 64; sub -= 64) {
    const long xslab = sub * 128L, htile = sub * 64L;
    const long fixed = (long)nslabs * kWSlab + (small1 ? kWSmall + 2L * sub * 32 : 0L) + 2L * C * htile + kChunkN * 4 + 1024;
    long xs = (kSmemLimit - 1024 - fixed) / xslab;
    if (xs > kMaxXStages) xs = kMaxXStages;
    if (xs >= 2) {
      pl.ok = true; pl.sub = sub; pl.xstages = (int)xs; pl.nxs = nxs; pl.small1 = small1;
      pl.smem = (size_t)fixed + (size_t)xs * xslab;
      return pl;
    }
  }
  return pl;
}

// FNSSL_TC_TRACE buffers: device memory (stores into host-mapped memory stall the traced threads for ~2 k cycles per slot),
// one per device, created under a lock
constexpr int kMaxDevices = 64;
static long long* g_trace_dev[kMaxDevices] = {};
static int g_trace_last = -1;
static std::mutex g_trace_mu;
static long long* trace_buffer() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  std::lock_guard<std::mutex> lk(g_trace_mu);
  if (!g_trace_dev[dev]) {
    if (cudaMalloc(&g_trace_dev[dev], 256 * sizeof(long long)) != cudaSuccess) { g_trace_dev[dev] = nullptr; return nullptr; }
    cudaMemset(g_trace_dev[dev], 0, 256 * sizeof(long long));
  }
  g_trace_last = dev;
  return g_trace_dev[dev];
}

template <int H, int SUB, bool TRACE, bool NARROW>
static int launch_t(const fnssl_lstm_args* a, const Plan& pl, cudaStream_t st) {
  constexpr int C = H / kChunkUnits, NHS = H / kSlabK;
  constexpr int kTileRows = 2 * SUB;
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
  const int nslabs = nxs + NHS;
  const int64_t wbytes = (int64_t)a->num_dirs * C * kChunkN * nslabs * kSlabK * 2;
  const int64_t need = wbytes + (int64_t)a->num_dirs * 4 * H * 4;
  FNSSL_REQUIRE(a->weights_bytes == need, "lstm(tcgen05): packed weight buffer is %lld bytes, expected %lld",
                (long long)a->weights_bytes, (long long)need);
  FNSSL_REQUIRE((reinterpret_cast<uintptr_t>(a->weights) & 15) == 0, "lstm(tcgen05): weights not 16-byte aligned");
  p.h_state = a->h_state; p.c_state = a->c_state; p.state_flags = a->state_flags;
  FNSSL_REQUIRE(!a->state_flags || (reinterpret_cast<uintptr_t>(a->h_state) & 15) == 0, "lstm(tcgen05): h_state not 16-byte aligned");
  p.axis = a->axis; p.nf = a->nf; p.nt = a->nt;
  int tiles;
  if (a->axis == FNSSL_ALONG_FREQ) {
    p.rows = (long long)a->nb * a->nt; p.steps = a->nf; p.tiles_per_b = 0;
    tiles = (int)((p.rows + kTileRows - 1) / kTileRows);
  } else {
    p.rows = (long long)a->nb * a->nf; p.steps = a->nt; p.tiles_per_b = (a->nf + kTileRows - 1) / kTileRows;
    tiles = a->nb * p.tiles_per_b;
  }
  p.bias = reinterpret_cast<const float*>(reinterpret_cast<const char*>(a->weights) + wbytes);
  p.out0 = (__half*)a->out0; p.out0_ld = a->out0_ld; p.out0_off = a->out0_off;
  p.addend = (const __half*)a->addend; p.addend_ld = a->addend_ld;
  p.out1 = (__half*)a->out1; p.out1_ld = a->out1_ld;
  if (SUB == 128) {   // 16-byte epilogue accesses
    FNSSL_REQUIRE(!a->out0 || ((reinterpret_cast<uintptr_t>(a->out0) & 15) == 0 && a->out0_ld % 8 == 0 && a->out0_off % 8 == 0),
                  "lstm(tcgen05): out0 must be 16-byte aligned (ld, offset multiples of 8)");
    FNSSL_REQUIRE(!a->out1 || ((reinterpret_cast<uintptr_t>(a->out1) & 15) == 0 && a->out1_ld % 8 == 0 &&
                               (reinterpret_cast<uintptr_t>(a->addend) & 15) == 0 && a->addend_ld % 8 == 0),
                  "lstm(tcgen05): out1/addend must be 16-byte aligned");
  } else {            // 4-byte (half2) epilogue accesses
    FNSSL_REQUIRE(!a->out0 || ((reinterpret_cast<uintptr_t>(a->out0) & 3) == 0 && a->out0_ld % 2 == 0 && a->out0_off % 2 == 0),
                  "lstm(tcgen05): out0 must be 4-byte aligned (even ld / offset)");
    FNSSL_REQUIRE(!a->out1 || ((reinterpret_cast<uintptr_t>(a->out1) & 3) == 0 && a->out1_ld % 2 == 0 &&
                               (reinterpret_cast<uintptr_t>(a->addend) & 3) == 0 && a->addend_ld % 2 == 0),
                  "lstm(tcgen05): out1/addend must be 4-byte aligned");
  }
  p.error_flag = tc_wait_timeout_enabled() ? tc_error_flag() : nullptr;
  p.debug = tc_debug_bits(32 | 64);
  if (getenv("FNSSL_TC_TRACE")) p.trace = trace_buffer();

  CUtensorMap m0, m1, mw, mw2;
  if (make_grid_map(&m0, a->src0, a->c0, a->ld0, a->nb, a->nt, a->nf, a->axis, SUB)) return 1;
  if (pl.small1) { if (make_small_grid_map(&m1, a->src1, a->c1, a->ld1, a->nb, a->nt, a->nf, a->axis, SUB)) return 1; }
  else if (a->c1 > 0) { if (make_grid_map(&m1, a->src1, a->c1, a->ld1, a->nb, a->nt, a->nf, a->axis, SUB)) return 1; }
  else m1 = m0;
  if (make_weight_map(&mw, a->weights, nslabs, a->num_dirs * C)) return 1;
  mw2 = mw;
  if (pl.small1 && make_small_weight_map(&mw2, a->weights, nslabs, a->num_dirs * C)) return 1;
  // outputs through TMA: out0 as tile stores, out1 as an in-place reduce-add when it aliases the residual operand
  const bool no_tma_out = getenv("FNSSL_TC_NO_TMA_OUT") != nullptr;
  const bool two_copies = a->out0 && a->out1 && !a->addend && !no_tma_out && a->out0_off % 8 == 0;      // -> whole-tile output boxes
  p.store_tile = (TC4_PUB == 1 && (two_copies || (SUB == 64 && H != 256))) ? 1 : 0;      // (SUB == 64 with H < 256: the small-grid plan)
  const int kOutBoxRows = p.store_tile ? SUB : SUB / 4;
  CUtensorMap mo0 = m0, mo1 = m0;
  if (a->out0 && !no_tma_out && a->out0_off % 8 == 0) {
    if (make_out_map(&mo0, a->out0, a->out0_ld, a->nb, a->nt, a->nf, a->axis, kOutBoxRows)) return 1;
    p.tma_out |= 1;
  }
  if (a->out1 && a->out1 == a->addend && a->out1_ld == a->addend_ld && !no_tma_out) {
    if (make_out_map(&mo1, a->out1, a->out1_ld, a->nb, a->nt, a->nf, a->axis, kOutBoxRows)) return 1;
    p.tma_out |= 2;
  } else if (a->out1 && !a->addend && !no_tma_out) {      // out1 = second copy of h: plain tile stores
    if (make_out_map(&mo1, a->out1, a->out1_ld, a->nb, a->nt, a->nf, a->axis, kOutBoxRows)) return 1;
    p.tma_out |= 4;
  }

  auto kern = lstm_tc4_kernel<H, SUB, TRACE, NARROW>;
  FNSSL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)tiles * C, (unsigned)a->num_dirs, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  FNSSL_CUDA(cudaLaunchKernelEx(&cfg, kern, m0, m1, mw, mw2, mo0, mo1, p));
  FNSSL_LAUNCH_CHECK("lstm_tc4_kernel");
  return 0;
}

template <int H, int SUB>
static int launch(const fnssl_lstm_args* a, const Plan& pl, cudaStream_t st) {
  if (pl.small1) return getenv("FNSSL_TC_TRACE") ? launch_t<H, SUB, true, true>(a, pl, st) : launch_t<H, SUB, false, true>(a, pl, st);
  return getenv("FNSSL_TC_TRACE") ? launch_t<H, SUB, true, false>(a, pl, st) : launch_t<H, SUB, false, false>(a, pl, st);
}

}  // namespace tc4

// diagnostic: copy the last trace (16 slots x 16 clock64 stamps) recorded with FNSSL_TC_TRACE=1
long long* lstm_tc5_trace_buffer();      // lstm_tc5.cu (we are inside namespace fnssl)
extern "C" int fnssl_lstm_tc4_trace(long long* out256) {
  if (long long* t5 = lstm_tc5_trace_buffer())      // the pair kernel ran with FNSSL_TC_TRACE: its timeline takes precedence
    return cudaMemcpy(out256, t5, 256 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 1 : 0;
  std::lock_guard<std::mutex> lk(tc4::g_trace_mu);
  if (tc4::g_trace_last < 0 || !tc4::g_trace_dev[tc4::g_trace_last]) return 0;
  return cudaMemcpy(out256, tc4::g_trace_dev[tc4::g_trace_last], 256 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 1 : 0;
}

bool lstm_tc4_supports(int hidden, int c0, int c1) { return tc4::make_plan(hidden, c0, c1).ok; }

int lstm_forward_tc4(const fnssl_lstm_args* a, cudaStream_t st) {
  // cluster tiles of the 64-row plan (2 x 64 rows each); one wave = 132 co-resident CTAs for clusters of 4 (33 clusters)
  const long long tiles64 = a->axis == FNSSL_ALONG_FREQ ? ((long long)a->nb * a->nt + 127) / 128 : (long long)a->nb * ((a->nf + 127) / 128);
  const bool small_grid = tiles64 * a->num_dirs * (a->hidden / 32) <= 132;
  const tc4::Plan pl = tc4::make_plan(a->hidden, a->c0, a->c1, a->axis, small_grid);
  FNSSL_REQUIRE(pl.ok, "lstm(tcgen05 two-chain kernel): unsupported layer (H=%d c0=%d c1=%d)", a->hidden, a->c0, a->c1);
  if (a->hidden == 64) return pl.sub == 128 ? tc4::launch<64, 128>(a, pl, st) : tc4::launch<64, 64>(a, pl, st);
  if (a->hidden == 128) return pl.sub == 128 ? tc4::launch<128, 128>(a, pl, st) : tc4::launch<128, 64>(a, pl, st);
  FNSSL_REQUIRE(pl.sub == 64, "lstm(tcgen05 two-chain kernel): H = 256 needs 64-row sub-tiles");
  return tc4::launch<256, 64>(a, pl, st);
}

}  // namespace fnssl
