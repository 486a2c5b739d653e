// Interleaved cluster-resident tcgen05 LSTM kernel (third generation of FNSSL_ENGINE_TCGEN05), H in {64, 128}.
//
// Same decomposition as lstm_tc2.cu -- the 4H gate columns are split over a cluster of C = H/32 CTAs, each keeping its
// weight slice resident in shared memory, h_t exchanged by DSMEM bulk copies -- but the 128 sequences of a cluster tile
// are treated as TWO independent sub-tiles (A = rows 0..63, B = rows 64..127) whose recurrences run half a step apart.
// The per-step serial chain of one sub-tile (h-part MMA -> gate math -> DSMEM exchange -> barrier hand-offs) leaves the
// MUFU pipe, the tensor pipe and the DSMEM fabric idle most of the time (measured timeline: tools/tc_trace.py); with
// two chains interleaved, B's gate math overlaps A's exchange and MMA and vice versa.
//
//   * M = 64 MMAs: an accumulator occupies lanes 0-15 of every TMEM quadrant, so sub-tile B's accumulators and cell
//     state live in lanes 16-31 of the SAME columns (D address lane offset 16).
//   * all 16 epilogue warps serve A, then B, then A ... with 16x256b TMEM accesses (thread T owns rows {T/4, T/4+8} of
//     the quadrant's 16 rows and two adjacent hidden units), so every lane is busy on either sub-tile.
//   * one x ring; one MMA-issuing thread per sub-tile (warps 1 and 18), each running  h-part(t), x-part(t+1)  for its
//     own chain, so the issue back-pressure and the multicast commits of one chain never delay the other.
//
// Replaces nn.LSTM at FN-SSL/Lightning/Model.py:38,46 and IPDnet/FixedAarryIPDnet.py:32,36 (+ glue :35-37,41-45,49).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace fnssl {
namespace tc3 {

constexpr int kThreads = 608;          // producer warp + MMA warp (sub-tile A) + 16 epilogue warps + MMA warp (sub-tile B)
constexpr int kEpiThreads = 512;
constexpr int kSlabK = 64;
constexpr int kWSlab = 128 * 128;      // [128 gate columns x 64] fp16
constexpr int kChunkUnits = 32;
constexpr int kChunkN = 128;
constexpr int kSubRows = 64;           // rows of one sub-tile (UMMA M)
constexpr int kTileRows = 128;         // rows of a cluster tile (two sub-tiles)
constexpr int kXSlab = kSubRows * 128; // one [64 x 64] fp16 x slab
constexpr int kHTile = kSubRows * 64;  // one [64 x 32] fp16 h tile (one chunk, 64B swizzle)
constexpr int kMaxXSlabs = 6;
constexpr int kMaxXStages = 8;
constexpr int kSmemLimit = 232448;

struct Params {
  int nxs;
  int xs_src[kMaxXSlabs];
  int xs_k0[kMaxXSlabs];
  int xs_nk16[kMaxXSlabs];
  int xstages;
  int steps, axis, nf, nt;
  long long rows;
  int tiles_per_b;
  const float* bias;               // [dirs][4H], accumulator column order [chunk][gate][unit]
  __half* out0; int out0_ld; int out0_off;
  const __half* addend; int addend_ld;
  __half* out1; int out1_ld;
  int* error_flag;
  int debug;                       // timing experiments only (FNSSL_TC_DEBUG): 1 = skip gate math, 2 = skip MMA issue
};

constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kChunkN >> 3) << 17) | ((uint32_t)(kSubRows >> 4) << 24);

template <int H>
__global__ void __launch_bounds__(kThreads, 1)
lstm_tc3_kernel(const __grid_constant__ CUtensorMap map_src0, const __grid_constant__ CUtensorMap map_src1,
                const __grid_constant__ CUtensorMap map_w, const Params p) {
  constexpr int C = H / kChunkUnits;      // cluster size == number of 32-unit chunks
  constexpr int NHS = H / kSlabK;         // K slabs of h in the weight layout
  static_assert(H == 64 || H == 128, "H in {64,128}");

  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) unsigned long long bars[1 + 2 * kMaxXStages + 12];
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;
  const uint32_t rank = cluster_ctarank();          // == chunk owned by this CTA
  const int tile = blockIdx.x / C;
  const uint16_t mask = (uint16_t)((1u << C) - 1u);
  const int nxs = p.nxs, XS = p.xstages, L = p.steps;
  const int nslabs = nxs + NHS;

  const uint32_t dyn0 = (smem_addr(smem_dyn) + 1023u) & ~1023u;
  const uint32_t w_base = dyn0;                                     // resident weights: nslabs tiles
  const uint32_t hs_base = w_base + (uint32_t)nslabs * kWSlab;      // h operand: [sub][buffer][chunk] tiles of 4 KB
  const uint32_t xr_base = hs_base + 4u * C * kHTile;               // x ring: XS slabs of 8 KB
  const uint32_t bias_base = xr_base + (uint32_t)XS * kXSlab;       // 128 floats
  float* bias_s = reinterpret_cast<float*>(smem_dyn + (bias_base - smem_addr(smem_dyn)));

  const uint32_t bar0 = smem_addr(bars);
  const uint32_t W_FULL = bar0;
  auto X_FULL = [&](int i) { return bar0 + 8u * (1 + i); };
  auto X_EMPTY = [&](int i) { return bar0 + 8u * (1 + kMaxXStages + i); };
  auto ACC_FULL = [&](int sub, int i) { return bar0 + 8u * (1 + 2 * kMaxXStages + sub * 2 + i); };
  auto ACC_EMPTY = [&](int sub, int i) { return bar0 + 8u * (1 + 2 * kMaxXStages + 4 + sub * 2 + i); };
  auto H_FULL = [&](int sub, int i) { return bar0 + 8u * (1 + 2 * kMaxXStages + 8 + sub * 2 + i); };

  if (tid == 0) {
    mbar_init(W_FULL, 1);
    for (int i = 0; i < kMaxXStages; ++i) { mbar_init(X_FULL(i), 1); mbar_init(X_EMPTY(i), C); }
    for (int sub = 0; sub < 2; ++sub)
      for (int i = 0; i < 2; ++i) {
        mbar_init(ACC_FULL(sub, i), 1);
        mbar_init(ACC_EMPTY(sub, i), kEpiThreads);
        mbar_init(H_FULL(sub, i), 5);   // MMA thread's expect_tx + 4 local quadrants
      }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&map_src0); prefetch_tmap(&map_src1); prefetch_tmap(&map_w); }
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
  const uint32_t tmem_c = tmem;             // cell state: columns [0, 32); sub-tile A lanes 0-15, B lanes 16-31 of a quadrant
  const uint32_t tmem_acc = tmem + 128;     // gate accumulators: 2 buffers x 128 columns, same lane split

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

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      mbar_expect_tx(W_FULL, (uint32_t)nslabs * kWSlab);
      for (int j = 0; j < nslabs; ++j)
        tma_load_2d(w_base + j * kWSlab, &map_w, W_FULL, j * kSlabK, (dir * C + (int)rank) * kChunkN);
      int n = 0;
      // L2 prefetch of step t+2's slabs (see lstm_tc2.cu): the ring's real loads then hit L2
      auto prefetch_step = [&](int tt) {
        const int ss = dir ? (L - 1 - tt) : tt;
        for (int sub = 0; sub < 2; ++sub)
          for (int j = 0; j < nxs; ++j) {
            if ((uint32_t)(((tt * 2 + sub) * nxs + j) % C) != rank) continue;
            const CUtensorMap* m = p.xs_src[j] ? &map_src1 : &map_src0;
            const int r0 = coord_r0 + sub * kSubRows;
            if (p.axis == FNSSL_ALONG_FREQ) tma_prefetch_l2_4d(m, p.xs_k0[j], ss, r0, 0);
            else tma_prefetch_l2_4d(m, p.xs_k0[j], r0, ss, coord_b);
          }
      };
      if (!(p.debug & 16)) { if (L > 1) prefetch_step(1); }
      for (int t = 0; t < L; ++t) {
        const int s = dir ? (L - 1 - t) : t;
        if (!(p.debug & 16) && t + 2 < L) prefetch_step(t + 2);
        for (int sub = 0; sub < 2; ++sub) {
          for (int j = 0; j < nxs; ++j, ++n) {
            const int stage = n % XS, use = n / XS;
            if (use > 0) mbar_wait(X_EMPTY(stage), (uint32_t)((use - 1) & 1), p.error_flag, 100 + stage);
            mbar_expect_tx(X_FULL(stage), kXSlab);
            if ((uint32_t)(n % C) == rank) {   // one CTA fetches the slab for the whole cluster
              const CUtensorMap* m = p.xs_src[j] ? &map_src1 : &map_src0;
              const uint32_t dst = xr_base + (uint32_t)stage * kXSlab;
              const int r0 = coord_r0 + sub * kSubRows;
              if (p.axis == FNSSL_ALONG_FREQ) tma_load_4d_mc(dst, m, X_FULL(stage), p.xs_k0[j], s, r0, 0, mask);
              else tma_load_4d_mc(dst, m, X_FULL(stage), p.xs_k0[j], r0, s, coord_b, mask);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 18) {
    // ============================== MMA issuers (warp 1: sub-tile A, warp 18: sub-tile B) ==============================
    if (lane == 0) {
      const int sub = (warp == 1) ? 0 : 1;
      mbar_wait(W_FULL, 0, p.error_flag, 200);
      auto x_part = [&](int s) {   // G_x of step s of this sub-tile -> accumulator buffer s & 1
        const int b = s & 1;
        if (s >= 2) mbar_wait(ACC_EMPTY(sub, b), (uint32_t)(((s >> 1) - 1) & 1), p.error_flag, 201 + sub * 2 + b);
        tc_fence_after();
        const uint32_t d_tmem = tmem_acc + (uint32_t)b * kChunkN + ((uint32_t)(sub * 16) << 16);
        uint32_t accumulate = 0;
        for (int j = 0; j < nxs; ++j) {
          const int n = (2 * s + sub) * nxs + j;           // position of the slab in the ring's sequence
          const int stage = n % XS;
          mbar_wait(X_FULL(stage), (uint32_t)((n / XS) & 1), p.error_flag, 210 + stage);
          tc_fence_after();
          const uint64_t a_desc = make_sw128_desc(xr_base + (uint32_t)stage * kXSlab);
          const uint64_t b_desc = make_sw128_desc(w_base + (uint32_t)j * kWSlab);
          const int nk16 = p.xs_nk16[j];
          if (!(p.debug & 2)) {
            for (int k = 0; k < nk16; ++k) {
              umma_f16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, kIdesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit_mc(X_EMPTY(stage), mask);    // this CTA is done with the slab: tell every CTA's ring
        }
      };
      auto h_part = [&](int t) {   // += h_{t-1} W_h^T, then hand the accumulator to the epilogue
        if (t > 0) {
          const int pb = (t - 1) & 1;
          mbar_expect_tx(H_FULL(sub, pb), (uint32_t)((C - 1) * kHTile));   // C-1 remote tiles (tx) + 4 local arrives
          mbar_wait_cluster(H_FULL(sub, pb), (uint32_t)(((t - 1) >> 1) & 1), p.error_flag, 220 + sub);
          tc_fence_after();
          const uint32_t d_tmem = tmem_acc + (uint32_t)(t & 1) * kChunkN + ((uint32_t)(sub * 16) << 16);
#pragma unroll
          for (int kc = 0; kc < C; ++kc) {   // K = 32 units of chunk kc: two K=16 steps; W columns inside 128B-swizzled slab kc/2
            const uint64_t a_desc = make_sw64_desc(hs_base + (uint32_t)((sub * 2 + pb) * C + kc) * kHTile);
            const uint64_t b_desc = make_sw128_desc(w_base + (uint32_t)(nxs + (kc >> 1)) * kWSlab) + 4u * (kc & 1);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              if (!(p.debug & 2)) umma_f16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, kIdesc, 1u);
          }
        }
        umma_commit(ACC_FULL(sub, t & 1));
      };
      x_part(0);
      for (int t = 0; t < L; ++t) {
        h_part(t);
        if (t + 1 < L) x_part(t + 1);
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue warps ==============================
    const int q = warp & 3;                    // TMEM lane quadrant of this warp
    const int sg = (warp - 2) >> 2;            // 8-unit group of the CTA's 32 units
    const int u0 = sg * 8;
    const int uo = 2 * (lane & 3);             // unit offset inside the 8-unit group (16x256b: two adjacent columns)
    const int ua = (int)rank * kChunkUnits + u0 + uo;   // absolute hidden unit of the thread's first column
    const float kL2E = 1.4426950408889634f;
    const float bias_i[2] = {bias_s[u0 + uo], bias_s[u0 + uo + 1]};
    const float bias_f[2] = {bias_s[kChunkUnits + u0 + uo], bias_s[kChunkUnits + u0 + uo + 1]};
    const float bias_g[2] = {bias_s[2 * kChunkUnits + u0 + uo], bias_s[2 * kChunkUnits + u0 + uo + 1]};
    const float bias_o[2] = {bias_s[3 * kChunkUnits + u0 + uo], bias_s[3 * kChunkUnits + u0 + uo + 1]};
    // One (row, unit): c' = sigmoid(f) c + sigmoid(i) tanh(g), h = sigmoid(o) tanh(c'); 5 ex2 + 2 rcp (see lstm_tc2.cu)
    auto lstm_cell = [&](float gi, float gf, float gg, float go, float& c) -> float {
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
    };
    const long long sstride = (p.axis == FNSSL_ALONG_FREQ) ? 1 : p.nf;
    // per sub-tile, per owned row (2 rows): grid position base, validity, h-piece offset
    long long base[2][2];
    bool valid[2][2];
    uint32_t hpiece[2][2];
    uint32_t addn[2][2];
#pragma unroll
    for (int sub = 0; sub < 2; ++sub)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int rs = q * 16 + (lane >> 2) + 8 * i;       // row inside the sub-tile
        const int r = sub * kSubRows + rs;                 // row inside the cluster tile
        valid[sub][i] = r < valid_rows;
        base[sub][i] = (p.axis == FNSSL_ALONG_FREQ) ? (row0 + r) * p.nf : ((long long)coord_b * p.nt * p.nf + coord_r0 + r);
        hpiece[sub][i] = (uint32_t)rank * kHTile + (uint32_t)(rs >> 3) * 512u + (uint32_t)(rs & 7) * 64u +
                         (uint32_t)((sg ^ ((rs >> 1) & 3)) << 4) + (uint32_t)uo * 2u;
        addn[sub][i] = 0u;
        if (p.out1 && valid[sub][i]) {
          const long long pos0 = base[sub][i] + (long long)(dir ? (L - 1) : 0) * sstride;
          addn[sub][i] = __ldg(reinterpret_cast<const unsigned int*>(p.addend + pos0 * p.addend_ld + dir * H + ua));
        }
      }
    const uint32_t hquad = (uint32_t)rank * kHTile + (uint32_t)q * 1024u;   // the quadrant's 16 rows x 64 B
    {
      float z[4] = {0.f, 0.f, 0.f, 0.f};
      tmem_st4_16x256(tmem_c + ((uint32_t)(q * 32) << 16) + u0, z);
      tmem_st4_16x256(tmem_c + ((uint32_t)(q * 32 + 16) << 16) + u0, z);
      tmem_wait_st();
    }

#pragma unroll 1
    for (int t = 0; t < L; ++t) {
      const int b = t & 1;
      const int s = dir ? (L - 1 - t) : t;
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const uint32_t lane_off = (uint32_t)(q * 32 + sub * 16) << 16;
        const uint32_t addc[2] = {addn[sub][0], addn[sub][1]};
        if (p.out1 && t + 1 < L) {     // residual operand of the next layer, fetched one step ahead
#pragma unroll
          for (int i = 0; i < 2; ++i)
            if (valid[sub][i]) {
              const long long posn = base[sub][i] + (long long)(dir ? (L - 2 - t) : (t + 1)) * sstride;
              addn[sub][i] = __ldg(reinterpret_cast<const unsigned int*>(p.addend + posn * p.addend_ld + dir * H + ua));
            }
        }
        mbar_wait(ACC_FULL(sub, b), (uint32_t)((t >> 1) & 1), p.error_flag, 300 + sub * 2 + b);
        tc_fence_after();
        const uint32_t acc = tmem_acc + (uint32_t)b * kChunkN + lane_off + u0;
        float gi[4], gf[4], gg[4], go[4], cs[4];
        tmem_ld4_16x256(acc + 0 * kChunkUnits, gi);
        tmem_ld4_16x256(acc + 1 * kChunkUnits, gf);
        tmem_ld4_16x256(acc + 2 * kChunkUnits, gg);
        tmem_ld4_16x256(acc + 3 * kChunkUnits, go);
        tmem_ld4_16x256(tmem_c + lane_off + u0, cs);
        tmem_wait_ld();
        tmem_ld_dep4(gi); tmem_ld_dep4(gf); tmem_ld_dep4(gg); tmem_ld_dep4(go); tmem_ld_dep4(cs);
        tc_fence_before();
        mbar_arrive(ACC_EMPTY(sub, b));      // accumulator drained: the MMA warp may produce G_x of step t+2 into it
        float hv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int w = e & 1;
          if (p.debug & 1) hv[e] = gi[e] + gf[e] + gg[e] + go[e] + cs[e];
          else hv[e] = lstm_cell(gi[e] + bias_i[w], gf[e] + bias_f[w], gg[e] + bias_g[w], go[e] + bias_o[w], cs[e]);
        }
        __half2 hp[2] = {__floats2half2_rn(hv[0], hv[1]), __floats2half2_rn(hv[2], hv[3])};
        if (t + 1 < L) {
          // publish h_t of this sub-tile: own tile locally, then one DSMEM bulk copy per peer and quadrant
          const uint32_t buf = hs_base + (uint32_t)((sub * 2 + b) * C) * kHTile;
          st_shared_b32(buf + hpiece[sub][0], *reinterpret_cast<uint32_t*>(&hp[0]));
          st_shared_b32(buf + hpiece[sub][1], *reinterpret_cast<uint32_t*>(&hp[1]));
          fence_async_smem();
          named_bar_sync(1 + q, 128);
          if (sg == 0 && lane == 0) {
            const uint32_t hb = H_FULL(sub, b);
#pragma unroll
            for (int dd = 1; dd < C; ++dd) {
              const uint32_t d = (rank + (uint32_t)dd) % C;
              bulk_copy_s2c(mapa_shared(buf + hquad, d), buf + hquad, 1024u, mapa_shared(hb, d));
            }
            mbar_arrive(hb);     // the local copy of this quadrant is in place
          }
        }
        tmem_st4_16x256(tmem_c + lane_off + u0, cs);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (!valid[sub][i]) continue;
          const long long pos = base[sub][i] + (long long)s * sstride;
          if (p.out0) *reinterpret_cast<__half2*>(p.out0 + pos * p.out0_ld + p.out0_off + dir * H + ua) = hp[i];
          if (p.out1) {
            const __half2 av = *reinterpret_cast<const __half2*>(&addc[i]);
            *reinterpret_cast<__half2*>(p.out1 + pos * p.out1_ld + dir * H + ua) =
                __floats2half2_rn(hv[2 * i] + __low2float(av), hv[2 * i + 1] + __high2float(av));
          }
        }
        tmem_wait_st();
      }
    }
  }

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

struct Plan { bool ok; int xstages; int nxs; size_t smem; };

static Plan make_plan(int H, int c0, int c1) {
  Plan pl{false, 0, 0, 0};
  if (H != 64 && H != 128) return pl;
  if (c0 % 16 || c1 % 16 || c0 <= 0) return pl;
  const int nxs = (c0 + 63) / 64 + (c1 + 63) / 64;
  if (nxs > kMaxXSlabs) return pl;
  const int C = H / kChunkUnits, NHS = H / 64, nslabs = nxs + NHS;
  const long fixed = (long)nslabs * kWSlab + 4L * C * kHTile + kChunkN * 4 + 1024;
  long xs = (kSmemLimit - 1024 - fixed) / kXSlab;
  if (xs > kMaxXStages) xs = kMaxXStages;
  if (xs < 2) return pl;
  pl.ok = true; pl.xstages = (int)xs; pl.nxs = nxs;
  pl.smem = (size_t)fixed + (size_t)xs * kXSlab;
  return pl;
}

template <int H>
static int launch(const fnssl_lstm_args* a, const Plan& pl, cudaStream_t st) {
  constexpr int C = H / kChunkUnits, NHS = H / kSlabK;
  Params p{};
  int nxs = 0;
  for (int src = 0; src < 2; ++src) {
    const int c = src ? a->c1 : a->c0;
    for (int k0 = 0; k0 < c; k0 += kSlabK) {
      p.xs_src[nxs] = src; p.xs_k0[nxs] = k0;
      p.xs_nk16[nxs] = (((c - k0) < kSlabK ? (c - k0) : kSlabK) + 15) / 16;
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
  // 4-byte (half2) epilogue accesses
  FNSSL_REQUIRE(!a->out0 || ((reinterpret_cast<uintptr_t>(a->out0) & 3) == 0 && a->out0_ld % 2 == 0 && a->out0_off % 2 == 0),
                "lstm(tcgen05): out0 must be 4-byte aligned (even ld / offset)");
  FNSSL_REQUIRE(!a->out1 || ((reinterpret_cast<uintptr_t>(a->out1) & 3) == 0 && a->out1_ld % 2 == 0 &&
                             (reinterpret_cast<uintptr_t>(a->addend) & 3) == 0 && a->addend_ld % 2 == 0),
                "lstm(tcgen05): out1/addend must be 4-byte aligned");
  p.error_flag = tc_error_flag();
  if (const char* e = getenv("FNSSL_TC_DEBUG")) p.debug = atoi(e);

  CUtensorMap m0, m1, mw;
  if (make_grid_map(&m0, a->src0, a->c0, a->ld0, a->nb, a->nt, a->nf, a->axis, kSubRows)) return 1;
  if (a->c1 > 0) { if (make_grid_map(&m1, a->src1, a->c1, a->ld1, a->nb, a->nt, a->nf, a->axis, kSubRows)) return 1; }
  else m1 = m0;
  if (make_weight_map(&mw, a->weights, nslabs, a->num_dirs * C)) return 1;

  auto kern = lstm_tc3_kernel<H>;
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
  FNSSL_CUDA(cudaLaunchKernelEx(&cfg, kern, m0, m1, mw, p));
  FNSSL_LAUNCH_CHECK("lstm_tc3_kernel");
  return 0;
}

}  // namespace tc3

bool lstm_tc3_supports(int hidden, int c0, int c1) { return tc3::make_plan(hidden, c0, c1).ok; }

int lstm_forward_tc3(const fnssl_lstm_args* a, cudaStream_t st) {
  const tc3::Plan pl = tc3::make_plan(a->hidden, a->c0, a->c1);
  FNSSL_REQUIRE(pl.ok, "lstm(tcgen05 interleaved kernel): unsupported layer (H=%d c0=%d c1=%d)", a->hidden, a->c0, a->c1);
  return a->hidden == 64 ? tc3::launch<64>(a, pl, st) : tc3::launch<128>(a, pl, st);
}

}  // namespace fnssl
