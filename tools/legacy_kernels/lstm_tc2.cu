// Cluster-resident tcgen05 LSTM kernel (second generation of FNSSL_ENGINE_TCGEN05) for sm_100a.
//
// Why: the first kernel (lstm_tc.cu) lets one CTA compute all 4H gate columns of its row tile, so the layer's whole
// weight matrix (400-600 KB, more than one SM's shared memory) is re-streamed from L2 through a small TMA ring every
// time step; measured on B200 that stream -- not the MMA, not the gate math -- bounds the step (profiles/r1_*).
//
// Here the 4H gate columns are split across a thread-block cluster: CTA k of a cluster of C = H/32 CTAs owns hidden
// units [32k, 32k+32) (N = 128 gate columns i,f,g,o) of a shared tile of MR sequences, and keeps its 1/C slice of the
// weights RESIDENT in shared memory for the whole launch (loaded once by TMA).  Per step and CTA:
//     x-part   G  = x_t . W_x^T      x_t slabs are TMA-multicast to the whole cluster through a small ring; issued a
//                                    step ahead into the other accumulator buffer, off the critical path
//     h-part   G += h_{t-1} . W_h^T  needs all of h_{t-1}: h lives as C tiles [MR x 32 units] (64B-swizzled K-major); a
//                                    CTA's epilogue writes its own tile locally and each 32-row quadrant is pushed to
//                                    every peer with one DSMEM bulk copy (cp.async.bulk shared::cta -> shared::cluster)
//                                    that completes tx bytes on the peer's "h_t complete" mbarrier
//     epilogue 16 warps: tcgen05.ld gates -> sigmoid/tanh -> c (fp32, resident in TMEM) -> h_t (fp16) -> DSMEM + HBM
// The serial chain of a step is: h-part MMA (K = H) -> gate math of 32 units -> DSMEM exchange; weights never move.
//
// Replaces nn.LSTM at FN-SSL/Lightning/Model.py:38,46 and IPDnet/FixedAarryIPDnet.py:32,36 (+ glue :35-37,41-45,49).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace fnssl {
namespace tc2 {

constexpr int kThreads = 576;          // producer warp + MMA warp + 16 epilogue warps
constexpr int kEpiThreads = 512;
constexpr int kEpiWarps = 16;
constexpr int kSlabK = 64;
constexpr int kWSlab = 128 * 128;      // [128 gate columns x 64] fp16
constexpr int kChunkUnits = 32;
constexpr int kChunkN = 128;
constexpr int kMaxXSlabs = 6;
constexpr int kMaxXStages = 6;
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
  float* h_state; float* c_state; int state_flags;   // optional carried state, fp32 (rows, H); dirs == 1
  int* error_flag;
  long long* trace;                // FNSSL_TC_TRACE: clock64 stamps of cluster 0 / CTA 0, steps [8, 16): [step][16 events]
  int debug;                       // timing experiments only (FNSSL_TC_DEBUG): 1 = skip gate math, 2 = skip MMA issue
};

template <int MR>
constexpr uint32_t make_idesc() { return (1u << 4) | ((uint32_t)(kChunkN >> 3) << 17) | ((uint32_t)(MR >> 4) << 24); }

template <int H, int MR>
__global__ void __launch_bounds__(kThreads, 1)
lstm_tc2_kernel(const __grid_constant__ CUtensorMap map_src0, const __grid_constant__ CUtensorMap map_src1,
                const __grid_constant__ CUtensorMap map_w, const Params p) {
  constexpr int C = H / kChunkUnits;      // cluster size == number of 32-unit chunks
  constexpr int NHS = H / kSlabK;         // K slabs of h
  constexpr int kASlab = MR * 128;        // one [MR x 64] fp16 A tile (x slabs, 128B swizzle)
  constexpr int kHTile = MR * 64;         // one [MR x 32] fp16 h tile (one chunk, 64B swizzle)
  static_assert(H == 64 || H == 128 || H == 256, "H in {64,128,256}");
  static_assert(MR == 64 || MR == 128, "MR in {64,128}");

  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) unsigned long long bars[1 + 2 * kMaxXStages + 7];
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
  const uint32_t hs_base = w_base + (uint32_t)nslabs * kWSlab;      // h operand: 2 buffers x NHS tiles
  const uint32_t xr_base = hs_base + 2u * C * kHTile;               // x ring: XS tiles (2*C*kHTile == 2*NHS*kASlab)
  const uint32_t bias_base = xr_base + (uint32_t)XS * kASlab;       // 128 floats
  float* bias_s = reinterpret_cast<float*>(smem_dyn + (bias_base - smem_addr(smem_dyn)));

  const uint32_t bar0 = smem_addr(bars);
  const uint32_t W_FULL = bar0;
  auto X_FULL = [&](int i) { return bar0 + 8u * (1 + i); };
  auto X_EMPTY = [&](int i) { return bar0 + 8u * (1 + kMaxXStages + i); };
  auto ACC_FULL = [&](int i) { return bar0 + 8u * (1 + 2 * kMaxXStages + i); };
  auto ACC_EMPTY = [&](int i) { return bar0 + 8u * (1 + 2 * kMaxXStages + 2 + i); };
  auto H_FULL = [&](int i) { return bar0 + 8u * (1 + 2 * kMaxXStages + 4 + i); };

  if (tid == 0) {
    mbar_init(W_FULL, 1);
    for (int i = 0; i < kMaxXStages; ++i) { mbar_init(X_FULL(i), 1); mbar_init(X_EMPTY(i), C); }
    for (int i = 0; i < 2; ++i) { mbar_init(ACC_FULL(i), 1); mbar_init(ACC_EMPTY(i), kEpiThreads); mbar_init(H_FULL(i), 5); }   // MMA thread's expect_tx + 4 local quadrants
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
  cluster_sync_all();      // every CTA's barriers are initialised before any multicast / remote arrive can reach them
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t tmem_c = tmem;             // cell state of this CTA's 32 units: columns [0, 32)
  const uint32_t tmem_acc = tmem + 128;     // gate accumulators, 2 buffers x 128 columns (G_x of step t+1 is produced
                                            // into the other buffer while step t is being finished)

  int coord_b = 0, coord_r0 = 0;
  long long row0;
  int valid_rows;
  if (p.axis == FNSSL_ALONG_FREQ) {
    row0 = (long long)tile * MR;
    coord_r0 = (int)row0;
    valid_rows = (int)min((long long)MR, p.rows - row0);
  } else {
    coord_b = tile / p.tiles_per_b;
    coord_r0 = (tile % p.tiles_per_b) * MR;
    row0 = (long long)coord_b * p.nf + coord_r0;
    valid_rows = min(MR, p.nf - coord_r0);
  }

  if (p.state_flags & 1) {
    // Resume from a carried state: h_{-1} takes the place step -1 would have written, i.e. all C tiles of buffer 1
    // (every CTA needs the whole h vector of its rows); c_{-1} is loaded by the epilogue warps below.
    constexpr int PPR = H / 8;                 // 16-byte pieces (8 units) per row
    for (int idx = tid; idx < MR * PPR; idx += kThreads) {
      const int r = idx / PPR, pc = idx % PPR, kc = pc >> 2, sb = pc & 3;
      uint4 pk = make_uint4(0, 0, 0, 0);
      if (r < valid_rows) {
        const float4* g = reinterpret_cast<const float4*>(p.h_state + (row0 + r) * H + pc * 8);
        const float4 a = __ldg(g), b = __ldg(g + 1);
        __half2 h01 = __floats2half2_rn(a.x, a.y), h23 = __floats2half2_rn(a.z, a.w);
        __half2 h45 = __floats2half2_rn(b.x, b.y), h67 = __floats2half2_rn(b.z, b.w);
        pk.x = *reinterpret_cast<uint32_t*>(&h01); pk.y = *reinterpret_cast<uint32_t*>(&h23);
        pk.z = *reinterpret_cast<uint32_t*>(&h45); pk.w = *reinterpret_cast<uint32_t*>(&h67);
      }
      st_shared_v4(hs_base + (uint32_t)(C + kc) * kHTile + (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u +
                       (uint32_t)((sb ^ ((r >> 1) & 3)) << 4), pk);
    }
    fence_async_smem();
    __syncthreads();
    cluster_sync_all();   // all peers have read the old state before anyone's last step overwrites it (steps == 1)
  }

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      mbar_expect_tx(W_FULL, (uint32_t)nslabs * kWSlab);
      for (int j = 0; j < nslabs; ++j)
        tma_load_2d(w_base + j * kWSlab, &map_w, W_FULL, j * kSlabK, (dir * C + (int)rank) * kChunkN);
      int n = 0;
      // x_t comes from HBM (~3.4 k cycles per TMA round trip, far more than the ring can cover): every CTA pulls its
      // share of the slabs of step t+2 into L2 while the ring works on step t / t+1, so the real loads are L2 hits
      auto prefetch_step = [&](int tt) {
        const int ss = dir ? (L - 1 - tt) : tt;
        for (int j = 0; j < nxs; ++j) {
          if ((uint32_t)((tt * nxs + j) % C) != rank) continue;
          const CUtensorMap* m = p.xs_src[j] ? &map_src1 : &map_src0;
          if (p.axis == FNSSL_ALONG_FREQ) tma_prefetch_l2_4d(m, p.xs_k0[j], ss, coord_r0, 0);
          else tma_prefetch_l2_4d(m, p.xs_k0[j], coord_r0, ss, coord_b);
        }
      };
      if (!(p.debug & 16)) { if (L > 1) prefetch_step(1); }
      for (int t = 0; t < L; ++t) {
        const int s = dir ? (L - 1 - t) : t;
        if (!(p.debug & 16) && t + 2 < L) prefetch_step(t + 2);
        for (int j = 0; j < nxs; ++j, ++n) {
          const int stage = n % XS, use = n / XS;
          if (use > 0) mbar_wait(X_EMPTY(stage), (uint32_t)((use - 1) & 1), p.error_flag, 100 + stage);
          mbar_expect_tx(X_FULL(stage), kASlab);
          if ((uint32_t)(n % C) == rank) {   // one CTA fetches the slab for the whole cluster
            const CUtensorMap* m = p.xs_src[j] ? &map_src1 : &map_src0;
            const uint32_t dst = xr_base + (uint32_t)stage * kASlab;
            if (p.axis == FNSSL_ALONG_FREQ) tma_load_4d_mc(dst, m, X_FULL(stage), p.xs_k0[j], s, coord_r0, 0, mask);
            else tma_load_4d_mc(dst, m, X_FULL(stage), p.xs_k0[j], coord_r0, s, coord_b, mask);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    // Per iteration the thread first issues the h-part of step t (accumulating onto G_x(t), which it produced during
    // the previous iteration) and only then the x-part of step t+1 into the other buffer, so the pacing of the x ring
    // stays off the recurrence's critical path.
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<MR>();
      const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0;
      mbar_wait(W_FULL, 0, p.error_flag, 200);
      int n = 0;
      auto x_part = [&](int s) {   // G_x of step s
        const int b = s & 1;
        if (s >= 2) mbar_wait(ACC_EMPTY(b), (uint32_t)(((s >> 1) - 1) & 1), p.error_flag, 201 + b);
        tc_fence_after();
        const uint32_t d_tmem = tmem_acc + (uint32_t)b * kChunkN;
        uint32_t accumulate = 0;
        long long w_acc = 0, i_acc = 0;
        for (int j = 0; j < nxs; ++j, ++n) {
          const int stage = n % XS;
          const long long c0 = tr ? clock64() : 0;
          mbar_wait(X_FULL(stage), (uint32_t)((n / XS) & 1), p.error_flag, 210 + stage);
          const long long c1 = tr ? clock64() : 0;
          w_acc += c1 - c0;
          tc_fence_after();
          const uint64_t a_desc = make_sw128_desc(xr_base + (uint32_t)stage * kASlab);
          const uint64_t b_desc = make_sw128_desc(w_base + (uint32_t)j * kWSlab);
          const int nk16 = p.xs_nk16[j];
          if (!(p.debug & 2)) {
            for (int k = 0; k < nk16; ++k) {
              umma_f16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit_mc(X_EMPTY(stage), mask);    // this CTA is done with the slab: tell every CTA's ring
          if (tr) i_acc += clock64() - c1;
        }
        if (tr && s >= 8 && s < 16) { p.trace[(s - 8) * 16 + 5] = w_acc; p.trace[(s - 8) * 16 + 6] = i_acc; }
      };
      x_part(0);
      for (int t = 0; t < L; ++t) {
        long long* tp = (tr && t >= 8 && t < 16) ? p.trace + (t - 8) * 16 : nullptr;
        if (tp) tp[0] = clock64();
        if (t > 0 || (p.state_flags & 1)) {
          if (t > 0) {
            // h_{t-1}: the C-1 remote tiles arrive as DSMEM bulk copies (tx bytes), the local tile by plain arrives
            mbar_expect_tx(H_FULL((t - 1) & 1), (uint32_t)((C - 1) * kHTile));
            if (tp) tp[1] = clock64();
            mbar_wait_cluster(H_FULL((t - 1) & 1), (uint32_t)(((t - 1) >> 1) & 1), p.error_flag, 220);
            if (tp) tp[2] = clock64();
          }   // t == 0 with a carried state: h_{-1} was placed in buffer 1 before the roles split
          tc_fence_after();
#pragma unroll
          for (int kc = 0; kc < C; ++kc) {   // K = 32 units of chunk kc: two K=16 steps; W columns inside 128B-swizzled slab kc/2
            const uint64_t a_desc = make_sw64_desc(hs_base + (uint32_t)(((t - 1) & 1) * C + kc) * kHTile);
            const uint64_t b_desc = make_sw128_desc(w_base + (uint32_t)(nxs + (kc >> 1)) * kWSlab) + 4u * (kc & 1);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              if (!(p.debug & 2)) umma_f16(tmem_acc + (uint32_t)(t & 1) * kChunkN, a_desc + 2u * k, b_desc + 2u * k, idesc, 1u);
          }
        }
        umma_commit(ACC_FULL(t & 1));     // fires when the h-part -- and G_x(t), issued earlier by this thread -- are complete
        if (tp) tp[4] = clock64();
        if (t + 1 < L) x_part(t + 1);
        if (tp) tp[3] = clock64();
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue warps ==============================
    const int q = warp & 3;                    // TMEM lane quadrant of this warp
    const int sub = (warp - 2) >> 2;           // 8-unit group of the CTA's 32 units
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int u0 = sub * 8;
    const float kL2E = 1.4426950408889634f;
    const float* bsp = bias_s + u0;            // bias of (gate, unit) at bsp[gate * 32 + e] (broadcast shared-memory reads)
    // One (row, unit): c' = sigmoid(f) c + sigmoid(i) tanh(g), h = sigmoid(o) tanh(c').  sigmoid(x) = 1/(1+2^(-x log2 e)),
    // tanh as (1-E)/(1+E); the cell update runs over ONE reciprocal: [c A B + (1-Eg) F] / (F A B), A = 1+Ei, B = 1+Eg,
    // F = 1+Ef (5 ex2 + 2 rcp per element).  Clamps keep the products finite: sigmoid(-20) = 2e-9, tanh(15) = 1 - 2e-13.
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
    // publish the quadrant's region of the CTA's own h tile: 4 warps meet, one thread pushes it to every peer
    auto publish_quadrant = [&](uint32_t buf, uint32_t hquad, uint32_t nbytes, int b) {
      fence_async_smem();
      named_bar_sync(1 + q, 128);
      if (sub == 0 && lane == 0) {
        const uint32_t hb = H_FULL(b);
#pragma unroll
        for (int dd = 1; dd < C; ++dd) {
          const uint32_t d = (rank + (uint32_t)dd) % C;
          bulk_copy_s2c(mapa_shared(buf + hquad, d), buf + hquad, nbytes, mapa_shared(hb, d));
        }
        mbar_arrive(hb);     // the local copy of this quadrant is in place
      }
    };

    if constexpr (MR == 64) {
      // ---- M = 64: the accumulator occupies lanes 0-15 of each quadrant.  16x256b TMEM accesses keep all 32 threads
      // busy: thread T owns rows {T/4, T/4+8} of the quadrant's 16 rows and units {2(T%4), 2(T%4)+1} of the 8-unit group.
      const int uo = 2 * (lane & 3);                       // unit offset inside the 8-unit group
      const int ua = (int)rank * kChunkUnits + u0 + uo;    // absolute hidden unit of the thread's first column
      long long base[2];
      long long sstride = 1;
      bool valid[2];
      uint32_t hpiece[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = q * 16 + (lane >> 2) + 8 * i;
        valid[i] = r < valid_rows;
        if (p.axis == FNSSL_ALONG_FREQ) { base[i] = (row0 + r) * p.nf; sstride = 1; }
        else { base[i] = (long long)coord_b * p.nt * p.nf + coord_r0 + r; sstride = p.nf; }
        hpiece[i] = (uint32_t)rank * kHTile + (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u +
                    (uint32_t)((sub ^ ((r >> 1) & 3)) << 4) + (uint32_t)uo * 2u;
      }
      const uint32_t hquad = (uint32_t)rank * kHTile + (uint32_t)q * 1024u;   // 16 rows x 64 B
      {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.state_flags & 1) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (valid[e >> 1]) z[e] = __ldg(p.c_state + (row0 + q * 16 + (lane >> 2) + 8 * (e >> 1)) * H + ua + (e & 1));
        }
        tmem_st4_16x256(tmem_c + lane_off + u0, z);
        tmem_wait_st();
      }
      uint32_t addn[2] = {0u, 0u};
      if (p.out1) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (valid[i]) {
            const long long pos0 = base[i] + (long long)(dir ? (L - 1) : 0) * sstride;
            addn[i] = __ldg(reinterpret_cast<const unsigned int*>(p.addend + pos0 * p.addend_ld + dir * H + ua));
          }
      }
      const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && warp == 2 && lane == 0;
      for (int t = 0; t < L; ++t) {
        const int b = t & 1;
        const int s = dir ? (L - 1 - t) : t;
        long long* tp = (tr && t >= 8 && t < 16) ? p.trace + (t - 8) * 16 : nullptr;
        if (tp) tp[8] = clock64();
        const uint32_t addc[2] = {addn[0], addn[1]};
        if (p.out1 && t + 1 < L) {     // residual operand of the next layer, fetched one step ahead
#pragma unroll
          for (int i = 0; i < 2; ++i)
            if (valid[i]) {
              const long long posn = base[i] + (long long)(dir ? (L - 2 - t) : (t + 1)) * sstride;
              addn[i] = __ldg(reinterpret_cast<const unsigned int*>(p.addend + posn * p.addend_ld + dir * H + ua));
            }
        }
        mbar_wait(ACC_FULL(b), (uint32_t)((t >> 1) & 1), p.error_flag, 300 + b);
        if (tp) tp[9] = clock64();
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
        if (tp) tp[10] = clock64();
        tc_fence_before();
        mbar_arrive(ACC_EMPTY(b));      // accumulator drained: the MMA warp may produce G_x of step t+2 into it
        float hv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int uu = uo + (e & 1);
          if (p.debug & 1) hv[e] = gi[e] + gf[e] + gg[e] + go[e] + cs[e];
          else hv[e] = lstm_cell(gi[e] + bsp[uu], gf[e] + bsp[kChunkUnits + uu], gg[e] + bsp[2 * kChunkUnits + uu],
                                 go[e] + bsp[3 * kChunkUnits + uu], cs[e]);
        }
        __half2 hp[2] = {__floats2half2_rn(hv[0], hv[1]), __floats2half2_rn(hv[2], hv[3])};
        if (tp) tp[11] = clock64();
        if (t + 1 < L) {
          const uint32_t buf = hs_base + (uint32_t)(b * C) * kHTile;
          st_shared_b32(buf + hpiece[0], *reinterpret_cast<uint32_t*>(&hp[0]));
          st_shared_b32(buf + hpiece[1], *reinterpret_cast<uint32_t*>(&hp[1]));
          publish_quadrant(buf, hquad, 1024u, b);
        }
        if (tp) tp[12] = clock64();
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && t == 12) p.trace[128 + warp] = clock64();
        tmem_st4_16x256(tmem_c + lane_off + u0, cs);
        if ((p.state_flags & 2) && t + 1 == L) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (valid[e >> 1]) {
              const long long so = (row0 + q * 16 + (lane >> 2) + 8 * (e >> 1)) * H + ua + (e & 1);
              p.c_state[so] = cs[e];
              p.h_state[so] = __half2float(__float2half_rn(hv[e]));   // the fp16 value the next step would have read
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (!valid[i]) continue;
          const long long pos = base[i] + (long long)s * sstride;
          if (p.out0) *reinterpret_cast<__half2*>(p.out0 + pos * p.out0_ld + p.out0_off + dir * H + ua) = hp[i];
          if (p.out1) {
            const __half2 av = *reinterpret_cast<const __half2*>(&addc[i]);
            *reinterpret_cast<__half2*>(p.out1 + pos * p.out1_ld + dir * H + ua) =
                __floats2half2_rn(hv[2 * i] + __low2float(av), hv[2 * i + 1] + __high2float(av));
          }
        }
        tmem_wait_st();
      }
    } else {
      // ---- M = 128: TMEM lane == row; thread = (row, 8 hidden units)
      const int r = q * 32 + lane;
      const bool valid = r < valid_rows;
      const int ua = (int)rank * kChunkUnits + u0;      // absolute hidden unit of this thread's first element
      long long base, sstride;
      if (p.axis == FNSSL_ALONG_FREQ) { base = (row0 + r) * p.nf; sstride = 1; }
      else { base = (long long)coord_b * p.nt * p.nf + coord_r0 + r; sstride = p.nf; }
      // this thread's 16-byte h piece inside the CTA's own [128 x 32] tile (64B swizzle: chunk ^= (row >> 1) & 3)
      const uint32_t hpiece = (uint32_t)rank * kHTile + (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u +
                              (uint32_t)((sub ^ ((r >> 1) & 3)) << 4);
      const uint32_t hquad = (uint32_t)rank * kHTile + (uint32_t)q * 2048u;   // the quadrant's 32 rows x 64 B
      {
        float z[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = 0.0f;
        if ((p.state_flags & 1) && valid) {
#pragma unroll
          for (int i = 0; i < 8; ++i) z[i] = __ldg(p.c_state + (row0 + r) * H + ua + i);
        }
        tmem_st8(tmem_c + lane_off + u0, z);
        tmem_wait_st();
      }
      const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && warp == 2 && lane == 0;
      // residual operand of the next layer (h + addend): fetched one whole step ahead so its HBM latency never sits
      // between the gate math and the h exchange
      uint4 addv_next = make_uint4(0, 0, 0, 0);
      if (p.out1 && valid) {
        const long long pos0 = base + (long long)(dir ? (L - 1) : 0) * sstride;
        addv_next = __ldg(reinterpret_cast<const uint4*>(p.addend + pos0 * p.addend_ld + dir * H + ua));
      }
      for (int t = 0; t < L; ++t) {
        const int b = t & 1;
        const int s = dir ? (L - 1 - t) : t;
        const long long pos = base + (long long)s * sstride;
        long long* tp = (tr && t >= 8 && t < 16) ? p.trace + (t - 8) * 16 : nullptr;
        if (tp) tp[8] = clock64();
        const uint4 addv = addv_next;
        if (p.out1 && valid && t + 1 < L) {
          const long long posn = base + (long long)(dir ? (L - 2 - t) : (t + 1)) * sstride;
          addv_next = __ldg(reinterpret_cast<const uint4*>(p.addend + posn * p.addend_ld + dir * H + ua));
        }
        mbar_wait(ACC_FULL(b), (uint32_t)((t >> 1) & 1), p.error_flag, 300 + b);
        if (tp) tp[9] = clock64();
        tc_fence_after();
        const uint32_t acc = tmem_acc + (uint32_t)b * kChunkN + lane_off + u0;
        float gti[8], gtf[8], gtg[8], gto[8], cs[8];
        tmem_ld8(acc + 0 * kChunkUnits, gti);
        tmem_ld8(acc + 1 * kChunkUnits, gtf);
        tmem_ld8(acc + 2 * kChunkUnits, gtg);
        tmem_ld8(acc + 3 * kChunkUnits, gto);
        tmem_ld8(tmem_c + lane_off + u0, cs);
        tmem_wait_ld();
        tmem_ld_dep(gti); tmem_ld_dep(gtf); tmem_ld_dep(gtg); tmem_ld_dep(gto); tmem_ld_dep(cs);
        if (tp) tp[10] = clock64();
        tc_fence_before();
        mbar_arrive(ACC_EMPTY(b));      // accumulator drained: the MMA warp may produce G_x of step t+2 into it
        float hv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (p.debug & 1) hv[e] = gti[e] + gtf[e] + gtg[e] + gto[e] + cs[e];
          else hv[e] = lstm_cell(gti[e] + bsp[e], gtf[e] + bsp[kChunkUnits + e], gtg[e] + bsp[2 * kChunkUnits + e],
                                 gto[e] + bsp[3 * kChunkUnits + e], cs[e]);
        }
        __half2 h01 = __floats2half2_rn(hv[0], hv[1]), h23 = __floats2half2_rn(hv[2], hv[3]);
        __half2 h45 = __floats2half2_rn(hv[4], hv[5]), h67 = __floats2half2_rn(hv[6], hv[7]);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h01); pk.y = *reinterpret_cast<uint32_t*>(&h23);
        pk.z = *reinterpret_cast<uint32_t*>(&h45); pk.w = *reinterpret_cast<uint32_t*>(&h67);
        if (tp) tp[11] = clock64();
        if (t + 1 < L) {
          // publish h_t: own tile locally (generic stores -> proxy fence), then one DSMEM bulk copy per peer and quadrant
          const uint32_t buf = hs_base + (uint32_t)(b * C) * kHTile;
          st_shared_v4(buf + hpiece, pk);
          publish_quadrant(buf, hquad, 2048u, b);
        }
        if (tp) tp[12] = clock64();
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && t == 12) p.trace[128 + warp] = clock64();
        tmem_st8(tmem_c + lane_off + u0, cs);
        if ((p.state_flags & 2) && t + 1 == L && valid) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            p.c_state[(row0 + r) * H + ua + i] = cs[i];
            p.h_state[(row0 + r) * H + ua + i] = __half2float(__float2half_rn(hv[i]));
          }
        }
        if (valid) {
          if (p.out0) *reinterpret_cast<uint4*>(p.out0 + pos * p.out0_ld + p.out0_off + dir * H + ua) = pk;
          if (p.out1) {
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
        tmem_wait_st();
        if (tp) tp[13] = clock64();
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

struct Plan { bool ok; int mr; int xstages; int nxs; size_t smem; };

static Plan make_plan(int H, int c0, int c1) {
  Plan pl{false, 0, 0, 0, 0};
  if (H != 64 && H != 128 && H != 256) return pl;
  if (c0 % 16 || c1 % 16 || c0 <= 0) return pl;
  const int nxs = (c0 + 63) / 64 + (c1 + 63) / 64;
  if (nxs > kMaxXSlabs) return pl;
  const int NHS = H / 64, nslabs = nxs + NHS;
  int mr_first = 128;
  if (const char* e = getenv("FNSSL_TC_ROWS")) { if (atoi(e) == 64) mr_first = 64; }   // tests / profiling
  for (int mr = mr_first; mr >= 64; mr -= 64) {
    const int aslab = mr * 128;
    const long fixed = (long)nslabs * kWSlab + 2L * NHS * aslab + kChunkN * 4 + 1024;
    long xs = (kSmemLimit - 1024 - fixed) / aslab;
    if (xs > kMaxXStages) xs = kMaxXStages;
    if (xs >= 2) {
      pl.ok = true; pl.mr = mr; pl.xstages = (int)xs; pl.nxs = nxs;
      pl.smem = (size_t)fixed + (size_t)xs * aslab;
      return pl;
    }
  }
  return pl;
}

long long* g_trace_host = nullptr;
static long long* tc2_trace_buffer() {
  static long long* dev = nullptr;
  if (!dev) {
    if (cudaHostAlloc(&g_trace_host, 160 * sizeof(long long), cudaHostAllocMapped) != cudaSuccess) return nullptr;
    for (int i = 0; i < 160; ++i) g_trace_host[i] = 0;
    if (cudaHostGetDevicePointer(&dev, g_trace_host, 0) != cudaSuccess) dev = nullptr;
  }
  return dev;
}

template <int H, int MR>
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
    tiles = (int)((p.rows + MR - 1) / MR);
  } else {
    p.rows = (long long)a->nb * a->nf; p.steps = a->nt; p.tiles_per_b = (a->nf + MR - 1) / MR;
    tiles = a->nb * p.tiles_per_b;
  }
  p.bias = reinterpret_cast<const float*>(reinterpret_cast<const char*>(a->weights) + wbytes);
  p.out0 = (__half*)a->out0; p.out0_ld = a->out0_ld; p.out0_off = a->out0_off;
  p.addend = (const __half*)a->addend; p.addend_ld = a->addend_ld;
  p.out1 = (__half*)a->out1; p.out1_ld = a->out1_ld;
  FNSSL_REQUIRE(!a->out0 || ((reinterpret_cast<uintptr_t>(a->out0) & 15) == 0 && a->out0_ld % 8 == 0 && a->out0_off % 8 == 0),
                "lstm(tcgen05): out0 must be 16-byte aligned (ld, offset multiples of 8)");
  FNSSL_REQUIRE(!a->out1 || ((reinterpret_cast<uintptr_t>(a->out1) & 15) == 0 && a->out1_ld % 8 == 0 &&
                             (reinterpret_cast<uintptr_t>(a->addend) & 15) == 0 && a->addend_ld % 8 == 0),
                "lstm(tcgen05): out1/addend must be 16-byte aligned");
  p.h_state = a->h_state; p.c_state = a->c_state; p.state_flags = a->state_flags;
  FNSSL_REQUIRE(!a->state_flags || (reinterpret_cast<uintptr_t>(a->h_state) & 15) == 0, "lstm(tcgen05): h_state not 16-byte aligned");
  p.error_flag = tc_error_flag();
  if (const char* e = getenv("FNSSL_TC_DEBUG")) p.debug = atoi(e);
  if (getenv("FNSSL_TC_TRACE")) p.trace = tc2_trace_buffer();

  CUtensorMap m0, m1, mw;
  if (make_grid_map(&m0, a->src0, a->c0, a->ld0, a->nb, a->nt, a->nf, a->axis, MR)) return 1;
  if (a->c1 > 0) { if (make_grid_map(&m1, a->src1, a->c1, a->ld1, a->nb, a->nt, a->nf, a->axis, MR)) return 1; }
  else m1 = m0;
  if (make_weight_map(&mw, a->weights, nslabs, a->num_dirs * C)) return 1;

  auto kern = lstm_tc2_kernel<H, MR>;
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
  FNSSL_LAUNCH_CHECK("lstm_tc2_kernel");
  return 0;
}

}  // namespace tc2

// diagnostic: copy the last trace (8 steps x 16 clock64 stamps) recorded with FNSSL_TC_TRACE=1
extern "C" int fnssl_lstm_tc_trace(long long* out160) {
  if (!tc2::g_trace_host) return 0;
  for (int i = 0; i < 160; ++i) out160[i] = tc2::g_trace_host[i];
  return 1;
}

bool lstm_tc2_supports(int hidden, int c0, int c1) { return tc2::make_plan(hidden, c0, c1).ok; }

int lstm_forward_tc2(const fnssl_lstm_args* a, cudaStream_t st) {
  const tc2::Plan pl = tc2::make_plan(a->hidden, a->c0, a->c1);
  FNSSL_REQUIRE(pl.ok, "lstm(tcgen05 cluster kernel): unsupported layer (H=%d c0=%d c1=%d)", a->hidden, a->c0, a->c1);
  if (a->hidden == 64) return pl.mr == 128 ? tc2::launch<64, 128>(a, pl, st) : tc2::launch<64, 64>(a, pl, st);
  if (a->hidden == 128) return pl.mr == 128 ? tc2::launch<128, 128>(a, pl, st) : tc2::launch<128, 64>(a, pl, st);
  return pl.mr == 128 ? tc2::launch<256, 128>(a, pl, st) : tc2::launch<256, 64>(a, pl, st);
}

}  // namespace fnssl
