// tcgen05 / TMEM / TMA LSTM engine (FNSSL_ENGINE_TCGEN05) for sm_100a.
//
// Replaces nn.LSTM at FN-SSL/Lightning/Model.py:38,46 and IPDnet/FixedAarryIPDnet.py:32,36 plus the layout
// glue around it (Model.py:35-37,41-45,49): one persistent launch time-steps ALL sequences of a layer.
//
// Work decomposition.  A CTA owns a tile of MR = 128 or 64 sequences ("rows": (b,t) pairs for the full-band
// pass, (b,f) pairs for the narrow-band pass) of one direction for all L steps -- there is no inter-CTA
// communication.  Per step the gate pre-activations  G[MR, 4H] = [x_t | h_{t-1}] . W^T  are produced by
// tcgen05.mma (kind::f16: fp16 operands, fp32 accumulation in TMEM) in chunks of 32 hidden units
// (N = 128 gate columns = i,f,g,o of those units); the cell state c (fp32) lives in TMEM for the whole
// launch, h_t goes to shared memory as the next step's A operand (fp16, 128B-swizzled K-major) and to HBM.
//
// MR = 128: one chunk per accumulator buffer, TMEM lane = row.  MR = 64 (small batches: twice the CTAs; and
// H = 256, whose state does not fit one SM at 128 rows): an M=64 accumulator occupies lanes 0-15 of each
// 32-lane quadrant, so two chunks are interleaved in one buffer (second chunk at lane offset 16) and a
// 32x32b TMEM load hands threads 0-15 the rows of chunk A and threads 16-31 the same rows of chunk B.
//
//   warp 0        TMA producer: x_t slabs (HBM -> smem, once per step) and the weight slab ring (L2 -> smem)
//   warp 1        MMA issuer (one elected lane), TMEM allocator
//   warps 2..17   epilogue (4 warps per TMEM lane quadrant, 8 hidden units of the chunk each): tcgen05.ld gates,
//                 sigmoid/tanh, c/h update, tcgen05.st c, h -> smem + global
//
// Pipelines (all mbarrier based): weight ring full/empty, per-slab x full/empty, 3 accumulator buffers
// full/empty (MMA of chunk k+1.. overlaps the epilogue of chunk k), h_t complete -> MMA of step t+1.
// The x-part of a chunk is issued before its h-part so the tensor pipe has work while h_t is finished.
//
// Every mbarrier wait is bounded (~1 s) and traps on timeout: a protocol bug fails the launch instead of
// hanging the GPU.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace fnssl {

constexpr int kTcThreads = 576;          // 2 control warps + 16 epilogue warps
constexpr int kEpiThreads = 512;
constexpr int kSlabK = 64;             // fp16 elements per 128-byte swizzle row
constexpr int kWSlabBytes = 128 * 128;  // one [128 gate columns x 64] fp16 weight tile
constexpr int kChunkUnits = 32;
constexpr int kChunkN = 4 * kChunkUnits;  // 128 gate columns per chunk (UMMA N)
constexpr int kNumAcc = 3;             // accumulator buffers in TMEM
constexpr int kMaxXSlabs = 6;
constexpr int kMaxWStages = 6;
constexpr int kSmemLimit = 232448;     // 227 KB

struct TcParams {
  int nxs;                        // x slabs
  int xs_src[kMaxXSlabs];         // 0 = src0, 1 = src1
  int xs_k0[kMaxXSlabs];          // first channel of the slab within its source
  int xs_nk16[kMaxXSlabs];        // valid K=16 steps in the slab (1..4)
  int wstages;
  int nxb;                        // x_t buffers: 2 = x_{t+1} is prefetched a whole step ahead, 1 = single-buffered
  int steps, axis, nf, nt;
  long long rows;
  int tiles_per_b;                // narrow-band: row tiles per utterance
  const float* bias;              // [dirs][4H] in accumulator column order [chunk][gate][unit]
  __half* out0; int out0_ld; int out0_off;
  const __half* addend; int addend_ld;
  __half* out1; int out1_ld;
  int* error_flag;
  int debug;                      // timing experiments only (FNSSL_TC_DEBUG): 1 = skip gate math, 2 = skip MMA issue
};

// instruction descriptor, kind::f16: D = f32 (bit 4), A = B = f16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
template <int MR>
constexpr uint32_t make_idesc() { return (1u << 4) | ((uint32_t)(kChunkN >> 3) << 17) | ((uint32_t)(MR >> 4) << 24); }

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------

template <int H, int MR>
__global__ void __launch_bounds__(kTcThreads, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap map_src0, const __grid_constant__ CUtensorMap map_src1,
               const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  constexpr int NCH = H / kChunkUnits;  // chunks per step
  constexpr int NHS = H / kSlabK;       // h slabs (H = 64 -> 1, 128 -> 2, 256 -> 4)
  constexpr int CPG = (MR == 64) ? 2 : 1;   // chunks per accumulator buffer ("group")
  constexpr int NG = NCH / CPG;             // groups per step
  constexpr int kASlab = MR * 128;          // bytes of one [MR x 64] fp16 A-operand tile
  static_assert((MR == 128 && (H == 64 || H == 128)) || (MR == 64 && (H == 64 || H == 128 || H == 256)),
                "128-row tiles hold H <= 128; 64-row tiles hold H <= 256");

  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) unsigned long long bars[2 * kMaxWStages + 4 * kMaxXSlabs + 2 * kNumAcc + 2];
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;
  const int tile = blockIdx.x;
  const int nxs = p.nxs, S = p.wstages, L = p.steps;

  // carve dynamic smem (1024-byte aligned operand tiles)
  const uint32_t dyn0 = (smem_addr(smem_dyn) + 1023u) & ~1023u;
  const uint32_t xs_base = dyn0;                                   // nxs tiles
  const uint32_t hs_base = xs_base + (uint32_t)(p.nxb * nxs) * kASlab;   // 2 * NHS tiles
  const uint32_t ws_base = hs_base + 2u * NHS * kASlab;            // S tiles
  const uint32_t bias_base = ws_base + (uint32_t)S * kWSlabBytes;  // 4H floats
  unsigned char* dyn_gen = smem_dyn + (dyn0 - smem_addr(smem_dyn));
  float* bias_s = reinterpret_cast<float*>(dyn_gen + (bias_base - dyn0));
  unsigned char* hs_gen = dyn_gen + (hs_base - dyn0);

  // barrier map
  const uint32_t bar0 = smem_addr(bars);
  auto W_FULL = [&](int i) { return bar0 + 8u * i; };
  auto W_EMPTY = [&](int i) { return bar0 + 8u * (kMaxWStages + i); };
  auto X_FULL = [&](int b, int i) { return bar0 + 8u * (2 * kMaxWStages + b * kMaxXSlabs + i); };
  auto X_EMPTY = [&](int b, int i) { return bar0 + 8u * (2 * kMaxWStages + 2 * kMaxXSlabs + b * kMaxXSlabs + i); };
  auto ACC_FULL = [&](int i) { return bar0 + 8u * (2 * kMaxWStages + 4 * kMaxXSlabs + i); };
  auto ACC_EMPTY = [&](int i) { return bar0 + 8u * (2 * kMaxWStages + 4 * kMaxXSlabs + kNumAcc + i); };
  auto H_FULL = [&](int i) { return bar0 + 8u * (2 * kMaxWStages + 4 * kMaxXSlabs + 2 * kNumAcc + i); };

  if (tid == 0) {
    for (int i = 0; i < kMaxWStages; ++i) { mbar_init(W_FULL(i), 1); mbar_init(W_EMPTY(i), 1); }
    for (int b = 0; b < 2; ++b)
      for (int i = 0; i < kMaxXSlabs; ++i) { mbar_init(X_FULL(b, i), 1); mbar_init(X_EMPTY(b, i), 1); }
    for (int i = 0; i < kNumAcc; ++i) { mbar_init(ACC_FULL(i), 1); mbar_init(ACC_EMPTY(i), kEpiThreads); }
    for (int i = 0; i < 2; ++i) mbar_init(H_FULL(i), kEpiThreads * NG);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&map_src0); prefetch_tmap(&map_src1); prefetch_tmap(&map_w); }
  if (warp == 1) {  // TMEM: all 512 columns (1 CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_base_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 4 * H; i += kTcThreads) bias_s[i] = p.bias[dir * 4 * H + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t tmem_c = tmem;                 // cell state: columns [0, H / CPG)
  const uint32_t tmem_acc = tmem + 128;         // accumulators: 3 x 128 columns

  // tile coordinates
  int coord_b = 0, coord_r0 = 0;
  long long row0;
  int valid_rows;
  if (p.axis == FNSSL_ALONG_FREQ) {
    row0 = (long long)tile * MR;
    coord_r0 = (int)row0;
    valid_rows = (int)min((long long)MR, p.rows - row0);
  } else {
    coord_b = tile / p.tiles_per_b;
    coord_r0 = (tile % p.tiles_per_b) * MR;     // first bin of the tile
    row0 = (long long)coord_b * p.nf + coord_r0;
    valid_rows = min(MR, p.nf - coord_r0);
  }
  const int nslabs = nxs + NHS;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      const int NXB = p.nxb;
      // x_tt slab j -> buffer tt % NXB (waits until the MMAs of step tt - NXB have released it)
      auto load_x = [&](int tt, int j) {
        const int b = tt % NXB, use = tt / NXB;
        const int ss = dir ? (L - 1 - tt) : tt;
        if (use > 0) mbar_wait(X_EMPTY(b, j), (uint32_t)((use - 1) & 1), p.error_flag, 100 + j);
        mbar_expect_tx(X_FULL(b, j), kASlab);
        const CUtensorMap* m = p.xs_src[j] ? &map_src1 : &map_src0;
        const uint32_t dst = xs_base + (uint32_t)(b * nxs + j) * kASlab;
        if (p.axis == FNSSL_ALONG_FREQ) tma_load_4d(dst, m, X_FULL(b, j), p.xs_k0[j], ss, coord_r0, 0);
        else tma_load_4d(dst, m, X_FULL(b, j), p.xs_k0[j], coord_r0, ss, coord_b);
      };
      int ws = 0;
      uint32_t wphase = 0;
      if (NXB == 2) for (int j = 0; j < nxs; ++j) load_x(0, j);
      for (int t = 0; t < L; ++t) {
        if (NXB == 2 && t + 1 < L) for (int j = 0; j < nxs; ++j) load_x(t + 1, j);   // a whole step ahead
        for (int c = 0; c < NCH; ++c) {
          for (int j = 0; j < nslabs; ++j) {
            if (j >= nxs && t == 0) continue;  // h_{-1} = 0: no recurrent term at the first step
            if (NXB == 1 && c == 0 && j < nxs) load_x(t, j);
            mbar_wait(W_EMPTY(ws), wphase ^ 1u, p.error_flag, 110);
            mbar_expect_tx(W_FULL(ws), kWSlabBytes);
            tma_load_2d(ws_base + ws * kWSlabBytes, &map_w, W_FULL(ws), j * kSlabK, (dir * NCH + c) * kChunkN);
            if (++ws == S) { ws = 0; wphase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      int ws = 0;
      uint32_t wphase = 0;
      int g = 0;
      constexpr uint32_t idesc = make_idesc<MR>();
      for (int t = 0; t < L; ++t) {
        const int xb = t % p.nxb;
        for (int gi = 0; gi < NG; ++gi, ++g) {
          const int buf = g % kNumAcc;
          const int use = g / kNumAcc;
          if (use > 0) mbar_wait(ACC_EMPTY(buf), (uint32_t)((use - 1) & 1), p.error_flag, 200 + buf);
          tc_fence_after();
#pragma unroll
          for (int cc = 0; cc < CPG; ++cc) {
            const int c = gi * CPG + cc;
            // M = 64: the second chunk of the pair lives in lanes 16-31 of every quadrant
            const uint32_t d_tmem = tmem_acc + (uint32_t)buf * kChunkN + ((uint32_t)(cc * 16) << 16);
            uint32_t accumulate = 0;
            for (int j = 0; j < nslabs; ++j) {
              if (j >= nxs && t == 0) continue;
              uint32_t a_tile;
              int nk16;
              if (j < nxs) {
                if (c == 0) mbar_wait(X_FULL(xb, j), (uint32_t)((t / p.nxb) & 1), p.error_flag, 210 + j);
                a_tile = xs_base + (uint32_t)(xb * nxs + j) * kASlab;
                nk16 = p.xs_nk16[j];
              } else {
                if (c == 0 && j == nxs) mbar_wait(H_FULL((t - 1) & 1), (uint32_t)(((t - 1) >> 1) & 1), p.error_flag, 220);
                a_tile = hs_base + (uint32_t)(((t - 1) & 1) * NHS + (j - nxs)) * kASlab;
                nk16 = 4;
              }
              mbar_wait(W_FULL(ws), wphase, p.error_flag, 230);
              tc_fence_after();
              const uint64_t a_desc = make_sw128_desc(a_tile);
              const uint64_t b_desc = make_sw128_desc(ws_base + ws * kWSlabBytes);
              if (!(p.debug & 2)) {
                for (int k = 0; k < nk16; ++k) {
                  umma_f16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, accumulate);  // +32 B per K=16 step
                  accumulate = 1;
                }
              }
              umma_commit(W_EMPTY(ws));                               // weight stage free once these MMAs retire
              if (c == NCH - 1 && j < nxs) umma_commit(X_EMPTY(xb, j));   // x_t slab free for a later step
              if (++ws == S) { ws = 0; wphase ^= 1u; }
            }
          }
          umma_commit(ACC_FULL(buf));
        }
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue warps ==============================
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int sub = (warp - 2) >> 2;        // which 8 of the chunk's 32 units (0..3)
    // MR = 128: row == TMEM lane.  MR = 64: lanes 0-15 of the quadrant hold rows 16q..16q+15 of the pair's first
    // chunk, lanes 16-31 the same rows of its second chunk.
    const int r = (MR == 128) ? (q * 32 + lane) : (q * 16 + (lane & 15));
    const int cc_lane = (MR == 128) ? 0 : (lane >> 4);
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const bool valid = r < valid_rows;
    long long base;
    long long sstride;
    if (p.axis == FNSSL_ALONG_FREQ) { base = (row0 + r) * p.nf; sstride = 1; }
    else { base = (long long)coord_b * p.nt * p.nf + coord_r0 + r; sstride = p.nf; }

    {  // c_0 = 0
      float z[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) z[i] = 0.0f;
      constexpr int CC = H / CPG;   // cell-state columns
      for (int col = sub * (CC / 4); col < (sub + 1) * (CC / 4); col += 8) tmem_st8(tmem_c + lane_off + col, z);
      tmem_wait_st();
    }
    const float kL2E = 1.4426950408889634f;
    int g = 0;
    for (int t = 0; t < L; ++t) {
      const int s = dir ? (L - 1 - t) : t;
      const long long pos = base + (long long)s * sstride;
      unsigned char* hbuf = hs_gen + (size_t)((t & 1) * NHS) * kASlab;
      for (int gi = 0; gi < NG; ++gi, ++g) {
        const int buf = g % kNumAcc;
        const int c = gi * CPG + cc_lane;            // this thread's chunk
        const int u0 = sub * 8;                      // unit offset inside the chunk
        const int ua = c * kChunkUnits + u0;         // absolute hidden unit
        // residual operand of the next layer: independent of the MMA, fetched before waiting for it
        uint4 addv = make_uint4(0, 0, 0, 0);
        if (p.out1 && valid) addv = __ldg(reinterpret_cast<const uint4*>(p.addend + pos * p.addend_ld + dir * H + ua));
        mbar_wait(ACC_FULL(buf), (uint32_t)((g / kNumAcc) & 1), p.error_flag, 300 + buf);
        tc_fence_after();
        const uint32_t acc = tmem_acc + (uint32_t)buf * kChunkN + lane_off;
        {
          const uint32_t ccol = tmem_c + lane_off + (uint32_t)(gi * kChunkUnits + u0);
          float gti[8], gtf[8], gtg[8], gto[8], cs[8];
          tmem_ld8(acc + 0 * kChunkUnits + u0, gti);
          tmem_ld8(acc + 1 * kChunkUnits + u0, gtf);
          tmem_ld8(acc + 2 * kChunkUnits + u0, gtg);
          tmem_ld8(acc + 3 * kChunkUnits + u0, gto);
          tmem_ld8(ccol, cs);
          tmem_wait_ld();
          tmem_ld_dep(gti); tmem_ld_dep(gtf); tmem_ld_dep(gtg); tmem_ld_dep(gto); tmem_ld_dep(cs);
          const float* bs = bias_s + c * kChunkN + u0;
          float hv[8];
          if (p.debug & 1) {
#pragma unroll
            for (int e = 0; e < 8; ++e) hv[e] = gti[e] + gtf[e] + gtg[e] + gto[e] + cs[e];
          } else
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            // sigmoid(x) = 1/(1+2^(-x log2 e)) needs no clamp (2^big = inf -> 1/inf = 0); tanh(g) and tanh(c') are
            // written as (1-E)/(1+E) and share a reciprocal with a sigmoid, so E must stay finite: clamp to +-15.
            const float xg = fminf(fmaxf(gtg[e] + bs[2 * kChunkUnits + e], -15.f), 15.f);
            const float ei = ex2_approx(-kL2E * (gti[e] + bs[e]));
            const float ef = ex2_approx(-kL2E * (gtf[e] + bs[kChunkUnits + e]));
            const float eg = ex2_approx(-2.0f * kL2E * xg);
            const float eo = ex2_approx(-kL2E * (gto[e] + bs[3 * kChunkUnits + e]));
            // c' = sigmoid(f) c + sigmoid(i) tanh(g);  h = sigmoid(o) tanh(c')
            const float cn = cs[e] * rcp_approx(1.0f + ef) + (1.0f - eg) * rcp_approx((1.0f + ei) * (1.0f + eg));
            cs[e] = cn;
            const float ec = ex2_approx(-2.0f * kL2E * fminf(fmaxf(cn, -15.f), 15.f));
            hv[e] = (1.0f - ec) * rcp_approx((1.0f + eo) * (1.0f + ec));
          }
          tmem_st8(ccol, cs);
          // h_t -> fp16: operand tile of the next step (128B swizzle: 16-byte chunk index ^= row & 7) and HBM
          __half2 h01 = __floats2half2_rn(hv[0], hv[1]), h23 = __floats2half2_rn(hv[2], hv[3]);
          __half2 h45 = __floats2half2_rn(hv[4], hv[5]), h67 = __floats2half2_rn(hv[6], hv[7]);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h01); pk.y = *reinterpret_cast<uint32_t*>(&h23);
          pk.z = *reinterpret_cast<uint32_t*>(&h45); pk.w = *reinterpret_cast<uint32_t*>(&h67);
          {
            const int slab = ua >> 6;
            const int chunk16 = (ua & 63) >> 3;
            unsigned char* dst = hbuf + (size_t)slab * kASlab + (r >> 3) * 1024 + (r & 7) * 128 + ((chunk16 ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = pk;
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
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(ACC_EMPTY(buf));   // accumulator drained (all tcgen05.ld of this thread completed)
        fence_async_smem();            // h_t writes (generic proxy) -> visible to the tensor core (async proxy)
        mbar_arrive(H_FULL(t & 1));
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

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// 4-D fp16 tensor map over a grid, 128B swizzle, zero OOB fill
static int make_map4(CUtensorMap* m, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                     const uint32_t box[4]) {
  auto enc = get_encode();
  FNSSL_REQUIRE(enc, "lstm(tcgen05): cuTensorMapEncodeTiled is unavailable in this driver");
  const uint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FNSSL_REQUIRE(r == CUDA_SUCCESS, "lstm(tcgen05): cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

int make_grid_map(CUtensorMap* m, const void* base, int c, int ld, int nb, int nt, int nf, int axis, int mr) {
  FNSSL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld % 8) == 0, "lstm(tcgen05): grid base / channel stride not 16-byte aligned");
  if (axis == FNSSL_ALONG_FREQ) {
    const uint64_t dims[4] = {(uint64_t)c, (uint64_t)nf, (uint64_t)nb * nt, 1};
    const uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)nf * ld * 2, (uint64_t)nb * nt * nf * ld * 2};
    const uint32_t box[4] = {kSlabK, 1, (uint32_t)mr, 1};
    return make_map4(m, base, dims, str, box);
  }
  const uint64_t dims[4] = {(uint64_t)c, (uint64_t)nf, (uint64_t)nt, (uint64_t)nb};
  const uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)nf * ld * 2, (uint64_t)nt * nf * ld * 2};
  const uint32_t box[4] = {kSlabK, (uint32_t)mr, 1, 1};
  return make_map4(m, base, dims, str, box);
}

int make_small_grid_map(CUtensorMap* m, const void* base, int c, int ld, int nb, int nt, int nf, int axis, int mr) {
  auto enc = get_encode();
  FNSSL_REQUIRE(enc, "lstm(tcgen05): cuTensorMapEncodeTiled is unavailable in this driver");
  FNSSL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld % 8) == 0 && c <= 16,
                "lstm(tcgen05): small-source grid base / channel stride not 16-byte aligned or wider than 16 channels");
  uint64_t dims[4], str[3];
  uint32_t box[4];
  if (axis == FNSSL_ALONG_FREQ) {
    dims[0] = (uint64_t)c; dims[1] = (uint64_t)nf; dims[2] = (uint64_t)nb * nt; dims[3] = 1;
    str[0] = (uint64_t)ld * 2; str[1] = (uint64_t)nf * ld * 2; str[2] = (uint64_t)nb * nt * nf * ld * 2;
    box[0] = 16; box[1] = 1; box[2] = (uint32_t)mr; box[3] = 1;
  } else {
    dims[0] = (uint64_t)c; dims[1] = (uint64_t)nf; dims[2] = (uint64_t)nt; dims[3] = (uint64_t)nb;
    str[0] = (uint64_t)ld * 2; str[1] = (uint64_t)nf * ld * 2; str[2] = (uint64_t)nt * nf * ld * 2;
    box[0] = 16; box[1] = (uint32_t)mr; box[2] = 1; box[3] = 1;
  }
  const uint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, str, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FNSSL_REQUIRE(r == CUDA_SUCCESS, "lstm(tcgen05): small-source tensor map failed (%d)", (int)r);
  return 0;
}

int make_small_weight_map(CUtensorMap* m, const void* weights, int nslabs, int nchunks_total) {
  auto enc = get_encode();
  FNSSL_REQUIRE(enc, "lstm(tcgen05): cuTensorMapEncodeTiled is unavailable in this driver");
  const uint64_t dims[2] = {(uint64_t)nslabs * kSlabK, (uint64_t)nchunks_total * kChunkN};
  const uint64_t str[1] = {(uint64_t)nslabs * kSlabK * 2};
  const uint32_t box[2] = {16, kChunkN};
  const uint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(weights), dims, str, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FNSSL_REQUIRE(r == CUDA_SUCCESS, "lstm(tcgen05): small weight tensor map failed (%d)", (int)r);
  return 0;
}

int make_out_map(CUtensorMap* m, const void* base, int ld, int nb, int nt, int nf, int axis, int rows) {
  auto enc = get_encode();
  FNSSL_REQUIRE(enc, "lstm(tcgen05): cuTensorMapEncodeTiled is unavailable in this driver");
  FNSSL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld % 8) == 0, "lstm(tcgen05): output grid base / channel stride not 16-byte aligned");
  uint64_t dims[4], str[3];
  uint32_t box[4];
  if (axis == FNSSL_ALONG_FREQ) {
    dims[0] = (uint64_t)ld; dims[1] = (uint64_t)nf; dims[2] = (uint64_t)nb * nt; dims[3] = 1;
    str[0] = (uint64_t)ld * 2; str[1] = (uint64_t)nf * ld * 2; str[2] = (uint64_t)nb * nt * nf * ld * 2;
    box[0] = 32; box[1] = 1; box[2] = (uint32_t)rows; box[3] = 1;
  } else {
    dims[0] = (uint64_t)ld; dims[1] = (uint64_t)nf; dims[2] = (uint64_t)nt; dims[3] = (uint64_t)nb;
    str[0] = (uint64_t)ld * 2; str[1] = (uint64_t)nf * ld * 2; str[2] = (uint64_t)nt * nf * ld * 2;
    box[0] = 32; box[1] = (uint32_t)rows; box[2] = 1; box[3] = 1;
  }
  const uint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, str, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FNSSL_REQUIRE(r == CUDA_SUCCESS, "lstm(tcgen05): output tensor map failed (%d)", (int)r);
  return 0;
}

// 2-D fp16 map over the packed weights [nchunks_total * 128 rows][nslabs * 64], box = one [128 x 64] slab
int make_weight_map(CUtensorMap* m, const void* weights, int nslabs, int nchunks_total) {
  auto enc = get_encode();
  FNSSL_REQUIRE(enc, "lstm(tcgen05): cuTensorMapEncodeTiled is unavailable in this driver");
  const uint64_t dims[2] = {(uint64_t)nslabs * kSlabK, (uint64_t)nchunks_total * kChunkN};
  const uint64_t str[1] = {(uint64_t)nslabs * kSlabK * 2};
  const uint32_t box[2] = {kSlabK, kChunkN};
  const uint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(weights), dims, str, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FNSSL_REQUIRE(r == CUDA_SUCCESS, "lstm(tcgen05): weight tensor map failed (%d)", (int)r);
  return 0;
}

int* g_flag_host = nullptr;
int* tc_error_flag() {   // one host-mapped int per process; written by mbar_timeout, readable after a trap
  static int* flag_dev = nullptr;
  if (!flag_dev) {
    if (cudaHostAlloc(&g_flag_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess) return nullptr;
    *g_flag_host = 0;
    if (cudaHostGetDevicePointer(&flag_dev, g_flag_host, 0) != cudaSuccess) { flag_dev = nullptr; return nullptr; }
  }
  return flag_dev;
}

bool lstm_tc_supports(const fnssl_lstm_args* a) {
  if (a->dtype != FNSSL_F16) return false;
  if (a->hidden != 64 && a->hidden != 128 && a->hidden != 256) return false;
  if (a->c0 % 16 || a->c1 % 16) return false;
  const int nxs = (a->c0 + 63) / 64 + (a->c1 + 63) / 64;
  return nxs <= kMaxXSlabs;
}

template <int H, int MR>
static int launch_tc(const fnssl_lstm_args* a, cudaStream_t st) {
  constexpr int NCH = H / kChunkUnits, NHS = H / kSlabK;
  constexpr int kASlab = MR * 128;
  TcParams p{};
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
  const int nslabs = nxs + NHS;
  const int64_t wbytes = (int64_t)a->num_dirs * NCH * kChunkN * nslabs * kSlabK * 2;
  const int64_t need = wbytes + (int64_t)a->num_dirs * 4 * H * 4;
  FNSSL_REQUIRE(a->weights_bytes == need, "lstm(tcgen05): packed weight buffer is %lld bytes, expected %lld",
                (long long)a->weights_bytes, (long long)need);
  FNSSL_REQUIRE((reinterpret_cast<uintptr_t>(a->weights) & 15) == 0, "lstm(tcgen05): weights not 16-byte aligned");
  // x_t double-buffered (prefetch a whole step ahead) when that still leaves a 4-stage weight ring
  int nxb = 2;
  int fixed = (nxb * nxs + 2 * NHS) * kASlab + 4 * H * 4 + 1024;
  if ((kSmemLimit - 2048 - fixed) / kWSlabBytes < 4) {
    nxb = 1;
    fixed = (nxb * nxs + 2 * NHS) * kASlab + 4 * H * 4 + 1024;
  }
  p.nxb = nxb;
  int S = (kSmemLimit - 2048 - fixed) / kWSlabBytes;
  if (S > kMaxWStages) S = kMaxWStages;
  FNSSL_REQUIRE(S >= 2, "lstm(tcgen05): layer does not fit in shared memory (c0=%d c1=%d H=%d)", a->c0, a->c1, H);
  p.wstages = S;
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
  // 16-byte vector stores / loads in the epilogue
  FNSSL_REQUIRE(!a->out0 || ((reinterpret_cast<uintptr_t>(a->out0) & 15) == 0 && a->out0_ld % 8 == 0 && a->out0_off % 8 == 0),
                "lstm(tcgen05): out0 must be 16-byte aligned (ld, offset multiples of 8)");
  FNSSL_REQUIRE(!a->out1 || ((reinterpret_cast<uintptr_t>(a->out1) & 15) == 0 && a->out1_ld % 8 == 0 &&
                             (reinterpret_cast<uintptr_t>(a->addend) & 15) == 0 && a->addend_ld % 8 == 0),
                "lstm(tcgen05): out1/addend must be 16-byte aligned");
  p.error_flag = tc_error_flag();
  if (const char* e = getenv("FNSSL_TC_DEBUG")) p.debug = atoi(e);

  CUtensorMap m0, m1, mw;
  if (make_grid_map(&m0, a->src0, a->c0, a->ld0, a->nb, a->nt, a->nf, a->axis, MR)) return 1;
  if (a->c1 > 0) { if (make_grid_map(&m1, a->src1, a->c1, a->ld1, a->nb, a->nt, a->nf, a->axis, MR)) return 1; }
  else m1 = m0;
  if (make_weight_map(&mw, a->weights, nslabs, a->num_dirs * NCH)) return 1;
  const size_t smem = (size_t)fixed + (size_t)S * kWSlabBytes;
  FNSSL_CUDA(cudaFuncSetAttribute(lstm_tc_kernel<H, MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(tiles, a->num_dirs);
  lstm_tc_kernel<H, MR><<<grid, kTcThreads, smem, st>>>(m0, m1, mw, p);
  FNSSL_LAUNCH_CHECK("lstm_tc_kernel");
  return 0;
}

bool lstm_tc2_supports(int hidden, int c0, int c1);
int lstm_forward_tc2(const fnssl_lstm_args* a, cudaStream_t st);
bool lstm_tc3_supports(int hidden, int c0, int c1);
int lstm_forward_tc3(const fnssl_lstm_args* a, cudaStream_t st);
bool lstm_tc4_supports(int hidden, int c0, int c1);
int lstm_forward_tc4(const fnssl_lstm_args* a, cudaStream_t st);

int lstm_forward_tc(const fnssl_lstm_args* a, cudaStream_t st) {
  FNSSL_REQUIRE(a->dtype == FNSSL_F16, "lstm(tcgen05): grids must be fp16");
  {
    // Kernel generation: 4 = cluster-resident weights, two independent row tiles per cluster run half a step apart
    // (lstm_tc4.cu; every H), 3 = cluster-resident weights with two interleaved 64-row sub-tiles (lstm_tc3.cu; H <= 128),
    // 2 = cluster-resident weights, one tile (lstm_tc2.cu; also H = 256), 1 = weight-streaming kernel below.
    // Default: the highest generation that supports the layer.  FNSSL_TC_KERNEL caps it (tests / profiling).
    const char* e = getenv("FNSSL_TC_KERNEL");
    const int want = e ? atoi(e) : 4;
    const bool aligned = a->c0 % 16 == 0 && a->c1 % 16 == 0;
    // Generation 3 issues its x-part as M = 64 MMAs (twice the tensor-pipe time and twice the x-ring hand-shakes of
    // generation 2), which only pays off when the input projection is small: measured on cfg2 it is 1.06 vs 1.46 ms
    // for the 16-channel first layer but 1.80 vs 1.59 ms for the 256-channel layers -> use it for <= 2 input slabs.
    const int nxs = (a->c0 + 63) / 64 + (a->c1 + 63) / 64;
    const bool force3 = e && atoi(e) == 3;
    if (want >= 4 && aligned && lstm_tc4_supports(a->hidden, a->c0, a->c1)) return lstm_forward_tc4(a, st);
    if (a->state_flags) {   // carried (h, c): generations 4 (above) and 2
      FNSSL_REQUIRE(aligned && lstm_tc2_supports(a->hidden, a->c0, a->c1),
                    "lstm(tcgen05): recurrent state needs the cluster kernel, which does not support H=%d c0=%d c1=%d", a->hidden, a->c0, a->c1);
      return lstm_forward_tc2(a, st);
    }
    if (want >= 3 && aligned && (nxs <= 2 || force3) && lstm_tc3_supports(a->hidden, a->c0, a->c1)) return lstm_forward_tc3(a, st);
    if (want >= 2 && aligned && lstm_tc2_supports(a->hidden, a->c0, a->c1)) return lstm_forward_tc2(a, st);
  }
  FNSSL_REQUIRE(a->c0 % 16 == 0 && a->c1 % 16 == 0, "lstm(tcgen05): channel counts must be multiples of 16 (got %d, %d); pad the grid",
                a->c0, a->c1);
  FNSSL_REQUIRE((a->c0 + 63) / 64 + (a->c1 + 63) / 64 <= kMaxXSlabs, "lstm(tcgen05): too many input channels (%d + %d)", a->c0, a->c1);
  // Row-tile size: 128 rows per CTA is the efficient shape (full-rate M=128 MMA) but needs enough tiles to fill
  // the 148 SMs; below ~3/4 of a wave (small batches) 64-row tiles double the CTA count, and H = 256 only fits
  // one SM with 64 rows.  FNSSL_TC_ROWS=64|128 overrides (tests / profiling).
  const long long rows = (a->axis == FNSSL_ALONG_FREQ) ? (long long)a->nb * a->nt : (long long)a->nb * a->nf;
  const long long tiles128 = (a->axis == FNSSL_ALONG_FREQ) ? (rows + 127) / 128 : (long long)a->nb * ((a->nf + 127) / 128);
  bool use64 = a->hidden == 256 || tiles128 * a->num_dirs < 111;
  if (const char* e = getenv("FNSSL_TC_ROWS")) {
    if (atoi(e) == 64) use64 = true;
    if (atoi(e) == 128 && a->hidden != 256) use64 = false;
  }
  switch (a->hidden) {
    case 64: return use64 ? launch_tc<64, 64>(a, st) : launch_tc<64, 128>(a, st);
    case 128: return use64 ? launch_tc<128, 64>(a, st) : launch_tc<128, 128>(a, st);
    case 256: return launch_tc<256, 64>(a, st);
    default: FNSSL_FAIL("lstm(tcgen05): hidden size %d is not built (64, 128, 256)", a->hidden);
  }
}

}  // namespace fnssl

extern "C" int fnssl_lstm_tc_supported(int hidden, int c0, int c1) {
  fnssl_lstm_args a{};
  a.dtype = FNSSL_F16; a.hidden = hidden; a.c0 = c0; a.c1 = c1;
  return fnssl::lstm_tc_supports(&a) ? 1 : 0;
}

extern "C" int fnssl_lstm_tc_error_site(void) {
  // site code recorded by a timed-out mbarrier wait (0 = none); resets the flag
  if (!fnssl::g_flag_host) return 0;
  const int v = *reinterpret_cast<volatile int*>(fnssl::g_flag_host);
  *fnssl::g_flag_host = 0;
  return v;
}
