// Host side of the tcgen05 LSTM engine (FNSSL_ENGINE_TCGEN05): TMA tensor maps over the channels-last grids and the packed
// weights, the per-process error flag of the bounded pipeline waits, and the dispatch to the cluster kernel (lstm_tc4.cu).
//
// The engine replaces nn.LSTM at FN-SSL/Lightning/Model.py:38,46 and IPDnet/FixedAarryIPDnet.py:32,36 plus the layout glue
// around it (Model.py:35-37,41-45,49).  Earlier kernel generations (weight streaming; one tile per cluster; two interleaved
// 64-row sub-tiles) are kept under tools/legacy_kernels/ for reference only -- they are not part of the library.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "tc_common.cuh"

namespace fnssl {

namespace {
constexpr int kSlabK = 64;     // fp16 elements per 128-byte swizzle row
constexpr int kChunkN = 128;   // gate columns (i,f,g,o of 32 hidden units) per accumulator chunk

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

int encode(CUtensorMap* m, int rank, const void* base, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
           CUtensorMapSwizzle swz, CUtensorMapL2promotion l2, const char* what) {
  auto enc = get_encode();
  FNSSL_REQUIRE(enc, "lstm(tcgen05): cuTensorMapEncodeTiled is unavailable in this driver");
  const uint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FNSSL_REQUIRE(r == CUDA_SUCCESS, "lstm(tcgen05): %s tensor map failed (%d)", what, (int)r);
  return 0;
}

// (c, f, t, b) view of a grid whose sequences run along f (axis = ALONG_FREQ: rows = (b,t) pairs, folded into one dimension)
// or along t (ALONG_TIME: rows = f inside one utterance).  `rows` consecutive sequences x `cbox` channels per box.
int grid_map(CUtensorMap* m, const void* base, int c, int ld, int nb, int nt, int nf, int axis, int rows, int cbox,
             CUtensorMapSwizzle swz, CUtensorMapL2promotion l2, const char* what) {
  uint64_t dims[4], str[3];
  uint32_t box[4];
  if (axis == FNSSL_ALONG_FREQ) {
    dims[0] = (uint64_t)c; dims[1] = (uint64_t)nf; dims[2] = (uint64_t)nb * nt; dims[3] = 1;
    str[0] = (uint64_t)ld * 2; str[1] = (uint64_t)nf * ld * 2; str[2] = (uint64_t)nb * nt * nf * ld * 2;
    box[0] = (uint32_t)cbox; box[1] = 1; box[2] = (uint32_t)rows; box[3] = 1;
  } else {
    dims[0] = (uint64_t)c; dims[1] = (uint64_t)nf; dims[2] = (uint64_t)nt; dims[3] = (uint64_t)nb;
    str[0] = (uint64_t)ld * 2; str[1] = (uint64_t)nf * ld * 2; str[2] = (uint64_t)nt * nf * ld * 2;
    box[0] = (uint32_t)cbox; box[1] = (uint32_t)rows; box[2] = 1; box[3] = 1;
  }
  return encode(m, 4, base, dims, str, box, swz, l2, what);
}
}  // namespace

// 4-D fp16 map over an input grid: box = [mr sequences x 64 channels], 128B swizzle, zero fill outside the tensor
int make_grid_map(CUtensorMap* m, const void* base, int c, int ld, int nb, int nt, int nf, int axis, int mr) {
  FNSSL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld % 8) == 0, "lstm(tcgen05): grid base / channel stride not 16-byte aligned");
  return grid_map(m, base, c, ld, nb, nt, nf, axis, mr, kSlabK, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "grid");
}

// narrow (<= 16 channel) second source: box = [mr x 16], 32B swizzle (one K = 16 step)
int make_small_grid_map(CUtensorMap* m, const void* base, int c, int ld, int nb, int nt, int nf, int axis, int mr) {
  FNSSL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld % 8) == 0 && c <= 16,
                "lstm(tcgen05): small-source grid base / channel stride not 16-byte aligned or wider than 16 channels");
  return grid_map(m, base, c, ld, nb, nt, nf, axis, mr, 16, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "small-source");
}

// output grid: box = [rows x 32 channels] in the 64B-swizzled layout of an h exchange tile (TMA stores / reduce-adds)
int make_out_map(CUtensorMap* m, const void* base, int ld, int nb, int nt, int nf, int axis, int rows) {
  FNSSL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld % 8) == 0, "lstm(tcgen05): output grid base / channel stride not 16-byte aligned");
  return grid_map(m, base, ld, ld, nb, nt, nf, axis, rows, 32, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, "output");
}

// 2-D fp16 map over the packed weights [nchunks_total * 128 rows][nslabs * 64], box = one [128 x 64] slab
int make_weight_map(CUtensorMap* m, const void* weights, int nslabs, int nchunks_total) {
  const uint64_t dims[2] = {(uint64_t)nslabs * kSlabK, (uint64_t)nchunks_total * kChunkN};
  const uint64_t str[1] = {(uint64_t)nslabs * kSlabK * 2};
  const uint32_t box[2] = {kSlabK, kChunkN};
  return encode(m, 2, weights, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "weight");
}

// ... box = one [128 x 16] slab (32B swizzle) for the narrow second source
int make_small_weight_map(CUtensorMap* m, const void* weights, int nslabs, int nchunks_total) {
  const uint64_t dims[2] = {(uint64_t)nslabs * kSlabK, (uint64_t)nchunks_total * kChunkN};
  const uint64_t str[1] = {(uint64_t)nslabs * kSlabK * 2};
  const uint32_t box[2] = {16, kChunkN};
  return encode(m, 2, weights, dims, str, box, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "small weight");
}

// One host-mapped int per process, written by mbar_timeout (debug builds of the waits) and readable after a trap.
// Portable + mapped: the same pointer is valid in every device's context (unified addressing); initialised exactly once.
static int* g_flag_host = nullptr;
int* tc_error_flag() {
  static int* flag_dev = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    if (cudaHostAlloc(&g_flag_host, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { g_flag_host = nullptr; return; }
    *g_flag_host = 0;
    if (cudaHostGetDevicePointer(&flag_dev, g_flag_host, 0) != cudaSuccess) flag_dev = nullptr;
  });
  return flag_dev;
}

// Bounded pipeline waits are a debugging aid (a protocol bug traps instead of hanging the GPU) but a trap poisons the whole
// CUDA context, and preemption / MPS time-slicing / a debugger can legitimately stretch a wait: production launches spin
// without a bound unless FNSSL_TC_WAIT_TIMEOUT is set (read once; the GPU test-suite sets it).
bool tc_wait_timeout_enabled() {
  static int v = -1;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* e = getenv("FNSSL_TC_WAIT_TIMEOUT");
    v = (e && atoi(e) != 0) ? 1 : 0;
  });
  return v == 1;
}

// FNSSL_TC_DEBUG selects timing experiments inside the kernels.  Bits outside `safe_mask` make a kernel skip work (WRONG results)
// or drop a hand-shake; they are honoured only together with FNSSL_TC_UNSAFE_EXPERIMENTS=1 (tools/ sets both).
int tc_debug_bits(int safe_mask) {
  const char* e = getenv("FNSSL_TC_DEBUG");
  if (!e) return 0;
  const int v = atoi(e);
  const char* u = getenv("FNSSL_TC_UNSAFE_EXPERIMENTS");
  return (u && atoi(u) != 0) ? v : (v & safe_mask);
}

bool lstm_tc4_supports(int hidden, int c0, int c1);
int lstm_forward_tc4(const fnssl_lstm_args* a, cudaStream_t st);
bool lstm_tc5_wants(const fnssl_lstm_args* a);      // CTA-pair (cta_group::2) kernel: H = 128 layers with enough rows
int lstm_forward_tc5(const fnssl_lstm_args* a, cudaStream_t st);
bool lstm_tc6_wants(const fnssl_lstm_args* a);      // CTA-pair kernel with M = 128 MMAs (64 rows per CTA): H = 256, mid-size H = 128 layers
int lstm_forward_tc6(const fnssl_lstm_args* a, cudaStream_t st);

int lstm_forward_tc(const fnssl_lstm_args* a, cudaStream_t st) {
  FNSSL_REQUIRE(a->dtype == FNSSL_F16, "lstm(tcgen05): grids must be fp16");
  FNSSL_REQUIRE(a->c0 % 16 == 0 && a->c1 % 16 == 0, "lstm(tcgen05): channel counts must be multiples of 16 (got %d, %d); pad the grid",
                a->c0, a->c1);
  FNSSL_REQUIRE(lstm_tc4_supports(a->hidden, a->c0, a->c1),
                "lstm(tcgen05): layer shape H=%d c0=%d c1=%d is not built (H in {64,128,256}, <= 6 input slabs of 64 channels); "
                "use FNSSL_ENGINE_SIMT for it", a->hidden, a->c0, a->c1);
  if (lstm_tc5_wants(a)) return lstm_forward_tc5(a, st);      // H = 128, at least one wave of 512-row cluster tiles
  if (lstm_tc6_wants(a)) return lstm_forward_tc6(a, st);      // H = 256 by wave count; mid-size H = 128 layers
  return lstm_forward_tc4(a, st);
}

}  // namespace fnssl

extern "C" int fnssl_lstm_tc_supported(int hidden, int c0, int c1) {
  if (c0 <= 0 || c0 % 16 || c1 < 0 || c1 % 16) return 0;
  return fnssl::lstm_tc4_supports(hidden, c0, c1) ? 1 : 0;
}

extern "C" int fnssl_lstm_tc_kernel_for(const fnssl_lstm_args* a) {
  if (!a || a->dtype != FNSSL_F16 || a->c0 <= 0 || a->c0 % 16 || a->c1 < 0 || a->c1 % 16) return 0;
  if (!fnssl::lstm_tc4_supports(a->hidden, a->c0, a->c1)) return 0;
  if (fnssl::lstm_tc5_wants(a)) return 5;
  if (fnssl::lstm_tc6_wants(a)) return 6;
  return 4;
}

extern "C" int fnssl_lstm_tc_error_site(void) {
  // site code recorded by a timed-out mbarrier wait (0 = none); resets the flag
  if (!fnssl::g_flag_host) return 0;
  const int v = *reinterpret_cast<volatile int*>(fnssl::g_flag_host);
  *fnssl::g_flag_host = 0;
  return v;
}
