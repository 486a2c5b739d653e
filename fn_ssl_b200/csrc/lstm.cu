// fnssl_lstm_forward: argument validation and engine dispatch (see include/fnssl_b200.h).
#include "common.cuh"

namespace fnssl {
int lstm_forward_simt(const fnssl_lstm_args* a, cudaStream_t st);
int lstm_forward_tc(const fnssl_lstm_args* a, cudaStream_t st);
}  // namespace fnssl

using namespace fnssl;

extern "C" int fnssl_lstm_forward(const fnssl_lstm_args* a, void* stream) {
  FNSSL_REQUIRE(a != nullptr, "lstm: null args");
  FNSSL_REQUIRE(a->axis == FNSSL_ALONG_FREQ || a->axis == FNSSL_ALONG_TIME, "lstm: bad axis %d", a->axis);
  FNSSL_REQUIRE(a->nb > 0 && a->nt > 0 && a->nf > 0, "lstm: bad grid %d x %d x %d", a->nb, a->nt, a->nf);
  FNSSL_REQUIRE(a->num_dirs == 1 || a->num_dirs == 2, "lstm: num_dirs must be 1 or 2 (got %d)", a->num_dirs);
  FNSSL_REQUIRE(a->dtype == FNSSL_F32 || a->dtype == FNSSL_F16, "lstm: bad dtype %d", a->dtype);
  FNSSL_REQUIRE(a->src0 && a->c0 > 0 && a->ld0 >= a->c0, "lstm: bad src0 (c0=%d ld0=%d)", a->c0, a->ld0);
  FNSSL_REQUIRE(a->c1 >= 0 && (a->c1 == 0 || (a->src1 && a->ld1 >= a->c1)), "lstm: bad src1 (c1=%d ld1=%d)", a->c1, a->ld1);
  FNSSL_REQUIRE(a->weights && (a->out0 || a->out1), "lstm: null weights / no output");
  const int oc = a->num_dirs * a->hidden;
  FNSSL_REQUIRE(!a->out0 || (a->out0_off >= 0 && a->out0_ld >= a->out0_off + oc), "lstm: out0 window (off %d + %d channels) exceeds ld %d",
                a->out0_off, oc, a->out0_ld);
  FNSSL_REQUIRE(!a->out1 || (a->out1_ld >= oc && (!a->addend || a->addend_ld >= oc)), "lstm: bad out1/addend");      // (addend == NULL: out1 = h)
  if (a->state_flags) {
    FNSSL_REQUIRE((a->state_flags & ~3) == 0, "lstm: bad state_flags %d", a->state_flags);
    FNSSL_REQUIRE(a->num_dirs == 1, "lstm: recurrent state is carried for uni-directional layers only");
    FNSSL_REQUIRE(a->h_state && a->c_state, "lstm: state_flags set but h_state/c_state is null");
  }
  if (a->engine == FNSSL_ENGINE_SIMT) return lstm_forward_simt(a, (cudaStream_t)stream);
  if (a->engine == FNSSL_ENGINE_TCGEN05) return lstm_forward_tc(a, (cudaStream_t)stream);
  FNSSL_FAIL("lstm: unknown engine %d", a->engine);
}
