/*
 * fnssl_b200 -- C ABI of the B200-native (sm_100a) FN-SSL / IPDnet / IPDnet2 forward hot path.
 *
 * The reference (Audio-WestlakeU/FN-SSL) has no FFI: its "plugin surface" for this path is the
 * Python nn.Module contract.  This header is the drop-in boundary one level
 * below it: plain device pointers and sizes, no torch types.  Each entry point names the reference
 * call site it replaces.  fn_ssl_b200/{Module,Model,FixedAarryIPDnet,IPDnet2}.py bind these through ctypes
 * and re-create the reference's module classes on top.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated; `stream` is a cudaStream_t passed as void*;
 *   - all functions enqueue work on `stream` and return immediately; 0 = ok, non-zero = error
 *     (message via fnssl_last_error(), thread-local); nothing aborts, nothing falls back to the CPU;
 *   - "grid" tensors are channels-last  (nb, nt, nf, C)  with a channel stride `ld >= C`
 *     (element (b,t,f,c) at ((b*nt + t)*nf + f)*ld + c) in fp32 or fp16 (FNSSL_F32 / FNSSL_F16);
 *   - the reference's own layouts ((nb,C,nf,nt) network input, (nb,nf,nt,nch) complex STFT,
 *     (nb,nt//12,2nf) IPD output ...) are produced / consumed by the entry points below as documented.
 */
#ifndef FNSSL_B200_H_
#define FNSSL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FNSSL_ABI_VERSION 5 /* 2: carried LSTM state; 3: IPDnet2 entry points (fnssl_reflect_pad, fnssl_sn_*); 4: one tcgen05 LSTM kernel
                             (fnssl_lstm_tc_trace of the retired generations removed), training-side targets / losses,
                             fused fnssl_stft_features_forward; 5: training forward / backward of the LSTM layer and the
                             DP-IPD head (fnssl_lstm_forward_train, fnssl_lstm_backward, fnssl_ipd_head_backward), of the causal conv (fnssl_conv3x3_*)
                             and of the DOA linear (fnssl_linear_backward) */

/* element types of grid tensors */
#define FNSSL_F32 0
#define FNSSL_F16 1

/* recurrence axis of an LSTM pass over a (nb, nt, nf, C) grid */
#define FNSSL_ALONG_FREQ 0 /* full-band:   sequences = (b,t) rows, steps = f   (Model.py:35-38)  */
#define FNSSL_ALONG_TIME 1 /* narrow-band: sequences = (b,f) rows, steps = t   (Model.py:41-46)  */

/* LSTM engine */
#define FNSSL_ENGINE_SIMT 0    /* fp32 CUDA-core kernel, fp32 or fp16 grids, any shape            */
#define FNSSL_ENGINE_TCGEN05 1 /* tcgen05/TMEM/TMA kernel, fp16 operands, fp32 accumulate + state */

/* channel pairing of the feature front end */
#define FNSSL_PAIRS_M 0   /* AddChToBatch('M'):  rows (ref 0, m)          Module.py:390-396 */
#define FNSSL_PAIRS_MM 1  /* AddChToBatch('MM'): all i<j pairs            Module.py:398-404 */
#define FNSSL_PAIRS_ALL 2 /* IPDnet: no re-batching, all mics as channels runIPDnetOn.py:246 */

/* magnitude normalisation */
#define FNSSL_NORM_NONE 0
#define FNSSL_NORM_FORGETTING 1 /* utils_.py:9-55  */
#define FNSSL_NORM_GLOBAL 2     /* runIPDnetOff.py:248-251 */
#define FNSSL_NORM_GIVEN 3      /* fnssl_features_forward only: mu is an INPUT (e.g. from fnssl_norm_stream_forward) */

int fnssl_abi_version(void);
const char* fnssl_last_error(void);

/* ---- front end ------------------------------------------------------------------------------ */

/* nt = floor((nsample - win_len) / hop + 1), FN-SSL/Module.py:56 */
int fnssl_stft_num_frames(int nsample, int win_len, int hop);

/* Replaces STFT.forward (FN-SSL/Lightning/Module.py:48-68; IPDnet/Module.py:45-63): framed,
 * periodic-Hann-windowed, un-normalised one-sided 512-point FFT, center=False.
 *   signal : (nb, nsample, nch) f32        spec : (nb, 257, nt, nch) complex64 (interleaved re,im)
 *   magsum : optional (nb, nch, nt) f32 = sum over the 257 bins of |X| (feeds the normaliser), or NULL
 * Only win_len = nfft = 512 is implemented (the value hard-coded by every caller, main.py:38-44). */
int fnssl_stft_forward(const float* signal, int nb, int nsample, int nch, int win_len, int hop, int nfft,
                       float* spec, float* magsum, void* stream);

/* The normaliser alone (forgetting_norm, FN-SSL/Lightning/utils_.py:9-55; global mean, runIPDnetOff.py:248-251).
 *   magsum : (nb, nch, nt) f32 sums of |X| over `nbins` bins;  mu : (R, nt) f32, R = fnssl_feature_rows() */
int fnssl_norm_forward(const float* magsum, int nb, int nch, int nt, int nbins, int pairing, int norm,
                       int sample_length, float* mu, void* stream);

/* forgetting_norm continued across chunks of one stream (chunked inference; the reference runs whole clips):
 * frames [t0, t0+nt) of the recursion utils_.py:27-44.  mu_state : (R) f32, read as mu_{t0-1} when t0 > 0 and
 * always written with mu_{t0+nt-1}; mu : (R, nt) f32.  Feed mu to fnssl_features_forward(norm = FNSSL_NORM_GIVEN). */
int fnssl_norm_stream_forward(const float* magsum, int nb, int nch, int nt, int nbins, int pairing, int sample_length,
                              long long t0, float* mu_state, float* mu, void* stream);

/* rows of the feature tensor for a pairing mode: nb*(nch-1), nb*nch*(nch-1)/2 or nb */
int fnssl_feature_rows(int nb, int nch, int pairing);
/* channels of the feature tensor: 4 for pair modes, 2*nch for FNSSL_PAIRS_ALL */
int fnssl_feature_channels(int nch, int pairing);

/* Replaces the body of data_preprocess (FN-SSL/Lightning/main.py:206-225; IPDnet/runIPDnetOn.py:240-254;
 * runIPDnetOff.py:248-251): pair re-batching + |X| + forgetting_norm / global mean + re,im / (mu+eps) +
 * cat + bins 1..256.
 *   spec, magsum : outputs of fnssl_stft_forward
 *   mu    : workspace AND output, (R, nt) f32 -- the normaliser the reference returns from forgetting_norm
 *   feat  : grid (R, nt, 256, ld) of `dtype`; channels [re_0..re_{C/2-1}, im_0..im_{C/2-1}], channels
 *           C..ld-1 are written as zero (padding for the tensor-core path)
 *   feat_cfirst : optional (R, C, 256, nt) f32 in the reference's own layout, or NULL */
int fnssl_features_forward(const float* spec, const float* magsum, int nb, int nt, int nch, int pairing,
                           int norm, int sample_length, float eps, float* mu, void* feat, int dtype, int ld,
                           float* feat_cfirst, void* stream);

/* FUSED front end for the pipelines: signal -> normalised feature grid, without ever writing the complex spectrum to HBM
 * (STFT.forward + data_preprocess in one entry point: FN-SSL/Lightning/Module.py:48-68 + main.py:206-225; IPDnet/
 * runIPDnetOn.py:240-254).  Two passes over the (L2-resident) signal: pass 1 = FFT -> per-frame sums of |X| only -> the
 * normaliser recursion; pass 2 = FFT again -> re/(mu+eps), im/(mu+eps), bins 1..256, written straight into the grid.
 *   magsum : workspace (nb, nch, nt) f32 (unused for FNSSL_NORM_NONE / _GIVEN);  mu : (R, nt) f32, output (input for _GIVEN)
 *   feat   : 16-byte aligned grid (R, nt, 256, ld) of `dtype`, ld a multiple of 16 bytes; channels C..ld-1 are written as zero
 * Same numbers as fnssl_stft_forward + fnssl_features_forward (the same FFT code runs twice). */
int fnssl_stft_features_forward(const float* signal, int nb, int nsample, int nch, int win_len, int hop, int nfft, int pairing,
                                int norm, int sample_length, float eps, float* magsum, float* mu, void* feat, int dtype, int ld,
                                void* stream);

/* (nb, C, nf, nt) f32  <->  grid (nb, nt, nf, ld) of dtype at channel offset `ch_off`
 * (FN_SSL.forward's x.permute(0,3,2,1), Model.py:73; IPDnet :93,:111). */
int fnssl_cfirst_to_grid(const float* src, int nb, int C, int nf, int nt, void* dst, int dtype, int ld,
                         int ch_off, void* stream);
int fnssl_grid_to_cfirst(const void* src, int dtype, int ld, int ch_off, int nb, int C, int nf, int nt,
                         float* dst, void* stream);
/* grid copy / convert / zero-pad: dst[..., dst_off + c] = src[..., src_off + c] for c < C */
int fnssl_grid_copy(const void* src, int src_dtype, int src_ld, int src_off, void* dst, int dst_dtype,
                    int dst_ld, int dst_off, int64_t npos, int C, void* stream);

/* dst = a + b, elementwise over n elements of `dtype` (FNblock's standalone residual add, Model.py:36-37,44-45) */
int fnssl_grid_add(const void* a, const void* b, void* dst, int dtype, int64_t n, void* stream);

/* ---- LSTM ----------------------------------------------------------------------------------- */

/* One uni- or bi-directional LSTM layer over a grid; replaces nn.LSTM as called at
 * FN-SSL/Lightning/Model.py:38,46 and IPDnet/FixedAarryIPDnet.py:32,36 *including* the surrounding
 * layout glue (reshape/permute :35,:41,:49), the channel concat of a skip (:42-43; IPDnet :34,:38) and
 * the residual add feeding the NEXT layer (:36-37,:44-45).
 *
 *   input   x_t  = concat(src0[c0 channels], src1[c1 channels])         (src1 may be NULL, c1 = 0)
 *   output  out0 = h                  (dirs*hidden channels at channel offset out0_off, stride out0_ld)
 *           out1 = h + addend         (optional; same channel count, its own stride)  -- the operand of
 *                                      the next layer in FN-SSL's additive-skip blocks.  With addend == NULL
 *                                      out1 is a second copy of h (block 1's narrow-band layer reads h as its
 *                                      input AND accumulates its residual sum onto a copy of it, Model.py:41-45)
 *   weights: engine-specific packed buffer produced by fn_ssl_b200.packing (layout in DESIGN.md) from
 *            nn.LSTM-shaped weight_ih_l0 (4H,in), weight_hh_l0 (4H,H), bias_ih_l0, bias_hh_l0 [+ _reverse];
 *            gate order i,f,g,o; zero initial state.
 */
typedef struct fnssl_lstm_args {
  int32_t engine;   /* FNSSL_ENGINE_* */
  int32_t axis;     /* FNSSL_ALONG_*  */
  int32_t nb, nt, nf;
  int32_t hidden;   /* H per direction */
  int32_t num_dirs; /* 1 or 2; direction 1 runs the sequence in reverse */
  int32_t dtype;    /* element type of every grid in this call */
  const void* src0; int32_t c0; int32_t ld0;
  const void* src1; int32_t c1; int32_t ld1;
  const void* weights; int64_t weights_bytes;
  void* out0; int32_t out0_ld; int32_t out0_off;
  const void* addend; int32_t addend_ld;
  void* out1; int32_t out1_ld;
  /* Optional recurrent state for chunked (streaming) inference -- nn.LSTM's (h_0, c_0) argument / (h_n, c_n)
   * result, which the reference leaves at zero / discards (Model.py:38,46) because it runs whole clips.
   * fp32 (rows, hidden), rows = nb*nf for FNSSL_ALONG_TIME, nb*nt for FNSSL_ALONG_FREQ; num_dirs == 1 only.
   * state_flags bit 0: start from h_state/c_state instead of zeros; bit 1: write the state after the last
   * step back into the same buffers.  0 / NULL = the reference's behaviour. */
  float* h_state; float* c_state; int32_t state_flags;
} fnssl_lstm_args;

int fnssl_lstm_forward(const fnssl_lstm_args* args, void* stream);

/* 1 if the tcgen05 engine is built for this layer shape (fp16 grids; hidden, c0, c1 as in fnssl_lstm_args) */
int fnssl_lstm_tc_supported(int hidden, int c0, int c1);
/* Which tensor-core kernel fnssl_lstm_forward would launch for these arguments (no launch, no GPU needed): 4 = lstm_tc4.cu
 * (cluster kernel: small grids, carried state), 5 = lstm_tc5.cu (CTA pairs, H = 128, >= 30 clusters), 6 = lstm_tc6.cu (CTA pairs
 * with M = 128: H = 256 by wave count, mid-size two-source H = 128 layers); 0 if the engine is not built for the shape (the
 * fp32 kernel runs).  Honours FNSSL_TC_PAIR / FNSSL_TC_PAIR256 / *_MIN like the dispatcher itself. */
int fnssl_lstm_tc_kernel_for(const fnssl_lstm_args* args);
/* diagnostic: site code written by a timed-out pipeline wait inside the tcgen05 kernel (0 = none) */
int fnssl_lstm_tc_error_site(void);
/* diagnostic: last in-kernel timeline of the cluster kernel (lstm_tc4.cu; 16 slots x 16 SM-clock stamps) recorded when
 * FNSSL_TC_TRACE is set; 0 if none */
int fnssl_lstm_tc4_trace(long long* out256);

/* ---- LSTM, training side ---------------------------------------------------------------------- */

/* What autograd does behind nn.LSTM in the reference's training_step (FN-SSL/Lightning/main.py:95-109 through Model.py:38,46;
 * IPDnet/runIPDnetOn.py:110-125 through FixedAarryIPDnet.py:32,36), as explicit entry points on the fp32 engine
 * (args->engine == FNSSL_ENGINE_SIMT, args->dtype == FNSSL_F32, no carried state; addend / out1 are honoured by the forward
 * but have no gradient path here -- the host adds the residuals, fn_ssl_b200/training.py).
 *
 * fnssl_lstm_forward_train: fnssl_lstm_forward + `saved`: per (direction, grid position, unit) the activated gates (i,f,g,o)
 *   and the cell state c_t -- fnssl_lstm_train_saved_bytes() bytes, 16-byte aligned.
 * fnssl_lstm_backward: args = the forward's arguments (src0 / src1 / weights as then, out0 = the h grid it produced);
 *   saved   : the forward's buffer; CONSUMED (the gates are overwritten with the gate pre-activation gradients)
 *   whh_t   : weight_hh transposed for the recurrent product: [dirs][H (unit j)][H (input k)][4 (gate)] f32
 *             = weight_hh_l0[gate*H + j][k]   (fn_ssl_b200.packing.pack_lstm_whh_t)
 *   dout    : gradient w.r.t. h, grid (nb, nt, nf, dout_ld) f32, dirs*H channels
 *   dsrc0/1 : OUT gradient w.r.t. the two input sources, grids with c0 / c1 channels (either may be NULL: not needed)
 *   dweights: OUT gradient in the layout of the packed forward weights (args->weights_bytes bytes: [dirs][Kp][H][4] for
 *             weight_ih | weight_hh rows, then [dirs][H][4] for the bias = d bias_ih = d bias_hh); zeroed by the call. */
int64_t fnssl_lstm_train_saved_bytes(int nb, int nt, int nf, int hidden, int num_dirs);
int fnssl_lstm_forward_train(const fnssl_lstm_args* args, void* saved, int64_t saved_bytes, void* stream);
int fnssl_lstm_backward(const fnssl_lstm_args* args, void* saved, int64_t saved_bytes, const float* whh_t, const float* dout,
                        int dout_ld, float* dsrc0, int dsrc0_ld, float* dsrc1, int dsrc1_ld, float* dweights, void* stream);

/* ---- heads ---------------------------------------------------------------------------------- */

/* FN_SSL head (Model.py:79-87): AvgPool over 12 frames -> Linear(C,2) -> tanh -> [ch0 over f | ch1 over f].
 *   x : grid (nb, nt, nf, ld) of dtype, C channels used;  w : (2, C) f32, b : (2) f32
 *   out : (nb, nt/12, 2*nf) f32 */
int fnssl_ipd_head_forward(const void* x, int dtype, int ld, int nb, int nt, int nf, int C, const float* w,
                           const float* b, float* out, void* stream);
/* Backward of fnssl_ipd_head_forward on an fp32 grid (training side): y = the forward's output, dy its gradient, both
 * (nb, nt/12, 2*nf);  dx : grid (nb, nt, nf, dld) -- frames [12*(nt/12), nt) are not written;  dw : (2, C), db : (2), zeroed
 * by the call.  C <= 512. */
int fnssl_ipd_head_backward(const float* x, int ld, int nb, int nt, int nf, int C, const float* w, const float* y, const float* dy,
                            float* dx, int dld, float* dw, float* db, void* stream);

/* y = x @ w^T + b for small row counts; the DOA classifier Linear(512,180) (Model.py:71,88-89).
 *   x : (rows, in) f32, w : (out, in) f32, b : (out) f32, y : (rows, out) f32 */
int fnssl_linear_forward(const float* x, const float* w, const float* b, int rows, int in_features,
                         int out_features, float* y, void* stream);
/* Backward of fnssl_linear_forward (training side, the DOA classifier): dx (rows, in) [may be NULL], dw (out, in), db (out). */
int fnssl_linear_backward(const float* x, const float* w, const float* dy, int rows, int in_features, int out_features, float* dx,
                          float* dw, float* db, void* stream);

/* CausCnnBlock.forward (IPDnet/FixedAarryIPDnet.py:61-73): 3x (Conv2d 3x3, pad (1,2), no bias, crop 2)
 * with ReLU+AvgPool(1,3), ReLU+AvgPool(1,4), tanh.
 *   input  = concat(src0[c0], src1[c1]) grids (nb, nt, nf, ld*) of dtype
 *   w1 : (hid, c0+c1, 3, 3), w2 : (hid, hid, 3, 3), w3 : (cout, hid, 3, 3)  f32, PyTorch layout
 *   work : scratch, fnssl_causcnn_workspace_bytes() bytes
 *   out : (nb, cout, nf, nt/12) f32 -- the reference's layout */
size_t fnssl_causcnn_workspace_bytes(int nb, int nt, int nf, int cin, int hid, int cout);
int fnssl_causcnn_forward(const void* src0, int c0, int ld0, const void* src1, int c1, int ld1, int dtype,
                          int nb, int nt, int nf, const float* w1, const float* w2, const float* w3, int hid,
                          int cout, void* work, float* out, void* stream);

/* Training side of CausCnnBlock: the three products autograd runs behind each nn.Conv2d(kernel 3x3, padding (1,2), bias=False)
 * + crop of the last two frames (IPDnet/FixedAarryIPDnet.py:50-52,61-72; training_step IPDnet/runIPDnetOn.py:110-125), on fp32
 * channels-last grids (nb, nt, nf, C); ReLU / AvgPool / tanh between them are host-side elementwise ops
 * (fn_ssl_b200.training.causcnn_train).  w, dw : (cout, c0+c1, 3, 3) f32 in PyTorch's layout.
 *   forward         out[b,t,f,o] = sum w[o,c,kf,kt] in[b, t+kt-2, f+kf-1, c],   in = concat(in0[c0], in1[c1])
 *   backward_data   din[b,t,f,c] = sum w[o,c,kf,kt] dy[b, t-kt+2, f-kf+1, o]    (din0: c < c0, din1: the rest; either may be NULL)
 *   backward_weight dw[o,c,kf,kt] = sum_{b,t,f} dy[b,t,f,o] in[b, t+kt-2, f+kf-1, c]
 * work : scratch of fnssl_conv3x3_train_workspace_bytes(c0+c1, cout) bytes (a re-packed copy of w / the unreduced dw). */
size_t fnssl_conv3x3_train_workspace_bytes(int cin, int cout);
int fnssl_conv3x3_forward(const float* in0, int c0, int ld0, const float* in1, int c1, int ld1, int nb, int nt, int nf, const float* w,
                          int cout, float* work, float* out, int out_ld, void* stream);
int fnssl_conv3x3_backward_data(const float* dy, int cout, int dy_ld, int nb, int nt, int nf, const float* w, int c0, int c1, float* work,
                                float* din0, int din0_ld, float* din1, int din1_ld, void* stream);
int fnssl_conv3x3_backward_weight(const float* in0, int c0, int ld0, const float* in1, int c1, int ld1, const float* dy, int cout, int dy_ld,
                                  int nb, int nt, int nf, float* work, float* dw, void* stream);

/* ---- IPD -> DOA decoding ("next" row of the scope contract) ------------------------------------- */

/* SourceDetectLocalize.forward, meth_mode 'IDL' (FN-SSL/Lightning/Module.py:525-581): max_sources rounds of
 * spatial spectrum (pred_ipd . templates / (K/2)) -> first-maximum argmax -> projection ratio -> residual update.
 *   pred_ipd : (R, K) f32, R = nb*nt, K = 2nf*npairs in (2nf, pair) order       templ : (ncand, K), templ_t : (K, ncand)
 *   cur, map : workspaces (R, K) and (R, ncand)      ss : (R, ncand) spectrum of the first round
 *   idx_out : (R, max_sources) int32 candidate index (ele*nazi + azi)            vad_out : (R, max_sources)
 *   vad_mode : 0 -> 0, 1 -> 1 ('kNum'), 2 -> projection ratio ('unkNum') */
int fnssl_doa_decode_idl(const float* pred_ipd, const float* templ, const float* templ_t, int R, int K, int ncand,
                         int max_sources, int vad_mode, float* cur, float* map, float* ss, int* idx_out,
                         float* vad_out, void* stream);

/* ---- training-side forward pieces --------------------------------------------- */

/* DP-IPD regression targets: DPIPD.forward(source_doa) + the ground-truth branch of data_preprocess
 * (FN-SSL/Lightning/Module.py:464-497, main.py:227-265; IPDnet/runIPDnetOn.py:256-290).
 *   source_doa : (nb, nt, 2, ns) f32 [elevation, azimuth] radians;  vad : (nb, nt, ns) f32 or NULL (no gating)
 *   mic_pos : (nmic, 3) f32;  pairs : (P, 2) int32 microphone pairs (m1, m2) in DPIPD.data_adjust order
 *   phase(b,t,s,p,k) = 2 pi f_k r(b,t,s).(mic[m1] - mic[m2]) / speed,  f_k = fre_max * (bin_lo + k) / (nf - 1), computed in double
 *   per_source = 0 : out (nb, nt, 2*nbins, P)     = sum_s gate_s [cos | sin]      gate_s = vad > vad_threshold (1 if vad NULL)
 *   per_source = 1 : out (nb, nt, 2*nbins, P, ns) = gate_s [cos | sin]; where gate_s = 0 and nonsrc != NULL the (2*nbins, P)
 *                    non-source target is written instead (IPDnet's silent-source target) */
int fnssl_dpipd_targets(const float* source_doa, const float* vad, const float* mic_pos, const int* pairs, int nb, int nt,
                        int ns, int nmic, int P, int nf, float fre_max, float speed, int bin_lo, int nbins,
                        float vad_threshold, int per_source, const float* nonsrc, float* out, void* stream);

/* cal_loss of FN-SSL (main.py:191-198): pred (nb*P, nt, nf2) [row b*P + p], gt (nb, nt, nf2, P) -> mean squared error.
 *   workspace : nb*nt floats;  loss : 1 float (device).  Deterministic (fixed-order two-stage sum). */
int fnssl_ipd_mse_loss(const float* pred, const float* gt, int nb, int P, int nt, int nf2, float* workspace, float* loss,
                       void* stream);

/* Frame-level PIT loss of IPDnet (runIPDnetOn.py:188-206): pred, gt (rows, K, ns), rows = nb*nt frames, ns <= 4 sources.
 * Per frame the permutation of the predicted sources with the smallest squared error (exhaustive, itertools order, first
 * minimum) is applied; loss = mean over rows*K*ns.  best_perm (rows, ns) int32 or NULL: prediction index per target source.
 *   workspace : rows floats */
int fnssl_ipd_pit_mse_loss(const float* pred, const float* gt, int rows, int K, int ns, float* workspace, float* loss,
                           int* best_perm, void* stream);

/* ---- IPDnet2: OnlineSpatialNet with Mamba time modules (scope row a11) --------------------------- */

/* torch.stft(center=True)'s reflect padding (IPDnet2/Module.py:61-62): (nb, nsample, nch) -> (nb, nsample + 2 pad, nch).
 * Followed by fnssl_stft_forward(hop = 320) this is IPDnet2's STFT: nt = floor(nsample / hop) + 1. */
int fnssl_reflect_pad(const float* signal, int nb, int nsample, int nch, int pad, float* out, void* stream);

/* LayerNorm + grouped Conv1d along frequency + PReLU (SpatialNetLayer.fconv1 / fconv2, IPDnet2.py:105-109,119-123). */
typedef struct fnssl_sn_fconv_weights {
  const float* ln_w;     /* (H)                                                                   */
  const float* ln_b;     /* (H)                                                                   */
  const float* conv_wp;  /* packed [group][k][in-channel of group][out-channel of group]         */
  const float* conv_b;   /* (H)                                                                   */
  const float* prelu;    /* (H)                                                                   */
} fnssl_sn_fconv_weights;

/* Frequency-axis half of one SpatialNetLayer (IPDnet2.py:146-153) [+ the CausalConv1d encoder (:335) when is_first]:
 *   x = x + fconv1(x); [pool F/2]; x = x + unsqueeze(full(squeeze(LN(x)))); x = x + fconv2(x); [pool F/8]
 * is_first: x is the feature grid (nb, nt, 256, x_ld) f32 with cin real channels (zero padded to x_ld, x_ld % 4 == 0);
 *           out is (nb, nt, 16, H).   else: x and out are (nb, nt, nf = 16, H). */
typedef struct fnssl_sn_freq_args {
  int32_t nb, nt, nf;
  int32_t hidden, squeeze, groups, fkernel;
  int32_t is_first;
  const float* x; int32_t cin; int32_t x_ld;
  const float* enc_wp;   /* packed [k][c < x_ld][H] (zero rows for c >= cin)                      */
  const float* enc_b;    /* (H)                                                                   */
  int32_t enc_kernel;
  fnssl_sn_fconv_weights fconv1;
  const float* lnf_w; const float* lnf_b;    /* norm_full                                          */
  const float* sq_wt;    /* squeeze weight TRANSPOSED (H, squeeze)                                */
  const float* sq_b;
  const float* full_wt;  /* full.weight TRANSPOSED (in f', out f)                                 */
  const float* full_b;
  const float* usq_w;    /* unsqueeze weight (H, squeeze)                                         */
  const float* usq_b;
  fnssl_sn_fconv_weights fconv2;
  float* out;
  int32_t t_begin;       /* is_first only: frames [0, t_begin) of x are history for the causal encoder (streaming); outputs
                            are produced for frames [t_begin, nt): out is (nb, nt - t_begin, 16, H).  0 = whole clip.     */
} fnssl_sn_freq_args;
int fnssl_sn_freq_forward(const fnssl_sn_freq_args* args, void* stream);

/* mamba_ssm.Mamba parameters (names as in IPDnet2/checkpoints/ipdnet2_small.ckpt) + the LayerNorm in front of it. */
typedef struct fnssl_mamba_weights {
  const float* ln_w; const float* ln_b;
  const float* in_proj_wt;   /* in_proj.weight TRANSPOSED (d_model, 2 d_inner)                     */
  const float* conv_w;       /* conv1d.weight (d_inner, d_conv)                                    */
  const float* conv_b;
  const float* x_proj_wt;    /* x_proj.weight TRANSPOSED (d_inner, dt_rank + 2 d_state)            */
  const float* dt_proj_w;    /* (d_inner, dt_rank)                                                 */
  const float* dt_proj_b;
  const float* A_log;        /* (d_inner, d_state)                                                 */
  const float* D;
  const float* out_proj_wt;  /* out_proj.weight TRANSPOSED (d_inner, d_model)                      */
} fnssl_mamba_weights;

/* Time-axis half of one SpatialNetLayer (IPDnet2.py:155-163,166-181) [+ AvgPool over `pool` frames, :347]:
 *   x = x + Mamba_0(LN_0(x)); x = x + Mamba_1(LN_1(x));   x: (nb, nt, nf, H) -> out: (nb, nt / pool, nf, H)
 * Two launches (one per Mamba block); work: scratch of the size of x holding the intermediate. */
typedef struct fnssl_sn_time_args {
  int32_t nb, nt, nf, hidden;
  int32_t d_inner, d_state, dt_rank, d_conv;
  int32_t pool;
  const float* x;
  float* work;
  float* out;
  fnssl_mamba_weights m[2];
  /* Optional carried state for chunked (streaming) inference, one buffer per Mamba block: (nb * nf, d_state + d_conv - 1,
   * d_inner) f32 = the selective-scan state and the last d_conv - 1 raw inner-channel frames of the causal conv (what
   * mamba_ssm keeps in InferenceParams, IPDnet2.py:170-177).  state_flags bit 0: resume from it; bit 1: write it back. */
  float* state[2];
  int32_t state_flags;
} fnssl_sn_time_args;
int fnssl_sn_time_forward(const fnssl_sn_time_args* args, void* stream);

/* FreqInverse (IPDnet2.py:37-43) + decoder Linear (:361) + output reshape (:363-364, literal 2 generalised to n_src):
 *   x (nb, nt, nfc, H) -> out (nb, nt, 2 * nfc * ratio, dim_out / (2 n_src), n_src)
 *   trans_wt : trans2.weight TRANSPOSED (H, ratio * dim_out), channel = o * ratio + j */
int fnssl_sn_head_forward(const float* x, int nb, int nt, int nfc, int hidden, const float* trans_wt, const float* trans_b,
                          const float* dec_w, const float* dec_b, int dim_out, int ratio, int n_src, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FNSSL_B200_H_ */
