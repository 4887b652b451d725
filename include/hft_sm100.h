/* libhft_sm100.so -- C ABI of the B200 (sm_100a) implementation of nylon-amt's hFT-Transformer hot path.
 *
 * The reference (d-f/nylon-amt) is 100 % Python and has no FFI layer; its "operator API" for this path is two
 * Python call signatures.  Each entry point below names the reference interface it replaces; the host-side mirror
 * of those signatures lives in nylon_amt_b200/{amt.py,model_spec2midi.py} and INTEGRATION.md shows the binding a
 * reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a cudaError_t or an HFT_ERR_* code otherwise; hft_last_error() gives
 *     the message of the last failure on the calling thread.  Nothing throws across the ABI.
 *   - POD arguments only.  `*_dev` pointers are CUDA device pointers owned by the caller (PyTorch in the mirror);
 *     `*_host` pointers are host memory.  `stream` is a cudaStream_t passed as void* (NULL = legacy default
 *     stream).  Launches are asynchronous on that stream unless stated otherwise.
 *   - no CPU fallback and no silent shape fallback: an unsupported configuration is an error.
 *   - handles are not thread-safe; use one handle per device per thread (the reference is single-device too:
 *     hftt_code/training/m_training.py:113, hftt_code/model/amt.py:14-17).
 */
#ifndef HFT_SM100_H
#define HFT_SM100_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HFT_ERR_ARG 10001
#define HFT_ERR_UNSUPPORTED 10002
#define HFT_ERR_STATE 10003

/* ABI version of this header (bumped on incompatible change). */
int hft_version(void);
/* Message of the last error raised on this thread ("" if none). */
const char* hft_last_error(void);
/* Device facts the host side uses for sharding and bench bookkeeping. */
int hft_device_sm_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Log-mel front end.
 * Replaces: torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, win_length=2048, hop_length=256,
 *           pad_mode='constant', n_mels=256, norm='slaney') followed by torch.log(mel + log_offset).T in
 *           AMT.wav2feature  -- reference hftt_code/model/amt.py:59-61 (parameters hftt_code/corpus/config.json:2-12;
 *           second call site dataset_creation.py:45-53,24).
 * Fixed geometry: n_fft 2048, hop 256, 1025 FFT bins, 256 mel bins.  T = 1 + n_samples / 256 frames.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct hft_logmel_plan hft_logmel_plan;

/* window_host: 2048 floats (torch.hann_window(2048)) or NULL for the built-in periodic Hann.
 * fb_host: dense filterbank [1025][256] row-major exactly as torchaudio's MelScale.fb, or NULL for the built-in
 *          HTK/slaney 0..8 kHz table.  Each mel column must be a band of <= 31 consecutive FFT bins. */
int hft_logmel_create(hft_logmel_plan** plan, const float* window_host, const float* fb_host, float log_offset);
int hft_logmel_destroy(hft_logmel_plan* plan);
int64_t hft_logmel_num_frames(int64_t n_samples);

/* One clip resident in HBM: wav_dev [n_samples] fp32 mono 16 kHz -> out_dev [n_frames][256] fp32 (amt.py:61 layout). */
int hft_logmel_f32(hft_logmel_plan* plan, const float* wav_dev, int64_t n_samples, float* out_dev, int64_t n_frames, void* stream);

/* A ragged batch of clips in ONE launch (the file loop of hftt_code/corpus/conv_wav2fe.py:41-48): clip c is
 * wav_dev[clip_start_host[c] .. clip_start_host[c] + clip_len_host[c]); its T_c rows follow those of clip c-1 in
 * out_dev.  Clips whose start is a multiple of 4 samples (16 bytes) are staged by TMA. */
int hft_logmel_batch_f32(hft_logmel_plan* plan, const float* wav_dev, const int64_t* clip_start_host, const int64_t* clip_len_host,
                         int n_clips, float* out_dev, void* stream);

/* Host-buffer form of AMT.wav2feature's arithmetic (amt.py:59-63 returns a CPU tensor): H2D copy, kernel, D2H copy,
 * stream synchronised on return. */
int hft_logmel_host_f32(hft_logmel_plan* plan, const float* wav_host, int64_t n_samples, float* out_host, int64_t n_frames,
                        void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Channel mix-down + resampling in front of the log-mel kernel.
 * Replaces: wave_mono = torch.mean(wave, dim=0); torchaudio.transforms.Resample(sr, 16000)(wave_mono)
 *           -- reference hftt_code/model/amt.py:56-58 (same call in dataset_creation.py:21-24).
 * kernel_host is the polyphase table [new_reduced][2*width + orig_reduced] fp32 that torchaudio builds
 * (torchaudio.functional.functional._get_sinc_resample_kernel: sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99),
 * with orig_reduced = sr / gcd, new_reduced = 16000 / gcd.  Output length = ceil(new * n_in / orig).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct hft_resample_plan hft_resample_plan;
int hft_resample_create(hft_resample_plan** plan, const float* kernel_host, int32_t orig_reduced, int32_t new_reduced, int32_t width);
/* The same table built by the library (double-precision restatement of torchaudio's formula, fp32 result within 1e-7 of torchaudio's):
 * hft_resample_build_table is host-only (no device needed) and returns the number of floats the table takes ([new_reduced][2*width+orig_reduced]);
 * it fills table_host when capacity suffices.  hft_resample_create_hz = build + hft_resample_create. */
int64_t hft_resample_build_table(int32_t orig_hz, int32_t new_hz, float* table_host, int64_t capacity, int32_t* orig_reduced, int32_t* new_reduced, int32_t* width);
int hft_resample_create_hz(hft_resample_plan** plan, int32_t orig_hz, int32_t new_hz);
int hft_resample_destroy(hft_resample_plan* plan);
int64_t hft_resample_num_samples(const hft_resample_plan* plan, int64_t n_in);
/* wav_dev [channels][n_in] fp32 (channel-major, like torchaudio.load) -> out_dev [n_out] fp32 mono at the new rate. */
int hft_resample_mono_f32(hft_resample_plan* plan, const float* wav_dev, int32_t channels, int64_t n_in, float* out_dev, int64_t n_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * hFT-Transformer forward.
 * Replaces: Model_SPEC2MIDI.forward(input_spec) -- reference hftt_code/model/model_spec2midi.py:15-35
 *           (Encoder_SPEC2MIDI.forward :60-106, Decoder_SPEC2MIDI.forward :145-216, EncoderLayer :230-245,
 *           DecoderLayer_Zero :255-272, DecoderLayer :283-306, MultiHeadAttentionLayer :322-360,
 *           PositionwiseFeedforwardLayer :369-378).  Eval semantics (dropout = identity).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct hft_model hft_model;

/* Constructor arguments of Encoder_SPEC2MIDI / Decoder_SPEC2MIDI (model_spec2midi.py:42,113). */
typedef struct hft_dims {
  int32_t n_margin;      /* 32  */
  int32_t n_frame;       /* 128 */
  int32_t n_bin;         /* 256 */
  int32_t cnn_channel;   /* 4   */
  int32_t cnn_kernel;    /* 5   */
  int32_t hid_dim;       /* 256 paper / 64 reduced */
  int32_t pf_dim;        /* 512 / 128 */
  int32_t n_enc_layers;  /* 3 / 2 */
  int32_t n_dec_layers;  /* 3 / 2 */
  int32_t n_heads;       /* 4 / 2 (encoder and decoder use the same count) */
  int32_t n_note;        /* 88  */
  int32_t n_velocity;    /* 128 */
} hft_dims;

/* Arithmetic the forward runs in. */
#define HFT_PREC_F32 0      /* fp32 CUDA-core kernels: the 2e-3 parity path */
#define HFT_PREC_BF16 1     /* bf16 operands on tcgen05 tensor cores, fp32 accumulate, fp32 softmax/LayerNorm/sigmoid */
#define HFT_PREC_F16 2      /* fp16 operands on tcgen05 tensor cores (3 more mantissa bits than bf16) */
#define HFT_PREC_F16X3 3    /* split fp16 operands (hi + lo): a*w = ah*wh + al*wh + ah*wl on tcgen05, ~22 mantissa bits:
                               the tensor-core path that meets the fp32 parity budget (2e-3) */
#define HFT_PREC_MIXED 4    /* the 16-bit-class mode: split-fp16 storage with a per-GEMM plan of ONE or THREE products (single: every P V
                               product outside encoder layer 0, the decoder's score products, time layers >= 1, both head GEMMs); meets
                               the 16-bit parity budget (2e-2) on all eight head outputs; paper-size geometry (head_dim 64) only */

int hft_model_create(hft_model** model, const hft_dims* dims);
int hft_model_destroy(hft_model* model);

/* The state_dict schema (reference Appendix: every key of Model_SPEC2MIDI.state_dict()).  The host passes one fp32
 * device pointer per tensor in this order. */
int hft_model_num_weights(const hft_model* model);
const char* hft_model_weight_name(const hft_model* model, int index);
int64_t hft_model_weight_numel(const hft_model* model, int index);

/* Register (or re-register after load_state_dict) the parameters.  The library copies them and derives what the
 * kernels use (collapsed 65-tap front filter, fused QKV / head matrices, 16-bit operand copies). */
int hft_model_set_weights(hft_model* model, const float* const* weights_dev, int n_weights, void* stream);

/* The 9-tuple Model_SPEC2MIDI.forward returns (model_spec2midi.py:35).  All fp32, contiguous:
 *   onset/offset/mpe  [B][n_frame][n_note]  sigmoid probabilities
 *   velocity          [B][n_frame][n_note][n_velocity]  raw logits
 *   attention         [B][n_frame][n_heads][n_note][n_bin]  softmax of the last cross-attention
 *   velocity_*_argmax [B][n_frame][n_note] int8   argmax over the velocity axis (first maximum, like torch.argmax), the only
 *                                                 thing AMT.transcript keeps of the logits (reference amt.py:107,113: `.argmax(2)`);
 *                                                 computed in the epilogue of the heads GEMM, so a caller that passes velocity_* =
 *                                                 NULL never moves the 2 x 5.8 MB of logits per segment
 * Any pointer may be NULL to skip writing that output. */
typedef struct hft_outputs {
  float* onset_A;
  float* offset_A;
  float* mpe_A;
  float* velocity_A;
  float* attention;
  float* onset_B;
  float* offset_B;
  float* mpe_B;
  float* velocity_B;
  int8_t* velocity_A_argmax;
  int8_t* velocity_B_argmax;
} hft_outputs;

/* spec_dev: [B][n_bin][n_margin + n_frame + n_margin] fp32 with element strides (the reference calls forward with
 * a non-contiguous .T view, amt.py:89, so strides are part of the interface). */
int hft_forward(hft_model* model, int precision, const float* spec_dev, int64_t stride_b, int64_t stride_bin, int64_t stride_t,
                int32_t batch, const hft_outputs* outputs, void* stream);

/* The two halves as separate calls -- Encoder_SPEC2MIDI.forward (model_spec2midi.py:60-106): spec [B, n_bin, n_frame + 2 margin] ->
 * enc_out_dev fp32 [B, n_frame, n_bin, hid];  Decoder_SPEC2MIDI.forward (:145-216): that memory -> the nine outputs.  fp32 CUDA-core
 * kernels (eval semantics); hft_forward is the fused hot path. */
int hft_forward_encoder(hft_model* model, const float* spec_dev, int64_t stride_b, int64_t stride_bin, int64_t stride_t, int32_t batch,
                        float* enc_out_dev, void* stream);
int hft_forward_decoder(hft_model* model, const float* enc_dev, int32_t batch, const hft_outputs* outputs, void* stream);

/* Largest batch one hft_forward call processes at once (bigger batches are looped internally). */
int hft_model_set_max_batch(hft_model* model, int32_t max_batch);
/* Give the activation work spaces back to the driver (they are sized by max_batch and only grow; lowering max_batch releases them too).
 * Synchronises the device.  Weights stay registered; the next hft_forward re-allocates what it needs. */
int hft_model_release_workspace(hft_model* model);

/* Number of kernels the last hft_forward / hft_logmel call on this thread launched (bench bookkeeping). */
int64_t hft_last_launch_count(void);

/* Per-kernel-class device timing for the roofline report (bench.py).  While enabled, every launch is bracketed by
 * CUDA events on the launching stream; hft_profile_read synchronises those events and returns the accumulated
 * milliseconds and launch count of one class, then clears it. */
#define HFT_KCLASS_LOGMEL 0
#define HFT_KCLASS_FRONT 1
#define HFT_KCLASS_GEMM 2
#define HFT_KCLASS_ATTENTION 3
#define HFT_KCLASS_NORM 4
#define HFT_KCLASS_HEADS 5
#define HFT_KCLASS_COUNT 6
int hft_profile_enable(int on);

/* Measurement probe (bench.py): launches a kernel of independent fp32 FMA chains on every SM (8 CTAs of 256 threads each, 16 chains per
 * thread, `iters` steps) and reports the flop it issues; timed by the caller with CUDA events it gives the chip's fp32 FMA rate under
 * the current clocks -- the measured denominator of `roofline.fp32` for the CUDA-core kernels (log-mel, training step). */
int hft_probe_fp32_fma(int32_t iters, float* scratch_dev, double* flop_out, void* stream);
int hft_profile_read(int kclass, double* ms, int64_t* launches);

/* ------------------------------------------------------------------------------------------------------------
 * Component entry points of the tensor-core path (used by the parity tests of the building blocks; a host may
 * also call them directly).  16-bit tensors are bf16 (bf16 != 0) or fp16, row-major, device memory.
 * ---------------------------------------------------------------------------------------------------------- */
/* out[M,N] = epilogue(a[M,K] * w[N,K]^T + bias).  epi: 0 store, 1 ReLU, 2 LayerNorm(x + resid)*gamma+beta (N <= 256).
 * M % 128 == 0, K % 64 == 0, N % 64 == 0.  Replaces nn.Linear (+ReLU / +residual+LayerNorm), model_spec2midi.py:328-330,
 * :357, :372-375, :236, :242. */
int hft_tc_linear(int bf16, int epi, const void* a16_dev, const void* w16_dev, const float* bias_dev, int64_t M, int32_t N, int32_t K,
                  void* out16_dev, const void* resid16_dev, const float* gamma_dev, const float* beta_dev, void* stream);
/* out[M,256] = LayerNorm(x + fc_2(relu(fc_1(x)))) * gamma + beta for hid_dim 256 / pf_dim 512 (w1 [512,256], w2 [256,512], 16-bit),
 * M % 256 == 0: the fused FFN kernel (CTA pairs, hidden activation kept in tensor memory).  Replaces
 * PositionwiseFeedforwardLayer.forward + the residual LayerNorm, model_spec2midi.py:369-378, :242. */
int hft_tc_ffn(int bf16, const void* x16_dev, const void* w1_16_dev, const float* b1_dev, const void* w2_16_dev, const float* b2_dev,
               const float* gamma_dev, const float* beta_dev, int64_t M, void* out16_dev, void* stream);
/* Self-attention over n_seq sequences of L tokens (L in {256, 128, 88}) stored as qkv[n_seq*L, 3*heads*dh]
 * (q | k | v): ctx[n_seq*L, heads*dh] = softmax(q k^T / sqrt(dh)) v, optional probs[n_seq, heads, L, L] fp32 (L = 256
 * only).  Replaces MultiHeadAttentionLayer.forward :335-355. */
int hft_tc_attention(int bf16, int32_t dh, int32_t heads, const void* qkv16_dev, int64_t n_seq, int32_t L, void* ctx16_dev,
                     float* probs_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Note decoding: the O(T x n_note) scans of AMT.mpe2note -- reference hftt_code/model/amt.py:179-344 (pure Python there).
 * Activation maps are [T][n_note] fp32 row-major device arrays (what AMT.transcript returns, uploaded).  The host mirror
 * (nylon_amt_b200/notes.py, mpe2note_device) compacts the flags and assembles the note dictionaries.
 * ---------------------------------------------------------------------------------------------------------- */
/* flags_dev [n_note][T] (uint8): 1 where frame t is a peak of column p: a[t][p] >= thr and the first differing value on each side
 * is lower (amt.py:196-212; all points of a plateau count). */
int hft_note_peaks(const float* a_dev, int64_t T, int32_t n_note, float thr, uint8_t* flags_dev, void* stream);
/* For n peaks idx_dev[n][2] = (pitch, frame): kind[i] = 1 and t32[i] = sub-frame peak time (amt.py:213-222, float32 arithmetic of
 * numpy >= 2) when the two direct neighbours differ, else kind[i] = 0 (the time is frame * hop_sec). */
int hft_note_peak_times(const float* a_dev, int64_t T, int32_t n_note, const int64_t* idx_dev, int64_t n, double hop_sec, uint8_t* kind_dev,
                        float* t32_dev, void* stream);
/* out[i] = first frame f in (frame_i, limit_i) with mpe[f][pitch_i] < thr, or -1 (amt.py:262-271). */
int hft_note_first_below(const float* mpe_dev, int64_t T, int32_t n_note, const int64_t* idx_dev, const int64_t* limit_dev, int64_t n, float thr,
                         int64_t* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Training step (BASELINE config 5: reduced hFT, batch 8 per GPU, Adam lr 1e-4, data parallel).
 * Replaces: the body of train() -- reference hftt_code/training/train.py:89-160 (model(input) in train mode, BCELoss on
 *           the six sigmoid outputs + CrossEntropyLoss on the two velocity logit tensors, loss = weight_A * loss_A +
 *           weight_B * loss_B, loss.backward(), optimizer.step()) with torch.optim.Adam(lr) of
 *           hftt_code/training/m_training.py:146.  fp32 CUDA-core kernels; dropout through hft_trainer_set_dropout.
 * Parameters and gradients are ONE flat fp32 vector each: the tensors of the state_dict in schema order, every tensor
 * padded to a multiple of 4 floats (hft_model_param_offset).  The flat gradient is the single bucket the data-parallel
 * configuration all-reduces over NCCL between hft_train_forward_backward and hft_adam_step.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct hft_trainer hft_trainer;

/* Length of the flat parameter / gradient vectors and the offset of tensor `index` inside them. */
int64_t hft_model_param_floats(const hft_model* model);
int64_t hft_model_param_offset(const hft_model* model, int index);
/* Device pointer to the model's own flat parameter vector (what hft_model_set_weights filled). */
float* hft_model_params(hft_model* model);
/* Copy the flat parameters into a caller-owned device buffer of hft_model_param_floats floats (checkpointing: the host
 * mirror writes them back into its nn.Parameters, m_training.py:275). */
int hft_model_get_params(hft_model* model, float* params_out_dev, void* stream);
/* Re-derive what the kernels consume (collapsed front filter, fused QKV, 16-bit copies) after the flat parameters were
 * updated in place (hft_adam_step on hft_model_params). */
int hft_model_refresh(hft_model* model, void* stream);

/* Activation tape and gradient work space for `batch` segments per step. */
int hft_trainer_create(hft_trainer** trainer, hft_model* model, int32_t batch);
int hft_trainer_destroy(hft_trainer* trainer);

/* Forward (train.py:90), loss (:139-151) and backward (:157) for one batch of `batch` segments:
 *   spec_dev [batch][n_bin][n_margin + n_frame + n_margin] fp32 (element strides given),
 *   label_onset/offset/mpe_dev [batch][n_frame][n_note] fp32 in [0, 1], label_velocity_dev [batch][n_frame][n_note] int64 in [0, n_velocity).
 * loss_dev[0] receives the scalar loss; grads_dev (hft_model_param_floats floats) is overwritten with dLoss/dParam. */
int hft_train_forward_backward(hft_trainer* trainer, const float* spec_dev, int64_t stride_b, int64_t stride_bin, int64_t stride_t,
                               const float* label_onset_dev, const float* label_offset_dev, const float* label_mpe_dev,
                               const int64_t* label_velocity_dev, float weight_A, float weight_B, float* loss_dev, float* grads_dev, void* stream);

/* The same step on the first `batch` <= capacity segments (the reference's DataLoader keeps the last, partial batch of an epoch:
 * drop_last=False, m_training.py; every mean of the loss is taken over the rows actually present). */
int hft_train_forward_backward_n(hft_trainer* trainer, int32_t batch, const float* spec_dev, int64_t stride_b, int64_t stride_bin, int64_t stride_t,
                                 const float* label_onset_dev, const float* label_offset_dev, const float* label_mpe_dev,
                                 const int64_t* label_velocity_dev, float weight_A, float weight_B, float* loss_dev, float* grads_dev, void* stream);

/* The two halves of the step for a host that computes the loss itself (the reference does: train.py:90 `model(input_spec)` in train mode,
 * :139-151 its own criteria, :157 `loss.backward()`, then a stock torch optimiser):
 *   hft_train_forward   Model_SPEC2MIDI.forward with model.train(): fills the tape, applies dropout, writes the eight head outputs
 *                       (onset/offset/mpe A and B as sigmoid probabilities, velocity A and B as raw logits; outputs->attention and the
 *                       argmax members are ignored -- train.py never reads the attention);
 *   hft_train_backward  dLoss/dParam from dLoss/d(output) of those eight tensors (out_grads: same struct, NULL member = zero gradient),
 *                       for the LAST hft_train_forward on this trainer and the same spec_dev.  grads_dev is overwritten. */
int hft_train_forward(hft_trainer* trainer, int32_t batch, const float* spec_dev, int64_t stride_b, int64_t stride_bin, int64_t stride_t,
                      const hft_outputs* outputs, void* stream);
int hft_train_backward(hft_trainer* trainer, const float* spec_dev, int64_t stride_b, int64_t stride_bin, int64_t stride_t, const hft_outputs* out_grads,
                       float* grads_dev, void* stream);

/* Dropout of the training forward (nn.Dropout(p) of the reference modules; the reference trains with p = 0.1, m_training.py).
 * Masks are counter based: element idx of site `site` is kept iff hash(seed, site, idx) >= p * 2^32 and scaled by 1 / (1 - p);
 * the backward regenerates them.  Call before every step with a fresh seed; p = 0 (default) disables every mask.
 * Site numbering (Le encoder layers, Ld decoder layers incl. layer zero):
 *   0                      encoder embedding dropout (model_spec2midi.py:95), idx = row * hid + col of [B*F*n_bin, hid]
 *   1 + 4 l + {0,1,2,3}    encoder layer l: attention probabilities (idx = ((seq * heads + head) * Lq + i) * Lk + j), attention sub-layer
 *                          output [rows, hid], FFN hidden [rows, pf], FFN sub-layer output [rows, hid]
 *   D0 = 1 + 4 Le, + {0,1,2,3}           decoder layer zero: cross probabilities, cross output, FFN hidden, FFN output
 *   D0 + 4 + 6 (l - 1) + {0..5}, l >= 1  decoder layer l: self probabilities, self output, cross probabilities, cross output, FFN hidden, FFN output
 *   T0 = D0 + 4 + 6 (Ld - 1)             time embedding dropout (:191), rows in (b, note, frame) order
 *   T0 + 1 + 4 l + {0,1,2,3}             time layer l (same four sites as an encoder layer) */
int hft_trainer_set_dropout(hft_trainer* trainer, float p, uint32_t seed);
/* mask_out_dev[i] = kept ? 1 / (1 - p) : 0 for i < n of one site: the multiplier the kernels apply (test / restatement aid). */
int hft_dropout_mask(float p, uint32_t seed, int32_t site, int64_t n, float* mask_out_dev, void* stream);

/* Component entry: the multi-head attention of the training step alone -- softmax(Q K^T / sqrt(dh)) V with dropout on the
 * probabilities (MultiHeadAttentionLayer.forward, model_spec2midi.py:342-348) and, when d_ctx_dev is not NULL, its backward
 * (what loss.backward(), training/train.py:158, runs through that node).  All tensors fp32 on the device:
 *   q [n_seq * lq, ldq] (q_seq_stride = elements between consecutive sequences; 0 = the same queries for every sequence),
 *   k, v [n_seq * lk, ldkv]; head h occupies columns h * dh .. of each; ctx, d_ctx [n_seq * lq, heads * dh]; lse, d_buf [n_seq, heads, lq];
 *   dq [n_seq * lq, lddq], dk, dv [n_seq * lk, lddkv].
 * use_tc: 1 = the tcgen05 kernels (split-fp16 products, fp32 accumulate; head_dim 32, lq, lk <= 256), 0 = the fp32 CUDA-core kernels,
 * -1 = what the training step picks (tcgen05 where supported unless HFT_TRAIN_TC=0). */
int hft_train_attention(int32_t use_tc, int32_t dh, int32_t heads, const float* q_dev, int32_t ldq, int64_t q_seq_stride, const float* k_dev,
                        const float* v_dev, int32_t ldkv, int64_t n_seq, int32_t lq, int32_t lk, float p_drop, uint32_t seed, int32_t site,
                        float* ctx_dev, float* lse_dev, const float* d_ctx_dev, float* dq_dev, int32_t lddq, float* dk_dev, float* dv_dev,
                        int32_t lddkv, float* d_buf_dev, void* stream);

/* Component entry: one Linear of the training step on fp32 device tensors.
 *   w_kn = 0: c[m, n] = a[m, k] w[n, k]^T + bias (optional ReLU)            -- nn.Linear forward (model_spec2midi.py:322-378)
 *   w_kn = 1: c[m, n] = a[m, k] w[k, n], then c = mask > 0 ? c * mask_scale : 0 when mask_dev is given   -- its input gradient under loss.backward()
 * accum != 0 adds the result to c instead of overwriting it.  use_tc: 1 = the tcgen05 kernel (split-fp16 products, fp32 accumulate;
 * k = 64 or 128, n % 16 == 0, n <= 256), 0 = the fp32 CUDA-core kernels, -1 = what the training step picks. */
int hft_train_linear(int32_t use_tc, int32_t w_kn, const float* a_dev, int32_t lda, const float* w_dev, int32_t ldw, const float* bias_dev,
                     float* c_dev, int32_t ldc, int64_t m, int32_t n, int32_t k, int32_t relu, int32_t accum, const float* mask_dev, int32_t ldm,
                     float mask_scale, void* stream);

/* Component entry: the weight gradient of one Linear of the training step: dw[n, k] += dy[m, n]^T x[m, k], db[n] += sum over m of dy[m, n]
 * (db_dev may be NULL), on fp32 device tensors (what loss.backward(), training/train.py:158, accumulates into nn.Linear.weight.grad / .bias.grad).
 * use_tc: 1 = the tcgen05 kernel (three-piece bf16 operands, fp32 accumulate; k = 64 or 128, n % 8 == 0, n + k <= 256), 0 = the fp32 CUDA-core
 * kernel, -1 = what the training step picks. */
int hft_train_linear_wgrad(int32_t use_tc, const float* dy_dev, int32_t ldy, const float* x_dev, int32_t ldx, float* dw_dev, int32_t ldw, float* db_dev,
                           int64_t m, int32_t n, int32_t k, void* stream);

/* torch.optim.Adam (no weight decay, no amsgrad) on flat vectors: grads are multiplied by grad_scale first (1 / world size
 * after a sum all-reduce); step counts from 1. */
int hft_adam_step(float* params_dev, const float* grads_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n, float lr, float beta1, float beta2,
                  float eps, int64_t step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HFT_SM100_H */
