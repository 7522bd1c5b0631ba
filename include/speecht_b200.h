/* speecht_b200 -- C ABI of the B200-native Wav2Letter hot path.
 *
 * louiskirsch/speechT has no FFI / plugin boundary of its own: its hot path is a handful of TensorFlow-1 and
 * librosa library calls made from Python.  Each entry point below replaces ONE of those
 * library calls; the comment above it names the reference call site (file:line under the speechT checkout) it
 * stands in for.  The reference-side binding a maintainer would add is the ctypes stub in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - the caller owns every buffer (including workspaces, sized by the *_workspace_bytes functions);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises the host except
 *     st_device_sync and functions documented as host-synchronous;
 *   - return value: 0 = ST_OK, negative = error (st_last_error() gives a thread-local message).  No C++ exception
 *     crosses the boundary; the Python mirror raises.
 *   - layouts are the reference's: activations [batch, time, channels] (NWC), filters [width, cin, cout]
 *     (the `export` .npy layout, exporting.py:30-40), logits addressed through explicit (stride_t, stride_b)
 *     element strides so the time-major view of speech_model.py:295 costs no copy.
 */
#ifndef SPEECHT_B200_H_
#define SPEECHT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ST_OK 0
#define ST_ERR_INVALID_ARG (-1)
#define ST_ERR_CTC_LABELS (-2)     /* tf.nn.ctc_loss "Not enough time for target transition sequence" / bad label id */
#define ST_ERR_CUDA (-3)
#define ST_ERR_UNSUPPORTED (-4)

typedef void* st_stream_t;         /* cudaStream_t */

/* conv precision modes of the tensor-core path (st_tc_*): number of bf16 planes an fp32 value is split into. */
#define ST_PREC_BF16 1             /* 1 plane : plain bf16 operands, fp32 accumulate                      */
#define ST_PREC_BF16X3 2           /* 2 planes: hi+lo split, 3 tcgen05 products -> ~2^-17 operand error     */
#define ST_PREC_BF16X6 3           /* 3 planes: 6 products -> fp32-equivalent operands (error = accumulation) */

int st_version(void);
const char* st_last_error(void);
int st_device_sync(void);

/* ---- a12: tf.nn.ctc_greedy_decoder(logits, seq_len, merge_repeated)          speech_model.py:113-115 ----
 * logits[t*stride_t + b*stride_b + c], t < T, b < B, c < C.  For every utterance: per frame t < seq_len[b] argmax
 * over raw logits (first maximum wins), emit unless blank or (merge_repeated and equal to previous frame's argmax).
 * out_values [B, T] int32 (row b holds out_counts[b] labels), out_counts [B], neg_sum_logits [B] (= TF's
 * log_probabilities column).  The sparse (indices, values, dense_shape) triple is assembled by the host mirror. */
int st_ctc_greedy_decode(const float* logits, int64_t stride_t, int64_t stride_b, int T, int B, int C,
                         const int32_t* seq_len, int blank, int merge_repeated,
                         int32_t* out_values, int32_t* out_counts, float* neg_sum_logits, st_stream_t stream);

/* ---- a8/a9: tf.nn.ctc_loss(labels, logits, seq_len) + its gradient           speech_model.py:74-75 ----
 * st_ctc_validate_labels_host is host-side and mirrors TF's InvalidArgument checks (label ids in [0, blank),
 * len + adjacent repeats <= seq_len, seq_len <= T); it returns ST_ERR_CTC_LABELS and sets st_last_error.
 * st_ctc_loss: loss[b] = -log p(label_b | logits_b); grad (nullable) gets grad_scale * dloss_b/dlogits with the
 * SAME strides as logits and exact zeros for frames t >= seq_len[b] (grad_scale = 1/B folds tf.reduce_mean,
 * speech_model.py:75).  grad_planes (nullable) additionally receives the gradient split into bf16 planes
 * [n_planes][B][T][c_pad] for the tensor-core backward (c_pad = channel stride of the planes, >= C).
 * status [B] int32: 0 ok, 1 = infeasible/invalid label seen on the device (loss set to +inf, grad zero). */
int st_ctc_validate_labels_host(const int32_t* labels_host, const int32_t* label_offsets_host,
                                const int32_t* seq_len_host, int B, int T, int blank);
size_t st_ctc_workspace_bytes(int T, int B, int C, int max_label_len);
int st_ctc_loss(const float* logits, int64_t stride_t, int64_t stride_b, int T, int B, int C,
                const int32_t* labels, const int32_t* label_offsets, int max_label_len,
                const int32_t* seq_len, int blank, float* loss, float* grad, float grad_scale,
                void* grad_planes, int n_planes, int c_pad,
                int32_t* status, void* workspace, size_t workspace_bytes, st_stream_t stream);

/* ---- a6: tf.nn.conv1d(value, filters, stride, 'SAME') + bias_add + relu       speech_model.py:155,173,177 ----
 * Exact-fp32 CUDA-core path (FFMA, fp32 accumulate).  x [B,T,Cin], w [K,Cin,Cout], bias [Cout] (nullable),
 * y [B,ceil(T/stride),Cout].  TF 'SAME': pad_total = max((out-1)*stride+K-T,0), left = pad_total/2. */
int st_conv1d_fwd_f32(const float* x, const float* w, const float* bias, float* y,
                      int B, int T, int Cin, int Cout, int K, int stride, int relu, st_stream_t stream);
/* Gradients of the same op (what TF autodiff produces for speech_model.py:78).  y_act (nullable) is the layer's
 * post-ReLU output: when given, dy is masked with (y_act > 0) on load (relu backward fused).
 * bwd_data: dx [B,T,Cin].  bwd_filter: dw [K,Cin,Cout] and db [Cout] are OVERWRITTEN. */
int st_conv1d_bwd_data_f32(const float* dy, const float* y_act, const float* w, float* dx,
                           int B, int T, int Cin, int Cout, int K, int stride, st_stream_t stream);
int st_conv1d_bwd_filter_f32(const float* x, const float* dy, const float* y_act, float* dw, float* db,
                             int B, int T, int Cin, int Cout, int K, int stride, st_stream_t stream);

/* ---- a10/a11: tf.clip_by_global_norm + tf.train.AdamOptimizer(eps=1e-3)       speech_model.py:77-82 ----
 * All 22 tensors live in ONE flat fp32 buffer (params / grads / m / v each [n]).  st_sumsq accumulates
 * sum(g^2) into *accum (double, device; the caller zeroes it -- several calls may add into it).
 * st_clip_adam: g' = g * grad_prescale * clip*min(1/norm, 1/clip), norm = sqrt(*normsq)*grad_prescale;
 * TF1 Adam: lr_t = lr*sqrt(1-b2^step)/(1-b1^step); p -= lr_t*m/(sqrt(v)+eps).  step counts from 1. */
int st_sumsq(const float* g, int64_t n, double* accum, int zero_first, st_stream_t stream);
int st_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                 float eps, int64_t step, float max_norm, const double* normsq, float grad_prescale,
                 st_stream_t stream);

/* ---- a1-a3: calc_power_spectrogram                                            preprocessing.py:36-58 ----
 * librosa.feature.melspectrogram(n_fft=512, hop, n_mels) -> power_to_db(ref=max, top_db=80) -> (x-mean)/std.
 * wav [B, wav_stride] fp32, n_samples [B]; mel_basis [n_mels, n_fft/2+1] fp32 (host mirror builds the Slaney
 * filterbank); out [B, T_max, n_mels] with T_b = 1 + n_samples[b]/hop frames valid (rest zero = batch padding,
 * speech_input.py:39-43), out_frames [B].  n_fft must be 512. */
size_t st_melspec_workspace_bytes(int B, int max_samples, int n_fft, int hop, int n_mels);
int st_melspec(const float* wav, int64_t wav_stride, const int32_t* n_samples, int B, int max_samples,
               const float* mel_basis, int n_fft, int hop, int n_mels,
               float* out, int T_max, int32_t* out_frames, void* workspace, size_t workspace_bytes,
               st_stream_t stream);

/* ---- a6 on the tensor cores + a13: the whole step as a plan                     speech_model.py:235,275-295 ----
 * For a fixed (batch B, time T) shape the 11-layer stack is a fixed launch sequence of tcgen05/TMEM/TMA
 * implicit-GEMM kernels (csrc/conv_tc.cu) over bf16 operand planes (n_planes = ST_PREC_BF16, _BF16X3 or _BF16X6).
 * The plan owns no device memory: the caller provides one arena (st_plan_arena_bytes, 1024-byte aligned) and the
 * flat fp32 parameter / gradient buffers (layout: per layer filters [K,Cin,Cout] then bias [Cout], each start
 * rounded up to 64 floats; st_plan_param_floats gives the total).
 *   st_plan_pack_weights : fp32 parameters -> bf16 operand planes (call after every parameter change)
 *   st_plan_forward      : inputs [B,T,input_size] fp32 -> logits fp32 [B,T',32] at st_plan_logits (classes 0..C-1
 *                          valid; view it time-major with stride_t=32, stride_b=T'*32, T' = st_plan_logit_frames)
 *   st_plan_backward     : consumes d(loss)/d(logits) from the bf16 planes [n_planes][B][T'][64] at
 *                          st_plan_dlogits_planes (st_ctc_loss writes them with c_pad=64) -> flat gradient buffer
 *   st_plan_get_activation: output of layer 0..9 merged back to fp32 [B,T',Cout] (parity tests); -1 = input planes
 * Stream semantics: every call only enqueues work; results are ordered on `stream`.  A plan may run HBM-bound passes
 * (part of the packing, the zeroing of the gradient buffer, one combine pass of backward) on a side stream of its own
 * underneath tensor-core launches -- it forks behind what `stream` holds at the call and the consumer (st_plan_forward in
 * front of layer 8, st_plan_backward at its start / end) makes `stream` wait for it, so callers never see the side
 * stream (SPEECHT_B200_OVERLAP=0 keeps everything on `stream`).  st_plan_destroy synchronises it. */
typedef struct st_plan st_plan;
int st_plan_create(st_plan** out, int B, int T, int input_size, int num_classes, int n_planes);
int st_plan_destroy(st_plan* plan);
size_t st_plan_arena_bytes(const st_plan* plan);
int64_t st_plan_param_floats(const st_plan* plan);
int st_plan_logit_frames(const st_plan* plan);
int st_plan_bind(st_plan* plan, void* arena, size_t arena_bytes, float* params, float* grads);
int st_plan_pack_weights(st_plan* plan, st_stream_t stream);
int st_plan_forward(st_plan* plan, const float* inputs, st_stream_t stream);
int st_plan_backward(st_plan* plan, st_stream_t stream);
int st_plan_backward_range(st_plan* plan, int layer_hi, int layer_lo, st_stream_t stream);  /* 10 first, downwards */
/* Data parallelism (no reference counterpart: the reference is single-process).  Tensor-core grids launched after the
 * call leave n_sms SMs free for a concurrent collective (the NCCL allreduce overlapped with backward); 0 = all SMs. */
int st_plan_reserve_sms(st_plan* plan, int n_sms);
/* Optional, before st_plan_forward of a train step: zeroes the flat gradient buffer on the plan's side stream underneath
 * the forward pass (TF zero-initialises gradient accumulators per sess.run, speech_model.py:78); st_plan_backward then
 * waits for it instead of zeroing in line. */
int st_plan_prepare_backward(st_plan* plan, st_stream_t stream);
float* st_plan_logits(st_plan* plan);
void* st_plan_dlogits_planes(st_plan* plan);
int st_plan_get_activation(st_plan* plan, int layer, float* dst, st_stream_t stream);
int st_plan_launches(const st_plan* plan);
/* Packed filters live in a leading arena region whose offsets do not depend on (B, T): plans bound into the SAME arena
 * share them, and a plan needs st_plan_pack_weights again only after the parameters changed or when no plan of its
 * filter set (0 / 1 / 2 = fast-FIR level of layer 8) has packed since. */
int st_plan_filter_set(const st_plan* plan);
/* bench support: CUDA-event pair around every tensor-core launch (kind 0 = forward conv, 1 = data gradient,
 * 2 = filter gradient); st_plan_read_timings is host-synchronous and returns the number of records written. */
int st_plan_set_timing(st_plan* plan, int enable);
int st_plan_read_timings(st_plan* plan, int* kind, int* layer, double* flops, float* ms, int max_records);

/* ---- caller of a1: audio loading                                                 preprocessing.py:169 ----
 * librosa.load(audio_file) decodes LibriSpeech's .flac files through soundfile / audioread; neither exists in this
 * image, so the native FLAC decoder is part of the library.  HOST functions (no GPU, no stream): `data` is the whole
 * file in host memory.  st_flac_info_host: info = {sample_rate, channels, bits_per_sample}, samples per channel
 * (0 = unknown) and the MD5 of the unencoded PCM (STREAMINFO).  st_flac_decode_host: interleaved int32 samples
 * out[sample][channel]; *decoded = samples per channel.  Frame CRC-8 / CRC-16 are verified. */
int st_flac_info_host(const uint8_t* data, size_t nbytes, int32_t* info, int64_t* total_samples, uint8_t* md5);
int st_flac_decode_host(const uint8_t* data, size_t nbytes, int32_t* out, int64_t capacity, int64_t* decoded);

/* Debug: per-CTA %globaltimer timeline of ONE tensor-core conv launch (the launch_index-th forward / data-gradient
 * launch after this call): buf[grid][8] int64 device memory -- 0 entry, 1 previous grid complete, 2 first operands
 * landed, 3 last MMA issued, 4 accumulator complete, 5 epilogue issued, 6 staging tiles drained, 7 exit.
 * buf = NULL switches it off.  Used by tools/conv_timeline.py; no reference counterpart. */
int st_debug_conv_timeline(int64_t* buf, int launch_index);

#ifdef __cplusplus
}
#endif
#endif  /* SPEECHT_B200_H_ */
