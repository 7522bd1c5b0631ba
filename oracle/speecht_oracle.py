"""CPU oracle for the speechT Wav2Letter hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, in plain numpy, the arithmetic that louiskirsch/speechT delegates to
TensorFlow 1.x and librosa for the path SURVEY.md section 8(a) lists.  It is the *checker*:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.  Nothing under speecht_b200/ imports it; the product path fails loudly without the
CUDA library.

PARITY: PINNED TO PUBLISHED KNOWN-ANSWER VECTORS ONLY, NOT TO A RUN OF THE REFERENCE ("parity
unpinned" in the strict sense of the build rules).  The reference's own test-suite
(speecht/tests/test_speechCorpusReader.py) holds one value on this path -- the 114 881 samples its
LibriSpeech FLAC fixture loads to, reproduced by the native FLAC decoder + resampler -- and neither
TensorFlow 1.x nor librosa can be imported in the build container or on the GPU box (no wheels, no
network).  Library semantics below are restated from the published behaviour of
  * tensorflow 1.x (requirements.txt:2, unpinned ">=1.0.1"): tf.nn.conv1d 'SAME',
    tf.nn.ctc_loss, tf.nn.ctc_greedy_decoder, tf.clip_by_global_norm, tf.train.AdamOptimizer
  * librosa 0.5-0.7 (requirements.txt:5, unpinned ">=0.5.0"): feature.melspectrogram,
    power_to_db, filters.mel
and are checked in tests/test_oracle.py against
  * the literals TensorFlow's own kernel tests publish (tests/golden/tf_published_vectors.py, typed in
    from the published test sources): ctc_loss_op_test testBasic (loss + gradient, six digits),
    ctc_decoder_ops_test testCTCGreedyDecoder, conv_ops_test (1x1, stride-2 SAME, kernel smaller than
    stride, stride-2 data / filter gradients -- integer valued, exact), clip_ops_test
    testClipByGlobalNormClipped, adam_test's adam_update_numpy;
  * the mel-scale examples printed in librosa's docstrings (mel_frequencies(n_mels=40), hz_to_mel,
    mel_to_hz);
  * independent implementations available here: torch conv1d / ctc_loss + autograd, brute-force CTC
    path enumeration, torchaudio's librosa-compatible mel pipeline, transformers' TF-port padding.
What stays unpinned: the 11-layer stack end to end, exact-tie behaviour of the greedy arg-max, librosa's
triangle construction / Slaney normalisation / power_to_db beyond the torchaudio cross-check.  Claims
about those read "vs CPU restatement of the reference".

Every function cites the reference file:line whose behaviour it follows.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

# ----------------------------------------------------------------------------------------
# vocabulary  (reference: speecht/vocabulary.py:16-21, speech_model.py:301)
# ----------------------------------------------------------------------------------------
VOCAB_SIZE = 28                 # a-z, apostrophe, space
NUM_CLASSES = VOCAB_SIZE + 1    # + CTC blank
BLANK = NUM_CLASSES - 1         # TF1 ctc ops use the LAST class as blank

# Wav2Letter layer table (reference: speech_model.py:275-292): (filter_width, stride, cin, cout, relu)
def layer_table(input_size: int = 128, num_classes: int = NUM_CLASSES):
  layers = [(48, 2, input_size, 250, True)]
  layers += [(7, 1, 250, 250, True)] * 7
  layers += [(32, 1, 250, 2000, True), (1, 1, 2000, 2000, True), (1, 1, 2000, num_classes, False)]
  return layers


# ----------------------------------------------------------------------------------------
# a14: Xavier-uniform initialisation (reference: speech_model.py:150-152)
# ----------------------------------------------------------------------------------------
def xavier_weights(rng: np.random.Generator, input_size: int = 128, num_classes: int = NUM_CLASSES,
                   dtype=np.float32, layers=None):
  """tf.contrib.layers.xavier_initializer() on a [K, Cin, Cout] filter: uniform(+-sqrt(6/(fan_in+fan_out)))
  with fan_in = K*Cin, fan_out = K*Cout; bias = 0 (speech_model.py:152).  TF's RNG stream cannot be
  reproduced, so parity runs inject these weights into both sides."""
  weights = []
  for (k, _s, cin, cout, _r) in (layers or layer_table(input_size, num_classes)):
    limit = math.sqrt(6.0 / (k * cin + k * cout))
    w = rng.uniform(-limit, limit, size=(k, cin, cout)).astype(dtype)
    b = np.zeros((cout,), dtype=dtype)
    weights.append((w, b))
  return weights


# ----------------------------------------------------------------------------------------
# a1-a3: feature extraction (reference: preprocessing.py:29-58)
# ----------------------------------------------------------------------------------------
def _hz_to_mel_slaney(f):
  f = np.asarray(f, dtype=np.float64)
  f_sp = 200.0 / 3
  mels = f / f_sp
  min_log_hz = 1000.0
  min_log_mel = min_log_hz / f_sp
  logstep = np.log(6.4) / 27.0
  with np.errstate(divide='ignore', invalid='ignore'):
    log_t = min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep
  return np.where(f >= min_log_hz, log_t, mels)


def _mel_to_hz_slaney(m):
  m = np.asarray(m, dtype=np.float64)
  f_sp = 200.0 / 3
  freqs = f_sp * m
  min_log_hz = 1000.0
  min_log_mel = min_log_hz / f_sp
  logstep = np.log(6.4) / 27.0
  return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr: float, n_fft: int = 512, n_mels: int = 128, fmin: float = 0.0, fmax=None):
  """librosa.filters.mel(sr, n_fft, n_mels, fmin=0, fmax=sr/2, htk=False, norm=1): Slaney mel scale,
  triangular filters, each scaled by 2/(f[i+2]-f[i]) (area normalisation).  -> [n_mels, 1+n_fft//2]."""
  if fmax is None:
    fmax = sr / 2.0
  n_bins = 1 + n_fft // 2
  fftfreqs = np.linspace(0.0, sr / 2.0, n_bins)
  mel_f = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2))
  fdiff = np.diff(mel_f)
  ramps = mel_f[:, None] - fftfreqs[None, :]
  weights = np.zeros((n_mels, n_bins))
  for i in range(n_mels):
    lower = -ramps[i] / fdiff[i]
    upper = ramps[i + 2] / fdiff[i + 1]
    weights[i] = np.maximum(0.0, np.minimum(lower, upper))
  enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
  weights *= enorm[:, None]
  return weights


def hann_periodic(n: int):
  """scipy.signal.get_window('hann', n, fftbins=True) as librosa's stft uses it."""
  return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def stft_power(audio: np.ndarray, n_fft: int = 512, hop_length: int = 160, dtype=np.float64):
  """|librosa.stft(y, n_fft, hop_length, center=True, pad_mode='reflect', window='hann')|**2.
  Frame count = 1 + len(y)//hop.  -> [1+n_fft//2, T]."""
  y = np.asarray(audio, dtype=dtype)
  if y.shape[0] <= n_fft // 2:
    raise ValueError('audio shorter than n_fft/2 cannot be reflect-padded')
  y = np.pad(y, n_fft // 2, mode='reflect')
  n_frames = 1 + (y.shape[0] - n_fft) // hop_length
  idx = np.arange(n_fft)[None, :] + hop_length * np.arange(n_frames)[:, None]
  frames = y[idx] * hann_periodic(n_fft).astype(dtype)[None, :]
  spec = np.fft.rfft(frames, axis=1)
  return (spec.real ** 2 + spec.imag ** 2).T.astype(dtype)


def power_to_db(S: np.ndarray, amin: float = 1e-10, top_db: float = 80.0):
  """librosa.power_to_db(S, ref=np.max) (preprocessing.py:53)."""
  ref = np.max(S)
  log_spec = 10.0 * np.log10(np.maximum(amin, S)) - 10.0 * np.log10(np.maximum(amin, ref))
  return np.maximum(log_spec, log_spec.max() - top_db)


def normalize(values: np.ndarray):
  """preprocessing.py:29-33: one scalar mean and one scalar (population) std over the whole matrix."""
  return (values - np.mean(values)) / np.std(values)


def calc_power_spectrogram(audio_data, samplerate, n_mels=128, n_fft=512, hop_length=160, dtype=np.float64):
  """preprocessing.py:36-58 -> [time, n_mels]."""
  S = stft_power(audio_data, n_fft, hop_length, dtype=dtype)
  mel = mel_filterbank(samplerate, n_fft, n_mels).astype(dtype) @ S
  return normalize(power_to_db(mel)).T.astype(dtype)


# ----------------------------------------------------------------------------------------
# a4-a5: batch assembly (reference: speech_input.py:27-69)
# ----------------------------------------------------------------------------------------
def pad_batch(input_list: Sequence[np.ndarray], input_size: int, dtype=np.float32):
  """speech_input.py:27-45: zero-pad to the batch max; no masking anywhere downstream."""
  lengths = np.array([x.shape[0] for x in input_list], dtype=np.int32)
  out = np.zeros((len(input_list), int(lengths.max()), input_size), dtype=dtype)
  for i, x in enumerate(input_list):
    out[i, :x.shape[0]] = x
  return out, lengths


def sparse_labels(label_list: Sequence[Sequence[int]], max_time: int):
  """speech_input.py:48-69: COO (indices [N,2], values [N], dense_shape [B, max_time])."""
  idx, val = [], []
  for b, lab in enumerate(label_list):
    for j, v in enumerate(lab):
      idx.append((b, j))
      val.append(int(v))
  return (np.array(idx, dtype=np.int64).reshape(-1, 2), np.array(val, dtype=np.int64),
          np.array([len(label_list), max_time], dtype=np.int64))


# ----------------------------------------------------------------------------------------
# a6-a7: conv stack (reference: speech_model.py:128-181, 275-295)
# ----------------------------------------------------------------------------------------
def same_padding(t_in: int, k: int, stride: int) -> Tuple[int, int, int]:
  """TF 'SAME': out = ceil(T/s); pad_total = max((out-1)*s + K - T, 0); left = total//2, rest right."""
  out = -(-t_in // stride)
  total = max((out - 1) * stride + k - t_in, 0)
  return out, total // 2, total - total // 2


def conv1d_same(x: np.ndarray, w: np.ndarray, bias: np.ndarray, stride: int, relu: bool):
  """tf.nn.conv1d(value[B,T,Cin], filters[K,Cin,Cout], stride, 'SAME') (cross-correlation) + bias_add
  + optional relu (speech_model.py:155,173,177).  One BLAS GEMM per filter tap."""
  k, cin, cout = w.shape
  b, t, _ = x.shape
  out, left, right = same_padding(t, k, stride)
  xp = np.pad(x, ((0, 0), (left, right), (0, 0)))
  y = np.zeros((b, out, cout), dtype=x.dtype)
  span = (out - 1) * stride + 1
  for j in range(k):
    y += np.matmul(xp[:, j:j + span:stride], w[j])
  y = y + bias
  return np.maximum(y, 0) if relu else y


def conv1d_same_backward(x, w, stride, dy):
  """Gradients of y = conv1d_same(x, w) (pre-activation) wrt x, w, bias."""
  k, cin, cout = w.shape
  b, t, _ = x.shape
  out, left, right = same_padding(t, k, stride)
  xp = np.pad(x, ((0, 0), (left, right), (0, 0)))
  span = (out - 1) * stride + 1
  dy2 = dy.reshape(-1, cout)
  dw = np.empty_like(w, dtype=dy.dtype)
  dxp = np.zeros((b, t + left + right, cin), dtype=dy.dtype)
  for j in range(k):
    dw[j] = np.ascontiguousarray(xp[:, j:j + span:stride]).reshape(-1, cin).T @ dy2
    dxp[:, j:j + span:stride] += np.matmul(dy, w[j].T)
  db = dy.sum(axis=(0, 1))
  dx = dxp[:, left:left + t]
  return dx, dw, db


def wav2letter_forward(inputs: np.ndarray, weights, layers=None, keep_activations=False):
  """speech_model.py:275-295.  inputs [B,T,input_size] -> logits time-major [T',B,num_classes]."""
  layers = layers or layer_table(inputs.shape[2], weights[-1][0].shape[2])
  acts = [inputs]
  x = inputs
  for (k, s, cin, cout, relu), (w, b) in zip(layers, weights):
    x = conv1d_same(x, w, b, s, relu)
    if keep_activations:
      acts.append(x)
  logits = np.transpose(x, (1, 0, 2))
  return (logits, acts) if keep_activations else logits


def wav2letter_backward(acts, weights, dlogits_tm, layers=None):
  """Backprop dL/dlogits (time-major) through the stack; returns [(dw, db)] per layer."""
  layers = layers or layer_table(acts[0].shape[2], weights[-1][0].shape[2])
  dy = np.transpose(dlogits_tm, (1, 0, 2))
  grads = [None] * len(layers)
  for li in reversed(range(len(layers))):
    (k, s, cin, cout, relu), (w, b) = layers[li], weights[li]
    if relu:
      dy = dy * (acts[li + 1] > 0)
    dx, dw, db = conv1d_same_backward(acts[li], w, s, dy)
    grads[li] = (dw, db)
    dy = dx
  return grads


# ----------------------------------------------------------------------------------------
# a8-a9: CTC loss and its gradient (reference: speech_model.py:74-75 -> tf.nn.ctc_loss)
# ----------------------------------------------------------------------------------------
class CTCLabelError(ValueError):
  pass


def _logsumexp2(a, b):
  m = np.maximum(a, b)
  with np.errstate(invalid='ignore', divide='ignore'):
    r = m + np.log(np.exp(a - m) + np.exp(b - m))
  return np.where(np.isneginf(m), -np.inf, r)


def ctc_loss_and_grad(logits_tm: np.ndarray, labels: Sequence[Sequence[int]], seq_len: Sequence[int],
                      blank: int = None, dtype=np.float64):
  """tf.nn.ctc_loss(labels, logits, seq_len) with TF1 defaults (preprocess_collapse_repeated=False,
  ctc_merge_repeated=True, time_major=True, blank = num_classes-1) and the gradient TF registers for it.

  logits_tm [T,B,C] unnormalised.  Returns (loss [B] = -log p(label|x), grad [T,B,C] = dloss_b/dlogits).
  Frames t >= seq_len[b] get zero gradient.  Raises CTCLabelError ("Not enough time for target transition
  sequence") when a label cannot be emitted in seq_len frames, as TF does
  (ignore_longer_outputs_than_inputs=False), and for label ids outside [0, blank)."""
  T, B, C = logits_tm.shape
  blank = C - 1 if blank is None else blank
  x = logits_tm.astype(dtype)
  loss = np.zeros((B,), dtype=dtype)
  grad = np.zeros_like(x)
  NEG = -np.inf
  for b in range(B):
    lab = np.asarray(labels[b], dtype=np.int64)
    L = lab.shape[0]
    tb = int(seq_len[b])
    if tb > T:
      raise CTCLabelError('sequence_length(%d) > max_time' % b)
    if np.any(lab < 0) or np.any(lab >= blank):
      raise CTCLabelError('label id outside [0, num_classes-1) in batch %d' % b)
    repeats = int(np.sum(lab[1:] == lab[:-1])) if L > 1 else 0
    if L + repeats > tb:
      raise CTCLabelError('Not enough time for target transition sequence (required: %d, available: %d)'
                          % (L + repeats, tb))
    if tb == 0:
      continue
    S = 2 * L + 1
    ext = np.full((S,), blank, dtype=np.int64)
    ext[1::2] = lab
    # skip transition s-2 -> s allowed when ext[s] != blank and ext[s] != ext[s-2]
    can_skip = np.zeros((S,), dtype=bool)
    if S > 2:
      can_skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
    xb = x[:tb, b, :]
    lsm = xb - xb.max(axis=1, keepdims=True)
    lsm = lsm - np.log(np.exp(lsm).sum(axis=1, keepdims=True))
    lp = lsm[:, ext]                                                   # [tb, S]
    alpha = np.full((tb, S), NEG, dtype=dtype)
    alpha[0, 0] = lp[0, 0]
    if S > 1:
      alpha[0, 1] = lp[0, 1]
    for t in range(1, tb):
      prev = alpha[t - 1]
      acc = prev.copy()
      acc[1:] = _logsumexp2(acc[1:], prev[:-1])
      sk = np.full((S,), NEG, dtype=dtype)
      sk[2:] = np.where(can_skip[2:], prev[:-2], NEG)
      acc = _logsumexp2(acc, sk)
      alpha[t] = acc + lp[t]
    # beta excludes the emission at t (TF's convention) so alpha+beta is the path mass through (t,s)
    beta = np.full((tb, S), NEG, dtype=dtype)
    beta[tb - 1, S - 1] = 0.0
    if S > 1:
      beta[tb - 1, S - 2] = 0.0
    for t in range(tb - 2, -1, -1):
      nxt = beta[t + 1] + lp[t + 1]
      acc = nxt.copy()
      acc[:-1] = _logsumexp2(acc[:-1], nxt[1:])
      sk = np.full((S,), NEG, dtype=dtype)
      sk[:-2] = np.where(can_skip[2:], nxt[2:], NEG)
      acc = _logsumexp2(acc, sk)
      beta[t] = acc
    log_p = _logsumexp2(alpha[tb - 1, S - 1], alpha[tb - 1, S - 2] if S > 1 else np.array(NEG, dtype=dtype))
    loss[b] = -log_p
    ab = alpha + beta                                                  # [tb, S]
    occ = np.zeros((tb, C), dtype=dtype)
    with np.errstate(invalid='ignore'):
      m = ab.max(axis=1, keepdims=True)
      e = np.where(np.isneginf(ab), 0.0, np.exp(ab - np.where(np.isneginf(m), 0.0, m)))
    for s in range(S):
      occ[:, ext[s]] += e[:, s]
    with np.errstate(divide='ignore'):
      occ = np.where(occ > 0, np.exp(np.log(np.where(occ > 0, occ, 1.0)) + m - log_p), 0.0)
    grad[:tb, b, :] = np.exp(lsm) - occ
  return loss, grad


def ctc_brute_force_loss(logits_tm: np.ndarray, label: Sequence[int], blank: int = None):
  """-log sum over every alignment that collapses to `label` (enumeration; tiny T only).
  Independent known-answer generator for the alpha-beta recursion above."""
  import itertools
  T, C = logits_tm.shape
  blank = C - 1 if blank is None else blank
  x = logits_tm.astype(np.float64)
  lsm = x - x.max(axis=1, keepdims=True)
  lsm = lsm - np.log(np.exp(lsm).sum(axis=1, keepdims=True))
  total = 0.0
  target = list(label)
  for path in itertools.product(range(C), repeat=T):
    out, prev = [], None
    for c in path:
      if c != prev and c != blank:
        out.append(c)
      prev = c
    if out == target:
      total += math.exp(sum(lsm[t, c] for t, c in enumerate(path)))
  return -math.log(total) if total > 0 else math.inf


# ----------------------------------------------------------------------------------------
# a12: greedy decode (reference: speech_model.py:113-115 -> tf.nn.ctc_greedy_decoder)
# ----------------------------------------------------------------------------------------
def ctc_greedy_decoder(logits_tm: np.ndarray, seq_len: Sequence[int], merge_repeated: bool = True,
                       blank: int = None):
  """Per frame t < seq_len[b]: argmax over RAW logits (first maximum wins); emit unless blank or
  (merge_repeated and equal to the previous frame's argmax); previous updates every frame.
  Returns ((indices int64 [N,2], values int64 [N], dense_shape int64 [2]), neg_sum_logits [B,1])."""
  T, B, C = logits_tm.shape
  blank = C - 1 if blank is None else blank
  idx, val = [], []
  neg_sum = np.zeros((B, 1), dtype=logits_tm.dtype)
  max_len = 0
  for b in range(B):
    prev = -1
    n = 0
    for t in range(int(seq_len[b])):
      row = logits_tm[t, b]
      c = int(np.argmax(row))
      neg_sum[b, 0] += -row[c]
      if c != blank and not (merge_repeated and c == prev):
        idx.append((b, n))
        val.append(c)
        n += 1
      prev = c
    max_len = max(max_len, n)
  return ((np.array(idx, dtype=np.int64).reshape(-1, 2), np.array(val, dtype=np.int64),
           np.array([B, max_len], dtype=np.int64)), neg_sum)


def extract_decoded_ids(indices: np.ndarray, values: np.ndarray):
  """evaluation.py:161-171 verbatim semantics, including the quirk that an utterance decoding to the
  empty string produces no entry (so later strings shift)."""
  ids, last = [], 0
  out = []
  for i, (batch_id, _char) in enumerate(indices):
    if batch_id > last:
      out.append(ids)
      ids = []
      last = batch_id
    ids.append(int(values[i]))
  out.append(ids)
  return out


# ----------------------------------------------------------------------------------------
# a10-a11: clip by global norm + TF1 Adam (reference: speech_model.py:77-82)
# ----------------------------------------------------------------------------------------
def clip_by_global_norm(grads: Sequence[np.ndarray], clip_norm: float):
  """tf.clip_by_global_norm: norm over ALL tensors; scale = clip * min(1/norm, 1/clip)."""
  acc = np.float64(0.0)
  for g in grads:
    acc += np.sum(np.square(g.astype(np.float64)))
  norm = np.sqrt(acc)
  with np.errstate(divide='ignore'):
    scale = clip_norm * min(1.0 / norm if norm > 0 else np.inf, 1.0 / clip_norm)
  return [g * g.dtype.type(scale) for g in grads], norm


def adam_tf1(params, grads, m, v, lr: float, step: int, beta1=0.9, beta2=0.999, eps=1e-3):
  """tf.train.AdamOptimizer(lr, epsilon=1e-3) (speech_model.py:77): step counts from 1;
  lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; theta -= lr_t * m / (sqrt(v) + eps) -- eps is added to the
  UNcorrected sqrt(v) (unlike torch.optim.Adam).  Updates in place, returns nothing."""
  for p, g, mi, vi in zip(params, grads, m, v):
    dt = p.dtype.type
    lr_t = dt(lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step))
    mi *= dt(beta1); mi += dt(1.0 - beta1) * g
    vi *= dt(beta2); vi += dt(1.0 - beta2) * (g * g)
    p -= lr_t * mi / (np.sqrt(vi) + dt(eps))


# ----------------------------------------------------------------------------------------
# a13: one model.step (reference: speech_model.py:197-235; training.py:63; evaluation.py:132-137)
# ----------------------------------------------------------------------------------------
def evaluate_step(inputs, seq_lengths, labels, weights, dtype=np.float32):
  """model.step(update=False, decode=True): avg_loss + greedy decode from ONE forward pass."""
  inputs = inputs.astype(dtype)
  weights = [(w.astype(dtype), b.astype(dtype)) for w, b in weights]
  logits = wav2letter_forward(inputs, weights)
  ctc_len = np.asarray(seq_lengths) // 2                                # speech_model.py:74
  loss, _ = ctc_loss_and_grad(logits, labels, ctc_len, dtype=dtype)
  decoded, neg_sum = ctc_greedy_decoder(logits, ctc_len)
  return {'logits': logits, 'loss': loss, 'avg_loss': loss.mean(dtype=dtype), 'decoded': decoded,
          'neg_sum_logits': neg_sum}


def train_step(inputs, seq_lengths, labels, weights, m, v, step, lr=1e-4, max_gradient_norm=5.0,
               dtype=np.float32):
  """model.step(update=True): forward, CTC, backward, clip_by_global_norm(5.0), Adam(eps=1e-3).
  `weights`, `m`, `v` are lists of (w, b) pairs and are updated in place.  Returns avg_loss and the
  unclipped gradients."""
  inputs = inputs.astype(dtype)
  logits, acts = wav2letter_forward(inputs, weights, keep_activations=True)
  ctc_len = np.asarray(seq_lengths) // 2
  loss, dlogits = ctc_loss_and_grad(logits, labels, ctc_len, dtype=dtype)
  B = inputs.shape[0]
  dlogits = dlogits / dtype(B)                                          # reduce_mean (speech_model.py:75)
  grads = wav2letter_backward(acts, weights, dlogits)
  flat_p = [t for pair in weights for t in pair]
  flat_g = [t.astype(dtype) for pair in grads for t in pair]
  flat_m = [t for pair in m for t in pair]
  flat_v = [t for pair in v for t in pair]
  clipped, norm = clip_by_global_norm(flat_g, max_gradient_norm)
  adam_tf1(flat_p, clipped, flat_m, flat_v, lr, step)
  return {'avg_loss': loss.mean(dtype=dtype), 'loss': loss, 'grads': grads, 'grad_norm': norm,
          'logits': logits}


# ----------------------------------------------------------------------------------------
# synthetic workload (BASELINE.md "Synthetic inputs")
# ----------------------------------------------------------------------------------------
def frames_for_seconds(seconds: float, sr: int = 16000, hop: int = 160) -> int:
  return 1 + int(sr * seconds) // hop


def synthetic_labels(rng: np.random.Generator, n_chars: int, ctc_len: int):
  """uniform ids in [0,27]; regenerated until len + adjacent repeats <= ctc_len (feasible for CTC)."""
  while True:
    lab = rng.integers(0, VOCAB_SIZE, size=n_chars)
    if n_chars + int(np.sum(lab[1:] == lab[:-1])) <= ctc_len:
      return lab.astype(np.int32)


def synthetic_batch(seed: int, batch: int, seconds, chars_per_second: int = 15, n_mels: int = 128):
  """N(0,1) mel inputs [B,T,128] f32 zero-padded to the batch max, lengths, labels."""
  rng = np.random.default_rng(seed)
  secs = [seconds] * batch if np.isscalar(seconds) else list(seconds)
  feats, labels = [], []
  for s in secs:
    t = frames_for_seconds(s)
    feats.append(rng.standard_normal((t, n_mels), dtype=np.float32))
    labels.append(synthetic_labels(rng, int(chars_per_second * s), t // 2))
  inputs, lengths = pad_batch(feats, n_mels)
  return inputs, lengths, labels
