"""Operator-level host mirror: one function per TensorFlow-1 / librosa call on the reference's hot path.

Each function takes CUDA torch tensors (torch is only the allocator / stream owner), calls the C ABI in
libspeecht_b200.so through ctypes and returns torch tensors.  Names and argument meaning follow the reference's
call sites (file:line in the docstrings); errors are raised as Python exceptions like the reference's ops do.
There is no CPU implementation here: without the CUDA library or a GPU these raise.
"""
from collections import namedtuple

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

SparseTensorValue = namedtuple('SparseTensorValue', ['indices', 'values', 'dense_shape'])


def _require_cuda(*tensors):
  for t in tensors:
    if t is not None and not t.is_cuda:
      raise ValueError('speecht_b200 ops need CUDA tensors (no CPU fallback)')


def same_padding(t_in, k, stride):
  """TF 'SAME' geometry: (out, pad_left, pad_right)."""
  out = -(-t_in // stride)
  total = max((out - 1) * stride + k - t_in, 0)
  return out, total // 2, total - total // 2


# ------------------------------------------------------------------------------------------------
# conv (exact fp32 path)
# ------------------------------------------------------------------------------------------------
def conv1d(value, filters, bias=None, stride=1, relu=False, out=None):
  """tf.nn.conv1d(value, filters, stride, 'SAME') + tf.nn.bias_add + tf.nn.relu (speech_model.py:155,173,177).
  value [B,T,Cin] f32, filters [K,Cin,Cout] f32 -> [B,ceil(T/stride),Cout]."""
  _require_cuda(value, filters, bias)
  value, filters = value.contiguous(), filters.contiguous()
  B, T, Cin = value.shape
  K, Cin2, Cout = filters.shape
  if Cin2 != Cin:
    raise ValueError('filter input channels %d != value channels %d' % (Cin2, Cin))
  To = -(-T // stride)
  if out is None:
    out = torch.empty((B, To, Cout), device=value.device, dtype=torch.float32)
  check(lib().st_conv1d_fwd_f32(ptr(value), ptr(filters), ptr(bias), ptr(out), B, T, Cin, Cout, K, stride,
                                int(bool(relu)), stream_ptr()))
  return out


def conv1d_backprop_input(dy, filters, input_shape, stride=1, y_act=None, out=None):
  """Gradient of conv1d wrt its input; y_act (the layer's post-ReLU output) fuses the ReLU backward mask."""
  _require_cuda(dy, filters, y_act)
  dy, filters = dy.contiguous(), filters.contiguous()
  B, T, Cin = input_shape
  K, _, Cout = filters.shape
  if out is None:
    out = torch.empty((B, T, Cin), device=dy.device, dtype=torch.float32)
  check(lib().st_conv1d_bwd_data_f32(ptr(dy), ptr(y_act), ptr(filters), ptr(out), B, T, Cin, Cout, K, stride,
                                     stream_ptr()))
  return out


def conv1d_backprop_filter(value, dy, filter_width, stride=1, y_act=None, dw=None, db=None):
  """Gradient of conv1d+bias wrt filters [K,Cin,Cout] and bias [Cout]."""
  _require_cuda(value, dy, y_act)
  value, dy = value.contiguous(), dy.contiguous()
  B, T, Cin = value.shape
  Cout = dy.shape[2]
  if dw is None:
    dw = torch.empty((filter_width, Cin, Cout), device=value.device, dtype=torch.float32)
  if db is None:
    db = torch.empty((Cout,), device=value.device, dtype=torch.float32)
  check(lib().st_conv1d_bwd_filter_f32(ptr(value), ptr(dy), ptr(y_act), ptr(dw), ptr(db), B, T, Cin, Cout,
                                       filter_width, stride, stream_ptr()))
  return dw, db


# ------------------------------------------------------------------------------------------------
# CTC
# ------------------------------------------------------------------------------------------------
def flatten_labels(labels):
  """list of int sequences (or a SparseTensorValue as speech_input.py:48-69 builds) -> (flat int32, offsets int32)."""
  if isinstance(labels, SparseTensorValue) or (hasattr(labels, 'indices') and hasattr(labels, 'dense_shape')):
    B = int(labels.dense_shape[0])
    rows = [[] for _ in range(B)]
    for (b, _j), v in zip(np.asarray(labels.indices), np.asarray(labels.values)):
      rows[int(b)].append(int(v))
    labels = rows
  lens = [len(l) for l in labels]
  offsets = np.zeros((len(labels) + 1,), dtype=np.int32)
  offsets[1:] = np.cumsum(lens)
  flat = np.concatenate([np.asarray(l, dtype=np.int32).reshape(-1) for l in labels]) if sum(lens) else \
    np.zeros((0,), dtype=np.int32)
  return np.ascontiguousarray(flat, dtype=np.int32), offsets


_ctc_ws_cache = {}


def _workspace(key, nbytes, device):
  buf = _ctc_ws_cache.get(key)
  if buf is None or buf.numel() < nbytes or buf.device != device:
    buf = torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=device)
    _ctc_ws_cache[key] = buf
  return buf


class _PinnedPool:
  """Page-locked staging buffers for the per-step label upload.  cudaHostAlloc costs ~0.1 ms, so buffers are
  recycled -- but only once the asynchronous copy that last read them has finished (its event has fired), because
  a caller that does not synchronise every step would otherwise overwrite labels still on their way to the device."""

  def __init__(self):
    self.items = []                                   # [buffer, event or None]

  def acquire(self, n):
    for item in self.items:
      buf, ev = item
      if buf.numel() >= n and (ev is None or ev.query()):
        item[1] = None
        return item
    item = [torch.empty((max(n, 1024),), dtype=torch.int32).pin_memory(), None]
    if len(self.items) < 64:
      self.items.append(item)
    return item


_pinned_pool = _PinnedPool()


def upload_int32_async(values, device):
  """Small int32 host array -> device tensor through the pinned pool, WITHOUT blocking the host: a plain
  torch.from_numpy(x).to(device) of pageable memory synchronises the stream, i.e. waits for every kernel enqueued
  before it (a whole forward pass when it sits between the forward and the decode kernel)."""
  arr = np.ascontiguousarray(np.asarray(values, dtype=np.int32)).reshape(-1)
  n = max(int(arr.size), 1)
  item = _pinned_pool.acquire(n)
  host = item[0][:n]
  host.numpy()[:arr.size] = arr
  dev = host.to(device, non_blocking=True)
  ev = torch.cuda.Event()
  ev.record()
  item[1] = ev
  return dev[:arr.size]


class CTCBatch:
  """Device-side label / length tensors of one batch, uploaded asynchronously from pinned memory so that no host
  synchronisation sits between the forward pass and the loss kernels."""

  def __init__(self, labels, sequence_length, T, C, device, validate=True):
    flat, offsets = flatten_labels(labels)
    seq_host = np.ascontiguousarray(np.asarray(sequence_length.cpu() if torch.is_tensor(sequence_length)
                                               else sequence_length, dtype=np.int32))
    self.B = int(seq_host.shape[0])
    if validate:
      check(lib().st_ctc_validate_labels_host(flat.ctypes.data, offsets.ctypes.data, seq_host.ctypes.data, self.B, T,
                                              C - 1))
    self.max_len = int(np.max(np.diff(offsets))) if self.B else 0
    n_lab = max(int(flat.size), 1)
    n = n_lab + 2 * self.B + 1
    # one pinned staging buffer, one async copy: [labels | offsets | seq_len]
    item = _pinned_pool.acquire(n)
    host = item[0][:n]
    view = host.numpy()
    view[:flat.size] = flat
    view[n_lab:n_lab + self.B + 1] = offsets
    view[n_lab + self.B + 1:] = seq_host
    self._host = host
    dev = host.to(device, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    item[1] = ev
    self.labels = dev[:n_lab]
    self.offsets = dev[n_lab:n_lab + self.B + 1]
    self.seq_len = dev[n_lab + self.B + 1:]
    self.T, self.C = T, C


def ctc_loss(labels, logits, sequence_length=None, want_grad=True, grad_scale=1.0, grad_planes=None, time_major=True,
             validate=True):
  """tf.nn.ctc_loss(labels, logits, sequence_length) (speech_model.py:74) and its gradient.

  logits: [T,B,C] (time_major, may be a transposed VIEW of a [B,T,C] buffer -- strides are honoured, no copy).
  labels: list of int lists, the sparse triple of speech_input.py:48-69, or a prepared CTCBatch.  Blank = C-1.
  Returns (loss [B], grad like logits or None).  Raises CTCLabelError like TF's InvalidArgumentError when a label
  does not fit its sequence."""
  _require_cuda(logits)
  if not time_major:
    logits = logits.transpose(0, 1)
  T, B, C = logits.shape
  if logits.stride(2) != 1 or logits.dtype != torch.float32:
    raise ValueError('logits must be float32 with unit class stride')
  dev = logits.device
  batch = labels if isinstance(labels, CTCBatch) else CTCBatch(labels, sequence_length, T, C, dev, validate)
  if batch.B != B or batch.T != T:
    raise ValueError('prepared CTC batch does not match the logits shape')
  loss = torch.empty((B,), dtype=torch.float32, device=dev)
  status = torch.empty((B,), dtype=torch.int32, device=dev)
  grad = None
  if want_grad:
    grad = torch.empty_strided(logits.shape, logits.stride(), dtype=torch.float32, device=dev)
  nbytes = lib().st_ctc_workspace_bytes(T, B, C, batch.max_len)
  ws = _workspace(('ctc', dev), nbytes, dev)
  n_planes, c_pad = (0, 0) if grad_planes is None else (grad_planes.shape[0], grad_planes.shape[-1])
  check(lib().st_ctc_loss(ptr(logits), logits.stride(0), logits.stride(1), T, B, C, ptr(batch.labels),
                          ptr(batch.offsets), batch.max_len, ptr(batch.seq_len), C - 1, ptr(loss), ptr(grad),
                          float(grad_scale), ptr(grad_planes), n_planes, c_pad, ptr(status), ptr(ws), ws.numel(),
                          stream_ptr()))
  return loss, grad


_decode_bufs = {}


def ctc_greedy_decode_device(logits, d_seq, merge_repeated=True, fresh=False):
  """The decode kernel alone: logits [T,B,C] (strided view allowed), d_seq int32 [B] on the device ->
  (label rows [B, max(T,1)] int32, counts [B] int32, neg_sum_logits [B] f32), all on the device.  The buffers are
  reused per shape (copy what must survive the next call) unless fresh=True."""
  _require_cuda(logits)
  T, B, C = logits.shape
  if logits.stride(2) != 1 or logits.dtype != torch.float32:
    raise ValueError('logits must be float32 with unit class stride')
  dev = logits.device
  key = (T, B, dev)
  buf = None if fresh else _decode_bufs.get(key)
  if buf is None:
    buf = (torch.empty((B, max(T, 1)), dtype=torch.int32, device=dev),
           torch.empty((B,), dtype=torch.int32, device=dev), torch.empty((B,), dtype=torch.float32, device=dev))
    if not fresh:
      if len(_decode_bufs) > 16:
        _decode_bufs.clear()
      _decode_bufs[key] = buf
  values, counts, neg = buf
  check(lib().st_ctc_greedy_decode(ptr(logits), logits.stride(0), logits.stride(1), T, B, C, ptr(d_seq), C - 1,
                                   int(bool(merge_repeated)), ptr(values), ptr(counts), ptr(neg), stream_ptr()))
  return values, counts, neg


def ctc_greedy_decoder(logits, sequence_length, merge_repeated=True):
  """tf.nn.ctc_greedy_decoder(logits [T,B,C], sequence_length, merge_repeated) (speech_model.py:113-115).
  Returns ([SparseTensorValue(indices int64 [N,2], values int64 [N], dense_shape int64 [2])], neg_sum_logits [B,1])
  as numpy, like sess.run would hand it to evaluation.py:144,161-171."""
  _require_cuda(logits)
  T, B, C = logits.shape
  dev = logits.device
  if torch.is_tensor(sequence_length):
    d_seq = sequence_length.to(device=dev, dtype=torch.int32)
  else:
    d_seq = upload_int32_async(sequence_length, dev)
  values, counts, neg = ctc_greedy_decode_device(logits, d_seq, merge_repeated)
  return [sparse_from_rows(values, counts)], neg.cpu().numpy().reshape(B, 1)


class PendingDecode:
  """A greedy decode whose kernel is enqueued and whose results are on their way to pinned host memory: finish() waits
  for THAT copy (an event recorded right behind it, not the whole stream) and builds what ctc_greedy_decoder returns.
  Lets the caller put host work -- fetching and uploading the next batch, even enqueueing the next step's kernels --
  between the launch and the read (speech_model.SpeechModel.step, evaluate path)."""

  _free = []                                        # pinned int32 staging buffers not owned by a live PendingDecode

  def __init__(self, logits, sequence_length, merge_repeated=True, fresh=False):
    _require_cuda(logits)
    dev = logits.device
    if torch.is_tensor(sequence_length):
      d_seq = sequence_length.to(device=dev, dtype=torch.int32)
    else:
      d_seq = upload_int32_async(sequence_length, dev)
    self.B = B = logits.shape[1]
    values, counts, neg = ctc_greedy_decode_device(logits, d_seq, merge_repeated, fresh=fresh)
    self.width = W = values.shape[1]
    n = B * W + 2 * B                               # [label rows | counts | neg_sum_logits bits]
    self._host = None
    for i, buf in enumerate(PendingDecode._free):
      if buf.numel() >= n:
        self._host = PendingDecode._free.pop(i)
        break
    if self._host is None:
      self._host = torch.empty((n,), dtype=torch.int32).pin_memory()
    h = self._host
    h[:B * W].view(B, W).copy_(values, non_blocking=True)
    h[B * W:B * W + B].copy_(counts, non_blocking=True)
    h[B * W + B:n].view(torch.float32).copy_(neg, non_blocking=True)
    self._event = torch.cuda.Event()
    self._event.record()
    self._keep = (values, counts, neg)              # alive until the copies have run

  def finish(self):
    self._event.synchronize()
    B, W = self.B, self.width
    h = self._host.numpy()
    out = [sparse_from_host_rows(h[:B * W].reshape(B, W), h[B * W:B * W + B])], \
        h[B * W + B:B * W + 2 * B].view(np.float32).reshape(B, 1).copy()
    if len(PendingDecode._free) < 8:
      PendingDecode._free.append(self._host)
    self._host = self._keep = None
    return out


def sparse_from_host_rows(values_h, counts_h):
  """Host-side half of sparse_from_rows: [B,T] int32 rows + counts (numpy) -> SparseTensorValue (row-major order)."""
  counts_h = counts_h.astype(np.int64)
  B = counts_h.shape[0]
  width = int(counts_h.max()) if B else 0
  n = int(counts_h.sum())
  rows = np.repeat(np.arange(B, dtype=np.int64), counts_h)
  starts = np.cumsum(counts_h) - counts_h
  cols = np.arange(n, dtype=np.int64) - np.repeat(starts, counts_h)
  indices = np.stack([rows, cols], axis=1).reshape(n, 2)
  vals = values_h[rows, cols].astype(np.int64) if n else np.zeros((0,), dtype=np.int64)
  return SparseTensorValue(indices, vals, np.array([B, width], dtype=np.int64))


def sparse_from_rows(values, counts):
  """[B,T] int32 rows + counts -> the SparseTensor triple TF returns (row-major order).  Only the counts and the
  first max(count) columns of the rows cross to the host."""
  counts_h = counts.cpu().numpy().astype(np.int64)
  B = counts_h.shape[0]
  width = int(counts_h.max()) if B else 0
  values_h = values[:, :max(width, 1)].cpu().numpy()
  n = int(counts_h.sum())
  rows = np.repeat(np.arange(B, dtype=np.int64), counts_h)
  starts = np.cumsum(counts_h) - counts_h
  cols = np.arange(n, dtype=np.int64) - np.repeat(starts, counts_h)
  indices = np.stack([rows, cols], axis=1).reshape(n, 2)
  vals = values_h[rows, cols].astype(np.int64) if n else np.zeros((0,), dtype=np.int64)
  shape = np.array([B, width], dtype=np.int64)
  return SparseTensorValue(indices, vals, shape)


# ------------------------------------------------------------------------------------------------
# optimiser
# ------------------------------------------------------------------------------------------------
def global_norm_sq(flat_grads, accum=None, zero=False):
  """sum(g^2) over the flat gradient buffer (first half of tf.clip_by_global_norm, speech_model.py:80).
  accum: device double that the sum is ADDED to (zero=True clears it first, on the stream); None: a fresh one."""
  _require_cuda(flat_grads)
  if accum is None:
    accum, zero = torch.empty((1,), dtype=torch.float64, device=flat_grads.device), True
  check(lib().st_sumsq(ptr(flat_grads), flat_grads.numel(), ptr(accum), int(bool(zero)), stream_ptr()))
  return accum


def clip_adam(params, grads, m, v, step, lr, max_norm=5.0, beta1=0.9, beta2=0.999, eps=1e-3, normsq=None,
              grad_prescale=1.0):
  """tf.clip_by_global_norm(grads, max_norm) + AdamOptimizer(lr, epsilon=1e-3).apply_gradients
  (speech_model.py:77-82) on flat buffers, in place.  `normsq` = device double from global_norm_sq (None: no clip)."""
  _require_cuda(params, grads, m, v)
  check(lib().st_clip_adam(ptr(params), ptr(grads), ptr(m), ptr(v), params.numel(), float(lr), float(beta1),
                           float(beta2), float(eps), int(step), float(max_norm), ptr(normsq), float(grad_prescale),
                           stream_ptr()))


# ------------------------------------------------------------------------------------------------
# features
# ------------------------------------------------------------------------------------------------
def _hz_to_mel(f):
  f = np.asarray(f, dtype=np.float64)
  f_sp = 200.0 / 3
  min_log_hz = 1000.0
  logstep = np.log(6.4) / 27.0
  return np.where(f >= min_log_hz, min_log_hz / f_sp + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
  m = np.asarray(m, dtype=np.float64)
  f_sp = 200.0 / 3
  min_log_hz = 1000.0
  min_log_mel = min_log_hz / f_sp
  logstep = np.log(6.4) / 27.0
  return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft=512, n_mels=128):
  """librosa.filters.mel(sr, n_fft, n_mels) defaults: Slaney scale, fmin 0, fmax sr/2, area-normalised."""
  n_bins = 1 + n_fft // 2
  fft_f = np.linspace(0.0, sr / 2.0, n_bins)
  mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
  fdiff = np.diff(mel_f)
  ramps = mel_f[:, None] - fft_f[None, :]
  w = np.maximum(0.0, np.minimum(-ramps[:-2] / fdiff[:-1, None], ramps[2:] / fdiff[1:, None]))
  w *= (2.0 / (mel_f[2:] - mel_f[:-2]))[:, None]
  return w.astype(np.float32)


_mel_cache = {}


def power_spectrogram(wav, n_samples, samplerate, n_mels=128, n_fft=512, hop_length=160):
  """Batched calc_power_spectrogram (preprocessing.py:36-58) on the GPU.
  wav [B, max_samples] f32 CUDA (rows zero-padded), n_samples int list -> (features [B,T_max,n_mels], frames [B])."""
  _require_cuda(wav)
  wav = wav.contiguous()
  B, max_samples = wav.shape
  n_host = np.ascontiguousarray(np.asarray(n_samples, dtype=np.int32))
  if n_host.min() <= n_fft // 2:
    raise ValueError('audio shorter than n_fft/2 cannot be reflect-padded')
  dev = wav.device
  key = (float(samplerate), n_fft, n_mels, dev)
  basis = _mel_cache.get(key)
  if basis is None:
    basis = torch.from_numpy(mel_filterbank(samplerate, n_fft, n_mels)).to(dev)
    _mel_cache[key] = basis
  T_max = 1 + int(n_host.max()) // hop_length
  out = torch.empty((B, T_max, n_mels), dtype=torch.float32, device=dev)
  frames = torch.empty((B,), dtype=torch.int32, device=dev)
  d_n = torch.from_numpy(n_host).to(dev)
  nbytes = lib().st_melspec_workspace_bytes(B, max_samples, n_fft, hop_length, n_mels)
  ws = _workspace(('mel', dev), nbytes, dev)
  check(lib().st_melspec(ptr(wav), wav.stride(0), ptr(d_n), B, int(n_host.max()), ptr(basis), n_fft, hop_length,
                         n_mels, ptr(out), T_max, ptr(frames), ptr(ws), ws.numel(), stream_ptr()))
  return out, frames
