"""W2LEngine -- device-side state and step sequencing of the Wav2Letter hot path.

What the reference does in one `sess.run` (speech_model.py:235) -- forward through 11 conv layers
(speech_model.py:275-295), CTC loss (:74-75), gradients (:78), clip_by_global_norm (:80), Adam (:77,:81), greedy
decode (:113) -- is sequenced here as calls into libspeecht_b200.so.  The engine owns:

  * ONE flat fp32 buffer each for parameters, gradients, Adam m and Adam v (the 22 tensors of the model at
    256-byte aligned offsets), so clip/Adam/allreduce are single streaming passes;
  * per-shape activation buffers [B,T,C] (NWC, the reference layout);
  * for the tensor-core precisions, a native plan object (csrc/w2l_plan.cu) that holds the bf16 operand planes and
    enqueues the whole forward/backward.

precision:
  'fp32'   -- exact fp32 on the CUDA cores (FFMA).  Strict-accuracy path.
  'bf16x3' -- tcgen05 tensor cores, every fp32 operand split into hi+lo bf16 planes, 3 products, fp32 accumulate in
              TMEM.  Meets the 1e-4 activation / loss parity gates (measured in tests/test_gpu_parity.py).
  'bf16x6' -- tcgen05, three bf16 planes per operand, the six products above 2^-24: fp32-equivalent operands, the
              remaining error is the tensor core's fp32 accumulation.  Twice the MMA work of bf16x3.
  'bf16'   -- tcgen05, single bf16 plane, fp32 accumulate (BASELINE configs 3-4).  Does NOT meet the 1e-4 gate.
"""
import math

import numpy as np
import torch

from . import ops
from ._lib import check, lib, ptr, stream_ptr

PRECISIONS = ('fp32', 'bf16x6', 'bf16x3', 'bf16')


class BatchMean:
  """tf.reduce_mean(cost) (speech_model.py:75) of the per-utterance losses, evaluated WHEN READ: the step itself
  launches no reduction kernel; .item() averages the [B] losses on the host in float32.

  early=True (train steps): the losses are copied to pinned host memory right after the CTC kernels and an event is
  recorded THERE, so .item() returns as soon as the loss exists -- while backward and Adam of the same step are still
  running -- and the caller's next step is enqueued without the GPU ever waiting for the host."""

  _free = []                                        # pinned [>= B] float32 buffers not owned by a live BatchMean

  def __init__(self, loss, early=False):
    self.loss = loss
    self._host = self._event = None
    if early and loss.is_cuda:
      n = loss.numel()
      for i, buf in enumerate(BatchMean._free):
        if buf.numel() >= n:
          self._host = BatchMean._free.pop(i)
          break
      if self._host is None:
        self._host = torch.empty((max(n, 64),), dtype=torch.float32).pin_memory()
      self._host[:n].copy_(loss.detach(), non_blocking=True)
      self._event = torch.cuda.Event()
      self._event.record()

  def item(self):
    if self._event is not None:
      self._event.synchronize()
      return float(self._host[:self.loss.numel()].numpy().mean(dtype=np.float32))
    return float(self.loss.detach().cpu().numpy().mean(dtype=np.float32))

  def __float__(self):
    return self.item()

  def tensor(self):
    """Device scalar (for collectives such as parallel.mean_scalar)."""
    return self.loss.mean()

  def __del__(self):
    try:
      if self._host is not None and self._event is not None and self._event.query() and len(BatchMean._free) < 16:
        BatchMean._free.append(self._host)
    except Exception:
      pass


def _on_engine_device(fn):
  """The native calls launch on torch's CURRENT device and stream: run the method with the engine's own device
  current, so an engine created for cuda:1 works while cuda:0 is the process default."""
  import functools

  @functools.wraps(fn)
  def wrapped(self, *args, **kwargs):
    with torch.cuda.device(self.device):
      return fn(self, *args, **kwargs)
  return wrapped


def layer_table(input_size=128, num_classes=29):
  """(filter_width, stride, cin, cout, relu) per layer -- reference speech_model.py:275-292."""
  layers = [(48, 2, input_size, 250, True)]
  layers += [(7, 1, 250, 250, True)] * 7
  layers += [(32, 1, 250, 2000, True), (1, 1, 2000, 2000, True), (1, 1, 2000, num_classes, False)]
  return layers


class ParamLayout:
  """Offsets (in floats) of filters / bias of every layer inside the flat buffers; 64-float (256 B) aligned."""

  ALIGN = 64

  def __init__(self, layers):
    self.layers = layers
    self.w_off, self.b_off = [], []
    off = 0
    for (k, _s, cin, cout, _r) in layers:
      self.w_off.append(off)
      off = self._up(off + k * cin * cout)
      self.b_off.append(off)
      off = self._up(off + cout)
    self.total = off
    self.n_params = sum(k * cin * cout + cout for (k, _s, cin, cout, _r) in layers)

  def _up(self, x):
    return (x + self.ALIGN - 1) // self.ALIGN * self.ALIGN

  def views(self, flat):
    out = []
    for i, (k, _s, cin, cout, _r) in enumerate(self.layers):
      w = flat[self.w_off[i]:self.w_off[i] + k * cin * cout].view(k, cin, cout)
      b = flat[self.b_off[i]:self.b_off[i] + cout]
      out.append((w, b))
    return out


class W2LEngine:

  def __init__(self, input_size=128, num_classes=29, device=None, precision='fp32', process_group=None):
    if not torch.cuda.is_available():
      raise RuntimeError('speecht_b200 needs a CUDA device: there is no CPU fallback')
    if precision not in PRECISIONS:
      raise ValueError('precision must be one of %s' % (PRECISIONS,))
    lib()                                                     # fail loudly if the native library is missing
    self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
    self.input_size, self.num_classes = input_size, num_classes
    self.precision = precision
    self.layers = layer_table(input_size, num_classes)
    self.layout = ParamLayout(self.layers)
    n = self.layout.total
    self.params = torch.zeros((n,), dtype=torch.float32, device=self.device)
    self.grads = torch.zeros((n,), dtype=torch.float32, device=self.device)
    self.adam_m = torch.zeros((n,), dtype=torch.float32, device=self.device)
    self.adam_v = torch.zeros((n,), dtype=torch.float32, device=self.device)
    self.weights = self.layout.views(self.params)
    self.weight_grads = self.layout.views(self.grads)
    self.global_step = 0
    self.process_group = process_group
    self.world_size = 1
    if process_group is not None:
      import torch.distributed as dist
      self.world_size = dist.get_world_size(process_group)
    self._normsq = torch.zeros((1,), dtype=torch.float64, device=self.device)
    self._bufs = {}
    self._plan = None
    self._weights_version = 0
    self.launches = 0                                         # native kernel launches issued (bench gpu_launches)
    self.record_kernel_times = False                          # bench.py: CUDA events around the conv kernels
    self.kernel_times = []                                    # (kernel name, algorithmic flops, start, stop)

  # ---------------------------------------------------------------- parameters
  @_on_engine_device
  def init_xavier(self, seed=0):
    """tf.contrib.layers.xavier_initializer on [K,Cin,Cout] + zero bias (speech_model.py:150-152)."""
    rng = np.random.default_rng(seed)
    for (k, _s, cin, cout, _r), (w, b) in zip(self.layers, self.weights):
      limit = math.sqrt(6.0 / (k * cin + k * cout))
      w.copy_(torch.from_numpy(rng.uniform(-limit, limit, size=(k, cin, cout)).astype(np.float32)))
      b.zero_()
    self.mark_weights_changed()

  @_on_engine_device
  def load_weights(self, weights):
    """weights: list of (filters [K,Cin,Cout], bias [Cout]) numpy arrays in the `export` .npy layout."""
    if len(weights) != len(self.layers):
      raise ValueError('expected %d layers, got %d' % (len(self.layers), len(weights)))
    for (w, b), (wn, bn) in zip(self.weights, weights):
      if tuple(wn.shape) != tuple(w.shape) or tuple(bn.shape) != tuple(b.shape):
        raise ValueError('weight shape mismatch: %s vs %s' % (wn.shape, tuple(w.shape)))
      w.copy_(torch.from_numpy(np.ascontiguousarray(wn, dtype=np.float32)))
      b.copy_(torch.from_numpy(np.ascontiguousarray(bn, dtype=np.float32)))
    self.mark_weights_changed()

  def export_weights(self):
    return [(w.cpu().numpy().copy(), b.cpu().numpy().copy()) for w, b in self.weights]

  def mark_weights_changed(self):
    self._weights_version += 1

  def reset_optimizer(self):
    self.adam_m.zero_()
    self.adam_v.zero_()
    self.global_step = 0

  # ---------------------------------------------------------------- per-kernel timing (bench.py roofline)
  class _Timed:
    def __init__(self, eng, name, flops):
      self.eng, self.name, self.flops = eng, name, flops

    def __enter__(self):
      if self.eng.record_kernel_times:
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e1 = torch.cuda.Event(enable_timing=True)
        self.e0.record()

    def __exit__(self, *exc):
      if self.eng.record_kernel_times:
        self.e1.record()
        self.eng.kernel_times.append((self.name, self.flops, self.e0, self.e1))
      return False

  def _timed(self, name, flops):
    return W2LEngine._Timed(self, name, flops)

  def start_kernel_timing(self):
    self.kernel_times = []
    self.record_kernel_times = True
    if self.precision != 'fp32':
      self._tc().set_timing(True)

  def stop_kernel_timing(self):
    """-> list of (kernel, tag, layer, flops, ms) for every timed launch since start_kernel_timing."""
    self.record_kernel_times = False
    torch.cuda.synchronize(self.device)
    out = [(name, 'conv', -1, flops, e0.elapsed_time(e1)) for name, flops, e0, e1 in self.kernel_times]
    if self.precision != 'fp32':
      out += self._tc().read_timings()
      self._tc().set_timing(False)
    self.kernel_times = []
    return out

  def roofline_report(self, timings, steps, peaks_path=None, region_seconds=None, train=True):
    """Roofline of the dominant kernel from CUDA events recorded around its launches INSIDE the timed steps:
    achieved = algorithmic FLOPs of those launches / their summed duration.  Denominator: MEASURED_PEAKS.json --
    the BURST cuBLAS bf16 figure when the timed region is shorter than a second (the power cap has not settled, the
    SM clock is still near its maximum), the SUSTAINED one otherwise; both fractions are reported.  Without the file:
    the B200_PROFILING.md fallback (1.59 PFLOP/s burst, ~1.4 sustained), labelled as such."""
    import json
    import os
    by, per_layer = {}, {}
    for name, tag, layer, flops, ms in timings:
      acc = by.setdefault(name, [0.0, 0.0, 0])
      acc[0] += flops; acc[1] += ms; acc[2] += 1
      if layer >= 0:
        pl = per_layer.setdefault('L%d.%s' % (layer, tag), [0.0, 0.0])
        pl[0] += flops; pl[1] += ms
    if not by:
      return None
    burst, sustained, src = 1590.0, 1400.0, 'fallback (B200_PROFILING.md: 1.59 PFLOP/s burst, ~1.4 sustained)'
    if peaks_path and os.path.exists(peaks_path):
      pk = json.load(open(peaks_path))
      burst = float(pk.get('bf16_tflops', burst))
      sustained = float(pk.get('bf16_tflops_sustained', sustained))
      src = 'measured (MEASURED_PEAKS.json: cuBLAS bf16 8192^3, burst best-of-10 / sustained 4 s back to back)'
    use_burst = region_seconds is None or region_seconds < 1.0
    peak = burst if use_burst else sustained
    name = max(by, key=lambda k: by[k][1])
    flops, ms, n = by[name]
    achieved = flops / (ms * 1e-3) / 1e12
    passes = {'fp32': None, 'bf16x6': 6, 'bf16x3': 3, 'bf16': 1}[self.precision]
    rep = {'kernel': name, 'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
           'frac': achieved / peak, 'traffic': None,
           'peak_kind': ('burst' if use_burst else 'sustained') + ' (timed region %.2f s)' % (region_seconds or 0.0),
           'peak_burst': burst, 'peak_sustained': sustained, 'frac_burst': achieved / burst,
           'frac_sustained': achieved / sustained, 'peak_source': src,
           'launches_per_step': n / steps, 'avg_launch_ms': ms / n, 'ms_per_step': ms / steps,
           'kernels': {k: {'tflops': v[0] / (v[1] * 1e-3) / 1e12, 'ms_per_step': v[1] / steps,
                           'launches_per_step': v[2] / steps} for k, v in by.items()},
           'layers_ms_per_step': {k: round(v[1] / steps, 4) for k, v in sorted(per_layer.items())}}
    if passes:
      rep['mma_passes'] = passes
      rep['tensor_pipe_frac'] = passes * achieved / peak
      rep['note'] = ('achieved = algorithmic FLOPs (unpadded, 1 pass) / event time; in %s mode the tensor pipe '
                     'executes %d MMAs per algorithmic MAC, so the ceiling of frac is 1/%d and tensor_pipe_frac = '
                     '%d*frac' % (self.precision, passes, passes, passes)) if passes > 1 else \
                    'plain bf16: one MMA pass'
    else:
      rep['note'] = 'exact-fp32 FFMA path: compared with the bf16 tensor peak for reference only'
    return rep

  # ---------------------------------------------------------------- buffers
  def _buf(self, name, shape, dtype=torch.float32):
    key = (name, tuple(shape), dtype)
    t = self._bufs.get(key)
    if t is None:
      # drop same-name buffers of other shapes so ragged batches do not accumulate memory
      for k in [k for k in self._bufs if k[0] == name]:
        del self._bufs[k]
      t = torch.empty(shape, dtype=dtype, device=self.device)
      self._bufs[key] = t
    return t

  def conv_flops_forward(self, B, T):
    """Algorithmic forward FLOPs of the stack: sum_l 2*K*Cin*Cout*T_out*B with the true (unpadded) channel counts."""
    total, t = 0.0, T
    for (k, s, cin, cout, _r) in self.layers:
      t = -(-t // s)
      total += 2.0 * k * cin * cout * t * B
    return total

  def _out_lengths(self, T):
    outs = []
    t = T
    for (k, s, _ci, _co, _r) in self.layers:
      t = -(-t // s)
      outs.append(t)
    return outs

  # ---------------------------------------------------------------- forward
  @_on_engine_device
  def forward(self, inputs, keep_activations=False):
    """inputs [B,T,input_size] f32 CUDA -> logits [T',B,num_classes] (time-major VIEW, speech_model.py:295)."""
    if inputs.dim() != 3 or inputs.shape[2] != self.input_size:
      raise ValueError('inputs must be [batch, time, %d]' % self.input_size)
    inputs = inputs.contiguous()
    if self.precision != 'fp32':
      return self._tc().forward(inputs, keep_activations)
    B, T, _ = inputs.shape
    acts = [inputs]
    x = inputs
    for li, ((k, s, cin, cout, relu), (w, b)) in enumerate(zip(self.layers, self.weights)):
      to = -(-x.shape[1] // s)
      y = self._buf('act%d' % (li + 1), (B, to, cout))
      with self._timed('conv_gemm_f32_kernel', 2.0 * k * cin * cout * to * B):
        ops.conv1d(x, w, b, stride=s, relu=relu, out=y)
      self.launches += 1
      x = y
      if keep_activations:
        acts.append(y)
    self._acts = acts if keep_activations else None
    self._logits_bm = x                                       # [B,T',C] batch-major storage
    return x.transpose(0, 1)

  # ---------------------------------------------------------------- backward (fp32 path)
  def _backward_fp32(self, dlogits_bm):
    acts = self._acts
    dy = dlogits_bm
    for li in reversed(range(len(self.layers))):
      (k, s, cin, cout, relu) = self.layers[li]
      dw, db = self.weight_grads[li]
      y_act = acts[li + 1] if relu else None
      flops = 2.0 * k * cin * cout * dy.shape[0] * dy.shape[1]
      with self._timed('conv_gemm_f32_kernel', flops):
        ops.conv1d_backprop_filter(acts[li], dy, k, stride=s, y_act=y_act, dw=dw, db=db)
      self.launches += 2
      if li > 0:
        dx = self._buf('dx%d' % (li & 1), tuple(acts[li].shape))
        with self._timed('conv_gemm_f32_kernel', flops):
          ops.conv1d_backprop_input(dy, self.weights[li][0], tuple(acts[li].shape), stride=s, y_act=y_act, out=dx)
        self.launches += 1
        dy = dx

  # ---------------------------------------------------------------- steps
  def _tc(self):
    if self._plan is None:
      from .tc_plan import TCPlan
      self._plan = TCPlan(self)
    return self._plan

  @staticmethod
  def length_buckets(sequence_lengths, buckets):
    """Utterance indices of `buckets` groups of similar length (sorted by length, split evenly): evaluating each
    group padded to its OWN maximum skips most of the padding frames of a ragged batch."""
    order = np.argsort(np.asarray(sequence_lengths), kind='stable')
    return [g for g in np.array_split(order, max(1, min(int(buckets), len(order)))) if len(g)]

  @_on_engine_device
  def evaluate_step(self, inputs, sequence_lengths, labels=None, decode=True, buckets=1, defer_decode=False,
                    fresh_decode=False):
    """model.step(update=False, decode=True): returns dict(loss [B] tensor|None, avg_loss, decoded, logits).
    defer_decode: 'decoded' is an ops.PendingDecode (kernel enqueued, nothing read back yet); pass the dict to
    finish_evaluate() once the host has nothing better to do than to wait (fresh_decode: the decoded rows get their own
    buffers, so that a second step may be enqueued before this one is read).

    buckets > 1 (optional; the reference always pads the whole batch to its longest utterance, speech_input.py:38-43):
    the batch is evaluated as that many length-sorted groups, each padded to its own maximum.  Results come back in
    the original utterance order; `logits` is then None.  Because the network does not mask padding
    (speech_model.py:128-181), the last ~20 logit frames of an utterance depend on how much padding follows it, so a
    bucketed evaluation is NOT bit-identical to the reference batch -- parity is checked unbucketed."""
    if buckets > 1 and inputs.shape[0] > 1:
      return self._evaluate_bucketed(inputs, sequence_lengths, labels, decode, buckets)
    logits = self.forward(inputs, keep_activations=False)
    ctc_len = np.asarray(sequence_lengths, dtype=np.int32) // 2      # speech_model.py:74,114
    out = {'logits': logits, 'loss': None, 'avg_loss': None, 'decoded': None}
    d_seq = ctc_len
    if labels is not None:
      batch = labels if isinstance(labels, ops.CTCBatch) else ops.CTCBatch(labels, ctc_len, logits.shape[0],
                                                                           self.num_classes, self.device)
      loss, _ = ops.ctc_loss(batch, logits, want_grad=False)
      d_seq = batch.seq_len                                     # already on the device (pinned, asynchronous upload)
      self.launches += 2
      out['loss'] = loss
      out['avg_loss'] = BatchMean(loss, early=defer_decode)     # deferred: the [B] losses start their trip to the host now
    if decode:
      if defer_decode:
        out['decoded'] = ops.PendingDecode(logits, d_seq, fresh=fresh_decode)
      else:
        out['decoded'], out['neg_sum_logits'] = ops.ctc_greedy_decoder(logits, d_seq)
      self.launches += 1
    return out

  @staticmethod
  def finish_evaluate(out):
    """Completes an evaluate_step(defer_decode=True) result in place (the blocking device -> host read)."""
    if isinstance(out.get('decoded'), ops.PendingDecode):
      out['decoded'], out['neg_sum_logits'] = out['decoded'].finish()
    return out

  def _evaluate_bucketed(self, inputs, sequence_lengths, labels, decode, buckets):
    lengths = np.asarray(sequence_lengths, dtype=np.int32)
    B = int(lengths.shape[0])
    rows = None
    if labels is not None:
      flat, offsets = ops.flatten_labels(labels)
      rows = [flat[offsets[b]:offsets[b + 1]] for b in range(B)]
    loss = torch.zeros((B,), dtype=torch.float32, device=self.device) if labels is not None else None
    pending = []
    # every group's kernels are enqueued first (the next group's forward reuses the arena, which stream order makes
    # safe); results cross to the host once, at the end
    for group in self.length_buckets(lengths, buckets):
      idx = ops.upload_int32_async(group, self.device).long()       # pinned + asynchronous: no stream synchronisation
      t_max = max(int(lengths[group].max()), 2)
      x = inputs.index_select(0, idx)[:, :t_max].contiguous()
      logits = self.forward(x, keep_activations=False)
      ctc_len = lengths[group] // 2
      batch = ops.CTCBatch([rows[b] for b in group] if rows is not None else [[] for _ in group], ctc_len,
                           logits.shape[0], self.num_classes, self.device, validate=rows is not None)
      if loss is not None:
        group_loss, _ = ops.ctc_loss(batch, logits, want_grad=False)
        loss.index_copy_(0, idx, group_loss)
        self.launches += 2
      if decode:
        pending.append((group,) + ops.ctc_greedy_decode_device(logits, batch.seq_len, fresh=True))
        self.launches += 1
    out = {'logits': None, 'loss': loss, 'avg_loss': BatchMean(loss) if loss is not None else None, 'decoded': None}
    if decode:
      decoded_rows = [None] * B
      neg = np.zeros((B, 1), dtype=np.float32)
      for group, values, counts, group_neg in pending:
        sp = ops.sparse_from_rows(values, counts)
        starts = np.searchsorted(sp.indices[:, 0], np.arange(len(group) + 1))
        for j, b in enumerate(group):
          decoded_rows[b] = sp.values[starts[j]:starts[j + 1]]
        neg[group, 0] = group_neg.cpu().numpy()
      counts = np.array([len(r) for r in decoded_rows], dtype=np.int64)
      n = int(counts.sum())
      row_idx = np.repeat(np.arange(B, dtype=np.int64), counts)
      col_idx = np.arange(n, dtype=np.int64) - np.repeat(np.cumsum(counts) - counts, counts)
      values = np.concatenate(decoded_rows).astype(np.int64) if n else np.zeros((0,), dtype=np.int64)
      out['decoded'] = [ops.SparseTensorValue(np.stack([row_idx, col_idx], axis=1).reshape(n, 2), values,
                                              np.array([B, int(counts.max()) if B else 0], dtype=np.int64))]
      out['neg_sum_logits'] = neg
    return out

  @_on_engine_device
  def evaluate_step_device(self, inputs, ctc_batch, merge_repeated=True):
    """The kernels of model.step(update=False, decode=True) with every result left on the device: forward, CTC loss
    (no gradient) and greedy decode.  ctc_batch: ops.CTCBatch of this batch (labels + ctc lengths already uploaded).
    -> (loss [B], label rows [B, T'] int32, counts [B] int32, neg_sum_logits [B])."""
    logits = self.forward(inputs, keep_activations=False)
    loss, _ = ops.ctc_loss(ctc_batch, logits, want_grad=False)
    values, counts, neg = ops.ctc_greedy_decode_device(logits, ctc_batch.seq_len, merge_repeated)
    self.launches += 3
    return loss, values, counts, neg

  @_on_engine_device
  def train_step(self, inputs, sequence_lengths, labels, learning_rate, max_gradient_norm=5.0, decode=False):
    """model.step(update=True): forward, CTC, backward, [allreduce], clip_by_global_norm, Adam.
    Returns dict(avg_loss device scalar (LOCAL batch mean), loss [B], decoded|None)."""
    if self.precision != 'fp32':
      return self._tc().train_step(inputs, sequence_lengths, labels, learning_rate, max_gradient_norm, decode)
    B, T = inputs.shape[0], inputs.shape[1]
    ctc_len = np.asarray(sequence_lengths, dtype=np.int32) // 2
    batch = ops.CTCBatch(labels, ctc_len, -(-T // 2), self.num_classes, self.device)
    logits = self.forward(inputs, keep_activations=True)
    scale = 1.0 / (B * self.world_size)                                # tf.reduce_mean folded into the gradient
    loss, dlogits = ops.ctc_loss(batch, logits, want_grad=True, grad_scale=scale)
    self.launches += 3
    out = {'loss': loss, 'avg_loss': BatchMean(loss, early=True), 'decoded': None, 'logits': logits}
    if decode:
      out['decoded'], out['neg_sum_logits'] = ops.ctc_greedy_decoder(logits, ctc_len)
      self.launches += 1
    self._backward_fp32(dlogits.transpose(0, 1))                        # batch-major storage, contiguous
    self.apply_gradients(learning_rate, max_gradient_norm)
    return out

  def allreduce_gradients(self, async_ranges=None):
    """NCCL sum-allreduce of the flat gradient buffer (the only collective on the path).
    async_ranges: list of (start, stop) float ranges -> returns work handles instead of blocking."""
    from . import parallel
    if self.world_size == 1:
      return []
    handles = parallel.allreduce_flat(self.grads, self.process_group, async_ranges)
    if async_ranges is None:
      for h in handles:
        h.wait()
      return []
    return handles

  @_on_engine_device
  def apply_gradients(self, learning_rate, max_gradient_norm=5.0, reduced=False):
    """[allreduce] + tf.clip_by_global_norm + Adam(eps=1e-3) on the flat buffers (speech_model.py:77-82)."""
    if not reduced:
      self.allreduce_gradients()
    self.global_step += 1
    ops.global_norm_sq(self.grads, self._normsq, zero=True)
    ops.clip_adam(self.params, self.grads, self.adam_m, self.adam_v, self.global_step, learning_rate,
                  max_norm=max_gradient_norm, normsq=self._normsq)
    self.launches += 2
    self.mark_weights_changed()

  @_on_engine_device
  def grad_norm(self):
    return float(torch.sqrt(ops.global_norm_sq(self.grads)).item())
