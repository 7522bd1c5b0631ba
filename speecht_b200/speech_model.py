"""SpeechModel / Wav2LetterModel / create_default_model -- mirror of reference speecht/speech_model.py.

Same classes, method names, argument names and defaults; `step` returns its results in the reference's fixed order
(avg_loss, decoded, labels, update, summary -- speech_model.py:216-229).  What was a TensorFlow-1 graph executed by
`sess.run` is a W2LEngine (engine.py) calling hand-written sm_100a kernels; `sess` is a lightweight Session object
(device + stream) so caller code such as training.py:57-90 / evaluation.py:126-158 reads the same.

Out of scope, and raising instead of silently differing: beam search with the KenLM TensorFlow fork
(speech_model.py:101-111) and TensorBoard summaries (fetched as None).
"""
import abc
import glob
import os
import re
import sys
import time

import numpy as np
import torch

from . import vocabulary
from .engine import W2LEngine
from .errors import OutOfRangeError  # noqa: F401  (re-exported for callers)
from .speech_input import BaseInputLoader


_STEP_TRACE = os.environ.get('SPEECHT_B200_STEP_TRACE') == '1'      # host-side phase timing of evaluate steps -> stderr


class Session:
  """Opaque execution context handed to model.step (the reference passes a tf.Session, training.py:46)."""

  def __init__(self, device=None):
    if not torch.cuda.is_available():
      raise RuntimeError('speecht_b200 needs a CUDA device: there is no CPU fallback')
    self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())

  def run(self, op):
    """sess.run(model.learning_rate_decay_op) / sess.run(variable.assign(..)) equivalents: ops are callables."""
    if isinstance(op, (list, tuple)):
      return [self.run(o) for o in op]
    return op()

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    torch.cuda.synchronize(self.device)
    return False


class HostVariable:
  """tf.Variable stand-in for the scalars callers read with .eval() (training.py:70,75)."""

  def __init__(self, value, name):
    self.value, self.name = value, name

  def eval(self, session=None):
    return self.value

  def assign(self, value):
    def op():
      self.value = value() if callable(value) else value
      return self.value
    return op


class _GlobalStep(HostVariable):
  def __init__(self, engine):
    self._engine = engine
    self.name = 'global_step'

  @property
  def value(self):
    return self._engine.global_step

  @value.setter
  def value(self, v):
    self._engine.global_step = int(v)


class NullSummaryWriter:
  """TensorBoard summaries are out of scope; keeps caller code (training.py:79) running."""

  def __init__(self, logdir):
    self.logdir = logdir

  def add_summary(self, summary, global_step=None):
    pass

  def add_graph(self, graph):
    pass


class Saver:
  """tf.train.Saver stand-in with an own single-file format (TF bundles are unreadable without TF).

  save(sess, path, global_step=) writes `<path>-<step>.npz` holding the flat parameter / Adam buffers, global_step
  and learning_rate, plus a `checkpoint` text file naming the latest one (like tf.train.get_checkpoint_state)."""

  def __init__(self, model, max_to_keep=5):
    self.model = model
    self.max_to_keep = max_to_keep              # tf.train.Saver default: the five most recent checkpoints stay
    self._kept = []

  def save(self, sess, save_path, global_step=None):
    step = global_step.eval() if hasattr(global_step, 'eval') else global_step
    path = '%s-%d' % (save_path, step) if step is not None else save_path
    eng = self.model.engine
    # written under a temporary name and renamed: a crash mid-write never leaves a truncated file under the final name
    tmp = path + '.tmp.npz'
    np.savez(tmp, params=eng.params.cpu().numpy(), adam_m=eng.adam_m.cpu().numpy(),
             adam_v=eng.adam_v.cpu().numpy(), global_step=np.int64(eng.global_step),
             learning_rate=np.float64(self.model.learning_rate.value if hasattr(self.model, 'learning_rate') else 0))
    os.replace(tmp, path + '.npz')
    marker = os.path.join(os.path.dirname(path) or '.', 'checkpoint')
    with open(marker + '.tmp', 'w') as f:
      f.write('model_checkpoint_path: "%s"\n' % os.path.basename(path))
    os.replace(marker + '.tmp', marker)
    if path in self._kept:
      self._kept.remove(path)
    self._kept.append(path)
    while self.max_to_keep and len(self._kept) > self.max_to_keep:
      old = self._kept.pop(0)
      try:
        os.remove(old + '.npz')
      except OSError:
        pass
    return path

  def restore(self, sess, path):
    with np.load(path + '.npz') as data:
      eng = self.model.engine
      if data['params'].shape[0] != eng.params.numel():
        raise ValueError('checkpoint holds %d floats, model has %d' % (data['params'].shape[0], eng.params.numel()))
      eng.params.copy_(torch.from_numpy(data['params']))
      eng.adam_m.copy_(torch.from_numpy(data['adam_m']))
      eng.adam_v.copy_(torch.from_numpy(data['adam_v']))
      eng.global_step = int(data['global_step'])
      if hasattr(self.model, 'learning_rate'):
        self.model.learning_rate.value = float(data['learning_rate'])
      eng.mark_weights_changed()


def latest_checkpoint(checkpoint_directory):
  marker = os.path.join(checkpoint_directory, 'checkpoint')
  if os.path.exists(marker):
    m = re.search(r'model_checkpoint_path: "(.*)"', open(marker).read())
    if m and os.path.exists(os.path.join(checkpoint_directory, m.group(1) + '.npz')):
      return os.path.join(checkpoint_directory, m.group(1))
  return None


def load_exported_weights(directory, n_layers=11):
  """Reads the layout `speecht-cli export --weights DIR` writes (exporting.py:30-40):
  DIR/convolution_layer_{i}/filters:0.npy [K,Cin,Cout] and DIR/convolution_layer_{i}/bias:0.npy [Cout]."""
  weights = []
  for i in range(n_layers):
    base = os.path.join(directory, 'convolution_layer_%d' % i)
    f = glob.glob(os.path.join(base, 'filters*.npy'))
    b = glob.glob(os.path.join(base, 'bias*.npy'))
    if not f or not b:
      raise FileNotFoundError('no exported weights for layer %d under %s' % (i, directory))
    weights.append((np.load(f[0]), np.load(b[0])))
  return weights


def save_exported_weights(directory, weights):
  """Writes the same layout (np.save appends .npy to 'filters:0' / 'bias:0')."""
  for i, (w, b) in enumerate(weights):
    base = os.path.join(directory, 'convolution_layer_%d' % i)
    os.makedirs(base, exist_ok=True)
    np.save(os.path.join(base, 'filters:0'), w)
    np.save(os.path.join(base, 'bias:0'), b)


# noinspection PyAttributeOutsideInit
class SpeechModel:

  def __init__(self, input_loader: BaseInputLoader, input_size: int, num_classes: int, precision: str = 'bf16x3',
               process_group=None, device=None, engine=None):
    """
    Args:
      input_loader: the object that provides input batches
      input_size: the number of values per time step
      num_classes: the number of output classes (vocabulary_size + 1 for blank label)
      precision: conv-stack arithmetic, see engine.PRECISIONS (not a reference argument; default meets its 1e-4 gate)
    """
    self.input_loader = input_loader
    self.input_size = input_size
    self.convolution_count = 0
    self.inputs, self.sequence_lengths, self.labels = input_loader.get_inputs()
    self.engine = engine if engine is not None else self._create_network(num_classes, precision, process_group,
                                                                         device)
    self.global_step = _GlobalStep(self.engine)
    self._training = False
    self._decoding = False
    self.max_gradient_norm = 5.0
    self._prefetched = None                      # next batch, already on its way to the device
    self._copy_stream = None
    # evaluate steps are software-pipelined: the kernels of the NEXT batch are enqueued before the host blocks on the
    # results of this one (SPEECHT_B200_EVAL_PIPELINE=0 disables it)
    self._speculated = None
    self._pipeline_eval = os.environ.get('SPEECHT_B200_EVAL_PIPELINE', '1') != '0'

  def add_training_ops(self, learning_rate: float = 1e-3, learning_rate_decay_factor: float = 0,
                       max_gradient_norm: float = 5.0, momentum: float = 0.9):
    """speech_model.py:53-82.  `momentum` is accepted and unused, exactly like the reference (Adam ignores it)."""
    self.learning_rate = HostVariable(float(learning_rate), 'learning_rate')
    self.learning_rate_decay_op = self.learning_rate.assign(
      lambda: self.learning_rate.value * learning_rate_decay_factor)
    self.max_gradient_norm = max_gradient_norm
    self._training = self.labels is not None

  def add_decoding_ops(self, language_model: str = None, lm_weight: float = 0.8, word_count_weight: float = 0.0,
                       valid_word_count_weight: float = 2.3):
    """speech_model.py:84-115.  Greedy decoding only: the language-model branch needs the author's TensorFlow
    fork with KenLM (README.md:89) and is out of scope."""
    self.lm_weight, self.word_count_weight, self.valid_word_count_weight = \
      lm_weight, word_count_weight, valid_word_count_weight
    if language_model:
      raise NotImplementedError('beam search with a KenLM language model is out of scope (needs the TF fork)')
    self._decoding = True

  def finalize(self, log_dir: str, run_name: str, run_type: str):
    self.saver = Saver(self)
    self.merged_summaries = None
    self.summary_writer = NullSummaryWriter('{}/{}_{}'.format(log_dir, run_name, run_type))

  def init_session(self, sess, init_variables=True):
    if init_variables:
      self.engine.init_xavier(seed=int(os.environ.get('SPEECHT_B200_SEED', '0')))
      self.engine.reset_optimizer()

  def _to_device(self, inputs, stream=None, lengths=None):
    """Host batch -> device tensor (async from pinned memory); returns (tensor, ready event or None).
    A ragged batch (the reference zero-pads every utterance to the batch maximum, speech_input.py:38-43) is uploaded
    utterance by utterance into a zeroed device buffer when more than a quarter of it is padding: the zeros are
    produced on the device instead of crossing PCIe (an evaluate batch of 1-30 s utterances is half padding)."""
    if torch.is_tensor(inputs) and inputs.is_cuda:
      return inputs, None
    host = inputs if torch.is_tensor(inputs) else torch.from_numpy(np.ascontiguousarray(inputs, dtype=np.float32))

    def upload():
      if lengths is not None and host.dim() == 3 and host.shape[0] > 1:
        lens = np.minimum(np.asarray(lengths, dtype=np.int64), host.shape[1])
        if int(lens.sum()) < 0.75 * host.shape[0] * host.shape[1]:
          dev = torch.zeros(host.shape, dtype=host.dtype, device=self.engine.device)
          for b, n in enumerate(lens):
            if n > 0:
              dev[b, :n].copy_(host[b, :n], non_blocking=True)
          return dev
      return host.to(self.engine.device, non_blocking=True)

    if stream is None:
      return upload(), None
    with torch.cuda.stream(stream):
      dev = upload()
      ev = torch.cuda.Event()
      ev.record(stream)
    return dev, ev

  def _fetch(self, feed_dict, stream=None):
    """(device inputs, lengths, labels, ready event) of the next batch, or an exception instance at end of data."""
    feed = self.input_loader.get_feed_dict() or {}
    if feed_dict is not None:
      feed.update(feed_dict)
    if self.inputs in feed:
      inputs, lengths = feed[self.inputs], feed[self.sequence_lengths]
      labels = feed.get(self.labels) if self.labels is not None else None
    else:
      batch = self.input_loader.dequeue()
      if batch is None:
        raise OutOfRangeError('no input available')
      inputs, lengths, labels = batch
    dev, ev = self._to_device(inputs, stream, lengths)
    return dev, lengths, labels, ev

  def _next_batch(self, feed_dict):
    """This step's batch.  Batches that come from the loader's queue are prefetched one step ahead: their
    host->device copy runs on a side stream underneath the previous step's kernels (the reference's FIFOQueue
    plays the same role on the host side, speech_input.py:147-156)."""
    if feed_dict is None and self._prefetched is not None:
      item, self._prefetched = self._prefetched, None
      if isinstance(item, Exception):
        raise item
      dev, lengths, labels, ev = item
      if ev is not None:
        torch.cuda.current_stream().wait_event(ev)
        dev.record_stream(torch.cuda.current_stream())    # allocated on the copy stream, consumed on this one
      return dev, lengths, labels
    dev, lengths, labels, _ = self._fetch(feed_dict)
    return dev, lengths, labels

  def _prefetch_next(self):
    if getattr(self.input_loader, 'queue', None) is None and not getattr(self.input_loader, 'prefetchable', False):
      return
    if torch.device(self.engine.device).type != 'cuda':
      # host-only use (the control flow is tested on the CPU with a stub engine): nothing to overlap, fetch in line
      try:
        self._prefetched = self._fetch(None, None)
      except OutOfRangeError as e:
        self._prefetched = e
      return
    if self._copy_stream is None:
      self._copy_stream = torch.cuda.Stream(device=self.engine.device)
    try:
      self._prefetched = self._fetch(None, self._copy_stream)
    except OutOfRangeError as e:
      self._prefetched = e

  def _speculate(self, flags, buckets):
    """Software pipelining of evaluate steps: takes the prefetched next batch and enqueues its forward / loss / decode
    kernels now, so that the GPU works on it while the host reads back and post-processes the current results (the
    decoded rows and losses of a step live in their own buffers; only `logits` of the previous step is overwritten).
    The result is handed out by the next step() if it asks for the same outputs; a training step just takes the batch."""
    item = self._prefetched
    if not self._pipeline_eval or buckets > 1 or item is None or isinstance(item, Exception):
      return
    self._prefetched = None
    dev, lengths, labels, ev = item
    if ev is not None:
      torch.cuda.current_stream().wait_event(ev)
      dev.record_stream(torch.cuda.current_stream())
    res = None
    try:
      if not flags[0] or labels is not None:
        res = self.engine.evaluate_step(dev, lengths, labels if flags[0] else None, decode=flags[1], buckets=1,
                                        defer_decode=True, fresh_decode=True)
    except Exception:
      res = None                                # e.g. a label that does not fit: raised by the step that owns the batch
    self._speculated = {'batch': (dev, lengths, labels), 'flags': flags, 'res': res}

  def input_exhausted(self):
    """True when the NEXT step would find no batch (the prefetch already hit the end, or the loader says so)."""
    if self._speculated is not None:
      return False
    if self._prefetched is not None:
      return isinstance(self._prefetched, Exception)
    at_end = getattr(self.input_loader, 'at_end', None)
    return bool(at_end()) if at_end is not None else False

  def step(self, sess, loss=True, update=True, decode=False, return_label=False, summary=False, feed_dict=None):
    """speech_model.py:197-235.  Returns: avg_loss (optional), decoded (optional), label (optional),
    update (optional, None), summary (optional, None) -- in that order."""
    spec = None
    if feed_dict is None and self._speculated is not None:
      # the batch an earlier evaluate step already took from the loader (and enqueued kernels for)
      spec, self._speculated = self._speculated, None
      d_inputs, lengths, labels = spec['batch']
    else:
      d_inputs, lengths, labels = self._next_batch(feed_dict)
    if (loss or update) and (labels is None or not self._training):
      raise ValueError('loss/update requested but the model has no labels / training ops')
    if decode and not self._decoding:
      raise ValueError('decode requested but add_decoding_ops was not called')
    if update:
      trace = [time.perf_counter()] if _STEP_TRACE else None
      res = self.engine.train_step(d_inputs, lengths, labels, self.learning_rate.value, self.max_gradient_norm,
                                   decode=decode)
      if trace: trace.append(time.perf_counter())
      if feed_dict is None:
        self._prefetch_next()                   # next batch's H2D overlaps this step's kernels
      if trace:
        trace.append(time.perf_counter())
        if loss:
          res['avg_loss'].item()
        trace.append(time.perf_counter())
        sys.stderr.write('train step trace (ms): enqueue %.2f prefetch %.2f loss wait %.2f | since last %.2f\n'
                         % tuple([1e3 * (b - a) for a, b in zip(trace[:-1], trace[1:])] +
                                 [1e3 * (trace[0] - getattr(self, '_trace_last', trace[0]))]))
        self._trace_last = trace[-1]
    else:
      # an evaluate step ends with a blocking read of the decoded labels.  Its kernels are enqueued FIRST; fetching
      # the next batch and starting its upload (host work: a ragged batch is 256 small copies) then runs underneath
      # them, and only after that does the host block on the results
      buckets = getattr(self, 'eval_buckets', 1)
      flags = (bool(loss), bool(decode))
      trace = [time.perf_counter()] if _STEP_TRACE else None
      if spec is not None and spec['res'] is not None and spec['flags'] == flags:
        res = spec['res']                       # enqueued during the previous step
      else:
        res = self.engine.evaluate_step(d_inputs, lengths, labels if loss else None, decode=decode, buckets=buckets,
                                        defer_decode=True)
      if trace: trace.append(time.perf_counter())
      if feed_dict is None:
        self._prefetch_next()
        if trace: trace.append(time.perf_counter())
        self._speculate(flags, buckets)         # the next batch's kernels go in BEFORE the blocking read below
      if trace: trace.append(time.perf_counter())
      self.engine.finish_evaluate(res)
      if trace:
        trace.append(time.perf_counter())
        if loss:
          res['avg_loss'].item()
        trace.append(time.perf_counter())
        sys.stderr.write('evaluate step trace (ms): enqueue %.2f prefetch %.2f speculate %.2f finish %.2f loss %.2f | since last %.2f\n'
                         % tuple([1e3 * (b - a) for a, b in zip(trace[:-1], trace[1:])] +
                                 [1e3 * (trace[0] - getattr(self, '_trace_last', trace[0]))]))
        self._trace_last = trace[-1]
    self.last_result = res
    output = []
    if loss:
      output.append(np.float32(res['avg_loss'].item()))       # the device->host read of the step's result ([B] losses)
    if decode:
      output.append(res['decoded'])
    if return_label:
      output.append(labels)
    if update:
      output.append(None)
    if summary:
      output.append(None)
    return output

  @abc.abstractmethod
  def _create_network(self, num_classes, precision, process_group, device):
    raise NotImplementedError()

  @property
  def logits(self):
    """[time, batch_size, num_classes] view of the last forward pass (speech_model.py:47,295)."""
    return self.last_result['logits']

  def restore(self, session, checkpoint_directory: str, reset_learning_rate: float = None):
    ckpt = latest_checkpoint(checkpoint_directory)
    if ckpt:
      print('Reading model parameters from {}'.format(ckpt))
      self.saver.restore(session, ckpt)
      self.init_session(session, init_variables=False)
      if reset_learning_rate:
        self.learning_rate.value = reset_learning_rate
    elif os.path.isdir(os.path.join(checkpoint_directory, 'convolution_layer_0')):
      print('Reading exported weights from {}'.format(checkpoint_directory))
      self.engine.load_weights(load_exported_weights(checkpoint_directory, len(self.engine.layers)))
      self.init_session(session, init_variables=False)
    else:
      raise FileNotFoundError('No checkpoint for evaluation found')

  def restore_or_create(self, session, checkpoint_directory: str, reset_learning_rate: float = None):
    try:
      self.restore(session, checkpoint_directory, reset_learning_rate)
    except FileNotFoundError:
      print('Created model with fresh parameters.')
      self.init_session(session, init_variables=True)


class Wav2LetterModel(SpeechModel):

  def _create_network(self, num_classes, precision, process_group, device):
    """The 11-layer stack of speech_model.py:275-295 (layer table in engine.layer_table)."""
    engine = W2LEngine(self.input_size, num_classes, device=device, precision=precision,
                       process_group=process_group)
    self.convolution_count = len(engine.layers)
    return engine


def create_default_model(flags, input_size: int, speech_input: BaseInputLoader) -> SpeechModel:
  """speech_model.py:298-324: same flag names; `flags.precision` (optional) selects the conv arithmetic."""
  model = Wav2LetterModel(input_loader=speech_input, input_size=input_size, num_classes=vocabulary.SIZE + 1,
                          precision=getattr(flags, 'precision', 'bf16x3'),
                          process_group=getattr(flags, 'process_group', None),
                          engine=getattr(flags, 'engine', None))
  if flags.command == 'train':
    model.add_training_ops(learning_rate=flags.learning_rate,
                           learning_rate_decay_factor=flags.learning_rate_decay_factor,
                           max_gradient_norm=flags.max_gradient_norm,
                           momentum=flags.momentum)
    model.add_decoding_ops()
  elif flags.command == 'export':
    model.add_training_ops()
    model.add_decoding_ops()
  else:
    model.add_training_ops()
    model.add_decoding_ops(language_model=getattr(flags, 'language_model', None),
                           lm_weight=getattr(flags, 'lm_weight', 0.8),
                           word_count_weight=getattr(flags, 'word_count_weight', 0.0),
                           valid_word_count_weight=getattr(flags, 'valid_word_count_weight', 2.3))
  # optional (not a reference flag): evaluate a ragged batch as length-sorted groups, see W2LEngine.evaluate_step
  model.eval_buckets = int(getattr(flags, 'eval_buckets', 1) or 1)
  model.finalize(log_dir=flags.log_dir, run_name=flags.run_name, run_type=flags.run_type)
  return model
