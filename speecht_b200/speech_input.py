"""Input loaders (mirror of reference speecht/speech_input.py).

Same class and method names -- BaseInputLoader._get_inputs_feed_item / _get_labels_feed_item, get_inputs,
get_feed_dict, SingleInputLoader.set_input, InputBatchLoader.start_threads -- but the TF1 FIFOQueue(100) fed by
enqueue ops (speech_input.py:147-156,181-207) is a plain bounded queue.Queue of ready host batches; model.step
dequeues from it.  The feed layout is the reference's: inputs zero-padded to the batch max length with NO masking
downstream (speech_input.py:38-43), labels as the COO triple with dense_shape [B, max input time]
(speech_input.py:59-69).
"""
import queue
import threading
from abc import abstractmethod

import numpy as np

from .errors import OutOfRangeError
from .ops import SparseTensorValue


class Placeholder:
  """Stands in for the tf.placeholder handles callers use as feed_dict keys."""

  def __init__(self, name):
    self.name = name

  def __repr__(self):
    return '<Placeholder %s>' % self.name


class Coordinator:
  """Minimal tf.train.Coordinator: should_stop / request_stop / register_thread / join."""

  def __init__(self):
    self._stop = threading.Event()
    self._threads = []

  def should_stop(self):
    return self._stop.is_set()

  def request_stop(self):
    self._stop.set()

  def register_thread(self, t):
    self._threads.append(t)

  def join(self, timeout=5.0):
    for t in self._threads:
      t.join(timeout)


class BaseInputLoader:

  def __init__(self, input_size):
    self.input_size = input_size

  def _get_inputs_feed_item(self, input_list):
    """speech_input.py:27-45 -> (input_tensor [B,max_time,input_size] f32, sequence_lengths, max_time).
    (The reference builds float64 and lets the f32 placeholder cast; building f32 directly is value-identical.)"""
    sequence_lengths = np.array([inp.shape[0] for inp in input_list], dtype=np.int32)
    max_time = int(sequence_lengths.max())
    shape = (len(input_list), max_time, self.input_size)
    input_tensor = None
    try:
      import torch
      if torch.cuda.is_available():
        # page-locked staging buffer: the host->device copy of model.step is then truly asynchronous
        input_tensor = torch.zeros(shape, dtype=torch.float32, pin_memory=True).numpy()
    except Exception:
      input_tensor = None
    if input_tensor is None:
      input_tensor = np.zeros(shape, dtype=np.float32)
    for idx, inp in enumerate(input_list):
      input_tensor[idx, :inp.shape[0], :] = inp
    return input_tensor, sequence_lengths, max_time

  @staticmethod
  def _get_labels_feed_item(label_list, max_time):
    """Sparse COO labels as the reference feeds them (speech_input.py:48-69): indices [N,2] = (utterance, position),
    values [N], dense_shape [batch, max_time] -- the dense width is the INPUT max_time, not the longest label."""
    lengths = [len(label) for label in label_list]
    total = int(sum(lengths))
    indices = np.zeros((total, 2), dtype=np.int64)
    values = np.zeros((total,), dtype=np.int64)
    cursor = 0
    for row, (label, n) in enumerate(zip(label_list, lengths)):
      indices[cursor:cursor + n, 0] = row
      indices[cursor:cursor + n, 1] = np.arange(n)
      values[cursor:cursor + n] = np.asarray(label, dtype=np.int64).reshape(-1)
      cursor += n
    return SparseTensorValue(indices, values, np.array([len(label_list), max_time], dtype=np.int64))

  @abstractmethod
  def get_inputs(self):
    raise NotImplementedError()

  def get_feed_dict(self):
    return None

  def dequeue(self):
    """Next ready batch (inputs, sequence_lengths, labels) or None when the loader feeds through get_feed_dict."""
    return None


class SingleInputLoader(BaseInputLoader):
  """Feeds single inputs through the feed dict (speech_input.py:79-127)."""

  def __init__(self, input_size):
    super().__init__(input_size)
    self.speech_input = None
    self.inputs = Placeholder('inputs')
    self.sequence_lengths = Placeholder('sequence_lengths')

  def get_inputs(self):
    return self.inputs, self.sequence_lengths, None

  def get_feed_dict(self):
    if self.speech_input is None:
      raise ValueError('Speech input must be provided using `set_input` first!')
    input_tensor, sequence_lengths, _max_time = self._get_inputs_feed_item([self.speech_input])
    self.speech_input = None
    return {self.inputs: input_tensor, self.sequence_lengths: sequence_lengths}

  def set_input(self, speech_input):
    self.speech_input = speech_input


class InputBatchLoader(BaseInputLoader):
  """Background threads assemble batches into a bounded queue (speech_input.py:130-218)."""

  _END = object()

  def __init__(self, input_size, batch_size, data_generator_creator, max_steps=None, capacity=100):
    super().__init__(input_size)
    self.batch_size = batch_size
    self.data_generator_creator = data_generator_creator
    self.steps_left = max_steps
    self.inputs = Placeholder('inputs')
    self.sequence_lengths = Placeholder('sequence_lengths')
    self.labels = Placeholder('labels')
    self.queue = queue.Queue(maxsize=capacity)
    self._lock = threading.Lock()
    self._live_threads = 0
    self._closed = False
    self._error = None                      # first exception raised inside a feeder thread
    self._held = None                       # item taken off the queue by at_end(), handed out by the next dequeue

  def get_inputs(self):
    return self.inputs, self.sequence_lengths, self.labels

  def _batch(self, iterable):
    args = [iter(iterable)] * self.batch_size
    return zip(*args)

  def _put(self, item, coord):
    while True:
      try:
        self.queue.put(item, timeout=0.1)
        return True
      except queue.Full:
        if coord is not None and coord.should_stop():
          return False

  def _close(self, coord):
    """Append the end marker.  When the consumer has stopped (full queue, stop requested) one stale batch is
    dropped to make room, so a feeder thread never blocks forever on a queue nobody reads."""
    while True:
      try:
        self.queue.put_nowait(self._END)
        return
      except queue.Full:
        if coord is None or coord.should_stop():
          try:
            self.queue.get_nowait()
          except queue.Empty:
            pass
        else:
          import time
          time.sleep(0.05)

  def _enqueue(self, sess, coord):
    try:
      data_generator = self.data_generator_creator()
      for sample_batch in self._batch(data_generator):
        input_list, label_list = zip(*sample_batch)
        input_tensor, sequence_lengths, max_time = self._get_inputs_feed_item(input_list)
        labels = self._get_labels_feed_item(label_list, max_time)
        if not self._put((input_tensor, sequence_lengths, labels), coord):
          break
        # the reference decrements steps_left unsynchronised across feeder threads (speech_input.py:199-202);
        # here the counter is locked so exactly max_steps batches are produced
        with self._lock:
          if self.steps_left is not None:
            self.steps_left -= 1
            if self.steps_left <= 0:
              break
        if coord is not None and coord.should_stop():
          break
    except BaseException as e:              # bad .npz, KeyError, ...: surfaces in the consumer, not a silent "done"
      with self._lock:
        if self._error is None:
          self._error = e
      if coord is not None:
        coord.request_stop()
    finally:
      with self._lock:
        self._live_threads -= 1
        last = self._live_threads == 0
      if last:
        self._closed = True
        self._close(coord)

  def start_threads(self, sess, coord, n_threads=1):
    threads = []
    with self._lock:
      self._live_threads += n_threads
    for _ in range(n_threads):
      t = threading.Thread(target=self._enqueue, args=(sess, coord))
      t.daemon = True
      t.start()
      coord.register_thread(t)
      threads.append(t)
    return threads

  def raise_if_failed(self):
    if self._error is not None:
      raise RuntimeError('input feeder thread failed') from self._error

  def at_end(self):
    """Blocks until the next batch or the end marker is available; True when the stream has ended (or a feeder
    failed).  Data-parallel ranks use it to agree on termination BEFORE entering a step (training.py)."""
    if self._error is not None:
      return True
    if self._held is None:
      self._held = self.queue.get()
    return self._held is self._END

  def dequeue(self):
    self.raise_if_failed()
    if self._held is not None:
      item, self._held = self._held, None
    else:
      item = self.queue.get()
    if item is self._END:
      try:
        self.queue.put_nowait(self._END)   # stay closed for any later step
      except queue.Full:
        pass
      if self._error is not None:
        raise RuntimeError('input feeder thread failed') from self._error
      raise OutOfRangeError('input queue is closed and has insufficient elements')
    return item
