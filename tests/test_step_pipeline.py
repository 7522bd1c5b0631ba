"""Control flow of SpeechModel.step around prefetching and the software-pipelined evaluate steps, on the CPU with a
stub engine: which batch every call works on, in which order kernels are "enqueued" and results read, what happens
when the requested outputs change, when a training step follows an evaluate step, when a speculatively enqueued batch
fails, and at the end of the data.  (The arithmetic is covered by the GPU suite; this is the bookkeeping that must
never skip, repeat or reorder a batch.)"""
import numpy as np
import pytest
import torch

from speecht_b200 import speech_input, speech_model
from speecht_b200.errors import OutOfRangeError


class Feed(speech_input.BaseInputLoader):
  """Batches tagged 0, 1, 2, ...: inputs[0, 0, 0] carries the tag."""

  def __init__(self, n):
    super().__init__(4)
    self.inputs, self.sequence_lengths, self.labels = (speech_input.Placeholder(x) for x in 'ilt')
    self.n, self.i, self.prefetchable = n, 0, True

  def get_inputs(self):
    return self.inputs, self.sequence_lengths, self.labels

  def dequeue(self):
    if self.i >= self.n:
      return None
    tag = self.i
    self.i += 1
    return np.full((2, 6, 4), tag, np.float32), np.array([6, 6], np.int32), [[1, 2], [3]]


class _Loss:
  def __init__(self, v):
    self.v = v

  def item(self):
    return self.v


class StubEngine:
  device = torch.device('cpu')
  global_step = 0
  layers = []

  def __init__(self, fail_on=None):
    self.log, self.fail_on = [], fail_on

  def evaluate_step(self, inputs, lengths, labels=None, decode=True, buckets=1, defer_decode=False, fresh_decode=False):
    tag = int(inputs[0, 0, 0])
    if self.fail_on is not None and tag == self.fail_on[0] and self.fail_on[1] > 0:
      self.fail_on = (tag, self.fail_on[1] - 1)
      self.log.append(('fail', tag))
      raise ValueError('label does not fit batch %d' % tag)
    self.log.append(('eval', tag, labels is not None, decode))
    return {'avg_loss': _Loss(100.0 + tag) if labels is not None else None,
            'decoded': ('pending', tag) if decode else None, 'logits': None, 'loss': None}

  def finish_evaluate(self, out):
    if isinstance(out.get('decoded'), tuple) and out['decoded'][0] == 'pending':
      self.log.append(('read', out['decoded'][1]))
      out['decoded'] = ('rows', out['decoded'][1])
    return out

  def train_step(self, inputs, lengths, labels, lr, max_norm, decode=False):
    tag = int(inputs[0, 0, 0])
    self.log.append(('train', tag))
    return {'avg_loss': _Loss(200.0 + tag), 'decoded': None}


def make(n, monkeypatch=None, pipeline=True, fail_on=None):
  if monkeypatch is not None:
    monkeypatch.setenv('SPEECHT_B200_EVAL_PIPELINE', '1' if pipeline else '0')
  eng = StubEngine(fail_on)
  model = speech_model.Wav2LetterModel(Feed(n), 4, 29, engine=eng)
  model.add_training_ops(learning_rate=1e-4)
  model.add_decoding_ops()
  return model, eng


def test_evaluate_steps_are_pipelined_in_order_and_end_cleanly(monkeypatch):
  model, eng = make(4, monkeypatch)
  seen = []
  for _ in range(4):
    loss, decoded, _labels = model.step(None, update=False, decode=True, return_label=True)
    seen.append((float(loss), decoded))
  assert seen == [(100.0 + k, ('rows', k)) for k in range(4)]
  # batch k+1 is enqueued BEFORE batch k is read back; every batch is evaluated exactly once
  assert eng.log == [('eval', 0, True, True), ('eval', 1, True, True), ('read', 0), ('eval', 2, True, True), ('read', 1),
                     ('eval', 3, True, True), ('read', 2), ('read', 3)]
  assert model.input_exhausted()
  with pytest.raises(OutOfRangeError):
    model.step(None, update=False, decode=True)


def test_pipeline_can_be_switched_off(monkeypatch):
  model, eng = make(3, monkeypatch, pipeline=False)
  for k in range(3):
    assert model.step(None, update=False, decode=True)[1] == ('rows', k)
  assert eng.log == [('eval', 0, True, True), ('read', 0), ('eval', 1, True, True), ('read', 1), ('eval', 2, True, True),
                     ('read', 2)]


def test_changed_outputs_discard_the_speculated_result_but_keep_its_batch(monkeypatch):
  model, eng = make(3, monkeypatch)
  assert model.step(None, update=False, decode=True)[1] == ('rows', 0)
  out = model.step(None, loss=True, update=False, decode=False)          # batch 1 again, now without decode
  assert [float(x) for x in out] == [101.0]
  assert ('eval', 1, True, True) in eng.log and ('eval', 1, True, False) in eng.log
  assert model.step(None, update=False, decode=True)[1] == ('rows', 2)    # nothing skipped, nothing repeated


def test_training_step_takes_the_batch_an_evaluate_step_had_speculated_on(monkeypatch):
  model, eng = make(3, monkeypatch)
  model.step(None, update=False, decode=True)
  out = model.step(None)                                                 # train on batch 1
  assert float(out[0]) == 201.0 and ('train', 1) in eng.log
  out = model.step(None)
  assert float(out[0]) == 202.0
  with pytest.raises(OutOfRangeError):
    model.step(None)


def test_a_failing_speculation_surfaces_at_the_step_that_owns_the_batch(monkeypatch):
  model, eng = make(3, monkeypatch, fail_on=(1, 2))
  assert model.step(None, update=False, decode=True)[1] == ('rows', 0)    # batch 1 failed behind the scenes: silent
  with pytest.raises(ValueError, match='batch 1'):
    model.step(None, update=False, decode=True)                          # ... and fails where it belongs
  assert eng.log.count(('fail', 1)) == 2


def test_feed_dict_steps_never_speculate(monkeypatch):
  model, eng = make(3, monkeypatch)
  feed = {model.inputs: np.full((1, 6, 4), 9, np.float32), model.sequence_lengths: np.array([6], np.int32),
          model.labels: [[1]]}
  assert model.step(None, update=False, decode=True, feed_dict=feed)[1] == ('rows', 9)
  assert eng.log == [('eval', 9, True, True), ('read', 9)]
  assert model.step(None, update=False, decode=True)[1] == ('rows', 0)    # the loader is untouched by the feed_dict step
