"""Evaluation loop + LER/WER statistics (mirror of reference speecht/evaluation.py).

`editdistance` is not installed, so the Levenshtein distance is a small DP here.  `extract_decoded_ids` keeps the
reference's behaviour bit for bit -- including that an utterance decoding to the empty string yields no entry, so
later decodings pair with the wrong expected strings (evaluation.py:161-171)."""
import itertools

import numpy as np

from . import vocabulary
from .errors import OutOfRangeError
from .execution import DatasetExecutor
from .speech_model import Session


def edit_distance(a, b):
  """Levenshtein distance between two sequences (what editdistance.eval computes)."""
  a, b = list(a), list(b)
  if len(a) < len(b):
    a, b = b, a
  prev = list(range(len(b) + 1))
  for i, ca in enumerate(a, 1):
    cur = [i]
    for j, cb in enumerate(b, 1):
      cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
    prev = cur
  return prev[-1]


class EvalStatistics:
  """Running letter / word edit distances and error rates (evaluation.py:27-64)."""

  def __init__(self):
    self.decodings_counter = 0
    self.sum_letter_edit_distance = 0
    self.sum_letter_error_rate = 0
    self.sum_word_edit_distance = 0
    self.sum_word_error_rate = 0
    self.letter_edit_distance = 0
    self.letter_error_rate = 0
    self.word_edit_distance = 0
    self.word_error_rate = 0

  def track_decoding(self, decoded_str, expected_str):
    self.letter_edit_distance = edit_distance(expected_str, decoded_str)
    self.letter_error_rate = self.letter_edit_distance / len(expected_str)
    self.word_edit_distance = edit_distance(expected_str.split(), decoded_str.split())
    self.word_error_rate = self.word_edit_distance / len(expected_str.split())
    self.sum_letter_edit_distance += self.letter_edit_distance
    self.sum_letter_error_rate += self.letter_error_rate
    self.sum_word_edit_distance += self.word_edit_distance
    self.sum_word_error_rate += self.word_error_rate
    self.decodings_counter += 1

  @property
  def global_letter_edit_distance(self):
    return self.sum_letter_edit_distance / self.decodings_counter

  @property
  def global_letter_error_rate(self):
    return self.sum_letter_error_rate / self.decodings_counter

  @property
  def global_word_edit_distance(self):
    return self.sum_word_edit_distance / self.decodings_counter

  @property
  def global_word_error_rate(self):
    return self.sum_word_error_rate / self.decodings_counter


class Evaluation(DatasetExecutor):

  def create_sample_generator(self, limit_count: int):
    return self.reader.load_samples(self.flags.dataset, loop_infinitely=False, limit_count=limit_count,
                                    feature_type=self.flags.feature_type)

  def get_loader_limit_count(self):
    return self.flags.step_count * self.flags.batch_size

  def get_max_steps(self):
    return self.flags.step_count if self.flags.step_count else None

  def run(self):
    stats = EvalStatistics()
    with Session() as sess:
      model = self.create_model(sess)
      print('Starting input pipeline')
      coord = self.start_pipeline(sess)
      try:
        print('Begin evaluation')
        step_iter = range(self.flags.step_count) if self.flags.step_count else itertools.count()
        for step in step_iter:
          if coord.should_stop():
            break
          self.run_step(model, sess, stats, self.flags.should_save and step == 0)
      except OutOfRangeError:
        print('Done evaluating -- step limit reached')
      finally:
        coord.request_stop()
      self.print_global_statistics(stats)
      coord.join()
    return stats

  @staticmethod
  def print_global_statistics(stats):
    print('Global statistics')
    print('LED: {} LER: {:.2f} WED: {} WER: {:.2f}'.format(stats.global_letter_edit_distance,
                                                           stats.global_letter_error_rate,
                                                           stats.global_word_edit_distance,
                                                           stats.global_word_error_rate))

  def run_step(self, model, sess, stats, save, verbose=True, feed_dict=None):
    global_step = model.global_step.eval()
    if save:
      avg_loss, decoded, label, summary = model.step(sess, update=False, decode=True, return_label=True,
                                                     summary=True, feed_dict=feed_dict)
      model.summary_writer.add_summary(summary, global_step)
    else:
      avg_loss, decoded, label = model.step(sess, update=False, decode=True, return_label=True,
                                            feed_dict=feed_dict)
    if verbose:
      perplexity = np.exp(float(avg_loss)) if avg_loss < 300 else float('inf')
      print('validation average loss {:.2f} perplexity {:.2f}'.format(avg_loss, perplexity))
    decoded_ids_paths = [Evaluation.extract_decoded_ids(path) for path in decoded]
    for label_ids in Evaluation.extract_decoded_ids(label):
      expected_str = vocabulary.ids_to_sentence(label_ids)
      if verbose:
        print('expected: {}'.format(expected_str))
      for decoded_path in decoded_ids_paths:
        decoded_ids = next(decoded_path)
        decoded_str = vocabulary.ids_to_sentence(decoded_ids)
        stats.track_decoding(decoded_str, expected_str)
        if verbose:
          print('decoded: {}'.format(decoded_str))
          print('LED: {} LER: {:.2f} WED: {} WER: {:.2f}'.format(stats.letter_edit_distance, stats.letter_error_rate,
                                                                 stats.word_edit_distance, stats.word_error_rate))

  @staticmethod
  def extract_decoded_ids(sparse_tensor):
    ids = []
    last_batch_id = 0
    for i, index in enumerate(sparse_tensor.indices):
      batch_id, _char_id = index
      if batch_id > last_batch_id:
        yield ids
        ids = []
        last_batch_id = batch_id
      ids.append(sparse_tensor.values[i])
    yield ids
