"""`evaluate`: greedy-decode a data set and report letter / word error statistics.

Behaviour follows reference speecht/evaluation.py: per step one `model.step(update=False, decode=True,
return_label=True)`, per utterance LED / LER / WED / WER (Levenshtein distances; `editdistance` is not installed, so a
small DP is used), global averages at the end, the same console lines.  `extract_decoded_ids` keeps the reference's
observable quirk: it starts a new utterance only when the batch index of the sparse entries INCREASES, so an utterance
that decodes to the empty string yields no entry and later decodings pair with the wrong expected strings
(evaluation.py:161-171).
"""
import itertools
import math

from . import vocabulary
from .errors import OutOfRangeError
from .execution import DatasetExecutor
from .speech_model import Session


def edit_distance(a, b):
  """Levenshtein distance between two sequences (insert / delete / substitute cost 1)."""
  a, b = list(a), list(b)
  if len(b) > len(a):
    a, b = b, a
  row = list(range(len(b) + 1))
  for i, item_a in enumerate(a, start=1):
    diagonal, row[0] = row[0], i
    for j, item_b in enumerate(b, start=1):
      substitute = diagonal + (item_a != item_b)
      diagonal = row[j]
      row[j] = min(row[j] + 1, row[j - 1] + 1, substitute)
  return row[-1]


class EvalStatistics:
  """Last-utterance values (`letter_edit_distance`, ...) and running sums / global means over all utterances."""

  METRICS = ('letter_edit_distance', 'letter_error_rate', 'word_edit_distance', 'word_error_rate')

  def __init__(self):
    self.decodings_counter = 0
    for name in self.METRICS:
      setattr(self, name, 0)
      setattr(self, 'sum_' + name, 0)

  def track_decoding(self, decoded_str, expected_str):
    expected_words, decoded_words = expected_str.split(), decoded_str.split()
    letters = edit_distance(expected_str, decoded_str)
    words = edit_distance(expected_words, decoded_words)
    current = {'letter_edit_distance': letters, 'letter_error_rate': letters / len(expected_str),
               'word_edit_distance': words, 'word_error_rate': words / len(expected_words)}
    for name, value in current.items():
      setattr(self, name, value)
      setattr(self, 'sum_' + name, getattr(self, 'sum_' + name) + value)
    self.decodings_counter += 1

  def _mean(self, name):
    return getattr(self, 'sum_' + name) / self.decodings_counter

  global_letter_edit_distance = property(lambda self: self._mean('letter_edit_distance'))
  global_letter_error_rate = property(lambda self: self._mean('letter_error_rate'))
  global_word_edit_distance = property(lambda self: self._mean('word_edit_distance'))
  global_word_error_rate = property(lambda self: self._mean('word_error_rate'))

  def last_line(self):
    return 'LED: {} LER: {:.2f} WED: {} WER: {:.2f}'.format(self.letter_edit_distance, self.letter_error_rate,
                                                            self.word_edit_distance, self.word_error_rate)

  def global_line(self):
    return 'LED: {} LER: {:.2f} WED: {} WER: {:.2f}'.format(self.global_letter_edit_distance,
                                                            self.global_letter_error_rate,
                                                            self.global_word_edit_distance,
                                                            self.global_word_error_rate)


class Evaluation(DatasetExecutor):

  def create_sample_generator(self, limit_count: int):
    return self.reader.load_samples(self.flags.dataset, loop_infinitely=False, limit_count=limit_count,
                                    feature_type=self.flags.feature_type)

  def get_loader_limit_count(self):
    return self.flags.batch_size * self.flags.step_count

  def get_max_steps(self):
    return self.flags.step_count or None

  def run(self):
    stats = EvalStatistics()
    steps = range(self.flags.step_count) if self.flags.step_count else itertools.count()
    with Session() as sess:
      model = self.create_model(sess)
      print('Starting input pipeline')
      coord = self.start_pipeline(sess)
      print('Begin evaluation')
      try:
        for step in steps:
          if coord.should_stop():
            break
          self.run_step(model, sess, stats, save=self.flags.should_save and step == 0)
      except OutOfRangeError:
        print('Done evaluating -- step limit reached')
      finally:
        coord.request_stop()
      self.print_global_statistics(stats)
      coord.join()
      self.speech_input.raise_if_failed()
    return stats

  @staticmethod
  def print_global_statistics(stats):
    print('Global statistics')
    print(stats.global_line())

  def run_step(self, model, sess, stats, save, verbose=True, feed_dict=None):
    fetched = model.step(sess, update=False, decode=True, return_label=True, summary=bool(save), feed_dict=feed_dict)
    avg_loss, decoded, label = fetched[:3]
    if save:
      model.summary_writer.add_summary(fetched[3], model.global_step.eval())
    if verbose:
      perplexity = math.exp(float(avg_loss)) if avg_loss < 300 else float('inf')
      print('validation average loss {:.2f} perplexity {:.2f}'.format(avg_loss, perplexity))
    decoded_streams = [self.extract_decoded_ids(path) for path in decoded]
    for expected_ids in self.extract_decoded_ids(label):
      expected_str = vocabulary.ids_to_sentence(expected_ids)
      if verbose:
        print('expected: {}'.format(expected_str))
      for stream in decoded_streams:
        decoded_str = vocabulary.ids_to_sentence(next(stream))
        stats.track_decoding(decoded_str, expected_str)
        if verbose:
          print('decoded: {}'.format(decoded_str))
          print(stats.last_line())

  @staticmethod
  def extract_decoded_ids(sparse_tensor):
    """Generator over the id list of each batch row present in a sparse (indices, values) pair -- rows WITHOUT
    entries are skipped, not yielded as empty lists (the reference's behaviour, see the module docstring)."""
    current_row, ids = 0, []
    for (row, _position), value in zip(sparse_tensor.indices, sparse_tensor.values):
      if row > current_row:
        yield ids
        current_row, ids = row, []
      ids.append(value)
    yield ids
