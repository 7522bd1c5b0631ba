"""CPU tests of the rows SURVEY.md 8(f) marks "next": the .npz sample store, transcripts, LER/WER statistics with
the extract_decoded_ids quirk, checkpoint/export layouts, CLI flag surface.  Known answers come from the
reference's own test (speecht/tests/test_speechCorpusReader.py:25-35) where it has them."""
import importlib.machinery
import importlib.util
import os
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_edit_distance_and_eval_statistics():
  from speecht_b200.evaluation import EvalStatistics, edit_distance
  assert edit_distance('kitten', 'sitting') == 3
  assert edit_distance('', 'abc') == 3 and edit_distance('abc', 'abc') == 0
  assert edit_distance('a b c'.split(), 'a x c d'.split()) == 2
  s = EvalStatistics()
  s.track_decoding('he hoped there would be stew', 'he hoped there would be stew')
  s.track_decoding('he hopd their', 'he hoped there')
  assert s.letter_edit_distance == 3 and s.word_edit_distance == 2
  assert abs(s.letter_error_rate - 3 / 14) < 1e-12 and abs(s.word_error_rate - 2 / 3) < 1e-12
  assert abs(s.global_letter_edit_distance - 1.5) < 1e-12 and abs(s.global_word_error_rate - 1 / 3) < 1e-12


def test_extract_decoded_ids_keeps_the_reference_quirk():
  from speecht_b200.evaluation import Evaluation
  from speecht_b200.ops import SparseTensorValue
  sp = SparseTensorValue(np.array([[0, 0], [0, 1], [2, 0]]), np.array([5, 6, 7]), np.array([3, 2]))
  assert [list(map(int, x)) for x in Evaluation.extract_decoded_ids(sp)] == [[5, 6], [7]]


def test_transcripts_match_the_reference_fixture_known_answers(tmp_path):
  """First / last entries of the reference's own transcript fixture (test_speechCorpusReader.py:28-32), re-typed
  here as a two-line file (the 38-line fixture is not copied)."""
  from speecht_b200 import vocabulary
  from speecht_b200.preprocessing import SpeechCorpusReader
  d = tmp_path / 'train' / '1089' / '134686'
  d.mkdir(parents=True)
  first = ('HE HOPED THERE WOULD BE STEW FOR DINNER TURNIPS AND CARROTS AND BRUISED POTATOES AND FAT MUTTON '
           'PIECES TO BE LADLED OUT IN THICK PEPPERED FLOUR FATTENED SAUCE')
  last = 'IN THE SILENCE THEIR DARK FIRE KINDLED THE DUSK INTO A TAWNY GLOW'
  (d / '1089-134686.trans.txt').write_text('1089-134686-0000 %s\n1089-134686-0037 %s\n' % (first, last))
  entries = list(SpeechCorpusReader._get_transcript_entries(str(tmp_path)))
  assert entries[0] == ['1089-134686-0000', first] and entries[-1] == ['1089-134686-0037', last]
  reader = SpeechCorpusReader(str(tmp_path))
  ids = reader._transcript_dict['1089-134686-0037']
  assert vocabulary.ids_to_sentence(ids) == last.lower() and len(ids) == len(last)


def test_npz_store_layout_and_load_samples(tmp_path):
  from speecht_b200.preprocessing import SpeechCorpusReader, calc_power_spectrogram
  reader = SpeechCorpusReader(str(tmp_path))
  assert reader._get_directory('power', 'train').endswith('/preprocessed-power/train')
  assert reader._get_directory(calc_power_spectrogram, 'dev').endswith('/preprocessed-power/dev')
  assert reader._get_directory('mfcc', 'test').endswith('/preprocessed/test')
  out = tmp_path / 'preprocessed-power' / 'train'
  out.mkdir(parents=True)
  rng = np.random.default_rng(0)
  for i, t in enumerate((7, 12, 30)):
    np.savez(out / ('utt%d' % i), audio_fragments=rng.standard_normal((t, 128)).astype(np.float32),
             transcript=np.arange(i + 1))
  got = sorted((a.shape[0], len(tr)) for a, tr in reader.load_samples('train', feature_type='power'))
  assert got == [(7, 1), (12, 2), (30, 3)]
  assert len(list(reader.load_samples('train', feature_type='power', limit_count=2))) == 2
  assert sorted(a.shape[0] for a, _ in reader.load_samples('train', feature_type='power', max_size=12)) == [7, 12]
  it = reader.load_samples('train', feature_type='power', loop_infinitely=True)
  assert len([next(it) for _ in range(7)]) == 7
  with pytest.raises(ValueError):
    next(reader.load_samples('nope', feature_type='power'))


def test_wav_loader_roundtrip(tmp_path):
  import wave
  from speecht_b200.preprocessing import read_wav as load_wav
  x = (np.sin(np.arange(1600) * 0.1) * 12000).astype(np.int16)
  with wave.open(str(tmp_path / 'a.wav'), 'wb') as f:
    f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000); f.writeframes(x.tobytes())
  data, sr = load_wav(str(tmp_path / 'a.wav'))
  assert sr == 16000 and data.shape == (1600,) and np.allclose(data, x / 32768.0)


def test_export_layout_roundtrip(tmp_path):
  from speecht_b200.speech_model import load_exported_weights, save_exported_weights
  rng = np.random.default_rng(1)
  weights = [(rng.standard_normal((3, 4, 5)).astype(np.float32), rng.standard_normal(5).astype(np.float32))
             for _ in range(11)]
  save_exported_weights(str(tmp_path), weights)
  assert os.path.exists(tmp_path / 'convolution_layer_3' / 'filters:0.npy')    # exporting.py:33-40 naming
  assert os.path.exists(tmp_path / 'convolution_layer_10' / 'bias:0.npy')
  back = load_exported_weights(str(tmp_path))
  for (w, b), (w2, b2) in zip(weights, back):
    assert np.array_equal(w, w2) and np.array_equal(b, b2)


def _load_cli():
  path = os.path.join(ROOT, 'speecht-cli-b200')
  loader = importlib.machinery.SourceFileLoader('speecht_cli_b200', path)
  spec = importlib.util.spec_from_loader('speecht_cli_b200', loader)
  mod = importlib.util.module_from_spec(spec)
  loader.exec_module(mod)
  return mod


def test_cli_flag_surface_matches_reference_defaults():
  cli = _load_cli()
  f = cli.parse(['train'])
  assert (f.batch_size, f.learning_rate, f.max_gradient_norm, f.steps_per_checkpoint, f.feature_type) == \
         (64, 1e-4, 5.0, 1000, 'power')                      # speecht-cli:43,53,67,76,80
  assert f.run_type == 'train' and f.run_train_dir == 'train/noname' and f.momentum == 0.9
  f = cli.parse(['evaluate', '--dev', '--step-count', '1', '--run-name', 'x', '--batch-size', '4'])
  assert f.run_type == 'dev' and f.step_count == 1 and f.run_train_dir == 'train/x' and f.should_save is True
  f = cli.parse(['evaluate'])
  assert f.dataset == 'test' and f.lm_weight == 0.8 and f.valid_word_count_weight == 2.3
  f = cli.parse(['preprocess', '--train-only'])
  assert f.train_only and not f.test_only and f.run_type == 'other'


def test_learning_rate_decay_policy_matches_reference_training_loop():
  """reference training.py:81-83: multiply the learning rate by learning_rate_decay_factor when the factor is > 0,
  more than two window losses are on record and this window's loss is worse than EACH of the last three."""
  from speecht_b200.speech_model import HostVariable
  from speecht_b200.training import Training

  class Sess:
    def run(self, op):
      return op()

  def run_policy(factor, losses):
    t = object.__new__(Training)
    t.flags = types.SimpleNamespace(learning_rate_decay_factor=factor)
    lr = HostVariable(1e-3, 'learning_rate')
    model = types.SimpleNamespace(learning_rate=lr, learning_rate_decay_op=lr.assign(lambda: lr.value * factor))
    history, rates = [], []
    for loss in losses:
      t._maybe_decay(Sess(), model, loss, history)
      history.append(loss)                      # training.py:84 appends AFTER the decision
      rates.append(lr.value)
    return rates

  # fewer than three recorded windows: never decays, whatever the loss does
  assert run_policy(0.5, [1.0, 2.0, 3.0]) == [1e-3] * 3
  # 4th window worse than each of the last three -> decay once; 5th is better than the 4th -> unchanged
  np.testing.assert_allclose(run_policy(0.5, [3.0, 2.0, 1.0, 3.5, 3.4]), [1e-3, 1e-3, 1e-3, 5e-4, 5e-4])
  # worse than two of the last three but not all -> unchanged (strict "> max")
  assert run_policy(0.5, [3.0, 2.0, 1.0, 2.5]) == [1e-3] * 4
  assert run_policy(0.5, [1.0, 1.0, 1.0, 1.0]) == [1e-3] * 4       # equal is not worse
  # keeps decaying while the loss keeps climbing
  np.testing.assert_allclose(run_policy(0.5, [1.0, 1.0, 1.0, 2.0, 3.0, 4.0])[-3:], [5e-4, 2.5e-4, 1.25e-4])
  # factor 0 (the CLI default, speecht-cli:69-72) disables the policy
  assert run_policy(0.0, [1.0, 1.0, 1.0, 9.0]) == [1e-3] * 4


def test_sharded_load_samples_partitions_one_commonly_shuffled_list(tmp_path):
  """Data-parallel reading (new; the reference is single-process): every rank shuffles the file list with the same
  private generator and keeps every world-th file -- disjoint, complete, and no rank loads what it discards."""
  import random
  from speecht_b200.preprocessing import SpeechCorpusReader
  out = tmp_path / 'preprocessed-power' / 'train'
  out.mkdir(parents=True)
  for i in range(10):
    np.savez(out / ('utt%02d' % i), audio_fragments=np.full((i + 1, 4), i, np.float32), transcript=np.array([i]))
  reader = SpeechCorpusReader(str(tmp_path))
  parts = [[int(tr[0]) for _a, tr in reader.load_samples('train', feature_type='power', shard=(r, 3),
                                                          rng=random.Random(7))] for r in range(3)]
  assert sorted(sum(parts, [])) == list(range(10))
  assert [len(p) for p in parts] == [4, 3, 3]
  whole = [int(tr[0]) for _a, tr in reader.load_samples('train', feature_type='power', rng=random.Random(7))]
  assert whole[0::3] == parts[0] and whole[1::3] == parts[1] and whole[2::3] == parts[2]
  assert whole != sorted(whole)                                   # it IS shuffled


def test_sharded_load_samples_never_starves_a_rank(tmp_path):
  """Regression (2-rank CLI run hung at 'Determine input size from first sample'): DatasetExecutor peeks ONE sample
  through the training generator -- limit_count=1, loop_infinitely=True, shard=(rank, world).  Rank 1's every-2nd
  slice of a one-file list is empty, and the infinite generator then spun forever without yielding."""
  import itertools
  import random
  from speecht_b200.preprocessing import SpeechCorpusReader
  out = tmp_path / 'preprocessed-power' / 'train'
  out.mkdir(parents=True)
  for i in range(5):
    np.savez(out / ('utt%02d' % i), audio_fragments=np.full((i + 1, 4), i, np.float32), transcript=np.array([i]))
  reader = SpeechCorpusReader(str(tmp_path))
  for rank in range(2):
    gen = reader.load_samples('train', loop_infinitely=True, limit_count=1, feature_type='power', shard=(rank, 2),
                              rng=random.Random(3))
    first, _tr = next(iter(gen))                                  # must return (used to hang for rank 1)
    assert first.shape[1] == 4
  # ... and an infinite generator over an empty directory ends instead of spinning
  empty = tmp_path / 'preprocessed-power' / 'dev'
  empty.mkdir(parents=True)
  assert list(itertools.islice(reader.load_samples('dev', loop_infinitely=True, feature_type='power'), 3)) == []
  # the usual case is untouched: with at least `world` files the slices stay disjoint
  parts = [[int(tr[0]) for _a, tr in reader.load_samples('train', feature_type='power', shard=(r, 2),
                                                          rng=random.Random(3))] for r in range(2)]
  assert sorted(parts[0] + parts[1]) == list(range(5)) and not set(parts[0]) & set(parts[1])


def test_feeder_thread_failure_surfaces_and_full_queue_does_not_block_shutdown():
  from speecht_b200 import speech_input

  def broken():
    yield np.zeros((3, 4), np.float32), [1]
    raise KeyError('audio_fragments')

  loader = speech_input.InputBatchLoader(4, 1, broken)
  coord = speech_input.Coordinator()
  loader.start_threads(None, coord, n_threads=1)
  with pytest.raises(RuntimeError, match='feeder thread failed'):   # fails fast: at the first or the second dequeue
    assert loader.dequeue()[0].shape == (1, 3, 4)
    loader.dequeue()
  assert coord.should_stop()
  coord.join()

  # a full queue nobody reads: the last feeder must still terminate once stop is requested
  def endless():
    while True:
      yield np.zeros((2, 4), np.float32), [0]

  loader = speech_input.InputBatchLoader(4, 1, endless, capacity=2)
  coord = speech_input.Coordinator()
  (thread,) = loader.start_threads(None, coord, n_threads=1)
  import time
  time.sleep(0.3)                                                  # queue fills up
  coord.request_stop()
  thread.join(timeout=3.0)
  assert not thread.is_alive()
  # at_end() lets data-parallel ranks agree on termination before a step
  loader = speech_input.InputBatchLoader(4, 1, lambda: iter([(np.zeros((2, 4), np.float32), [0])]))
  coord = speech_input.Coordinator()
  loader.start_threads(None, coord, n_threads=1)
  assert loader.at_end() is False
  loader.dequeue()
  assert loader.at_end() is True
  with pytest.raises(speech_input.OutOfRangeError):
    loader.dequeue()
  coord.join()


def test_saver_keeps_the_last_five_checkpoints_and_writes_atomically(tmp_path):
  import torch
  from speecht_b200.speech_model import HostVariable, Saver, latest_checkpoint
  eng = types.SimpleNamespace(params=torch.arange(8, dtype=torch.float32), adam_m=torch.zeros(8),
                              adam_v=torch.ones(8), global_step=0, mark_weights_changed=lambda: None)
  model = types.SimpleNamespace(engine=eng, learning_rate=HostVariable(1e-3, 'learning_rate'))
  saver = Saver(model)
  for step in range(1, 8):
    eng.global_step = step
    saver.save(None, str(tmp_path / 'speechT.ckpt'), global_step=step)
  kept = sorted(f for f in os.listdir(tmp_path) if f.endswith('.npz'))
  assert kept == ['speechT.ckpt-%d.npz' % s for s in range(3, 8)]      # tf.train.Saver max_to_keep = 5
  assert not [f for f in os.listdir(tmp_path) if '.tmp' in f]
  assert latest_checkpoint(str(tmp_path)).endswith('speechT.ckpt-7')
  eng.params.zero_()
  saver.restore(None, latest_checkpoint(str(tmp_path)))
  assert eng.params.tolist() == list(range(8)) and eng.global_step == 7
