"""CPU tests of the host-side mirror and of the C-ABI library's exports (no GPU compute)."""
import ctypes
import os
import queue

import numpy as np
import pytest

from speecht_b200 import _lib, vocabulary


def test_library_builds_loads_and_exports_every_declared_symbol():
  from speecht_b200 import build
  path = build.build()
  assert os.path.exists(path)
  handle = _lib.lib()
  declared = _lib.declared_symbols()
  assert len(declared) >= 14
  for name in declared:
    assert hasattr(handle, name), name
  for name in _lib._SIGNATURES:
    assert name in declared, '%s bound in _lib.py but not declared in include/speecht_b200.h' % name
  assert handle.st_version() >= 100


def test_ctc_label_validation_is_host_side_and_mirrors_tf():
  lib = _lib.lib()
  def run(labels, offsets, seq, T=10, blank=28):
    l = np.asarray(labels, np.int32); o = np.asarray(offsets, np.int32); s = np.asarray(seq, np.int32)
    return lib.st_ctc_validate_labels_host(l.ctypes.data, o.ctypes.data, s.ctypes.data, len(seq), T, blank)
  assert run([1, 2, 3], [0, 3], [3]) == 0
  assert run([1, 1, 3], [0, 3], [3]) == _lib.ST_ERR_CTC_LABELS
  assert 'Not enough time for target transition sequence (required: 4, available: 3)' in _lib.last_error()
  assert run([1, 28], [0, 2], [5]) == _lib.ST_ERR_CTC_LABELS
  assert run([1], [0, 1], [11]) == _lib.ST_ERR_CTC_LABELS
  assert run([], [0, 0], [0]) == 0
  with pytest.raises(_lib.CTCLabelError):
    _lib.check(_lib.ST_ERR_CTC_LABELS)


def test_null_pointer_is_rejected_without_touching_the_gpu():
  lib = _lib.lib()
  rc = lib.st_conv1d_fwd_f32(None, None, None, None, 1, 1, 1, 1, 1, 1, 0, None)
  assert rc == _lib.ST_ERR_INVALID_ARG and 'null pointer' in _lib.last_error()
  with pytest.raises(ValueError):
    _lib.check(rc)


def test_vocabulary_matches_reference_table():
  assert vocabulary.SIZE == 28
  assert vocabulary.sentence_to_ids("Ab z'") == [0, 1, 27, 25, 26]
  assert vocabulary.ids_to_sentence([7, 8, 27, 26]) == "hi '"
  for i in range(28):
    assert vocabulary.letter_to_id(vocabulary.id_to_letter(i)) == i


def test_feed_items_follow_reference_layout():
  from speecht_b200.speech_input import BaseInputLoader
  loader = BaseInputLoader(3)
  a, b = np.ones((4, 3)), 2 * np.ones((2, 3))
  x, lens, max_time = loader._get_inputs_feed_item([a, b])
  assert x.shape == (2, 4, 3) and x.dtype == np.float32 and lens.tolist() == [4, 2] and max_time == 4
  assert np.all(x[1, 2:] == 0) and np.all(x[1, :2] == 2)
  sp = BaseInputLoader._get_labels_feed_item([[5, 6], [], [7]], max_time)
  assert sp.indices.tolist() == [[0, 0], [0, 1], [2, 0]] and sp.values.tolist() == [5, 6, 7]
  assert sp.dense_shape.tolist() == [3, 4]        # dense_shape uses the INPUT max_time (speech_input.py:59)


def test_input_batch_loader_threads_and_out_of_range():
  from speecht_b200.errors import OutOfRangeError
  from speecht_b200.speech_input import Coordinator, InputBatchLoader
  samples = [(np.full((3 + i, 2), i, np.float32), [i, i + 1]) for i in range(7)]
  loader = InputBatchLoader(2, 2, lambda: iter(samples))
  coord = Coordinator()
  loader.start_threads(None, coord, n_threads=1)
  seen = []
  with pytest.raises(OutOfRangeError):
    while True:
      x, lens, labels = loader.dequeue()
      seen.append((x.shape, lens.tolist(), labels.values.tolist()))
  assert len(seen) == 3                             # 7 samples -> 3 full batches, remainder dropped (zip semantics)
  assert seen[0] == ((2, 4, 2), [3, 4], [0, 1, 1, 2])
  with pytest.raises(OutOfRangeError):
    loader.dequeue()
  loader2 = InputBatchLoader(2, 2, lambda: iter(samples * 10), max_steps=2)
  loader2.start_threads(None, coord, n_threads=2)
  n = 0
  with pytest.raises(OutOfRangeError):
    while True:
      loader2.dequeue(); n += 1
  assert 2 <= n <= 3                                # exactly max_steps, +1 if both feeders raced the last slot
  coord.request_stop(); coord.join()


def test_mel_filterbank_matches_oracle_restatement():
  from oracle import speecht_oracle as O
  from speecht_b200.ops import mel_filterbank
  for sr in (16000, 22050):
    np.testing.assert_allclose(mel_filterbank(sr), O.mel_filterbank(sr), rtol=1e-6, atol=1e-9)


def test_param_layout_counts_and_alignment():
  from speecht_b200.engine import ParamLayout, layer_table
  lay = ParamLayout(layer_table())
  assert lay.n_params == 24662529
  assert all(o % 64 == 0 for o in lay.w_off + lay.b_off) and lay.total >= lay.n_params


def test_no_product_module_imports_the_oracle():
  import pathlib
  root = pathlib.Path(__file__).resolve().parents[1] / 'speecht_b200'
  for path in root.rglob('*.py'):
    for line in path.read_text().splitlines():
      if 'import' in line:
        assert 'oracle' not in line, (path, line)
