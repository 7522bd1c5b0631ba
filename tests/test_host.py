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


def test_native_plan_host_logic_without_a_gpu(monkeypatch):
  """st_plan_create is pure host code (shapes, SAME padding, arena offsets): it must agree with the Python side's
  flat parameter layout and frame arithmetic for every shape class, on a box without a GPU."""
  import ctypes
  from speecht_b200._lib import check, lib
  from speecht_b200.engine import ParamLayout, layer_table
  layout = ParamLayout(layer_table(128, 29))
  sizes = {}
  for (B, T, npl) in [(4, 101, 2), (32, 1001, 2), (32, 1000, 1), (1, 2, 2), (3, 37, 3), (32, 3001, 1)]:
    h = ctypes.c_void_p()
    check(lib().st_plan_create(ctypes.byref(h), B, T, 128, 29, npl))
    try:
      assert lib().st_plan_param_floats(h) == layout.total
      assert lib().st_plan_logit_frames(h) == -(-T // 2)                  # ceil(T/2), speech_model.py:275 stride 2
      n = lib().st_plan_arena_bytes(h)
      assert n > 0 and n % 1024 == 0
      sizes[(B, T, npl)] = n
    finally:
      check(lib().st_plan_destroy(h))
  assert sizes[(32, 1001, 2)] > sizes[(4, 101, 2)]
  assert sizes[(32, 3001, 1)] > sizes[(32, 1000, 1)]
  # bad arguments are rejected with the C ABI's error code, not a crash
  h = ctypes.c_void_p()
  assert lib().st_plan_create(ctypes.byref(h), 4, 101, 100, 29, 2) != 0    # input_size not a multiple of 64
  assert lib().st_plan_create(ctypes.byref(h), 4, 101, 128, 29, 4) != 0    # n_planes outside 1..3
  # the fast-FIR buffers of layer 8 (default on for one / two planes) disappear with SPEECHT_B200_FFA=0 (read at plan
  # creation); three-plane plans never carry them
  monkeypatch.setenv('SPEECHT_B200_FFA', '0')
  check(lib().st_plan_create(ctypes.byref(h), 32, 1001, 128, 29, 2))
  try:
    assert lib().st_plan_arena_bytes(h) < sizes[(32, 1001, 2)]
  finally:
    check(lib().st_plan_destroy(h))


def test_sparse_from_host_rows_is_the_tf_sparse_triple():
  """Host half of the greedy decoder's output (ops.PendingDecode.finish): padded label rows + counts -> the
  SparseTensor triple tf.nn.ctc_greedy_decoder returns (row-major indices, dense_shape = [B, longest row]); compared
  with the oracle's decoder output format on the same rows."""
  from oracle import speecht_oracle as O
  from speecht_b200.ops import sparse_from_host_rows
  rows = np.array([[3, 1, 4, 0, 0], [0, 0, 0, 0, 0], [2, 7, 1, 8, 2]], dtype=np.int32)
  counts = np.array([3, 0, 5], dtype=np.int32)
  sp = sparse_from_host_rows(rows, counts)
  np.testing.assert_array_equal(sp.indices, [[0, 0], [0, 1], [0, 2], [2, 0], [2, 1], [2, 2], [2, 3], [2, 4]])
  np.testing.assert_array_equal(sp.values, [3, 1, 4, 2, 7, 1, 8, 2])
  np.testing.assert_array_equal(sp.dense_shape, [3, 5])
  assert sp.indices.dtype == np.int64 and sp.values.dtype == np.int64
  # the same rows through the oracle's decoder: one-hot logits whose arg-max path is row r separated by blanks
  C, blank = 10, 9
  T = 2 * rows.shape[1]
  logits = np.full((T, 3, C), -5.0)
  seq = np.zeros(3, dtype=np.int64)
  for b in range(3):
    path = []
    for v in rows[b, :counts[b]]:
      path += [int(v), blank]
    seq[b] = max(len(path), 1)
    for t in range(T):
      logits[t, b, path[t] if t < len(path) else blank] = 5.0
  (ri, rv, rs), _neg = O.ctc_greedy_decoder(logits, seq)
  np.testing.assert_array_equal(sp.indices, ri)
  np.testing.assert_array_equal(sp.values, rv)
  np.testing.assert_array_equal(sp.dense_shape, rs)
  empty = sparse_from_host_rows(np.zeros((2, 1), np.int32), np.zeros(2, np.int32))
  assert empty.indices.shape == (0, 2) and empty.values.shape == (0,) and list(empty.dense_shape) == [2, 0]


def test_kernel_resource_invariants_of_the_built_library():
  """Resource facts the design relies on, read from the built library with cuobjdump (skipped without the toolkit):
  * background kernels (run on a side stream beside resident tensor-core CTAs): at most 40
    registers and no static shared memory, so that 256 threads x 40 + 320 x 168 registers and 1 KB + 227 KB of shared
    memory fit one SM;
  * tensor-core kernels: 10 warps x 168 registers (or 18 warps x 112 for the sixteen-epilogue-warp instantiations)
    must fit the 64 K register file -- a build that needs more fails to LAUNCH, not to compile."""
  import re
  import shutil
  import subprocess
  from speecht_b200 import _lib
  tool = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
  if not os.path.exists(tool) or not os.path.exists(_lib.LIB_PATH):
    pytest.skip('cuobjdump or the built library is not available')
  out = subprocess.run([tool, '-res-usage', _lib.LIB_PATH], capture_output=True, text=True).stdout
  usage = {}
  for name, reg, shared in re.findall(r'Function (\S+):\s*\n\s*REG:(\d+) STACK:\d+ SHARED:(\d+)', out):
    usage[name] = (int(reg), int(shared))
  assert len(usage) > 50, 'no resource records parsed'
  background = [n for n in usage if re.search(r'pack_bwd_kernel|zero_f32_kernel|ffa2_dw_combine_kernel', n)]
  assert len(background) >= 4
  for n in background:
    reg, shared = usage[n]
    assert reg <= 40 and shared <= 1024, (n, reg, shared)        # 1024 = the per-block reservation, no static smem
  tensor = [n for n in usage if re.search(r'tc_conv_kernel|tc_wgrad_kernel', n)]
  assert len(tensor) >= 30
  for n in tensor:
    reg, shared = usage[n]
    wide_epilogue = re.search(r'tc_conv_kernelILi128ELi[12]E', n) is not None      # ConvCfg::EPW == 16: 576 threads
    assert reg <= (112 if wide_epilogue else 168), (n, reg)
    assert shared <= 1024, (n, shared)                                             # pipeline stages are dynamic smem
