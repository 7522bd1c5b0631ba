"""The committed fixtures (tests/golden/*.npz, produced by make_golden.py from the float64 oracle) still match the
oracle, and the float32 oracle (the cpu_baseline arithmetic) stays within the parity tolerances of it."""
import os
import sys

import numpy as np

from oracle import speecht_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
sys.path.insert(0, GOLDEN)
import make_golden as G  # noqa: E402


def rel(a, b):
  """BASELINE.md metric: max|a-b| / max|b| per tensor."""
  return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))) / np.max(np.abs(b)))


def test_config1_fixture_matches_oracle_and_fp32_oracle():
  g = np.load(os.path.join(GOLDEN, 'config1_eval.npz'))
  inputs, lengths, labels, weights, res = G.config1(np.float64)
  assert abs(inputs.astype(np.float64).sum() - g['inputs_checksum']) < 1e-9
  np.testing.assert_allclose(res['logits'], g['logits'], rtol=0, atol=1e-12)
  np.testing.assert_allclose(res['loss'], g['loss'], rtol=1e-12)
  np.testing.assert_array_equal(res['decoded'][1], g['decoded_values'])
  np.testing.assert_array_equal(res['decoded'][0], g['decoded_indices'])
  r32 = O.evaluate_step(inputs, lengths, labels, weights, dtype=np.float32)
  assert rel(r32['logits'], g['logits']) < 1e-4
  assert rel(r32['loss'], g['loss']) < 1e-4
  np.testing.assert_array_equal(r32['decoded'][1], g['decoded_values'])


def test_ctc_fixture():
  g = np.load(os.path.join(GOLDEN, 'ctc_case.npz'))
  logits, seq, labels, loss, grad, dec, neg = G.ctc_case()
  np.testing.assert_allclose(loss, g['loss'], rtol=1e-12)
  np.testing.assert_allclose(grad, g['grad'], rtol=0, atol=1e-12)
  np.testing.assert_array_equal(dec[1], g['decoded_values'])
  assert g['loss'][3] > 0 and g['label_lengths'][3] == 0      # empty transcript: loss = -log p(all blank)


def test_feature_fixture():
  g = np.load(os.path.join(GOLDEN, 'features_1s.npz'))
  wav, feat = G.features()
  np.testing.assert_allclose(feat, g['feat'], rtol=0, atol=1e-10)
  f32 = O.calc_power_spectrogram(wav, 16000, dtype=np.float32)
  assert rel(f32, g['feat']) < 1e-4
