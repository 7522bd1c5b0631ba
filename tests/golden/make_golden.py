"""Generates tests/golden/*.npz from the oracle (NOT from the reference: TF1/librosa cannot be imported here,
see oracle/speecht_oracle.py header -- "parity unpinned").  Re-run:  python tests/golden/make_golden.py

The fixtures freeze the oracle's float64 outputs on seeded inputs so that (a) an accidental change to the oracle is
caught by the CPU suite and (b) the GPU parity tests have a second, committed anchor besides the live oracle.
Weights are NOT stored (98.7 MB): they are regenerated from numpy's PCG64 stream, which is stable across versions.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import speecht_oracle as O  # noqa: E402


def config1(dtype=np.float64):
  """BASELINE.json configs[0]: evaluate --step-count 1 on 4 synthetic 1 s utterances."""
  inputs, lengths, labels = O.synthetic_batch(seed=0, batch=4, seconds=1)
  weights = O.xavier_weights(np.random.default_rng(1234), dtype=np.float32)
  res = O.evaluate_step(inputs, lengths, labels, weights, dtype=dtype)
  return inputs, lengths, labels, weights, res


def ragged(dtype=np.float64):
  """Variable-length batch (1,2,1,3 s): padding tail is NOT masked."""
  inputs, lengths, labels = O.synthetic_batch(seed=7, batch=4, seconds=[1, 2, 1, 3])
  weights = O.xavier_weights(np.random.default_rng(1234), dtype=np.float32)
  res = O.evaluate_step(inputs, lengths, labels, weights, dtype=dtype)
  return inputs, lengths, labels, weights, res


def features():
  rng = np.random.default_rng(11)
  wav = (0.1 * rng.standard_normal(16000)).astype(np.float32)
  return wav, O.calc_power_spectrogram(wav, 16000)


def ctc_case():
  rng = np.random.default_rng(21)
  T, B, C = 60, 6, 29
  logits = (rng.standard_normal((T, B, C)) * 3).astype(np.float32)
  seq = np.array([60, 51, 60, 30, 2, 44])
  labels = [O.synthetic_labels(rng, n, s) for n, s in zip([20, 11, 29, 0, 1, 22], seq)]
  loss, grad = O.ctc_loss_and_grad(logits, labels, seq)
  dec, neg = O.ctc_greedy_decoder(logits, seq)
  return logits, seq, labels, loss, grad, dec, neg


def main():
  inputs, lengths, labels, _, res = config1()
  np.savez_compressed(os.path.join(HERE, 'config1_eval.npz'), lengths=lengths,
                      labels=np.array(labels), logits=res['logits'], loss=res['loss'],
                      decoded_indices=res['decoded'][0], decoded_values=res['decoded'][1],
                      decoded_shape=res['decoded'][2], neg_sum_logits=res['neg_sum_logits'],
                      inputs_checksum=np.float64(inputs.astype(np.float64).sum()))
  inputs, lengths, labels, _, res = ragged()
  np.savez_compressed(os.path.join(HERE, 'ragged_eval.npz'), lengths=lengths,
                      labels=np.concatenate(labels), label_lengths=np.array([len(l) for l in labels]),
                      logits=res['logits'], loss=res['loss'], decoded_indices=res['decoded'][0],
                      decoded_values=res['decoded'][1], decoded_shape=res['decoded'][2])
  wav, feat = features()
  np.savez_compressed(os.path.join(HERE, 'features_1s.npz'), feat=feat)
  logits, seq, labels, loss, grad, dec, neg = ctc_case()
  np.savez_compressed(os.path.join(HERE, 'ctc_case.npz'), logits=logits, seq=seq,
                      labels=np.concatenate(labels), label_lengths=np.array([len(l) for l in labels]),
                      loss=loss, grad=grad, decoded_indices=dec[0], decoded_values=dec[1], decoded_shape=dec[2],
                      neg_sum_logits=neg)


if __name__ == '__main__':
  main()
