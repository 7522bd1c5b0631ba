"""Writes a synthetic preprocessed corpus in the reference's .npz layout (preprocessing.py:178,199-206):
  python tools/make_synth_data.py <data_dir> [n_utterances] [split]
N(0,1) 128-mel features of 1-3 s with random feasible transcripts -- for CLI smoke runs without LibriSpeech."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
  data_dir = sys.argv[1]
  n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
  split = sys.argv[3] if len(sys.argv) > 3 else 'train'
  out = os.path.join(data_dir, 'preprocessed-power', split)
  os.makedirs(out, exist_ok=True)
  rng = np.random.default_rng(0)
  for i in range(n):
    secs = int(rng.integers(1, 4))
    T = 1 + 16000 * secs // 160
    feats = rng.standard_normal((T, 128)).astype(np.float32)
    while True:
      lab = rng.integers(0, 28, size=10 * secs)
      if len(lab) + int(np.sum(lab[1:] == lab[:-1])) <= T // 2:
        break
    np.savez(os.path.join(out, 'synth-%04d' % i), audio_fragments=feats, transcript=lab)
  print('wrote %d samples to %s' % (n, out))


if __name__ == '__main__':
  main()
