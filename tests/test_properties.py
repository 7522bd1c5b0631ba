"""Property tests (hypothesis): CTC and greedy-decode invariants on the oracle (CPU) and, on a GPU, the CUDA kernels
against the oracle over randomly drawn shapes, lengths and label sets (SURVEY.md section 4's implied test plan)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import speecht_oracle as O


@st.composite
def ctc_case(draw, max_t=24, max_b=4, classes=(3, 29)):
  C = draw(st.sampled_from(classes))
  T = draw(st.integers(1, max_t))
  B = draw(st.integers(1, max_b))
  seed = draw(st.integers(0, 2 ** 31 - 1))
  rng = np.random.default_rng(seed)
  scale = draw(st.sampled_from([0.1, 1.0, 5.0]))
  logits = (rng.standard_normal((T, B, C)) * scale).astype(np.float32)
  seq = rng.integers(0, T + 1, size=B).astype(np.int32)
  labels = []
  for b in range(B):
    want = int(rng.integers(0, max(1, seq[b] // 2 + 1)))
    lab = rng.integers(0, C - 1, size=want)
    while len(lab) + int(np.sum(lab[1:] == lab[:-1])) > seq[b]:
      lab = lab[:-1]
    labels.append(lab.astype(np.int32))
  return logits, seq, labels


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(ctc_case())
def test_oracle_ctc_invariants(case):
  logits, seq, labels = case
  loss, grad = O.ctc_loss_and_grad(logits, labels, seq)
  assert np.all(loss >= -1e-9)                                  # -log p with p <= 1
  assert np.abs(grad.sum(axis=2)).max() < 1e-8                  # softmax - occupancy sums to zero over classes
  for b in range(len(seq)):
    assert np.all(grad[seq[b]:, b] == 0)                        # frames beyond the sequence carry no gradient
    if seq[b] > 0 and seq[b] <= 6 and logits.shape[2] <= 3:
      bf = O.ctc_brute_force_loss(logits[:seq[b], b], list(labels[b]))
      assert abs(loss[b] - bf) < 1e-8


@settings(max_examples=60, deadline=None)
@given(ctc_case())
def test_oracle_greedy_decode_invariants(case):
  logits, seq, _ = case
  C = logits.shape[2]
  (idx, val, shape), neg = O.ctc_greedy_decoder(logits, seq)
  assert np.all(val != C - 1) and np.all(val >= 0)               # never emits the blank
  rows = [val[idx[:, 0] == b] for b in range(len(seq))]
  for b, r in enumerate(rows):
    assert len(r) <= seq[b]
    assert np.array_equal(idx[idx[:, 0] == b, 1], np.arange(len(r)))
  assert shape[1] == max([len(r) for r in rows] + [0])
  # decoding is idempotent on its own one-hot rendering (labels separated by blanks)
  for b, r in enumerate(rows):
    if len(r):
      onehot = np.full((2 * len(r), 1, C), -1.0, np.float32)
      for i, c in enumerate(r):
        onehot[2 * i, 0, c] = 1.0
        onehot[2 * i + 1, 0, C - 1] = 1.0
      (_, v2, _), _ = O.ctc_greedy_decoder(onehot, [2 * len(r)])
      assert np.array_equal(v2, r)


@pytest.mark.gpu
@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(ctc_case(max_t=70, max_b=6, classes=(29,)))
def test_gpu_ctc_and_decode_match_oracle_on_random_cases(case):
  import torch
  from speecht_b200 import ops
  logits, seq, labels = case
  T, B, C = logits.shape
  store = torch.from_numpy(np.ascontiguousarray(logits.transpose(1, 0, 2))).cuda()
  view = store.transpose(0, 1)
  dec, neg = ops.ctc_greedy_decoder(view, seq)
  (ri, rv, rs), rneg = O.ctc_greedy_decoder(logits, seq)
  np.testing.assert_array_equal(dec[0].values, rv)
  np.testing.assert_array_equal(dec[0].indices, ri)
  np.testing.assert_array_equal(dec[0].dense_shape, rs)
  rloss, rgrad = O.ctc_loss_and_grad(logits, labels, seq)
  loss, grad = ops.ctc_loss(labels, view, seq)
  np.testing.assert_allclose(loss.cpu().numpy(), rloss, rtol=1e-5, atol=1e-4)
  g = grad.cpu().numpy()
  assert np.max(np.abs(g - rgrad)) <= 1e-4 * max(np.max(np.abs(rgrad)), 1e-3)
  assert np.abs(g.sum(axis=2)).max() < 1e-4
