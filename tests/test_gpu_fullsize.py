"""GPU parity at BASELINE.json's full sizes, through properties / cross-checks that do not need the (slow) CPU
oracle for the conv stack: the tensor-core path against the exact-fp32 CUDA path on identical inputs, and the
CTC loss + greedy decode against the oracle on the GPU's own logits."""
import numpy as np
import pytest
import torch

from oracle import speecht_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def _engines(*precisions, seed=0):
  from speecht_b200.engine import W2LEngine
  engs = []
  for p in precisions:
    e = W2LEngine(precision=p)
    e.init_xavier(seed=seed)
    engs.append(e)
  return engs


def test_config2_batch32_10s_tensor_path_matches_fp32_path():
  """configs[1]: batch 32 x 10 s (T=1001 -> T'=501).  bf16x3 logits / loss vs the exact-fp32 path, labels equal."""
  inputs, lengths, labels = O.synthetic_batch(seed=11, batch=32, seconds=10)
  x = torch.from_numpy(inputs).cuda()
  tc, ref = _engines('bf16x3', 'fp32')
  a = tc.evaluate_step(x, lengths, labels)
  b = ref.evaluate_step(x, lengths, labels)
  assert a['logits'].shape == (501, 32, 29)
  assert rel(a['logits'].cpu().numpy(), b['logits'].cpu().numpy()) < 1e-4
  assert rel(a['loss'].cpu().numpy(), b['loss'].cpu().numpy()) < 1e-4
  # Greedy labels: the two paths sum in different orders, so frames whose two largest logits are closer than the
  # arithmetic tolerance may pick the other class (report such flips, do not reseed around
  # them).  Every differing frame must be such a near-tie, and there may only be a handful in 16032 frames.
  la, lb = a['logits'].cpu().numpy(), b['logits'].cpu().numpy()
  flips = np.argwhere(la.argmax(axis=2) != lb.argmax(axis=2))
  scale = np.abs(lb).max()
  for t, bb in flips:
    top2 = np.sort(lb[t, bb])[-2:]
    assert top2[1] - top2[0] < 2e-4 * scale, ('argmax flip that is not a near-tie', t, bb, top2)
  assert len(flips) <= 8, len(flips)
  print('near-tie argmax flips bf16x3 vs fp32 at config 2: %d of %d frames' % (len(flips), la.shape[0] * la.shape[1]))
  if len(flips) == 0:
    np.testing.assert_array_equal(a['decoded'][0].values, b['decoded'][0].values)
    np.testing.assert_array_equal(a['decoded'][0].indices, b['decoded'][0].indices)
  # CTC loss and greedy decode of the GPU's own logits against the CPU oracle at full size (bit-exact)
  logits = a['logits'].cpu().numpy()
  oloss, _ = O.ctc_loss_and_grad(logits, labels, lengths // 2)
  assert rel(a['loss'].cpu().numpy(), oloss) < 1e-5
  (oi, ov, osh), _ = O.ctc_greedy_decoder(logits, lengths // 2)
  np.testing.assert_array_equal(a['decoded'][0].values, ov)
  np.testing.assert_array_equal(a['decoded'][0].indices, oi)


def test_config2_train_step_gradients_tensor_path_vs_fp32_path():
  """One train step at batch 32 x 10 s on both GPU paths from identical weights: global gradient norm and every
  layer's filter gradient agree.  The per-layer bound is rms-relative 3e-2, not 1e-4: a forward difference of
  eps flips the ReLU of a fraction ~eps*density of the units (pre-activation within eps of zero), and removing a
  fraction f of random-sign terms from a gradient sum changes it by ~sqrt(f) -- sqrt(1e-4) = 1e-2 is what is
  measured at layer 0.  With the on/off pattern held equal the tensor path agrees to 3e-4
  (tests/test_gpu_model.py::test_train_step_parity)."""
  inputs, lengths, labels = O.synthetic_batch(seed=12, batch=32, seconds=10)
  x = torch.from_numpy(inputs).cuda()
  tc, ref = _engines('bf16x3', 'fp32')
  ra = tc.train_step(x, lengths, labels, 1e-4)
  rb = ref.train_step(x, lengths, labels, 1e-4)
  assert abs(ra['avg_loss'].item() - rb['avg_loss'].item()) < 1e-4 * abs(rb['avg_loss'].item())
  na, nb = tc.grad_norm(), ref.grad_norm()
  assert abs(na - nb) < 2e-3 * nb, (na, nb)
  for li, ((dw, db), (rw, rbias)) in enumerate(zip(tc.weight_grads, ref.weight_grads)):
    err = (dw - rw).norm().item() / rw.norm().item()
    assert err < 3e-2, (li, err)
    errb = (db - rbias).norm().item() / max(rbias.norm().item(), 1e-30)
    assert errb < 3e-2, (li, errb)
  for (wa, _), (wb, _) in zip(tc.export_weights(), ref.export_weights()):
    assert rel(wa, wb) < 1e-4


def test_config4_batch32_30s_bf16_train_step_runs_and_matches_bf16x3_loss():
  """configs[3] per-GPU shape: batch 32 x 30 s (T=3001 -> T'=1501), bf16 conv stack + fp32 CTC."""
  inputs, lengths, labels = O.synthetic_batch(seed=13, batch=32, seconds=30)
  x = torch.from_numpy(inputs).cuda()
  lo, hi = _engines('bf16', 'bf16x3')
  a = lo.train_step(x, lengths, labels, 1e-4, decode=True)
  b = hi.evaluate_step(x, lengths, labels)
  assert a['logits'].shape == (1501, 32, 29) and lo.global_step == 1
  la, lb = a['loss'].cpu().numpy(), b['loss'].cpu().numpy()
  assert np.all(np.isfinite(la)) and rel(la, lb) < 2e-2       # bf16: reported, not gated
  assert np.isfinite(lo.grad_norm())


def test_config5_batch256_variable_length_greedy_decode():
  """configs[4]: evaluate, batch 256, variable 1-30 s utterances, greedy CTC decode; decode + loss of the GPU
  logits checked against the oracle at full size; repeated evaluation is bit-identical (deterministic forward)."""
  rng = np.random.default_rng(5)
  secs = rng.integers(1, 31, size=256).tolist()
  secs[0] = 30
  inputs, lengths, labels = O.synthetic_batch(seed=14, batch=256, seconds=secs)
  x = torch.from_numpy(inputs).cuda()
  (tc,) = _engines('bf16x3')
  a = tc.evaluate_step(x, lengths, labels)
  logits = a['logits'].cpu().numpy()
  assert logits.shape == (1501, 256, 29)
  (oi, ov, osh), oneg = O.ctc_greedy_decoder(logits, lengths // 2)
  np.testing.assert_array_equal(a['decoded'][0].values, ov)
  np.testing.assert_array_equal(a['decoded'][0].indices, oi)
  np.testing.assert_array_equal(a['decoded'][0].dense_shape, osh)
  np.testing.assert_allclose(a['neg_sum_logits'], oneg, rtol=1e-4, atol=1e-3)
  oloss, _ = O.ctc_loss_and_grad(logits, labels, lengths // 2)
  assert rel(a['loss'].cpu().numpy(), oloss) < 1e-5
  first = torch.from_numpy(logits).cuda()                    # the logits tensor is a view of the plan's arena
  b = tc.evaluate_step(x, lengths, labels)
  assert torch.equal(first, b['logits'])
  np.testing.assert_array_equal(a['decoded'][0].values, b['decoded'][0].values)


@pytest.mark.parametrize('precision', ['bf16x3', 'fp32'])
def test_full_length_10s_every_layer_against_the_float64_oracle(precision):
  """Closes the chain at FULL LENGTH (T=1001 mel frames, T'=501): activations of all 11 layers, logits, CTC loss and
  greedy labels of the tensor-core path directly against the float64 numpy oracle -- not against the repo's own fp32
  CUDA path.  Batch 2 keeps the oracle at a few seconds; the per-row arithmetic (K = 32 x 250 and 2000-deep
  contractions, 501-row time axis with its ragged last tile) is that of config 2."""
  from speecht_b200.engine import W2LEngine
  inputs, lengths, labels = O.synthetic_batch(seed=21, batch=2, seconds=10)
  weights = O.xavier_weights(np.random.default_rng(77), dtype=np.float32)
  w64 = [(w.astype(np.float64), b.astype(np.float64)) for w, b in weights]
  logits64, acts64 = O.wav2letter_forward(inputs.astype(np.float64), w64, keep_activations=True)
  ref = O.evaluate_step(inputs, lengths, labels, weights, dtype=np.float64)
  eng = W2LEngine(precision=precision)
  eng.load_weights(weights)
  res = eng.evaluate_step(torch.from_numpy(inputs).cuda(), lengths, labels)
  if precision == 'fp32':
    eng.forward(torch.from_numpy(inputs).cuda(), keep_activations=True)
    gpu_acts = [a.cpu().numpy() for a in eng._acts[1:11]]
  else:
    gpu_acts = [eng._tc().activation(l).cpu().numpy() for l in range(10)]
  errs = [rel(a, acts64[l + 1]) for l, a in enumerate(gpu_acts)]
  errs.append(rel(res['logits'].cpu().numpy(), logits64))
  print('%s vs float64 oracle at T=1001: per-layer rel err %s' % (precision, ' '.join('%.2e' % e for e in errs)))
  assert res['logits'].shape == (501, 2, 29)
  assert max(errs) < 1e-4, errs
  assert rel(res['loss'].cpu().numpy(), ref['loss']) < 1e-4
  # greedy labels: bit-exact unless a frame's two best logits are closer than the arithmetic tolerance (reported)
  lg, lo = res['logits'].cpu().numpy(), ref['logits']
  flips = np.argwhere(lg.argmax(axis=2) != lo.argmax(axis=2))
  for t, b in flips:
    top2 = np.sort(lo[t, b])[-2:]
    assert top2[1] - top2[0] < 2e-4 * np.abs(lo).max(), ('argmax flip that is not a near-tie', t, b, top2)
  if len(flips) == 0:
    np.testing.assert_array_equal(res['decoded'][0].values, ref['decoded'][1])
    np.testing.assert_array_equal(res['decoded'][0].indices, ref['decoded'][0])
  assert len(flips) <= 1, flips


def test_evaluate_step_device_equals_evaluate_step_on_ragged_batch():
  """bench.py --config 5 times evaluate_step_device (results left on the device): same loss, same label rows as the
  host-facing evaluate_step, on a ragged 1-4 s batch."""
  from speecht_b200 import ops
  from speecht_b200.engine import W2LEngine
  inputs, lengths, labels = O.synthetic_batch(seed=3, batch=6, seconds=[1, 4, 2, 3, 1, 2])
  eng = W2LEngine(precision='bf16x3')
  eng.init_xavier(seed=2)
  x = torch.from_numpy(inputs).cuda()
  host = eng.evaluate_step(x, lengths, labels)
  host_loss = host['loss'].cpu().numpy()
  host_vals, host_idx = host['decoded'][0].values.copy(), host['decoded'][0].indices.copy()
  To = (inputs.shape[1] + 1) // 2
  batch = ops.CTCBatch(labels, lengths // 2, To, 29, x.device)
  loss, values, counts, _neg = eng.evaluate_step_device(x, batch)
  np.testing.assert_array_equal(loss.cpu().numpy(), host_loss)
  sp = ops.sparse_from_rows(values, counts)
  np.testing.assert_array_equal(sp.values, host_vals)
  np.testing.assert_array_equal(sp.indices, host_idx)
