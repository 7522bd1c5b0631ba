"""GPU parity tests, operator level: every C-ABI entry point against the CPU oracle on the same seeded inputs.
Bars: integer outputs bit-exact; floating point within 1e-4 of max|ref| per tensor (BASELINE.md metric) unless a
tighter bound is stated.  Run on the B200 box:  python -m pytest tests -m gpu"""
import os

import numpy as np
import pytest
import torch

from oracle import speecht_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def rel(a, b):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def dev(x, dtype=torch.float32):
  return torch.from_numpy(np.ascontiguousarray(x)).to('cuda', dtype=dtype)


@pytest.fixture(scope='module')
def ops():
  from speecht_b200 import ops as _ops
  return _ops


# ----------------------------------------------------------------------------------------------- greedy decode
@pytest.mark.parametrize('T,B,C,seed', [(51, 4, 29, 0), (501, 32, 29, 1), (7, 3, 29, 2), (1, 1, 29, 3), (300, 5, 5, 4)])
def test_greedy_decode_bit_exact(ops, T, B, C, seed):
  rng = np.random.default_rng(seed)
  logits = rng.standard_normal((T, B, C)).astype(np.float32)
  logits[rng.random((T, B)) < 0.4, C - 1] += 4.0           # plenty of blanks
  logits[1::3] = logits[0:-1:3][:logits[1::3].shape[0]]     # and repeated frames
  seq = rng.integers(0, T + 1, size=B).astype(np.int32)
  seq[0] = T
  (ri, rv, rs), rneg = O.ctc_greedy_decoder(logits, seq)
  # batch-major storage viewed time-major, exactly what the engine hands over
  store = dev(logits.transpose(1, 0, 2))
  dec, neg = ops.ctc_greedy_decoder(store.transpose(0, 1), seq)
  np.testing.assert_array_equal(dec[0].indices, ri)
  np.testing.assert_array_equal(dec[0].values, rv)
  np.testing.assert_array_equal(dec[0].dense_shape, rs)
  assert dec[0].values.dtype == np.int64 and dec[0].indices.dtype == np.int64
  np.testing.assert_allclose(neg, rneg, rtol=1e-5, atol=1e-4)
  dec2, _ = ops.ctc_greedy_decoder(dev(logits), seq, merge_repeated=False)
  (_, rv2, _), _ = O.ctc_greedy_decoder(logits, seq, merge_repeated=False)
  np.testing.assert_array_equal(dec2[0].values, rv2)


def test_greedy_decode_ties_and_all_blank(ops):
  x = np.zeros((6, 2, 29), dtype=np.float32)                # exact ties everywhere: first max (class 0) wins
  x[:, 1, 28] = 1.0                                         # utterance 1 is all blank
  dec, _ = ops.ctc_greedy_decoder(dev(x), [6, 6])
  assert dec[0].values.tolist() == [0] and dec[0].indices.tolist() == [[0, 0]]
  assert dec[0].dense_shape.tolist() == [2, 1]
  dec, _ = ops.ctc_greedy_decoder(dev(x), [0, 6])           # empty + all blank -> no entries at all
  assert dec[0].values.shape == (0,) and dec[0].dense_shape.tolist() == [2, 0]


def test_greedy_decode_golden(ops):
  g = np.load(os.path.join(GOLDEN, 'ctc_case.npz'))
  dec, neg = ops.ctc_greedy_decoder(dev(g['logits']), g['seq'])
  np.testing.assert_array_equal(dec[0].values, g['decoded_values'])
  np.testing.assert_array_equal(dec[0].indices, g['decoded_indices'])
  np.testing.assert_allclose(neg, g['neg_sum_logits'], rtol=1e-5, atol=1e-4)


# ----------------------------------------------------------------------------------------------- CTC
def _ctc_case(rng, T, B, C, scale=2.0, chars=None):
  logits = (rng.standard_normal((T, B, C)) * scale).astype(np.float32)
  seq = rng.integers(max(1, T // 2), T + 1, size=B).astype(np.int32)
  seq[0] = T
  labels = [O.synthetic_labels(rng, int(rng.integers(0, max(1, s // 3))) if chars is None else chars, int(s))
            for s in seq]
  return logits, seq, labels


@pytest.mark.parametrize('T,B,seed,chars', [(51, 4, 0, 15), (60, 6, 1, None), (501, 8, 2, 150), (5, 2, 3, 1),
                                            (1501, 2, 4, 450)])
def test_ctc_loss_and_grad(ops, T, B, seed, chars):
  rng = np.random.default_rng(seed)
  logits, seq, labels = _ctc_case(rng, T, B, 29, chars=chars)
  if chars is not None:
    seq[:] = T
  rloss, rgrad = O.ctc_loss_and_grad(logits, labels, seq)
  store = dev(logits.transpose(1, 0, 2))                    # batch-major storage, time-major view
  loss, grad = ops.ctc_loss(labels, store.transpose(0, 1), seq)
  np.testing.assert_allclose(loss.cpu().numpy(), rloss, rtol=1e-5, atol=1e-4)
  g = grad.cpu().numpy()
  assert g.shape == (T, B, 29)
  assert rel(g, rgrad) < 1e-4, rel(g, rgrad)
  for b in range(B):
    assert np.all(g[seq[b]:, b] == 0)
  assert np.abs(g.sum(axis=2)).max() < 1e-4                 # softmax - occupancy sums to 0 per frame
  # grad_scale folds reduce_mean
  _, g2 = ops.ctc_loss(labels, dev(logits), seq, grad_scale=0.125)
  assert rel(g2.cpu().numpy(), rgrad * 0.125) < 1e-4


def test_ctc_golden_and_empty_label(ops):
  g = np.load(os.path.join(GOLDEN, 'ctc_case.npz'))
  lens = g['label_lengths']
  offs = np.concatenate([[0], np.cumsum(lens)])
  labels = [g['labels'][offs[i]:offs[i + 1]] for i in range(len(lens))]
  loss, grad = ops.ctc_loss(labels, dev(g['logits']), g['seq'])
  np.testing.assert_allclose(loss.cpu().numpy(), g['loss'], rtol=1e-5, atol=1e-4)
  assert rel(grad.cpu().numpy(), g['grad']) < 1e-4
  assert lens[3] == 0 and loss[3].item() > 0


def test_ctc_rejects_like_tf(ops):
  from speecht_b200._lib import CTCLabelError
  logits = dev(np.zeros((4, 1, 29), np.float32))
  with pytest.raises(CTCLabelError, match='Not enough time'):
    ops.ctc_loss([[1, 1, 1]], logits, [4])
  with pytest.raises(CTCLabelError):
    ops.ctc_loss([[28]], logits, [4])
  with pytest.raises(CTCLabelError):
    ops.ctc_loss([[1]], logits, [5])
  # device-side flag when host validation is skipped: loss = +inf, zero gradient, no crash
  loss, grad = ops.ctc_loss([[1, 1, 1]], logits, [4], validate=False)
  assert np.isinf(loss.cpu().numpy()[0]) and np.all(grad.cpu().numpy() == 0)
  loss, _ = ops.ctc_loss([[1, 1]], logits, [3])             # exactly enough frames
  assert np.isfinite(loss.cpu().numpy()[0])


def test_ctc_bf16_planes_reconstruct_gradient(ops):
  rng = np.random.default_rng(9)
  logits, seq, labels = _ctc_case(rng, 40, 3, 29, chars=8)
  store = dev(logits.transpose(1, 0, 2))
  planes = torch.full((2, 3, 40, 32), 7.0, dtype=torch.bfloat16, device='cuda')
  loss, grad = ops.ctc_loss(labels, store.transpose(0, 1), seq, grad_planes=planes)
  rec = planes.float().sum(0)[:, :, :29].permute(1, 0, 2).cpu().numpy()
  assert rel(rec, grad.cpu().numpy()) < 2e-5                # hi+lo split keeps ~16 mantissa bits
  assert torch.all(planes[:, :, :, 29:] == 0)


# ----------------------------------------------------------------------------------------------- conv (fp32 path)
LAYER_SHAPES = [(48, 2, 128, 250), (7, 1, 250, 250), (32, 1, 250, 2000), (1, 1, 2000, 2000), (1, 1, 2000, 29)]


@pytest.mark.parametrize('k,s,cin,cout', LAYER_SHAPES + [(3, 1, 5, 7), (6, 2, 9, 130), (4, 3, 17, 33)])
@pytest.mark.parametrize('T', [101, 100, 5])
def test_conv_forward_fp32(ops, k, s, cin, cout, T):
  rng = np.random.default_rng(k * 1000 + cin + T)
  B = 3
  x = rng.standard_normal((B, T, cin)).astype(np.float32)
  w = (rng.standard_normal((k, cin, cout)) / np.sqrt(k * cin)).astype(np.float32)
  b = rng.standard_normal((cout,)).astype(np.float32)
  for relu in (True, False):
    ref = O.conv1d_same(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), s, relu)
    y = ops.conv1d(dev(x), dev(w), dev(b), stride=s, relu=relu).cpu().numpy()
    assert y.shape == ref.shape
    assert rel(y, ref) < 1e-5, rel(y, ref)
  y = ops.conv1d(dev(x), dev(w), None, stride=s, relu=False).cpu().numpy()
  assert rel(y, O.conv1d_same(x.astype(np.float64), w.astype(np.float64), np.zeros(cout), s, False)) < 1e-5


@pytest.mark.parametrize('k,s,cin,cout', LAYER_SHAPES + [(3, 1, 5, 7), (6, 2, 9, 130), (4, 3, 17, 33)])
@pytest.mark.parametrize('T', [101, 6])
def test_conv_backward_fp32(ops, k, s, cin, cout, T):
  rng = np.random.default_rng(k * 77 + cout + T)
  B = 2
  x = rng.standard_normal((B, T, cin)).astype(np.float32)
  w = (rng.standard_normal((k, cin, cout)) / np.sqrt(k * cin)).astype(np.float32)
  b = rng.standard_normal((cout,)).astype(np.float32)
  y = O.conv1d_same(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), s, True)
  dy = rng.standard_normal(y.shape).astype(np.float32)
  rdx, rdw, rdb = O.conv1d_same_backward(x.astype(np.float64), w.astype(np.float64), s, dy * (y > 0))
  d_y = dev(y)
  dx = ops.conv1d_backprop_input(dev(dy), dev(w), (B, T, cin), stride=s, y_act=d_y).cpu().numpy()
  dw, db = ops.conv1d_backprop_filter(dev(x), dev(dy), k, stride=s, y_act=d_y)
  assert rel(dx, rdx) < 1e-5 and rel(dw.cpu().numpy(), rdw) < 1e-5 and rel(db.cpu().numpy(), rdb) < 1e-5
  # without the fused mask
  rdx, rdw, rdb = O.conv1d_same_backward(x.astype(np.float64), w.astype(np.float64), s, dy.astype(np.float64))
  dx = ops.conv1d_backprop_input(dev(dy), dev(w), (B, T, cin), stride=s).cpu().numpy()
  dw, db = ops.conv1d_backprop_filter(dev(x), dev(dy), k, stride=s)
  assert rel(dx, rdx) < 1e-5 and rel(dw.cpu().numpy(), rdw) < 1e-5 and rel(db.cpu().numpy(), rdb) < 1e-5


def test_conv_linearity_at_full_size(ops):
  """Size-independent property at BASELINE config-2 size (B=32, T'=501, the k32 250->2000 layer):
  conv(a*x1 + x2) == a*conv(x1) + conv(x2) without bias/relu."""
  g = torch.Generator(device='cuda').manual_seed(0)
  x1 = torch.randn((32, 501, 250), device='cuda', generator=g)
  x2 = torch.randn((32, 501, 250), device='cuda', generator=g)
  w = torch.randn((32, 250, 2000), device='cuda', generator=g) / 90.0
  y1 = ops.conv1d(x1, w); y2 = ops.conv1d(x2, w)
  y12 = ops.conv1d(2.5 * x1 + x2, w)
  err = (y12 - (2.5 * y1 + y2)).abs().max().item() / y12.abs().max().item()
  assert err < 1e-5, err


# ----------------------------------------------------------------------------------------------- clip + Adam
def test_clip_adam_matches_tf1_semantics(ops):
  rng = np.random.default_rng(5)
  n = 100003
  p = rng.standard_normal(n).astype(np.float32); g = (rng.standard_normal(n) * 0.05).astype(np.float32)
  m = (rng.standard_normal(n) * 0.01).astype(np.float32); v = (rng.random(n) * 0.01).astype(np.float32)
  for clip, step in ((5.0, 1), (1.0, 7)):
    rp, rm, rv = p.astype(np.float64), m.astype(np.float64), v.astype(np.float64)
    (cg,), norm = O.clip_by_global_norm([g.astype(np.float64)], clip)
    O.adam_tf1([rp], [cg], [rm], [rv], lr=1e-2, step=step)
    dp, dg, dm, dv = dev(p), dev(g), dev(m), dev(v)
    nsq = ops.global_norm_sq(dg)
    assert abs(np.sqrt(nsq.item()) - norm) < 1e-6 * norm
    ops.clip_adam(dp, dg, dm, dv, step, 1e-2, max_norm=clip, normsq=nsq)
    assert rel(dp.cpu().numpy(), rp) < 1e-6 and rel(dm.cpu().numpy(), rm) < 1e-6 and rel(dv.cpu().numpy(), rv) < 1e-6


# ----------------------------------------------------------------------------------------------- features
def test_power_spectrogram_vs_oracle_and_golden(ops):
  g = np.load(os.path.join(GOLDEN, 'features_1s.npz'))
  rng = np.random.default_rng(11)
  wav = (0.1 * rng.standard_normal(16000)).astype(np.float32)
  feat, frames = ops.power_spectrogram(dev(wav[None]), [16000], 16000)
  assert frames.tolist() == [101]
  assert rel(feat[0].cpu().numpy(), g['feat']) < 1e-4
  # ragged batch incl. odd frame count, 22050 Hz (what librosa.load gives the reference), zero batch padding
  lens = [16000, 12345, 4000]
  wavs = np.zeros((3, 16000), np.float32)
  refs = []
  for i, n in enumerate(lens):
    wavs[i, :n] = 0.05 * rng.standard_normal(n) + 0.02 * np.sin(np.arange(n) * 0.05)
    refs.append(O.calc_power_spectrogram(wavs[i, :n], 22050))
  feat, frames = ops.power_spectrogram(dev(wavs), lens, 22050)
  for i, r in enumerate(refs):
    assert frames[i].item() == r.shape[0]
    assert rel(feat[i, :r.shape[0]].cpu().numpy(), r) < 1e-4
    assert torch.all(feat[i, r.shape[0]:] == 0)
  with pytest.raises(ValueError):
    ops.power_spectrogram(dev(wavs[:, :200]), [200, 200, 200], 16000)


# ----------------------------------------------------------------------------------------------- TF published vectors
def test_ctc_kernels_reproduce_tensorflow_published_test_vectors(ops):
  """The literals of TensorFlow's own ctc_loss / ctc_greedy_decoder kernel tests (tests/golden/
  tf_published_vectors.py), through the C ABI: pins blank = last class, softmax inside the op, the gradient and the
  decoder's merge / ignore-beyond-length rules on the CUDA path itself, not only on the oracle."""
  import sys
  sys.path.insert(0, GOLDEN)
  import tf_published_vectors as TFV
  logits, targets, seq_len, loss_truth, grad_truth = TFV.ctc_case()
  loss, grad = ops.ctc_loss(targets, dev(logits), np.asarray(seq_len, np.int32), want_grad=True)
  np.testing.assert_allclose(loss.cpu().numpy(), loss_truth, rtol=0, atol=5e-6)
  np.testing.assert_allclose(grad.cpu().numpy(), grad_truth, rtol=0, atol=2e-6)
  glogits, gseq, indices, values, shape, neg = TFV.greedy_case()
  dec, gneg = ops.ctc_greedy_decoder(dev(glogits), np.asarray(gseq, np.int32), merge_repeated=True)
  np.testing.assert_array_equal(dec[0].indices, indices)
  np.testing.assert_array_equal(dec[0].values, values)
  np.testing.assert_array_equal(dec[0].dense_shape, shape)
  np.testing.assert_allclose(gneg, neg, rtol=1e-6)


def test_conv_kernels_reproduce_tensorflow_published_test_vectors(ops):
  """conv_ops_test.py literals (tests/golden/tf_published_vectors.py) through the C ABI: SAME with a stride (padding
  column on the right), the [width, in, out] filter layout, stride-2 data and filter gradients.  Integer valued and
  small: fp32 is exact."""
  import sys
  sys.path.insert(0, GOLDEN)
  import tf_published_vectors as TFV
  for name, x, w, stride, expected in TFV.conv_forward_cases():
    y = ops.conv1d(dev(x), dev(w), None, stride=stride, relu=False).cpu().numpy()
    assert y.shape == expected.shape, name
    np.testing.assert_array_equal(y, expected, err_msg=name)
  for name, x, w, stride, dy, fold, dx_lit, dw_lit in TFV.conv_backward_cases():
    dx = ops.conv1d_backprop_input(dev(dy), dev(w), x.shape, stride=stride).cpu().numpy()
    dw, db = ops.conv1d_backprop_filter(dev(x), dev(dy), w.shape[0], stride=stride)
    np.testing.assert_array_equal(fold(dx), dx_lit, err_msg=name)
    np.testing.assert_array_equal(dw.cpu().numpy(), dw_lit, err_msg=name)
  # clip_ops_test.py::testClipByGlobalNormClipped: global norm 5
  flat = dev(np.concatenate([a.ravel() for a in TFV.CLIP_INPUTS]))
  assert ops.global_norm_sq(flat).item() == TFV.CLIP_GLOBAL_NORM ** 2
