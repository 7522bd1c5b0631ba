"""CPU tests of the oracle itself: independent cross-checks (torch CPU ops, brute-force CTC, direct loops).
These pin the restatement's *mathematics*; they cannot pin TF1/librosa quirks (parity unpinned, see oracle header)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import speecht_oracle as O


def test_same_padding_matches_tf_rule():
  assert O.same_padding(1001, 48, 2) == (501, 23, 24)
  assert O.same_padding(1000, 48, 2) == (500, 23, 23)
  assert O.same_padding(501, 7, 1) == (501, 3, 3)
  assert O.same_padding(501, 32, 1) == (501, 15, 16)
  assert O.same_padding(501, 1, 1) == (501, 0, 0)
  assert O.same_padding(5, 48, 2) == (3, 23, 24)


@pytest.mark.parametrize('k,s,cin,cout,t', [(48, 2, 16, 10, 101), (48, 2, 16, 10, 100), (7, 1, 10, 10, 51),
                                             (32, 1, 10, 24, 51), (1, 1, 24, 29, 51), (32, 1, 6, 8, 5)])
def test_conv_forward_backward_vs_torch(k, s, cin, cout, t):
  rng = np.random.default_rng(1)
  x = rng.standard_normal((3, t, cin))
  w = rng.standard_normal((k, cin, cout)) * 0.1
  b = rng.standard_normal((cout,))
  y = O.conv1d_same(x, w, b, s, relu=False)
  out, left, right = O.same_padding(t, k, s)
  xt = torch.tensor(x, requires_grad=True)
  wt = torch.tensor(w, requires_grad=True)
  bt = torch.tensor(b, requires_grad=True)
  xp = F.pad(xt.permute(0, 2, 1), (left, right))
  yt = F.conv1d(xp, wt.permute(2, 1, 0), bt, stride=s).permute(0, 2, 1)
  assert yt.shape[1] == out
  np.testing.assert_allclose(y, yt.detach().numpy(), rtol=1e-10, atol=1e-10)
  dy = rng.standard_normal(y.shape)
  yt.backward(torch.tensor(dy))
  dx, dw, db = O.conv1d_same_backward(x, w, s, dy)
  np.testing.assert_allclose(dx, xt.grad.numpy(), rtol=1e-9, atol=1e-9)
  np.testing.assert_allclose(dw, wt.grad.numpy(), rtol=1e-9, atol=1e-9)
  np.testing.assert_allclose(db, bt.grad.numpy(), rtol=1e-9, atol=1e-9)


def test_ctc_vs_brute_force_enumeration():
  rng = np.random.default_rng(2)
  C = 4
  for label in ([0], [1, 1], [0, 1, 2], [2, 2, 1], []):
    T = 5
    logits = rng.standard_normal((T, 1, C))
    loss, _ = O.ctc_loss_and_grad(logits, [label], [T])
    bf = O.ctc_brute_force_loss(logits[:, 0, :], label)
    assert abs(loss[0] - bf) < 1e-10, (label, loss, bf)


def test_ctc_vs_torch_loss_and_grad():
  rng = np.random.default_rng(3)
  T, B, C = 40, 5, 29
  logits = rng.standard_normal((T, B, C)) * 2
  seq = [40, 33, 40, 25, 12]
  labels = [O.synthetic_labels(rng, n, s) for n, s in zip([10, 7, 15, 1, 5], seq)]
  labels[3] = np.array([], dtype=np.int32)          # empty transcript
  loss, grad = O.ctc_loss_and_grad(logits, labels, seq)
  lt = torch.tensor(logits, requires_grad=True)
  lsm = F.log_softmax(lt, dim=2)
  tl = F.ctc_loss(lsm, torch.tensor(np.concatenate(labels).astype(np.int64)), torch.tensor(seq),
                  torch.tensor([len(l) for l in labels]), blank=C - 1, reduction='none')
  np.testing.assert_allclose(loss, tl.detach().numpy(), rtol=1e-9, atol=1e-9)
  tl.sum().backward()
  np.testing.assert_allclose(grad, lt.grad.numpy(), rtol=1e-7, atol=1e-9)
  # frames beyond seq_len carry no gradient; in-range rows sum to ~0 over classes
  assert np.all(grad[33:, 1] == 0)
  assert np.abs(grad.sum(axis=2)).max() < 1e-9


def test_ctc_rejects_infeasible_and_bad_labels():
  logits = np.zeros((4, 1, 5))
  with pytest.raises(O.CTCLabelError):
    O.ctc_loss_and_grad(logits, [[1, 1, 1]], [4])        # needs 5 frames
  with pytest.raises(O.CTCLabelError):
    O.ctc_loss_and_grad(logits, [[4]], [4])              # blank id as label
  O.ctc_loss_and_grad(logits, [[1, 1]], [3])             # exactly enough


def test_greedy_decoder_rules():
  C = 4
  def onehot(seq):
    x = np.full((len(seq), 1, C), -1.0, dtype=np.float32)
    for t, c in enumerate(seq):
      x[t, 0, c] = 1.0
    return x
  (idx, val, shape), neg = O.ctc_greedy_decoder(onehot([0, 0, 3, 0, 1, 1, 3, 3, 2]), [9])
  assert val.tolist() == [0, 0, 1, 2] and shape.tolist() == [1, 4]
  assert idx.tolist() == [[0, 0], [0, 1], [0, 2], [0, 3]]
  assert neg[0, 0] == -9.0
  (idx, val, shape), _ = O.ctc_greedy_decoder(onehot([0, 0, 3, 0, 1, 1, 3, 3, 2]), [9], merge_repeated=False)
  assert val.tolist() == [0, 0, 0, 1, 1, 2]
  # ties: first maximum wins; only frames < seq_len count
  x = np.zeros((3, 1, C), dtype=np.float32)
  (idx, val, shape), _ = O.ctc_greedy_decoder(x, [2])
  assert val.tolist() == [0]
  # all-blank utterance produces no entries and a zero-width dense shape
  (idx, val, shape), _ = O.ctc_greedy_decoder(onehot([3, 3]), [2])
  assert idx.shape == (0, 2) and shape.tolist() == [1, 0]


def test_extract_decoded_ids_quirk():
  idx = np.array([[0, 0], [0, 1], [2, 0]])
  val = np.array([5, 6, 7])
  assert O.extract_decoded_ids(idx, val) == [[5, 6], [7]]    # utterance 1 (empty) leaves no entry


def test_adam_tf1_matches_closed_form_and_differs_from_torch():
  p = np.array([1.0, -2.0], dtype=np.float64); g = np.array([0.5, 0.25])
  m = np.zeros(2); v = np.zeros(2)
  O.adam_tf1([p], [g], [m], [v], lr=0.1, step=1)
  lr_t = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
  exp = np.array([1.0, -2.0]) - lr_t * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-3)
  np.testing.assert_allclose(p, exp, rtol=1e-12)


def test_clip_by_global_norm():
  g = [np.full((3,), 4.0), np.full((4,), 3.0)]
  out, norm = O.clip_by_global_norm(g, 5.0)
  assert abs(norm - math.sqrt(48 + 36)) < 1e-12
  np.testing.assert_allclose(out[0], g[0] * 5.0 / norm)
  out, norm = O.clip_by_global_norm([np.array([0.3])], 5.0)
  np.testing.assert_allclose(out[0], [0.3])


def test_mel_filterbank_and_spectrogram_properties():
  fb = O.mel_filterbank(16000, 512, 128)
  assert fb.shape == (128, 257) and np.all(fb >= 0)
  # Slaney area normalisation: each triangle integrates to ~1 over Hz -> sum * bin_width ~ 1
  area = fb.sum(axis=1) * (8000 / 256)
  assert np.all(np.abs(area[10:] - 1.0) < 0.35)
  rng = np.random.default_rng(4)
  wav = (0.1 * rng.standard_normal(16000)).astype(np.float32)
  feat = O.calc_power_spectrogram(wav, 16000)
  assert feat.shape == (101, 128)
  assert abs(feat.mean()) < 1e-9 and abs(feat.std() - 1) < 1e-9
  # STFT against scipy's own short-time FFT of the same reflect-padded frames
  import scipy.signal
  S = O.stft_power(wav, 512, 160)
  f, t, Z = scipy.signal.stft(np.pad(wav.astype(np.float64), 256, mode='reflect'), window='hann', nperseg=512,
                              noverlap=512 - 160, boundary=None, padded=False)
  ref = np.abs(Z * 256.0) ** 2     # scipy scales by 1/sum(window) = 1/256
  np.testing.assert_allclose(S, ref, rtol=1e-8, atol=1e-12)


def test_train_step_gradients_vs_torch_autograd_tiny_net():
  """Whole step (forward, CTC, backward) on a scaled-down layer table vs torch autograd."""
  rng = np.random.default_rng(5)
  layers = [(6, 2, 8, 12, True), (3, 1, 12, 12, True), (4, 1, 12, 16, True), (1, 1, 16, 6, False)]
  weights = O.xavier_weights(rng, layers=layers, dtype=np.float64)
  weights = [(w, rng.standard_normal(b.shape) * 0.1) for w, b in weights]
  x = rng.standard_normal((3, 21, 8))
  lengths = np.array([21, 17, 21])
  labels = [[0, 1, 1], [2], [4, 3, 2, 1]]
  logits, acts = O.wav2letter_forward(x, weights, layers=layers, keep_activations=True)
  loss, dlog = O.ctc_loss_and_grad(logits, labels, lengths // 2)
  grads = O.wav2letter_backward(acts, weights, dlog / 3, layers=layers)
  tw = [(torch.tensor(w, requires_grad=True), torch.tensor(b, requires_grad=True)) for w, b in weights]
  h = torch.tensor(x)
  for (k, s, cin, cout, relu), (w, b) in zip(layers, tw):
    out, left, right = O.same_padding(h.shape[1], k, s)
    h = F.conv1d(F.pad(h.permute(0, 2, 1), (left, right)), w.permute(2, 1, 0), b, stride=s).permute(0, 2, 1)
    if relu:
      h = torch.relu(h)
  lt = h.permute(1, 0, 2)
  np.testing.assert_allclose(logits, lt.detach().numpy(), rtol=1e-9, atol=1e-9)
  tl = F.ctc_loss(F.log_softmax(lt, 2), torch.tensor(sum(labels, [])), torch.tensor(lengths // 2),
                  torch.tensor([len(l) for l in labels]), blank=5, reduction='none')
  np.testing.assert_allclose(loss, tl.detach().numpy(), rtol=1e-9)
  tl.mean().backward()
  for (dw, db), (w, b) in zip(grads, tw):
    np.testing.assert_allclose(dw, w.grad.numpy(), rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(db, b.grad.numpy(), rtol=1e-7, atol=1e-10)


def test_features_vs_torchaudio_librosa_compatible_pipeline():
  """Second opinion on the librosa semantics the oracle restates from memory (reference preprocessing.py:50-58):
  torchaudio's MelSpectrogram(norm='slaney', mel_scale='slaney', center=True, pad_mode='reflect', power=2) is
  torchaudio's documented librosa.feature.melspectrogram equivalent and AmplitudeToDB('power', top_db=80) its
  power_to_db; ref=np.max only subtracts the maximum.  Agreement pins: reflect padding, periodic Hann, the Slaney
  mel scale + area normalisation, fmax = sr/2, amin = 1e-10, the 80 dB floor.  (Not the reference itself: librosa
  is not installable here.)"""
  torch = pytest.importorskip('torch')
  ta = pytest.importorskip('torchaudio')
  rng = np.random.default_rng(11)
  wav = (0.1 * rng.standard_normal(16000 + 123)).astype(np.float32)
  wav[4000:6000] *= 1e-4                      # a quiet stretch so that the 80 dB floor actually clips something
  old = torch.get_default_dtype()
  torch.set_default_dtype(torch.float64)      # torchaudio builds window and filterbank in the default dtype
  try:
    fb = ta.functional.melscale_fbanks(n_freqs=257, f_min=0.0, f_max=8000.0, n_mels=128, sample_rate=16000,
                                       norm='slaney', mel_scale='slaney').numpy().T
    mel = ta.transforms.MelSpectrogram(sample_rate=16000, n_fft=512, hop_length=160, n_mels=128, f_min=0.0,
                                       f_max=8000.0, power=2.0, center=True, pad_mode='reflect', norm='slaney',
                                       mel_scale='slaney')
    S = mel(torch.from_numpy(wav).double())                               # [128, T]
    db = ta.transforms.AmplitudeToDB(stype='power', top_db=80.0)(S)       # ref = 1.0, amin = 1e-10
  finally:
    torch.set_default_dtype(old)
  assert fb.dtype == np.float64
  np.testing.assert_allclose(O.mel_filterbank(16000, 512, 128), fb, rtol=1e-9, atol=1e-12)
  db = db - 10.0 * torch.log10(torch.clamp(S.max(), min=1e-10))           # ref = np.max
  x = db.numpy()
  ref = ((x - x.mean()) / x.std()).T                                      # normalize(), then .T (preprocessing.py:29-33,58)
  got = O.calc_power_spectrogram(wav, 16000)
  assert got.shape == ref.shape == (1 + len(wav) // 160, 128)
  assert (x <= x.max() - 80.0 + 1e-9).any()                               # the floor was active in this case
  np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9)


def test_same_padding_vs_transformers_tf_port():
  """Second opinion on tf.nn.conv1d(padding='SAME') (reference speech_model.py:155): Hugging Face's MobileNetV2 port
  carries `apply_tf_padding`, its restatement of TensorFlow's SAME rule (validated there against TF checkpoints).
  The oracle's (left, right) padding must equal what that function pads along the height axis, for the layer shapes
  of the network and for odd/even lengths.  (Not TensorFlow itself: TF is not installable here.)"""
  torch = pytest.importorskip('torch')
  mod = pytest.importorskip('transformers.models.mobilenet_v2.modeling_mobilenet_v2')
  for k, s in [(48, 2), (7, 1), (32, 1), (1, 1), (6, 2), (4, 3), (3, 1)]:
    conv = torch.nn.Conv2d(1, 1, kernel_size=(k, 1), stride=(s, 1))
    for t in [1, 2, 5, 6, 37, 100, 101, 501, 1001, 3001]:
      x = torch.ones((1, 1, t, 1))
      y = mod.apply_tf_padding(x, conv)
      col = y[0, 0, :, 0]
      nz = torch.nonzero(col).flatten()
      left = int(nz[0])
      right = int(col.numel() - 1 - nz[-1])
      out_len, pl, pr = O.same_padding(t, k, s)
      assert (left, right) == (pl, pr), (k, s, t, left, right, pl, pr)
      assert (t + left + right - k) // s + 1 == out_len == -(-t // s)


# ---------------------------------------------------------------------------------------------------------------
# Known-answer vectors published in TensorFlow's own kernel tests (tests/golden/tf_published_vectors.py): the only
# third-party golden data for the two CTC ops the reference delegates to TF1.  They pin blank = last class, softmax
# inside the op, the gradient definition and the greedy decoder's merge / ignore rules.
# ---------------------------------------------------------------------------------------------------------------
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), 'golden'))
import tf_published_vectors as TFV  # noqa: E402


def test_oracle_ctc_loss_matches_tensorflow_published_vectors():
  logits, targets, seq_len, loss_truth, grad_truth = TFV.ctc_case()
  loss, grad = O.ctc_loss_and_grad(logits, targets, seq_len)          # blank defaults to the LAST class
  np.testing.assert_allclose(loss, loss_truth, rtol=0, atol=5e-6)     # literals carry six digits
  np.testing.assert_allclose(grad, grad_truth, rtol=0, atol=2e-6)
  # torch's CTC (blank=5, log-softmax applied by hand) is a second independent witness of the same literals ...
  lp = torch.log_softmax(torch.tensor(logits), dim=-1)
  lens = torch.tensor([len(r) for r in targets])

  def torch_ctc(label_rows, blank):
    return F.ctc_loss(lp, torch.tensor([t for row in label_rows for t in row]), torch.tensor(seq_len), lens,
                      blank=blank, reduction='none').numpy()

  tl = torch_ctc(targets, 5)
  # ... and the other common convention (blank = class 0, what warp-ctc / torch default to) cannot reproduce them
  assert np.max(np.abs(torch_ctc([[t + 1 for t in row] for row in targets], 0) - loss_truth)) > 0.1
  np.testing.assert_allclose(tl, loss_truth, rtol=0, atol=5e-6)


def test_oracle_greedy_decoder_matches_tensorflow_published_vectors():
  logits, seq_len, indices, values, shape, neg = TFV.greedy_case()
  (ri, rv, rs), rneg = O.ctc_greedy_decoder(logits, seq_len, merge_repeated=True)
  np.testing.assert_array_equal(ri, indices)
  np.testing.assert_array_equal(rv, values)
  np.testing.assert_array_equal(rs, shape)
  np.testing.assert_allclose(rneg, neg, rtol=1e-6)


def test_oracle_adam_is_tensorflows_adam_update_numpy():
  rng = np.random.default_rng(3)
  p0 = rng.standard_normal(7); g = rng.standard_normal(7)
  m0 = rng.standard_normal(7) * 0.1; v0 = rng.random(7) * 0.1
  for t in (1, 2, 50):
    p, m, v = [p0.copy()], [m0.copy()], [v0.copy()]
    O.adam_tf1(p, [g], m, v, lr=1e-4, step=t, eps=1e-3)
    pt, mt, vt = TFV.adam_update_numpy(p0, g, t, m0, v0, alpha=1e-4, epsilon=1e-3)
    np.testing.assert_allclose(p[0], pt, rtol=1e-13)
    np.testing.assert_allclose(m[0], mt, rtol=1e-13)
    np.testing.assert_allclose(v[0], vt, rtol=1e-13)


def test_oracle_conv_matches_tensorflow_published_vectors():
  """conv_ops_test.py literals restated as the 1-D problems they contain: SAME with a stride (padding column on the
  right), filter layout, stride-2 data and filter gradients -- integer valued, so exact."""
  for name, x, w, stride, expected in TFV.conv_forward_cases():
    y = O.conv1d_same(x, w, np.zeros(w.shape[2]), stride, False)
    assert y.shape == expected.shape, name
    np.testing.assert_array_equal(y, expected, err_msg=name)
  for name, x, w, stride, dy, fold, dx_lit, dw_lit in TFV.conv_backward_cases():
    dx, dw, db = O.conv1d_same_backward(x, w, stride, dy)
    np.testing.assert_array_equal(fold(dx), dx_lit, err_msg=name)
    np.testing.assert_array_equal(dw, dw_lit, err_msg=name)
  # the padding column on the LEFT (the other way to make SAME) cannot produce the stride-2 literals
  name, x, w, stride, expected = TFV.conv_forward_cases()[1]
  flipped = O.conv1d_same(x[:, ::-1], w[::-1], np.zeros(3), stride, False)[:, ::-1]
  assert np.max(np.abs(flipped - expected)) > 100


def test_oracle_clip_matches_tensorflow_published_vector():
  clipped, norm = O.clip_by_global_norm(TFV.CLIP_INPUTS, TFV.CLIP_NORM)
  assert norm == TFV.CLIP_GLOBAL_NORM
  for c, e in zip(clipped, TFV.CLIP_OUTPUTS):
    np.testing.assert_allclose(c, e, rtol=1e-15)


def test_oracle_mel_scale_matches_librosa_published_docstring_values():
  hz, mel = TFV.LIBROSA_HZ_TO_MEL
  np.testing.assert_allclose(O._hz_to_mel_slaney(np.array(hz)), mel, rtol=0, atol=1e-12)
  mel, hz = TFV.LIBROSA_MEL_TO_HZ
  np.testing.assert_allclose(O._mel_to_hz_slaney(np.array(mel)), hz, rtol=0, atol=5e-4)
  edges = O._mel_to_hz_slaney(np.linspace(O._hz_to_mel_slaney(0.0), O._hz_to_mel_slaney(11025.0), 40))
  np.testing.assert_allclose(edges, TFV.LIBROSA_MEL_FREQUENCIES_40, rtol=0, atol=5e-4)   # three printed decimals


def test_oracle_mel_filter_normalisation_matches_librosa_docstring_values():
  """The two filterbank entries librosa's own documentation prints (weak: two significant digits, but an
  un-normalised triangle has 0.40 where the Slaney-normalised one has 0.016)."""
  for case in (TFV.LIBROSA_MEL_FILTER_0_1, TFV.LIBROSA_MEL_FILTER_0_1_FMAX8000):
    fb = O.mel_filterbank(case['sr'], case['n_fft'], case['n_mels'], fmax=case.get('fmax'))
    assert fb.shape == (case['n_mels'], 1 + case['n_fft'] // 2)
    assert round(float(fb[0, 1]), case['decimals']) == case['value'], fb[0, :3]
    assert abs(float(fb[0, 0])) == 0.0 and fb[-1, -1] == 0.0
