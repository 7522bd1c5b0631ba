"""Known-answer vectors PUBLISHED in TensorFlow's own unit tests for the two CTC ops on the hot path.

The reference (louiskirsch/speechT) delegates `tf.nn.ctc_loss` (speech_model.py:74) and `tf.nn.ctc_greedy_decoder`
(speech_model.py:113) to TensorFlow 1.x, which is neither vendored under /root/reference nor installable here, and
its own tests hold no value for this path.  The only third-party golden data for these two ops are
the literals of TensorFlow's kernel tests (tensorflow/python/kernel_tests/ctc_loss_op_test.py::testBasic and
ctc_decoder_ops_test.py::testCTCGreedyDecoder, TF 1.x tree; the CTC numbers also circulate in warp-ctc / PyTorch
tests).  They are typed in here from the published test source -- NOT produced by running TensorFlow -- and they are
self-authenticating: six-digit loss and gradient literals that the independent numpy oracle reproduces to the
rounding of the literals (4e-7) cannot agree by accident.

What they pin ([L] items of SURVEY.md 8a that were resting on memory):
  * blank label = LAST class (depth 6 -> class 5; depth 4 -> class 3);
  * the op applies softmax itself (inputs are log-probabilities of rows that already sum to 1, and the gradient
    literal equals softmax - occupancy);
  * repeated labels need a blank between them (targets_1 = [0,1,1,0]: the -0.797544 sits in the blank column);
  * greedy decode: per-frame argmax, merge repeats BEFORE dropping blanks, frames t >= seq_len ignored,
    neg_sum_logits = -sum of the per-frame maxima.
"""
import numpy as np

# ---------------------------------------------------------------------------------------------- ctc_loss (depth 6)
CTC_DEPTH = 6
CTC_SEQ_LEN = [5, 5]
CTC_TARGETS = [[0, 1, 2, 1, 0], [0, 1, 1, 0]]
CTC_LOSS = np.array([3.34211, 5.42262])                     # -loss_log_prob_{0,1}

CTC_PROB_0 = np.array(
    [[0.633766, 0.221185, 0.0917319, 0.0129757, 0.0142857, 0.0260553],
     [0.111121, 0.588392, 0.278779, 0.0055756, 0.00569609, 0.010436],
     [0.0357786, 0.633813, 0.321418, 0.00249248, 0.00272882, 0.0037688],
     [0.0663296, 0.643849, 0.280111, 0.00283995, 0.0035545, 0.00331533],
     [0.458235, 0.396634, 0.123377, 0.00648837, 0.00903441, 0.00623107]])
CTC_GRAD_0 = np.array(
    [[-0.366234, 0.221185, 0.0917319, 0.0129757, 0.0142857, 0.0260553],
     [0.111121, -0.411608, 0.278779, 0.0055756, 0.00569609, 0.010436],
     [0.0357786, 0.633813, -0.678582, 0.00249248, 0.00272882, 0.0037688],
     [0.0663296, -0.356151, 0.280111, 0.00283995, 0.0035545, 0.00331533],
     [-0.541765, 0.396634, 0.123377, 0.00648837, 0.00903441, 0.00623107]])
CTC_PROB_1 = np.array(
    [[0.30176, 0.28562, 0.0831517, 0.0862751, 0.0816851, 0.161508],
     [0.24082, 0.397533, 0.0557226, 0.0546814, 0.0557528, 0.19549],
     [0.230246, 0.450868, 0.0389607, 0.038309, 0.0391602, 0.202456],
     [0.280884, 0.429522, 0.0326593, 0.0339046, 0.0326856, 0.190345],
     [0.423286, 0.315517, 0.0338439, 0.0393744, 0.0339315, 0.154046]])
CTC_GRAD_1 = np.array(
    [[-0.69824, 0.28562, 0.0831517, 0.0862751, 0.0816851, 0.161508],
     [0.24082, -0.602467, 0.0557226, 0.0546814, 0.0557528, 0.19549],
     [0.230246, 0.450868, 0.0389607, 0.038309, 0.0391602, -0.797544],
     [0.280884, -0.570478, 0.0326593, 0.0339046, 0.0326856, 0.190345],
     [-0.576714, 0.315517, 0.0338439, 0.0393744, 0.0339315, 0.154046]])


def ctc_case():
  """-> (logits [T=5, B=2, C=6] float64 log-probabilities, targets, seq_len, loss [2], grad [5,2,6])."""
  logits = np.stack([np.log(CTC_PROB_0), np.log(CTC_PROB_1)], axis=1)
  grad = np.stack([CTC_GRAD_0, CTC_GRAD_1], axis=1)
  return logits, CTC_TARGETS, CTC_SEQ_LEN, CTC_LOSS, grad


# ---------------------------------------------------------------------------------------------- greedy decoder (depth 4)
GREEDY_SEQ_LEN = [4, 5]
GREEDY_PROB_0 = np.array([[1.0, 0.0, 0.0, 0.0],
                          [0.0, 0.0, 0.4, 0.6],
                          [0.0, 0.0, 0.4, 0.6],
                          [0.0, 0.9, 0.1, 0.0],
                          [0.0, 0.0, 0.0, 0.0],     # t = 4, 5: beyond seq_len_0 = 4, ignored
                          [0.0, 0.0, 0.0, 0.0]])
GREEDY_PROB_1 = np.array([[0.1, 0.9, 0.0, 0.0],
                          [0.0, 0.9, 0.1, 0.0],
                          [0.0, 0.0, 0.1, 0.9],
                          [0.0, 0.9, 0.1, 0.1],
                          [0.9, 0.1, 0.0, 0.0],
                          [0.0, 0.0, 0.0, 0.0]])     # t = 5: beyond seq_len_1 = 5, ignored
GREEDY_INDICES = np.array([[0, 0], [0, 1], [1, 0], [1, 1], [1, 2]], dtype=np.int64)
GREEDY_VALUES = np.array([0, 1, 1, 1, 0], dtype=np.int64)
GREEDY_SHAPE = np.array([2, 3], dtype=np.int64)
GREEDY_NEG_SUM_LOGITS = np.array([[np.sum(-np.log([1.0, 0.6, 0.6, 0.9]))],
                                  [np.sum(-np.log([0.9, 0.9, 0.9, 0.9, 0.9]))]])


def greedy_case():
  """-> (logits [T=6, B=2, C=4] float32 log-probabilities (log 0 = -inf like the TF test), seq_len, indices, values,
  dense_shape, neg_sum_logits [2,1])."""
  with np.errstate(divide='ignore'):
    logits = np.stack([np.log(GREEDY_PROB_0), np.log(GREEDY_PROB_1)], axis=1).astype(np.float32)
  return logits, GREEDY_SEQ_LEN, GREEDY_INDICES, GREEDY_VALUES, GREEDY_SHAPE, GREEDY_NEG_SUM_LOGITS


# ---------------------------------------------------------------------------------------------- Adam (formula only)
def adam_update_numpy(param, g_t, t, m, v, alpha=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
  """The numpy restatement TensorFlow's adam_test.py checks tf.train.AdamOptimizer against (no literals there):
  epsilon is added to the UNcorrected sqrt(v_t), the bias correction lives in alpha_t."""
  alpha_t = alpha * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
  m_t = beta1 * m + (1 - beta1) * g_t
  v_t = beta2 * v + (1 - beta2) * g_t * g_t
  return param - alpha_t * m_t / (np.sqrt(v_t) + epsilon), m_t, v_t


# ---------------------------------------------------------------------------------------------- conv2d (as conv1d)
# Literals of tensorflow/python/kernel_tests/conv_ops_test.py (TF 1.x; `_VerifyValues` / `_RunAndVerifyBackprop*`
# fill input, filter and output-gradient with 1, 2, 3, ... in row-major order).  `tf.nn.conv1d` -- what the reference
# calls (speech_model.py:139-147) -- IS conv2d on a height-1 image, so every case below is restated as the 1-D
# problem it contains: a filter as tall as the rows it touches turns those rows into extra input channels; an
# output row becomes a batch entry.  Typed in from the published test source, not produced by running TensorFlow;
# integer-valued, so the match is exact.  What they pin: the SAME rule with a stride (output length ceil(T / s),
# the odd padding column goes to the RIGHT -- the 1230 / 1305 / 1380 triple below exists only under that rule),
# the [width, in, out] filter layout, and the stride-2 data / filter gradients.

def _seq(shape):
  return np.arange(1, int(np.prod(shape)) + 1, dtype=np.float64).reshape(shape)


def conv_forward_cases():
  """-> list of (name, x [B,T,Cin], w [K,Cin,Cout], stride, expected [B,T',Cout]) -- all under SAME padding."""
  cases = []
  # testConv2D1x1Filter: input [1,2,3,3], filter [1,1,3,3]
  cases.append(('1x1Filter', _seq((1, 6, 3)), _seq((1, 3, 3)), 1,
                np.array([30.0, 36.0, 42.0, 66.0, 81.0, 96.0, 102.0, 126.0, 150.0, 138.0, 171.0, 204.0, 174.0, 216.0,
                          258.0, 210.0, 261.0, 312.0]).reshape(1, 6, 3)))
  # testConv2D2x2FilterStride2Same: input [1,2,3,3], filter [2,2,3,3], strides [2,2], SAME.  Height: in 2, filter 2,
  # one output row, no padding -> the two rows are six channels; width: in 3 -> out 2, ONE padding column, right.
  x = _seq((1, 2, 3, 3)).transpose(0, 2, 1, 3).reshape(1, 3, 6)        # [B, W, (h, c)]
  w = _seq((2, 2, 3, 3)).transpose(1, 0, 2, 3).reshape(2, 6, 3)        # [kw, (kh, c), out]
  cases.append(('2x2FilterStride2Same', x, w, 2,
                np.array([2271.0, 2367.0, 2463.0, 1230.0, 1305.0, 1380.0]).reshape(1, 2, 3)))
  # testConv2DKernelSmallerThanStrideSame: [1,3,3,1] and [1,4,4,1], filter [1,1,1,1], strides [2,2], SAME: the image
  # rows the vertical stride selects (0 and 2) are batch entries
  cases.append(('KernelSmallerThanStrideSame3', _seq((3, 3, 1))[::2], _seq((1, 1, 1)), 2,
                np.array([1.0, 3.0, 7.0, 9.0]).reshape(2, 2, 1)))
  cases.append(('KernelSmallerThanStrideSame4', _seq((4, 4, 1))[::2], _seq((1, 1, 1)), 2,
                np.array([1.0, 3.0, 9.0, 11.0]).reshape(2, 2, 1)))
  # testConv2D2x2FilterStride1x2: input [1,3,6,1], filter [2,2,1,1], strides [1,2], VALID.  Width 6, filter 2,
  # stride 2 needs no padding (SAME == VALID); output row r reads image rows r, r+1 as two channels
  img = _seq((3, 6))
  x = np.stack([np.stack([img[r], img[r + 1]], axis=-1) for r in range(2)])          # [2, 6, 2]
  w = _seq((2, 2)).T.reshape(2, 2, 1)                                              # [kw, kh, 1]
  cases.append(('2x2FilterStride1x2', x, w, 2,
                np.array([58.0, 78.0, 98.0, 118.0, 138.0, 158.0]).reshape(2, 3, 1)))
  return cases


def conv_backward_cases():
  """-> list of (name, x [B,T,Cin], w [K,Cin,Cout], stride, dy [B,T',Cout], fold(dx) -> literal dx or None,
  expected dx (or None), expected dw [K,Cin,Cout] (or None))."""
  cases = []
  # testConv2D2x2Depth3ValidBackprop{Input,Filter}Stride1x2: input [1,3,6,1], filter [2,2,1,1], output [1,2,3,1],
  # strides [1,2] (the 1-D problem of '2x2FilterStride1x2' above).  dx of batch entry r lands on image rows r, r+1.
  img = _seq((3, 6))
  x = np.stack([np.stack([img[r], img[r + 1]], axis=-1) for r in range(2)])
  w = _seq((2, 2)).T.reshape(2, 2, 1)
  dy = _seq((2, 3, 1))

  def fold(dx):
    out = np.zeros((3, 6))
    for r in range(2):
      out[r] += dx[r, :, 0]
      out[r + 1] += dx[r, :, 1]
    return out.ravel()

  dx_lit = np.array([1.0, 2.0, 2.0, 4.0, 3.0, 6.0, 7.0, 12.0, 11.0, 18.0, 15.0, 24.0, 12.0, 16.0, 15.0, 20.0, 18.0,
                     24.0])
  dw_lit = np.array([161.0, 182.0, 287.0, 308.0]).reshape(2, 2).T.reshape(2, 2, 1)  # literal is [kh, kw]
  cases.append(('2x2Depth3ValidBackpropStride1x2', x, w, 2, dy, fold, dx_lit, dw_lit))
  # testConv2DStrideTwoFilterOneSameBackprop{Input,Filter}: input [1,4,4,1], filter [1,1,1,1], output [1,2,2,1],
  # strides [2,2], SAME.  Image rows 0 and 2 are the batch entries; rows 1 and 3 receive no gradient.
  x = _seq((4, 4, 1))[::2]
  w = _seq((1, 1, 1))
  dy = _seq((2, 2, 1))

  def fold2(dx):
    out = np.zeros((4, 4))
    out[0] = dx[0, :, 0]
    out[2] = dx[1, :, 0]
    return out.ravel()

  dx_lit = np.array([1.0, 0.0, 2.0, 0.0, 0.0, 0.0, 0.0, 0.0, 3.0, 0.0, 4.0, 0.0, 0.0, 0.0, 0.0, 0.0])
  cases.append(('StrideTwoFilterOneSameBackprop', x, w, 2, dy, fold2, dx_lit, np.array([78.0]).reshape(1, 1, 1)))
  return cases


# ---------------------------------------------------------------------------------------------- clip_by_global_norm
# tensorflow/python/kernel_tests/clip_ops_test.py::testClipByGlobalNormClipped: global norm 5, clip_norm 4.
CLIP_INPUTS = [np.array([[-2.0, 0.0, 0.0], [4.0, 0.0, 0.0]]), np.array([1.0, -2.0])]
CLIP_NORM = 4.0
CLIP_GLOBAL_NORM = 5.0
CLIP_OUTPUTS = [np.array([[-1.6, 0.0, 0.0], [3.2, 0.0, 0.0]]), np.array([0.8, -1.6])]


# ---------------------------------------------------------------------------------------------- librosa mel scale
# The reference's features are librosa.feature.melspectrogram (preprocessing.py:51), not installable here.  The
# literals below are the examples printed in librosa's own docstrings (librosa/core/convert.py, 0.5-0.10):
# `librosa.mel_frequencies(n_mels=40)` (fmin 0, fmax 11025, htk=False), `hz_to_mel(60)`, `hz_to_mel([110, 220, 440])`,
# `mel_to_hz([1, 2, 3, 4, 5])`.  Typed in from the published documentation.  Forty six-digit band edges pin the
# Slaney scale (linear below 1 kHz at 200/3 Hz per mel, log above with step log(6.4)/27) that places the 130 edges of
# the reference's 128 filters; they do not pin the triangle construction or the Slaney area normalisation.
LIBROSA_MEL_FREQUENCIES_40 = np.array(
    [0., 85.317, 170.635, 255.952, 341.269, 426.586, 511.904, 597.221, 682.538, 767.855, 853.173, 938.49, 1024.856,
     1119.114, 1222.042, 1334.436, 1457.167, 1591.187, 1737.532, 1897.337, 2071.84, 2262.393, 2470.47, 2697.686,
     2945.799, 3216.731, 3512.582, 3835.643, 4188.417, 4573.636, 4994.285, 5453.621, 5955.205, 6502.92, 7101.009,
     7754.107, 8467.272, 9246.028, 10096.408, 11025.])
LIBROSA_HZ_TO_MEL = ([60.0, 110.0, 220.0, 440.0], [0.9, 1.65, 3.3, 6.6])
LIBROSA_MEL_TO_HZ = ([1.0, 2.0, 3.0, 4.0, 5.0], [66.667, 133.333, 200.0, 266.667, 333.333])
# librosa.filters.mel docstring (0.5-0.10): `librosa.filters.mel(sr=22050, n_fft=2048)` prints
# array([[ 0.   ,  0.016, ...,  0.   ,  0.   ], ...]) and, clipped to fmax=8000, array([[ 0.  ,  0.02, ...,  0.  ,  0.  ], ...]).
# Two significant digits only -- but an un-normalised triangle would have 0.40 there, so they do pin the Slaney area
# normalisation 2 / (f[i+2] - f[i]) that turns it into 0.016.
LIBROSA_MEL_FILTER_0_1 = {'sr': 22050, 'n_fft': 2048, 'n_mels': 128, 'value': 0.016, 'decimals': 3}
LIBROSA_MEL_FILTER_0_1_FMAX8000 = {'sr': 22050, 'n_fft': 2048, 'n_mels': 128, 'fmax': 8000.0, 'value': 0.02, 'decimals': 2}
