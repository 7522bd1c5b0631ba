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
