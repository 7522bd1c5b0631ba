"""Prints per-layer forward / gradient errors of the tensor-core path against the oracle (diagnostics, GPU box)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import speecht_oracle as O
from speecht_b200.engine import W2LEngine


def rel(a, b):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def rms(a, b):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.sqrt(np.mean(b ** 2)), 1e-30))


def bf16_round(x):
  return torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).to(torch.float32).numpy()


def accumulation_probe():
  """Operands exactly representable in bf16 -> the hi/lo split is exact and the only error left is the fp32
  accumulation inside the tensor core (round-to-nearest: ~1e-7*sqrt(n); truncation: biased, ~n*3e-8)."""
  inputs, lengths, labels = O.synthetic_batch(seed=3, batch=2, seconds=1)
  inputs = bf16_round(inputs)
  weights = O.xavier_weights(np.random.default_rng(99), dtype=np.float32)
  weights = [(bf16_round(w), b) for w, b in weights]
  w64 = [(w.astype(np.float64), b.astype(np.float64)) for w, b in weights]
  ref0 = O.conv1d_same(inputs.astype(np.float64), w64[0][0], w64[0][1], 2, True)
  for precision in ('bf16x6', 'bf16x3', 'bf16', 'fp32'):
    eng = W2LEngine(precision=precision)
    eng.load_weights(weights)
    eng.forward(torch.from_numpy(inputs).cuda(), keep_activations=True)
    a = eng._tc().activation(0).cpu().numpy() if precision != 'fp32' else eng._acts[1].cpu().numpy()
    d = a.astype(np.float64) - ref0
    pos = ref0 > 1e-3
    print('accumulation probe layer0 [%s]: max rel %.3e rms rel %.3e  mean signed rel err on positive outputs %.3e'
          % (precision, rel(a, ref0), rms(a, ref0), float(np.mean(d[pos] / ref0[pos]))))


def main():
  batch = int(sys.argv[1]) if len(sys.argv) > 1 else 3
  seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 1
  inputs, lengths, labels = O.synthetic_batch(seed=3, batch=batch, seconds=seconds)
  weights = O.xavier_weights(np.random.default_rng(99), dtype=np.float32)
  weights = [(w, (0.01 * np.random.default_rng(i).standard_normal(b.shape)).astype(np.float32))
             for i, (w, b) in enumerate(weights)]
  w64 = [(w.astype(np.float64), b.astype(np.float64)) for w, b in weights]
  logits, acts = O.wav2letter_forward(inputs.astype(np.float64), w64, keep_activations=True)
  loss, dlog = O.ctc_loss_and_grad(logits, labels, lengths // 2)
  grads = O.wav2letter_backward(acts, w64, dlog / batch)
  accumulation_probe()
  for precision in sys.argv[3:] or ['bf16x6', 'bf16x3', 'bf16']:
    eng = W2LEngine(precision=precision)
    eng.load_weights(weights)
    out = eng.forward(torch.from_numpy(inputs).cuda(), keep_activations=True)
    torch.cuda.synchronize()
    plan = eng._tc()
    print('== %s forward' % precision)
    for l in range(10):
      a = plan.activation(l).cpu().numpy()
      print('  layer %2d out  rel err %.3e   (max|ref| %.3f, nonzero frac %.3f)' % (
        l, rel(a, acts[l + 1]), np.abs(acts[l + 1]).max(), float((a != 0).mean())))
    print('  logits        rel err %.3e' % rel(out.cpu().numpy(), logits))
    eng2 = W2LEngine(precision=precision)
    eng2.load_weights(weights)
    res = eng2.train_step(torch.from_numpy(inputs).cuda(), lengths, labels, 1e-4)
    torch.cuda.synchronize()
    print('== %s backward: loss rel err %.3e' % (precision, rel(res['loss'].cpu().numpy(), loss)))
    # oracle backward with the ReLU masks the GPU forward produced (a sign flip of a ~1e-6 pre-activation changes
    # a gradient element by O(1): compare like with like)
    plan2 = eng2._tc()
    acts_h = [acts[0]]
    for l in range(10):
      g = plan2.activation(l).cpu().numpy() > 0
      acts_h.append(np.where(g, np.maximum(acts[l + 1], 1e-30), 0.0))
    acts_h.append(acts[11])
    grads_h = O.wav2letter_backward(acts_h, w64, dlog / batch)
    for l, ((dw, db), (rdw, rdb), (hdw, hdb)) in enumerate(zip(eng2.weight_grads, grads, grads_h)):
      print('  layer %2d dW max-rel %.3e rms-rel %.3e | same-mask: dW max-rel %.3e rms-rel %.3e db max-rel %.3e' % (
        l, rel(dw.cpu().numpy(), rdw), rms(dw.cpu().numpy(), rdw), rel(dw.cpu().numpy(), hdw),
        rms(dw.cpu().numpy(), hdw), rel(db.cpu().numpy(), hdb)))


if __name__ == '__main__':
  main()
