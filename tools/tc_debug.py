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
  for precision in sys.argv[3:] or ['bf16x3', 'bf16']:
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
    for l, ((dw, db), (rdw, rdb)) in enumerate(zip(eng2.weight_grads, grads)):
      print('  layer %2d dW rel err %.3e   db rel err %.3e' % (l, rel(dw.cpu().numpy(), rdw), rel(db.cpu().numpy(), rdb)))


if __name__ == '__main__':
  main()
