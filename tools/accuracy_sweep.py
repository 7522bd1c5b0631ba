"""Per-layer activation error of the tensor-core path against the exact-fp32 CUDA path (itself 2e-6 from the float64
oracle) over several seeds at BASELINE config-2 size -- how much margin the 1e-4 gate has.  GPU box diagnostics."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from speecht_b200.engine import W2LEngine


def main():
  B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
  T = int(sys.argv[2]) if len(sys.argv) > 2 else 1001
  worst = {}
  for seed in range(int(sys.argv[3]) if len(sys.argv) > 3 else 5):
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.randn((B, T, 128), device='cuda', generator=g)
    ref = W2LEngine(precision='fp32'); ref.init_xavier(seed=seed)
    ref.forward(x, keep_activations=True)
    ref_acts = [a.clone() for a in ref._acts[1:11]] + [ref._logits_bm.clone()]
    for precision in ('bf16x6', 'bf16x3', 'bf16'):
      eng = W2LEngine(precision=precision); eng.init_xavier(seed=seed)
      out = eng.forward(x, keep_activations=True)
      errs = []
      for l in range(10):
        a = eng._tc().activation(l)
        errs.append(((a - ref_acts[l]).abs().max() / ref_acts[l].abs().max()).item())
      lg = out.transpose(0, 1)
      errs.append(((lg - ref_acts[10]).abs().max() / ref_acts[10].abs().max()).item())
      print('seed %d %-7s' % (seed, precision), ' '.join('%.2e' % e for e in errs), flush=True)
      worst[precision] = max(worst.get(precision, 0.0), max(errs))
      del eng
    del ref, ref_acts
    torch.cuda.empty_cache()
  print('worst', worst)


if __name__ == '__main__':
  main()
