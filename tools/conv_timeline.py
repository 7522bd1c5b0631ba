"""Per-CTA timeline of single tensor-core conv launches inside a real (pipelined, un-profiled) train step.

  python tools/conv_timeline.py [launch indices ...]      (default: 0 4 11 13 17 20)

Launch index = position among the tc_conv_kernel launches of one step: 0..10 forward of layers 0..10, 11..20 data
gradient of layers 10..1 (8, 9 and 12 use the two-phase epilogue build, which carries no stamps).  Uses the st_debug_conv_timeline hook: eight %globaltimer stamps per CTA (first tile).
Prints, per stamp, min / median / max over the CTAs in microseconds after the earliest CTA entry.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from speecht_b200._lib import check, lib, ptr
from speecht_b200.engine import W2LEngine

NAMES = ['entry', 'prev grid done', 'first operands', 'last MMA issued', 'accumulator done', 'epilogue issued',
         'staging drained', 'exit']


def main():
  targets = [int(a) for a in sys.argv[1:]] or [0, 4, 11, 13, 17, 20]
  dev = torch.device('cuda', 0)
  eng = W2LEngine(precision=os.environ.get('SPEECHT_B200_PRECISION', 'bf16x3'), device=dev)
  eng.init_xavier(seed=0)
  inputs, lengths, labels = bench.make_batch(0, 32, 10.0)
  x = torch.from_numpy(inputs).to(dev)
  for _ in range(3):
    eng.train_step(x, lengths, labels, 1e-4)
  torch.cuda.synchronize()
  buf = torch.zeros((148 * 8,), dtype=torch.int64, device=dev)
  for idx in targets:
    buf.zero_()
    torch.cuda.synchronize()
    check(lib().st_debug_conv_timeline(ptr(buf), idx))
    eng.train_step(x, lengths, labels, 1e-4)
    torch.cuda.synchronize()
    check(lib().st_debug_conv_timeline(None, -1))
    t = buf.cpu().numpy().reshape(148, 8)
    used = t[:, 0] > 0
    t = t[used].astype(np.float64)
    t0 = t[:, 0].min()
    print('launch %d: %d CTAs, span entry->exit %.1f us' % (idx, int(used.sum()), (t[:, 7].max() - t0) / 1e3))
    for k, name in enumerate(NAMES):
      col = (t[:, k] - t0) / 1e3
      col = col[t[:, k] > 0]
      if col.size:
        print('  %-18s min %7.2f  median %7.2f  max %7.2f' % (name, col.min(), np.median(col), col.max()))
    d = t[:, 3] - t[:, 2]
    print('  MMA phase (first operands -> last MMA issued): median %.2f us;  epilogue (accumulator -> issued): median '
          '%.2f us;  tail (issued -> exit): median %.2f us' % (np.median(d) / 1e3, np.median(t[:, 5] - t[:, 4]) / 1e3,
                                                              np.median(t[:, 7] - t[:, 5]) / 1e3))


if __name__ == '__main__':
  main()
