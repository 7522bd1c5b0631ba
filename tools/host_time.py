"""Host-side cost of a train step: wall time the Python thread spends ENQUEUEING one step (ctypes launches, label
flattening, pinned staging) against the device time of the step.  If the first stays well under the second, the GPU
never waits for the host.  Also prints the top cumulative entries of a cProfile run of the same loop.

  python tools/host_time.py [steps]
"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speecht_b200.engine import W2LEngine  # noqa: E402


def main():
  steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
  dev = torch.device('cuda', 0)
  eng = W2LEngine(precision='bf16x3', device=dev)
  eng.init_xavier(0)
  rng = np.random.default_rng(0)
  inputs = rng.standard_normal((32, 1001, 128), dtype=np.float32)
  lengths = np.full((32,), 1001, dtype=np.int32)
  labels = [list(rng.integers(0, 28, size=150)) for _ in range(32)]
  x = torch.from_numpy(inputs).to(dev)
  for _ in range(5):
    eng.train_step(x, lengths, labels, 1e-4)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0 = time.perf_counter()
  e0.record()
  for _ in range(steps):
    eng.train_step(x, lengths, labels, 1e-4)
  e1.record()
  t_enq = time.perf_counter() - t0
  torch.cuda.synchronize()
  t_all = time.perf_counter() - t0
  print('steps %d: host enqueue %.3f ms/step, device %.3f ms/step, wall %.3f ms/step' % (
    steps, 1e3 * t_enq / steps, e0.elapsed_time(e1) / steps, 1e3 * t_all / steps))
  # host-only cost with an idle GPU in front (synchronise every step so that no call ever blocks on a full queue)
  host = []
  for _ in range(20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.train_step(x, lengths, labels, 1e-4)
    host.append(time.perf_counter() - t0)
  print('enqueue of one step on an idle stream: median %.3f ms, min %.3f ms' % (1e3 * float(np.median(host)), 1e3 * min(host)))
  prof = cProfile.Profile()
  prof.enable()
  for _ in range(steps):
    eng.train_step(x, lengths, labels, 1e-4)
  prof.disable()
  torch.cuda.synchronize()
  pstats.Stats(prof).sort_stats('cumulative').print_stats(22)


if __name__ == '__main__':
  main()
