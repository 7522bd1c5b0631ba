"""Runs st_ctc_loss a few times at config-2 size (T'=501, B=32, 150-char labels) -- target for `ncu -k regex:ctc_`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from speecht_b200 import ops
rng = np.random.default_rng(0)
T, B, C = 501, 32, 29
logits = torch.randn((B, T, 32), device='cuda')[:, :, :C].transpose(0, 1)
labels = []
for _ in range(B):
  while True:
    lab = rng.integers(0, 28, size=150)
    if 150 + int(np.sum(lab[1:] == lab[:-1])) <= T:
      break
  labels.append(lab.astype(np.int32))
batch = ops.CTCBatch(labels, [T] * B, T, C, logits.device)
for _ in range(4):
  ops.ctc_loss(batch, logits, want_grad=True)
torch.cuda.synchronize()
print('ok')
