"""CTC loss + gradient (st_ctc_loss: log-softmax, alpha/beta recursion, gradient) timed alone at the config-2 and
config-4 shapes; under ncu this is the command whose launch list / source view the CTC work is read from."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from speecht_b200 import ops


def labels_for(rng, n, ctc_len):
  while True:
    lab = rng.integers(0, 28, size=n)
    if n + int(np.sum(lab[1:] == lab[:-1])) <= ctc_len:
      return lab.astype(np.int32)


def main():
  reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
  rng = np.random.default_rng(0)
  for B, To, L in ((32, 501, 150), (64, 501, 150), (32, 1501, 450), (256, 1501, 450)):
    logits = torch.randn((B, To, 32), device='cuda')[:, :, :29].transpose(0, 1)
    seq = np.full((B,), To - 1, dtype=np.int32)
    labs = [labels_for(rng, L, To - 1) for _ in range(B)]
    batch = ops.CTCBatch(labs, seq, To, 29, logits.device)
    planes = torch.empty((2, B, To, 64), dtype=torch.bfloat16, device='cuda')
    for want in ('loss+grad', 'loss'):
      fn = (lambda: ops.ctc_loss(batch, logits, want_grad=False, grad_scale=1.0 / B, grad_planes=planes)) \
        if want == 'loss+grad' else (lambda: ops.ctc_loss(batch, logits, want_grad=False))
      for _ in range(3):
        fn()
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(reps):
        fn()
      e1.record()
      torch.cuda.synchronize()
      ms = e0.elapsed_time(e1) / reps
      print('B=%3d T\'=%4d L=%3d %-9s %.4f ms  (%.0f ns per time step)' % (B, To, L, want, ms, ms * 1e6 / To), flush=True)


if __name__ == '__main__':
  main()
