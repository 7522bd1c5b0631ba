"""world_size-2 gloo tests (CPU) of the data-parallel host logic: sharding, gradient scaling + allreduce giving the
single-process gradient, identical updates on every rank, decode gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world, port, out_dir):
  sys.path.insert(0, ROOT)
  os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1',
                    MASTER_PORT=str(port))
  from oracle import speecht_oracle as O
  from speecht_b200 import parallel
  r, l, w = parallel.init_from_env('gloo')
  assert (r, w) == (rank, world)
  # a small conv+CTC model on the oracle: each rank differentiates its shard, scaled by 1/(B_local*world)
  rng = np.random.default_rng(0)
  layers = [(6, 2, 8, 12, True), (3, 1, 12, 12, True), (1, 1, 12, 6, False)]
  weights = O.xavier_weights(rng, layers=layers, dtype=np.float64)
  x = rng.standard_normal((4, 21, 8))
  lengths = np.array([21, 21, 21, 21])
  labels = [[0, 1, 1], [2], [4, 3, 2, 1], [1, 0]]
  a, b = parallel.shard_batch(4, rank, world)
  logits, acts = O.wav2letter_forward(x[a:b], weights, layers=layers, keep_activations=True)
  loss, dlog = O.ctc_loss_and_grad(logits, labels[a:b], lengths[a:b] // 2)
  grads = O.wav2letter_backward(acts, weights, dlog * parallel.gradient_scale(b - a, world), layers=layers)
  flat = torch.from_numpy(np.concatenate([g.ravel() for pair in grads for g in pair]))
  n = flat.numel()
  for h in parallel.allreduce_flat(flat, buckets=[(0, n // 2), (n // 2, n)]):
    h.wait()
  avg = parallel.mean_scalar(torch.tensor(loss.mean()))
  rows = parallel.gather_decoded([[rank, rank + 1]] * (b - a))
  t = parallel.max_scalar(1.0 + rank)
  # tensor-based gather of the sparse decode triple (config-5 evaluate at N > 1): rank 0 decodes rows [7,8],[],[9],
  # rank 1 decodes nothing in its first row and [5] in its second
  from speecht_b200.ops import SparseTensorValue
  if rank == 0:
    sp = SparseTensorValue(np.array([[0, 0], [0, 1], [2, 0]], np.int64), np.array([7, 8, 9], np.int64),
                           np.array([3, 2], np.int64))
  else:
    sp = SparseTensorValue(np.array([[1, 0]], np.int64), np.array([5], np.int64), np.array([2, 1], np.int64))
  g = parallel.gather_decoded_sparse(sp)
  # the per-step termination word of the training loop travels over a SEPARATE host (gloo) group (training.py):
  # a subgroup created next to the default one, CPU tensor, logical OR over ranks
  flag_group = dist.new_group(backend='gloo')
  any_true = [parallel.any_rank_true(rank == 1, device='cpu', group=flag_group),
              parallel.any_rank_true(False, device='cpu', group=flag_group)]
  same = parallel.identical_across_ranks(flat)                       # the reduced gradient is identical everywhere
  differs = parallel.identical_across_ranks(torch.full((5,), float(rank)))
  np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), flat=flat.numpy(), avg=avg.numpy(), rows=np.array(rows), t=t,
           g_idx=g.indices, g_val=g.values, g_shape=g.dense_shape, same=same, differs=differs,
           any_true=np.array(any_true))
  dist.destroy_process_group()


def test_sharding_is_contiguous_and_balanced():
  sys.path.insert(0, ROOT)
  from speecht_b200.parallel import shard_batch
  for n, w in ((32, 8), (7, 2), (5, 8), (256, 8)):
    cuts = [shard_batch(n, r, w) for r in range(w)]
    assert cuts[0][0] == 0 and cuts[-1][1] == n
    assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
    sizes = [b - a for a, b in cuts]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_two_rank_gradient_allreduce_matches_single_process(tmp_path):
  port = _free_port()
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  from oracle import speecht_oracle as O
  rng = np.random.default_rng(0)
  layers = [(6, 2, 8, 12, True), (3, 1, 12, 12, True), (1, 1, 12, 6, False)]
  weights = O.xavier_weights(rng, layers=layers, dtype=np.float64)
  x = rng.standard_normal((4, 21, 8))
  lengths = np.array([21, 21, 21, 21])
  labels = [[0, 1, 1], [2], [4, 3, 2, 1], [1, 0]]
  logits, acts = O.wav2letter_forward(x, weights, layers=layers, keep_activations=True)
  loss, dlog = O.ctc_loss_and_grad(logits, labels, lengths // 2)
  grads = O.wav2letter_backward(acts, weights, dlog / 4, layers=layers)
  ref = np.concatenate([g.ravel() for pair in grads for g in pair])
  r0 = np.load(tmp_path / 'rank0.npz'); r1 = np.load(tmp_path / 'rank1.npz')
  np.testing.assert_allclose(r0['flat'], ref, rtol=1e-10, atol=1e-13)
  np.testing.assert_array_equal(r0['flat'], r1['flat'])           # identical on every rank -> identical Adam update
  assert abs(float(r0['avg']) - loss.mean()) < 1e-12
  assert r0['rows'].tolist() == [[0, 1], [0, 1], [1, 2], [1, 2]] == r1['rows'].tolist()
  assert float(r0['t']) == 2.0 == float(r1['t'])
  for r in (r0, r1):
    assert r['g_idx'].tolist() == [[0, 0], [0, 1], [2, 0], [4, 0]]
    assert r['g_val'].tolist() == [7, 8, 9, 5] and r['g_shape'].tolist() == [5, 2]
    assert bool(r['same']) and not bool(r['differs'])
    assert r['any_true'].tolist() == [True, False]                  # one rank out of input -> every rank stops
