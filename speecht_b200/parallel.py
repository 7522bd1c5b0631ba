"""Data parallelism over the utterance batch -- new work, the reference is single-process.

One process per GPU.  Every op on the path is independent per utterance except tf.reduce_mean (speech_model.py:75),
the global-norm clip (:80) and the weight update (:81), so the ONLY collective is a sum-allreduce of the flat
gradient buffer (24.66 M floats, 98.7 MB) over NCCL/NVLink; the 1/(B_local*world) factor is folded into the CTC
gradient, the global norm is computed on the reduced gradient (identical on all ranks -> no second collective) and
every rank applies the identical Adam update.  These helpers are device-agnostic so the host logic is tested with
gloo on CPU (tests/test_parallel_gloo.py).
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend=None):
  """torchrun contract: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT.  Returns (rank, local, world)."""
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  if world > 1 and not dist.is_initialized():
    if backend is None:
      # SPEECHT_B200_DIST_BACKEND=gloo: host-staged collectives, which also allow several ranks to share one GPU (a
      # smoke run of the data-parallel loop on a single-GPU box; NCCL refuses two ranks on one device)
      backend = os.environ.get('SPEECHT_B200_DIST_BACKEND') or ('nccl' if torch.cuda.is_available() else 'gloo')
    if backend == 'nccl':
      torch.cuda.set_device(local)
      dist.init_process_group(backend, device_id=torch.device('cuda', local))
    else:
      if torch.cuda.is_available():
        torch.cuda.set_device(local % torch.cuda.device_count())
      dist.init_process_group(backend)
  return rank, local, world


def shard_batch(n_items, rank, world):
  """Contiguous shard [start, stop) of a global batch of n_items utterances; sizes differ by at most one."""
  base, extra = divmod(n_items, world)
  start = rank * base + min(rank, extra)
  return start, start + base + (1 if rank < extra else 0)


def gradient_scale(local_batch, world):
  """Factor folded into d(loss)/d(logits): mean over the GLOBAL batch when every rank holds local_batch items."""
  return 1.0 / (local_batch * world)


def allreduce_flat(flat, group=None, buckets=None):
  """Sum-allreduce of the flat gradient buffer.  `buckets` = list of (start, stop) float ranges launched as
  separate collectives (async) so that early-finished layers overlap with the rest of backward; None = one call."""
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return []
  if not buckets:
    return [dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)]
  return [dist.all_reduce(flat[a:b], op=dist.ReduceOp.SUM, group=group, async_op=True) for a, b in buckets]


def mean_scalar(value, group=None):
  """Average of a per-rank scalar tensor (the reported avg_loss of the global batch)."""
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return value
  out = value.clone()
  dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
  return out / dist.get_world_size(group)


def max_scalar(x, device=None, group=None):
  """Max over ranks of a python float (bench timing: the slowest rank defines the step time)."""
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return float(x)
  t = torch.tensor([float(x)], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
  return float(t.item())


def gather_decoded(decoded_rows, group=None):
  """Decode (config 5) is embarrassingly parallel: only the compacted label lists travel.  decoded_rows is this
  rank's list of int lists; returns the concatenation in rank order on every rank."""
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return list(decoded_rows)
  gathered = [None] * dist.get_world_size(group)
  dist.all_gather_object(gathered, [list(map(int, r)) for r in decoded_rows], group=group)
  return [row for part in gathered for row in part]


def gather_decoded_sparse(sparse, group=None, device=None):
  """Decode gather over tensors instead of pickled objects: every rank contributes the (indices, values, dense_shape)
  triple of its greedy decode; returns the triple of the GLOBAL batch (rows offset by the batch sizes of the lower
  ranks) on every rank.  Two collectives: the sizes, then the padded entries."""
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return sparse
  world = dist.get_world_size(group)
  n, b = int(sparse.values.shape[0]), int(sparse.dense_shape[0])
  meta = torch.tensor([n, b, int(sparse.dense_shape[1])], dtype=torch.int64, device=device)
  metas = [torch.empty_like(meta) for _ in range(world)]
  dist.all_gather(metas, meta, group=group)
  metas = torch.stack(metas).cpu().numpy()
  cap = int(metas[:, 0].max())
  mine = torch.zeros((max(cap, 1), 3), dtype=torch.int64, device=device)
  if n:
    mine[:n, :2] = torch.from_numpy(np.ascontiguousarray(sparse.indices)).to(device)
    mine[:n, 2] = torch.from_numpy(np.ascontiguousarray(sparse.values)).to(device)
  parts = [torch.empty_like(mine) for _ in range(world)]
  dist.all_gather(parts, mine, group=group)
  indices, values, row0 = [], [], 0
  for r in range(world):
    part = parts[r][:int(metas[r, 0])].cpu().numpy()
    idx = part[:, :2].copy()
    idx[:, 0] += row0
    indices.append(idx)
    values.append(part[:, 2])
    row0 += int(metas[r, 1])
  from .ops import SparseTensorValue
  return SparseTensorValue(np.concatenate(indices).reshape(-1, 2), np.concatenate(values),
                           np.array([row0, int(metas[:, 2].max())], dtype=np.int64))


def identical_across_ranks(tensor, group=None):
  """True when every rank holds bit-identical contents (data-parallel invariant: identical parameters after every
  update).  Compares the elementwise MAX and MIN over ranks of the raw bit patterns."""
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return True
  bits = tensor.detach().contiguous().view(torch.int32).clone()
  hi, lo = bits.clone(), bits
  dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
  dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
  return bool(torch.equal(hi, lo))


def any_rank_true(flag, device=None, group=None):
  """Logical OR of a per-rank python bool (data-parallel ranks agree on termination before entering a step, so that
  no rank is left waiting in the gradient allreduce)."""
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return bool(flag)
  t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
  return bool(t.item())
