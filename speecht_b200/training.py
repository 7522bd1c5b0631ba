"""`train`: the optimisation loop around model.step.

Policy follows reference speecht/training.py:44-99 -- time every step, and every `steps_per_checkpoint` steps: report
(global step, learning rate, mean step time, loss, perplexity), multiply the learning rate by
`learning_rate_decay_factor` when the windowed mean loss is worse than each of the last three windows, append the
window loss to the history and write `speechT.ckpt-<global_step>`.

New (the reference is single-process): under torchrun (WORLD_SIZE > 1) every rank runs this loop on its own GPU and
reads every WORLD_SIZE-th sample of one commonly seeded shuffle stream; the gradient allreduce happens inside
model.step (speecht_b200/parallel.py); the learning-rate decision uses the loss averaged over ranks so that all ranks
take it together; rank 0 alone prints and writes checkpoints.
"""
import itertools
import math
import os
import random
import time

from . import parallel
from .errors import OutOfRangeError
from .execution import DatasetExecutor
from .speech_model import Session, create_default_model


class _Window:
  """Running means over one checkpoint window."""

  def __init__(self, length):
    self.length = length
    self.reset()

  def reset(self):
    self.step_time = 0.0
    self.loss = 0.0

  def add(self, seconds, loss):
    self.step_time += seconds / self.length
    self.loss += loss / self.length


class Training(DatasetExecutor):

  def __init__(self, flags):
    self.rank, self.local_rank, self.world = parallel.init_from_env()
    if self.world > 1:
      import torch
      import torch.distributed as dist
      torch.cuda.set_device(self.local_rank % torch.cuda.device_count())
      flags.process_group = dist.group.WORLD
      # the per-step "has any rank run out of input?" word travels over a HOST (gloo) group: on the NCCL group it is
      # ordered behind the previous step's gradient allreduce and Adam, and reading it would stall the host until that
      # step has completely finished -- the GPU would then idle while the next step is enqueued
      self.flag_group = None
      if dist.get_backend() == 'nccl':
        try:
          self.flag_group = dist.new_group(backend='gloo')
        except Exception as e:            # no gloo in this build: fall back to the (stalling) word on the NCCL group
          print('speecht_b200: no host group for the termination word (%s); using the NCCL group' % e)
    super().__init__(flags)

  # ---- DatasetExecutor hooks ---------------------------------------------------------------------
  def create_sample_generator(self, limit_count: int):
    if self.world == 1:
      return self.reader.load_samples('train', loop_infinitely=True, limit_count=limit_count,
                                      feature_type=self.flags.feature_type)
    # data parallel: every rank shuffles the FILE LIST with the same private generator and reads only its own
    # every-world-th file -- no rank loads utterances it will discard, and the order does not depend on the global
    # `random` state staying in step across ranks
    seed = int(os.environ.get('SPEECHT_B200_DATA_SEED', '1234'))
    return self.reader.load_samples('train', loop_infinitely=True, limit_count=limit_count,
                                    feature_type=self.flags.feature_type, shard=(self.rank, self.world),
                                    rng=random.Random(seed))

  def get_loader_limit_count(self) -> int:
    return self.flags.limit_training_set

  def create_model(self, sess):
    """Resume from the run directory if it holds a checkpoint, otherwise start from Xavier weights."""
    model = create_default_model(self.flags, self.input_size, self.speech_input)
    reset_to = self.flags.learning_rate if self.flags.reset_learning_rate else None
    model.restore_or_create(sess, self.flags.run_train_dir, reset_to)
    return model

  # ---- checkpoint-time policy --------------------------------------------------------------------
  def _window_loss(self, model, window):
    if self.world == 1:
      return window.loss
    import torch
    local = torch.tensor([window.loss], device=model.engine.device)
    return float(parallel.mean_scalar(local).item())

  def _report(self, model, window, last_loss):
    perplexity = math.exp(float(last_loss)) if last_loss < 300 else float('inf')
    print('global step {:d} learning rate {:.4f} step-time {:.2f} average loss {:.2f} perplexity {:.2f}'
          .format(model.global_step.eval(), model.learning_rate.eval(), window.step_time, last_loss, perplexity))

  def _maybe_decay(self, sess, model, window_loss, history):
    factor = self.flags.learning_rate_decay_factor
    if factor > 0 and len(history) > 2 and window_loss > max(history[-3:]):
      sess.run(model.learning_rate_decay_op)

  def _save(self, sess, model):
    path = os.path.join(self.flags.run_train_dir, 'speechT.ckpt')
    model.saver.save(sess, path, global_step=model.global_step)
    print('Model saved')

  def _any_rank_exhausted(self, model):
    if getattr(self, 'flag_group', None) is not None:
      return parallel.any_rank_true(model.input_exhausted(), device='cpu', group=self.flag_group)
    return parallel.any_rank_true(model.input_exhausted(), device=model.engine.device)

  # ---- the loop ----------------------------------------------------------------------------------
  def run(self, max_steps=None):
    per_checkpoint = self.flags.steps_per_checkpoint
    window = _Window(per_checkpoint)
    history = []
    with Session() as sess:
      model = self.create_model(sess)
      # two feeder threads like the reference (training.py:49); one per rank under data parallelism, where the
      # threads of a rank must not race for the commonly seeded shuffle stream
      coord = self.start_pipeline(sess, n_threads=1 if self.world > 1 else 2)
      print('Begin training')
      try:
        for step in itertools.count(1):
          if coord.should_stop() or (max_steps is not None and step > max_steps):
            break
          if self.world > 1 and self._any_rank_exhausted(model):
            print('Done training -- a rank ran out of input')      # all ranks leave together: nobody waits in NCCL
            break
          at_checkpoint = step % per_checkpoint == 0
          began = time.time()
          fetched = model.step(sess, summary=at_checkpoint)
          window.add(time.time() - began, fetched[0])
          if not at_checkpoint:
            continue
          window_loss = self._window_loss(model, window)
          if self.rank == 0:
            self._report(model, window, fetched[0])
            model.summary_writer.add_summary(fetched[2], model.global_step.eval())
          self._maybe_decay(sess, model, window_loss, history)
          history.append(window_loss)
          if self.rank == 0:
            self._save(sess, model)
          window.reset()
      except OutOfRangeError:
        print('Done training -- step limit reached')
      finally:
        coord.request_stop()
      coord.join()
      self.speech_input.raise_if_failed()       # a feeder-thread exception must not look like a clean end of data
    return model
