"""Training loop -- mirror of reference speecht/training.py:26-99 (step timing, LR decay on no improvement over the
last three checkpoints, checkpoint cadence).

New (the reference is single-process): under torchrun (WORLD_SIZE > 1) every rank runs this loop on its own GPU,
reads every WORLD_SIZE-th sample of one commonly shuffled stream, and the gradient is allreduced inside model.step
(speecht_b200/parallel.py); rank 0 alone prints and writes checkpoints."""
import itertools
import os
import random
import time

import numpy as np

from . import parallel
from .errors import OutOfRangeError
from .execution import DatasetExecutor
from .speech_model import Session, create_default_model


class Training(DatasetExecutor):

  def __init__(self, flags):
    self.rank, self.local_rank, self.world = parallel.init_from_env()
    if self.world > 1:
      import torch
      import torch.distributed as dist
      torch.cuda.set_device(self.local_rank)
      flags.process_group = dist.group.WORLD
      random.seed(int(os.environ.get('SPEECHT_B200_DATA_SEED', '1234')))   # same shuffle order on every rank
    super().__init__(flags)

  def create_sample_generator(self, limit_count: int):
    gen = self.reader.load_samples('train', loop_infinitely=True, limit_count=limit_count,
                                   feature_type=self.flags.feature_type)
    if self.world > 1:
      gen = itertools.islice(gen, self.rank, None, self.world)
    return gen

  def get_loader_limit_count(self) -> int:
    return self.flags.limit_training_set

  def create_model(self, sess):
    model = create_default_model(self.flags, self.input_size, self.speech_input)
    model.restore_or_create(sess, self.flags.run_train_dir,
                            self.flags.learning_rate if self.flags.reset_learning_rate else None)
    return model

  def run(self, max_steps=None):
    with Session() as sess:
      model = self.create_model(sess)
      # two feeder threads like the reference (training.py:49) -- one under data parallelism, where the threads of
      # a rank must not race for the commonly seeded shuffle stream
      coord = self.start_pipeline(sess, n_threads=2 if self.world == 1 else 1)
      step_time, loss = 0.0, 0.0
      current_step = 0
      previous_losses = []
      try:
        print('Begin training')
        while not coord.should_stop():
          current_step += 1
          is_checkpoint_step = current_step % self.flags.steps_per_checkpoint == 0
          start_time = time.time()
          step_result = model.step(sess, summary=is_checkpoint_step)
          avg_loss = step_result[0]
          step_time += (time.time() - start_time) / self.flags.steps_per_checkpoint
          loss += avg_loss / self.flags.steps_per_checkpoint
          if is_checkpoint_step:
            global_step = model.global_step.eval()
            if self.world > 1:
              # every rank must take the same learning-rate decision: use the global mean of the running loss
              import torch
              loss = float(parallel.mean_scalar(torch.tensor([loss], device=model.engine.device)).item())
            if self.rank == 0:
              perplexity = np.exp(float(avg_loss)) if avg_loss < 300 else float('inf')
              print('global step {:d} learning rate {:.4f} step-time {:.2f} average loss {:.2f} perplexity {:.2f}'
                    .format(global_step, model.learning_rate.eval(), step_time, avg_loss, perplexity))
              model.summary_writer.add_summary(step_result[2], global_step)
            # decrease the learning rate if no improvement was seen over the last 3 checkpoints
            if self.flags.learning_rate_decay_factor > 0 and len(previous_losses) > 2 \
                and loss > max(previous_losses[-3:]):
              sess.run(model.learning_rate_decay_op)
            previous_losses.append(loss)
            if self.rank == 0:
              checkpoint_path = os.path.join(self.flags.run_train_dir, 'speechT.ckpt')
              model.saver.save(sess, checkpoint_path, global_step=model.global_step)
              print('Model saved')
            step_time, loss = 0.0, 0.0
          if max_steps is not None and current_step >= max_steps:
            break
      except OutOfRangeError:
        print('Done training -- step limit reached')
      finally:
        coord.request_stop()
      coord.join()
    return model
