"""DatasetExecutor: shared wiring of corpus reader, batch loader and model for the dataset-driven commands.
Same protocol as the reference's execution.DatasetExecutor (execution.py:26-63): subclasses provide
`create_sample_generator(limit_count)` and `get_loader_limit_count()`, optionally `get_max_steps()`;
`start_pipeline(sess, n_threads)` returns a coordinator and `create_model(sess)` a restored model."""
import abc
import functools

from . import speech_input
from .preprocessing import SpeechCorpusReader
from .speech_model import create_default_model


class DatasetExecutor(abc.ABC):

  def __init__(self, flags):
    self.flags = flags
    self.reader = SpeechCorpusReader(flags.data_dir)
    print('Determine input size from first sample')
    self.input_size = self.determine_input_size()
    print('Initialize InputBatchLoader')
    generator_factory = functools.partial(self.create_sample_generator, self.get_loader_limit_count())
    self.speech_input = speech_input.InputBatchLoader(self.input_size, flags.batch_size, generator_factory,
                                                      max_steps=self.get_max_steps())

  # ---- hooks -------------------------------------------------------------------------------------
  @abc.abstractmethod
  def create_sample_generator(self, limit_count: int):
    """Iterator of (features [T, input_size], transcript ids)."""

  @abc.abstractmethod
  def get_loader_limit_count(self) -> int:
    """How many samples the loader may use (0 = all)."""

  def get_max_steps(self):
    """Number of batches after which the loader closes its queue (None = unbounded)."""
    return None

  # ---- shared behaviour --------------------------------------------------------------------------
  def determine_input_size(self):
    first_features, _first_transcript = next(iter(self.create_sample_generator(limit_count=1)))
    return first_features.shape[1]

  def start_pipeline(self, sess, n_threads=1):
    coordinator = speech_input.Coordinator()
    self.speech_input.start_threads(sess=sess, coord=coordinator, n_threads=n_threads)
    return coordinator

  def create_model(self, sess):
    """Evaluation-style creation: a checkpoint (or exported weights) MUST exist (raises FileNotFoundError)."""
    model = create_default_model(self.flags, self.input_size, self.speech_input)
    model.restore(sess, self.flags.run_train_dir)
    return model
