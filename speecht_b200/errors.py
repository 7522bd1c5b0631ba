"""Exceptions the reference's callers catch (tf.errors.OutOfRangeError at training.py:92, evaluation.py:109)."""


class OutOfRangeError(Exception):
  """Raised by model.step when the input pipeline is exhausted (the closed-FIFOQueue condition in TF1)."""
