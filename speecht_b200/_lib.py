"""ctypes binding of libspeecht_b200.so (the C ABI declared in include/speecht_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this module raises.  The product path never
touches oracle/.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPEECHT_B200_LIB selects another build of the same C ABI (A/B timing of kernel variants); default: in-tree build
LIB_PATH = os.environ.get('SPEECHT_B200_LIB') or os.path.join(_HERE, 'libspeecht_b200.so')

ST_OK = 0
ST_ERR_INVALID_ARG = -1
ST_ERR_CTC_LABELS = -2
ST_ERR_CUDA = -3
ST_ERR_UNSUPPORTED = -4


class NativeLibraryMissing(ImportError):
  pass


class NativeError(RuntimeError):
  def __init__(self, code, message):
    super().__init__('speecht_b200 native call failed (%d): %s' % (code, message))
    self.code = code


class CTCLabelError(ValueError):
  """tf.nn.ctc_loss InvalidArgumentError equivalent ("Not enough time for target transition sequence")."""


P = c_void_p
_SIGNATURES = {
  'st_version': (c_int, []),
  'st_last_error': (c_char_p, []),
  'st_device_sync': (c_int, []),
  'st_ctc_greedy_decode': (c_int, [P, c_int64, c_int64, c_int, c_int, c_int, P, c_int, c_int, P, P, P, P]),
  'st_ctc_validate_labels_host': (c_int, [P, P, P, c_int, c_int, c_int]),
  'st_ctc_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
  'st_ctc_loss': (c_int, [P, c_int64, c_int64, c_int, c_int, c_int, P, P, c_int, P, c_int, P, P, c_float,
                          P, c_int, c_int, P, P, c_size_t, P]),
  'st_conv1d_fwd_f32': (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
  'st_conv1d_bwd_data_f32': (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
  'st_conv1d_bwd_filter_f32': (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
  'st_sumsq': (c_int, [P, c_int64, P, c_int, P]),
  'st_clip_adam': (c_int, [P, P, P, P, c_int64, c_float, c_float, c_float, c_float, c_int64, c_float, P, c_float, P]),
  'st_melspec_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
  'st_melspec': (c_int, [P, c_int64, P, c_int, c_int, P, c_int, c_int, c_int, P, c_int, P, P, c_size_t, P]),
  'st_plan_create': (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int]),
  'st_plan_destroy': (c_int, [P]),
  'st_plan_arena_bytes': (c_size_t, [P]),
  'st_plan_param_floats': (c_int64, [P]),
  'st_plan_logit_frames': (c_int, [P]),
  'st_plan_bind': (c_int, [P, P, c_size_t, P, P]),
  'st_plan_pack_weights': (c_int, [P, P]),
  'st_plan_forward': (c_int, [P, P, P]),
  'st_plan_backward': (c_int, [P, P]),
  'st_plan_reserve_sms': (c_int, [P, c_int]),
  'st_plan_prepare_backward': (c_int, [P, P]),
  'st_plan_backward_range': (c_int, [P, c_int, c_int, P]),
  'st_plan_logits': (c_void_p, [P]),
  'st_plan_dlogits_planes': (c_void_p, [P]),
  'st_plan_get_activation': (c_int, [P, c_int, P, P]),
  'st_plan_launches': (c_int, [P]),
  'st_plan_filter_set': (c_int, [P]),
  'st_plan_set_timing': (c_int, [P, c_int]),
  'st_plan_read_timings': (c_int, [P, P, P, P, P, c_int]),
  'st_debug_conv_timeline': (c_int, [P, c_int]),
  'st_flac_info_host': (c_int, [P, c_size_t, P, P, P]),
  'st_flac_decode_host': (c_int, [P, c_size_t, P, c_int64, P]),
}

_lib = None


def declared_symbols():
  """Every function name include/speecht_b200.h declares (parsed from the header, used by the CPU tests)."""
  import re
  header = os.path.join(os.path.dirname(_HERE), 'include', 'speecht_b200.h')
  text = open(header).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(st_[a-z0-9_]+)\s*\(', text)))


def register(name, restype, argtypes):
  _SIGNATURES[name] = (restype, argtypes)
  if _lib is not None:
    fn = getattr(_lib, name)
    fn.restype, fn.argtypes = restype, argtypes


def lib():
  """Load (once) and return the ctypes handle.  Raises NativeLibraryMissing when the .so has not been built."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise NativeLibraryMissing(
        '%s not found: build it with `python -m speecht_b200.build` (there is no CPU fallback)' % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
      fn = getattr(handle, name)
      fn.restype, fn.argtypes = restype, argtypes
    _lib = handle
  return _lib


def last_error():
  return lib().st_last_error().decode('utf-8', 'replace')


def check(code):
  if code == ST_OK:
    return
  msg = last_error()
  if code == ST_ERR_CTC_LABELS:
    raise CTCLabelError(msg)
  if code == ST_ERR_INVALID_ARG:
    raise ValueError(msg)
  raise NativeError(code, msg)


def ptr(t):
  """Device/host pointer of a torch tensor (or None -> NULL)."""
  return None if t is None else c_void_p(t.data_ptr())


def stream_ptr(stream=None):
  import torch
  s = stream if stream is not None else torch.cuda.current_stream()
  return c_void_p(s.cuda_stream)
