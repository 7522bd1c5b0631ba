"""Feature extraction + the preprocessed-sample store (mirror of reference speecht/preprocessing.py).

`calc_power_spectrogram(audio_data, samplerate, n_mels=128, n_fft=512, hop_length=160)` keeps the reference
signature (preprocessing.py:36) and runs on the GPU (csrc/melspec.cu).  `SpeechCorpusReader` keeps the on-disk
contract between `preprocess` and `train/evaluate`: `<data>/preprocessed-power/<split>/<audio_id>.npz` with keys
`audio_fragments [T, n_mels]` and `transcript [L]` (preprocessing.py:178,199-206,243-279).

Audio loading follows `librosa.load(audio_file)` (preprocessing.py:169): mono float32 resampled to 22050 Hz.  The
image has no soundfile / audioread / ffmpeg, so .flac files go through the library's own FLAC decoder
(csrc/flac_host.cu, MD5-verified) and .wav files through the stdlib; resampling is a polyphase Kaiser FIR (scipy).
The MFCC feature type (preprocessing.py:61-84) is out of scope (not the default, not on the hot path).
"""
import fnmatch
import itertools
import logging
import os
import random
import wave

import numpy as np

from . import vocabulary


def normalize(values):
  """(x - mean) / std over the whole array (preprocessing.py:29-33)."""
  return (values - np.mean(values)) / np.std(values)


def calc_power_spectrogram(audio_data, samplerate, n_mels=128, n_fft=512, hop_length=160):
  """preprocessing.py:36-58 on the GPU.  audio_data: 1-D float array -> ndarray [time, n_mels] float32."""
  import torch
  from . import ops
  wav = torch.from_numpy(np.ascontiguousarray(audio_data, dtype=np.float32)).cuda()[None]
  feat, frames = ops.power_spectrogram(wav, [wav.shape[1]], samplerate, n_mels=n_mels, n_fft=n_fft,
                                       hop_length=hop_length)
  return feat[0, :int(frames[0].item())].cpu().numpy()


def calc_power_spectrogram_batch(audio_list, samplerate, n_mels=128, n_fft=512, hop_length=160):
  """Batched variant: list of 1-D arrays -> list of [time_i, n_mels] arrays, one kernel sequence for all."""
  import torch
  from . import ops
  lens = [len(a) for a in audio_list]
  host = np.zeros((len(audio_list), max(lens)), dtype=np.float32)
  for i, a in enumerate(audio_list):
    host[i, :lens[i]] = a
  feat, frames = ops.power_spectrogram(torch.from_numpy(host).cuda(), lens, samplerate, n_mels=n_mels, n_fft=n_fft,
                                       hop_length=hop_length)
  feat, frames = feat.cpu().numpy(), frames.cpu().numpy()
  return [feat[i, :frames[i]].copy() for i in range(len(audio_list))]


LOAD_SAMPLERATE = 22050          # librosa.load's default `sr`: the reference resamples EVERY file to it


def resample(audio, orig_sr, target_sr=LOAD_SAMPLERATE):
  """Band-limited polyphase resampling (Kaiser-windowed FIR, scipy.signal.resample_poly) in place of librosa's
  'kaiser_best' windowed-sinc interpolation (resampy): same output length ceil(n * target / orig), the same kind
  of filter; sample values agree to the filters' stop-band level, not bit for bit."""
  if int(orig_sr) == int(target_sr):
    return np.ascontiguousarray(audio, dtype=np.float32)
  from math import gcd
  from scipy.signal import resample_poly
  g = gcd(int(orig_sr), int(target_sr))
  out = resample_poly(np.asarray(audio, dtype=np.float64), int(target_sr) // g, int(orig_sr) // g)
  n_out = -(-len(audio) * int(target_sr) // int(orig_sr))          # ceil, librosa's n_samples
  return np.ascontiguousarray(out[:n_out], dtype=np.float32)


def read_wav(path):
  """Minimal PCM16 WAV reader -> (mono float32 in [-1, 1), native sample rate)."""
  with wave.open(path, 'rb') as f:
    if f.getsampwidth() != 2:
      raise ValueError('only 16-bit PCM WAV is supported')
    data = np.frombuffer(f.readframes(f.getnframes()), dtype=np.int16).astype(np.float32) / 32768.0
    if f.getnchannels() > 1:
      data = data.reshape(-1, f.getnchannels()).mean(axis=1)
    return data, f.getframerate()


def read_flac(path, verify_md5=True):
  """FLAC file -> (mono float32 in [-1, 1), native sample rate) through the library's native decoder
  (csrc/flac_host.cu); the decoded PCM is checked against the MD5 signature in STREAMINFO."""
  import ctypes
  import hashlib
  from ._lib import check, lib
  raw = np.fromfile(path, dtype=np.uint8)
  info = (ctypes.c_int32 * 3)()
  total = ctypes.c_int64()
  md5 = (ctypes.c_uint8 * 16)()
  check(lib().st_flac_info_host(raw.ctypes.data, raw.size, info, ctypes.byref(total), md5))
  rate, channels, bits = info[0], info[1], info[2]
  # unknown length (streamed encodes): a frame holds at most 65535 samples and at least ~10 bytes
  capacity = total.value if total.value > 0 else raw.size * 65535 // 10
  pcm = np.empty((max(capacity, 1), channels), dtype=np.int32)
  decoded = ctypes.c_int64()
  check(lib().st_flac_decode_host(raw.ctypes.data, raw.size, pcm.ctypes.data, capacity, ctypes.byref(decoded)))
  pcm = pcm[:decoded.value]
  if verify_md5 and any(md5):
    width = (bits + 7) // 8
    le = pcm.astype('<i4').view(np.uint8).reshape(-1, channels, 4)[:, :, :width]
    if hashlib.md5(np.ascontiguousarray(le).tobytes()).digest() != bytes(md5):
      raise ValueError('FLAC MD5 mismatch: %s does not decode to the audio it was encoded from' % path)
  audio = pcm.astype(np.float32).mean(axis=1) / float(1 << (bits - 1))
  return audio, rate


def load_audio(path, sr=LOAD_SAMPLERATE):
  """librosa.load(path) as the reference calls it (preprocessing.py:169): mono float32 RESAMPLED to 22050 Hz.
  -> (audio, sr).  .flac through the native decoder, .wav through the stdlib."""
  ext = os.path.splitext(path)[1].lower()
  if ext == '.flac':
    audio, rate = read_flac(path)
  elif ext == '.wav':
    audio, rate = read_wav(path)
  else:
    raise ValueError('unsupported audio file type: %s' % path)
  return resample(audio, rate, sr), sr


def load_wav(path):
  """Kept for callers of the first version: a WAV file loaded like load_audio (resampled to 22050 Hz)."""
  return load_audio(path)


def iglob_recursive(directory, file_pattern):
  for root, _dirs, file_names in os.walk(directory):
    for filename in fnmatch.filter(file_names, file_pattern):
      yield os.path.join(root, filename)


class SpeechCorpusReader:
  """Reads / writes the preprocessed corpus (preprocessing.py:103-279)."""

  AUDIO_PATTERNS = ('*.flac', '*.wav')

  def __init__(self, data_directory, load_audio=None):
    self._data_directory = data_directory
    self._transcript_dict_cache = None
    self._load_audio = load_audio or globals()['load_audio']

  @property
  def _transcript_dict(self):
    if not self._transcript_dict_cache:
      self._transcript_dict_cache = self._build_transcript()
    return self._transcript_dict_cache

  @staticmethod
  def _get_transcript_entries(transcript_directory):
    """Yields [audio_id, sentence] from every *.trans.txt (lines `ID WORD1 WORD2 ...`)."""
    for transcript_file in iglob_recursive(transcript_directory, '*.trans.txt'):
      with open(transcript_file, 'r') as f:
        for line in f:
          yield line.rstrip('\n').split(' ', 1)

  def _build_transcript(self):
    return {entry[0]: vocabulary.sentence_to_ids(entry[1])
            for entry in self._get_transcript_entries(self._data_directory)}

  @classmethod
  def _extract_audio_id(cls, audio_file):
    return os.path.splitext(os.path.basename(audio_file))[0]

  def _audio_files(self, directory):
    files = []
    for pattern in self.AUDIO_PATTERNS:
      files.extend(iglob_recursive(self._data_directory + '/' + directory, pattern))
    return files

  def _get_directory(self, feature_type, sub_directory):
    preprocess_directory = 'preprocessed'
    if feature_type == calc_power_spectrogram or feature_type == 'power':
      preprocess_directory += '-power'
    return self._data_directory + '/' + preprocess_directory + '/' + sub_directory

  def generate_samples(self, directory, preprocess_fnc):
    """(audio_id, audio_fragments, transcript) for every audio file below `directory`."""
    transcript_dict = self._transcript_dict
    for audio_file in self._audio_files(directory):
      audio_data, samplerate = self._load_audio(audio_file)
      audio_id = self._extract_audio_id(audio_file)
      yield audio_id, preprocess_fnc(audio_data, samplerate), transcript_dict[audio_id]

  def store_samples(self, directory, preprocess_fnc, batch_size=32):
    """Preprocess every audio file of `directory` into <preprocessed[-power]>/<directory>/<audio_id>.npz.
    The reference fans out over a multiprocessing Pool (preprocessing.py:229); here utterances are batched onto
    the GPU instead when the feature function is the power spectrogram."""
    out_directory = self._get_directory(preprocess_fnc, directory)
    os.makedirs(out_directory, exist_ok=True)
    transcript_dict = self._transcript_dict
    files = self._audio_files(directory)
    for i in range(0, len(files), batch_size):
      chunk = files[i:i + batch_size]
      loaded = [self._load_audio(f) for f in chunk]
      rates = {sr for _a, sr in loaded}
      if preprocess_fnc == calc_power_spectrogram and len(rates) == 1:
        feats = calc_power_spectrogram_batch([a for a, _sr in loaded], rates.pop())
      else:
        feats = [preprocess_fnc(a, sr) for a, sr in loaded]
      for audio_file, fragments in zip(chunk, feats):
        audio_id = self._extract_audio_id(audio_file)
        np.savez(out_directory + '/' + audio_id, audio_fragments=fragments, transcript=transcript_dict[audio_id])

  def load_samples(self, directory, max_size=False, loop_infinitely=False, limit_count=0, feature_type='mfcc',
                   shard=None, rng=None):
    """Generator of (audio_fragments [T, n_features], transcript [L]) over the stored .npz files of `directory`
    (preprocessing.py:243-279): shuffled once, optionally truncated to `limit_count` files, over-long utterances
    (more than `max_size` frames) skipped with a warning; with `loop_infinitely` reshuffled after every pass.
    `rng` (a random.Random) replaces the global generator; `shard` = (rank, world) keeps every world-th file."""
    load_directory = self._get_directory(feature_type, directory)
    if not os.path.exists(load_directory):
      raise ValueError('Directory {} does not exist'.format(load_directory))
    files = list(iglob_recursive(load_directory, '*.npz'))
    if rng is not None:
      files.sort()                              # os.walk order is not portable; the private generator defines the order
    shuffle = rng.shuffle if rng is not None else random.shuffle
    shuffle(files)
    files = files[:limit_count] if limit_count else files
    rank, world = shard if shard is not None else (0, 1)
    if len(files) < world:
      # fewer files than ranks (e.g. the single sample that determines the input size): a rank's every-world-th slice
      # would be empty -- and with loop_infinitely the generator would spin forever without yielding; every rank reads
      # the whole (tiny) list instead
      rank, world = 0, 1
    if not files:
      return
    for epoch in itertools.count():
      if epoch and not loop_infinitely:
        return
      if epoch:
        shuffle(files)
      # shard = (rank, world): this reader yields every world-th file of the commonly shuffled list (new; the
      # reference is single-process)
      for path in files[rank::world]:
        with np.load(path) as stored:
          fragments, transcript = stored['audio_fragments'], stored['transcript']
        if max_size and fragments.shape[0] > max_size:
          logging.warning('Audio snippet too long: {}'.format(fragments.shape[0]))
          continue
        yield fragments, transcript


class Preprocessing:
  """`speecht-cli preprocess` (preprocessing.py:282-311) without the corpus download (no network here)."""

  def __init__(self, flags):
    self.flags = flags

  def run(self):
    corpus_reader = SpeechCorpusReader(self.flags.data_dir)
    if self.flags.feature_type == 'power':
      preprocess_fnc = calc_power_spectrogram
    elif self.flags.feature_type == 'mfcc':
      raise ValueError('mfcc features are out of scope of speecht_b200; use --power')
    else:
      raise ValueError('Feature type must be mfcc or power.')
    preprocess_all = not (self.flags.train_only or self.flags.test_only or self.flags.dev_only)
    for enabled, split, title in ((self.flags.train_only, 'train', 'training'), (self.flags.test_only, 'test', 'test'),
                                  (self.flags.dev_only, 'dev', 'development')):
      if enabled or preprocess_all:
        print('Preprocessing {} data'.format(title))
        corpus_reader.store_samples(split, preprocess_fnc)
