"""CPU tests of the audio-loading caller of the feature path (reference preprocessing.py:169, librosa.load): the native
FLAC decoder (host function of libspeecht_b200.so) against streams produced by a small encoder written here, against
the MD5 signature the format carries, and -- when the reference checkout is present -- against the reference's own
known answer for its LibriSpeech fixture (test_speechCorpusReader.py:40-45: 114881 samples after resampling)."""
import hashlib
import os
import struct
import wave

import numpy as np
import pytest

REF_FLAC = '/root/reference/speecht/tests/data/train/1089-134686-0037.flac'


# ------------------------------------------------------------------------------------------------ a tiny FLAC encoder
class BitWriter:
  def __init__(self):
    self.bits = []

  def put(self, value, n):
    for i in reversed(range(n)):
      self.bits.append((value >> i) & 1)

  def put_signed(self, value, n):
    self.put(value & ((1 << n) - 1), n)

  def unary(self, q):
    self.bits.extend([0] * q + [1])

  def align(self):
    while len(self.bits) % 8:
      self.bits.append(0)

  def tobytes(self):
    assert len(self.bits) % 8 == 0
    return bytes(int(''.join(map(str, self.bits[i:i + 8])), 2) for i in range(0, len(self.bits), 8))


def crc(data, poly, width):
  c, top = 0, 1 << (width - 1)
  for byte in data:
    c ^= byte << (width - 8)
    for _ in range(8):
      c = ((c << 1) ^ poly) if c & top else (c << 1)
      c &= (1 << width) - 1
  return c


def rice(w, residuals, k):
  for r in residuals:
    u = (r << 1) ^ (r >> 63) if r >= 0 else ((-r) << 1) - 1
    w.unary(u >> k)
    w.put(u & ((1 << k) - 1), k)


def subframe(w, samples, bps, kind):
  n = len(samples)
  if kind == 'constant':
    w.put(0, 1); w.put(0, 6); w.put(0, 1)
    w.put_signed(int(samples[0]), bps)
  elif kind == 'verbatim':
    w.put(0, 1); w.put(1, 6); w.put(0, 1)
    for s in samples:
      w.put_signed(int(s), bps)
  elif kind == 'wasted':                                # verbatim with 3 wasted (zero) low bits
    w.put(0, 1); w.put(1, 6); w.put(1, 1); w.unary(2)
    for s in samples:
      w.put_signed(int(s) >> 3, bps - 3)
  elif kind.startswith('fixed'):
    order = int(kind[-1])
    w.put(0, 1); w.put(8 + order, 6); w.put(0, 1)
    for s in samples[:order]:
      w.put_signed(int(s), bps)
    x = [int(v) for v in samples]
    coef = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}[order]
    res = [x[i] - sum(c * x[i - 1 - j] for j, c in enumerate(coef)) for i in range(order, n)]
    porder = 1 if n % 2 == 0 and n // 2 > order else 0   # two Rice partitions when the block splits evenly
    w.put(0, 2); w.put(porder, 4)
    if porder:
      cut = n // 2 - order
      w.put(3, 4); rice(w, res[:cut], 3)
      w.put(15, 4); w.put(bps + 4, 5)                     # escape partition: raw residuals
      for r in res[cut:]:
        w.put_signed(r, bps + 4)
    else:
      w.put(4, 4); rice(w, res, 4)
  elif kind == 'lpc':
    order, precision, shift = 2, 12, 10
    coef = [1900, -900]
    w.put(0, 1); w.put(31 + order, 6); w.put(0, 1)
    for s in samples[:order]:
      w.put_signed(int(s), bps)
    w.put(precision - 1, 4); w.put_signed(shift, 5)
    for c in coef:
      w.put_signed(c, precision)
    x = [int(v) for v in samples]
    res = [x[i] - ((coef[0] * x[i - 1] + coef[1] * x[i - 2]) >> shift) for i in range(order, n)]
    w.put(1, 2); w.put(0, 4); w.put(6, 5); rice(w, res, 6)      # Rice2 coding, one partition
  else:
    raise ValueError(kind)


def encode_flac(pcm, rate, bps, blocks):
  """pcm [n, channels] ints; blocks = list of (blocksize, channel_assignment, [subframe kind per channel])."""
  n, channels = pcm.shape
  width = (bps + 7) // 8
  md5 = hashlib.md5(b''.join(int(v).to_bytes(4, 'little', signed=True)[:width] for v in pcm.reshape(-1))).digest()
  info = BitWriter()
  info.put(16, 16); info.put(65535, 16); info.put(0, 24); info.put(0, 24)
  info.put(rate, 20); info.put(channels - 1, 3); info.put(bps - 1, 5); info.put(n, 36)
  out = b'fLaC' + bytes([0x00, 0, 0, 34]) + info.tobytes() + md5
  out += bytes([0x84, 0, 0, 4]) + b'\x00' * 4            # a PADDING block, marked last
  pos = 0
  for number, (size, assignment, kinds) in enumerate(blocks):
    w = BitWriter()
    w.put(0x3ffe, 14); w.put(0, 1); w.put(0, 1)
    w.put(7, 4); w.put(0, 4)                               # 16-bit explicit block size, rate from STREAMINFO
    w.put(assignment if assignment >= 8 else channels - 1, 4)
    w.put({8: 1, 12: 2, 16: 4, 20: 5, 24: 6}[bps], 3); w.put(0, 1)
    w.put(number, 8)                                       # frame number < 128: one UTF-8 byte
    w.put(size - 1, 16)
    header = w.tobytes()
    w.put(crc(header, 0x07, 8), 8)
    blk = pcm[pos:pos + size].astype(np.int64)
    chans = [blk[:, c] for c in range(channels)]
    widths = [bps] * channels
    if assignment == 8:
      chans, widths = [chans[0], chans[0] - chans[1]], [bps, bps + 1]
    elif assignment == 9:
      chans, widths = [chans[0] - chans[1], chans[1]], [bps + 1, bps]
    elif assignment == 10:
      chans, widths = [(chans[0] + chans[1]) >> 1, chans[0] - chans[1]], [bps, bps + 1]
    for c in range(channels):
      subframe(w, chans[c], widths[c], kinds[c])
    w.align()
    body = w.tobytes()
    out += body + struct.pack('>H', crc(body, 0x8005, 16))
    pos += size
  assert pos == n
  return out


def _decode(path):
  from speecht_b200.preprocessing import read_flac
  return read_flac(path)


def test_flac_decoder_against_streams_from_a_reference_encoder(tmp_path):
  rng = np.random.default_rng(0)
  t = np.arange(1000)
  left = (6000 * np.sin(t * 0.05) + rng.integers(-40, 40, size=1000)).astype(np.int64)
  right = (left * 0.7 + 300 * np.cos(t * 0.11)).astype(np.int64)
  stereo = np.stack([left, right], axis=1)
  blocks = [(192, 1, ['verbatim', 'fixed2']), (200, 8, ['fixed1', 'fixed3']), (208, 9, ['fixed4', 'lpc']),
            (200, 10, ['lpc', 'fixed0']), (200, 1, ['fixed2', 'verbatim'])]
  path = tmp_path / 'stereo.flac'
  path.write_bytes(encode_flac(stereo, 16000, 16, blocks))
  audio, rate = _decode(str(path))                        # raises when the MD5 of the decoded PCM does not match
  assert rate == 16000 and audio.shape == (1000,)
  np.testing.assert_allclose(audio, stereo.mean(axis=1) / 32768.0, atol=1e-7)

  mono = np.concatenate([np.full(64, -1234), (rng.integers(-2000, 2000, size=136) // 8) * 8,
                         (3000 * np.sin(np.arange(300) * 0.2)).astype(np.int64)])[:, None]
  path = tmp_path / 'mono24.flac'
  path.write_bytes(encode_flac(mono, 22050, 24, [(64, 0, ['constant']), (136, 0, ['wasted']), (300, 0, ['lpc'])]))
  audio, rate = _decode(str(path))
  assert rate == 22050
  np.testing.assert_allclose(audio, mono[:, 0] / float(1 << 23), atol=1e-9)

  # corruption is detected (frame CRC), not decoded into noise
  data = bytearray(encode_flac(stereo, 16000, 16, blocks))
  data[len(data) // 2] ^= 0x10
  bad = tmp_path / 'bad.flac'
  bad.write_bytes(bytes(data))
  with pytest.raises(ValueError):
    _decode(str(bad))
  (tmp_path / 'not.flac').write_bytes(b'RIFF' + b'\x00' * 100)
  with pytest.raises(ValueError):
    _decode(str(tmp_path / 'not.flac'))


def test_load_audio_resamples_to_22050_like_librosa_load(tmp_path):
  from speecht_b200.preprocessing import SpeechCorpusReader, load_audio, resample
  n, sr = 16000, 16000
  tone = np.sin(2 * np.pi * 440.0 * np.arange(n) / sr)
  x = (tone * 12000).astype(np.int16)
  with wave.open(str(tmp_path / 'a.wav'), 'wb') as f:
    f.setnchannels(1); f.setsampwidth(2); f.setframerate(sr); f.writeframes(x.tobytes())
  audio, rate = load_audio(str(tmp_path / 'a.wav'))
  assert rate == 22050 and audio.dtype == np.float32
  assert audio.shape == (-(-n * 22050 // sr),)              # ceil(n * target / orig), librosa's output length
  # band-limited: the 440 Hz tone comes out as a 440 Hz tone of the same amplitude
  want = np.sin(2 * np.pi * 440.0 * np.arange(audio.shape[0]) / 22050.0) * (12000 / 32768.0)
  assert np.max(np.abs(audio[200:-200] - want[200:-200])) < 2e-3
  assert resample(np.ones(10, np.float32), 22050).shape == (10,)          # already at the target rate: untouched
  assert SpeechCorpusReader(str(tmp_path))._load_audio is load_audio       # the default loader of `preprocess`


@pytest.mark.skipif(not os.path.exists(REF_FLAC), reason='reference checkout not present')
def test_reference_librispeech_fixture_known_answers():
  """The reference's own fixture and known answer (test_speechCorpusReader.py:40-45): its LibriSpeech FLAC loads to
  114881 samples (librosa.load resamples 16 kHz -> 22050 Hz).  The decode itself is verified bit-exactly by the MD5
  signature inside the file (read_flac raises on mismatch): LPC subframes and partitioned Rice residuals of a real
  encoder, not of the toy encoder above."""
  from speecht_b200.preprocessing import load_audio, read_flac
  raw, rate = read_flac(REF_FLAC)
  assert rate == 16000 and raw.shape == (83360,)
  audio, sr = load_audio(REF_FLAC)
  assert sr == 22050 and audio.shape == (114881,)
  assert float(np.abs(audio).max()) < 1.0 and float(np.abs(audio).max()) > 0.05
