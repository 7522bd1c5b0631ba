// Host-side FLAC decoder for the `preprocess` caller of the hot path (reference preprocessing.py:169 loads LibriSpeech
// .flac files through librosa.load; the image has no soundfile / audioread / ffmpeg, so the decoder lives here).
// Plain C++, no CUDA: native FLAC streams (fLaC marker, STREAMINFO, frames with CONSTANT / VERBATIM / FIXED / LPC
// subframes, Rice and Rice2 residuals with escape partitions, independent / left-side / side-right / mid-side stereo,
// wasted bits, 4..32 bits per sample), frame CRC-8 / CRC-16 checked.  The decoded PCM can be verified against the MD5
// signature STREAMINFO carries (done by the Python caller with hashlib).
#include "st_common.cuh"
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

struct BitReader {
  const uint8_t* p;
  size_t n, pos;        // byte position
  uint64_t acc;         // bit accumulator (msb first)
  int bits;             // valid bits in acc
  bool bad;
  BitReader(const uint8_t* data, size_t len, size_t start) : p(data), n(len), pos(start), acc(0), bits(0), bad(false) {}
  void fill() {
    while (bits <= 56 && pos < n) { acc |= (uint64_t)p[pos++] << (56 - bits); bits += 8; }
  }
  uint32_t read(int k) {                      // k in [0, 32]
    if (k == 0) return 0;
    if (bits < k) fill();
    if (bits < k) { bad = true; return 0; }
    const uint32_t v = (uint32_t)(acc >> (64 - k));
    acc <<= k;
    bits -= k;
    return v;
  }
  int32_t read_signed(int k) {
    if (k == 0) return 0;
    const uint32_t v = read(k);
    return (int32_t)(v << (32 - k)) >> (32 - k);
  }
  uint32_t read_unary() {                     // number of 0 bits before the next 1 bit
    uint32_t count = 0;
    for (;;) {
      if (bits == 0) fill();
      if (bits == 0) { bad = true; return count; }
      if (acc == 0) { count += bits; bits = 0; continue; }
      const int lz = __builtin_clzll(acc);
      if (lz >= bits) { count += bits; acc = 0; bits = 0; continue; }
      count += lz;
      acc = lz == 63 ? 0 : acc << (lz + 1);          // a shift by 64 is undefined
      bits -= lz + 1;
      return count;
    }
  }
  void align_byte() { const int r = bits & 7; acc <<= r; bits -= r; }
  size_t byte_pos() const { return pos - (size_t)(bits >> 3); }   // valid after align_byte
};

uint8_t crc8(const uint8_t* d, size_t n) {
  uint8_t c = 0;
  for (size_t i = 0; i < n; ++i) {
    c ^= d[i];
    for (int b = 0; b < 8; ++b) c = (c & 0x80) ? (uint8_t)((c << 1) ^ 0x07) : (uint8_t)(c << 1);
  }
  return c;
}

uint16_t crc16(const uint8_t* d, size_t n) {
  uint16_t c = 0;
  for (size_t i = 0; i < n; ++i) {
    c ^= (uint16_t)d[i] << 8;
    for (int b = 0; b < 8; ++b) c = (c & 0x8000) ? (uint16_t)((c << 1) ^ 0x8005) : (uint16_t)(c << 1);
  }
  return c;
}

struct StreamInfo {
  int sample_rate, channels, bits;
  int64_t total_samples;
  uint8_t md5[16];
  size_t first_frame;
};

int parse_header(const uint8_t* d, size_t n, StreamInfo* si) {
  if (n < 42 || memcmp(d, "fLaC", 4) != 0) { st_set_error("not a native FLAC stream (no fLaC marker)"); return ST_ERR_INVALID_ARG; }
  size_t pos = 4;
  bool have = false;
  for (;;) {
    if (pos + 4 > n) { st_set_error("FLAC: truncated metadata"); return ST_ERR_INVALID_ARG; }
    const bool last = d[pos] & 0x80;
    const int type = d[pos] & 0x7f;
    const size_t len = ((size_t)d[pos + 1] << 16) | ((size_t)d[pos + 2] << 8) | d[pos + 3];
    pos += 4;
    if (pos + len > n) { st_set_error("FLAC: truncated metadata block"); return ST_ERR_INVALID_ARG; }
    if (type == 0 && len >= 34) {
      const uint8_t* s = d + pos;
      si->sample_rate = (s[10] << 12) | (s[11] << 4) | (s[12] >> 4);
      si->channels = ((s[12] >> 1) & 7) + 1;
      si->bits = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
      si->total_samples = ((int64_t)(s[13] & 0x0f) << 32) | ((int64_t)s[14] << 24) | (s[15] << 16) | (s[16] << 8) | s[17];
      memcpy(si->md5, s + 18, 16);
      have = true;
    }
    pos += len;
    if (last) break;
  }
  if (!have) { st_set_error("FLAC: no STREAMINFO block"); return ST_ERR_INVALID_ARG; }
  si->first_frame = pos;
  return ST_OK;
}

bool decode_residual(BitReader& br, int32_t* out, int blocksize, int order) {
  const int method = br.read(2);
  if (method > 1) return false;
  const int pbits = method == 0 ? 4 : 5;
  const int escape = method == 0 ? 15 : 31;
  const int porder = br.read(4);
  const int parts = 1 << porder;
  if ((blocksize >> porder) << porder != blocksize && porder > 0) return false;
  int idx = order;
  for (int pt = 0; pt < parts; ++pt) {
    int count = (blocksize >> porder) - (pt == 0 ? order : 0);
    if (count < 0 || idx + count > blocksize) return false;
    const int k = br.read(pbits);
    if (k == escape) {
      const int raw = br.read(5);
      for (int i = 0; i < count; ++i) out[idx++] = br.read_signed(raw);
    } else {
      for (int i = 0; i < count; ++i) {
        const uint32_t q = br.read_unary();
        const uint32_t u = (q << k) | br.read(k);
        out[idx++] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
      }
    }
    if (br.bad) return false;
  }
  return idx == blocksize;
}

bool decode_subframe(BitReader& br, int32_t* out, int blocksize, int bps) {
  if (br.read(1) != 0) return false;
  const int type = br.read(6);
  int wasted = 0;
  if (br.read(1)) wasted = (int)br.read_unary() + 1;
  bps -= wasted;
  if (bps <= 0) return false;
  if (type == 0) {                                   // CONSTANT
    const int32_t v = br.read_signed(bps);
    for (int i = 0; i < blocksize; ++i) out[i] = v;
  } else if (type == 1) {                            // VERBATIM
    for (int i = 0; i < blocksize; ++i) out[i] = br.read_signed(bps);
  } else if (type >= 8 && type <= 12) {              // FIXED, order type - 8
    const int order = type - 8;
    if (order > blocksize) return false;
    for (int i = 0; i < order; ++i) out[i] = br.read_signed(bps);
    if (!decode_residual(br, out, blocksize, order)) return false;
    for (int i = order; i < blocksize; ++i) {
      int64_t pred = 0;
      switch (order) {
        case 1: pred = out[i - 1]; break;
        case 2: pred = 2 * (int64_t)out[i - 1] - out[i - 2]; break;
        case 3: pred = 3 * (int64_t)out[i - 1] - 3 * (int64_t)out[i - 2] + out[i - 3]; break;
        case 4: pred = 4 * (int64_t)out[i - 1] - 6 * (int64_t)out[i - 2] + 4 * (int64_t)out[i - 3] - out[i - 4]; break;
        default: break;
      }
      out[i] = (int32_t)(out[i] + pred);
    }
  } else if (type >= 32) {                           // LPC, order type - 31
    const int order = type - 31;
    if (order > blocksize) return false;
    for (int i = 0; i < order; ++i) out[i] = br.read_signed(bps);
    const int precision = br.read(4) + 1;
    if (precision == 16) return false;
    const int shift = br.read_signed(5);
    if (shift < 0) return false;
    int32_t coef[32];
    for (int j = 0; j < order; ++j) coef[j] = br.read_signed(precision);
    if (!decode_residual(br, out, blocksize, order)) return false;
    for (int i = order; i < blocksize; ++i) {
      int64_t sum = 0;
      for (int j = 0; j < order; ++j) sum += (int64_t)coef[j] * out[i - 1 - j];
      out[i] = (int32_t)(out[i] + (sum >> shift));
    }
  } else {
    return false;                                    // reserved subframe type
  }
  if (wasted)
    for (int i = 0; i < blocksize; ++i) out[i] = (int32_t)((uint32_t)out[i] << wasted);
  return !br.bad;
}

}  // namespace

// STREAMINFO of a FLAC stream held in host memory: info = {sample_rate, channels, bits_per_sample}, total samples per
// channel (0 = unknown), MD5 of the unencoded PCM (16 bytes).
ST_API int st_flac_info_host(const uint8_t* data, size_t nbytes, int32_t* info, int64_t* total_samples, uint8_t* md5) {
  ST_CHECK_ARG(data && info && total_samples && md5, "st_flac_info_host: null pointer");
  StreamInfo si;
  const int rc = parse_header(data, nbytes, &si);
  if (rc) return rc;
  info[0] = si.sample_rate; info[1] = si.channels; info[2] = si.bits;
  *total_samples = si.total_samples;
  memcpy(md5, si.md5, 16);
  return ST_OK;
}

// Decodes every frame into out[sample][channel] (interleaved int32, capacity in samples per channel); returns the
// number of samples per channel decoded in *decoded.
ST_API int st_flac_decode_host(const uint8_t* data, size_t nbytes, int32_t* out, int64_t capacity, int64_t* decoded) {
  ST_CHECK_ARG(data && out && decoded, "st_flac_decode_host: null pointer");
  StreamInfo si;
  int rc = parse_header(data, nbytes, &si);
  if (rc) return rc;
  static const int kBlock[16] = {0, 192, 576, 1152, 2304, 4608, 0, 0, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768};
  static const int kBits[8] = {0, 8, 12, 0, 16, 20, 24, 32};
  std::vector<int32_t> chan[8];
  size_t pos = si.first_frame;
  int64_t done = 0;
  while (pos + 6 <= nbytes) {
    if (data[pos] != 0xff || (data[pos + 1] & 0xfe) != 0xf8) {
      st_set_error("FLAC: lost frame sync at byte %zu", pos);
      return ST_ERR_INVALID_ARG;
    }
    BitReader br(data, nbytes, pos);
    br.read(16);
    const int bs_code = br.read(4), sr_code = br.read(4), ch_code = br.read(4), ss_code = br.read(3);
    br.read(1);
    // UTF-8 style frame / sample number
    const uint32_t first = br.read(8);
    int ones = 0;
    while (ones < 8 && (first & (0x80u >> ones))) ++ones;
    if (ones == 1 || ones == 8) { st_set_error("FLAC: bad frame number coding at byte %zu", pos); return ST_ERR_INVALID_ARG; }
    for (int i = 0; i < (ones ? ones - 1 : 0); ++i) br.read(8);
    int blocksize = kBlock[bs_code];
    if (bs_code == 6) blocksize = br.read(8) + 1;
    else if (bs_code == 7) blocksize = br.read(16) + 1;
    if (sr_code == 12) br.read(8);
    else if (sr_code == 13 || sr_code == 14) br.read(16);
    const size_t hdr_end = br.byte_pos();
    const uint8_t want8 = (uint8_t)br.read(8);
    if (br.bad || blocksize <= 0 || crc8(data + pos, hdr_end - pos) != want8) {
      st_set_error("FLAC: bad frame header at byte %zu", pos);
      return ST_ERR_INVALID_ARG;
    }
    const int bps = ss_code == 0 ? si.bits : kBits[ss_code];
    const int nch = ch_code < 8 ? ch_code + 1 : 2;
    if (bps == 0 || nch != si.channels || ch_code > 10) {
      st_set_error("FLAC: unsupported frame (channels code %d, sample size code %d)", ch_code, ss_code);
      return ST_ERR_UNSUPPORTED;
    }
    for (int c = 0; c < nch; ++c) {
      chan[c].resize(blocksize);
      // the side channel of a stereo-decorrelated frame carries one more bit
      const bool side = (ch_code == 8 && c == 1) || (ch_code == 9 && c == 0) || (ch_code == 10 && c == 1);
      if (!decode_subframe(br, chan[c].data(), blocksize, bps + (side ? 1 : 0))) {
        st_set_error("FLAC: corrupt subframe in the frame at byte %zu", pos);
        return ST_ERR_INVALID_ARG;
      }
    }
    br.align_byte();
    const size_t body_end = br.byte_pos();
    const uint16_t want16 = (uint16_t)br.read(16);
    if (br.bad || crc16(data + pos, body_end - pos) != want16) {
      st_set_error("FLAC: frame CRC mismatch at byte %zu", pos);
      return ST_ERR_INVALID_ARG;
    }
    if (ch_code == 8) {                      // left, side
      for (int i = 0; i < blocksize; ++i) chan[1][i] = chan[0][i] - chan[1][i];
    } else if (ch_code == 9) {               // side, right
      for (int i = 0; i < blocksize; ++i) chan[0][i] = chan[0][i] + chan[1][i];
    } else if (ch_code == 10) {              // mid, side
      for (int i = 0; i < blocksize; ++i) {
        const int32_t side = chan[1][i];
        const int32_t mid = (int32_t)(((uint32_t)chan[0][i] << 1) | (side & 1));
        chan[0][i] = (mid + side) >> 1;
        chan[1][i] = (mid - side) >> 1;
      }
    }
    if (done + blocksize > capacity) {
      st_set_error("FLAC: output buffer too small (%lld samples)", (long long)capacity);
      return ST_ERR_INVALID_ARG;
    }
    for (int i = 0; i < blocksize; ++i)
      for (int c = 0; c < nch; ++c) out[(done + i) * nch + c] = chan[c][i];
    done += blocksize;
    pos = br.byte_pos();
    if (si.total_samples > 0 && done >= si.total_samples) break;
  }
  if (si.total_samples > 0 && done > si.total_samples) done = si.total_samples;   // last block may carry padding
  *decoded = done;
  return ST_OK;
}
