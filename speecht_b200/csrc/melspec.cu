// a1-a3 -- calc_power_spectrogram (reference preprocessing.py:36-58):
//   librosa.feature.melspectrogram(y, sr, n_fft=512, hop_length=160, n_mels=128)   (centered, reflect-padded
//   frames, periodic Hann, |rFFT|^2, Slaney mel basis)  ->  power_to_db(ref=np.max, amin=1e-10, top_db=80)
//   ->  (x - mean) / std over the whole utterance  ->  [time, n_mels].
//
// Kernels (the per-utterance global dependencies -- the dB reference is the utterance max, the z-score needs the
// utterance mean/std -- force three passes over a 0.5 MB/utterance matrix; everything stays L2-resident):
//   0. mel_ranges      : non-zero bin range of every mel filter (the Slaney triangles are ~2-15 bins wide).
//   1. mel_power       : one warp per PAIR of frames: the two real frames ride one 512-point complex FFT
//                        (in-place radix-2, twiddles in shared memory), are separated by Hermitian symmetry,
//                        squared, and contracted with the sparse mel filters; utterance max via atomicMax.
//   2. db_stats        : dB conversion + clip, double-precision sum / sum-of-squares per utterance.
//   3. normalize       : z-score in place, zero the batch-padding frames (speech_input.py:39-43).
// Algorithmic bytes per utterance = 4*n_samples + 4*T*n_mels; HBM-bound in principle,
// launch/latency-bound at these sizes.
#include "st_common.cuh"

namespace {

constexpr int NFFT = 512;
constexpr int NBINS = NFFT / 2 + 1;
constexpr int WARPS = 8;

struct MelWs {
  int* mel_start;     // [n_mels]
  int* mel_end;       // [n_mels]
  unsigned* umax;     // [B] float bits of the utterance max (power >= 0 so uint order == float order)
  double* sums;       // [B][2]
};

__global__ void mel_ranges_kernel(const float* __restrict__ basis, int n_mels, int* start, int* end) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_mels) return;
  int s = NBINS, e = 0;
  for (int k = 0; k < NBINS; ++k) {
    if (basis[(int64_t)m * NBINS + k] != 0.f) {
      if (k < s) s = k;
      e = k + 1;
    }
  }
  if (s > e) s = e = 0;
  start[m] = s;
  end[m] = e;
}

__device__ __forceinline__ float sample_reflect(const float* __restrict__ w, int n, int i) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return w[i];
}

// grid (ceil(Tmax/(2*WARPS)), B), block WARPS*32.  smem: float2 tw[256]; per warp: float2 z[512]; float P[2][NBINS+1]
__global__ void __launch_bounds__(WARPS * 32)
mel_power_kernel(const float* __restrict__ wav, int64_t wav_stride, const int32_t* __restrict__ n_samples, int hop,
                 const float* __restrict__ basis, const int* __restrict__ mel_start, const int* __restrict__ mel_end,
                 int n_mels, float* __restrict__ out, int T_max, unsigned* __restrict__ umax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tw = reinterpret_cast<float2*>(smem_raw);                       // [256]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* z = tw + 256 + warp * NFFT;                                      // [512]
  float* P = reinterpret_cast<float*>(tw + 256 + WARPS * NFFT) + warp * 2 * (NBINS + 1);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    float s, c;
    sincospif(-2.f * (float)i / (float)NFFT, &s, &c);
    tw[i] = make_float2(c, s);
  }
  __syncthreads();
  const int b = blockIdx.y;
  const int n = n_samples[b];
  const int T = 1 + n / hop;
  const int f0 = (blockIdx.x * WARPS + warp) * 2;
  if (f0 >= T || f0 >= T_max) return;
  const bool has_b = (f0 + 1 < T) && (f0 + 1 < T_max);
  const float* w = wav + (int64_t)b * wav_stride;

  // load: z[bitrev(i)] = hann[i] * (frameA[i] + j frameB[i])
  for (int i = lane; i < NFFT; i += 32) {
    float s, c;
    sincospif(2.f * (float)i / (float)NFFT, &s, &c);
    const float hann = 0.5f - 0.5f * c;
    const int pa = f0 * hop + i - NFFT / 2;
    const float xa = sample_reflect(w, n, pa) * hann;
    const float xb = has_b ? sample_reflect(w, n, pa + hop) * hann : 0.f;
    z[__brev((unsigned)i) >> 23] = make_float2(xa, xb);
  }
  __syncwarp();
#pragma unroll 1
  for (int half = 1; half < NFFT; half <<= 1) {
    const int tw_step = (NFFT / 2) / half;
    for (int j = lane; j < NFFT / 2; j += 32) {
      const int pos = j & (half - 1);
      const int i0 = ((j - pos) << 1) + pos;
      const int i1 = i0 + half;
      const float2 wv = tw[pos * tw_step];
      const float2 a = z[i0], c = z[i1];
      const float2 t = make_float2(c.x * wv.x - c.y * wv.y, c.x * wv.y + c.y * wv.x);
      z[i0] = make_float2(a.x + t.x, a.y + t.y);
      z[i1] = make_float2(a.x - t.x, a.y - t.y);
    }
    __syncwarp();
  }
  // separate the two real spectra: XA[k] = (Z[k] + conj(Z[N-k]))/2, XB[k] = (Z[k] - conj(Z[N-k]))/(2j)
  for (int k = lane; k < NBINS; k += 32) {
    const float2 zk = z[k];
    const float2 zn = z[(NFFT - k) & (NFFT - 1)];
    const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
    const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);
    P[k] = ar * ar + ai * ai;
    P[NBINS + 1 + k] = br * br + bi * bi;
  }
  __syncwarp();
  float local_max = 0.f;
  for (int m = lane; m < n_mels; m += 32) {
    const int s = mel_start[m], e = mel_end[m];
    const float* row = basis + (int64_t)m * NBINS;
    float accA = 0.f, accB = 0.f;
    for (int k = s; k < e; ++k) {
      const float wv = __ldg(row + k);
      accA = fmaf(wv, P[k], accA);
      accB = fmaf(wv, P[NBINS + 1 + k], accB);
    }
    float* o = out + ((int64_t)b * T_max + f0) * n_mels + m;
    o[0] = accA;
    local_max = fmaxf(local_max, accA);
    if (has_b) {
      o[n_mels] = accB;
      local_max = fmaxf(local_max, accB);
    }
  }
  local_max = warp_max(local_max);
  if (lane == 0) atomicMax(umax + b, __float_as_uint(local_max));
}

// grid (chunks, B)
__global__ void __launch_bounds__(256)
db_stats_kernel(float* __restrict__ out, const int32_t* __restrict__ n_samples, int hop, int n_mels, int T_max,
                const unsigned* __restrict__ umax, double* __restrict__ sums, float amin, float top_db) {
  __shared__ double part[8][2];
  const int b = blockIdx.y;
  const int T = min(1 + n_samples[b] / hop, T_max);
  const int64_t count = (int64_t)T * n_mels;
  float* x = out + (int64_t)b * T_max * n_mels;
  const float ref_db = 10.f * log10f(fmaxf(amin, __uint_as_float(umax[b])));
  // log_spec.max() = 10log10(max(amin, max S)) - ref_db = 0, so the floor is -top_db
  double s1 = 0.0, s2 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    float v = 10.f * log10f(fmaxf(amin, x[i])) - ref_db;
    v = fmaxf(v, -top_db);
    x[i] = v;
    s1 += (double)v;
    s2 += (double)v * (double)v;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { part[threadIdx.x >> 5][0] = s1; part[threadIdx.x >> 5][1] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int w = 0; w < 8; ++w) { a += part[w][0]; c += part[w][1]; }
    atomicAdd(sums + 2 * b, a);
    atomicAdd(sums + 2 * b + 1, c);
  }
}

__global__ void __launch_bounds__(256)
normalize_kernel(float* __restrict__ out, const int32_t* __restrict__ n_samples, int hop, int n_mels, int T_max,
                 const double* __restrict__ sums, int32_t* __restrict__ out_frames) {
  const int b = blockIdx.y;
  const int T = min(1 + n_samples[b] / hop, T_max);
  const int64_t count = (int64_t)T * n_mels, total = (int64_t)T_max * n_mels;
  const double mean = sums[2 * b] / (double)count;
  const double var = fmax(sums[2 * b + 1] / (double)count - mean * mean, 0.0);
  const float fm = (float)mean, inv = (float)(1.0 / sqrt(var));
  float* x = out + (int64_t)b * total;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] = i < count ? (x[i] - fm) * inv : 0.f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && out_frames) out_frames[b] = T;
}

size_t ws_layout(int B, int n_mels, MelWs* ws, char* base) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
  size_t o_start = take((size_t)n_mels * sizeof(int));
  size_t o_end = take((size_t)n_mels * sizeof(int));
  size_t o_max = take((size_t)B * sizeof(unsigned));
  size_t o_sums = take((size_t)B * 2 * sizeof(double));
  if (ws) {
    ws->mel_start = reinterpret_cast<int*>(base + o_start);
    ws->mel_end = reinterpret_cast<int*>(base + o_end);
    ws->umax = reinterpret_cast<unsigned*>(base + o_max);
    ws->sums = reinterpret_cast<double*>(base + o_sums);
  }
  return off;
}

}  // namespace

ST_API size_t st_melspec_workspace_bytes(int B, int max_samples, int n_fft, int hop, int n_mels) {
  (void)max_samples; (void)n_fft; (void)hop;
  return ws_layout(B, n_mels, nullptr, nullptr);
}

ST_API int st_melspec(const float* wav, int64_t wav_stride, const int32_t* n_samples, int B, int max_samples,
                      const float* mel_basis, int n_fft, int hop, int n_mels, float* out, int T_max,
                      int32_t* out_frames, void* workspace, size_t workspace_bytes, st_stream_t stream) {
  ST_CHECK_ARG(wav && n_samples && mel_basis && out && workspace, "st_melspec: null pointer");
  ST_CHECK_ARG(n_fft == NFFT, "st_melspec: only n_fft=512 is built (got %d)", n_fft);
  ST_CHECK_ARG(B > 0 && hop > 0 && n_mels > 0 && T_max > 0 && max_samples > n_fft / 2,
               "st_melspec: bad dimensions (audio must be longer than n_fft/2 for reflect padding)");
  ST_CHECK_ARG(T_max >= 1 + max_samples / hop, "st_melspec: T_max %d < 1 + max_samples/hop", T_max);
  MelWs ws;
  const size_t need = ws_layout(B, n_mels, &ws, static_cast<char*>(workspace));
  ST_CHECK_ARG(workspace_bytes >= need, "st_melspec: workspace %zu < required %zu", workspace_bytes, need);
  cudaStream_t s = st_cu(stream);
  ST_CUDA_CALL(cudaMemsetAsync(ws.umax, 0, (size_t)B * sizeof(unsigned), s));
  ST_CUDA_CALL(cudaMemsetAsync(ws.sums, 0, (size_t)B * 2 * sizeof(double), s));
  mel_ranges_kernel<<<(n_mels + 127) / 128, 128, 0, s>>>(mel_basis, n_mels, ws.mel_start, ws.mel_end);
  ST_CUDA_LAUNCH_CHECK("mel_ranges_kernel");
  const size_t smem = 256 * sizeof(float2) + (size_t)WARPS * NFFT * sizeof(float2) +
                      (size_t)WARPS * 2 * (NBINS + 1) * sizeof(float);
  ST_CUDA_CALL(cudaFuncSetAttribute(mel_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((T_max + 2 * WARPS - 1) / (2 * WARPS), B);
  mel_power_kernel<<<grid, WARPS * 32, smem, s>>>(wav, wav_stride, n_samples, hop, mel_basis, ws.mel_start, ws.mel_end,
                                                  n_mels, out, T_max, ws.umax);
  ST_CUDA_LAUNCH_CHECK("mel_power_kernel");
  int chunks = (int)(((int64_t)T_max * n_mels + 256 * 16 - 1) / (256 * 16));
  chunks = chunks < 1 ? 1 : (chunks > 64 ? 64 : chunks);
  db_stats_kernel<<<dim3(chunks, B), 256, 0, s>>>(out, n_samples, hop, n_mels, T_max, ws.umax, ws.sums, 1e-10f, 80.f);
  ST_CUDA_LAUNCH_CHECK("db_stats_kernel");
  normalize_kernel<<<dim3(chunks, B), 256, 0, s>>>(out, n_samples, hop, n_mels, T_max, ws.sums, out_frames);
  ST_CUDA_LAUNCH_CHECK("normalize_kernel");
  return ST_OK;
}
