// a8/a9 -- CTC loss forward + gradient (replaces tf.nn.ctc_loss and the gradient TF registers for it,
// reference speech_model.py:74-75,78).  TF1 defaults: ctc_merge_repeated=True, blank = num_classes-1,
// softmax applied internally, frames t >= seq_len get zero gradient.
//
// Three kernels:
//   1. ctc_softmax     : one warp per (b,t) row, warp-shuffle max / sum over the 29 classes -> class probabilities.
//   2. ctc_alpha_beta  : grid (B,2): CTA (b,0) runs the alpha recursion, CTA (b,1) the beta recursion (only alpha
//                        when no gradient is wanted), 500-1500 strictly serial steps with one __syncthreads per
//                        step, in the LINEAR domain on scaled floats p*2^e (fp32 mantissa in [1,2), int32 exponent).
//   3. ctc_grad        : one warp per (b,t) row: occupancy per class from alpha*beta/p(z|x), grad = softmax - occ.
// Algorithmic bytes = read logits + write grad = 2*T*B*C*4; the alpha/beta lattices (T*S x 8 bytes each) are
// workspace traffic on top.  The recursion is latency-bound by construction.
#include "st_common.cuh"
#include <math.h>
#include <stdlib.h>
#include <vector>

namespace {

// probs[b][t][c] = max(softmax(logits[t][b][:])[c], FLT_MIN): the emission probabilities the recursions multiply by,
// kept NORMAL so that products with a mantissa in [1,6) never go denormal.
__global__ void __launch_bounds__(256)
ctc_softmax_kernel(const float* __restrict__ logits, int64_t stride_t, int64_t stride_b, int T, int B, int C,
                   const int32_t* __restrict__ seq_len, float* __restrict__ probs) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * T) return;
  const int b = row / T, t = row - b * T;
  if (t >= seq_len[b]) return;
  const float* src = logits + (int64_t)t * stride_t + (int64_t)b * stride_b;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, src[c]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += expf(src[c] - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  float* dst = probs + (int64_t)row * C;
  for (int c = lane; c < C; c += 32) dst[c] = fmaxf(expf(src[c] - mx) * inv, 1.17549435e-38f);
}

// ---- scaled-float arithmetic for the alpha / beta recursions --------------------------------------------------
// A lattice value (a probability that reaches 1e-600 and below) is kept in the LINEAR domain as p * 2^e with p in
// [1,2) in fp32 and e an int32.  ZERO is not special: it is p = 1 with the exponent kZeroExp = -2^28, a number so
// small that it vanishes in every sum -- no zero tests on the serial chain.  One recursion step is
//   em  = max(e0, e1, e2)                        (one 3-input integer max)
//   sum = sum_i p_i * 2^max(e_i - em, -126)      (per term: subtract, clamp, build the power of two, multiply)
//   x   = sum * y                                (y = emission probability, a normal fp32 in [2^-126, 1])
//   (p, e) = (x / 2^floor(log2 x), em + floor(log2 x))
// ~30 instructions, no transcendental and no fp64; relative error 6e-8 per step, ~1e-6 after 1500 steps.  The term
// with the largest exponent enters with scale 1, so sum is in [1,6) and x is a normal number: no special cases.
struct SF { float p; int e; };
constexpr int kZeroExp = -(1 << 28);

__device__ __forceinline__ float sf_scaled(float p, int de) {          // p * 2^max(de, -126), de <= 0
  // p is in [1,2) (biased exponent 127), so the scaling is an integer add on the exponent field: the result stays a
  // normal number (127 + de >= 1) and is bit-identical to the multiplication by the exact power of two
  return __int_as_float(__float_as_int(p) + (int)((unsigned)max(de, -126) << 23));
}
__device__ __forceinline__ SF sf_norm(float x, int ebase) {            // x normal and positive
  const int bits = __float_as_int(x);
  const int ex = (bits >> 23) - 127;
  SF r;
  r.p = __int_as_float(bits - (ex << 23));
  r.e = ebase + ex;
  return r;
}
__device__ __forceinline__ SF sf_sum3(SF a, SF b, SF c) {              // a + b + c, normalised
  const int em = max(a.e, max(b.e, c.e));
  return sf_norm(sf_scaled(a.p, a.e - em) + sf_scaled(b.p, b.e - em) + sf_scaled(c.p, c.e - em), em);
}
__device__ __forceinline__ double sf_log(SF v) {                        // natural log; -inf for ZERO
  return v.e < kZeroExp / 2 ? -INFINITY : ((double)v.e + (double)log2f(v.p)) * 0.6931471805599453;
}
__device__ __forceinline__ SF sf_zero() { SF r; r.p = 1.f; r.e = kZeroExp; return r; }
__device__ __forceinline__ SF sf_load(const int2* p) { const int2 v = *p; SF r; r.p = __int_as_float(v.x); r.e = v.y; return r; }
__device__ __forceinline__ void sf_store(int2* p, SF v) { *p = make_int2(__float_as_int(v.p), v.e); }

// dynamic smem: int2 buf[2][s_pad + 4]; int ext[s_pad]; unsigned char skip[s_pad + 2]; [float probs_s[len*C]]
// NS = extended-label positions per thread (S <= NS*blockDim.x; the launcher sizes the block to the longest
// extended label so that NS == 1 up to 511 characters).  PROBS_SMEM: the utterance's probability rows are staged in
// shared memory up front (one coalesced pass) so that the serial recursion never waits on an L2 round trip.
// STORE: write the lattices for the gradient kernel (off for loss-only calls, which also skip the beta CTAs).
template <int NS, bool PROBS_SMEM, bool STORE>
__global__ void __launch_bounds__(1024)
ctc_alpha_beta_kernel(const float* __restrict__ probs, int T, int C, const int32_t* __restrict__ labels,
                      const int32_t* __restrict__ label_offsets, const int32_t* __restrict__ seq_len, int blank,
                      int s_pad, int2* __restrict__ alpha, int2* __restrict__ beta,
                      double* __restrict__ logp, float* __restrict__ loss, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nthr = blockDim.x;                      // = extended-label length rounded up to a warp (NS == 1)
  const int b = blockIdx.x;
  const bool is_beta = blockIdx.y == 1;
  const int l0 = label_offsets[b];
  const int L = label_offsets[b + 1] - l0;
  const int S = 2 * L + 1;
  int len = seq_len[b];
  len = len > T ? T : len;
  int2* buf = reinterpret_cast<int2*>(smem_raw);                        // [2][s_pad + 4]
  const int bstride = s_pad + 4;
  int* ext = reinterpret_cast<int*>(buf + 2 * bstride);                 // [s_pad]
  unsigned char* skip = reinterpret_cast<unsigned char*>(ext + s_pad);  // [s_pad + 2]
  float* probs_s = reinterpret_cast<float*>(smem_raw + (((size_t)2 * bstride * 8 + (size_t)s_pad * 4 + s_pad + 2 + 15) & ~(size_t)15));
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();

  // extended label sequence + feasibility (device-side mirror of st_ctc_validate_labels_host)
  int repeats = 0, bad = 0;
  for (int s = threadIdx.x; s < S + 2 && S <= s_pad; s += nthr) {
    int e = blank;
    if (s < S && (s & 1)) {
      e = labels[l0 + (s >> 1)];
      if (e < 0 || e >= C || e == blank) { bad = 1; e = blank; }
      if (s >= 3 && e == labels[l0 + (s >> 1) - 1]) repeats++;
    }
    if (s < S) ext[s] = e;
  }
  if (bad) atomicOr(&s_bad, 1);
  if (repeats) atomicAdd(&s_bad, repeats << 1);
  __syncthreads();
  const int n_rep = s_bad >> 1;
  const bool infeasible = (s_bad & 1) || (L + n_rep > len) || len < 0 || S > s_pad || S > NS * nthr;
  for (int s = threadIdx.x; s < S + 2 && S <= s_pad; s += nthr) {
    // skip[s]: transition s-2 -> s allowed
    skip[s] = (s >= 2 && s < S && ext[s] != blank && ext[s] != ext[s - 2]) ? 1 : 0;
  }
  const int2 zero = make_int2(__float_as_int(1.f), kZeroExp);
  for (int i = threadIdx.x; i < 2 * bstride; i += nthr) buf[i] = zero;
  const float* grow = probs + (int64_t)b * T * C;
  if (PROBS_SMEM && !infeasible && len > 0) {
    for (int i = threadIdx.x; i < len * C; i += nthr) probs_s[i] = grow[i];
  }
  __syncthreads();
  // emission probability y_t(c): a shared-memory load when the rows are staged (never a generic-address load)
  auto emis = [&](int t, int c) -> float {
    if constexpr (PROBS_SMEM) return probs_s[t * C + c];
    else return __ldg(grow + (int64_t)t * C + c);
  };

  if (infeasible || len <= 0) {
    if (!is_beta && threadIdx.x == 0) {
      status[b] = infeasible ? 1 : 0;
      loss[b] = infeasible ? INFINITY : 0.f;
      logp[b] = infeasible ? -INFINITY : 0.0;
    }
    return;
  }
  if (!is_beta && threadIdx.x == 0) status[b] = 0;

  int2* lat = (is_beta ? beta : alpha) + (int64_t)b * T * s_pad;
  int my_ext[NS];
  bool my_skip[NS], live[NS];
  float y[NS];                                      // emission probability of the NEXT step to consume
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int s = threadIdx.x + j * nthr;
    live[j] = s < S;
    my_ext[j] = live[j] ? ext[s] : blank;
    my_skip[j] = false;
    y[j] = 1.f;
  }

  if (!is_beta) {
    // alpha_0(s) = y_0(ext_s) for s < 2
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int s = threadIdx.x + j * nthr;
      if (live[j]) {
        my_skip[j] = skip[s] != 0;
        const SF v = s < 2 ? sf_norm(emis(0, my_ext[j]), 0) : sf_zero();
        sf_store(buf + 2 + s, v);                                       // slot 0, +2 pad so s-1, s-2 read ZERO
        if (STORE) sf_store(lat + s, v);
        if (len > 1) y[j] = emis(1, my_ext[j]);
      }
    }
    __syncthreads();
    int2* lat_t = lat + s_pad;                                          // row t of the lattice
    for (int t = 1; t < len; ++t, lat_t += s_pad) {
      const int2* prev = buf + ((t - 1) & 1) * bstride;
      int2* cur = buf + (t & 1) * bstride;
      float yn[NS];                                                      // emission of step t+1: off the chain
#pragma unroll
      for (int j = 0; j < NS; ++j) yn[j] = (live[j] && t + 1 < len) ? emis(t + 1, my_ext[j]) : 1.f;
      SF v[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int s = threadIdx.x + j * nthr;
        if (live[j]) {
          const SF a0 = sf_load(prev + 2 + s), a1 = sf_load(prev + 1 + s);
          SF a2 = sf_load(prev + s);
          if (!my_skip[j]) a2.e = kZeroExp;
          const int em = max(a0.e, max(a1.e, a2.e));
          const float sum = sf_scaled(a0.p, a0.e - em) + sf_scaled(a1.p, a1.e - em) + sf_scaled(a2.p, a2.e - em);
          v[j] = sf_norm(sum * y[j], em);
          sf_store(cur + 2 + s, v[j]);
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int s = threadIdx.x + j * nthr;
        if (STORE && live[j]) sf_store(lat_t + s, v[j]);                 // global store after the barrier: off the chain
        y[j] = yn[j];
      }
    }
    if (threadIdx.x == 0) {
      const int2* fin = buf + ((len - 1) & 1) * bstride;
      const SF total = sf_sum3(sf_load(fin + 2 + S - 1), S > 1 ? sf_load(fin + 2 + S - 2) : sf_zero(), sf_zero());
      const double lpz = sf_log(total);
      logp[b] = lpz;
      loss[b] = (float)(-lpz);
      if (total.e < kZeroExp / 2) status[b] = 1;
    }
  } else {
    // beta_{len-1}(s) = 1 for the last two positions; smem holds g_t(s) = beta_t(s) * y_t(s) for the step below
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int s = threadIdx.x + j * nthr;
      if (live[j]) {
        my_skip[j] = skip[s + 2] != 0;                                  // transition s -> s+2
        SF v = sf_zero();
        if (s >= S - 2) v.e = 0;
        if (STORE) sf_store(lat + (int64_t)(len - 1) * s_pad + s, v);
        sf_store(buf + ((len - 1) & 1) * bstride + s, sf_norm(v.p * emis(len - 1, my_ext[j]), v.e));
        if (len > 1) y[j] = emis(len - 2, my_ext[j]);
      }
    }
    __syncthreads();
    int2* lat_t = lat + (int64_t)(len - 2) * s_pad;
    for (int t = len - 2; t >= 0; --t, lat_t -= s_pad) {
      const int2* nxt = buf + ((t + 1) & 1) * bstride;                  // entries S..S+3 stay ZERO
      int2* cur = buf + (t & 1) * bstride;
      float yn[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) yn[j] = (live[j] && t > 0) ? emis(t - 1, my_ext[j]) : 1.f;
      SF v[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int s = threadIdx.x + j * nthr;
        if (live[j]) {
          const SF b0 = sf_load(nxt + s), b1 = sf_load(nxt + s + 1);
          SF b2 = sf_load(nxt + s + 2);
          if (!my_skip[j]) b2.e = kZeroExp;
          v[j] = sf_sum3(b0, b1, b2);
          sf_store(cur + s, sf_norm(v[j].p * y[j], v[j].e));
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int s = threadIdx.x + j * nthr;
        if (STORE && live[j]) sf_store(lat_t + s, v[j]);
        y[j] = yn[j];
      }
    }
  }
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One warp per (b,t) row.  grad row = grad_scale * (softmax - occupancy); zero for t >= seq_len[b].
// occupancy(s) = alpha*beta / p(z|x) = pa*pb * 2^(ea + eb - log2 p(z|x)): log2 p(z|x) is split ONCE per row into an
// integer and a fraction in [0,1), so the exponents subtract exactly in integers and nothing on the per-state path is
// double precision.  Label states accumulate into a PRIVATE row of bins per lane (no shared-memory atomics, whose
// same-class collisions serialise); the 32 rows are summed per class at the end.
__global__ void __launch_bounds__(256)
ctc_grad_kernel(const float* __restrict__ probs, const int2* __restrict__ alpha, const int2* __restrict__ beta,
                const double* __restrict__ logp, const int32_t* __restrict__ labels,
                const int32_t* __restrict__ label_offsets, const int32_t* __restrict__ seq_len,
                const int32_t* __restrict__ status, int T, int B, int C, int blank, int s_pad,
                float grad_scale, float* __restrict__ grad, int64_t stride_t, int64_t stride_b,
                __nv_bfloat16* __restrict__ planes, int n_planes, int c_pad) {
  extern __shared__ float bins_all[];                       // [8 warps][32 lanes][C + 1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= B * T) return;
  const int b = row / T, t = row - b * T;
  const int len = min(seq_len[b], T);
  const bool live = t < len && status[b] == 0;
  float g0 = 0.f, g1 = 0.f;                     // classes lane and lane+32 (C <= 64)
  if (live) {
    const int cw = C + 1;                       // odd row pitch for 29 classes: lanes land in different banks
    float* mine = bins_all + ((size_t)warp * 32 + lane) * cw;
    for (int c = 0; c < C; ++c) mine[c] = 0.f;
    const int l0 = label_offsets[b];
    const int S = 2 * (label_offsets[b + 1] - l0) + 1;
    const double lz2 = logp[b] * 1.4426950408889634;                      // log2 p(z|x)
    const double lzi = floor(lz2);
    const int li = (int)lzi;
    const float lf = (float)(lz2 - lzi);
    const int2* a = alpha + ((int64_t)b * T + t) * s_pad;
    const int2* be = beta + ((int64_t)b * T + t) * s_pad;
    float blank_sum = 0.f;
    for (int s = lane; s < S; s += 32) {
      const SF av = sf_load(a + s), bv = sf_load(be + s);
      const int de = max(av.e + bv.e - li, -200);                          // a ZERO factor (exponent -2^28) -> 0
      const float e = fast_exp2((float)de - lf) * (av.p * bv.p);
      if (s & 1) mine[labels[l0 + (s >> 1)]] += e;
      else blank_sum += e;
    }
    blank_sum = warp_sum(blank_sum);
    __syncwarp();
    const float* pr = probs + (int64_t)row * C;
    const float* wbins = bins_all + (size_t)warp * 32 * cw;
    if (lane < C) {
      float occ = lane == blank ? blank_sum : 0.f;
      for (int l = 0; l < 32; ++l) occ += wbins[((l + lane) & 31) * cw + lane];
      g0 = grad_scale * (pr[lane] - occ);
    }
    if (lane + 32 < C) {
      float occ = lane + 32 == blank ? blank_sum : 0.f;
      for (int l = 0; l < 32; ++l) occ += wbins[((l + lane) & 31) * cw + lane + 32];
      g1 = grad_scale * (pr[lane + 32] - occ);
    }
  }
  if (grad) {
    float* dst = grad + (int64_t)t * stride_t + (int64_t)b * stride_b;
    if (lane < C) dst[lane] = g0;
    if (lane + 32 < C) dst[lane + 32] = g1;
  }
  if (planes) {
    // planes[p][b][t][c_pad]: value = sum_p plane_p; columns >= C are zero
    const int64_t plane_stride = (int64_t)B * T * c_pad;
    for (int c = lane; c < c_pad; c += 32) {
      const float v = (c >= C) ? 0.f : (c < 32 ? g0 : g1);
      float rem = v;
      for (int p = 0; p < n_planes; ++p) {
        const __nv_bfloat16 h = __float2bfloat16_rn(rem);
        planes[p * plane_stride + ((int64_t)b * T + t) * c_pad + c] = h;
        rem -= __bfloat162float(h);
      }
    }
  }
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

ST_API int st_ctc_validate_labels_host(const int32_t* labels_host, const int32_t* label_offsets_host,
                                       const int32_t* seq_len_host, int B, int T, int blank) {
  ST_CHECK_ARG(label_offsets_host && seq_len_host, "st_ctc_validate_labels_host: null pointer");
  for (int b = 0; b < B; ++b) {
    const int l0 = label_offsets_host[b], l1 = label_offsets_host[b + 1];
    ST_CHECK_ARG(l1 >= l0, "st_ctc_validate_labels_host: label_offsets not monotone at batch %d", b);
    if (seq_len_host[b] > T || seq_len_host[b] < 0) {
      st_set_error("sequence_length(%d) = %d outside [0, max_time=%d]", b, seq_len_host[b], T);
      return ST_ERR_CTC_LABELS;
    }
    int repeats = 0;
    for (int i = l0; i < l1; ++i) {
      if (labels_host[i] < 0 || labels_host[i] >= blank) {
        st_set_error("label id %d outside [0, num_classes-1=%d) in batch %d", labels_host[i], blank, b);
        return ST_ERR_CTC_LABELS;
      }
      if (i > l0 && labels_host[i] == labels_host[i - 1]) repeats++;
    }
    if ((l1 - l0) + repeats > seq_len_host[b]) {
      st_set_error("Not enough time for target transition sequence (required: %d, available: %d) in batch %d",
                   (l1 - l0) + repeats, seq_len_host[b], b);
      return ST_ERR_CTC_LABELS;
    }
  }
  return ST_OK;
}

static size_t ctc_offsets(int T, int B, int C, int max_label_len, size_t* off_alpha, size_t* off_beta,
                          size_t* off_logp, int* s_pad_out) {
  const int s_pad = round_up(2 * max_label_len + 1, 4);
  size_t off = 0;
  off += (size_t)B * T * C * sizeof(float);
  off = (off + 255) / 256 * 256;
  *off_alpha = off;
  off += (size_t)B * T * s_pad * sizeof(double);
  *off_beta = off;
  off += (size_t)B * T * s_pad * sizeof(double);
  *off_logp = off;
  off += (size_t)B * sizeof(double);
  *s_pad_out = s_pad;
  return (off + 255) / 256 * 256;
}

ST_API size_t st_ctc_workspace_bytes(int T, int B, int C, int max_label_len) {
  size_t a, b, l;
  int sp;
  return ctc_offsets(T, B, C, max_label_len, &a, &b, &l, &sp);
}

ST_API int st_ctc_loss(const float* logits, int64_t stride_t, int64_t stride_b, int T, int B, int C,
                       const int32_t* labels, const int32_t* label_offsets, int max_label_len,
                       const int32_t* seq_len, int blank, float* loss, float* grad, float grad_scale,
                       void* grad_planes, int n_planes, int c_pad, int32_t* status, void* workspace,
                       size_t workspace_bytes, st_stream_t stream) {
  ST_CHECK_ARG(logits && label_offsets && seq_len && loss && status && workspace, "st_ctc_loss: null pointer");
  ST_CHECK_ARG(T > 0 && B > 0 && C > 1 && C <= 64, "st_ctc_loss: need T,B > 0 and 1 < C <= 64 (C=%d)", C);
  ST_CHECK_ARG(blank >= 0 && blank < C, "st_ctc_loss: blank outside [0,C)");
  ST_CHECK_ARG(!grad_planes || (n_planes >= 1 && n_planes <= 3 && c_pad >= C && c_pad <= 64),
               "st_ctc_loss: bad plane arguments");
  ST_CHECK_ARG(max_label_len >= 0 && max_label_len <= 1023, "st_ctc_loss: max_label_len %d outside [0,1023]",
               max_label_len);
  size_t off_a, off_b, off_l;
  int s_pad = 0;
  const size_t need = ctc_offsets(T, B, C, max_label_len, &off_a, &off_b, &off_l, &s_pad);
  ST_CHECK_ARG(workspace_bytes >= need, "st_ctc_loss: workspace %zu < required %zu bytes", workspace_bytes, need);
  char* ws = static_cast<char*>(workspace);
  float* probs = reinterpret_cast<float*>(ws);      // class probabilities [B][T][C]
  int2* alpha = reinterpret_cast<int2*>(ws + off_a);          // scaled-float lattices, 8 bytes per entry
  int2* beta = reinterpret_cast<int2*>(ws + off_b);
  double* logp = reinterpret_cast<double*>(ws + off_l);
  cudaStream_t s = st_cu(stream);

  const int rows = B * T;
  ctc_softmax_kernel<<<(rows + 7) / 8, 256, 0, s>>>(logits, stride_t, stride_b, T, B, C, seq_len, probs);
  ST_CUDA_LAUNCH_CHECK("ctc_softmax_kernel");

  const bool want_grad = grad || grad_planes;
  const size_t smem_base = (((size_t)2 * (s_pad + 4) * sizeof(double) + (size_t)s_pad * sizeof(int) + (size_t)(s_pad + 2) + 15) &
                            ~(size_t)15);
  const size_t smem_probs = (size_t)T * C * sizeof(float);
  const bool probs_in_smem = smem_base + smem_probs + 64 <= 200 * 1024;
  const size_t smem = smem_base + (probs_in_smem ? smem_probs : 0) + 64;
  const int S_max = 2 * max_label_len + 1;
  const int threads = S_max >= 1024 ? 1024 : (S_max + 31) / 32 * 32;
  const int ns = S_max <= threads ? 1 : (S_max <= 2 * threads ? 2 : (S_max <= 4 * threads ? 4 : 8));
  // loss only: the alpha recursion alone gives p(z|x); no beta CTAs, no lattice stores
#define ST_CTC_LAUNCH2(NS_, SM_, ST_)                                                                                \
  do {                                                                                                               \
    if (smem > 48 * 1024)                                                                                            \
      ST_CUDA_CALL(cudaFuncSetAttribute(ctc_alpha_beta_kernel<NS_, SM_, ST_>,                                        \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
    ctc_alpha_beta_kernel<NS_, SM_, ST_><<<dim3(B, ST_ ? 2 : 1), threads, smem, s>>>(                                \
        probs, T, C, labels, label_offsets, seq_len, blank, s_pad, alpha, beta, logp, loss, status);                  \
  } while (0)
#define ST_CTC_LAUNCH(NS_, SM_)                                                                                      \
  do {                                                                                                               \
    if (want_grad) ST_CTC_LAUNCH2(NS_, SM_, true); else ST_CTC_LAUNCH2(NS_, SM_, false);                             \
  } while (0)
  if (probs_in_smem) {
    if (ns == 1) ST_CTC_LAUNCH(1, true); else if (ns == 2) ST_CTC_LAUNCH(2, true);
    else if (ns == 4) ST_CTC_LAUNCH(4, true); else ST_CTC_LAUNCH(8, true);
  } else {
    if (ns == 1) ST_CTC_LAUNCH(1, false); else if (ns == 2) ST_CTC_LAUNCH(2, false);
    else if (ns == 4) ST_CTC_LAUNCH(4, false); else ST_CTC_LAUNCH(8, false);
  }
#undef ST_CTC_LAUNCH
#undef ST_CTC_LAUNCH2
  ST_CUDA_LAUNCH_CHECK("ctc_alpha_beta_kernel");
  if (grad || grad_planes) {
    const size_t grad_smem = (size_t)8 * 32 * (C + 1) * sizeof(float);
    if (grad_smem > 48 * 1024)
      ST_CUDA_CALL(cudaFuncSetAttribute(ctc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grad_smem));
    ctc_grad_kernel<<<(rows + 7) / 8, 256, grad_smem, s>>>(probs, alpha, beta, logp, labels, label_offsets, seq_len, status,
                                                   T, B, C, blank, s_pad, grad_scale, grad, stride_t, stride_b,
                                                   static_cast<__nv_bfloat16*>(grad_planes), n_planes, c_pad);
    ST_CUDA_LAUNCH_CHECK("ctc_grad_kernel");
  }
  return ST_OK;
}
