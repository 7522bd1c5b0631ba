// a12 -- greedy CTC decode (replaces tf.nn.ctc_greedy_decoder, reference speech_model.py:113-115).
//
// One CTA per utterance.  Phase 1: one warp per frame, 29 classes on 29 lanes, shuffle arg-max with
// "first maximum wins" tie-break (lowest class index), written to a shared byte array.  Phase 2: block-wide
// stream compaction of the keep-flags (not blank, not a repeat of the previous frame's arg-max) with a
// ballot/popc scan, so label order is preserved.  Integer output -> bit-exact by construction.
// HBM-bound in principle (reads T*B*C*4 bytes once); at these sizes (<= 45 MB) it is launch/latency bound.
#include "st_common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
greedy_decode_kernel(const float* __restrict__ logits, int64_t stride_t, int64_t stride_b, int T, int C,
                     const int32_t* __restrict__ seq_len, int blank, int merge_repeated,
                     int32_t* __restrict__ out_values, int32_t* __restrict__ out_counts,
                     float* __restrict__ neg_sum_logits) {
  extern __shared__ unsigned char s_arg[];          // [len] arg-max class per frame
  __shared__ float s_part[kThreads / 32];
  __shared__ int s_warp_count[kThreads / 32];
  __shared__ int s_base;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int len = seq_len[b];
  len = len < 0 ? 0 : (len > T ? T : len);
  const float* base = logits + (int64_t)b * stride_b;

  float neg_sum = 0.f;
  for (int t = warp; t < len; t += kThreads / 32) {
    const float* row = base + (int64_t)t * stride_t;
    float best = -INFINITY;
    int best_c = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {            // C = 29 -> one pass
      float v = row[c];
      if (v > best || best_c == 0x7fffffff) { best = v; best_c = c; }   // strictly greater keeps the first max
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
      if (ov > best || (ov == best && oc < best_c)) { best = ov; best_c = oc; }
    }
    if (lane == 0) {
      s_arg[t] = (unsigned char)best_c;
      neg_sum -= best;
    }
  }
  if (lane == 0) s_part[warp] = neg_sum;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) acc += s_part[w];
    neg_sum_logits[b] = acc;
  }

  int32_t* out = out_values + (int64_t)b * T;
  for (int t0 = 0; t0 < len; t0 += kThreads) {
    const int t = t0 + threadIdx.x;
    int keep = 0, cls = 0;
    if (t < len) {
      cls = s_arg[t];
      const int prev = t > 0 ? (int)s_arg[t - 1] : -1;
      keep = (cls != blank) && !(merge_repeated && cls == prev);
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp_count[warp] = __popc(m);
    __syncthreads();
    int offset = s_base;
    for (int w = 0; w < warp; ++w) offset += s_warp_count[w];
    if (keep) out[offset + __popc(m & ((1u << lane) - 1u))] = cls;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < kThreads / 32; ++w) tot += s_warp_count[w];
      s_base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out_counts[b] = s_base;
}

}  // namespace

ST_API int st_ctc_greedy_decode(const float* logits, int64_t stride_t, int64_t stride_b, int T, int B, int C,
                                const int32_t* seq_len, int blank, int merge_repeated, int32_t* out_values,
                                int32_t* out_counts, float* neg_sum_logits, st_stream_t stream) {
  ST_CHECK_ARG(logits && seq_len && out_values && out_counts && neg_sum_logits, "st_ctc_greedy_decode: null pointer");
  ST_CHECK_ARG(T >= 0 && B >= 0 && C > 0 && C <= 255, "st_ctc_greedy_decode: need 0 < C <= 255 (got %d)", C);
  ST_CHECK_ARG(blank >= 0 && blank < C, "st_ctc_greedy_decode: blank %d outside [0,%d)", blank, C);
  ST_CHECK_ARG((size_t)T <= 200 * 1024, "st_ctc_greedy_decode: T=%d exceeds the shared-memory frame buffer", T);
  if (B == 0) return ST_OK;
  size_t smem = (size_t)(T > 0 ? T : 1);
  if (smem > 48 * 1024) {
    ST_CUDA_CALL(cudaFuncSetAttribute(greedy_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  greedy_decode_kernel<<<B, kThreads, smem, st_cu(stream)>>>(logits, stride_t, stride_b, T, C, seq_len, blank,
                                                             merge_repeated, out_values, out_counts, neg_sum_logits);
  ST_CUDA_LAUNCH_CHECK("greedy_decode_kernel");
  return ST_OK;
}
