// a10/a11 -- tf.clip_by_global_norm + tf.train.AdamOptimizer(epsilon=1e-3).apply_gradients
// (reference speech_model.py:77-82).  All 22 parameter tensors live in one flat fp32 buffer, so both ops are a
// single streaming pass each: HBM-bound, algorithmic bytes = 4n (norm) and 28n (read g,p,m,v; write p,m,v).
#include "st_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ accum) {
  __shared__ double part[8];
  const int64_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(g4 + i);
    acc += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[(n4 << 2) + threadIdx.x];
    acc += (double)v * v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += part[w];
    atomicAdd(accum, s);
  }
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float gscale, float lr_t, float b1,
                                         float b2, float eps) {
  g *= gscale;
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * (g * g);
  p -= lr_t * m / (sqrtf(v) + eps);
}

__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, float lr_t, float b1, float b2, float eps, float max_norm,
                 const double* __restrict__ normsq, float prescale) {
  // tf.clip_by_global_norm: scale = clip * min(1/norm, 1/clip)
  float gscale = prescale;
  if (normsq) {
    const float norm = (float)(sqrt(*normsq) * (double)prescale);
    gscale = prescale * max_norm * fminf(1.f / norm, 1.f / max_norm);
  }
  const int64_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = __ldg(g4 + i);
    adam_one(pp.x, gg.x, mm.x, vv.x, gscale, lr_t, b1, b2, eps);
    adam_one(pp.y, gg.y, mm.y, vv.y, gscale, lr_t, b1, b2, eps);
    adam_one(pp.z, gg.z, mm.z, vv.z, gscale, lr_t, b1, b2, eps);
    adam_one(pp.w, gg.w, mm.w, vv.w, gscale, lr_t, b1, b2, eps);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    adam_one(p[i], g[i], m[i], v[i], gscale, lr_t, b1, b2, eps);
  }
}

}  // namespace

ST_API int st_sumsq(const float* g, int64_t n, double* accum, int zero_first, st_stream_t stream) {
  ST_CHECK_ARG(g && accum && n >= 0, "st_sumsq: bad argument");
  ST_CHECK_ARG((reinterpret_cast<uintptr_t>(g) & 15) == 0, "st_sumsq: buffer must be 16-byte aligned");
  // zero_first: *accum = 0 on the stream before the blocks add into it (a memset node, not a separate kernel)
  if (zero_first) ST_CUDA_CALL(cudaMemsetAsync(accum, 0, sizeof(double), st_cu(stream)));
  if (n == 0) return ST_OK;
  int blocks = (int)((n / 4 + 255) / 256);
  const int cap = 8 * st_num_sms();
  blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
  sumsq_kernel<<<blocks, 256, 0, st_cu(stream)>>>(g, n, accum);
  ST_CUDA_LAUNCH_CHECK("sumsq_kernel");
  return ST_OK;
}

ST_API int st_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                        float eps, int64_t step, float max_norm, const double* normsq, float grad_prescale,
                        st_stream_t stream) {
  ST_CHECK_ARG(p && g && m && v && n >= 0 && step >= 1, "st_clip_adam: bad argument (step counts from 1)");
  ST_CHECK_ARG(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v)) & 15) == 0, "st_clip_adam: buffers must be 16-byte aligned");
  ST_CHECK_ARG(!normsq || max_norm > 0.f, "st_clip_adam: max_norm must be positive");
  if (n == 0) return ST_OK;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
  int blocks = (int)((n / 4 + 255) / 256);
  const int cap = 8 * st_num_sms();
  blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
  clip_adam_kernel<<<blocks, 256, 0, st_cu(stream)>>>(p, g, m, v, n, (float)lr_t, beta1, beta2, eps, max_norm, normsq,
                                                      grad_prescale);
  ST_CUDA_LAUNCH_CHECK("clip_adam_kernel");
  return ST_OK;
}
