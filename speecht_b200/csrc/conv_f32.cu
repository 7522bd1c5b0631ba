// a6 -- exact-fp32 conv1d 'SAME' + bias + ReLU and its gradients on the CUDA cores (FFMA, fp32 accumulate).
// Replaces tf.nn.conv1d / tf.nn.bias_add / tf.nn.relu (reference speech_model.py:155,173,177) and the gradients
// TF autodiff derives for them (speech_model.py:78).
//
// This is the strict-precision path: one implicit-GEMM kernel (128x128x16 tiles, 8x8 register micro-tile,
// double-buffered shared memory) instantiated three times with different operand "views":
//   forward : Y[(b,t), co]   = sum_{(k,ci)} X[b, t*s+k-padL, ci] * W[k,ci,co]
//   bwd_data: dX[(b,ti), ci] = sum_{(k,co)} dY[b, (ti+padL-k)/s, co] * W[k,ci,co]
//   bwd_filt: dW[(k,ci), co] = sum_{(b,t)}  X[b, t*s+k-padL, ci] * dY[b,t,co]        (split-K + atomics)
// Out-of-range taps read as zero = TF 'SAME' zero padding (left = pad_total/2).  The tcgen05 path in conv_tc.cu
// is the fast one; this one bounds it from the accuracy side and serves shapes the tensor path does not tile.
#include "st_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

struct ConvGeom {
  int B, Ti, To, Cin, Cout, K, stride, padL;
};

// ---- operand views.  mn-state is fixed per thread for the whole K loop, kk-state changes every BK step. ----
struct FwdA {            // A(m=(b,t), kk=(k,ci)) = X[b, t*s+k-padL, ci]
  const float* x; ConvGeom g; int M, Kd;
  struct Mn { int b, t0; bool ok; };
  struct Kk { int k, ci; bool ok; };
  __device__ void mn(int m, Mn& s) const { s.ok = m < M; int b = m / g.To; s.b = b; s.t0 = (m - b * g.To) * g.stride - g.padL; }
  __device__ void kk(int q, Kk& s) const { s.ok = q < Kd; int k = q / g.Cin; s.k = k; s.ci = q - k * g.Cin; }
  __device__ float load(const Mn& a, const Kk& c) const {
    const int ti = a.t0 + c.k;
    if (!(a.ok && c.ok) || ti < 0 || ti >= g.Ti) return 0.f;
    return __ldg(x + ((int64_t)a.b * g.Ti + ti) * g.Cin + c.ci);
  }
};
struct FwdB {            // B(kk=(k,ci), n=co) = W[kk*Cout + co]
  const float* w; int N, Kd;
  struct Mn { int n; bool ok; };
  struct Kk { int q; bool ok; };
  __device__ void mn(int n, Mn& s) const { s.ok = n < N; s.n = n; }
  __device__ void kk(int q, Kk& s) const { s.ok = q < Kd; s.q = q; }
  __device__ float load(const Mn& a, const Kk& c) const {
    return (a.ok && c.ok) ? __ldg(w + (int64_t)c.q * N + a.n) : 0.f;
  }
};
struct DgradA {          // A(m=(b,ti), kk=(k,co)) = dY[b, (ti+padL-k)/s, co] * (y_act > 0)
  const float* dy; const float* act; ConvGeom g; int M, Kd;
  struct Mn { int b, t0; bool ok; };
  struct Kk { int k, co; bool ok; };
  __device__ void mn(int m, Mn& s) const { s.ok = m < M; int b = m / g.Ti; s.b = b; s.t0 = (m - b * g.Ti) + g.padL; }
  __device__ void kk(int q, Kk& s) const { s.ok = q < Kd; int k = q / g.Cout; s.k = k; s.co = q - k * g.Cout; }
  __device__ float load(const Mn& a, const Kk& c) const {
    int num = a.t0 - c.k;
    if (!(a.ok && c.ok) || num < 0) return 0.f;
    int to = num;
    if (g.stride > 1) { to = num / g.stride; if (to * g.stride != num) return 0.f; }
    if (to >= g.To) return 0.f;
    const int64_t idx = ((int64_t)a.b * g.To + to) * g.Cout + c.co;
    float v = __ldg(dy + idx);
    if (act && !(__ldg(act + idx) > 0.f)) v = 0.f;
    return v;
  }
};
struct DgradB {          // B(kk=(k,co), n=ci) = W[(k*Cin+ci)*Cout + co]
  const float* w; ConvGeom g; int N, Kd;
  struct Mn { int n; bool ok; };
  struct Kk { int k, co; bool ok; };
  __device__ void mn(int n, Mn& s) const { s.ok = n < N; s.n = n; }
  __device__ void kk(int q, Kk& s) const { s.ok = q < Kd; int k = q / g.Cout; s.k = k; s.co = q - k * g.Cout; }
  __device__ float load(const Mn& a, const Kk& c) const {
    return (a.ok && c.ok) ? __ldg(w + ((int64_t)c.k * g.Cin + a.n) * g.Cout + c.co) : 0.f;
  }
};
struct WgradA {          // A(m=(k,ci), kk=(b,t)) = X[b, t*s+k-padL, ci]
  const float* x; ConvGeom g; int M, Kd;
  struct Mn { int k, ci; bool ok; };
  struct Kk { int b, t0; bool ok; };
  __device__ void mn(int m, Mn& s) const { s.ok = m < M; int k = m / g.Cin; s.k = k; s.ci = m - k * g.Cin; }
  __device__ void kk(int q, Kk& s) const { s.ok = q < Kd; int b = q / g.To; s.b = b; s.t0 = (q - b * g.To) * g.stride - g.padL; }
  __device__ float load(const Mn& a, const Kk& c) const {
    const int ti = c.t0 + a.k;
    if (!(a.ok && c.ok) || ti < 0 || ti >= g.Ti) return 0.f;
    return __ldg(x + ((int64_t)c.b * g.Ti + ti) * g.Cin + a.ci);
  }
};
struct WgradB {          // B(kk=(b,t), n=co) = dY[kk*Cout + co] * (y_act > 0)
  const float* dy; const float* act; int N, Kd;
  struct Mn { int n; bool ok; };
  struct Kk { int q; bool ok; };
  __device__ void mn(int n, Mn& s) const { s.ok = n < N; s.n = n; }
  __device__ void kk(int q, Kk& s) const { s.ok = q < Kd; s.q = q; }
  __device__ float load(const Mn& a, const Kk& c) const {
    if (!(a.ok && c.ok)) return 0.f;
    const int64_t idx = (int64_t)c.q * N + a.n;
    float v = __ldg(dy + idx);
    if (act && !(__ldg(act + idx) > 0.f)) v = 0.f;
    return v;
  }
};

// Thread -> element mapping of one BMxBK (or BNxBK) operand tile, 8 elements per thread.
//   KC (k-contiguous in memory):  kk = tid%16, mn = tid/16 + 16*i
//   MC (mn-contiguous in memory): mn = tid%128, kk = tid/128 + 2*i
template <class Op, bool KC>
struct TileLoader {
  typename Op::Mn mn_state[KC ? 8 : 1];
  __device__ void init(const Op& op, int mn0, int tid) {
    if (KC) {
#pragma unroll
      for (int i = 0; i < 8; ++i) op.mn(mn0 + (tid >> 4) + 16 * i, mn_state[i]);
    } else {
      op.mn(mn0 + (tid & 127), mn_state[0]);
    }
  }
  __device__ void fetch(const Op& op, int k0, int tid, float (&r)[8]) const {
    if (KC) {
      typename Op::Kk ks;
      op.kk(k0 + (tid & 15), ks);
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = op.load(mn_state[i], ks);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        typename Op::Kk ks;
        op.kk(k0 + (tid >> 7) + 2 * i, ks);
        r[i] = op.load(mn_state[0], ks);
      }
    }
  }
  __device__ static void stash(float (*tile)[BM + 4], int tid, const float (&r)[8]) {
    if (KC) {
#pragma unroll
      for (int i = 0; i < 8; ++i) tile[tid & 15][(tid >> 4) + 16 * i] = r[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) tile[(tid >> 7) + 2 * i][tid & 127] = r[i];
    }
  }
};

enum EpiKind { EPI_STORE_BIAS_RELU = 0, EPI_STORE = 1, EPI_ATOMIC = 2 };

template <class AOp, class BOp, bool A_KC, bool B_KC, int EPI>
__global__ void __launch_bounds__(NT)
conv_gemm_f32_kernel(AOp aop, BOp bop, float* __restrict__ C, const float* __restrict__ bias, int relu,
                     int M, int N, int Kd, int k_per_split) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(Kd, k_begin + k_per_split);
  if (k_begin >= k_end) return;

  TileLoader<AOp, A_KC> la;
  TileLoader<BOp, B_KC> lb;
  la.init(aop, m0, tid);
  lb.init(bop, n0, tid);

  const int ty = tid >> 4, tx = tid & 15;          // 16x16 threads; rows ty*4+{0..3} and 64+ty*4+{0..3}
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  la.fetch(aop, k_begin, tid, ra);
  lb.fetch(bop, k_begin, tid, rb);
  TileLoader<AOp, A_KC>::stash(As[0], tid, ra);
  TileLoader<BOp, B_KC>::stash(Bs[0], tid, rb);
  __syncthreads();

  int buf = 0;
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    const bool more = k0 + BK < k_end;
    if (more) {
      la.fetch(aop, k0 + BK, tid, ra);
      lb.fetch(bop, k0 + BK, tid, rb);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      TileLoader<AOp, A_KC>::stash(As[buf ^ 1], tid, ra);
      TileLoader<BOp, B_KC>::stash(Bs[buf ^ 1], tid, rb);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float v = acc[i][j];
      if (EPI == EPI_STORE_BIAS_RELU) {
        if (bias) v += __ldg(bias + n);
        if (relu) v = fmaxf(v, 0.f);
        C[(int64_t)m * N + n] = v;
      } else if (EPI == EPI_STORE) {
        C[(int64_t)m * N + n] = v;
      } else {
        atomicAdd(C + (int64_t)m * N + n, v);
      }
    }
  }
}

// db[co] = sum_r dY[r][co] * (act > 0).  grid (ceil(N/32), row_chunks), block (32, 8).
__global__ void bias_grad_f32_kernel(const float* __restrict__ dy, const float* __restrict__ act,
                                     float* __restrict__ db, int64_t rows, int N, int rows_per_block) {
  __shared__ float part[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float acc = 0.f;
  if (n < N) {
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      float v = dy[r * N + n];
      if (act && !(act[r * N + n] > 0.f)) v = 0.f;
      acc += v;
    }
  }
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][threadIdx.x];
    atomicAdd(db + n, s);
  }
}

int check_geom(const char* fn, int B, int T, int Cin, int Cout, int K, int stride, ConvGeom* g) {
  ST_CHECK_ARG(B > 0 && T > 0 && Cin > 0 && Cout > 0 && K > 0 && stride > 0, "%s: non-positive dimension", fn);
  g->B = B; g->Ti = T; g->Cin = Cin; g->Cout = Cout; g->K = K; g->stride = stride;
  g->To = (T + stride - 1) / stride;
  int pad_total = (g->To - 1) * stride + K - T;
  if (pad_total < 0) pad_total = 0;
  g->padL = pad_total / 2;
  ST_CHECK_ARG((int64_t)B * g->To < (1ll << 31) && (int64_t)K * Cin < (1ll << 31) && (int64_t)K * Cout < (1ll << 31),
               "%s: GEMM dimension overflows int32", fn);
  return ST_OK;
}

}  // namespace

ST_API int st_conv1d_fwd_f32(const float* x, const float* w, const float* bias, float* y, int B, int T, int Cin,
                             int Cout, int K, int stride, int relu, st_stream_t stream) {
  ST_CHECK_ARG(x && w && y, "st_conv1d_fwd_f32: null pointer");
  ConvGeom g;
  int rc = check_geom("st_conv1d_fwd_f32", B, T, Cin, Cout, K, stride, &g);
  if (rc) return rc;
  const int M = B * g.To, N = Cout, Kd = K * Cin;
  FwdA a{x, g, M, Kd};
  FwdB b{w, N, Kd};
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, 1);
  conv_gemm_f32_kernel<FwdA, FwdB, true, false, EPI_STORE_BIAS_RELU>
      <<<grid, NT, 0, st_cu(stream)>>>(a, b, y, bias, relu, M, N, Kd, Kd);
  ST_CUDA_LAUNCH_CHECK("conv_gemm_f32_kernel<fwd>");
  return ST_OK;
}

ST_API int st_conv1d_bwd_data_f32(const float* dy, const float* y_act, const float* w, float* dx, int B, int T,
                                  int Cin, int Cout, int K, int stride, st_stream_t stream) {
  ST_CHECK_ARG(dy && w && dx, "st_conv1d_bwd_data_f32: null pointer");
  ConvGeom g;
  int rc = check_geom("st_conv1d_bwd_data_f32", B, T, Cin, Cout, K, stride, &g);
  if (rc) return rc;
  const int M = B * g.Ti, N = Cin, Kd = K * Cout;
  DgradA a{dy, y_act, g, M, Kd};
  DgradB b{w, g, N, Kd};
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, 1);
  conv_gemm_f32_kernel<DgradA, DgradB, true, true, EPI_STORE>
      <<<grid, NT, 0, st_cu(stream)>>>(a, b, dx, nullptr, 0, M, N, Kd, Kd);
  ST_CUDA_LAUNCH_CHECK("conv_gemm_f32_kernel<bwd_data>");
  return ST_OK;
}

ST_API int st_conv1d_bwd_filter_f32(const float* x, const float* dy, const float* y_act, float* dw, float* db,
                                    int B, int T, int Cin, int Cout, int K, int stride, st_stream_t stream) {
  ST_CHECK_ARG(x && dy && dw, "st_conv1d_bwd_filter_f32: null pointer");
  ConvGeom g;
  int rc = check_geom("st_conv1d_bwd_filter_f32", B, T, Cin, Cout, K, stride, &g);
  if (rc) return rc;
  const int M = K * Cin, N = Cout, Kd = B * g.To;
  cudaStream_t s = st_cu(stream);
  ST_CUDA_CALL(cudaMemsetAsync(dw, 0, (size_t)M * N * sizeof(float), s));
  WgradA a{x, g, M, Kd};
  WgradB b{dy, y_act, N, Kd};
  const int tiles = ((N + BN - 1) / BN) * ((M + BM - 1) / BM);
  int split = (4 * st_num_sms() + tiles - 1) / tiles;            // aim for >= 4 CTAs per SM worth of work
  const int max_split = (Kd + 4 * BK - 1) / (4 * BK);
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  int k_per_split = ((Kd + split - 1) / split + BK - 1) / BK * BK;
  split = (Kd + k_per_split - 1) / k_per_split;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, split);
  conv_gemm_f32_kernel<WgradA, WgradB, false, false, EPI_ATOMIC>
      <<<grid, NT, 0, s>>>(a, b, dw, nullptr, 0, M, N, Kd, k_per_split);
  ST_CUDA_LAUNCH_CHECK("conv_gemm_f32_kernel<bwd_filter>");
  if (db) {
    ST_CUDA_CALL(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), s));
    const int64_t rows = Kd;
    int chunks = (int)((rows + 255) / 256);
    if (chunks > 4 * st_num_sms()) chunks = 4 * st_num_sms();
    const int rows_per_block = (int)((rows + chunks - 1) / chunks);
    dim3 bgrid((N + 31) / 32, chunks);
    bias_grad_f32_kernel<<<bgrid, dim3(32, 8), 0, s>>>(dy, y_act, db, rows, N, rows_per_block);
    ST_CUDA_LAUNCH_CHECK("bias_grad_f32_kernel");
  }
  return ST_OK;
}
