// Step plan of the tensor-core path: what the reference runs as ONE sess.run over its TF1 graph
// (speech_model.py:235; network speech_model.py:275-295) is, for a fixed (batch, time) shape, a fixed sequence of
// kernel launches over buffers carved from one caller-owned arena.  The plan owns no device memory: it records
// offsets, TMA tensor maps and launch parameters, and enqueues
//   forward : split input -> 11 x tc_conv_kernel (bias+ReLU fused; the last layer writes fp32 logits)
//   backward: per layer bias-grad, tc_wgrad_kernel (filter grad into the flat fp32 gradient buffer) and
//             tc_conv_kernel in data-gradient mode with the ReLU mask fused into its epilogue.
// CTC (ctc.cu) sits between the two and writes d(loss)/d(logits) straight into the plan's bf16 planes.
#include "st_common.cuh"
#include "conv_tc.h"
#include <stdlib.h>
#include <vector>

namespace {

struct Layer {
  int K, stride, Cin, Cout, relu;
  int Ti, To, pad_left;
  int ld_in, ld_out;          // channel stride of the input / output activation planes
  int cin_p;                  // Cin rounded up to 64 (forward filter layout)
  int ld_co;                  // Cout rounded up (backward filter layout)
  size_t off_wfwd, off_wbwd;  // arena offsets (bytes)
  size_t off_out;             // arena offset of this layer's OUTPUT activation planes (layers 0..9)
  size_t off_mask;            // ... and of its ReLU-mask bit words [ceil(ld_out / 32)][B * To] (tc::ConvParams::mask_out)
  int64_t w_off, b_off;       // float offsets into the flat parameter / gradient buffers
  CUtensorMap tm_fwd_a, tm_fwd_b;     // forward
  CUtensorMap tm_dg_a[2], tm_dg_b;    // data gradient (A = dz ping or pong)
  CUtensorMap tm_dg_b_n128;           // layer 10 only: filter rows in boxes of 128 (128-wide data-gradient tiles)
  CUtensorMap tm_wg_x, tm_wg_dz[2];   // filter gradient
  bool pair_fwd, pair_dg;             // forward / data gradient on CTA pairs: their B maps hold half-height boxes
  CUtensorMap tm_fwd_out;             // store map of this layer's output planes (layers 0..9)
  CUtensorMap tm_dg_out[2];           // store map of the dz buffer the data gradient of this layer writes (1..10)
};

int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

struct st_plan {
  int B, T, Tpad, F, C, npl;
  int To;                      // logit frames
  std::vector<Layer> layers;
  size_t off_in, off_logits, off_dlogits, off_dz[2], arena_bytes;
  // gradients wrt the outputs of layers 0..7 each keep their OWN buffer (the 2000-channel ones of layers 8, 9 share the
  // ping / pong pair): the filter gradients of the seven 250-channel layers are deferred to ONE multi-problem launch
  // at the end of the backward pass and need all of them alive (SPEECHT_B200_MERGE_WGRAD=0: one launch per layer)
  size_t off_dzs[8];
  int range_lo;                // lowest layer of the backward range being enqueued (st_plan_backward_range)
  bool merge_wgrad;
  size_t filter_bytes;         // leading arena region holding the packed filters (shape independent)
  size_t dz_elems;             // elements per plane of a dz buffer
  char* arena;
  float* params;
  float* grads;
  int launches;
  bool bound;
  // optional per-launch CUDA-event timing of the tensor-core kernels (bench.py roofline)
  bool timing;
  std::vector<cudaEvent_t> ev_pool;
  struct Rec { int kind, layer; double flops; int e0, e1; };   // kind 0 = conv fwd, 1 = data grad, 2 = filter grad
  std::vector<Rec> recs;
  int ev_used;
  int cur_dz;                  // ping/pong buffer holding the gradient wrt the next layer to process
  bool tma_store;              // one / two planes: the epilogues write bf16 planes with TMA stores (store maps needed)
  bool trim;                   // MMAs over channel / time padding are not issued (SPEECHT_B200_TRIM=0 disables)
  // Layer-10 data gradient (dz9 = dlogits . W10^T, 29-deep contraction, 2000 outputs): an HBM / epilogue kernel, not a
  // GEMM.  128-wide tiles leave room for TWO accumulator stages in the split modes (2 x 2 x 128 TMEM columns), so the
  // epilogue of one tile runs under the loads + MMAs of the next, and their kernel instantiation needs few enough
  // registers for SIXTEEN epilogue warps (ConvCfg::EPW) (SPEECHT_B200_L10_N128=0 restores 256-wide tiles).
  bool l10_n128;
  // Fast-FIR split of the 32-tap layer 8 (65 % of the FLOPs): forward, data gradient and filter gradient each run
  // THREE half-rate 16-tap problems (75 % of the MMAs) in one persistent launch, plus elementwise prepare / combine
  // passes (index algebra and backward formulas checked on the CPU by tools/ffa_study.py).
  // TWO levels (the default): each of the three half-rate problems is split again -- nine quarter-rate 8-tap
  // problems, 56 % of the MMAs (tools/ffa2_study.py).  `ffa` is the level in use: SPEECHT_B200_FFA=1 keeps one level,
  // =0 restores the direct 32-tap kernels; bf16x6 (three planes) always uses the direct kernels.
  int ffa;
  // Forward launches of the two widest layers (the nine fast-FIR leaves of layer 8, layer 9) read their filters
  // MN-major from the BACKWARD layout (tc::ConvParams::b_mn): no forward layout is packed for them, and their packing
  // is a pure streaming pass (tc::launch_pack_bwd): packing 165 -> 81 us per step under ncu.  The MN-major operand costs
  // the forward kernels 2-4 % in bf16x3 and 10 % in plain bf16 (profiles/r02_bmn_ab_session19.txt), so it is the default
  // for two planes only (step -0.5 .. -1 %); SPEECHT_B200_BMN=0 / 1 forces the K-major forward layouts / MN-major.
  bool bmn;
  // Overlap of HBM-bound passes with tensor-core launches that leave the HBM idle (SPEECHT_B200_OVERLAP=0 disables it).
  // On a plan-owned side stream, as BACKGROUND kernels (one block per SM, <= 40 registers, no shared memory: they fit on
  // an SM beside a resident tensor-core CTA and never keep one from starting):
  //   * packing of layers 8-9 (94 % of the filter bytes) underneath the forward pass of layers 0-7,
  //   * zeroing of the flat gradient buffer underneath the forward pass,
  //   * the filter-gradient combine of layer 8 underneath the backward pass of layers 8 (data gradient) .. 1.
  // Every hand-over is an event pair: the side stream starts behind what the main stream has enqueued so far (the
  // previous step's backward and Adam), the main stream waits for the side stream's event before the first consumer.
  bool overlap;
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_pack, ev_zero, ev_dw;
  bool pack_pending, zero_pending, dw_pending;
  int ffa2_Tq, ffa2_Tqi;                         // rows of the leaf products / of the quarter-rate input sequences
  size_t off_ffa2_s[5], off_ffa2_w[9], off_ffa2_wb[9], off_ffa2_p[9], off_ffa2_dxp[9], off_ffa2_c[9];
  bool ffa2_pair_fwd, ffa2_pair_dg;
  CUtensorMap tm_ffa2_a[9], tm_ffa2_b[9], tm_ffa2_dg_a[9], tm_ffa2_dg_b[9], tm_ffa2_wg_x[9], tm_ffa2_wg_dz[9];
  int ffa_Tx, ffa_Tu;                            // rows of the pair-sum planes / of the partial products
  size_t off_ffa_xs, off_ffa_w[3], off_ffa_wb[3], off_ffa_p[3], off_ffa_dxp[3], off_ffa_cs;
  bool ffa_pair_fwd, ffa_pair_dg;                // the multi-problem launches run on CTA pairs
  CUtensorMap tm_ffa_a[3], tm_ffa_b[3];          // forward: A = odd / even row views of x, pair sums; B = w0, w1, ws
  CUtensorMap tm_ffa_dg_a[3], tm_ffa_dg_b[3];    // data gradient: A = dA00 / dA11 / dS planes; B = backward layouts
  CUtensorMap tm_ffa_wg_x[3], tm_ffa_wg_dz[3];   // filter gradient: the same operands in 64-row boxes
};

namespace {

int timed_begin(st_plan* p, cudaStream_t s) {
  if (!p->timing) return -1;
  while ((int)p->ev_pool.size() < p->ev_used + 2) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return -1;
    p->ev_pool.push_back(e);
  }
  const int i = p->ev_used;
  p->ev_used += 2;
  cudaEventRecord(p->ev_pool[i], s);
  return i;
}

void timed_end(st_plan* p, int i, int kind, int layer, double flops, cudaStream_t s) {
  if (i < 0) return;
  cudaEventRecord(p->ev_pool[i + 1], s);
  p->recs.push_back({kind, layer, flops, i, i + 1});
}

// N tile of the tensor-core kernels: three planes per operand only fit two pipeline stages with 128-wide tiles
// (bf16x3 keeps 256-wide tiles with single-buffered main+side accumulators: 128-wide, double-buffered tiles were
// measured 17 % slower over the step, profiles/r01_tile_width_ab.txt)
int wide_n(const st_plan* p) { return p->npl == 3 ? 128 : 256; }

__nv_bfloat16* bf(st_plan* p, size_t off) { return reinterpret_cast<__nv_bfloat16*>(p->arena + off); }

const __nv_bfloat16* act_in(st_plan* p, int l) { return l == 0 ? bf(p, p->off_in) : bf(p, p->layers[l - 1].off_out); }

uint32_t* mask_of(st_plan* p, int l) { return reinterpret_cast<uint32_t*>(p->arena + p->layers[l].off_mask); }

// Side stream of the plan (created on first use) positioned behind everything enqueued on `s` so far.
int side_fork(st_plan* p, cudaStream_t s) {
  if (!p->side) {
    ST_CUDA_CALL(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&p->ev_fork, &p->ev_pack, &p->ev_zero, &p->ev_dw})
      ST_CUDA_CALL(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  ST_CUDA_CALL(cudaEventRecord(p->ev_fork, s));
  ST_CUDA_CALL(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
  return ST_OK;
}
// The main stream waits for a side-stream hand-over (if one is outstanding).
int side_join(st_plan* p, bool* pending, cudaEvent_t ev, cudaStream_t s) {
  if (*pending) {
    ST_CUDA_CALL(cudaStreamWaitEvent(s, ev, 0));
    *pending = false;
  }
  return ST_OK;
}

}  // namespace

ST_API int st_plan_create(st_plan** out, int B, int T, int input_size, int num_classes, int n_planes) {
  ST_CHECK_ARG(out && B > 0 && T > 1, "st_plan_create: bad shape");
  ST_CHECK_ARG(n_planes >= 1 && n_planes <= 3, "st_plan_create: n_planes must be 1 (bf16), 2 (bf16x3) or 3 (bf16x6)");
  ST_CHECK_ARG(input_size % 64 == 0, "st_plan_create: input_size must be a multiple of 64 (got %d)", input_size);
  ST_CHECK_ARG(num_classes >= 2 && num_classes <= 32, "st_plan_create: num_classes must be in [2,32]");
  st_plan* p = new st_plan();
  p->B = B; p->T = T; p->Tpad = round_up(T, 2); p->F = input_size; p->C = num_classes; p->npl = n_planes;
  p->arena = nullptr; p->params = nullptr; p->grads = nullptr; p->launches = 0; p->bound = false; p->cur_dz = 0; p->timing = false; p->ev_used = 0;
  {
    p->tma_store = n_planes <= 2;
    const char* e = getenv("SPEECHT_B200_TRIM");
    p->trim = !(e && e[0] == '0');
    e = getenv("SPEECHT_B200_L10_N128");
    p->l10_n128 = !(e && e[0] == '0') && n_planes <= 2;
    e = getenv("SPEECHT_B200_MERGE_WGRAD");
    p->merge_wgrad = !(e && e[0] == '0') && n_planes <= 2;
    e = getenv("SPEECHT_B200_FFA");
    p->ffa = n_planes > 2 ? 0 : (e && e[0] >= '0' && e[0] <= '2' ? e[0] - '0' : 2);
    e = getenv("SPEECHT_B200_OVERLAP");
    p->overlap = !(e && e[0] == '0') && n_planes <= 2;
    p->side = nullptr;
    p->ev_fork = p->ev_pack = p->ev_zero = p->ev_dw = nullptr;
    p->pack_pending = p->zero_pending = p->dw_pending = false;
    e = getenv("SPEECHT_B200_BMN");
    p->bmn = n_planes <= 2 && ((e && (e[0] == '0' || e[0] == '1')) ? e[0] == '1' : n_planes == 2);
  }
  const int ffa_requested = p->ffa;
  // reference speech_model.py:275-292
  const int table[11][5] = {{48, 2, input_size, 250, 1}, {7, 1, 250, 250, 1}, {7, 1, 250, 250, 1}, {7, 1, 250, 250, 1},
                            {7, 1, 250, 250, 1},         {7, 1, 250, 250, 1}, {7, 1, 250, 250, 1}, {7, 1, 250, 250, 1},
                            {32, 1, 250, 2000, 1},       {1, 1, 2000, 2000, 1}, {1, 1, 2000, num_classes, 0}};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 1023) / 1024 * 1024; return o; };
  // Region 1 -- packed filters.  Their offsets depend on n_planes (and the requested fast-FIR level) only, so every
  // plan of an engine finds them at the same place in the shared arena: a change of batch shape does not invalidate
  // them (an evaluate run over ragged batches packs once, not once per shape).  Region 2 -- everything whose size
  // depends on (B, T) -- follows.
  int t = T, ld_prev = input_size;
  int64_t poff = 0;
  for (int l = 0; l < 11; ++l) {
    Layer L{};
    L.K = table[l][0]; L.stride = table[l][1]; L.Cin = table[l][2]; L.Cout = table[l][3]; L.relu = table[l][4];
    L.Ti = t;
    L.To = (t + L.stride - 1) / L.stride;
    int pad_total = (L.To - 1) * L.stride + L.K - L.Ti;
    if (pad_total < 0) pad_total = 0;
    L.pad_left = pad_total / 2;
    L.ld_in = ld_prev;
    L.ld_out = l == 10 ? 32 : round_up(L.Cout, 8);
    if (L.Cout == 250) L.ld_out = 256;
    L.cin_p = round_up(L.Cin, 64);
    L.ld_co = l == 10 ? 64 : round_up(L.Cout, 8);
    if (L.Cout == 250) L.ld_co = 256;
    L.off_wfwd = take((size_t)n_planes * L.Cout * L.K * L.cin_p * 2);
    L.off_wbwd = l > 0 ? take((size_t)n_planes * L.K * L.Cin * L.ld_co * 2) : 0;
    // flat parameter layout: must match engine.ParamLayout (64-float alignment)
    L.w_off = poff;
    poff = (poff + (int64_t)L.K * L.Cin * L.Cout + 63) / 64 * 64;
    L.b_off = poff;
    poff = (poff + L.Cout + 63) / 64 * 64;
    t = L.To;
    ld_prev = L.ld_out;
    p->layers.push_back(L);
  }
  p->To = t;
  {
    // leaf filters of the fast-FIR levels a plan of this engine may use (small shapes fall back to lower levels)
    const Layer& L8 = p->layers[8];
    if (ffa_requested >= 1)
      for (int i = 0; i < 3; ++i) {
        p->off_ffa_w[i] = take((size_t)n_planes * L8.Cout * (L8.K / 2) * L8.cin_p * 2);
        p->off_ffa_wb[i] = take((size_t)n_planes * (L8.K / 2) * L8.Cin * L8.ld_co * 2);
      }
    if (ffa_requested >= 2)
      for (int i = 0; i < 9; ++i) {
        p->off_ffa2_w[i] = take((size_t)n_planes * L8.Cout * (L8.K / 4) * L8.cin_p * 2);
        p->off_ffa2_wb[i] = take((size_t)n_planes * (L8.K / 4) * L8.Cin * L8.ld_co * 2);
      }
  }
  p->filter_bytes = off;
  // ---- region 2: shape dependent
  p->off_in = take((size_t)n_planes * B * p->Tpad * input_size * 2);
  for (int l = 0; l < 10; ++l) {
    Layer& L = p->layers[l];
    L.off_out = take((size_t)n_planes * B * L.To * L.ld_out * 2);
    L.off_mask = take((size_t)((L.ld_out + 31) / 32) * B * L.To * sizeof(uint32_t));
  }
  p->off_logits = take((size_t)B * p->To * 32 * sizeof(float));
  p->off_dlogits = take((size_t)n_planes * B * p->To * 64 * 2);
  p->dz_elems = (size_t)B * p->To * 2000;
  p->off_dz[0] = take((size_t)n_planes * p->dz_elems * 2);
  p->off_dz[1] = take((size_t)n_planes * p->dz_elems * 2);
  for (int l = 0; l < 8; ++l) {
    const Layer& L = p->layers[l];
    p->off_dzs[l] = take((size_t)n_planes * B * L.To * L.ld_out * 2);
  }
  if (p->To < 8) p->ffa = 0;                  // nothing to gain on a handful of frames (and the odd-row view may be empty)
  if (p->ffa == 2 && p->To < 16) p->ffa = 1;
  if (p->ffa == 2) {
    const Layer& L8 = p->layers[8];
    const int Tu = (L8.To + 1) / 2 + 1;
    p->ffa2_Tq = (Tu + 1) / 2 + 1;
    p->ffa2_Tqi = (L8.Ti + 3) / 4 + 1;
    for (int i = 0; i < 5; ++i) p->off_ffa2_s[i] = take((size_t)n_planes * B * p->ffa2_Tqi * L8.ld_in * 2);
    for (int i = 0; i < 9; ++i) {
      // forward: fp32 leaf product [B][Tq][Cout]; backward: the planes of its gradient live in the same bytes
      const size_t fwd_bytes = (size_t)B * p->ffa2_Tq * L8.Cout * sizeof(float);
      const size_t bwd_bytes = (size_t)n_planes * B * p->ffa2_Tq * L8.ld_out * 2;
      p->off_ffa2_p[i] = take(fwd_bytes > bwd_bytes ? fwd_bytes : bwd_bytes);
      p->off_ffa2_dxp[i] = take((size_t)B * p->ffa2_Tqi * 256 * sizeof(float));
      p->off_ffa2_c[i] = take((size_t)(L8.K / 4) * L8.Cin * L8.Cout * sizeof(float));
    }
  }
  if (p->ffa == 1) {
    const Layer& L8 = p->layers[8];
    p->ffa_Tx = (L8.Ti + 1) / 2;
    p->ffa_Tu = (L8.To + 1) / 2 + 1;
    p->off_ffa_xs = take((size_t)n_planes * B * p->ffa_Tx * L8.ld_in * 2);
    for (int i = 0; i < 3; ++i) {
      // forward: fp32 partial product [B][Tu][Cout]; backward: the planes of dA00 / dA11 / dS [npl][B][Tu][Cout] live
      // in the same bytes (the partial products are dead once the forward combine has run)
      const size_t fwd_bytes = (size_t)B * p->ffa_Tu * L8.Cout * sizeof(float);
      const size_t bwd_bytes = (size_t)n_planes * B * p->ffa_Tu * L8.ld_out * 2;
      p->off_ffa_p[i] = take(fwd_bytes > bwd_bytes ? fwd_bytes : bwd_bytes);
      p->off_ffa_dxp[i] = take((size_t)B * p->ffa_Tx * 256 * sizeof(float));
    }
    p->off_ffa_cs = take((size_t)(L8.K / 2) * L8.Cin * L8.Cout * sizeof(float));
  }
  p->arena_bytes = off;
  *out = p;
  return ST_OK;
}

ST_API int st_plan_destroy(st_plan* p) {
  if (p) {
    for (cudaEvent_t e : p->ev_pool) cudaEventDestroy(e);
    if (p->side) {
      cudaStreamSynchronize(p->side);
      for (cudaEvent_t e : {p->ev_fork, p->ev_pack, p->ev_zero, p->ev_dw})
        if (e) cudaEventDestroy(e);
      cudaStreamDestroy(p->side);
    }
  }
  delete p;
  return ST_OK;
}

ST_API size_t st_plan_arena_bytes(const st_plan* p) { return p ? p->arena_bytes : 0; }
ST_API int64_t st_plan_param_floats(const st_plan* p) {
  if (!p) return 0;
  const Layer& L = p->layers.back();
  return (L.b_off + L.Cout + 63) / 64 * 64;
}
ST_API int st_plan_logit_frames(const st_plan* p) { return p ? p->To : 0; }
ST_API float* st_plan_logits(st_plan* p) { return p && p->arena ? reinterpret_cast<float*>(p->arena + p->off_logits) : nullptr; }
ST_API void* st_plan_dlogits_planes(st_plan* p) { return p && p->arena ? p->arena + p->off_dlogits : nullptr; }
ST_API int st_plan_launches(const st_plan* p) { return p ? p->launches : 0; }
// Which packed filters this plan reads: the fast-FIR level of layer 8 it runs (0 = direct 32-tap kernels).  The packed
// filters sit at shape-independent arena offsets, so plans of one level share them (speecht_b200/tc_plan.py).
ST_API int st_plan_filter_set(const st_plan* p) { return p ? p->ffa : 0; }

namespace {

// Leaves of the two-level fast-FIR split of layer 8 (order XX XY XZ YX YY YZ ZX ZY ZZ; tools/ffa2_study.py): the
// quarter-rate input sequence (a row view x[4r + view_c] of the layer's input, or the summed sequence s[seq]), the
// left padding of its 8-tap correlation and which of the source taps w[4i + c] its filter sums (bit c of tap_mask).
struct Leaf { int view_c, seq, pad, tap_mask; };
const Leaf kLeaves[9] = {{1, -1, 4, 0x1}, {3, -1, 4, 0x4}, {-1, 0, 3, 0x5},
                         {2, -1, 4, 0x2}, {0, -1, 3, 0x8}, {-1, 1, 3, 0xa},
                         {-1, 2, 4, 0x3}, {-1, 3, 3, 0xc}, {-1, 4, 3, 0xf}};

int bind_ffa2(st_plan* p) {
  Layer& L8 = p->layers[8];
  const int B = p->B, npl = p->npl, J = L8.K / 4;
  const __nv_bfloat16* x = act_in(p, 8);
  p->ffa2_pair_fwd = tc::want_pair(B * ((p->ffa2_Tq + tc::kTileM - 1) / tc::kTileM), 1, wide_n(p), npl, true,
                                   J * (L8.cin_p / 64));
  p->ffa2_pair_dg = tc::want_pair(B * ((p->ffa2_Tqi + tc::kTileM - 1) / tc::kTileM), 1, wide_n(p), npl, true,
                                  J * ((L8.Cout + 63) / 64));
  int rc;
  for (int l = 0; l < 9; ++l) {
    const Leaf& lf = kLeaves[l];
    // A operand (forward) / X operand (filter gradient): the leaf's quarter-rate input sequence
    for (int box_t = 128; box_t >= 64; box_t -= 64) {
      CUtensorMap* m = box_t == 128 ? &p->tm_ffa2_a[l] : &p->tm_ffa2_wg_x[l];
      if (lf.view_c >= 0) {
        const int rows = (L8.Ti - lf.view_c + 3) / 4;
        rc = tc::make_map_3d(m, x + (int64_t)lf.view_c * L8.ld_in, L8.Cin, rows, npl * B, 4 * L8.ld_in,
                             (int64_t)L8.Ti * L8.ld_in, 64, box_t);
      } else {
        rc = tc::make_map_3d(m, bf(p, p->off_ffa2_s[lf.seq]), L8.Cin, p->ffa2_Tqi, npl * B, L8.ld_in,
                             (int64_t)p->ffa2_Tqi * L8.ld_in, 64, box_t);
      }
      if (rc) return rc;
    }
    if (p->bmn) {
      rc = tc::make_map_2d(&p->tm_ffa2_b[l], bf(p, p->off_ffa2_wb[l]), L8.Cout, npl * J * L8.Cin, L8.ld_co, 64, 64);
    } else {
      rc = tc::make_map_2d(&p->tm_ffa2_b[l], bf(p, p->off_ffa2_w[l]), J * L8.cin_p, npl * L8.Cout,
                           (int64_t)J * L8.cin_p, 64, p->ffa2_pair_fwd ? wide_n(p) / 2 : wide_n(p));
    }
    if (rc) return rc;
    // backward: the gradient of the leaf product as planes [npl*B][Tq][Cout]
    const __nv_bfloat16* dA = bf(p, p->off_ffa2_p[l]);
    rc = tc::make_map_3d(&p->tm_ffa2_dg_a[l], dA, L8.Cout, p->ffa2_Tq, npl * B, L8.ld_out,
                         (int64_t)p->ffa2_Tq * L8.ld_out, 64, 128);
    if (rc) return rc;
    rc = tc::make_map_3d(&p->tm_ffa2_wg_dz[l], dA, L8.Cout, p->ffa2_Tq, npl * B, L8.ld_out,
                         (int64_t)p->ffa2_Tq * L8.ld_out, 64, 64);
    if (rc) return rc;
    rc = tc::make_map_2d(&p->tm_ffa2_dg_b[l], bf(p, p->off_ffa2_wb[l]), L8.Cout, npl * J * L8.Cin, L8.ld_co, 64,
                         p->ffa2_pair_dg ? wide_n(p) / 2 : wide_n(p));
    if (rc) return rc;
  }
  return ST_OK;
}

}  // namespace

ST_API int st_plan_bind(st_plan* p, void* arena, size_t arena_bytes, float* params, float* grads) {
  ST_CHECK_ARG(p && arena && params && grads, "st_plan_bind: null pointer");
  ST_CHECK_ARG(arena_bytes >= p->arena_bytes, "st_plan_bind: arena %zu < required %zu bytes", arena_bytes, p->arena_bytes);
  ST_CHECK_ARG((reinterpret_cast<uintptr_t>(arena) & 1023) == 0, "st_plan_bind: arena must be 1024-byte aligned");
  ST_CHECK_ARG((reinterpret_cast<uintptr_t>(params) & 15) == 0 && (reinterpret_cast<uintptr_t>(grads) & 15) == 0,
               "st_plan_bind: parameter / gradient buffers must be 16-byte aligned");
  p->arena = static_cast<char*>(arena);
  p->params = params;
  p->grads = grads;
  const int B = p->B, npl = p->npl;
  int rc;
  for (int l = 0; l < 11; ++l) {
    Layer& L = p->layers[l];
    const __nv_bfloat16* xin = act_in(p, l);
    const int block_n = l == 10 ? 32 : wide_n(p);
    // ---- forward: A = input activation planes, B = forward filter planes [Cout rows][K*cin_p]
    if (L.stride == 1) {
      rc = tc::make_map_3d(&L.tm_fwd_a, xin, L.Cin, L.Ti, npl * B, L.ld_in, (int64_t)L.Ti * L.ld_in, 64, 128);
    } else {
      // pair view [B, Tpad/2, 2*Cin] of the (even-padded) input planes
      rc = tc::make_map_3d(&L.tm_fwd_a, xin, 2 * L.Cin, p->Tpad / 2, npl * B, 2 * L.ld_in, (int64_t)p->Tpad * L.ld_in,
                           64, 128);
    }
    if (rc) return rc;
    L.pair_fwd = tc::want_pair(B * ((L.To + tc::kTileM - 1) / tc::kTileM), (L.Cout + block_n - 1) / block_n, block_n,
                               npl, false, L.K * (L.cin_p / 64));
    if (l == 9 && p->bmn) {
      rc = tc::make_map_2d(&L.tm_fwd_b, bf(p, L.off_wbwd), L.Cout, npl * L.K * L.Cin, L.ld_co, 64, 64);
    } else {
      rc = tc::make_map_2d(&L.tm_fwd_b, bf(p, L.off_wfwd), L.K * L.cin_p, npl * L.Cout, (int64_t)L.K * L.cin_p, 64,
                           L.pair_fwd ? block_n / 2 : block_n);
    }
    if (rc) return rc;
    // ---- filter gradient: X = input activation planes (boxes of 64 rows), dZ = gradient wrt this layer's output
    if (L.stride == 1) {
      rc = tc::make_map_3d(&L.tm_wg_x, xin, L.Cin, L.Ti, npl * B, L.ld_in, (int64_t)L.Ti * L.ld_in, 64, 64);
    } else {
      rc = tc::make_map_3d(&L.tm_wg_x, xin, 2 * L.Cin, p->Tpad / 2, npl * B, 2 * L.ld_in, (int64_t)p->Tpad * L.ld_in,
                           64, 64);
    }
    if (rc) return rc;
    for (int s = 0; s < 2; ++s) {
      const __nv_bfloat16* dz = l == 10 ? bf(p, p->off_dlogits) : (l <= 7 ? bf(p, p->off_dzs[l]) : bf(p, p->off_dz[s]));
      const int ld_dz = l == 10 ? 64 : L.ld_out;
      const int64_t plane_rows = (int64_t)L.To * ld_dz;
      // NOTE: planes of a dz buffer are p->dz_elems apart only for the shared ping/pong buffers; expressing the
      // plane index as an extra "batch" needs a uniform stride, so dz planes are laid out [npl][B][To][ld] densely
      // inside their buffer (plane stride = B*To*ld) -- see dz_plane_stride().
      rc = tc::make_map_3d(&L.tm_wg_dz[s], dz, l == 10 ? 64 : L.Cout, L.To, npl * B, ld_dz, plane_rows, 64, 64);
      if (rc) return rc;
      // ---- data gradient (layers 1..10): A = dz of this layer, B = backward filter planes [K*Cin rows][ld_co]
      if (l > 0) {
        rc = tc::make_map_3d(&L.tm_dg_a[s], dz, l == 10 ? 64 : L.Cout, L.To, npl * B, ld_dz, plane_rows, 64, 128);
        if (rc) return rc;
      }
    }
    if (l > 0) {
      L.pair_dg = !(l == 10 && p->l10_n128) &&
                  tc::want_pair(B * ((L.Ti + tc::kTileM - 1) / tc::kTileM), (L.Cin + wide_n(p) - 1) / wide_n(p),
                                wide_n(p), npl, false, L.K * (l == 10 ? 1 : (L.Cout + 63) / 64));
      rc = tc::make_map_2d(&L.tm_dg_b, bf(p, L.off_wbwd), l == 10 ? 64 : L.Cout, npl * L.K * L.Cin, L.ld_co, 64,
                           L.pair_dg ? wide_n(p) / 2 : wide_n(p));
      if (rc) return rc;
      if (l == 10 && p->l10_n128) {
        rc = tc::make_map_2d(&L.tm_dg_b_n128, bf(p, L.off_wbwd), 64, npl * L.K * L.Cin, L.ld_co, 64, 128);
        if (rc) return rc;
      }
    }
    // ---- store maps for the epilogues (planes are dense [npl][B][T][ld] inside their buffer)
    if (p->tma_store) {
      if (l < 10) {
        rc = tc::make_map_3d_store(&L.tm_fwd_out, bf(p, L.off_out), L.ld_out, L.To, npl * B, L.ld_out,
                                   (int64_t)L.To * L.ld_out);
        if (rc) return rc;
      }
      if (l > 0) {
        const Layer& Lb = p->layers[l - 1];
        for (int s = 0; s < 2; ++s) {
          rc = tc::make_map_3d_store(&L.tm_dg_out[s], l - 1 <= 7 ? bf(p, p->off_dzs[l - 1]) : bf(p, p->off_dz[s]),
                                     Lb.ld_out, Lb.To, npl * B, Lb.ld_out, (int64_t)Lb.To * Lb.ld_out);
          if (rc) return rc;
        }
      }
    }
  }
  if (p->ffa == 2) {
    rc = bind_ffa2(p);
    if (rc) return rc;
  }
  if (p->ffa == 1) {
    // A operands of the three half-rate convolutions: row views of the layer-7 output x (odd rows, even rows) and
    // the pair-sum planes; B operands: the packed even-tap, odd-tap and summed filters (forward layout, 16 taps)
    Layer& L8 = p->layers[8];
    const __nv_bfloat16* x = act_in(p, 8);
    p->ffa_pair_fwd = tc::want_pair(B * ((p->ffa_Tu + tc::kTileM - 1) / tc::kTileM), 1, wide_n(p), npl, true,
                                    (L8.K / 2) * (L8.cin_p / 64));
    p->ffa_pair_dg = tc::want_pair(B * ((p->ffa_Tx + tc::kTileM - 1) / tc::kTileM), 1, wide_n(p), npl, true,
                                   (L8.K / 4) * ((L8.Cout + 63) / 64));
    rc = tc::make_map_3d(&p->tm_ffa_a[0], x + L8.ld_in, L8.Cin, L8.Ti / 2, npl * B, 2 * L8.ld_in,
                         (int64_t)L8.Ti * L8.ld_in, 64, 128);
    if (rc) return rc;
    rc = tc::make_map_3d(&p->tm_ffa_a[1], x, L8.Cin, (L8.Ti + 1) / 2, npl * B, 2 * L8.ld_in,
                         (int64_t)L8.Ti * L8.ld_in, 64, 128);
    if (rc) return rc;
    rc = tc::make_map_3d(&p->tm_ffa_a[2], bf(p, p->off_ffa_xs), L8.Cin, p->ffa_Tx, npl * B, L8.ld_in,
                         (int64_t)p->ffa_Tx * L8.ld_in, 64, 128);
    if (rc) return rc;
    // the same three A operands in 64-row boxes for the filter gradient
    rc = tc::make_map_3d(&p->tm_ffa_wg_x[0], x + L8.ld_in, L8.Cin, L8.Ti / 2, npl * B, 2 * L8.ld_in,
                         (int64_t)L8.Ti * L8.ld_in, 64, 64);
    if (rc) return rc;
    rc = tc::make_map_3d(&p->tm_ffa_wg_x[1], x, L8.Cin, (L8.Ti + 1) / 2, npl * B, 2 * L8.ld_in,
                         (int64_t)L8.Ti * L8.ld_in, 64, 64);
    if (rc) return rc;
    rc = tc::make_map_3d(&p->tm_ffa_wg_x[2], bf(p, p->off_ffa_xs), L8.Cin, p->ffa_Tx, npl * B, L8.ld_in,
                         (int64_t)p->ffa_Tx * L8.ld_in, 64, 64);
    if (rc) return rc;
    for (int i = 0; i < 3; ++i) {
      rc = tc::make_map_2d(&p->tm_ffa_b[i], bf(p, p->off_ffa_w[i]), (L8.K / 2) * L8.cin_p, npl * L8.Cout,
                           (int64_t)(L8.K / 2) * L8.cin_p, 64, p->ffa_pair_fwd ? wide_n(p) / 2 : wide_n(p));
      if (rc) return rc;
      // backward: gradients of the partial products as planes [npl*B][Tu][Cout]
      const __nv_bfloat16* dA = bf(p, p->off_ffa_p[i]);
      rc = tc::make_map_3d(&p->tm_ffa_dg_a[i], dA, L8.Cout, p->ffa_Tu, npl * B, L8.ld_out,
                           (int64_t)p->ffa_Tu * L8.ld_out, 64, 128);
      if (rc) return rc;
      rc = tc::make_map_3d(&p->tm_ffa_wg_dz[i], dA, L8.Cout, p->ffa_Tu, npl * B, L8.ld_out,
                           (int64_t)p->ffa_Tu * L8.ld_out, 64, 64);
      if (rc) return rc;
      rc = tc::make_map_2d(&p->tm_ffa_dg_b[i], bf(p, p->off_ffa_wb[i]), L8.Cout, npl * (L8.K / 2) * L8.Cin, L8.ld_co, 64,
                           p->ffa_pair_dg ? wide_n(p) / 2 : wide_n(p));
      if (rc) return rc;
    }
  }
  p->bound = true;
  return ST_OK;
}

// fp32 parameters -> bf16 operand planes (forward K-major layout for every layer, backward layout for layers 1..10)
namespace {

int pack_layers(st_plan* p, cudaStream_t s) {
  tc::PackTable tab{};
  Layer& L8 = p->layers[8];
  // layers 8-9 as background kernels on the side stream (MN-major forward filters: their packing is its own launches)
  const bool bg = p->overlap && p->bmn;
  cudaStream_t sb = s;
  if (bg) {
    // a previous hand-over nobody consumed (packing twice without a forward in between) must not be lost
    int rcj = side_join(p, &p->pack_pending, p->ev_pack, s);
    if (rcj) return rcj;
    rcj = side_fork(p, s);
    if (rcj) return rcj;
    sb = p->side;
  }
  bool side_used = false;
  for (int l = 0; l < 11; ++l) {
    Layer& L = p->layers[l];
    if (l == 8 && p->ffa == 2) {
      // the nine leaf filters of the two-level split: their own launch, one pass over the 32-tap tensor
      __nv_bfloat16* f9[9];
      __nv_bfloat16* b9[9];
      int m9[9];
      for (int i = 0; i < 9; ++i) {
        f9[i] = bf(p, p->off_ffa2_w[i]);
        b9[i] = bf(p, p->off_ffa2_wb[i]);
        m9[i] = kLeaves[i].tap_mask;
      }
      const int rc2 = p->bmn ? tc::launch_pack_bwd(p->params + L8.w_off, b9, m9, 9, 4, L8.K / 4, L8.Cin, L8.Cout, L8.ld_co,
                                                   p->npl, sb, bg)
                             : tc::launch_pack_ffa2(p->params + L8.w_off, f9, b9, m9, L8.K / 4, L8.Cin, L8.Cout, L8.cin_p,
                                                    L8.ld_co, p->npl, s);
      if (rc2) return rc2;
      side_used = side_used || (bg && p->bmn);
      p->launches++;
      continue;
    }
    if (l == 9 && p->bmn) {
      __nv_bfloat16* b1[1] = {bf(p, L.off_wbwd)};
      const int one = 1;
      const int rc2 = tc::launch_pack_bwd(p->params + L.w_off, b1, &one, 1, 1, L.K, L.Cin, L.Cout, L.ld_co, p->npl, sb, bg);
      if (rc2) return rc2;
      side_used = side_used || bg;
      p->launches++;
      continue;
    }
    if (l == 8 && p->ffa == 1) {
      // fast-FIR filters of layer 8: even taps, odd taps and their sum, each packed like a 16-tap filter in both
      // operand layouts straight from the 32-tap fp32 tensor; the direct 32-tap layouts are then not needed
      for (int i = 0; i < 3; ++i)
        tab.e[tab.n++] = tc::PackEntry{p->params + L8.w_off, bf(p, p->off_ffa_w[i]), bf(p, p->off_ffa_wb[i]), L8.K / 2,
                                       L8.Cin, L8.Cout, L8.cin_p, L8.ld_co, 0, 2, i + 1};
      continue;
    }
    tab.e[tab.n++] = tc::PackEntry{p->params + L.w_off, bf(p, L.off_wfwd), l > 0 ? bf(p, L.off_wbwd) : nullptr,
                                   L.K, L.Cin, L.Cout, L.cin_p, L.ld_co, 0, 1, 1};
  }
  if (side_used) {
    ST_CUDA_CALL(cudaEventRecord(p->ev_pack, p->side));
    p->pack_pending = true;                       // st_plan_forward waits for it in front of layer 8
  }
  int n = 0;
  const int rc = tc::launch_pack_filters(tab, p->npl, s, &n);
  if (rc) return rc;
  p->launches += n;
  return ST_OK;
}

}  // namespace

ST_API int st_plan_pack_weights(st_plan* p, st_stream_t stream) {
  ST_CHECK_ARG(p && p->bound, "st_plan_pack_weights: plan is not bound");
  return pack_layers(p, st_cu(stream));
}

namespace {

// Fast-FIR form of layer 8: y = x (*) w over 32 taps as
//   A00[u] = sum_j odd [u+j-8] w[2j],  A11[u] = sum_j even[u+j-7] w[2j+1],  S[u] = sum_j xs[u+j-7] (w[2j]+w[2j+1])
//   y[2u] = A00[u] + A11[u],           y[2u+1] = S[u] - A11[u] - A00[u+1]            (tools/ffa_study.py)
// with odd[r] = x[2r+1], even[r] = x[2r], xs[r] = x[2r] + x[2r+1]: three 16-tap problems in ONE launch of
// tc_conv_kernel writing fp32 partial products, then bias + ReLU + plane split in ffa_combine_kernel.
int forward_layer8_ffa(st_plan* p, cudaStream_t s) {
  Layer& L = p->layers[8];
  int rc = tc::launch_pair_sum_planes(act_in(p, 8), bf(p, p->off_ffa_xs), p->B, L.Ti, p->ffa_Tx, L.ld_in, p->npl, s);
  if (rc) return rc;
  p->launches++;
  float* part[3];
  tc::ConvParams c{};
  c.taps = L.K / 2;
  c.chunks_per_tap = L.cin_p / 64;
  c.a_sign = 1;
  c.a_stride = 1;
  c.a_cin = 0;
  c.b_row_step = 0;
  c.b_col_step = L.cin_p;
  c.b_plane_rows = L.Cout;
  c.B = p->B; c.To = p->ffa_Tu; c.N = L.Cout;
  c.m_tiles_per_utt = (p->ffa_Tu + tc::kTileM - 1) / tc::kTileM;
  c.n_tiles = (L.Cout + wide_n(p) - 1) / wide_n(p);
  c.n_fastest = 0;
  c.ld_f32 = L.Cout;
  c.k_cols = L.Cin;
  c.trim = p->trim;
  c.n_problems = 3;
  c.k_split = 1;
  c.pair = p->ffa_pair_fwd;
  for (int i = 0; i < 3; ++i) {
    part[i] = reinterpret_cast<float*>(p->arena + p->off_ffa_p[i]);
    c.pad_left_q[i] = i == 0 ? 8 : 7;
    c.out_f32_q[i] = part[i];
  }
  const int ti = timed_begin(p, s);
  rc = tc::launch_conv_multi(p->tm_ffa_a, p->tm_ffa_b, c, wide_n(p), p->npl, s);
  if (rc) return rc;
  timed_end(p, ti, 0, 8, 2.0 * L.K * L.Cin * L.Cout * (double)L.To * p->B, s);
  p->launches++;
  rc = tc::launch_ffa_combine(part[0], part[1], part[2], p->params + L.b_off, L.relu, bf(p, L.off_out), mask_of(p, 8),
                              p->B, L.To, p->ffa_Tu, L.Cout, L.Cout, L.ld_out, p->npl, s);
  if (rc) return rc;
  p->launches++;
  return ST_OK;
}

// Backward of the fast-FIR form (formulas: tools/ffa_study.py, checked against autograd).  dz = planes of
// d(loss)/d(y8); writes the filter gradient of layer 8, the planes of d(loss)/d(x) (masked by layer 7's ReLU) into
// dz_out and layer 7's bias gradient.
int backward_layer8_ffa(st_plan* p, const __nv_bfloat16* dz, __nv_bfloat16* dz_out, cudaStream_t s) {
  Layer& L = p->layers[8];
  Layer& Lb = p->layers[7];
  const int J = L.K / 2;
  __nv_bfloat16* dA[3];
  for (int i = 0; i < 3; ++i) dA[i] = bf(p, p->off_ffa_p[i]);
  int rc = tc::launch_ffa_dz_prep(dz, dA[0], dA[1], dA[2], p->B, L.To, p->ffa_Tu, L.ld_out, p->npl, s);
  if (rc) return rc;
  p->launches++;
  // ---- filter gradient: dw[2j] = <odd, dA00>_j + <xs, dS>_j, dw[2j+1] = <even, dA11>_j + <xs, dS>_j
  float* cs = reinterpret_cast<float*>(p->arena + p->off_ffa_cs);
  ST_CUDA_CALL(cudaMemsetAsync(cs, 0, (size_t)J * L.Cin * L.Cout * sizeof(float), s));
  tc::WgradParams w{};
  w.B = p->B; w.To = p->ffa_Tu; w.t_chunks = (p->ffa_Tu + 63) / 64;
  w.taps = J; w.a_stride = 1; w.a_cin = 0;
  w.m_tiles = (L.Cin + 127) / 128;
  w.n_tiles = (L.Cout + wide_n(p) - 1) / wide_n(p);
  w.Cin = L.Cin; w.Cout = L.Cout;
  w.trim = p->trim;
  w.n_problems = 3;
  float* dW = p->grads + L.w_off;
  w.pad_left_q[0] = 8; w.dW_q[0] = dW;                              w.tap_stride_q[0] = 2;   // even taps
  w.pad_left_q[1] = 7; w.dW_q[1] = dW + (int64_t)L.Cin * L.Cout;    w.tap_stride_q[1] = 2;   // odd taps
  w.pad_left_q[2] = 7; w.dW_q[2] = cs;                              w.tap_stride_q[2] = 1;
  int ti = timed_begin(p, s);
  rc = tc::launch_wgrad_multi(p->tm_ffa_wg_x, p->tm_ffa_wg_dz, w, wide_n(p), p->npl, s);
  if (rc) return rc;
  timed_end(p, ti, 2, 8, 2.0 * L.K * L.Cin * L.Cout * (double)L.To * p->B, s);
  p->launches++;
  rc = tc::launch_ffa_dw_combine(dW, cs, J, (int64_t)L.Cin * L.Cout, s);
  if (rc) return rc;
  p->launches++;
  // ---- data gradient: d_odd = dA00 (*)^T w0, d_even = dA11 (*)^T w1, d_xs = dS (*)^T ws as fp32 [B][Tx][256],
  // each tile's taps in two slices so that 384 work items fill 148 SMs
  float* dxp[3];
  for (int i = 0; i < 3; ++i) dxp[i] = reinterpret_cast<float*>(p->arena + p->off_ffa_dxp[i]);
  const size_t dxp_bytes = (size_t)p->B * p->ffa_Tx * 256 * sizeof(float);
  for (int i = 0; i < 3; ++i) ST_CUDA_CALL(cudaMemsetAsync(dxp[i], 0, dxp_bytes, s));
  tc::ConvParams c{};
  c.taps = J;
  c.chunks_per_tap = (L.Cout + 63) / 64;
  c.a_sign = -1;
  c.a_stride = 1;
  c.a_cin = 0;
  c.b_row_step = L.Cin;
  c.b_col_step = 0;
  c.b_plane_rows = J * L.Cin;
  c.B = p->B; c.To = p->ffa_Tx; c.N = L.Cin;
  c.m_tiles_per_utt = (p->ffa_Tx + tc::kTileM - 1) / tc::kTileM;
  c.n_tiles = (L.Cin + wide_n(p) - 1) / wide_n(p);
  c.n_fastest = 1;
  c.ld_f32 = 256;
  c.k_cols = L.Cout;
  c.trim = p->trim;
  c.n_problems = 3;
  c.k_split = 2;
  c.pair = p->ffa_pair_dg;
  for (int i = 0; i < 3; ++i) {
    c.pad_left_q[i] = i == 0 ? 8 : 7;
    c.out_f32_q[i] = dxp[i];
  }
  ti = timed_begin(p, s);
  rc = tc::launch_conv_multi(p->tm_ffa_dg_a, p->tm_ffa_dg_b, c, wide_n(p), p->npl, s);
  if (rc) return rc;
  timed_end(p, ti, 1, 8, 2.0 * L.K * L.Cin * L.Cout * (double)L.To * p->B, s);
  p->launches++;
  rc = tc::launch_ffa_dx_combine(dxp[0], dxp[1], dxp[2], mask_of(p, 7), dz_out, p->grads + Lb.b_off, p->B, L.Ti,
                                 p->ffa_Tx, L.Cin, 256, Lb.ld_out, p->npl, s);
  if (rc) return rc;
  p->launches++;
  return ST_OK;
}

// ---- two levels: nine quarter-rate 8-tap problems per pass (leaf table kLeaves, algebra in tools/ffa2_study.py)
int forward_layer8_ffa2(st_plan* p, cudaStream_t s) {
  Layer& L = p->layers[8];
  __nv_bfloat16* seq[5];
  for (int i = 0; i < 5; ++i) seq[i] = bf(p, p->off_ffa2_s[i]);
  int rc = tc::launch_ffa2_inputs(act_in(p, 8), seq, p->B, L.Ti, p->ffa2_Tqi, L.ld_in, p->npl, s);
  if (rc) return rc;
  p->launches++;
  float* part[9];
  tc::ConvParams c{};
  c.taps = L.K / 4;
  c.chunks_per_tap = L.cin_p / 64;
  c.a_sign = 1;
  c.a_stride = 1;
  c.a_cin = 0;
  c.b_row_step = 0;
  c.b_col_step = L.cin_p;
  c.b_plane_rows = L.Cout;
  c.B = p->B; c.To = p->ffa2_Tq; c.N = L.Cout;
  c.m_tiles_per_utt = (p->ffa2_Tq + tc::kTileM - 1) / tc::kTileM;
  c.n_tiles = (L.Cout + wide_n(p) - 1) / wide_n(p);
  c.n_fastest = 0;
  c.ld_f32 = L.Cout;
  c.k_cols = L.Cin;
  c.trim = p->trim;
  c.n_problems = 9;
  c.k_split = 1;
  c.pair = p->ffa2_pair_fwd;
  if (p->bmn) {
    c.b_mn = 1;
    c.b_row_step = L.Cin;
    c.b_col_step = 0;
    c.b_plane_rows = (L.K / 4) * L.Cin;
  }
  for (int i = 0; i < 9; ++i) {
    part[i] = reinterpret_cast<float*>(p->arena + p->off_ffa2_p[i]);
    c.pad_left_q[i] = kLeaves[i].pad;
    c.out_f32_q[i] = part[i];
  }
  const int ti = timed_begin(p, s);
  rc = tc::launch_conv_multi(p->tm_ffa2_a, p->tm_ffa2_b, c, wide_n(p), p->npl, s);
  if (rc) return rc;
  timed_end(p, ti, 0, 8, 2.0 * L.K * L.Cin * L.Cout * (double)L.To * p->B, s);
  p->launches++;
  rc = tc::launch_ffa2_combine(part, p->params + L.b_off, L.relu, bf(p, L.off_out), mask_of(p, 8), p->B, L.To,
                               p->ffa2_Tq, L.Cout, L.Cout, L.ld_out, p->npl, s);
  if (rc) return rc;
  p->launches++;
  return ST_OK;
}

int backward_layer8_ffa2(st_plan* p, const __nv_bfloat16* dz, __nv_bfloat16* dz_out, cudaStream_t s) {
  Layer& L = p->layers[8];
  Layer& Lb = p->layers[7];
  const int J = L.K / 4;
  __nv_bfloat16* dA[9];
  float* cw[9];
  float* dxp[9];
  for (int i = 0; i < 9; ++i) {
    dA[i] = bf(p, p->off_ffa2_p[i]);
    cw[i] = reinterpret_cast<float*>(p->arena + p->off_ffa2_c[i]);
    dxp[i] = reinterpret_cast<float*>(p->arena + p->off_ffa2_dxp[i]);
  }
  int rc = tc::launch_ffa2_dz_prep(dz, dA, p->B, L.To, p->ffa2_Tq, L.ld_out, p->npl, s);
  if (rc) return rc;
  p->launches++;
  // ---- filter gradient: nine leaf correlations into their own [J][Cin][Cout] buffers, then one combine into dW
  tc::WgradParams w{};
  w.B = p->B; w.To = p->ffa2_Tq; w.t_chunks = (p->ffa2_Tq + 63) / 64;
  w.taps = J; w.a_stride = 1; w.a_cin = 0;
  w.m_tiles = (L.Cin + 127) / 128;
  w.n_tiles = (L.Cout + wide_n(p) - 1) / wide_n(p);
  w.Cin = L.Cin; w.Cout = L.Cout;
  w.trim = p->trim;
  w.n_problems = 9;
  for (int i = 0; i < 9; ++i) { w.pad_left_q[i] = kLeaves[i].pad; w.dW_q[i] = cw[i]; w.tap_stride_q[i] = 1; }
  if (tc::wgrad_accumulates(9 * J * w.m_tiles * w.n_tiles, p->B * w.t_chunks,
                            tc::wgrad_pair(J, w.m_tiles, wide_n(p), p->npl))) {
    const size_t bytes = (size_t)J * L.Cin * L.Cout * sizeof(float);
    for (int i = 0; i < 9; ++i) ST_CUDA_CALL(cudaMemsetAsync(cw[i], 0, bytes, s));
  }
  int ti = timed_begin(p, s);
  rc = tc::launch_wgrad_multi(p->tm_ffa2_wg_x, p->tm_ffa2_wg_dz, w, wide_n(p), p->npl, s);
  if (rc) return rc;
  timed_end(p, ti, 2, 8, 2.0 * L.K * L.Cin * L.Cout * (double)L.To * p->B, s);
  p->launches++;
  if (p->overlap && p->range_lo == 0) {
    // the combine (HBM-bound, 0.2 GB) runs on the side stream underneath the data gradients of layers 8 .. 1; its
    // result is first needed behind the backward pass (st_plan_backward_range joins at its end).  Ranges that stop
    // above layer 0 (data parallel: the allreduce of layers 8-10 is launched right after them) keep it in line.
    rc = side_fork(p, s);
    if (rc) return rc;
    rc = tc::launch_ffa2_dw_combine(p->grads + L.w_off, cw, J, (int64_t)L.Cin * L.Cout, p->side, true);
    if (rc) return rc;
    ST_CUDA_CALL(cudaEventRecord(p->ev_dw, p->side));
    p->dw_pending = true;
  } else {
    rc = tc::launch_ffa2_dw_combine(p->grads + L.w_off, cw, J, (int64_t)L.Cin * L.Cout, s);
    if (rc) return rc;
  }
  p->launches++;
  // ---- data gradient: nine transposed 8-tap correlations -> fp32 [B][Tqi][256], whole contraction per tile
  tc::ConvParams c{};
  c.taps = J;
  c.chunks_per_tap = (L.Cout + 63) / 64;
  c.a_sign = -1;
  c.a_stride = 1;
  c.a_cin = 0;
  c.b_row_step = L.Cin;
  c.b_col_step = 0;
  c.b_plane_rows = J * L.Cin;
  c.B = p->B; c.To = p->ffa2_Tqi; c.N = L.Cin;
  c.m_tiles_per_utt = (p->ffa2_Tqi + tc::kTileM - 1) / tc::kTileM;
  c.n_tiles = (L.Cin + wide_n(p) - 1) / wide_n(p);
  c.n_fastest = 1;
  c.ld_f32 = 256;
  c.k_cols = L.Cout;
  c.trim = p->trim;
  c.n_problems = 9;
  c.k_split = 1;
  c.pair = p->ffa2_pair_dg;
  for (int i = 0; i < 9; ++i) { c.pad_left_q[i] = kLeaves[i].pad; c.out_f32_q[i] = dxp[i]; }
  ti = timed_begin(p, s);
  rc = tc::launch_conv_multi(p->tm_ffa2_dg_a, p->tm_ffa2_dg_b, c, wide_n(p), p->npl, s);
  if (rc) return rc;
  timed_end(p, ti, 1, 8, 2.0 * L.K * L.Cin * L.Cout * (double)L.To * p->B, s);
  p->launches++;
  rc = tc::launch_ffa2_dx_combine(dxp, mask_of(p, 7), dz_out, p->grads + Lb.b_off, p->B, L.Ti, p->ffa2_Tqi, L.Cin,
                                  256, Lb.ld_out, p->npl, s);
  if (rc) return rc;
  p->launches++;
  return ST_OK;
}

}  // namespace

ST_API int st_plan_forward(st_plan* p, const float* inputs, st_stream_t stream) {
  ST_CHECK_ARG(p && p->bound && inputs, "st_plan_forward: plan is not bound / null input");
  cudaStream_t s = st_cu(stream);
  int rc = tc::launch_split_input(inputs, bf(p, p->off_in), p->B, p->T, p->Tpad, p->F, p->npl, s);
  if (rc) return rc;
  p->launches++;
  for (int l = 0; l < 11; ++l) {
    Layer& L = p->layers[l];
    const int block_n = l == 10 ? 32 : wide_n(p);
    if (l == 8) {                                 // the filters of layers 8-9 may still be on their way (side stream)
      rc = side_join(p, &p->pack_pending, p->ev_pack, s);
      if (rc) return rc;
    }
    if (l == 8 && p->ffa) {
      rc = p->ffa == 2 ? forward_layer8_ffa2(p, s) : forward_layer8_ffa(p, s);
      if (rc) return rc;
      continue;
    }
    tc::ConvParams c{};
    c.taps = L.K;
    c.chunks_per_tap = L.cin_p / 64;
    c.pad_left = L.pad_left;
    c.a_sign = 1;
    c.a_stride = L.stride;
    c.a_cin = L.Cin;
    c.b_row_step = 0;
    c.b_col_step = L.cin_p;
    c.b_plane_rows = L.Cout;
    c.B = p->B; c.To = L.To; c.N = L.Cout;
    c.m_tiles_per_utt = (L.To + tc::kTileM - 1) / tc::kTileM;
    c.n_tiles = (L.Cout + block_n - 1) / block_n;
    c.n_fastest = (size_t)p->npl * p->B * L.Ti * L.ld_in * 2 > (size_t)48 << 20;   // A planes too big for L2
    c.bias = p->params + L.b_off;
    c.relu = L.relu;
    if (l < 10) {
      c.out_planes = bf(p, L.off_out);
      c.out_plane_stride = (int64_t)p->B * L.To * L.ld_out;
      c.ld_out = L.ld_out;
      c.mask_out = L.relu ? mask_of(p, l) : nullptr;
      c.mask_rows = (int64_t)p->B * L.To;
    } else {
      c.out_f32 = reinterpret_cast<float*>(p->arena + p->off_logits);
      c.ld_f32 = 32;
    }
    c.tma_store = p->tma_store && l < 10;
    c.k_cols = L.Cin;
    c.trim = p->trim;
    c.pair = L.pair_fwd;
    if (l == 9 && p->bmn) {
      c.b_mn = 1;
      c.b_row_step = L.Cin;
      c.b_col_step = 0;
      c.b_plane_rows = L.K * L.Cin;
    }
    const int ti = timed_begin(p, s);
    rc = tc::launch_conv(L.tm_fwd_a, L.tm_fwd_b, l < 10 ? &L.tm_fwd_out : nullptr, c, block_n, p->npl, s);
    if (rc) return rc;
    timed_end(p, ti, 0, l, 2.0 * L.K * L.Cin * L.Cout * (double)L.To * p->B, s);
    p->launches++;
  }
  return ST_OK;
}

// Consumes d(loss)/d(logits) from the dlogits planes (written by st_ctc_loss) and fills the flat gradient buffer.
// st_plan_backward_range runs layers hi..lo (hi >= lo); call with hi=10 first, then continue downwards -- the caller
// may launch the gradient allreduce of the finished layers in between (speecht_b200/tc_plan.py).
ST_API int st_plan_backward_range(st_plan* p, int hi, int lo, st_stream_t stream) {
  ST_CHECK_ARG(p && p->bound, "st_plan_backward: plan is not bound");
  ST_CHECK_ARG(hi <= 10 && lo >= 0 && hi >= lo, "st_plan_backward_range: need 10 >= hi >= lo >= 0");
  cudaStream_t s = st_cu(stream);
  if (hi == 10) {
    p->cur_dz = 0;
    // filter gradients of K-sliced tiles accumulate with atomics: the whole flat buffer is zeroed once per backward --
    // already done underneath the forward pass when the caller announced the backward (st_plan_prepare_backward)
    if (p->zero_pending) {
      const int rcz = side_join(p, &p->zero_pending, p->ev_zero, s);
      if (rcz) return rcz;
    } else {
      ST_CUDA_CALL(cudaMemsetAsync(p->grads, 0, (size_t)st_plan_param_floats(p) * sizeof(float), s));
    }
  }
  p->range_lo = lo;
  int cur = p->cur_dz;                          // ping / pong buffer holding the gradient wrt layer 8's / 9's output
  int deferred[8], n_deferred = 0;              // 250-channel layers whose filter gradient waits for the merged launch
  for (int l = hi; l >= lo; --l) {
    Layer& L = p->layers[l];
    const __nv_bfloat16* dz = l == 10 ? bf(p, p->off_dlogits) : (l <= 7 ? bf(p, p->off_dzs[l]) : bf(p, p->off_dz[cur]));
    const int ld_dz = l == 10 ? 64 : L.ld_out;
    const int64_t rows = (int64_t)p->B * L.To;
    int rc = ST_OK;
    if (l == 8 && p->ffa) {
      const int nxt = cur ^ 1;
      rc = p->ffa == 2 ? backward_layer8_ffa2(p, dz, bf(p, p->off_dzs[7]), s)
                       : backward_layer8_ffa(p, dz, bf(p, p->off_dzs[7]), s);
      if (rc) return rc;
      cur = nxt;
      continue;
    }
    if (l == 10) {
      // bias gradient of the last layer from the CTC gradient planes; layers 0..9 get theirs from the epilogue of
      // the data-gradient kernel that produces their dz (ConvParams::col_sum)
      rc = tc::launch_bias_grad(dz, rows, L.Cout, ld_dz, p->npl, p->grads + L.b_off, s);
      if (rc) return rc;
      p->launches++;
    }
    // ---- filter gradient (the seven 250-channel layers: deferred to one multi-problem launch after the loop)
    const bool defer = p->merge_wgrad && l >= 1 && l <= 7;
    if (defer) deferred[n_deferred++] = l;
    tc::WgradParams w{};
    w.B = p->B; w.To = L.To; w.t_chunks = (L.To + 63) / 64;
    w.taps = L.K; w.pad_left = L.pad_left; w.a_stride = L.stride; w.a_cin = L.Cin;
    w.m_tiles = (L.Cin + 127) / 128;
    w.n_tiles = l == 10 ? 1 : (L.Cout + wide_n(p) - 1) / wide_n(p);
    w.Cin = L.Cin; w.Cout = L.Cout;
    w.dW = p->grads + L.w_off;
    w.trim = p->trim;
    int ti = -1;
    if (!defer) {
      ti = timed_begin(p, s);
      rc = tc::launch_wgrad(L.tm_wg_x, L.tm_wg_dz[l == 10 ? 0 : cur], w, l == 10 ? 64 : wide_n(p), p->npl, s);
      if (rc) return rc;
      timed_end(p, ti, 2, l, 2.0 * L.K * L.Cin * L.Cout * (double)L.To * p->B, s);
      p->launches++;
    }
    // ---- data gradient, ReLU mask of the layer below fused
    if (l > 0) {
      Layer& Lb = p->layers[l - 1];
      const int nxt = l == 10 ? 0 : cur ^ 1;
      tc::ConvParams c{};
      c.taps = L.K;
      c.chunks_per_tap = l == 10 ? 1 : (L.Cout + 63) / 64;
      c.pad_left = L.pad_left;
      c.a_sign = -1;
      c.a_stride = 1;
      c.a_cin = 0;
      c.b_row_step = L.Cin;
      c.b_col_step = 0;
      c.b_plane_rows = L.K * L.Cin;
      const bool n128 = l == 10 && p->l10_n128;
      const int dg_block_n = n128 ? 128 : wide_n(p);
      c.B = p->B; c.To = L.Ti; c.N = L.Cin;
      c.m_tiles_per_utt = (L.Ti + tc::kTileM - 1) / tc::kTileM;
      c.n_tiles = (L.Cin + dg_block_n - 1) / dg_block_n;
      c.n_fastest = (size_t)p->npl * p->B * L.To * ld_dz * 2 > (size_t)48 << 20;
      c.bias = nullptr;
      c.relu = 0;
      c.out_planes = l - 1 <= 7 ? bf(p, p->off_dzs[l - 1]) : bf(p, p->off_dz[nxt]);
      c.out_plane_stride = (int64_t)p->B * Lb.To * Lb.ld_out;
      c.ld_out = Lb.ld_out;
      c.mask_bits = mask_of(p, l - 1);
      c.mask_rows = (int64_t)p->B * Lb.To;
      c.col_sum = p->grads + Lb.b_off;
      c.tma_store = p->tma_store;
      c.k_cols = L.Cout;
      c.trim = p->trim;
      c.pair = L.pair_dg;
      ti = timed_begin(p, s);
      rc = tc::launch_conv(L.tm_dg_a[l == 10 ? 0 : cur], n128 ? L.tm_dg_b_n128 : L.tm_dg_b, &L.tm_dg_out[nxt], c,
                           dg_block_n, p->npl, s);
      if (rc) return rc;
      timed_end(p, ti, 1, l, 2.0 * L.K * L.Cin * L.Cout * (double)L.To * p->B, s);
      p->launches++;
      cur = nxt;
    }
  }
  p->cur_dz = cur;
  if (n_deferred > 0) {
    // ONE launch for the filter gradients of the deferred layers (identical shapes: 250 -> 250 channels, 7 taps): 14
    // tiles per layer are far fewer than SMs, so every tile's time contraction is cut into slices (force_split) --
    // seven separate launches each paid their own prologue, pipeline fill and accumulation epilogue for ~26 us of MMAs
    const Layer& L = p->layers[deferred[0]];
    CUtensorMap tx[tc::kMaxProblems], tz[tc::kMaxProblems];
    tc::WgradParams w{};
    w.B = p->B; w.To = L.To; w.t_chunks = (L.To + 63) / 64;
    w.taps = L.K; w.pad_left = L.pad_left; w.a_stride = L.stride; w.a_cin = L.Cin;
    w.m_tiles = (L.Cin + 127) / 128;
    w.n_tiles = (L.Cout + wide_n(p) - 1) / wide_n(p);
    w.Cin = L.Cin; w.Cout = L.Cout;
    w.trim = p->trim;
    w.n_problems = n_deferred;
    double flops = 0;
    for (int i = 0; i < n_deferred; ++i) {
      const Layer& Li = p->layers[deferred[i]];
      tx[i] = Li.tm_wg_x;
      tz[i] = Li.tm_wg_dz[0];
      w.pad_left_q[i] = Li.pad_left;
      w.dW_q[i] = p->grads + Li.w_off;
      w.tap_stride_q[i] = 1;
      flops += 2.0 * Li.K * Li.Cin * Li.Cout * (double)Li.To * p->B;
    }
    int rc;
    const int ti = timed_begin(p, s);
    if (n_deferred == 1) {
      w.dW = w.dW_q[0];
      w.n_problems = 0;
      rc = tc::launch_wgrad(tx[0], tz[0], w, wide_n(p), p->npl, s);
    } else {
      w.force_split = tc::wgrad_best_split(n_deferred * w.taps * w.m_tiles * w.n_tiles, p->B * w.t_chunks,
                                           tc::wgrad_pair(w.taps, w.m_tiles, wide_n(p), p->npl));
      rc = tc::launch_wgrad_multi(tx, tz, w, wide_n(p), p->npl, s);
    }
    if (rc) return rc;
    timed_end(p, ti, 2, deferred[n_deferred - 1], flops, s);
    p->launches++;
  }
  return side_join(p, &p->dw_pending, p->ev_dw, s);
}

ST_API int st_plan_backward(st_plan* p, st_stream_t stream) { return st_plan_backward_range(p, 10, 0, stream); }

// Announces a backward pass: the flat gradient buffer is zeroed NOW on the plan's side stream (behind everything
// enqueued on `stream` so far -- the previous step's Adam), underneath the forward pass that the caller enqueues next;
// st_plan_backward[_range] then waits for it instead of zeroing in line.  Optional: without it the backward zeroes.
ST_API int st_plan_prepare_backward(st_plan* p, st_stream_t stream) {
  ST_CHECK_ARG(p && p->bound, "st_plan_prepare_backward: plan is not bound");
  if (!p->overlap) return ST_OK;
  cudaStream_t s = st_cu(stream);
  int rc = side_join(p, &p->zero_pending, p->ev_zero, s);     // an announced backward that never came
  if (rc) return rc;
  rc = side_fork(p, s);
  if (rc) return rc;
  rc = tc::launch_zero_f32(p->grads, st_plan_param_floats(p), p->side, true);
  if (rc) return rc;
  ST_CUDA_CALL(cudaEventRecord(p->ev_zero, p->side));
  p->zero_pending = true;
  p->launches++;
  return ST_OK;
}

// Data-parallel training: the tensor-core grids launched after this call leave `n_sms` SMs free (0 restores the full
// machine).  The NCCL allreduce of the layer-8..10 gradients runs underneath the backward pass of layers 7..0; its
// CTAs cannot share an SM with a tensor-core CTA (227 KB of shared memory, 54 K registers), and the persistent
// filter-gradient grids assign their work statically to 148 CTAs -- without free SMs the collective starts late and
// then delays whole CTAs of those grids.  Process-wide setting (one engine per process in DP runs).
ST_API int st_plan_reserve_sms(st_plan* p, int n_sms) {
  ST_CHECK_ARG(p && n_sms >= 0 && n_sms <= 64, "st_plan_reserve_sms: 0 <= n_sms <= 64");
  tc::set_reserved_sms(n_sms);
  return ST_OK;
}

// Debug / test access: activation planes of layer `layer` output (0..9) merged to fp32 [B][To][Cout];
// layer = -1 gives the split input [B][Tpad][F].
ST_API int st_plan_get_activation(st_plan* p, int layer, float* dst, st_stream_t stream) {
  ST_CHECK_ARG(p && p->bound && dst && layer >= -1 && layer < 10, "st_plan_get_activation: bad argument");
  if (layer < 0)
    return tc::launch_merge_planes(bf(p, p->off_in), (int64_t)p->B * p->Tpad, p->F, p->F, p->npl, dst, p->F,
                                   st_cu(stream));
  Layer& L = p->layers[layer];
  return tc::launch_merge_planes(bf(p, L.off_out), (int64_t)p->B * L.To, L.Cout, L.ld_out, p->npl, dst, L.Cout,
                                 st_cu(stream));
}

// Per-launch timing of the tensor-core kernels.  st_plan_set_timing(plan, 1) starts recording a CUDA event pair
// around every tc_conv / tc_wgrad launch on the launch stream; st_plan_read_timings synchronises the events and
// returns up to `max` records as (kind, layer, flops, milliseconds), clearing the log.  Host-synchronous.
ST_API int st_plan_set_timing(st_plan* p, int enable) {
  ST_CHECK_ARG(p, "st_plan_set_timing: null plan");
  p->timing = enable != 0;
  p->recs.clear();
  p->ev_used = 0;
  return ST_OK;
}

ST_API int st_plan_read_timings(st_plan* p, int* kind, int* layer, double* flops, float* ms, int max) {
  ST_CHECK_ARG(p && kind && layer && flops && ms, "st_plan_read_timings: null pointer");
  int n = 0;
  for (const st_plan::Rec& r : p->recs) {
    if (n >= max) break;
    ST_CUDA_CALL(cudaEventSynchronize(p->ev_pool[r.e1]));
    float t = 0.f;
    ST_CUDA_CALL(cudaEventElapsedTime(&t, p->ev_pool[r.e0], p->ev_pool[r.e1]));
    kind[n] = r.kind; layer[n] = r.layer; flops[n] = r.flops; ms[n] = t;
    ++n;
  }
  p->recs.clear();
  p->ev_used = 0;
  return n;
}

// Debug: the `launch_index`-th tensor-core conv launch after this call (forward and data-gradient launches, in
// stream order) writes eight %globaltimer stamps per CTA into buf[grid][8] (device memory, >= 148*8 int64):
// 0 entry, 1 previous grid complete, 2 first operands landed, 3 last MMA issued, 4 accumulator complete,
// 5 epilogue issued, 6 staging tiles drained, 7 exit.  buf = NULL switches it off.  tools/conv_timeline.py.
ST_API int st_debug_conv_timeline(int64_t* buf, int launch_index) {
  tc::set_conv_timeline(reinterpret_cast<long long*>(buf), launch_index);
  return ST_OK;
}
