// Internal interface between the tcgen05 conv kernels (conv_tc.cu) and the step plan (w2l_plan.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

constexpr int kTileM = 128;      // output rows (time steps of one utterance) per tile
constexpr int kChunkK = 64;      // bf16 elements per 128-byte swizzled shared-memory row

// ---- forward / data-gradient kernel: D[b,t,n] = sum_{tap j, chunk c} A[b, t+shift_j, acol_j + c] * Bm[n + j*brow, j*bcol + c]
struct ConvParams {
  // contraction
  int taps, chunks_per_tap, pad_left, a_sign, a_stride, a_cin;
  int b_row_step, b_col_step, b_plane_rows;
  // output tiling
  int B, To, N, m_tiles_per_utt, n_tiles;
  int n_fastest;                 // tile order: 1 = n tile index fastest, 0 = m tile index fastest
  // epilogue
  const float* bias;
  int relu;
  __nv_bfloat16* out_planes;
  int64_t out_plane_stride;      // elements between planes
  int ld_out;
  float* out_f32;
  int ld_f32;
  // ReLU masks as BIT words: word w of row r (r = b * To + t) holds "output channel 32 w + i was > 0" in bit i, stored
  // column-major [word][rows] so that the 32 lanes (= 32 rows) of an epilogue warp touch one 128-byte line.
  // mask_out (forward, layers with a ReLU): written beside the activation planes; mask_bits (data gradient): the words
  // the forward pass of the layer below wrote.  64 MB of bf16 read back per 2000-channel layer become 4 MB.
  uint32_t* mask_out;
  const uint32_t* mask_bits;
  int64_t mask_rows;             // rows of the mask arrays (B * To)
  float* col_sum;                // nullable: += column sums of the stored tile (bias gradient), [N] fp32
  int tma_store;                 // set by launch_conv: bf16 planes leave through the store tensor map tmOut
  long long* timeline;           // debug (st_debug_conv_timeline): [grid][8] %globaltimer stamps of the CTA's first tile
  int k_cols;                    // valid contraction columns per tap (A channels): zero-filled K steps are not issued
  int trim;                      // 1: issue only the K steps / N columns that hold real channels (0: full tiles)
  // ---- several problems of IDENTICAL shape in one persistent launch (launch_conv_multi; fast-FIR split of layer 8):
  // problem q uses tensor maps A[q] / B[q], its own left padding and fp32 output; the taps of every tile may be cut
  // into k_split slices that accumulate into the (pre-zeroed) fp32 output with vector reductions.  Multi launches
  // write fp32 only: no bias, ReLU, mask, bf16 planes or column sums (the combine pass that follows does those).
  // CTA pairs (cta_group::2): 1 = launch as clusters of two CTAs sharing 256-row tiles.  The B tensor map must then
  // have boxes of block_n / 2 rows (each CTA loads half of the B tile).  Decide with want_pair().
  int pair;
  // 1: the B operand is MN-major -- tmB maps the filter's BACKWARD layout [planes * K * Cin rows][ld_co] with boxes of
  // {64 channels, 64 rows}; b_row_step = rows per tap (Cin), b_plane_rows = K * Cin.  256-wide tiles, <= 2 planes.
  int b_mn;
  int n_problems;                // 0 / 1: single problem (the fields above); 2..kMaxProblems: use the arrays below
  int k_split;                   // 0 / 1: whole contraction per tile
  int pad_left_q[9];
  float* out_f32_q[9];
};
constexpr int kMaxProblems = 9;   // three for one level of the fast-FIR split, nine for two
struct TmSet1 { CUtensorMap m[1]; };
struct TmSetN { CUtensorMap m[kMaxProblems]; };

// ---- filter-gradient kernel: dW[j, ci, co] += sum_{b,t} X[b, t+shift_j, acol_j + ci] * dZ[b, t, co]
struct WgradParams {
  int B, To, t_chunks;
  int taps, pad_left, a_stride, a_cin;
  int m_tiles, n_tiles;
  int Cin, Cout;
  float* dW;                     // must be zero on entry (K-sliced tiles accumulate with atomics)
  int trim;                      // 1: N of the last n tile rounded to 16, K steps past the last time row skipped
  // ---- several problems of identical shape in one launch (launch_wgrad_multi; fast-FIR split of layer 8): problem q
  // correlates X[q] with dZ[q] using its own left padding and writes tap j to dW_q[q] + j * tap_stride_q[q] * Cin * Cout
  int n_problems;                // 0 / 1: single problem
  int pad_left_q[9];
  float* dW_q[9];
  int tap_stride_q[9];
  // > 1: EVERY tile's contraction is cut into this many slices (work items = tiles x slices, slice-major so that a
  // wave works on one K range in phase); all slices accumulate into the pre-zeroed dW.  For launches with fewer tiles
  // than SMs (the merged filter gradients of the 250-channel layers); 0 / 1 = the wave-aligned split.
  int force_split;
};
// slices per tile that balance `num_tiles` tiles of `total_iters` iterations over the SMs (1 = leave the wave-aligned
// split in charge)
int wgrad_best_split(int num_tiles, int total_iters, bool pair);
// does a filter-gradient launch with these tile counts run on CTA pairs (two tiles that share their dZ tile per
// cluster, tcgen05.mma.cta_group::2)?  The split helpers need to know: the scheduling unit is then a tile pair.
bool wgrad_pair(int taps, int m_tiles, int block_n, int n_planes);

int make_map_3d(CUtensorMap* map, const void* base, int C, int T, int Bn, int64_t ld, int64_t batch_stride,
                int box_c, int box_t);
int make_map_2d(CUtensorMap* map, const void* base, int cols, int rows, int64_t ld, int box_c, int box_r);
// store map over output planes [Bn][T][C] (C = padded channel count ld), box {32, 32, 1}
int make_map_3d_store(CUtensorMap* map, const void* base, int C, int T, int Bn, int64_t ld, int64_t batch_stride);

// SMs left free by the tensor-core grids launched from now on (0 = use every SM): room for a concurrent collective.
void set_reserved_sms(int n);
// Debug hook: the conv launch number `launch_index` (counted from the call) writes its per-CTA timeline into buf.
void set_conv_timeline(long long* buf, int launch_index);

// block_n: 32 or 256 (conv) / 64 or 256 (wgrad); n_planes: 1 or 2.
// tmOut: store map of p.out_planes, required when p.out_planes is set and n_planes <= 2 (see make_map_3d_store)
int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmOut, const ConvParams& p,
                int block_n, int n_planes, cudaStream_t stream);
// Should a forward / data-gradient launch of m_tiles x n_tiles tiles (256 wide, 1-2 planes) run on CTA pairs?
// On by default (SPEECHT_B200_PAIR=0 disables them); single launches need more tiles than SMs (the 250-channel layers
// have one tile per SM and nothing to share), multi-problem launches at least two m tiles; k_iters = pipeline
// iterations (64-deep K chunks) per tile.
bool want_pair(int m_tiles, int n_tiles, int block_n, int n_planes, bool multi, int k_iters);
int launch_wgrad(const CUtensorMap& tmX, const CUtensorMap& tmDZ, const WgradParams& p, int block_n, int n_planes,
                 cudaStream_t stream);
int launch_wgrad_multi(const CUtensorMap* tmX, const CUtensorMap* tmDZ, const WgradParams& p, int block_n,
                       int n_planes, cudaStream_t stream);
// p.n_problems problems (2..3) of identical shape in one launch; fp32 outputs p.out_f32_q[q] (zeroed by the caller
// when p.k_split > 1), no epilogue extras.  block_n = 256, n_planes 1 or 2.
int launch_conv_multi(const CUtensorMap* tmA, const CUtensorMap* tmB, const ConvParams& p, int block_n, int n_planes,
                      cudaStream_t stream);

// fp32 [B][T][F] -> planes [n_planes][B][Tpad][F] (rows t >= T zero)
int launch_split_input(const float* x, __nv_bfloat16* planes, int B, int T, int Tpad, int F, int n_planes,
                       cudaStream_t stream);
// W [K][Cin][Cout] fp32 -> forward layout planes [n][Cout][K*cin_p] (K-major), backward layout planes
// [n][K*Cin][ld_co] (bwd may be null); all layers of the table in one launch
struct PackEntry {
  const float* w;
  __nv_bfloat16* fwd;
  __nv_bfloat16* bwd;
  int K, Cin, Cout, cin_p, ld_co;
  int blk0;                      // first block of this layer, filled by launch_pack_filters
  // packed tap k = sum over the set bits c of tap_mask of source tap (tap_group * k + c): plain packing is group 1,
  // mask 1; the fast-FIR filters are group 2 (masks 1, 2, 3 = even taps, odd taps, their sum) or group 4
  int tap_group, tap_mask;
};
struct PackTable {
  int n;
  PackEntry e[20];
};
int launch_pack_filters(PackTable& tab, int n_planes, cudaStream_t stream, int* launches);
// ---- fast-FIR split of a stride-1 layer (experimental, see w2l_plan.cu)
// xs planes [n][B][Tx][ld] of x[b][2r] + x[b][2r+1] from x planes [n][B][T][ld]
int launch_pair_sum_planes(const __nv_bfloat16* x, __nv_bfloat16* xs, int B, int T, int Tx, int ld, int n_planes,
                           cudaStream_t stream);
// y[2u] = act(A00[u] + A11[u] + bias), y[2u+1] = act(S[u] - A11[u] - A00[u+1] + bias): fp32 [B][Tu][ld_p] partial
// products -> bf16 planes [n][B][To][ld_out]
int launch_ffa_combine(const float* a00, const float* a11, const float* sm, const float* bias, int relu,
                       __nv_bfloat16* out, uint32_t* mask_out, int B, int To, int Tu, int N, int ld_p, int ld_out,
                       int n_planes, cudaStream_t stream);
// backward of the same form: dy planes [n][B][To][ld] -> dA00 / dA11 / dS planes [n][B][Tu][ld]
int launch_ffa_dz_prep(const __nv_bfloat16* dy, __nv_bfloat16* d00, __nv_bfloat16* d11, __nv_bfloat16* ds, int B,
                       int To, int Tu, int ld, int n_planes, cudaStream_t stream);
// fp32 partials d_odd / d_even / d_xs [B][Tx][ldp] -> masked dx planes [n][B][T][ld] + column sums into db
int launch_ffa_dx_combine(const float* d_odd, const float* d_even, const float* d_xs, const uint32_t* mask_bits,
                          __nv_bfloat16* out, float* db, int B, int T, int Tx, int N, int ldp, int ld, int n_planes,
                          cudaStream_t stream);
// dW[2j] += cs[j], dW[2j+1] += cs[j] for j < J; tap_elems = Cin*Cout
int launch_ffa_dw_combine(float* dW, const float* cs, int J, int64_t tap_elems, cudaStream_t stream);
// ---- two levels of the split: nine quarter-rate 8-tap problems (leaf order XX XY XZ YX YY YZ ZX ZY ZZ, conv_tc.cu)
// the five summed input sequences s5[i] planes [n][B][Tq][ld] from x planes [n][B][T][ld]
int launch_ffa2_inputs(const __nv_bfloat16* x, __nv_bfloat16* const* s5, int B, int T, int Tq, int ld, int n_planes,
                       cudaStream_t stream);
// nine fp32 partial products [B][Tq][ld_p] -> act(y + bias) planes [n][B][To][ld_out]
int launch_ffa2_combine(float* const* part9, const float* bias, int relu, __nv_bfloat16* out, uint32_t* mask_out, int B,
                        int To, int Tq, int N, int ld_p, int ld_out, int n_planes, cudaStream_t stream);
// dy planes [n][B][To][ld] -> gradients of the nine leaf products, planes [n][B][Tq][ld]
int launch_ffa2_dz_prep(const __nv_bfloat16* dy, __nv_bfloat16* const* out9, int B, int To, int Tq, int ld, int n_planes,
                        cudaStream_t stream);
// nine fp32 input-gradient partials [B][Tqx][ldp] -> masked dx planes [n][B][T][ld] + column sums into db
int launch_ffa2_dx_combine(float* const* g9, const uint32_t* mask_bits, __nv_bfloat16* out, float* db, int B, int T,
                           int Tqx, int N, int ldp, int ld, int n_planes, cudaStream_t stream);
// nine fp32 leaf correlations [J][tap_elems] -> dW [4J][tap_elems]
// background = true (here and below): one block per SM, for launches on a side stream beside tensor-core CTAs
int launch_ffa2_dw_combine(float* dW, float* const* c9, int J, int64_t tap_elems, cudaStream_t stream,
                           bool background = false);
// dst[0..n) = 0 (n a multiple of 4, 16-byte aligned)
int launch_zero_f32(float* dst, int64_t n, cudaStream_t stream, bool background = false);
// the nine leaf filters (tap sums selected by masks9[l] over w[4i + c]) in both operand layouts, one pass over w
int launch_pack_ffa2(const float* w, __nv_bfloat16* const* fwd9, __nv_bfloat16* const* bwd9, const int* masks9, int J,
                     int Cin, int Cout, int cin_p, int ld_co, int n_planes, cudaStream_t stream);
// backward-layout planes only (layers whose forward reads the filter MN-major): leaf l = sum over the bits c of
// masks[l] of the source taps group * k + c, k < J; bwd[l] planes [n][J * Cin][ld_co]; Cout % 8 == 0
int launch_pack_bwd(const float* w, __nv_bfloat16* const* bwd, const int* masks, int n_leaves, int group, int J, int Cin,
                    int Cout, int ld_co, int n_planes, cudaStream_t stream, bool background = false);
// true when a filter-gradient launch of num_tiles tiles accumulates into its outputs (they must then be zeroed)
bool wgrad_accumulates(int num_tiles, int total_iters, bool pair);
// db[n] = sum over rows and planes of dz planes [n_planes][rows][ld]
int launch_bias_grad(const __nv_bfloat16* dz, int64_t rows, int N, int ld, int n_planes, float* db,
                     cudaStream_t stream);
// planes [n][rows][ld] -> fp32 [rows][cols] (debug / tests)
int launch_merge_planes(const __nv_bfloat16* planes, int64_t rows, int cols, int ld, int n_planes, float* dst,
                        int64_t ld_dst, cudaStream_t stream);

}  // namespace tc
