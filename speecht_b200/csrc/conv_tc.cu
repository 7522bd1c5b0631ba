// a6 on the 5th-generation tensor cores: conv1d 'SAME' + bias + ReLU, its data gradient and its filter gradient
// as TMA-fed tcgen05 implicit GEMMs with fp32 accumulators in TMEM.  Replaces tf.nn.conv1d / bias_add / relu
// (reference speech_model.py:155,173,177) and the gradients TF autodiff derives (speech_model.py:78).
//
// Operand format: every fp32 tensor is held as `NPL` bf16 planes whose sum is the value (NPL=1: plain bf16;
// NPL=2: hi + lo split, x = hi + lo + O(2^-17 |x|); NPL=3: three planes, x exact to 2^-25).  A product of two split
// operands is accumulated as hi*hi + hi*lo + lo*hi (NPL=2: 3 MMAs, dropped terms 2^-16 relative) or as the six
// products above 2^-24 (NPL=3, "bf16x6": fp32-equivalent operands) in ONE fp32 TMEM accumulator.
//
// Layout (NWC, the reference's): activations [plane][batch][time][channels], channels contiguous, so a tile of 128
// time steps x 64 channels is one 3-D TMA box {64, 128, 1}; a filter tap k just shifts the time coordinate by
// (k - pad_left) and TMA's out-of-bounds zero fill IS the 'SAME' zero padding (no halo buffers, no im2col).
// Stride 2 (layer 0) is expressed on the pair view [B, T/2, 2*Cin] of the same memory.
//
// tc_conv_kernel (forward and data gradient): persistent, one CTA per SM, 320 threads:
//   warp 0   : TMA producer  (one 64-wide K chunk per pipeline stage; in bf16x3 mode a stage is two independently
//              released load groups {A_hi,B_lo} / {A_lo,B_hi}, so the refill on the critical path is 48 KB)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (128 x BLOCK_N x 16 per instruction); hi*hi products
//              go to the `main` accumulator, every product with a lo plane to the `side` accumulator
//   warps 2-9: epilogue, two warps per TMEM lane quarter: tcgen05.ld 32 lanes x 32 columns of main (+ side),
//              +bias, ReLU (forward: the lane's 32 activity BITS are stored for the data gradient, which reads one
//              word per chunk instead of 64 bytes of bf16 activations), split to planes, staged in a per-warp
//              shared-memory tile and written with one TMA store per plane (ragged edges clipped by the tensor map),
//              bias-gradient column sums.  The 128-wide instantiation (layer-10 data gradient, all epilogue) runs
//              sixteen epilogue warps (ConvCfg::EPW).
//   Variants: NPROB = 9 (several problems of one shape per launch: the fast-FIR leaves of layer 8), PAIR (clusters of
//              two CTAs, tcgen05.mma.cta_group::2 with M = 256, each CTA holding half of the B tile), BMN (B read
//              MN-major from the backward filter layout: no forward layout has to be packed for that layer).
// tc_wgrad_kernel (filter gradient): same roles; both operands are MN-major (the contraction runs over time, the
// slow axis), expressed through the MN-major SWIZZLE_128B shared-memory descriptors; wave-aligned split-K; PAIR: two
// tiles that share their dZ tile per cluster.
// Both are launched with programmatic stream serialization (griddepcontrol.wait after the prologue).
// The elementwise passes around them (packing, fast-FIR prepare / combine, zeroing) are at the end of the kernel
// section; three of them are "background kernels" sized to run beside resident tensor-core CTAs (pack_bwd_kernel).
//
// Roofline: tensor-pipe bound (in practice: bound by the clock the power cap allows while the pipe is 95 % busy).
// Algorithmic FLOPs per launch = 2*K*Cin*Cout*T'*B (unpadded); the tensor pipe executes 3x (bf16x3) / 6x (bf16x6) /
// 1x (bf16) that, plus the padding of 250 channels to 256 and 2.2 % time padding -- MMAs that would only multiply
// the padding of 2000 channels (N beyond 208 in the last tile, K steps beyond the 16 channels of the last chunk) are
// not issued.
#include "st_common.cuh"
#include "conv_tc.h"
#include "tc_ptx.cuh"
#include <stdlib.h>
#include <type_traits>

namespace tc {

namespace {

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr int kEpilogueWarps = 8;              // two per TMEM lane quarter: they split the 32-column chunks
constexpr int kThreads = 64 + 32 * kEpilogueWarps;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kBarrierBytes = 1024;                // mbarriers + TMEM slot, after the pipeline stages
constexpr int kStageTileBytes = 2 * 32 * 32 * 2;   // per epilogue warp: 2 planes x 32 rows x 32 bf16 columns

// PAIR: two CTAs of one cluster share a 256-row tile (cta_group::2): each holds its own 128 rows of A and HALF of
// the B tile, so a stage is a third smaller and the B operand crosses L2 -> shared memory once per pair.
template <int BLOCK_N, int NPL, bool PAIR = false>
struct ConvCfg {
  static constexpr int A_BYTES = kTileM * kChunkK * 2;           // 16 KB: 128 rows x 128 B
  static constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;    // B rows (output channels) held by one CTA
  static constexpr int B_BYTES = B_ROWS * kChunkK * 2;
  static constexpr int STAGE_BYTES = NPL * (A_BYTES + B_BYTES);
  // epilogue staging tiles for the TMA stores (one [2 planes][32][32] bf16 tile per epilogue warp); the three-plane
  // mode has no shared memory left for them and keeps the direct stores
  // Epilogue warps.  128-wide one- / two-plane tiles are used by ONE launch, the layer-10 data gradient (29-deep contraction,
  // 2000 outputs: all epilogue, ~590 instructions per 32-column chunk and warp at 0.3 IPC): it needs 110 registers, so
  // SIXTEEN epilogue warps fit the register file (576 threads x 112) and every warp owns one chunk of a tile.
  static constexpr int EPW = (BLOCK_N == 128 && NPL <= 2 && !PAIR) ? 16 : kEpilogueWarps;
  static constexpr int THREADS = 64 + 32 * EPW;
  static constexpr int STAGING_BYTES = NPL <= 2 ? EPW * kStageTileBytes : 0;
  static constexpr int STAGES_RAW = (kSmemBudget - 1024 - kBarrierBytes - STAGING_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  // [<= 1023 alignment slack][STAGES x stage][barriers, 1 KB][staging tiles]
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + kBarrierBytes + STAGING_BYTES;
  // Split modes keep TWO fp32 accumulators per tile: `main` takes the hi*hi products only, `side` every product
  // that involves a lo plane.  The tensor core truncates its fp32 accumulation, one error of up to an ulp of the
  // ACCUMULATOR per MMA whatever the size of the addend -- with a single accumulator the 2 (or 5) small products per
  // K step cost as much accuracy as the hi*hi product itself (measured: bf16x6 was worse than bf16x3).  `side` stays
  // 2^-8 times smaller, so its truncation is negligible; the epilogue adds the two in round-to-nearest fp32.
  static constexpr int ACC_COLS = (NPL > 1 ? 2 : 1) * BLOCK_N;          // TMEM columns of one accumulator stage
  static constexpr int ACC_STAGES = 2 * ACC_COLS <= 512 ? 2 : 1;        // double-buffered when TMEM allows
  static constexpr int TMEM_RAW = ACC_STAGES * ACC_COLS;
  static constexpr int TMEM_COLS = TMEM_RAW <= 32 ? 32 : (TMEM_RAW <= 64 ? 64 : (TMEM_RAW <= 128 ? 128 : (TMEM_RAW <= 256 ? 256 : 512)));
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
  static_assert(TMEM_RAW <= 512, "TMEM columns");
  static_assert(SMEM_BYTES <= kSmemBudget, "shared memory budget");
  static_assert((STAGES * 2 * 2 + 4) * 8 + 8 <= kBarrierBytes, "barrier area");
};

__device__ __forceinline__ int floordiv(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 lo, __nv_bfloat16 hi) {
  return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}

// Splits 8 fp32 values into NPL planes and stores each plane's 8 bf16 as one 16-byte vector.  Pairs are
// converted with one cvt.rn.bf16x2.f32; the residual for the next plane is x - float(hi) (exact in fp32).
template <int NPL>
__device__ __forceinline__ void store_planes8(__nv_bfloat16* dst, int64_t plane_stride, const float* v) {
  float rem[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) rem[i] = v[i];
#pragma unroll
  for (int p = 0; p < NPL; ++p) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(rem[2 * i], rem[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
      if (p + 1 < NPL) {
        rem[2 * i] -= __uint_as_float(w[i] << 16);
        rem[2 * i + 1] -= __uint_as_float(w[i] & 0xffff0000u);
      }
    }
    *reinterpret_cast<uint4*>(dst + p * plane_stride) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ReLU-mask bits (ConvParams::mask_out layout: word [c / 32][row], bit c % 32) seen from a thread that owns the eight
// channels c8 * 8 .. + 7 of one row: they are byte (c8 & 3) of word c8 >> 2.
__device__ __forceinline__ void store_mask_byte(uint32_t* mask, int64_t rows, int64_t row, int c8, const float* v) {
  uint32_t bits = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) bits |= (uint32_t)(v[i] > 0.f) << i;
  reinterpret_cast<uint8_t*>(mask + (int64_t)(c8 >> 2) * rows + row)[c8 & 3] = (uint8_t)bits;
}
__device__ __forceinline__ uint32_t load_mask_byte(const uint32_t* mask, int64_t rows, int64_t row, int c8) {
  return reinterpret_cast<const uint8_t*>(mask + (int64_t)(c8 >> 2) * rows + row)[c8 & 3];
}

struct Ptr9f { float* p[9]; };
struct Ptr9c { const float* p[9]; };
struct Ptr9h { __nv_bfloat16* p[9]; };

// One 32-column chunk of the epilogue for this lane's output row, from the summed accumulators v[32]:
// + bias, ReLU or ReLU-mask, then
//   * bf16 planes: staged [plane][32 rows][32 cols] in the warp's shared-memory tile and written with ONE TMA store
//     per plane (box {32 ch, 32 t, 1}; rows t >= T' and channel padding beyond ld are clipped by the tensor map), or
//     -- bf16x6 only: no shared memory left -- direct 16-byte stores from the lane that owns the row;
//   * fp32 logits (last layer): direct stores;
//   * bias gradient of the layer below (data gradient): column sums by a 32x32 transpose-reduce;
//   * forward of a ReLU layer: the lane's 32 activity bits for the data gradient (ConvParams::mask_out).
template <int NPL>
__device__ __forceinline__ float epilogue_chunk(const ConvParams& p, const CUtensorMap* tmOut, uint8_t* stage,
                                               float (&v)[32], float bias_lane, uint32_t keep, int nc, int lane,
                                               bool row_ok, int64_t out_row, int t_warp, int b, float* out_f32,
                                               bool f32_add) {
  // bias_lane: bias[nc + lane] (0 beyond N), fetched with one coalesced load per chunk before the accumulator was
  // ready and broadcast by shuffles here -- 32 dependent uniform loads in the epilogue cost ~30 k cycles per tile.
  // keep: bit i set = column nc+i is a real channel AND (data gradient) the ReLU below it was active.
  if (p.bias) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += __shfl_sync(0xffffffffu, bias_lane, i);
  }
  if (p.relu) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = (keep & (1u << i)) ? v[i] : 0.f;
  if (p.mask_out && row_ok && nc < p.N) {
    // ReLU mask of this lane's row for the data gradient: bit i = channel nc + i is active (one coalesced store per warp)
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) bits |= (uint32_t)(v[i] > 0.f) << i;
    p.mask_out[(int64_t)(nc >> 5) * p.mask_rows + out_row] = bits;
  }
  if (p.out_planes) {
    if constexpr (NPL <= 2) {
      // the previous TMA store out of this warp's tile must have finished reading it
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      __nv_bfloat16* srow = reinterpret_cast<__nv_bfloat16*>(stage + lane * 64);
#pragma unroll
      for (int g = 0; g < 4; ++g) store_planes8<NPL>(srow + g * 8, 32 * 32, v + g * 8);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && t_warp < p.To && nc < p.ld_out) {
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl) tma_store_3d(tmOut, stage + pl * (32 * 32 * 2), nc, t_warp, pl * p.B + b);
        bulk_commit();
      }
    } else if (row_ok) {       // three planes: no shared memory left for staging tiles
      __nv_bfloat16* orow = p.out_planes + out_row * p.ld_out + nc;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (nc + g * 8 < p.ld_out) store_planes8<NPL>(orow + g * 8, p.out_plane_stride, v + g * 8);
    }
  }
  if (out_f32 && row_ok) {
    // f32_add: this work item holds a slice of the taps -- accumulate into the pre-zeroed output (two slices add
    // commutatively, so the sum does not depend on which one lands first)
    float* frow = out_f32 + out_row * p.ld_f32 + nc;
    if (nc + 32 <= p.ld_f32 && (p.ld_f32 & 3) == 0) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        if (f32_add) red_add_v4(frow + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
        else *reinterpret_cast<float4*>(frow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (nc + i < p.ld_f32) {
          if (f32_add) atomicAdd(frow + i, v[i]);
          else frow[i] = v[i];
        }
    }
  }
  if (p.col_sum) {
    // bias gradient of the layer below = column sums of what was just stored: 32x32 transpose-reduce with
    // 31 shuffles (each step halves the values a lane holds); the one atomic per column per warp is issued by the
    // caller after the last chunk (an atomic in flight makes the next chunk's proxy fence wait for its round trip)
    if (!row_ok) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int j = 0; j < off; ++j) {
        const float send = upper ? v[j] : v[j + off];
        const float keep = upper ? v[j + off] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    return v[0];            // column nc + lane; the caller adds it to col_sum once the tile's stores are issued
  }
  return 0.f;
}

// Products of the split operands, in issue order, as (A plane, B plane, load group to wait for before it,
// load group released after it; -1 = none).  Plane 0 = hi, 1 = lo, 2 = lo-lo.
//   NPL 1: hi*hi
//   NPL 2: hi*lo, hi*hi, lo*hi                       two load groups per stage, X = {A0,B1}, Y = {A1,B0}
//   NPL 3: the six products above 2^-24: (0,2) (1,1) (2,0) (0,1) (1,0) (0,0), smallest first, one load group
template <int NPL> struct Products;
template <> struct Products<1> {
  static constexpr int N = 1;
  __device__ static constexpr int a(int) { return 0; }
  __device__ static constexpr int b(int) { return 0; }
  __device__ static constexpr int wait(int) { return 0; }
  __device__ static constexpr int release(int) { return 0; }
};
template <> struct Products<2> {
  static constexpr int N = 3;
  __device__ static constexpr int a(int i) { return i == 2 ? 1 : 0; }
  __device__ static constexpr int b(int i) { return i == 0 ? 1 : 0; }
  __device__ static constexpr int wait(int i) { return i == 0 ? 0 : (i == 1 ? 1 : -1); }
  __device__ static constexpr int release(int i) { return i == 1 ? 0 : (i == 2 ? 1 : -1); }
};
template <> struct Products<3> {
  static constexpr int N = 6;
  __device__ static constexpr int a(int i) { return i == 0 ? 0 : (i == 1 ? 1 : (i == 2 ? 2 : (i == 3 ? 0 : (i == 4 ? 1 : 0)))); }
  __device__ static constexpr int b(int i) { return i == 0 ? 2 : (i == 1 ? 1 : (i == 2 ? 0 : (i == 3 ? 1 : 0))); }
  __device__ static constexpr int wait(int i) { return i == 0 ? 0 : -1; }
  __device__ static constexpr int release(int i) { return i == 5 ? 0 : -1; }
};

// EARLY: two-phase epilogue (drain the accumulators to registers, release TMEM, then store) -- chosen by the host
// for launches where a CTA processes several tiles and the split modes leave no second accumulator stage; the
// streaming epilogue (lower register pressure, TMEM loads overlapped with the stores) is used everywhere else.
// 168 registers is the ceiling for 10 warps: each SM sub-partition holds 16384 registers and gets 3 of the warps
// (a 200-register build fails to launch), so the two-phase epilogue's 64 live accumulators spill ~450 bytes.
// NPROB > 1: several problems of identical shape in one launch (ConvParams::n_problems, k_split): a work item is
// (problem, tap slice, tile); fp32 outputs only.  NPROB == 1 is the plain single-problem kernel.
// PAIR: launched as clusters of two CTAs; a work item is a PAIR of m tiles (rank r of the cluster owns m tile
// 2*pair + r), the leader (rank 0) issues tcgen05.mma.cta_group::2 with M = 256 for both.
// BMN: the B operand is MN-major -- the filter is read from its BACKWARD layout [K * Cin rows][ld_co] (output channels
// contiguous, boxes of 64 rows x 64 channels like the filter-gradient operands), so a forward launch needs no packed
// layout of its own (ConvParams::b_row_step = rows per tap; K rows past Cin meet zero channels of A).
template <int BLOCK_N, int NPL, bool EARLY, int NPROB, bool PAIR, bool BMN = false>
__global__ void __launch_bounds__((ConvCfg<BLOCK_N, NPL, PAIR>::THREADS), 1)
tc_conv_kernel(const __grid_constant__ std::conditional_t<NPROB == 1, TmSet1, TmSetN> tmAs,
               const __grid_constant__ std::conditional_t<NPROB == 1, TmSet1, TmSetN> tmBs,
               const __grid_constant__ CUtensorMap tmOut, const ConvParams p) {
  using Cfg = ConvCfg<BLOCK_N, NPL, PAIR>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment in the shared window
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  // Each pipeline stage holds two independently signalled load groups (NPL == 2):
  //   group X = {A_hi, B_lo}, group Y = {A_lo, B_hi};  products are issued hi*lo (X), hi*hi (X+Y), lo*hi (Y),
  // so X is released after the 2nd product and Y after the 3rd: the refill that is on the critical path is 48 KB
  // instead of 96 KB and a two-stage ring keeps the tensor pipe fed.  NPL == 1 uses group X only.
  constexpr int NG = NPL == 2 ? 2 : 1;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);   // [STAGES][NG]
  uint64_t* empty_bar = full_bar + STAGES * NG;                                         // [STAGES][NG]
  uint64_t* tmem_full = empty_bar + STAGES * NG;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;           // position in the CTA pair; rank 0 leads
  const int cta_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;        // work-loop start
  const int cta_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;        // work-loop stride
  const int m_tiles_real = p.B * p.m_tiles_per_utt;
  const int m_tiles = PAIR ? (m_tiles_real + 1) >> 1 : m_tiles_real;         // PAIR: pairs of m tiles
  const int num_tiles = m_tiles * p.n_tiles;
  const int ksplit = NPROB > 1 ? max(1, p.k_split) : 1;
  const int total_work = NPROB > 1 ? num_tiles * ksplit * p.n_problems : num_tiles;
  // work item -> (problem q, taps [j0, j1), tile); single-problem launches: work item == tile, all taps
  auto work_coords = [&](int work, int& q, int& j0, int& j1, int& tile) {
    if constexpr (NPROB > 1) {
      const int per_prob = num_tiles * ksplit;
      q = work / per_prob;
      const int r = work - q * per_prob;
      const int ks = r / num_tiles;
      tile = r - ks * num_tiles;
      j0 = p.taps * ks / ksplit;
      j1 = p.taps * (ks + 1) / ksplit;
    } else {
      q = 0; j0 = 0; j1 = p.taps; tile = work;
    }
  };
  // debug stamps of the first tile (streaming-epilogue instantiations only: the two-phase one has no register to spare)
  long long* tl = (!EARLY && p.timeline) ? p.timeline + (int64_t)blockIdx.x * 8 : nullptr;
  if (tl && threadIdx.x == 0) tl[0] = global_ns();                                   // 0: kernel entry

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmAs.m[0]);
    prefetch_tmap(&tmBs.m[0]);
    if (p.tma_store) prefetch_tmap(&tmOut);
    for (int s = 0; s < STAGES * NG; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full + a, 1);
      // PAIR: the leader's MMA thread waits for the epilogue warps of BOTH CTAs (the peer's arrive remotely)
      mbar_init(tmem_empty + a, Cfg::EPW * (PAIR ? 2 : 1));
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();         // barriers of both CTAs initialised before any remote arrive / TMA
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) touched no
  // global data, so it may overlap the tail of the previous kernel in the stream.  Let OUR dependents start
  // launching, then wait until the previous grid has completed and its writes are visible.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (tl && threadIdx.x == 0) tl[1] = global_ns();                                   // 1: previous grid complete
  // tile -> (b, t0, n0) of THIS CTA; PAIR: rank r owns m tile 2*pair + r, which may not exist when the count is odd:
  // its loads then aim at rows far outside the tensor (TMA zero fill) and its epilogue stores nothing
  constexpr int kNoRow = 1 << 20;
  auto cta_tile = [&](int tile, int& b, int& t0, int& n0) {
    const int nt = p.n_fastest ? tile % p.n_tiles : tile / m_tiles;
    int mt = p.n_fastest ? tile / p.n_tiles : tile % m_tiles;
    if constexpr (PAIR) mt = 2 * mt + (int)rank;
    b = mt / p.m_tiles_per_utt;
    t0 = (mt - b * p.m_tiles_per_utt) * kTileM;
    n0 = nt * BLOCK_N;
    if (PAIR && mt >= m_tiles_real) { b = 0; t0 = kNoRow; }
  };

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = cta_id; work < total_work; work += cta_step) {
        int q, j0, j1, tile;
        work_coords(work, q, j0, j1, tile);
        const CUtensorMap* tmA = &tmAs.m[NPROB > 1 ? q : 0];
        const CUtensorMap* tmB = &tmBs.m[NPROB > 1 ? q : 0];
        const int pad_left = NPROB > 1 ? p.pad_left_q[q] : p.pad_left;
        // n fastest (A rows read from HBM once, shared through L2 by their n tiles) when A is too big for L2;
        // m fastest (one filter slab hot in L2 for the whole wave) otherwise
        int b, t0, n0;
        cta_tile(tile, b, t0, n0);
        // PAIR: this CTA holds B rows [n0 + rank * n_mma / 2, ...): the halves of the N the instruction multiplies
        int b_half = 0;
        if constexpr (PAIR) {
          const int n_valid = min(BLOCK_N, p.N - n0);
          const int n_mma = p.trim ? max(16, (n_valid + 15) & ~15) : BLOCK_N;
          b_half = (int)rank * (n_mma >> 1);
        }
        const int jn = j1 - j0, nk = jn * p.chunks_per_tap;
        for (int it = 0; it < nk; ++it) {
          // channel chunk outer, filter tap inner: the taps of one chunk re-read (shifted) the same A rows, which
          // stay in L2, instead of sweeping the whole channel range once per tap
          const int cc = it / jn, j = j0 + it - cc * jn;
          const int m = p.a_sign * (j - pad_left);
          const int shift = p.a_stride == 1 ? m : floordiv(m, p.a_stride);
          const int a_col = (m - shift * p.a_stride) * p.a_cin + cc * kChunkK;
          uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
          // K-major B: box {64 K columns, B_ROWS rows};  MN-major B: B_ROWS / 64 boxes {64 N columns, 64 K rows}
          const int b_c0 = BMN ? n0 + b_half : j * p.b_col_step + cc * kChunkK;
          const int b_c1 = BMN ? j * p.b_row_step + cc * kChunkK : n0 + j * p.b_row_step + b_half;
          auto load_b = [&](uint64_t* fb, uint32_t fbc, int pb) {
            uint8_t* dst = st + NPL * Cfg::A_BYTES + pb * Cfg::B_BYTES;
            if constexpr (BMN) {
#pragma unroll
              for (int h = 0; h < Cfg::B_ROWS / 64; ++h) {
                if constexpr (PAIR) tma_load_2d_pair(tmB, fbc, dst + h * 8192, b_c0 + h * 64, pb * p.b_plane_rows + b_c1);
                else tma_load_2d(tmB, fb, dst + h * 8192, b_c0 + h * 64, pb * p.b_plane_rows + b_c1);
              }
            } else {
              if constexpr (PAIR) tma_load_2d_pair(tmB, fbc, dst, b_c0, pb * p.b_plane_rows + b_c1);
              else tma_load_2d(tmB, fb, dst, b_c0, pb * p.b_plane_rows + b_c1);
            }
          };
          if (NG == 2) {
#pragma unroll
            for (int g = 0; g < NG; ++g) {
              // group 0 (X): A plane 0 + B plane 1;  group 1 (Y): A plane 1 + B plane 0
              const int pa = g, pb = 1 - g;
              uint64_t* fb = full_bar + stage * NG + g;
              mbar_wait(empty_bar + stage * NG + g, phase ^ 1);
              if constexpr (PAIR) {
                // both CTAs' bytes complete on the LEADER's barrier; the leader alone arms it
                if (rank == 0) mbar_expect_tx(fb, 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
                const uint32_t fbc = mapa_u32(smem_u32(fb), 0);
                tma_load_3d_pair(tmA, fbc, st + pa * Cfg::A_BYTES, a_col, t0 + shift, pa * p.B + b);
                load_b(fb, fbc, pb);
              } else {
                mbar_expect_tx(fb, Cfg::A_BYTES + Cfg::B_BYTES);
                tma_load_3d(tmA, fb, st + pa * Cfg::A_BYTES, a_col, t0 + shift, pa * p.B + b);
                load_b(fb, 0u, pb);
              }
            }
          } else {
            uint64_t* fb = full_bar + stage;
            mbar_wait(empty_bar + stage, phase ^ 1);
            if constexpr (PAIR) {
              if (rank == 0) mbar_expect_tx(fb, 2 * Cfg::STAGE_BYTES);
              const uint32_t fbc = mapa_u32(smem_u32(fb), 0);
#pragma unroll
              for (int pl = 0; pl < NPL; ++pl)
                tma_load_3d_pair(tmA, fbc, st + pl * Cfg::A_BYTES, a_col, t0 + shift, pl * p.B + b);
#pragma unroll
              for (int pl = 0; pl < NPL; ++pl) load_b(fb, fbc, pl);
            } else {
              mbar_expect_tx(fb, Cfg::STAGE_BYTES);
#pragma unroll
              for (int pl = 0; pl < NPL; ++pl)
                tma_load_3d(tmA, fb, st + pl * Cfg::A_BYTES, a_col, t0 + shift, pl * p.B + b);
#pragma unroll
              for (int pl = 0; pl < NPL; ++pl) load_b(fb, 0u, pl);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one thread; PAIR: of the leader CTA only)
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int work = cta_id; work < total_work; work += cta_step, ++local) {
        int q, j0, j1, tile;
        work_coords(work, q, j0, j1, tile);
        const int acc = local % Cfg::ACC_STAGES;
        const uint32_t acc_phase = (local / Cfg::ACC_STAGES) & 1;
        // Only real channels are multiplied: the N of the instruction is the tile's channel count rounded up to 16
        // (2000 output channels = 7 tiles of 256 and one of 208), and K steps that would read only TMA zero fill
        // (the last 64-wide chunk of a 2000-channel contraction holds 16 channels) are not issued.  The big layers
        // are power-bound, so every MMA that is not executed is time.
        const int nt = p.n_fastest ? tile % p.n_tiles : tile / m_tiles;
        const int n_valid = min(BLOCK_N, p.N - nt * BLOCK_N);
        const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * kTileM : kTileM,
                                               p.trim ? max(16, (n_valid + 15) & ~15) : BLOCK_N, 0, BMN ? 1 : 0);
        mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_main = tmem_base + acc * Cfg::ACC_COLS;
        const uint32_t d_side = d_main + BLOCK_N;
        // one pipeline iteration = the products of one 64-deep K chunk; FULL: all four K steps (unrolled, the hot
        // path), otherwise only the first nkk (runtime loop) -- the issuing thread is close to critical (a division
        // and four extra branches per iteration cost 3 % of the step), so the trimmed path must not tax the full one
        auto iteration = [&](auto full_tag, int it, int nkk) {
          constexpr bool FULL = decltype(full_tag)::value;
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + NPL * Cfg::A_BYTES;
          uint32_t acc_main = it > 0 ? 1u : 0u, acc_side = acc_main;
          using PR = Products<NPL>;
#pragma unroll
          for (int pr = 0; pr < PR::N; ++pr) {
            const int pa = PR::a(pr), pb = PR::b(pr);
            if (PR::wait(pr) >= 0) { mbar_wait(full_bar + stage * NG + PR::wait(pr), phase); tc_fence_after(); }
            if (tl && local == 0 && it == 0 && pr == 0) tl[2] = global_ns();         // 2: first operands landed
            const uint64_t da = make_smem_desc_sw128(a_addr + pa * Cfg::A_BYTES, 16, 1024);
            // MN-major B (see tc_wgrad_kernel): LBO = distance between the 64-channel boxes, a K step of 16 rows = two
            // swizzle atoms of 8 rows x 128 B = 2048 bytes = +128 in the address field
            const uint64_t db = BMN ? make_smem_desc_sw128(b_addr + pb * Cfg::B_BYTES, 8192, 1024)
                                    : make_smem_desc_sw128(b_addr + pb * Cfg::B_BYTES, 16, 1024);
            constexpr int kBStep = BMN ? 128 : 2;
            auto step = [&](int kk) {
              // advancing 16 bf16 along K = 32 bytes inside the 128-byte swizzled row = +2 in the address field
              const uint32_t d = (pa == 0 && pb == 0) ? d_main : d_side;
              uint32_t& accf = (pa == 0 && pb == 0) ? acc_main : acc_side;
              if constexpr (PAIR) umma_bf16_pair(d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * kBStep), idesc, accf);
              else umma_bf16(d, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * kBStep), idesc, accf);
              accf = 1u;
            };
            if constexpr (FULL) {
#pragma unroll
              for (int kk = 0; kk < kChunkK / 16; ++kk) step(kk);
            } else {
#pragma unroll 1
              for (int kk = 0; kk < nkk; ++kk) step(kk);
            }
            // a group's smem is reusable once the MMAs issued so far have read it (PAIR: in both CTAs)
            if (PR::release(pr) >= 0) {
              if constexpr (PAIR) umma_commit_pair(empty_bar + stage * NG + PR::release(pr), 3);
              else umma_commit(empty_bar + stage * NG + PR::release(pr));
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        };
        // channel chunk outer, tap inner (the producer's order): the number of K steps only depends on the chunk
        int it = 0;
        for (int cc = 0; cc < p.chunks_per_tap; ++cc) {
          const int nkk = p.trim ? min(kChunkK / 16, (p.k_cols - cc * kChunkK + 15) >> 4) : kChunkK / 16;
          if (nkk == kChunkK / 16) {
            for (int j = j0; j < j1; ++j, ++it) iteration(std::true_type{}, it, nkk);
          } else {
            for (int j = j0; j < j1; ++j, ++it) iteration(std::false_type{}, it, nkk);
          }
        }
        if constexpr (PAIR) umma_commit_pair(tmem_full + acc, 3);   // accumulators complete -> both epilogues
        else umma_commit(tmem_full + acc);                // accumulator complete -> epilogue
        if (tl && local == 0) tl[3] = global_ns();                                   // 3: last MMA of the tile issued
      }
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue warps 2 .. 2 + EPW - 1
    const int quarter = warp & 3;                         // TMEM lane quarter this warp may read
    const int chunk0 = (warp - 2) >> 2;                   // warps w, w+4, ... share a quarter and split its chunks
    constexpr int kChunkStep = Cfg::EPW / 4;
    constexpr int kChunks = BLOCK_N / 32;
    const int row = quarter * 32 + lane;
    uint8_t* stage = smem + STAGES * Cfg::STAGE_BYTES + kBarrierBytes + (warp - 2) * kStageTileBytes;
    auto tile_coords = [&](int tile, int& b, int& t0, int& n0) { cta_tile(tile, b, t0, n0); };
    // accumulator stage drained: PAIR arrives on the LEADER's barrier (its MMA thread feeds both TMEMs)
    auto release_acc = [&](int acc) {
      if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(tmem_empty + acc), 0));
      else mbar_arrive(tmem_empty + acc);
    };
    int local = 0;
    for (int work = cta_id; work < total_work; work += cta_step, ++local) {
      int q, j0, j1, tile;
      work_coords(work, q, j0, j1, tile);
      const int acc = local % Cfg::ACC_STAGES;
      const uint32_t acc_phase = (local / Cfg::ACC_STAGES) & 1;
      int b, t0, n0;
      tile_coords(tile, b, t0, n0);
      const int t = t0 + row;
      const int t_warp = t0 + quarter * 32;               // first row of this warp's 32-row slab
      const bool row_ok = t < p.To;
      const int64_t out_row = (int64_t)b * p.To + t;
      float* out_f32 = NPROB > 1 ? p.out_f32_q[q] : p.out_f32;
      const bool f32_add = NPROB > 1 && ksplit > 1;
      // Per chunk of this warp, fetched while the MMAs of the tile are still running: the lane's bias element and
      // (data gradient) the row's word of ReLU-mask bits -- one keep-bit per column.
      constexpr int kMine = (kChunks + kChunkStep - 1) / kChunkStep;       // chunks per warp (compile time)
      float bias_l[kMine];
      uint32_t keep[kMine];
#pragma unroll
      for (int ci = 0; ci < kMine; ++ci) {
        const int nc = n0 + (chunk0 + ci * kChunkStep) * 32;
        bias_l[ci] = (p.bias && nc + lane < p.N) ? __ldg(p.bias + nc + lane) : 0.f;
        const int nvalid = p.N - nc;
        keep[ci] = nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
        if (p.mask_bits && nvalid > 0)
          keep[ci] &= row_ok ? __ldg(p.mask_bits + (int64_t)(nc >> 5) * p.mask_rows + out_row) : 0u;
      }
      mbar_wait(tmem_full + acc, acc_phase);
      tc_fence_after();
      if (tl && local == 0 && warp == 2 && lane == 0) tl[4] = global_ns();           // 4: accumulator complete
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * Cfg::ACC_COLS;
      float csum[kMine];
#pragma unroll
      for (int ci = 0; ci < kMine; ++ci) csum[ci] = 0.f;
      if constexpr (EARLY) {
        // Phase 1: drain this warp's share of the accumulators into registers (main + side summed in round-to-nearest
        // fp32) and hand TMEM back to the MMA warp at once -- with the two 256-column accumulators of the split modes
        // there is no second accumulator stage, so everything after this point overlaps the next tile's MMAs.
        float sum[kMine][32];
#pragma unroll
        for (int ci = 0; ci < kMine; ++ci) {
          const int c = chunk0 + ci * kChunkStep;
          if (c < kChunks) {
            uint32_t r[32];
            tmem_ld32(taddr + c * 32, r);
            if (NPL > 1) {
              uint32_t q[32];
              tmem_ld32(taddr + BLOCK_N + c * 32, q);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) sum[ci][i] = __uint_as_float(r[i]) + __uint_as_float(q[i]);
            } else {
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) sum[ci][i] = __uint_as_float(r[i]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(acc);

        // Phase 2: bias / ReLU / ReLU-mask / plane split / stores / bias-gradient column sums, from registers
#pragma unroll
        for (int ci = 0; ci < kMine; ++ci) {
          const int c = chunk0 + ci * kChunkStep;
          if (c >= kChunks) continue;
          csum[ci] = epilogue_chunk<NPL>(p, &tmOut, stage, sum[ci], bias_l[ci], keep[ci], n0 + c * 32, lane, row_ok,
                                         out_row, t_warp, b, out_f32, f32_add);
        }
      } else {
        // A real loop, not four unrolled copies: one chunk is ~500 instructions, and four copies per tile did not
        // fit the instruction cache (stall_no_inst all over the epilogue in ncu's source view).  The per-chunk
        // scalars are picked out of / put back into their registers with selects.
        static_assert(kMine <= 4, "chunk scalars are selected by hand");
#pragma unroll 1
        for (int ci = 0; ci < kMine; ++ci) {
          const int c = chunk0 + ci * kChunkStep;
          if (c >= kChunks) break;
          const float bias_c = ci == 0 ? bias_l[0] : (ci == 1 ? bias_l[1 % kMine] : (ci == 2 ? bias_l[2 % kMine] : bias_l[3 % kMine]));
          const uint32_t keep_c = ci == 0 ? keep[0] : (ci == 1 ? keep[1 % kMine] : (ci == 2 ? keep[2 % kMine] : keep[3 % kMine]));
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          uint32_t q[32];
          if (NPL > 1) tmem_ld32(taddr + BLOCK_N + c * 32, q);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = __uint_as_float(r[i]);
            if (NPL > 1) v[i] += __uint_as_float(q[i]);        // main + side accumulator, round-to-nearest
          }
          const float cs = epilogue_chunk<NPL>(p, &tmOut, stage, v, bias_c, keep_c, n0 + c * 32, lane, row_ok, out_row,
                                               t_warp, b, out_f32, f32_add);
#pragma unroll
          for (int k = 0; k < kMine; ++k)
            if (ci == k) csum[k] = cs;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(acc);
      }
      if (p.col_sum) {
#pragma unroll
        for (int ci = 0; ci < kMine; ++ci) {
          const int n = n0 + (chunk0 + ci * kChunkStep) * 32 + lane;
          if (chunk0 + ci * kChunkStep < kChunks && n < p.N) atomicAdd(p.col_sum + n, csum[ci]);
        }
      }
      if (tl && local == 0 && warp == 2 && lane == 0) tl[5] = global_ns();           // 5: warp 2's epilogue issued
    }
    // outstanding TMA stores of this warp must have finished reading the staging tile before the CTA (and its shared
    // memory) goes away; their global writes are complete when the grid is
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
    if (tl && warp == 2 && lane == 0) tl[6] = global_ns();                           // 6: staging tiles drained
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();         // neither CTA leaves (or frees TMEM) while its peer still works
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
  if (tl && threadIdx.x == 0) tl[7] = global_ns();                                   // 7: exit
}

// ------------------------------------------------------------------------------------------------ filter gradient
// PAIR: two CTAs of one cluster work on two tiles that share their dZ (B) tile -- consecutive tiles differ in the
// filter tap or the Cin tile only -- as ONE tcgen05.mma.cta_group::2 of M = 256: each CTA holds its own X tile and
// HALF of the dZ columns, so a stage is a third smaller (three stages instead of two in bf16x3) and dZ crosses
// L2 -> shared memory once per pair.  These launches sit at the L2 -> SM bandwidth cap (96 KB per 12 MMAs and SM).
template <int BLOCK_N, int NPL, bool PAIR = false>
struct WgradCfg {
  static constexpr int A_BYTES = kTileM * kChunkK * 2;           // 2 boxes of [64 rows][64 ch]
  static constexpr int B_COLS = PAIR ? BLOCK_N / 2 : BLOCK_N;    // dZ columns (output channels) held by one CTA
  static constexpr int B_BYTES = B_COLS * kChunkK * 2;           // B_COLS/64 boxes
  static constexpr int STAGE_BYTES = NPL * (A_BYTES + B_BYTES);
  static constexpr int STAGES_RAW = (kSmemBudget - 2048) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2048;
  // Split modes keep TWO fp32 accumulators per tile: `main` takes the hi*hi products only, `side` every product
  // that involves a lo plane.  The tensor core truncates its fp32 accumulation, one error of up to an ulp of the
  // ACCUMULATOR per MMA whatever the size of the addend -- with a single accumulator the 2 (or 5) small products per
  // K step cost as much accuracy as the hi*hi product itself (measured: bf16x6 was worse than bf16x3).  `side` stays
  // 2^-8 times smaller, so its truncation is negligible; the epilogue adds the two in round-to-nearest fp32.
  static constexpr int ACC_COLS = (NPL > 1 ? 2 : 1) * BLOCK_N;          // TMEM columns of one accumulator stage
  static constexpr int ACC_STAGES = 2 * ACC_COLS <= 512 ? 2 : 1;        // double-buffered when TMEM allows
  static constexpr int TMEM_RAW = ACC_STAGES * ACC_COLS;
  static constexpr int TMEM_COLS = TMEM_RAW <= 32 ? 32 : (TMEM_RAW <= 64 ? 64 : (TMEM_RAW <= 128 ? 128 : (TMEM_RAW <= 256 ? 256 : 512)));
  static constexpr int BOX_BYTES = 64 * 128;                     // one {64 ch, 64 rows} box
};

template <int BLOCK_N, int NPL, int NPROB, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
tc_wgrad_kernel(const __grid_constant__ std::conditional_t<NPROB == 1, TmSet1, TmSetN> tmXs,
                const __grid_constant__ std::conditional_t<NPROB == 1, TmSet1, TmSetN> tmDZs,
                const WgradParams p) {
  using Cfg = WgradCfg<BLOCK_N, NPL, PAIR>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  constexpr int NG = NPL == 2 ? 2 : 1;              // load groups per stage, see tc_conv_kernel
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES * NG;
  uint64_t* tmem_full = empty_bar + STAGES * NG;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Wave-aligned split-K.  Whole waves of gridDim.x tiles run data-parallel (one tile per CTA, all CTAs sweep the
  // contraction index in phase, so tiles that share an operand slab hit it in L2 at the same time); the remaining
  // R = tiles % gridDim.x tiles are cut into S = gridDim.x / R aligned K slices each so that the last wave also
  // fills the machine.  Sliced tiles accumulate with fp32 atomics into the pre-zeroed gradient, whole tiles store.
  // PAIR: the scheduling unit is a pair of consecutive tiles (2u, 2u + 1) -- same problem and n tile, the host checks
  // that taps * m_tiles is even -- handled by one cluster; rank r of the cluster owns tile 2u + r.
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int unit_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int tiles_per_problem = p.taps * tiles_mn;
  const int num_tiles = ((NPROB > 1 ? p.n_problems : 1) * tiles_per_problem) >> (PAIR ? 1 : 0);    // scheduling units
  const int total_iters = p.B * p.t_chunks;
  const int G = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int full_waves = num_tiles / G;
  const int tail_tiles = num_tiles - full_waves * G;
  const int tail_split = tail_tiles > 0 ? max(1, min(G / tail_tiles, total_iters / 4 > 0 ? total_iters / 4 : 1)) : 1;
  const int fsplit = p.force_split > 1 ? p.force_split : 0;
  const int forced_items = fsplit * num_tiles;
  const int my_items = fsplit ? (unit_id < forced_items ? (forced_items - unit_id + G - 1) / G : 0)
                              : full_waves + (unit_id < tail_tiles * tail_split ? 1 : 0);
  // item i of this CTA -> (tile of THIS CTA, q0, q1)
  auto item = [&](int i, int& tile, int& q0, int& q1) {
    if (fsplit) {
      const int g = i * G + unit_id;                       // slice-major: a wave sweeps one K range in phase
      const int slice = g / num_tiles;
      tile = g - slice * num_tiles;
      q0 = (int)((int64_t)total_iters * slice / fsplit);
      q1 = (int)((int64_t)total_iters * (slice + 1) / fsplit);
    } else if (i < full_waves) {
      tile = i * G + unit_id;
      q0 = 0;
      q1 = total_iters;
    } else {
      const int slice = unit_id / tail_tiles;             // CTAs of one slice are consecutive: same K range in phase
      tile = full_waves * G + unit_id % tail_tiles;
      q0 = (int)((int64_t)total_iters * slice / tail_split);
      q1 = (int)((int64_t)total_iters * (slice + 1) / tail_split);
    }
    if constexpr (PAIR) tile = 2 * tile + (int)rank;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmXs.m[0]);
    prefetch_tmap(&tmDZs.m[0]);
    for (int s = 0; s < STAGES * NG; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full + a, 1);
      // PAIR: the leader's MMA thread waits for the epilogue warps of BOTH CTAs (the peer's arrive remotely)
      mbar_init(tmem_empty + a, kEpilogueWarps * (PAIR ? 2 : 1));
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();         // barriers of both CTAs initialised before any remote arrive / TMA
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) touched no
  // global data, so it may overlap the tail of the previous kernel in the stream.  Let OUR dependents start
  // launching, then wait until the previous grid has completed and its writes are visible.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // tile -> (problem, n tile, tap, m tile), n slowest within a problem: a wave touches few dZ column slabs
  auto decode = [&](int tile, int& q, int& j, int& mt, int& nt) {
    q = NPROB > 1 ? tile / tiles_per_problem : 0;
    tile -= q * tiles_per_problem;
    const int per_n = p.taps * p.m_tiles;
    nt = tile / per_n;
    const int rem = tile - nt * per_n;
    j = rem / p.m_tiles;
    mt = rem - j * p.m_tiles;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_items; ++i) {
        int tile, q0, q1, pq, j, mt, nt;
        item(i, tile, q0, q1);
        decode(tile, pq, j, mt, nt);
        const CUtensorMap* tmX = &tmXs.m[NPROB > 1 ? pq : 0];
        const CUtensorMap* tmDZ = &tmDZs.m[NPROB > 1 ? pq : 0];
        const int m = j - (NPROB > 1 ? p.pad_left_q[pq] : p.pad_left);
        const int shift = p.a_stride == 1 ? m : floordiv(m, p.a_stride);
        const int a_col = (m - shift * p.a_stride) * p.a_cin + mt * kTileM;
        int n0 = nt * BLOCK_N;
        if constexpr (PAIR) {
          // this CTA holds dZ columns [n0 + rank * n_mma / 2, ...): the halves of the N the instruction multiplies
          const int n_valid = min(BLOCK_N, p.Cout - n0);
          const int n_mma = p.trim ? max(16, (n_valid + 15) & ~15) : BLOCK_N;
          n0 += (int)rank * (n_mma >> 1);
        }
        for (int q = q0; q < q1; ++q) {
          const int b = q / p.t_chunks;
          const int t0 = (q - b * p.t_chunks) * kChunkK;
          uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            // NG == 2: group X = {X_hi, dZ_lo}, group Y = {X_lo, dZ_hi};  NG == 1: every plane in one group
            uint64_t* fb = full_bar + stage * NG + g;
            mbar_wait(empty_bar + stage * NG + g, phase ^ 1);
            constexpr uint32_t kBytes = NG == 2 ? Cfg::A_BYTES + Cfg::B_BYTES : Cfg::STAGE_BYTES;
            uint32_t fbc = 0;
            if constexpr (PAIR) {
              // both CTAs' bytes complete on the LEADER's barrier; the leader alone arms it
              if (rank == 0) mbar_expect_tx(fb, 2 * kBytes);
              fbc = mapa_u32(smem_u32(fb), 0);
            } else {
              mbar_expect_tx(fb, kBytes);
            }
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) {
              if (NG == 2 && pl != g) continue;
#pragma unroll
              for (int h = 0; h < kTileM / 64; ++h) {
                uint8_t* dst = st + pl * Cfg::A_BYTES + h * Cfg::BOX_BYTES;
                if constexpr (PAIR) tma_load_3d_pair(tmX, fbc, dst, a_col + h * 64, t0 + shift, pl * p.B + b);
                else tma_load_3d(tmX, fb, dst, a_col + h * 64, t0 + shift, pl * p.B + b);
              }
            }
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) {
              if (NG == 2 && pl != 1 - g) continue;
#pragma unroll
              for (int h = 0; h < Cfg::B_COLS / 64; ++h) {
                uint8_t* dst = st + NPL * Cfg::A_BYTES + pl * Cfg::B_BYTES + h * Cfg::BOX_BYTES;
                if constexpr (PAIR) tma_load_3d_pair(tmDZ, fbc, dst, n0 + h * 64, t0, pl * p.B + b);
                else tma_load_3d(tmDZ, fb, dst, n0 + h * 64, t0, pl * p.B + b);
              }
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {                    // PAIR: the leader issues for both CTAs
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (; local < my_items; ++local) {
        int tile, q0, q1, tq, tj, tmt, tnt;
        item(local, tile, q0, q1);
        decode(tile, tq, tj, tmt, tnt);
        // both operands MN-major; N trimmed to the real output channels of this n tile (rounded up to 16)
        const int n_valid = min(BLOCK_N, p.Cout - tnt * BLOCK_N);
        const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * kTileM : kTileM,
                                               p.trim ? max(16, (n_valid + 15) & ~15) : BLOCK_N, 1, 1);
        const int acc = local % Cfg::ACC_STAGES;
        const uint32_t acc_phase = (local / Cfg::ACC_STAGES) & 1;
        mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_main = tmem_base + acc * Cfg::ACC_COLS;
        const uint32_t d_side = d_main + BLOCK_N;
        uint32_t acc_main = 0u, acc_side = 0u;
        auto iteration = [&](auto full_tag, int nkk) {
          constexpr bool FULL = decltype(full_tag)::value;
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + NPL * Cfg::A_BYTES;
          using PR = Products<NPL>;
#pragma unroll
          for (int pr = 0; pr < PR::N; ++pr) {
            const int pa = PR::a(pr), pb = PR::b(pr);
            if (PR::wait(pr) >= 0) { mbar_wait(full_bar + stage * NG + PR::wait(pr), phase); tc_fence_after(); }
            auto step = [&](int kk) {
              // MN-major SW128: a K step of 16 rows = 2 swizzle atoms of 8 rows x 128 B = 2048 bytes;
              // LBO = distance between 64-wide MN blocks (one TMA box), SBO = 1024 (next 8 K rows)
              const uint64_t da = make_smem_desc_sw128(a_addr + pa * Cfg::A_BYTES + kk * 2048, Cfg::BOX_BYTES, 1024);
              const uint64_t db = make_smem_desc_sw128(b_addr + pb * Cfg::B_BYTES + kk * 2048, Cfg::BOX_BYTES, 1024);
              if (pa == 0 && pb == 0) {
                if constexpr (PAIR) umma_bf16_pair(d_main, da, db, idesc, acc_main);
                else umma_bf16(d_main, da, db, idesc, acc_main);
                acc_main = 1u;
              } else {
                if constexpr (PAIR) umma_bf16_pair(d_side, da, db, idesc, acc_side);
                else umma_bf16(d_side, da, db, idesc, acc_side);
                acc_side = 1u;
              }
            };
            if constexpr (FULL) {
#pragma unroll
              for (int kk = 0; kk < kChunkK / 16; ++kk) step(kk);
            } else {
#pragma unroll 1
              for (int kk = 0; kk < nkk; ++kk) step(kk);
            }
            if (PR::release(pr) >= 0) {
              if constexpr (PAIR) umma_commit_pair(empty_bar + stage * NG + PR::release(pr), 3);
              else umma_commit(empty_bar + stage * NG + PR::release(pr));
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        };
        // K = time: rows past the utterance's last logit frame are TMA zero fill in dZ, their K steps are skipped
        // (only the last chunk of an utterance can be short)
        const int last_nkk = p.trim ? min(kChunkK / 16, (p.To - (p.t_chunks - 1) * kChunkK + 15) >> 4) : kChunkK / 16;
        int tc = q0 - (q0 / p.t_chunks) * p.t_chunks;
        for (int q = q0; q < q1; ++q) {
          if (tc == p.t_chunks - 1 && last_nkk != kChunkK / 16) iteration(std::false_type{}, last_nkk);
          else iteration(std::true_type{}, kChunkK / 16);
          if (++tc == p.t_chunks) tc = 0;
        }
        if constexpr (PAIR) umma_commit_pair(tmem_full + acc, 3);   // accumulators complete -> both epilogues
        else umma_commit(tmem_full + acc);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int chunk0 = (warp - 2) >> 2;
    constexpr int kChunkStep = kEpilogueWarps / 4;
    const int row = quarter * 32 + lane;
    int local = 0;
    for (; local < my_items; ++local) {
      int tile, q0, q1, pq, j, mt, nt;
      item(local, tile, q0, q1);
      decode(tile, pq, j, mt, nt);
      const bool whole_tile = q0 == 0 && q1 == total_iters;
      const int acc = local % Cfg::ACC_STAGES;
      const uint32_t acc_phase = (local / Cfg::ACC_STAGES) & 1;
      const int ci = mt * kTileM + row;
      const int n0 = nt * BLOCK_N;
      mbar_wait(tmem_full + acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * Cfg::ACC_COLS;
      float* wrow = NPROB > 1 ? p.dW_q[pq] + ((int64_t)j * p.tap_stride_q[pq] * p.Cin + ci) * p.Cout
                              : p.dW + ((int64_t)j * p.Cin + ci) * p.Cout;
#pragma unroll 1
      for (int c = chunk0; c < BLOCK_N / 32; c += kChunkStep) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        uint32_t q[32];
        if (NPL > 1) tmem_ld32(taddr + BLOCK_N + c * 32, q);
        tmem_ld_wait();
        if (ci < p.Cin) {
          const int cb = n0 + c * 32;
          float x[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(r[i]) + (NPL > 1 ? __uint_as_float(q[i]) : 0.f);
          // a lane owns one filter row (ci) and 32 consecutive output channels of it: they leave as 16- or 8-byte
          // vectors when the row pitch keeps them aligned (Cout % 4 / % 2) -- plain stores for whole tiles, vector
          // reductions (red.global.add.v4/.v2.f32) for K-sliced tiles -- and as scalars otherwise
          if (cb + 32 <= p.Cout && (p.Cout & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (whole_tile) *reinterpret_cast<float4*>(wrow + cb + i) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
              else red_add_v4(wrow + cb + i, x[i], x[i + 1], x[i + 2], x[i + 3]);
            }
          } else if (cb + 32 <= p.Cout && (p.Cout & 1) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              if (whole_tile) *reinterpret_cast<float2*>(wrow + cb + i) = make_float2(x[i], x[i + 1]);
              else red_add_v2(wrow + cb + i, x[i], x[i + 1]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int co = cb + i;
              if (co < p.Cout) {
                if (whole_tile) wrow[co] = x[i];
                else atomicAdd(wrow + co, x[i]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        // accumulator stage drained: PAIR arrives on the LEADER's barrier (its MMA thread feeds both TMEMs)
        if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(tmem_empty + acc), 0));
        else mbar_arrive(tmem_empty + acc);
      }
    }
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();         // neither CTA leaves (or frees TMEM) while its peer still works
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ small kernels
template <int NPL>
__global__ void __launch_bounds__(256)
split_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ planes, int B, int T, int Tpad, int F) {
  // one thread per 8 consecutive features
  const int64_t groups = (int64_t)B * Tpad * (F / 8);
  const int64_t plane_stride = (int64_t)B * Tpad * F;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    const int f8 = (int)(g % (F / 8));
    const int64_t bt = g / (F / 8);
    const int t = (int)(bt % Tpad);
    const int b = (int)(bt / Tpad);
    float v[8];
    if (t < T) {
      const float4* src = reinterpret_cast<const float4*>(x + ((int64_t)b * T + t) * F + f8 * 8);
      const float4 a = __ldg(src), c = __ldg(src + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    store_planes8<NPL>(planes + ((int64_t)b * Tpad + t) * F + f8 * 8, plane_stride, v);
  }
}

// ---- filter packing for ALL layers in one launch (PackTable lists the layers) ----
// Both layouts in ONE pass over W: a block owns a 64(ci) x 64(co) tile of one tap, stages it in shared memory and
// writes it out twice -- transposed for the forward layout (warp = one co row, lanes along ci, 128-byte bf16x2
// rows) and as is for the backward layout (warp = one (k,ci) row, lanes along co).  Padding (ci >= Cin in the
// forward rows, co >= Cout in the backward rows) is written as zeros every time: the arena is shared between shapes.
template <int NPL>
__global__ void __launch_bounds__(256)
pack_filter_both_kernel(const PackTable tab) {
  __shared__ float tile[64][65];
  int l = 0;
  while (l + 1 < tab.n && (int)blockIdx.x >= tab.e[l + 1].blk0) ++l;
  const PackEntry& e = tab.e[l];
  int lb = blockIdx.x - e.blk0;
  const int co_tiles = (e.ld_co + 63) / 64, ci_tiles = e.cin_p / 64;
  const int co0 = (lb % co_tiles) * 64;
  lb /= co_tiles;
  const int ci0 = (lb % ci_tiles) * 64;
  const int k = lb / ci_tiles;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int r = 0; r < 64; r += 8) {
    const int ci = ci0 + r + ty;
    // packed tap k = sum of the source taps tap_group * k + c over the set bits c of tap_mask (fast-FIR filters)
    const int64_t tap = (int64_t)e.Cin * e.Cout;
    const float* src = e.w + ((int64_t)(e.tap_group * k) * e.Cin + ci) * e.Cout + co0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = tx + 32 * h;
      float v = 0.f;
      if (ci < e.Cin && co0 + c < e.Cout) {
        for (int g = 0; g < e.tap_group; ++g)
          if (e.tap_mask & (1 << g)) v += __ldg(src + g * tap + c);
      }
      tile[r + ty][c] = v;
    }
  }
  __syncthreads();
  {
    const int64_t ld = (int64_t)e.K * e.cin_p;
    const int64_t plane_stride = (int64_t)e.Cout * ld;
#pragma unroll
    for (int r = 0; r < 64; r += 8) {
      const int c = r + ty, co = co0 + c;
      if (co < e.Cout) {
        float v0 = tile[2 * tx][c], v1 = tile[2 * tx + 1][c];
        __nv_bfloat16* dst = e.fwd + (int64_t)co * ld + (int64_t)k * e.cin_p + ci0 + 2 * tx;
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl) {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
          *reinterpret_cast<uint32_t*>(dst + pl * plane_stride) = pack_bf16x2(h0, h1);
          v0 -= __bfloat162float(h0);
          v1 -= __bfloat162float(h1);
        }
      }
    }
  }
  if (e.bwd && co0 + 2 * tx < e.ld_co) {
    const int64_t plane_stride = (int64_t)e.K * e.Cin * e.ld_co;
#pragma unroll
    for (int r = 0; r < 64; r += 8) {
      const int ci = ci0 + r + ty;
      if (ci < e.Cin) {
        float v0 = tile[r + ty][2 * tx], v1 = tile[r + ty][2 * tx + 1];
        __nv_bfloat16* dst = e.bwd + ((int64_t)k * e.Cin + ci) * e.ld_co + co0 + 2 * tx;
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl) {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
          *reinterpret_cast<uint32_t*>(dst + pl * plane_stride) = pack_bf16x2(h0, h1);
          v0 -= __bfloat162float(h0);
          v1 -= __bfloat162float(h1);
        }
      }
    }
  }
}

// The nine filters of the two-level fast-FIR split of a 4J-tap layer in ONE pass over its fp32 tensor: a block loads
// the four source taps w[4i + c] of a 64(ci) x 64(co) tile once and writes, for every leaf l, the tap sum selected by
// the bits of masks[l] in both operand layouts (leaf order and masks: w2l_plan.cu kLeaves).  Packing the leaves as
// nine independent table entries read the tensor four times over.
struct LeafMasks { int m[9]; };
template <int NPL>
__global__ void __launch_bounds__(256)
pack_ffa2_kernel(const float* __restrict__ w, const Ptr9h fwd, const Ptr9h bwd, const LeafMasks masks, int J, int Cin,
                 int Cout, int cin_p, int ld_co) {
  extern __shared__ float tiles[];                    // [4][64][65]
  int lb = blockIdx.x;
  const int co_tiles = (ld_co + 63) / 64, ci_tiles = cin_p / 64;
  const int co0 = (lb % co_tiles) * 64;
  lb /= co_tiles;
  const int ci0 = (lb % ci_tiles) * 64;
  const int k = lb / ci_tiles;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
#pragma unroll
    for (int r = 0; r < 64; r += 8) {
      const int ci = ci0 + r + ty;
      const float* src = w + ((int64_t)(4 * k + g) * Cin + ci) * Cout + co0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = tx + 32 * h;
        tiles[(g * 64 + r + ty) * 65 + c] = (ci < Cin && co0 + c < Cout) ? __ldg(src + c) : 0.f;
      }
    }
  }
  __syncthreads();
  auto leaf_sum = [&](int mask, int row, int col) {
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (mask & (1 << g)) v += tiles[(g * 64 + row) * 65 + col];
    return v;
  };
  const int64_t ld = (int64_t)J * cin_p;
  const int64_t fwd_plane = (int64_t)Cout * ld, bwd_plane = (int64_t)J * Cin * ld_co;
  for (int l = 0; l < 9; ++l) {
    const int mask = masks.m[l];
#pragma unroll
    for (int r = 0; r < 64; r += 8) {
      const int c = r + ty, co = co0 + c;
      if (co < Cout) {
        float v0 = leaf_sum(mask, 2 * tx, c), v1 = leaf_sum(mask, 2 * tx + 1, c);
        __nv_bfloat16* dst = fwd.p[l] + (int64_t)co * ld + (int64_t)k * cin_p + ci0 + 2 * tx;
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl) {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
          *reinterpret_cast<uint32_t*>(dst + pl * fwd_plane) = pack_bf16x2(h0, h1);
          v0 -= __bfloat162float(h0);
          v1 -= __bfloat162float(h1);
        }
      }
    }
    if (co0 + 2 * tx < ld_co) {
#pragma unroll
      for (int r = 0; r < 64; r += 8) {
        const int ci = ci0 + r + ty;
        if (ci < Cin) {
          float v0 = leaf_sum(mask, r + ty, 2 * tx), v1 = leaf_sum(mask, r + ty, 2 * tx + 1);
          __nv_bfloat16* dst = bwd.p[l] + ((int64_t)k * Cin + ci) * ld_co + co0 + 2 * tx;
#pragma unroll
          for (int pl = 0; pl < NPL; ++pl) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
            *reinterpret_cast<uint32_t*>(dst + pl * bwd_plane) = pack_bf16x2(h0, h1);
            v0 -= __bfloat162float(h0);
            v1 -= __bfloat162float(h1);
          }
        }
      }
    }
  }
}

// Backward-layout planes ONLY, for layers whose forward kernel reads the filter MN-major (ConvParams::b_mn): no
// transposed copy, so this is a pure streaming pass -- one thread per 4 output channels of one (packed tap, ci) row.
// Leaf l = sum of the source taps group * k + c over the set bits c of masks.m[l] (plain layers: one leaf, group 1,
// mask 1; the two-level fast-FIR split of layer 8: nine leaves, group 4).
// BACKGROUND kernels (this one, ffa2_dw_combine_kernel, zero_f32_kernel): no shared memory and at most 40 registers
// (__launch_bounds__(256, 6)), so that ONE such block fits on an SM beside a resident tensor-core CTA (54 K registers,
// 227 KB of shared memory).  The step plan launches them with one block per SM on a side stream underneath tensor-core
// launches that leave the HBM idle (w2l_plan.cu, "overlap").
template <int NPL>
__device__ __forceinline__ void store_planes4(__nv_bfloat16* dst, int64_t plane_stride, const float* v) {
  float rem[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
  for (int p = 0; p < NPL; ++p) {
    uint32_t w[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(rem[2 * i], rem[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
      if (p + 1 < NPL) {
        rem[2 * i] -= __uint_as_float(w[i] << 16);
        rem[2 * i + 1] -= __uint_as_float(w[i] & 0xffff0000u);
      }
    }
    *reinterpret_cast<uint2*>(dst + p * plane_stride) = make_uint2(w[0], w[1]);
  }
}

template <int NPL>
__global__ void __launch_bounds__(256, 6)
pack_bwd_kernel(const float* __restrict__ w, const Ptr9h bwd, const LeafMasks masks, int n_leaves, int group, int J,
                int Cin, int Cout, int ld_co) {
  const int cg = ld_co / 4;
  const int64_t rows = (int64_t)J * Cin;
  const int64_t groups = rows * cg;
  const int64_t tap = (int64_t)Cin * Cout;
  const int64_t plane = rows * ld_co;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(g % cg);
    const int64_t row = g / cg;
    const int ci = (int)(row % Cin);
    const int k = (int)(row / Cin);
    float4 t[4];
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) {
      t[gg] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gg < group && c4 * 4 < Cout)
        t[gg] = __ldg(reinterpret_cast<const float4*>(w + (int64_t)(group * k + gg) * tap + (int64_t)ci * Cout + c4 * 4));
    }
#pragma unroll 1
    for (int l = 0; l < n_leaves; ++l) {
      const int mask = masks.m[l];
      float v[4] = {0.f, 0.f, 0.f, 0.f};           // same order of additions as pack_ffa2_kernel (taps ascending)
#pragma unroll
      for (int gg = 0; gg < 4; ++gg)
        if (mask & (1 << gg)) { v[0] += t[gg].x; v[1] += t[gg].y; v[2] += t[gg].z; v[3] += t[gg].w; }
      store_planes4<NPL>(bwd.p[l] + row * ld_co + c4 * 4, plane, v);
    }
  }
}

__global__ void __launch_bounds__(256, 6)
zero_f32_kernel(float4* __restrict__ dst, int64_t n4) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) dst[i] = z;
}

// db[n] += sum_rows sum_planes dz[pl][row][n]; block (32 column octets, 8 row lanes): each thread streams 16-byte
// vectors (8 bf16 columns) down its rows; grid (ceil(ld/256), row chunks)
__global__ void __launch_bounds__(256)
bias_grad_planes_kernel(const __nv_bfloat16* __restrict__ dz, int64_t rows, int N, int ld, int n_planes,
                        float* __restrict__ db, int rows_per_block) {
  __shared__ float part[8][32 * 8 + 1];
  const int c8 = blockIdx.x * 32 + threadIdx.x;            // column octet index
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  const int64_t plane_stride = rows * ld;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (c8 * 8 < ld) {
    for (int pl = 0; pl < n_planes; ++pl) {
      const __nv_bfloat16* src = dz + pl * plane_stride + c8 * 8;
#pragma unroll 4
      for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + r * ld);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[2 * i] += __uint_as_float(w[i] << 16);
          acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[threadIdx.y][threadIdx.x * 8 + i] = acc[i];
  __syncthreads();
  const int tid = threadIdx.y * 32 + threadIdx.x;          // 256 threads <-> 256 columns of this block
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += part[i][tid];
  const int n = blockIdx.x * 256 + tid;
  if (n < N) atomicAdd(db + n, s);
}

__global__ void __launch_bounds__(256)
merge_planes_kernel(const __nv_bfloat16* __restrict__ planes, int64_t rows, int cols, int ld, int n_planes,
                    float* __restrict__ dst, int64_t ld_dst) {
  const int64_t total = rows * cols;
  const int64_t plane_stride = rows * ld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    const int64_t r = i / cols;
    float acc = 0.f;
    for (int pl = n_planes - 1; pl >= 0; --pl) acc += __bfloat162float(planes[pl * plane_stride + r * ld + c]);
    dst[r * ld_dst + c] = acc;
  }
}

// ---- fast-FIR split of a stride-1 K-tap layer (experimental, SPEECHT_B200_FFA=1) -------------
// xs[pl][b][r][:] = planes of x[b][2r][:] + x[b][2r+1][:] (x[b][T][:] = 0); one thread per 8 channels
template <int NPL>
__global__ void __launch_bounds__(256)
pair_sum_planes_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ xs, int B, int T, int Tx,
                       int ld) {
  const int64_t groups = (int64_t)B * Tx * (ld / 8);
  const int64_t in_plane = (int64_t)B * T * ld, out_plane = (int64_t)B * Tx * ld;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(g % (ld / 8));
    const int64_t br = g / (ld / 8);
    const int r = (int)(br % Tx);
    const int b = (int)(br / Tx);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = 2 * r + h;
      if (t >= T) continue;
      for (int pl = NPL - 1; pl >= 0; --pl) {
        const uint4 q = *reinterpret_cast<const uint4*>(x + pl * in_plane + ((int64_t)b * T + t) * ld + c8 * 8);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[2 * i] += __uint_as_float(w[i] << 16);
          v[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
        }
      }
    }
    store_planes8<NPL>(xs + ((int64_t)b * Tx + r) * ld + c8 * 8, out_plane, v);
  }
}

// y[b][2u] = act(A00[u] + A11[u] + bias), y[b][2u+1] = act(S[u] - A11[u] - A00[u+1] + bias) -> bf16 planes
// [NPL][B][To][ld_out]; the three partial products are fp32 [B][Tu][ld_p]; one thread per 8 channels of one u
template <int NPL>
__global__ void __launch_bounds__(256)
ffa_combine_kernel(const float* __restrict__ a00, const float* __restrict__ a11, const float* __restrict__ sm,
                   const float* __restrict__ bias, int relu, __nv_bfloat16* __restrict__ out,
                   uint32_t* __restrict__ mask_out, int B, int To, int Tu, int N, int ld_p, int ld_out) {
  const int cg = ld_out / 8;
  const int64_t groups = (int64_t)B * Tu * cg;
  const int64_t out_plane = (int64_t)B * To * ld_out;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(g % cg);
    const int64_t bu = g / cg;
    const int u = (int)(bu % Tu);
    const int b = (int)(bu / Tu);
    if (2 * u >= To) continue;
    const int64_t row = ((int64_t)b * Tu + u) * ld_p + c8 * 8;
    float p0[8], p0n[8], p1[8], ps[8], ye[8], yo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool ok = c8 * 8 + i < N;
      p0[i] = ok ? a00[row + i] : 0.f;
      p1[i] = ok ? a11[row + i] : 0.f;
      ps[i] = ok ? sm[row + i] : 0.f;
      p0n[i] = (ok && u + 1 < Tu) ? a00[row + ld_p + i] : 0.f;
      const float bv = (ok && bias) ? __ldg(bias + c8 * 8 + i) : 0.f;
      ye[i] = p0[i] + p1[i] + bv;
      yo[i] = ps[i] - p1[i] - p0n[i] + bv;
      if (relu) { ye[i] = fmaxf(ye[i], 0.f); yo[i] = fmaxf(yo[i], 0.f); }
      if (!ok) { ye[i] = 0.f; yo[i] = 0.f; }
    }
    __nv_bfloat16* o = out + ((int64_t)b * To + 2 * u) * ld_out + c8 * 8;
    store_planes8<NPL>(o, out_plane, ye);
    if (2 * u + 1 < To) store_planes8<NPL>(o + ld_out, out_plane, yo);
    if (mask_out && c8 * 8 < N) {
      const int64_t row = (int64_t)b * To + 2 * u;
      store_mask_byte(mask_out, (int64_t)B * To, row, c8, ye);
      if (2 * u + 1 < To) store_mask_byte(mask_out, (int64_t)B * To, row + 1, c8, yo);
    }
  }
}

// Backward of the fast-FIR form.  From dy = d(loss)/d(y) of the layer (planes [NPL][B][To][ld]) the gradients of the
// three partial products: dA00[u] = dy[2u] - dy[2u-1], dA11[u] = dy[2u] - dy[2u+1], dS[u] = dy[2u+1] (dy = 0 outside
// [0, To)), each as planes [NPL][B][Tu][ld]; one thread per 8 channels of one u
template <int NPL>
__global__ void __launch_bounds__(256)
ffa_dz_prep_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ d00,
                   __nv_bfloat16* __restrict__ d11, __nv_bfloat16* __restrict__ ds, int B, int To, int Tu, int ld) {
  const int cg = ld / 8;
  const int64_t groups = (int64_t)B * Tu * cg;
  const int64_t in_plane = (int64_t)B * To * ld, out_plane = (int64_t)B * Tu * ld;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(g % cg);
    const int64_t bu = g / cg;
    const int u = (int)(bu % Tu);
    const int b = (int)(bu / Tu);
    float y[3][8];                                  // rows 2u-1, 2u, 2u+1
#pragma unroll
    for (int h = 0; h < 3; ++h) {
      const int t = 2 * u - 1 + h;
#pragma unroll
      for (int i = 0; i < 8; ++i) y[h][i] = 0.f;
      if (t < 0 || t >= To) continue;
      for (int pl = NPL - 1; pl >= 0; --pl) {
        const uint4 q = *reinterpret_cast<const uint4*>(dy + pl * in_plane + ((int64_t)b * To + t) * ld + c8 * 8);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          y[h][2 * i] += __uint_as_float(w[i] << 16);
          y[h][2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
        }
      }
    }
    float a[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = y[1][i] - y[0][i]; c[i] = y[1][i] - y[2][i]; }
    const int64_t o = ((int64_t)b * Tu + u) * ld + c8 * 8;
    store_planes8<NPL>(d00 + o, out_plane, a);
    store_planes8<NPL>(d11 + o, out_plane, c);
    store_planes8<NPL>(ds + o, out_plane, y[2]);
  }
}

// dx[2r+1] = d_odd[r] + d_xs[r], dx[2r] = d_even[r] + d_xs[r], times the ReLU mask of the layer below, split to planes
// [NPL][B][T][ld]; column sums of what is stored go to db (bias gradient of the layer below).  Partials are fp32
// [B][Tx][ldp].  block (32 channel octets, 8 row lanes); a block owns a contiguous slab of (b, r) rows.
template <int NPL>
__global__ void __launch_bounds__(256)
ffa_dx_combine_kernel(const float* __restrict__ d_odd, const float* __restrict__ d_even, const float* __restrict__ d_xs,
                      const uint32_t* __restrict__ mask, __nv_bfloat16* __restrict__ out, float* __restrict__ db,
                      int B, int T, int Tx, int N, int ldp, int ld, int rows_per_block) {
  __shared__ float part[8][32 * 8 + 1];
  const int c8 = threadIdx.x;                                 // channel octet (ld <= 256)
  const int64_t rows = (int64_t)B * Tx;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  const int64_t out_plane = (int64_t)B * T * ld;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (c8 * 8 < ld) {
    for (int64_t br = r0 + threadIdx.y; br < r1; br += 8) {
      const int b = (int)(br / Tx), r = (int)(br - (int64_t)b * Tx);
      if (2 * r >= T) continue;
      const int64_t prow = br * ldp + c8 * 8;
      float xs[8], ev[8], od[8];
#pragma unroll
      for (int i = 0; i < 8; i += 4) {
        const float4 a = *reinterpret_cast<const float4*>(d_xs + prow + i);
        const float4 e = *reinterpret_cast<const float4*>(d_even + prow + i);
        const float4 o = *reinterpret_cast<const float4*>(d_odd + prow + i);
        xs[i] = a.x; xs[i + 1] = a.y; xs[i + 2] = a.z; xs[i + 3] = a.w;
        ev[i] = e.x; ev[i + 1] = e.y; ev[i + 2] = e.z; ev[i + 3] = e.w;
        od[i] = o.x; od[i + 1] = o.y; od[i + 2] = o.z; od[i + 3] = o.w;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int t = 2 * r + h;
        if (t >= T) continue;
        const int64_t orow = ((int64_t)b * T + t) * ld + c8 * 8;
        const uint32_t mbits = c8 * 8 < N ? load_mask_byte(mask, (int64_t)B * T, (int64_t)b * T + t, c8) : 0u;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool keep = (mbits >> i) & 1u;                                  // ReLU below was active (real channel)
          v[i] = keep ? (h ? od[i] : ev[i]) + xs[i] : 0.f;
          acc[i] += v[i];
        }
        store_planes8<NPL>(out + orow, out_plane, v);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[threadIdx.y][threadIdx.x * 8 + i] = acc[i];
  __syncthreads();
  const int tid = threadIdx.y * 32 + threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += part[i][tid];
  if (tid < N) atomicAdd(db + tid, s);
}

// dW[2j] += cs[j], dW[2j+1] += cs[j]  (filter gradient of the fast-FIR form: the pair-sum correlation feeds both taps)
__global__ void __launch_bounds__(256)
ffa_dw_combine_kernel(float* __restrict__ dW, const float* __restrict__ cs, int J, int64_t tap_elems4) {
  const int64_t total = (int64_t)J * tap_elems4;
  float4* w4 = reinterpret_cast<float4*>(dW);
  const float4* c4 = reinterpret_cast<const float4*>(cs);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i / tap_elems4, e = i - j * tap_elems4;
    const float4 c = c4[i];
    float4 a = w4[(2 * j) * tap_elems4 + e], b = w4[(2 * j + 1) * tap_elems4 + e];
    a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
    b.x += c.x; b.y += c.y; b.z += c.z; b.w += c.w;
    w4[(2 * j) * tap_elems4 + e] = a;
    w4[(2 * j + 1) * tap_elems4 + e] = b;
  }
}

// ---- two levels of the fast-FIR split (nine quarter-rate 8-tap problems; algebra: tools/ffa2_study.py) --------------
// Leaf order everywhere: 0 XX, 1 XY, 2 XZ, 3 YX, 4 YY, 5 YZ, 6 ZX, 7 ZY, 8 ZZ (first letter: level-1 product X1 = odd
// rows (*) even taps, Y1 = even rows (*) odd taps, Z1 = pair sums (*) tap sums; second letter: the same split inside it).

__device__ __forceinline__ void load_merged8(const __nv_bfloat16* base, int64_t plane, int npl, float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  for (int pl = npl - 1; pl >= 0; --pl) {
    const uint4 q = *reinterpret_cast<const uint4*>(base + pl * plane);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] += __uint_as_float(w[i] << 16);
      v[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
    }
  }
}

// The five quarter-rate input sequences that are sums (the other four leaves read row views x[4r+c] directly):
//   s[0][r] = x[4r-1] + x[4r+1]   (leaf XZ)      s[1][r] = x[4r] + x[4r+2]     (leaf YZ)
//   s[2][r] = x[4r+2] + x[4r+3]   (leaf ZX)      s[3][r] = x[4r] + x[4r+1]     (leaf ZY)
//   s[4][r] = s[3][r] + s[2][r]   (leaf ZZ)      x = 0 outside [0, T); planes [NPL][B][Tq][ld]
template <int NPL>
__global__ void __launch_bounds__(256)
ffa2_inputs_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* s0, __nv_bfloat16* s1, __nv_bfloat16* s2,
                   __nv_bfloat16* s3, __nv_bfloat16* s4, int B, int T, int Tq, int ld) {
  const int cg = ld / 8;
  const int64_t groups = (int64_t)B * Tq * cg;
  const int64_t in_plane = (int64_t)B * T * ld, out_plane = (int64_t)B * Tq * ld;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(g % cg);
    const int64_t br = g / cg;
    const int r = (int)(br % Tq);
    const int b = (int)(br / Tq);
    float v[5][8];                                   // rows 4r-1 .. 4r+3
#pragma unroll
    for (int h = 0; h < 5; ++h) {
      const int t = 4 * r - 1 + h;
      if (t >= 0 && t < T) load_merged8(x + ((int64_t)b * T + t) * ld + c8 * 8, in_plane, NPL, v[h]);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[h][i] = 0.f;
      }
    }
    float o[5][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[0][i] = v[0][i] + v[2][i];
      o[1][i] = v[1][i] + v[3][i];
      o[2][i] = v[3][i] + v[4][i];
      o[3][i] = v[1][i] + v[2][i];
      o[4][i] = o[3][i] + o[2][i];
    }
    const int64_t off = ((int64_t)b * Tq + r) * ld + c8 * 8;
    store_planes8<NPL>(s0 + off, out_plane, o[0]);
    store_planes8<NPL>(s1 + off, out_plane, o[1]);
    store_planes8<NPL>(s2 + off, out_plane, o[2]);
    store_planes8<NPL>(s3 + off, out_plane, o[3]);
    store_planes8<NPL>(s4 + off, out_plane, o[4]);
  }
}

// Forward combine: X1[2q] = XX[q] + XY[q], X1[2q+1] = XZ[q] - XY[q] - XX[q+1] (likewise Y1, Z1), then
// y[2u] = X1[u] + Y1[u], y[2u+1] = Z1[u] - Y1[u] - X1[u+1]; + bias, ReLU, plane split.  Partials fp32 [B][Tq][ld_p];
// one thread per 8 channels of one q (four output rows 4q .. 4q+3).
template <int NPL>
__global__ void __launch_bounds__(256)
ffa2_combine_kernel(const Ptr9c part, const float* __restrict__ bias, int relu, __nv_bfloat16* __restrict__ out,
                    uint32_t* __restrict__ mask_out, int B, int To, int Tq, int N, int ld_p, int ld_out) {
  const int cg = ld_out / 8;
  const int64_t groups = (int64_t)B * Tq * cg;
  const int64_t out_plane = (int64_t)B * To * ld_out;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(g % cg);
    const int64_t bq = g / cg;
    const int q = (int)(bq % Tq);
    const int b = (int)(bq / Tq);
    if (4 * q >= To) continue;
    const int64_t row = ((int64_t)b * Tq + q) * ld_p + c8 * 8;
    const bool nxt = q + 1 < Tq;
    // 16-byte loads: the partial products are read once, 8 channels per thread (ld_p and N are multiples of 8 for
    // every layer this is used on; a ragged last octet falls back to scalars)
    const bool full8 = c8 * 8 + 8 <= N && (ld_p & 3) == 0;
    float a[9][8], an[4][8];                         // an: rows q+1 of XX, XY, YX, ZX
    auto load8 = [&](const float* src, bool on, float* dst) {
      if (on && full8) {
        const float4 u = *reinterpret_cast<const float4*>(src), w = *reinterpret_cast<const float4*>(src + 4);
        dst[0] = u.x; dst[1] = u.y; dst[2] = u.z; dst[3] = u.w; dst[4] = w.x; dst[5] = w.y; dst[6] = w.z; dst[7] = w.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = (on && c8 * 8 + i < N) ? src[i] : 0.f;
      }
    };
#pragma unroll
    for (int l = 0; l < 9; ++l) load8(part.p[l] + row, true, a[l]);
    load8(part.p[0] + row + ld_p, nxt, an[0]);
    load8(part.p[1] + row + ld_p, nxt, an[1]);
    load8(part.p[3] + row + ld_p, nxt, an[2]);
    load8(part.p[6] + row + ld_p, nxt, an[3]);
    float y[4][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool ok = c8 * 8 + i < N;
      const float x1e = a[0][i] + a[1][i], x1o = a[2][i] - a[1][i] - an[0][i], x1n = an[0][i] + an[1][i];
      const float y1e = a[3][i] + a[4][i], y1o = a[5][i] - a[4][i] - an[2][i];
      const float z1e = a[6][i] + a[7][i], z1o = a[8][i] - a[7][i] - an[3][i];
      const float bv = (ok && bias) ? __ldg(bias + c8 * 8 + i) : 0.f;
      y[0][i] = x1e + y1e + bv;
      y[1][i] = z1e - y1e - x1o + bv;
      y[2][i] = x1o + y1o + bv;
      y[3][i] = z1o - y1o - x1n + bv;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        if (relu) y[h][i] = fmaxf(y[h][i], 0.f);
        if (!ok) y[h][i] = 0.f;
      }
    }
    __nv_bfloat16* o = out + ((int64_t)b * To + 4 * q) * ld_out + c8 * 8;
#pragma unroll
    for (int h = 0; h < 4; ++h)
      if (4 * q + h < To) {
        store_planes8<NPL>(o + h * ld_out, out_plane, y[h]);
        if (mask_out && c8 * 8 < N) store_mask_byte(mask_out, (int64_t)B * To, (int64_t)b * To + 4 * q + h, c8, y[h]);
      }
  }
}

// Backward prepare: gradients of the nine leaf products from dy (planes [NPL][B][To][ld]), planes [NPL][B][Tq][ld]:
// level 1: GX[u] = dy[2u] - dy[2u-1], GY[u] = dy[2u] - dy[2u+1], GZ[u] = dy[2u+1]; level 2 of each G:
// d?X[q] = G[2q] - G[2q-1], d?Y[q] = G[2q] - G[2q+1], d?Z[q] = G[2q+1]  (dy = 0 outside [0, To)).
// One thread per FOUR channels of one q: 8-byte loads / stores keep a warp on 256 contiguous bytes, and at ~60
// registers four blocks are resident per SM (eight channels per thread: 98 registers, two blocks, 46 % of HBM peak).
template <int NPL>
__global__ void __launch_bounds__(256, 4)
ffa2_dz_prep_kernel(const __nv_bfloat16* __restrict__ dy, const Ptr9h out, int B, int To, int Tq, int ld) {
  const int cg = ld / 4;
  const int64_t groups = (int64_t)B * Tq * cg;
  const int64_t in_plane = (int64_t)B * To * ld, out_plane = (int64_t)B * Tq * ld;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(g % cg);
    const int64_t bq = g / cg;
    const int q = (int)(bq % Tq);
    const int b = (int)(bq / Tq);
    float d[7][4];                                   // dy rows 4q-3 .. 4q+3
#pragma unroll
    for (int h = 0; h < 7; ++h) {
      const int t = 4 * q - 3 + h;
#pragma unroll
      for (int i = 0; i < 4; ++i) d[h][i] = 0.f;
      if (t >= 0 && t < To) {
        const __nv_bfloat16* src = dy + ((int64_t)b * To + t) * ld + c4 * 4;
#pragma unroll
        for (int pl = NPL - 1; pl >= 0; --pl) {
          const uint2 w = *reinterpret_cast<const uint2*>(src + pl * in_plane);
          d[h][0] += __uint_as_float(w.x << 16);
          d[h][1] += __uint_as_float(w.x & 0xffff0000u);
          d[h][2] += __uint_as_float(w.y << 16);
          d[h][3] += __uint_as_float(w.y & 0xffff0000u);
        }
      }
    }
    const int64_t off = ((int64_t)b * Tq + q) * ld + c4 * 4;
    float o[4];
    // G(2q-1), G(2q), G(2q+1) of the three level-1 gradients, from rows d[k+3] = dy[4q+k]
#pragma unroll
    for (int lvl = 0; lvl < 3; ++lvl) {
      float gm[4], g0[4], gp[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (lvl == 0) { gm[i] = d[1][i] - d[0][i]; g0[i] = d[3][i] - d[2][i]; gp[i] = d[5][i] - d[4][i]; }
        else if (lvl == 1) { gm[i] = d[1][i] - d[2][i]; g0[i] = d[3][i] - d[4][i]; gp[i] = d[5][i] - d[6][i]; }
        else { gm[i] = d[2][i]; g0[i] = d[4][i]; gp[i] = d[6][i]; }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = g0[i] - gm[i];
      store_planes4<NPL>(out.p[3 * lvl + 0] + off, out_plane, o);
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = g0[i] - gp[i];
      store_planes4<NPL>(out.p[3 * lvl + 1] + off, out_plane, o);
      store_planes4<NPL>(out.p[3 * lvl + 2] + off, out_plane, gp);
    }
  }
}

// Data-gradient combine of the nine leaf partials g[l] (fp32 [B][Tqx][ldp]):
//   dx[4r]   = gYY[r] + gYZ[r] + gZY[r] + gZZ[r]        dx[4r+1] = gXX[r] + gXZ[r]   + gZY[r] + gZZ[r]
//   dx[4r+2] = gYX[r] + gYZ[r] + gZX[r] + gZZ[r]        dx[4r+3] = gXY[r] + gXZ[r+1] + gZX[r] + gZZ[r]
// times the ReLU mask of the layer below, split to planes [NPL][B][T][ld]; column sums into db.
template <int NPL>
__global__ void __launch_bounds__(256)
ffa2_dx_combine_kernel(const Ptr9c gp, const uint32_t* __restrict__ mask, __nv_bfloat16* __restrict__ out,
                       float* __restrict__ db, int B, int T, int Tqx, int N, int ldp, int ld, int rows_per_block) {
  __shared__ float part[8][32 * 8 + 1];
  const int c8 = threadIdx.x;
  const int64_t rows = (int64_t)B * Tqx;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  const int64_t out_plane = (int64_t)B * T * ld;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (c8 * 8 < ld) {
    for (int64_t br = r0 + threadIdx.y; br < r1; br += 8) {
      const int b = (int)(br / Tqx), r = (int)(br - (int64_t)b * Tqx);
      if (4 * r >= T) continue;
      const int64_t prow = br * ldp + c8 * 8;
      float g[9][8], gxz_n[8];
#pragma unroll
      for (int l = 0; l < 9; ++l) {
#pragma unroll
        for (int i = 0; i < 8; i += 4) {
          const float4 a = *reinterpret_cast<const float4*>(gp.p[l] + prow + i);
          g[l][i] = a.x; g[l][i + 1] = a.y; g[l][i + 2] = a.z; g[l][i + 3] = a.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; i += 4) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r + 1 < Tqx) a = *reinterpret_cast<const float4*>(gp.p[2] + prow + ldp + i);
        gxz_n[i] = a.x; gxz_n[i + 1] = a.y; gxz_n[i + 2] = a.z; gxz_n[i + 3] = a.w;
      }
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int t = 4 * r + h;
        if (t >= T) continue;
        const int64_t orow = ((int64_t)b * T + t) * ld + c8 * 8;
        const uint32_t mbits = c8 * 8 < N ? load_mask_byte(mask, (int64_t)B * T, (int64_t)b * T + t, c8) : 0u;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool keep = (mbits >> i) & 1u;
          float s;
          if (h == 0) s = (g[4][i] + g[5][i]) + (g[7][i] + g[8][i]);
          else if (h == 1) s = (g[0][i] + g[2][i]) + (g[7][i] + g[8][i]);
          else if (h == 2) s = (g[3][i] + g[5][i]) + (g[6][i] + g[8][i]);
          else s = (g[1][i] + gxz_n[i]) + (g[6][i] + g[8][i]);
          v[i] = keep ? s : 0.f;
          acc[i] += v[i];
        }
        store_planes8<NPL>(out + orow, out_plane, v);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[threadIdx.y][threadIdx.x * 8 + i] = acc[i];
  __syncthreads();
  const int tid = threadIdx.y * 32 + threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += part[i][tid];
  if (tid < N) atomicAdd(db + tid, s);
}

// Filter-gradient combine of the nine leaf correlations c[l] ([J][tap] fp32, J = taps / 4):
//   dW[4i]   = cXX + cXZ + cZX + cZZ      dW[4i+1] = cYX + cYZ + cZX + cZZ
//   dW[4i+2] = cXY + cXZ + cZY + cZZ      dW[4i+3] = cYY + cYZ + cZY + cZZ
__global__ void __launch_bounds__(256, 6)      // a background kernel (see pack_bwd_kernel): <= 40 registers, no smem
ffa2_dw_combine_kernel(float* __restrict__ dW, const Ptr9c c, int J, int64_t tap_elems4) {
  const int64_t total = (int64_t)J * tap_elems4;
  float4* w4 = reinterpret_cast<float4*>(dW);
  auto ld = [&](int l, int64_t i) { return reinterpret_cast<const float4*>(c.p[l])[i]; };
  auto add = [](float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i / tap_elems4, e = i - j * tap_elems4;
    // every output is (a + b) + (c + d) with the pairs of the formulas above; few values are live at a time
    const float4 v8 = ld(8, i);
    const float4 s68 = add(ld(6, i), v8), s78 = add(ld(7, i), v8);
    const float4 v2 = ld(2, i), v5 = ld(5, i);
    w4[(4 * j + 0) * tap_elems4 + e] = add(add(ld(0, i), v2), s68);
    w4[(4 * j + 1) * tap_elems4 + e] = add(add(ld(3, i), v5), s68);
    w4[(4 * j + 2) * tap_elems4 + e] = add(add(ld(1, i), v2), s78);
    w4[(4 * j + 3) * tap_elems4 + e] = add(add(ld(4, i), v5), s78);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
  }
  return fn;
}

constexpr int kMaxDevices = 64;
int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

// SMs the persistent tensor-core grids may occupy: all of them, minus the ones a caller has set aside for a collective
// that runs concurrently (set_reserved_sms; data-parallel training reserves room for the NCCL allreduce kernel, whose
// CTAs cannot share an SM with a 227 KB tensor-core CTA and would otherwise start only when a wave drains)
int g_reserved_sms = 0;
int tc_sms() {
  const int n = st_num_sms() - g_reserved_sms;
  return n < 8 ? 8 : n;
}

int grid_for(int work_items) {
  const int sms = tc_sms();
  return work_items < sms ? work_items : sms;
}

// Launch with the programmatic-stream-serialization attribute (PDL): the kernel may begin while the previous kernel
// of the stream drains; it synchronises with `griddepcontrol.wait` before touching global memory.
template <class Kernel, class... Args>
cudaError_t launch_pdl_cluster_t(Kernel kernel, int grid, int threads, int cluster, int smem, cudaStream_t stream,
                                 const Args&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  static const bool pdl = []() { const char* e = getenv("SPEECHT_B200_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cluster > 1) {                       // CTA pairs: consecutive blocks form a cluster on the SMs of one TPC
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <class Kernel, class... Args>
cudaError_t launch_pdl_cluster(Kernel kernel, int grid, int cluster, int smem, cudaStream_t stream, const Args&... args) {
  return launch_pdl_cluster_t(kernel, grid, kThreads, cluster, smem, stream, args...);
}

template <class Kernel, class... Args>
cudaError_t launch_pdl(Kernel kernel, int grid, int smem, cudaStream_t stream, const Args&... args) {
  return launch_pdl_cluster(kernel, grid, 1, smem, stream, args...);
}

// CTA-pair (cta_group::2) tiles for the 256-wide forward / data-gradient launches (SPEECHT_B200_PAIR=0 disables them)
bool pair_enabled() {                        // read when a plan is bound (want_pair), not cached
  const char* e = getenv("SPEECHT_B200_PAIR");
  return !(e && e[0] == '0');
}

// grid of a CTA-pair launch: one cluster of two CTAs per work item, at most one cluster per TPC
int pair_grid(int pair_work) {
  const int clusters = tc_sms() / 2;
  return 2 * (pair_work < clusters ? pair_work : clusters);
}

template <int BLOCK_N, int NPL, bool EARLY, bool PAIR, bool BMN = false>
int launch_conv_e(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const ConvParams& p,
                  cudaStream_t stream) {
  using Cfg = ConvCfg<BLOCK_N, NPL, PAIR>;
  if constexpr (!BMN && BLOCK_N == 256 && NPL <= 2) {
    if (p.b_mn) return launch_conv_e<BLOCK_N, NPL, EARLY, PAIR, true>(tmA, tmB, tmOut, p, stream);
  }
  // the opt-in for > 48 KB of dynamic shared memory is a per-DEVICE function attribute
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (!configured[dev]) {
    ST_CUDA_CALL(cudaFuncSetAttribute(tc_conv_kernel<BLOCK_N, NPL, EARLY, 1, PAIR, BMN>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured[dev] = true;
  }
  const int m_tiles = p.B * p.m_tiles_per_utt;
  TmSet1 a, b;
  a.m[0] = tmA;
  b.m[0] = tmB;
  if constexpr (PAIR) {
    ST_CUDA_CALL(launch_pdl_cluster(tc_conv_kernel<BLOCK_N, NPL, EARLY, 1, true, BMN>,
                                    pair_grid(((m_tiles + 1) / 2) * p.n_tiles), 2, Cfg::SMEM_BYTES, stream, a, b, tmOut, p));
  } else {
    ST_CUDA_CALL(launch_pdl_cluster_t(tc_conv_kernel<BLOCK_N, NPL, EARLY, 1, false, BMN>, grid_for(m_tiles * p.n_tiles),
                                      Cfg::THREADS, 1, Cfg::SMEM_BYTES, stream, a, b, tmOut, p));
  }
  return ST_OK;
}

template <int BLOCK_N, int NPL, bool PAIR, bool BMN = false>
int launch_conv_multi_t(const CUtensorMap* tmA, const CUtensorMap* tmB, const ConvParams& p0, cudaStream_t stream) {
  using Cfg = ConvCfg<BLOCK_N, NPL, PAIR>;
  constexpr bool EARLY = Cfg::ACC_STAGES == 1;      // several work items per CTA, long main loops: release TMEM early
  if constexpr (!BMN) {
    if (p0.b_mn) return launch_conv_multi_t<BLOCK_N, NPL, PAIR, true>(tmA, tmB, p0, stream);
  }
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (!configured[dev]) {
    ST_CUDA_CALL(cudaFuncSetAttribute(tc_conv_kernel<BLOCK_N, NPL, EARLY, kMaxProblems, PAIR, BMN>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured[dev] = true;
  }
  ConvParams p = p0;
  p.tma_store = 0;
  p.timeline = nullptr;
  TmSetN a, b;
  for (int q = 0; q < kMaxProblems; ++q) {
    a.m[q] = tmA[q < p.n_problems ? q : 0];
    b.m[q] = tmB[q < p.n_problems ? q : 0];
  }
  const int m_tiles = p.B * p.m_tiles_per_utt;
  const int per_tile = (p.k_split > 1 ? p.k_split : 1) * p.n_problems;
  if constexpr (PAIR) {
    ST_CUDA_CALL(launch_pdl_cluster(tc_conv_kernel<BLOCK_N, NPL, EARLY, kMaxProblems, true, BMN>,
                                    pair_grid(((m_tiles + 1) / 2) * p.n_tiles * per_tile), 2, Cfg::SMEM_BYTES, stream, a, b,
                                    a.m[0], p));
  } else {
    ST_CUDA_CALL(launch_pdl(tc_conv_kernel<BLOCK_N, NPL, EARLY, kMaxProblems, false, BMN>,
                            grid_for(m_tiles * p.n_tiles * per_tile), Cfg::SMEM_BYTES, stream, a, b, a.m[0], p));
  }
  return ST_OK;
}

long long* g_timeline_buf = nullptr;
int g_timeline_index = -1, g_timeline_count = 0;

template <int BLOCK_N, int NPL>
int launch_conv_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmOut, const ConvParams& p0,
                  cudaStream_t stream) {
  // the TMA-store epilogue needs a store map, bf16 output planes and the staging tiles (<= 2 planes)
  ConvParams p = p0;
  // bf16 planes leave through the store map in the one- and two-plane builds (compile-time in the epilogue)
  p.tma_store = p.out_planes && NPL <= 2;
  if (p.tma_store && !tmOut) {
    st_set_error("launch_conv: bf16 output planes need their store tensor map (n_planes <= 2)");
    return ST_ERR_INVALID_ARG;
  }
  p.timeline = (g_timeline_buf && g_timeline_count++ == g_timeline_index) ? g_timeline_buf : nullptr;
  const CUtensorMap& tmO = p.tma_store ? *tmOut : tmA;          // placeholder when unused
  // early TMEM release pays when a CTA has several tiles, no second accumulator stage, and a main loop long enough
  // to hide the register-resident epilogue behind it
  const int m_tiles = p.B * p.m_tiles_per_utt;
  // CTA pairs for the wide launches with enough m tiles to pair up (the 250-channel layers have one tile per SM and
  // nothing to share; they stay on single CTAs)
  constexpr bool kCanPair = BLOCK_N == 256 && NPL <= 2;
  const bool pair = kCanPair && p.pair;
  const int work = pair ? ((m_tiles + 1) / 2) * p.n_tiles : m_tiles * p.n_tiles;
  const int ctas = pair ? tc_sms() / 2 : tc_sms();
  const bool early = ConvCfg<BLOCK_N, NPL>::ACC_STAGES == 1 && work > ctas && p.taps * p.chunks_per_tap >= 16;
  if constexpr (kCanPair) {
    if (pair) {
      if constexpr (ConvCfg<BLOCK_N, NPL>::ACC_STAGES == 1) {
        if (early) return launch_conv_e<BLOCK_N, NPL, true, true>(tmA, tmB, tmO, p, stream);
      }
      return launch_conv_e<BLOCK_N, NPL, false, true>(tmA, tmB, tmO, p, stream);
    }
  }
  if constexpr (ConvCfg<BLOCK_N, NPL>::ACC_STAGES == 1) {
    if (early) return launch_conv_e<BLOCK_N, NPL, true, false>(tmA, tmB, tmO, p, stream);
  }
  return launch_conv_e<BLOCK_N, NPL, false, false>(tmA, tmB, tmO, p, stream);
}

// persistent filter-gradient grid: one CTA per SM, or one cluster of two CTAs per TPC
template <int BLOCK_N, int NPL, int NPROB, bool PAIR, class Tm>
int launch_wgrad_k(const Tm& x, const Tm& dz, const WgradParams& p, cudaStream_t stream) {
  using Cfg = WgradCfg<BLOCK_N, NPL, PAIR>;
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (!configured[dev]) {
    ST_CUDA_CALL(cudaFuncSetAttribute(tc_wgrad_kernel<BLOCK_N, NPL, NPROB, PAIR>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured[dev] = true;
  }
  if constexpr (PAIR) {
    ST_CUDA_CALL(launch_pdl_cluster(tc_wgrad_kernel<BLOCK_N, NPL, NPROB, true>, 2 * (tc_sms() / 2), 2, Cfg::SMEM_BYTES,
                                    stream, x, dz, p));
  } else {
    ST_CUDA_CALL(launch_pdl(tc_wgrad_kernel<BLOCK_N, NPL, NPROB, false>, tc_sms(), Cfg::SMEM_BYTES, stream, x, dz, p));
  }
  return ST_OK;
}

template <int BLOCK_N, int NPL>
int launch_wgrad_t(const CUtensorMap& tmX, const CUtensorMap& tmDZ, const WgradParams& p, cudaStream_t stream) {
  TmSet1 x, dz;
  x.m[0] = tmX;
  dz.m[0] = tmDZ;
  if constexpr (BLOCK_N == 256 && NPL <= 2) {
    if (wgrad_pair(p.taps, p.m_tiles, BLOCK_N, NPL)) return launch_wgrad_k<BLOCK_N, NPL, 1, true>(x, dz, p, stream);
  }
  return launch_wgrad_k<BLOCK_N, NPL, 1, false>(x, dz, p, stream);
}

template <int BLOCK_N, int NPL>
int launch_wgrad_multi_t(const CUtensorMap* tmX, const CUtensorMap* tmDZ, const WgradParams& p, cudaStream_t stream) {
  TmSetN x, dz;
  for (int q = 0; q < kMaxProblems; ++q) {
    x.m[q] = tmX[q < p.n_problems ? q : 0];
    dz.m[q] = tmDZ[q < p.n_problems ? q : 0];
  }
  if (wgrad_pair(p.taps, p.m_tiles, BLOCK_N, NPL))
    return launch_wgrad_k<BLOCK_N, NPL, kMaxProblems, true>(x, dz, p, stream);
  return launch_wgrad_k<BLOCK_N, NPL, kMaxProblems, false>(x, dz, p, stream);
}

}  // namespace

bool want_pair(int m_tiles, int n_tiles, int block_n, int n_planes, bool multi, int k_iters) {
  if (!pair_enabled() || block_n != 256 || n_planes > 2 || m_tiles < 2) return false;
  // Measured same-box (profiles/r02_pair_ab_session7.txt, r02_pair_default_ab_session8.txt): split modes gain on every eligible launch (layer-8 / layer-9
  // data gradients -5 %, forward neutral); in plain bf16 a pair tile's hand-offs between the two CTAs only pay off on
  // long contractions -- layer 8 gains 3-10 %, the 32-iteration layer-9 tiles lose 5-9 % and stay on single CTAs.
  if (n_planes == 1 && k_iters < 64) return false;
  // One-wave launches (the 250-channel layers and layer 0: 128 tiles, every one with the SAME filter tile) are bound by
  // L2 -> shared-memory bandwidth, two thirds of it the B operand: a pair fetches B once (SPEECHT_B200_PAIR_SMALL=0:
  // only launches of more than one wave pair up)
  const char* e = getenv("SPEECHT_B200_PAIR_SMALL");
  const bool small_ok = e && e[0] == '1';
  return multi || small_ok || m_tiles * n_tiles > st_num_sms();
}

// Filter-gradient launches on CTA pairs: 256-wide tiles, at most two planes, and an even number of (tap, Cin tile)
// combinations per n tile so that tiles 2u and 2u + 1 always share their dZ tile.  Measured same-box
// (profiles/r02_wgrad_pair_ab_session17.txt): plain bf16 gains on every layer (layer 8 -16 %, layer 9 -11 %, the
// 250-channel layers -12 %; step -3.8 % at config 3); in bf16x3 only the one-tap layer 9 gains (-6 %), the others are
// neutral to +2 % and stay on single CTAs.  SPEECHT_B200_WGRAD_PAIR=0 / 1 forces none / all eligible launches.
bool wgrad_pair(int taps, int m_tiles, int block_n, int n_planes) {
  if (block_n != 256 || n_planes > 2 || ((taps * m_tiles) & 1) != 0) return false;
  const char* e = getenv("SPEECHT_B200_WGRAD_PAIR");
  if (e && (e[0] == '0' || e[0] == '1')) return e[0] == '1';
  return n_planes == 1 || taps == 1;
}

void set_reserved_sms(int n) { g_reserved_sms = n < 0 ? 0 : n; }

void set_conv_timeline(long long* buf, int launch_index) {
  g_timeline_buf = buf;
  g_timeline_index = launch_index;
  g_timeline_count = 0;
}

int make_map_3d(CUtensorMap* map, const void* base, int C, int T, int Bn, int64_t ld, int64_t batch_stride,
                int box_c, int box_t) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    st_set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return ST_ERR_CUDA;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)Bn};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)batch_stride * 2};
  const cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_t, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    st_set_error("cuTensorMapEncodeTiled(3d: C=%d T=%d B=%d ld=%lld) failed with CUresult %d", C, T, Bn,
                 (long long)ld, (int)r);
    return ST_ERR_CUDA;
  }
  return ST_OK;
}

// Store map of activation planes [Bn][T][C] for the epilogue: box {32 ch, 32 t, 1}, no swizzle (the staging tile in
// shared memory is dense [32 rows][32 cols] bf16).
int make_map_3d_store(CUtensorMap* map, const void* base, int C, int T, int Bn, int64_t ld, int64_t batch_stride) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    st_set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return ST_ERR_CUDA;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)Bn};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)batch_stride * 2};
  const cuuint32_t box[3] = {32, 32, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    st_set_error("cuTensorMapEncodeTiled(store 3d: C=%d T=%d B=%d ld=%lld) failed with CUresult %d", C, T, Bn,
                 (long long)ld, (int)r);
    return ST_ERR_CUDA;
  }
  return ST_OK;
}

int make_map_2d(CUtensorMap* map, const void* base, int cols, int rows, int64_t ld, int box_c, int box_r) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    st_set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return ST_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_r};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    st_set_error("cuTensorMapEncodeTiled(2d: cols=%d rows=%d ld=%lld) failed with CUresult %d", cols, rows,
                 (long long)ld, (int)r);
    return ST_ERR_CUDA;
  }
  return ST_OK;
}

int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmOut, const ConvParams& p,
                int block_n, int n_planes, cudaStream_t stream) {
  if (block_n == 256 && n_planes == 2) return launch_conv_t<256, 2>(tmA, tmB, tmOut, p, stream);
  if (block_n == 256 && n_planes == 1) return launch_conv_t<256, 1>(tmA, tmB, tmOut, p, stream);
  if (block_n == 32 && n_planes == 2) return launch_conv_t<32, 2>(tmA, tmB, tmOut, p, stream);
  if (block_n == 32 && n_planes == 1) return launch_conv_t<32, 1>(tmA, tmB, tmOut, p, stream);
  if (block_n == 128 && n_planes == 3) return launch_conv_t<128, 3>(tmA, tmB, tmOut, p, stream);
  if (block_n == 128 && n_planes == 2) return launch_conv_t<128, 2>(tmA, tmB, tmOut, p, stream);
  if (block_n == 128 && n_planes == 1) return launch_conv_t<128, 1>(tmA, tmB, tmOut, p, stream);
  if (block_n == 32 && n_planes == 3) return launch_conv_t<32, 3>(tmA, tmB, tmOut, p, stream);
  st_set_error("launch_conv: unsupported (block_n=%d, n_planes=%d)", block_n, n_planes);
  return ST_ERR_UNSUPPORTED;
}

int launch_conv_multi(const CUtensorMap* tmA, const CUtensorMap* tmB, const ConvParams& p, int block_n, int n_planes,
                      cudaStream_t stream) {
  ST_CHECK_ARG(p.n_problems >= 2 && p.n_problems <= kMaxProblems, "launch_conv_multi: 2..%d problems", kMaxProblems);
  ST_CHECK_ARG(!p.bias && !p.relu && !p.out_planes && !p.mask_bits && !p.mask_out && !p.col_sum,
               "launch_conv_multi: fp32 outputs only (no bias / ReLU / mask / planes / column sums)");
  ST_CHECK_ARG(p.k_split <= 1 || (p.taps % p.k_split) == 0, "launch_conv_multi: k_split must divide the taps");
  for (int q = 0; q < p.n_problems; ++q) ST_CHECK_ARG(p.out_f32_q[q] != nullptr, "launch_conv_multi: null output");
  const bool pair = p.pair != 0;
  if (block_n == 256 && n_planes == 2)
    return pair ? launch_conv_multi_t<256, 2, true>(tmA, tmB, p, stream) : launch_conv_multi_t<256, 2, false>(tmA, tmB, p, stream);
  if (block_n == 256 && n_planes == 1)
    return pair ? launch_conv_multi_t<256, 1, true>(tmA, tmB, p, stream) : launch_conv_multi_t<256, 1, false>(tmA, tmB, p, stream);
  st_set_error("launch_conv_multi: unsupported (block_n=%d, n_planes=%d)", block_n, n_planes);
  return ST_ERR_UNSUPPORTED;
}

int launch_wgrad(const CUtensorMap& tmX, const CUtensorMap& tmDZ, const WgradParams& p, int block_n, int n_planes,
                 cudaStream_t stream) {
  if (block_n == 256 && n_planes == 2) return launch_wgrad_t<256, 2>(tmX, tmDZ, p, stream);
  if (block_n == 256 && n_planes == 1) return launch_wgrad_t<256, 1>(tmX, tmDZ, p, stream);
  if (block_n == 64 && n_planes == 2) return launch_wgrad_t<64, 2>(tmX, tmDZ, p, stream);
  if (block_n == 64 && n_planes == 1) return launch_wgrad_t<64, 1>(tmX, tmDZ, p, stream);
  if (block_n == 128 && n_planes == 3) return launch_wgrad_t<128, 3>(tmX, tmDZ, p, stream);
  if (block_n == 128 && n_planes == 2) return launch_wgrad_t<128, 2>(tmX, tmDZ, p, stream);
  if (block_n == 64 && n_planes == 3) return launch_wgrad_t<64, 3>(tmX, tmDZ, p, stream);
  st_set_error("launch_wgrad: unsupported (block_n=%d, n_planes=%d)", block_n, n_planes);
  return ST_ERR_UNSUPPORTED;
}

int launch_wgrad_multi(const CUtensorMap* tmX, const CUtensorMap* tmDZ, const WgradParams& p, int block_n,
                       int n_planes, cudaStream_t stream) {
  ST_CHECK_ARG(p.n_problems >= 2 && p.n_problems <= kMaxProblems, "launch_wgrad_multi: 2..%d problems", kMaxProblems);
  for (int q = 0; q < p.n_problems; ++q)
    ST_CHECK_ARG(p.dW_q[q] != nullptr && p.tap_stride_q[q] >= 1, "launch_wgrad_multi: bad output of problem %d", q);
  if (block_n == 256 && n_planes == 2) return launch_wgrad_multi_t<256, 2>(tmX, tmDZ, p, stream);
  if (block_n == 256 && n_planes == 1) return launch_wgrad_multi_t<256, 1>(tmX, tmDZ, p, stream);
  st_set_error("launch_wgrad_multi: unsupported (block_n=%d, n_planes=%d)", block_n, n_planes);
  return ST_ERR_UNSUPPORTED;
}

int launch_split_input(const float* x, __nv_bfloat16* planes, int B, int T, int Tpad, int F, int n_planes,
                       cudaStream_t stream) {
  ST_CHECK_ARG(F % 8 == 0, "launch_split_input: feature count must be a multiple of 8");
  const int64_t groups = (int64_t)B * Tpad * (F / 8);
  int blocks = (int)((groups + 255) / 256);
  const int cap = 16 * st_num_sms();
  blocks = blocks > cap ? cap : blocks;
  if (n_planes == 3) split_input_kernel<3><<<blocks, 256, 0, stream>>>(x, planes, B, T, Tpad, F);
  else if (n_planes == 2) split_input_kernel<2><<<blocks, 256, 0, stream>>>(x, planes, B, T, Tpad, F);
  else split_input_kernel<1><<<blocks, 256, 0, stream>>>(x, planes, B, T, Tpad, F);
  ST_CUDA_LAUNCH_CHECK("split_input_kernel");
  return ST_OK;
}

int launch_pair_sum_planes(const __nv_bfloat16* x, __nv_bfloat16* xs, int B, int T, int Tx, int ld, int n_planes,
                           cudaStream_t stream) {
  ST_CHECK_ARG(ld % 8 == 0 && n_planes >= 1 && n_planes <= 2, "launch_pair_sum_planes: bad arguments");
  const int64_t groups = (int64_t)B * Tx * (ld / 8);
  int blocks = (int)((groups + 255) / 256);
  const int cap = 16 * st_num_sms();
  blocks = blocks > cap ? cap : blocks;
  if (n_planes == 2) pair_sum_planes_kernel<2><<<blocks, 256, 0, stream>>>(x, xs, B, T, Tx, ld);
  else pair_sum_planes_kernel<1><<<blocks, 256, 0, stream>>>(x, xs, B, T, Tx, ld);
  ST_CUDA_LAUNCH_CHECK("pair_sum_planes_kernel");
  return ST_OK;
}

int launch_ffa_combine(const float* a00, const float* a11, const float* sm, const float* bias, int relu,
                       __nv_bfloat16* out, uint32_t* mask_out, int B, int To, int Tu, int N, int ld_p, int ld_out,
                       int n_planes, cudaStream_t stream) {
  ST_CHECK_ARG(ld_out % 8 == 0 && n_planes >= 1 && n_planes <= 2, "launch_ffa_combine: bad arguments");
  const int64_t groups = (int64_t)B * Tu * (ld_out / 8);
  int blocks = (int)((groups + 255) / 256);
  const int cap = 16 * st_num_sms();
  blocks = blocks > cap ? cap : blocks;
  if (n_planes == 2)
    ffa_combine_kernel<2><<<blocks, 256, 0, stream>>>(a00, a11, sm, bias, relu, out, mask_out, B, To, Tu, N, ld_p, ld_out);
  else
    ffa_combine_kernel<1><<<blocks, 256, 0, stream>>>(a00, a11, sm, bias, relu, out, mask_out, B, To, Tu, N, ld_p, ld_out);
  ST_CUDA_LAUNCH_CHECK("ffa_combine_kernel");
  return ST_OK;
}

int launch_ffa_dz_prep(const __nv_bfloat16* dy, __nv_bfloat16* d00, __nv_bfloat16* d11, __nv_bfloat16* ds, int B,
                       int To, int Tu, int ld, int n_planes, cudaStream_t stream) {
  ST_CHECK_ARG(ld % 8 == 0 && n_planes >= 1 && n_planes <= 2, "launch_ffa_dz_prep: bad arguments");
  const int64_t groups = (int64_t)B * Tu * (ld / 8);
  int blocks = (int)((groups + 255) / 256);
  const int cap = 16 * st_num_sms();
  blocks = blocks > cap ? cap : blocks;
  if (n_planes == 2) ffa_dz_prep_kernel<2><<<blocks, 256, 0, stream>>>(dy, d00, d11, ds, B, To, Tu, ld);
  else ffa_dz_prep_kernel<1><<<blocks, 256, 0, stream>>>(dy, d00, d11, ds, B, To, Tu, ld);
  ST_CUDA_LAUNCH_CHECK("ffa_dz_prep_kernel");
  return ST_OK;
}

int launch_ffa_dx_combine(const float* d_odd, const float* d_even, const float* d_xs, const uint32_t* mask,
                          __nv_bfloat16* out, float* db, int B, int T, int Tx, int N, int ldp, int ld, int n_planes,
                          cudaStream_t stream) {
  ST_CHECK_ARG(ld % 8 == 0 && ld <= 256 && ldp % 4 == 0 && ldp >= ld && n_planes >= 1 && n_planes <= 2,
               "launch_ffa_dx_combine: bad arguments");
  const int64_t rows = (int64_t)B * Tx;
  int blocks = (int)((rows + 15) / 16);
  const int cap = 4 * st_num_sms();
  blocks = blocks > cap ? cap : (blocks < 1 ? 1 : blocks);
  const int rows_per_block = (int)((rows + blocks - 1) / blocks);
  if (n_planes == 2)
    ffa_dx_combine_kernel<2><<<blocks, dim3(32, 8), 0, stream>>>(d_odd, d_even, d_xs, mask, out, db, B, T, Tx, N, ldp,
                                                                 ld, rows_per_block);
  else
    ffa_dx_combine_kernel<1><<<blocks, dim3(32, 8), 0, stream>>>(d_odd, d_even, d_xs, mask, out, db, B, T, Tx, N, ldp,
                                                                 ld, rows_per_block);
  ST_CUDA_LAUNCH_CHECK("ffa_dx_combine_kernel");
  return ST_OK;
}

int launch_ffa_dw_combine(float* dW, const float* cs, int J, int64_t tap_elems, cudaStream_t stream) {
  ST_CHECK_ARG(tap_elems % 4 == 0, "launch_ffa_dw_combine: tap size must be a multiple of 4 floats");
  const int64_t total = (int64_t)J * (tap_elems / 4);
  int blocks = (int)((total + 255) / 256);
  const int cap = 16 * st_num_sms();
  blocks = blocks > cap ? cap : blocks;
  ffa_dw_combine_kernel<<<blocks, 256, 0, stream>>>(dW, cs, J, tap_elems / 4);
  ST_CUDA_LAUNCH_CHECK("ffa_dw_combine_kernel");
  return ST_OK;
}

namespace {
int ew_blocks(int64_t groups) {
  int blocks = (int)((groups + 255) / 256);
  const int cap = 16 * st_num_sms();
  return blocks > cap ? cap : (blocks < 1 ? 1 : blocks);
}
// grid of a background kernel: ONE block per SM when it is to run beside resident tensor-core CTAs
int bg_blocks(int64_t groups, bool background) {
  const int blocks = ew_blocks(groups);
  return background && blocks > st_num_sms() ? st_num_sms() : blocks;
}
}  // namespace

int launch_ffa2_inputs(const __nv_bfloat16* x, __nv_bfloat16* const* s5, int B, int T, int Tq, int ld, int n_planes,
                       cudaStream_t stream) {
  ST_CHECK_ARG(ld % 8 == 0 && n_planes >= 1 && n_planes <= 2, "launch_ffa2_inputs: bad arguments");
  const int blocks = ew_blocks((int64_t)B * Tq * (ld / 8));
  if (n_planes == 2) ffa2_inputs_kernel<2><<<blocks, 256, 0, stream>>>(x, s5[0], s5[1], s5[2], s5[3], s5[4], B, T, Tq, ld);
  else ffa2_inputs_kernel<1><<<blocks, 256, 0, stream>>>(x, s5[0], s5[1], s5[2], s5[3], s5[4], B, T, Tq, ld);
  ST_CUDA_LAUNCH_CHECK("ffa2_inputs_kernel");
  return ST_OK;
}

int launch_ffa2_combine(float* const* part9, const float* bias, int relu, __nv_bfloat16* out, uint32_t* mask_out, int B,
                        int To, int Tq, int N, int ld_p, int ld_out, int n_planes, cudaStream_t stream) {
  ST_CHECK_ARG(ld_out % 8 == 0 && n_planes >= 1 && n_planes <= 2, "launch_ffa2_combine: bad arguments");
  Ptr9c pp;
  for (int l = 0; l < 9; ++l) pp.p[l] = part9[l];
  const int blocks = ew_blocks((int64_t)B * Tq * (ld_out / 8));
  if (n_planes == 2) ffa2_combine_kernel<2><<<blocks, 256, 0, stream>>>(pp, bias, relu, out, mask_out, B, To, Tq, N, ld_p, ld_out);
  else ffa2_combine_kernel<1><<<blocks, 256, 0, stream>>>(pp, bias, relu, out, mask_out, B, To, Tq, N, ld_p, ld_out);
  ST_CUDA_LAUNCH_CHECK("ffa2_combine_kernel");
  return ST_OK;
}

int launch_ffa2_dz_prep(const __nv_bfloat16* dy, __nv_bfloat16* const* out9, int B, int To, int Tq, int ld, int n_planes,
                        cudaStream_t stream) {
  ST_CHECK_ARG(ld % 8 == 0 && n_planes >= 1 && n_planes <= 2, "launch_ffa2_dz_prep: bad arguments");
  Ptr9h pp;
  for (int l = 0; l < 9; ++l) pp.p[l] = out9[l];
  const int blocks = ew_blocks((int64_t)B * Tq * (ld / 4));
  if (n_planes == 2) ffa2_dz_prep_kernel<2><<<blocks, 256, 0, stream>>>(dy, pp, B, To, Tq, ld);
  else ffa2_dz_prep_kernel<1><<<blocks, 256, 0, stream>>>(dy, pp, B, To, Tq, ld);
  ST_CUDA_LAUNCH_CHECK("ffa2_dz_prep_kernel");
  return ST_OK;
}

int launch_ffa2_dx_combine(float* const* g9, const uint32_t* mask, __nv_bfloat16* out, float* db, int B, int T,
                           int Tqx, int N, int ldp, int ld, int n_planes, cudaStream_t stream) {
  ST_CHECK_ARG(ld % 8 == 0 && ld <= 256 && ldp % 4 == 0 && ldp >= ld && n_planes >= 1 && n_planes <= 2,
               "launch_ffa2_dx_combine: bad arguments");
  Ptr9c pp;
  for (int l = 0; l < 9; ++l) pp.p[l] = g9[l];
  const int64_t rows = (int64_t)B * Tqx;
  int blocks = (int)((rows + 7) / 8);
  const int cap = 4 * st_num_sms();
  blocks = blocks > cap ? cap : (blocks < 1 ? 1 : blocks);
  const int rows_per_block = (int)((rows + blocks - 1) / blocks);
  if (n_planes == 2)
    ffa2_dx_combine_kernel<2><<<blocks, dim3(32, 8), 0, stream>>>(pp, mask, out, db, B, T, Tqx, N, ldp, ld, rows_per_block);
  else
    ffa2_dx_combine_kernel<1><<<blocks, dim3(32, 8), 0, stream>>>(pp, mask, out, db, B, T, Tqx, N, ldp, ld, rows_per_block);
  ST_CUDA_LAUNCH_CHECK("ffa2_dx_combine_kernel");
  return ST_OK;
}

int launch_ffa2_dw_combine(float* dW, float* const* c9, int J, int64_t tap_elems, cudaStream_t stream, bool background) {
  ST_CHECK_ARG(tap_elems % 4 == 0, "launch_ffa2_dw_combine: tap size must be a multiple of 4 floats");
  Ptr9c pp;
  for (int l = 0; l < 9; ++l) pp.p[l] = c9[l];
  ffa2_dw_combine_kernel<<<bg_blocks((int64_t)J * (tap_elems / 4), background), 256, 0, stream>>>(dW, pp, J, tap_elems / 4);
  ST_CUDA_LAUNCH_CHECK("ffa2_dw_combine_kernel");
  return ST_OK;
}

// Does a filter-gradient launch of `num_tiles` tiles cut its last wave into K slices (which ACCUMULATE into dW, so the
// outputs must be zero on entry)?  Mirrors the wave-aligned split of tc_wgrad_kernel.
int wgrad_best_split(int num_tiles, int total_iters, bool pair) {
  int G = tc_sms();
  if (pair) { G /= 2; num_tiles /= 2; }          // scheduling units: tile pairs on clusters
  if (num_tiles >= G) return 1;
  int best = 1;
  int64_t best_cost = (int64_t)((num_tiles + G - 1) / G) * total_iters;
  for (int s = 2; s <= 8 && s * 4 <= total_iters; ++s) {
    // makespan in iterations + a charge for the extra accumulation epilogues
    const int64_t cost = (int64_t)((num_tiles * s + G - 1) / G) * ((total_iters + s - 1) / s + 4);
    if (cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

bool wgrad_accumulates(int num_tiles, int total_iters, bool pair) {
  int G = tc_sms();
  if (pair) { G /= 2; num_tiles /= 2; }
  const int tail_tiles = num_tiles % G;
  if (tail_tiles == 0) return false;
  const int cap = total_iters / 4 > 0 ? total_iters / 4 : 1;
  const int split = G / tail_tiles < cap ? G / tail_tiles : cap;
  return split > 1;
}

int launch_pack_filters(PackTable& tab, int n_planes, cudaStream_t stream, int* launches) {
  *launches = 0;
  int blocks = 0;
  for (int l = 0; l < tab.n; ++l) {
    PackEntry& e = tab.e[l];
    ST_CHECK_ARG(e.cin_p % 64 == 0 && e.ld_co % 4 == 0 && e.ld_co >= e.Cout, "launch_pack_filters: bad padded sizes");
    e.blk0 = blocks;
    blocks += e.K * (e.cin_p / 64) * ((e.ld_co + 63) / 64);
  }
  if (n_planes == 3) pack_filter_both_kernel<3><<<blocks, dim3(32, 8), 0, stream>>>(tab);
  else if (n_planes == 2) pack_filter_both_kernel<2><<<blocks, dim3(32, 8), 0, stream>>>(tab);
  else pack_filter_both_kernel<1><<<blocks, dim3(32, 8), 0, stream>>>(tab);
  ST_CUDA_LAUNCH_CHECK("pack_filter_both_kernel");
  *launches = 1;
  return ST_OK;
}

int launch_pack_ffa2(const float* w, __nv_bfloat16* const* fwd9, __nv_bfloat16* const* bwd9, const int* masks9, int J,
                     int Cin, int Cout, int cin_p, int ld_co, int n_planes, cudaStream_t stream) {
  ST_CHECK_ARG(cin_p % 64 == 0 && ld_co % 4 == 0 && ld_co >= Cout && n_planes >= 1 && n_planes <= 2,
               "launch_pack_ffa2: bad arguments");
  Ptr9h f, b;
  LeafMasks m;
  for (int l = 0; l < 9; ++l) { f.p[l] = fwd9[l]; b.p[l] = bwd9[l]; m.m[l] = masks9[l]; }
  const int blocks = J * (cin_p / 64) * ((ld_co + 63) / 64);
  const int smem = 4 * 64 * 65 * (int)sizeof(float);
  if (n_planes == 2) {
    ST_CUDA_CALL(cudaFuncSetAttribute(pack_ffa2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pack_ffa2_kernel<2><<<blocks, dim3(32, 8), smem, stream>>>(w, f, b, m, J, Cin, Cout, cin_p, ld_co);
  } else {
    ST_CUDA_CALL(cudaFuncSetAttribute(pack_ffa2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pack_ffa2_kernel<1><<<blocks, dim3(32, 8), smem, stream>>>(w, f, b, m, J, Cin, Cout, cin_p, ld_co);
  }
  ST_CUDA_LAUNCH_CHECK("pack_ffa2_kernel");
  return ST_OK;
}

int launch_zero_f32(float* dst, int64_t n, cudaStream_t stream, bool background) {
  ST_CHECK_ARG((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && n % 4 == 0, "launch_zero_f32: 16-byte granularity");
  zero_f32_kernel<<<bg_blocks(n / 4, background), 256, 0, stream>>>(reinterpret_cast<float4*>(dst), n / 4);
  ST_CUDA_LAUNCH_CHECK("zero_f32_kernel");
  return ST_OK;
}

int launch_pack_bwd(const float* w, __nv_bfloat16* const* bwd, const int* masks, int n_leaves, int group, int J, int Cin,
                    int Cout, int ld_co, int n_planes, cudaStream_t stream, bool background) {
  ST_CHECK_ARG(n_leaves >= 1 && n_leaves <= 9 && group >= 1 && group <= 4 && Cout % 8 == 0 && ld_co % 8 == 0 &&
               ld_co >= Cout && n_planes >= 1 && n_planes <= 2, "launch_pack_bwd: bad arguments");
  Ptr9h b;
  LeafMasks m;
  for (int l = 0; l < 9; ++l) { b.p[l] = bwd[l < n_leaves ? l : 0]; m.m[l] = masks[l < n_leaves ? l : 0]; }
  const int blocks = bg_blocks((int64_t)J * Cin * (ld_co / 4), background);
  if (n_planes == 2) pack_bwd_kernel<2><<<blocks, 256, 0, stream>>>(w, b, m, n_leaves, group, J, Cin, Cout, ld_co);
  else pack_bwd_kernel<1><<<blocks, 256, 0, stream>>>(w, b, m, n_leaves, group, J, Cin, Cout, ld_co);
  ST_CUDA_LAUNCH_CHECK("pack_bwd_kernel");
  return ST_OK;
}

int launch_bias_grad(const __nv_bfloat16* dz, int64_t rows, int N, int ld, int n_planes, float* db,
                     cudaStream_t stream) {
  // db must be zero on entry (the plan zeroes the whole flat gradient buffer once per backward)
  int chunks = (int)((rows + 63) / 64);
  const int cap = 8 * st_num_sms() / ((ld + 255) / 256);
  chunks = chunks > cap ? cap : (chunks < 1 ? 1 : chunks);
  const int rows_per_block = (int)((rows + chunks - 1) / chunks);
  dim3 grid((ld + 255) / 256, chunks);
  bias_grad_planes_kernel<<<grid, dim3(32, 8), 0, stream>>>(dz, rows, N, ld, n_planes, db, rows_per_block);
  ST_CUDA_LAUNCH_CHECK("bias_grad_planes_kernel");
  return ST_OK;
}

int launch_merge_planes(const __nv_bfloat16* planes, int64_t rows, int cols, int ld, int n_planes, float* dst,
                        int64_t ld_dst, cudaStream_t stream) {
  int blocks = (int)((rows * cols + 255) / 256);
  const int cap = 16 * st_num_sms();
  blocks = blocks > cap ? cap : (blocks < 1 ? 1 : blocks);
  merge_planes_kernel<<<blocks, 256, 0, stream>>>(planes, rows, cols, ld, n_planes, dst, ld_dst);
  ST_CUDA_LAUNCH_CHECK("merge_planes_kernel");
  return ST_OK;
}

}  // namespace tc
