// Tensor-core (tcgen05 + TMEM) convolution kernel for the TemporalUNet levels with horizon <= 13
// (94 % of the network's multiply-accumulates).
//
// GEMM view ("rows as M"):  D[row, (l_out, c_out)] = sum_{(l_in, c_in)} X[row, (l_in, c_in)] * W
//   * M = 128 trajectory rows per CTA (UMMA_M = 128, cta_group::1), accumulator in TMEM;
//   * the K axis is ordered (l_in, c_in) and cut into 128-byte chunks (32 tf32 or 64 bf16 input
//     channels of ONE input position l_in);
//   * the N axis of a CTA is ordered (l_out, channel-in-tile); a K chunk at l_in only touches the
//     output positions within the filter's reach, i.e. a contiguous column window of the
//     accumulator, so each chunk is ONE windowed tcgen05.mma (N_w = #positions x channels) and no
//     zero-padding tap is ever multiplied (exactly the non-padding MACs);
//   * operands are pre-tiled in global memory as ready-made UMMA shared-memory images (K-major; weight tiles 128-byte
//     rows with SWIZZLE_128B, activation blocks 128 rows x 128 B chunk-major without swizzle, see tc_swz_bytes below):
//     producer threads move them with 1-D bulk async copies (TMA engine) signalled on mbarriers -- no tensor maps;
//   * fp32 fidelity: every operand exists as a "hi" part and a "lo" remainder (TF32 or BF16) and
//     each product is three MMAs (lo*hi + hi*lo + hi*hi): "3xTF32" (~2^-22 operand error) or
//     "3xBF16" (~2^-17).  The single-pass modes use the hi part only;
//   * warp roles: warp 0 + lane 0 of three epilogue warps = bulk-copy producers (one thread per
//     operand stream), warp 1 = TMEM allocator + MMA issuer (warp-uniform schedule walk, one
//     elected lane issues), warps 2..17 = epilogue (TMEM -> registers; GroupNorm over the CTA's
//     own columns: a thread owns one row, so group statistics are private serial reductions that
//     meet in shared memory; Mish; time embedding / residual; hi/lo split; results staged in shared
//     memory and written out in the tiled format the next layer consumes).
//
// Reference ops covered: Conv1dBlock (blocks.py:13-34), ResidualConvolutionBlock (:137-166) incl.
// the 1x1 residual conv (second accumulator), stride-2 Conv1d (:211), ConvTranspose1d (:249).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace edmp {

// Trace-mode ablation (tools/tc_trace.py with EDMP_ABLATE, read only when the dbg buffer is set -- never on the product path):
// bit 0 = the epilogue only does its barrier handshakes, bit 1 = the MMA warp issues no MMAs (commits only),
// bit 2 = conv_tc2's producer copies nothing (every stage is "full" at once: MMAs on stale shared memory).
static __device__ int g_edmp_ablate;


constexpr int kTcRows = 128;                 // rows per CTA tile (UMMA M)
constexpr int kTcBlockBytes = kTcRows * 128; // bytes of one tiled operand block (128 rows x 128 B)
constexpr int kTcMaxLin = 13;
constexpr int kTcEpiWarps = 16;              // epilogue warps: 4 per 32-lane TMEM quarter
constexpr int kTcEpiThreads = kTcEpiWarps * 32;
constexpr int kTcThreads = 64 + kTcEpiThreads;  // producer warp, MMA warp, epilogue warps

struct TcOperand {      // tiled activation: blocks [row_tile][l * C/cpc + c/cpc][128 rows x 128 B swizzled]
  const void* hi;
  const void* lo;
  int C;
};

struct TcSched {        // what one K chunk at input position l_in contributes to
  int8_t slot_begin;    // first weight slot (row block of the packed weight tile)
  int8_t n_slots;       // number of consecutive output positions touched
  int8_t lo_begin;      // first output position
  int8_t pad;
};

struct TcPhase {
  TcOperand a, b;       // b.C == 0 unless the input is a skip concat (blocks.py:253)
  const void* w_hi;     // packed weights [n_tile][c_chunk][slots * ct rows][128 B] swizzled
  const void* w_lo;
  float acc_scale;      // exact inverse of the power-of-two weight scale (1 unless half operands)
  int lin;              // input positions
  int slots;            // slots stored per weight tile
  int d_col;            // accumulator column base
  TcSched sched[kTcMaxLin];
};

enum TcMode { TC_BIAS = 0, TC_GN = 1, TC_GN_RES_ID = 2, TC_GN_RES_PW = 3 };

struct TcArgs {
  TcPhase ph[2];
  int n_phases;
  int rows, lout, ct, cout;     // ct = channels per CTA column tile; N = lout * ct
  int cg;                       // channels per GroupNorm group (ct / cg groups per column tile, 1 or 2)
  int mode;
  int split;                    // 1 = three-MMA hi/lo split, 0 = single pass
  int a_stages, b_stages;
  int epi_units;                // 16-column units staged in shared memory per epilogue round (multiple of 4)
  int stage_bytes;              // bytes reserved for operand stages / epilogue staging (barriers follow)
  const float *bias, *gamma, *beta, *temb, *bres;
  TcOperand res;                // identity residual source (tiled)
  void *out_hi, *out_lo;        // tiled output, may be null
  int out_pm;                   // 1: out_hi/out_lo are position-major images (conv_pm.cuh) of (cout, lout)
  float* out_plain;             // plain [rows][cout][lout], may be null
  long long* dbg;               // optional [ctas][8] clock64 stamps (tools/tc_trace.py), normally null
};

// Element-type traits: TF32 operands are 4-byte floats (32 per 128-byte row), BF16 / F16 2-byte (64 per row)
enum TcEl { TC_EL_TF32 = 0, TC_EL_BF16 = 1, TC_EL_F16 = 2 };
template <int EL> struct TcElem {
  static constexpr bool k16 = true;
  static constexpr int kCpc = 64;
  static constexpr int kShift = 6;
  static constexpr int kFmt = (EL == TC_EL_BF16) ? 1 : 0;   // UMMA a/b format: 1 = BF16, 0 = F16
  static constexpr int kUnitChunks = 2;
};
template <> struct TcElem<TC_EL_TF32> {
  static constexpr bool k16 = false;
  static constexpr int kCpc = 32;      // channels per K chunk
  static constexpr int kShift = 5;
  static constexpr int kFmt = 2;       // UMMA a/b format TF32
  static constexpr int kUnitChunks = 4;  // 16-byte chunks that 16 channels occupy
};

// Byte offset of 16-byte chunk `chunk16` (0..7) of row `row_local` inside a tiled activation block.  Blocks are CHUNK-MAJOR
// [chunk][128 rows][16 B] -- the un-swizzled ("interleaved") K-major UMMA layout, core matrix = 8 rows x 16 B -- so that
// an epilogue warp (a thread = a row) touches 512 contiguous bytes per 16-byte store / residual load instead of 32
// different 128-byte lines as with row-major SWIZZLE_128B rows (the L1 tag stage, one line per cycle, was what made the
// identity-residual loads cost +10k cycles per tile).  Weight tiles keep SWIZZLE_128B.
constexpr int kTcChunkStride = kTcRows * 16;   // bytes between consecutive 16-byte K chunks of a block (UMMA LBO)
__device__ __forceinline__ int tc_swz_bytes(int row_local, int chunk16) {
  return chunk16 * kTcChunkStride + row_local * 16;
}
// UMMA descriptor of a tiled activation block in shared memory (without the start address)
__device__ __forceinline__ uint64_t tc_act_desc0() { return umma::make_desc_interleaved(0, kTcChunkStride, 128); }
constexpr int kTcActKStep = (2 * kTcChunkStride) >> 4;   // descriptor start-address increment per 32-byte K step

// Position-major ACTIVATION images (conv_pm*.cuh) are chunk-major like the tiled blocks above: per block of 8 trajectory
// rows [16-byte chunk][position -2 .. L+1][8 rows][16 B] -- the un-swizzled K-major UMMA layout whose core matrix is the
// 8 rows of one position; LBO = (L + 4) * 128 B (next chunk), SBO = 128 B (next position; doubled for a stride-2 conv),
// a filter tap = a start-address offset of 128 B.  An epilogue warp (4 positions x 8 rows) then writes 512 contiguous
// bytes per 16-byte store instead of 32 separate lines.  pm_act_off: byte offset of (position p, row r, chunk) in an image.
__device__ __forceinline__ size_t pm_act_off(int L, int p, int r, int chunk) {
  return ((size_t)chunk * (L + 4) + (size_t)(p + 2)) * 128 + (size_t)r * 16;
}
__device__ __forceinline__ size_t pm_img_bytes(int L, int C) { return (size_t)(L + 4) * 16 * C; }   // hi (or lo) part of one row block

// swizzle of 16-byte chunk `c` in row `r` of a position-major WEIGHT atom with `rby`-byte rows (32 / 64 / 128)
__device__ __forceinline__ int pm_swz(int rby, int r, int c) {
  return rby == 128 ? (c ^ (r & 7)) : (rby == 64 ? (c ^ ((r >> 1) & 3)) : (c ^ ((r >> 2) & 1)));
}

// Mish with hardware exp2 / reciprocal approximations (rel. error ~1e-6, below the split-MMA
// accumulation error); the CUDA-core path keeps the accurate version.
__device__ __forceinline__ float mish_fast(float x) {
  const float e = __expf(fminf(x, 20.0f));
  const float n = e * (e + 2.0f);
  return x > 20.0f ? x : x * __fdividef(n, n + 2.0f);
}

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo_f(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi_f(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  const __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
// two 16-bit elements (BF16 or IEEE half) <-> two floats
template <int EL> __device__ __forceinline__ uint32_t pack16x2(float a, float b) {
  return EL == TC_EL_BF16 ? pack_bf16x2(a, b) : pack_f16x2(a, b);
}
template <int EL> __device__ __forceinline__ float2 unpack16x2(uint32_t v) {
  if (EL == TC_EL_BF16) return make_float2(bf16_lo_f(v), bf16_hi_f(v));
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

// IEEE-half operand range (f16x3 / f16 modes): activations are split hi + lo UN-scaled, so a value beyond
// +-65504 becomes an infinite hi part and the forward silently turns to NaN.  The epilogues keep the largest |hi| bit
// pattern they store (two halves per word: one LOP3 + one VIMNMX per pair) and raise a device flag at the end of the
// kernel if it reached the infinity / NaN patterns; edmp_unet_range_status reads it.  BF16 / TF32 have fp32's range.
template <int EL> __device__ __forceinline__ void range_track(uint32_t& hmax, uint32_t packed_hi) {
  if (EL == TC_EL_F16) hmax = __vmaxu2(hmax, packed_hi & 0x7FFF7FFFu);
}
template <int EL> __device__ __forceinline__ void range_report(uint32_t hmax, unsigned* flag) {
  if (EL == TC_EL_F16 && flag != nullptr && ((hmax & 0xFFFFu) >= 0x7C00u || (hmax >> 16) >= 0x7C00u)) atomicOr(flag, 1u);
}

// split 16 fp32 values into hi/lo parts in the operand element type; writes kUnitChunks 16-byte chunks each
template <int EL>
__device__ __forceinline__ void tc_split_store(const float (&v)[16], bool want_lo, uint4* hi, uint4* lo) {
  if (EL != TC_EL_TF32) {
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      uint32_t h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float x0 = v[m * 8 + 2 * e], x1 = v[m * 8 + 2 * e + 1];
        h[e] = pack16x2<EL>(x0, x1);
        const float2 hf = unpack16x2<EL>(h[e]);
        l[e] = want_lo ? pack16x2<EL>(x0 - hf.x, x1 - hf.y) : 0u;
      }
      hi[m] = make_uint4(h[0], h[1], h[2], h[3]);
      lo[m] = make_uint4(l[0], l[1], l[2], l[3]);
    }
  } else {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      float h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float x = v[m * 4 + e];
        h[e] = want_lo ? umma::to_tf32(x) : x;
        l[e] = want_lo ? umma::to_tf32(x - h[e]) : 0.0f;
      }
      hi[m] = make_uint4(__float_as_uint(h[0]), __float_as_uint(h[1]), __float_as_uint(h[2]), __float_as_uint(h[3]));
      lo[m] = make_uint4(__float_as_uint(l[0]), __float_as_uint(l[1]), __float_as_uint(l[2]), __float_as_uint(l[3]));
    }
  }
}

// 16-bit element types: split 16 fp32 values held as 8 packed pairs into hi/lo parts; writes two 16-byte chunks each
template <int EL>
__device__ __forceinline__ void tc_split_store2(const f2::f32x2 (&v2)[8], bool want_lo, uint4* hi, uint4* lo) {
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float x0, x1, d0, d1;
      f2::upk(v2[m * 4 + e], x0, x1);
      h[e] = pack16x2<EL>(x0, x1);
      const float2 hf = unpack16x2<EL>(h[e]);
      f2::upk(f2::sub(v2[m * 4 + e], f2::pk(hf.x, hf.y)), d0, d1);
      l[e] = want_lo ? pack16x2<EL>(d0, d1) : 0u;
    }
    hi[m] = make_uint4(h[0], h[1], h[2], h[3]);
    lo[m] = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// sum of the hi and lo 16-byte chunks as fp32 values (4 for TF32, 8 for BF16)
template <int EL>
__device__ __forceinline__ void tc_chunk_sum(const uint4& h, const uint4& l, bool has_lo, float* out) {
  const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
  if (EL != TC_EL_TF32) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 a = unpack16x2<EL>(hh[e]), b = unpack16x2<EL>(ll[e]);
      out[2 * e] = a.x + (has_lo ? b.x : 0.0f);
      out[2 * e + 1] = a.y + (has_lo ? b.y : 0.0f);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) out[e] = __uint_as_float(hh[e]) + (has_lo ? __uint_as_float(ll[e]) : 0.0f);
  }
}

template <int EL>
__global__ void __launch_bounds__(kTcThreads, 1) conv_tc_kernel(const __grid_constant__ TcArgs a) {
  using E = TcElem<EL>;
  constexpr bool BF16 = E::k16;   // 16-bit operand elements (BF16 or IEEE half)
  constexpr int UC = E::kUnitChunks;           // 16-byte chunks per (row, 16-channel unit)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A stages][B stages] ... [barriers]; statically: parameters and GroupNorm partials
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nparts = a.split ? 2 : 1;
  const int a_stage_bytes = kTcBlockBytes * nparts;
  int max_slots = a.ph[0].slots;
  if (a.n_phases > 1 && a.ph[1].slots > max_slots) max_slots = a.ph[1].slots;
  const int b_part_bytes = max_slots * a.ct * 128;
  const int b_stage_bytes = b_part_bytes * nparts;
  uint8_t* a_smem = smem;
  uint8_t* b_smem = a_smem + a.a_stages * a_stage_bytes;
  uint64_t* bars = (uint64_t*)(smem + a.stage_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + a.a_stages;
  uint64_t* b_full = a_empty + a.a_stages;
  uint64_t* b_empty = b_full + a.b_stages;
  uint64_t* acc_full = b_empty + a.b_stages;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);
  __shared__ __align__(16) float s_par[5 * 64];   // bias | gamma | beta | temb | bres for this column tile
  __shared__ float s_stat[2 * 4 * 2 * 128];        // [pass][part][group][row] partial GroupNorm sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rt = blockIdx.x, nt = blockIdx.y;
  long long* dbg = a.dbg ? a.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    // one arrival per producer thread on the "full" barriers (hi and lo parts have their own thread)
    for (int i = 0; i < a.a_stages; ++i) { umma::mbar_init(a_full + i, nparts); umma::mbar_init(a_empty + i, 1); }
    for (int i = 0; i < a.b_stages; ++i) { umma::mbar_init(b_full + i, nparts); umma::mbar_init(b_empty + i, 1); }
    umma::mbar_init(acc_full, 1);
    umma::fence_barrier_init();
  }
  if (warp == 1) umma::tmem_alloc<512>(tmem_slot);
  if (warp >= 2) {
    const int e = threadIdx.x - 64;
    if (e < a.ct) {
      const int c = nt * a.ct + e;
      s_par[e] = a.bias[c];
      s_par[64 + e] = a.gamma ? a.gamma[c] : 1.0f;
      s_par[128 + e] = a.beta ? a.beta[c] : 0.0f;
      s_par[192 + e] = a.temb ? a.temb[c] : 0.0f;
      s_par[256 + e] = a.bres ? a.bres[c] : 0.0f;
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel; activations are read below
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();

  // ===== producers: bulk async copies of ready-made operand tiles.  One thread can keep only ~2
  // bulk copies in flight (~20 B/cycle), separate threads scale linearly (tools/micro/bulk_bw.cu), so
  // the activation-hi, activation-lo, weight-hi and weight-lo streams each get their own thread:
  // lane 0 of warp 0 and of the first three epilogue warps (idle until the accumulator is ready).
  int prod_kind = -1;   // 0: A hi, 1: A lo, 2: B hi, 3: B lo
  if (lane == 0) {
    if (warp == 0) prod_kind = 0;
    else if (warp == 2 && a.split) prod_kind = 1;
    else if (warp == 3) prod_kind = 2;
    else if (warp == 4 && a.split) prod_kind = 3;
  }
  if (prod_kind >= 0) {
    const bool is_lo = prod_kind & 1;
    int a_it = 0, b_it = 0;
    for (int p = 0; p < a.n_phases; ++p) {
      const TcPhase& ph = a.ph[p];
      const int ka = ph.a.C >> E::kShift, kb = ph.b.C >> E::kShift;
      const uint32_t wtile_bytes = (uint32_t)ph.slots * a.ct * 128u;
      for (int cc = 0; cc < ka + kb; ++cc) {
        if (prod_kind >= 2) {
          const int bs = b_it % a.b_stages;
          umma::mbar_wait(b_empty + bs, ((b_it / a.b_stages) & 1) ^ 1);
          umma::mbar_arrive_expect_tx(b_full + bs, wtile_bytes);
          const size_t woff = ((size_t)nt * (ka + kb) + cc) * wtile_bytes;
          uint8_t* dst = b_smem + bs * b_stage_bytes + (is_lo ? b_part_bytes : 0);
          umma::bulk_g2s(dst, (const uint8_t*)(is_lo ? ph.w_lo : ph.w_hi) + woff, wtile_bytes, b_full + bs);
          ++b_it;
        } else {
          for (int li = 0; li < ph.lin; ++li) {
            if (ph.sched[li].n_slots == 0) continue;
            const int as = a_it % a.a_stages;
            umma::mbar_wait(a_empty + as, ((a_it / a.a_stages) & 1) ^ 1);
            umma::mbar_arrive_expect_tx(a_full + as, (uint32_t)kTcBlockBytes);
            const bool first = cc < ka;
            const TcOperand& op = first ? ph.a : ph.b;
            const int c2 = first ? cc : cc - ka;
            const int kop = op.C >> E::kShift;
            const size_t blk = ((size_t)rt * (ph.lin * kop) + (size_t)li * kop + c2) * kTcBlockBytes;
            uint8_t* dst = a_smem + as * a_stage_bytes + (is_lo ? kTcBlockBytes : 0);
            umma::bulk_g2s(dst, (const uint8_t*)(is_lo ? op.lo : op.hi) + blk, kTcBlockBytes, a_full + as);
            ++a_it;
          }
        }
      }
    }
  }

  if (warp == 0) {
    // producer warp: nothing else to do
  } else if (warp == 1) {
    // ===== MMA issuer.  The whole warp walks the schedule so that every descriptor is computed on
    // the uniform datapath; only the tcgen05 instructions themselves are issued by one elected lane
    // (a divergent single-lane loop costs a ~14-instruction R2UR "waterfall" per MMA).  Measured
    // cost of one M=128 SS-mode MMA is ~40 + N/2 cycles for tf32 (K=8) and bf16 (K=16) alike
    // (tools/micro/mma_rate.cu), so bf16 operands halve the mainloop. =====
    int a_it = 0, b_it = 0;
    long long wait_a = 0, wait_b = 0;
    const uint64_t desc0 = umma::make_desc_sw128(0);
    for (int p = 0; p < a.n_phases; ++p) {
      const TcPhase& ph = a.ph[p];
      const int kc_total = (ph.a.C + ph.b.C) >> E::kShift;
      uint32_t touched = 0;   // output positions whose accumulator columns already hold a partial sum
      for (int cc = 0; cc < kc_total; ++cc) {
        const int bs = b_it % a.b_stages;
        { const long long tw = dbg ? clock64() : 0; umma::mbar_wait(b_full + bs, (b_it / a.b_stages) & 1); if (dbg) wait_b += clock64() - tw; }
        if (dbg && b_it == 0 && lane == 0) dbg[2] = clock64();
        const uint32_t b_base = umma::smem_u32(b_smem + bs * b_stage_bytes);
        for (int li = 0; li < ph.lin; ++li) {
          const TcSched s = ph.sched[li];
          if (s.n_slots == 0) continue;
          const int as = a_it % a.a_stages;
          { const long long tw = dbg ? clock64() : 0; umma::mbar_wait(a_full + as, (a_it / a.a_stages) & 1); if (dbg) wait_a += clock64() - tw; }
          umma::tc_fence_after();
          if (dbg && a_it == 0 && lane == 0) dbg[3] = clock64();
          const uint32_t a_base = umma::smem_u32(a_smem + as * a_stage_bytes);
          const uint64_t da_hi = tc_act_desc0() | (uint64_t)((a_base & 0x3FFFF) >> 4);
          const uint64_t da_lo = tc_act_desc0() | (uint64_t)(((a_base + kTcBlockBytes) & 0x3FFFF) >> 4);
          // first K chunk: the window's positions may differ in "already written", so issue one
          // MMA per position with its own accumulate flag; afterwards one windowed MMA.
          const int n_issue = (cc == 0) ? s.n_slots : 1;
          const int n_cols = ((cc == 0) ? 1 : s.n_slots) * a.ct;
          const uint32_t idesc = umma::make_idesc(E::kFmt, kTcRows, n_cols);
          for (int q = 0; q < n_issue; ++q) {
            const int lo = s.lo_begin + q;
            const uint32_t acc0 = (cc == 0) ? ((touched >> lo) & 1u) : 1u;
            const uint32_t d = tmem_base + (uint32_t)(ph.d_col + lo * a.ct);
            const uint32_t b_off = b_base + (uint32_t)((s.slot_begin + q) * a.ct * 128);
            const uint64_t db_hi = desc0 | (uint64_t)((b_off & 0x3FFFF) >> 4);
            const uint64_t db_lo = desc0 | (uint64_t)(((b_off + b_part_bytes) & 0x3FFFF) >> 4);
            if (umma::elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {   // 32-byte K steps inside the 128-byte swizzle atom: +2 in the address field
                const uint32_t acc = acc0 | (uint32_t)(ks > 0);
                if (BF16) {
                  if (a.split) {
                    umma::mma_bf16(d, da_lo + kTcActKStep * ks, db_hi + 2 * ks, idesc, acc);
                    umma::mma_bf16(d, da_hi + kTcActKStep * ks, db_lo + 2 * ks, idesc, 1u);
                    umma::mma_bf16(d, da_hi + kTcActKStep * ks, db_hi + 2 * ks, idesc, 1u);
                  } else {
                    umma::mma_bf16(d, da_hi + kTcActKStep * ks, db_hi + 2 * ks, idesc, acc);
                  }
                } else {
                  if (a.split) {
                    umma::mma_tf32(d, da_lo + kTcActKStep * ks, db_hi + 2 * ks, idesc, acc);
                    umma::mma_tf32(d, da_hi + kTcActKStep * ks, db_lo + 2 * ks, idesc, 1u);
                    umma::mma_tf32(d, da_hi + kTcActKStep * ks, db_hi + 2 * ks, idesc, 1u);
                  } else {
                    umma::mma_tf32(d, da_hi + kTcActKStep * ks, db_hi + 2 * ks, idesc, acc);
                  }
                }
              }
            }
            __syncwarp();
            if (cc == 0) touched |= 1u << lo;
          }
          if (umma::elect_one()) umma::mma_commit(a_empty + as);   // frees the A stage once these MMAs have read it
          __syncwarp();
          ++a_it;
        }
        if (umma::elect_one()) umma::mma_commit(b_empty + bs);
        __syncwarp();
        ++b_it;
      }
    }
    if (umma::elect_one()) umma::mma_commit(acc_full);
    __syncwarp();
    if (dbg && lane == 0) { dbg[4] = clock64(); dbg[8] = wait_a; dbg[9] = wait_b; dbg[10] = a_it; dbg[11] = b_it; }
  } else {
    // ===== epilogue: 16 warps; a thread owns one accumulator lane (trajectory row) and every 4th
    // 16-column unit of it (four warps share each 32-lane TMEM quarter) =====
    constexpr int kParts = kTcEpiWarps / 4;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int part = (warp - 2) >> 2;             // which share of the column units (0..kParts-1)
    const int row_local = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int N = a.lout * a.ct;
    const int n_units = N >> 4;
    const int gpt = a.ct / a.cg;                  // GroupNorm groups in this column tile (1 or 2)
    umma::mbar_wait(acc_full, 0);
    __syncwarp();
    umma::tc_fence_after();
    if (dbg && threadIdx.x == 64) dbg[5] = clock64();

    float mean[2] = {0.0f, 0.0f}, rstd[2] = {1.0f, 1.0f};
    const float sc0 = a.ph[0].acc_scale, sc1 = a.ph[1].acc_scale;   // exact power-of-two de-scaling of the weights
    if (a.mode != TC_BIAS) {
      // GroupNorm(8) over (cg channels x lout positions) of this row (blocks.py:24-26): two-pass,
      // partial sums of the column shares meet in shared memory
      const float inv_n = 1.0f / (float)(a.cg * a.lout);
      float s[2] = {0.0f, 0.0f};
      for (int u = part; u < n_units; u += kParts) {
        float v[16];
        umma::tmem_ld16(t_lane + a.ph[0].d_col + u * 16, v);
        const int c0 = (u * 16) % a.ct;
        if (gpt == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) s[0] += fmaf(v[i], sc0, s_par[c0 + i]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) { s[0] += fmaf(v[i], sc0, s_par[c0 + i]); s[1] += fmaf(v[8 + i], sc0, s_par[c0 + 8 + i]); }
        }
      }
      s_stat[(part * 2 + 0) * 128 + row_local] = s[0];
      s_stat[(part * 2 + 1) * 128 + row_local] = s[1];
      epi_barrier();
      {
        float t0 = 0.0f, t1 = 0.0f;
#pragma unroll
        for (int q = 0; q < kParts; ++q) { t0 += s_stat[(q * 2 + 0) * 128 + row_local]; t1 += s_stat[(q * 2 + 1) * 128 + row_local]; }
        mean[0] = t0 * inv_n;
        mean[1] = t1 * inv_n;
      }
      float ss[2] = {0.0f, 0.0f};
      for (int u = part; u < n_units; u += kParts) {
        float v[16];
        umma::tmem_ld16(t_lane + a.ph[0].d_col + u * 16, v);
        const int c0 = (u * 16) % a.ct;
        if (gpt == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float d = fmaf(v[i], sc0, s_par[c0 + i]) - mean[0]; ss[0] = fmaf(d, d, ss[0]); }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float d0 = fmaf(v[i], sc0, s_par[c0 + i]) - mean[0];
            const float d1 = fmaf(v[8 + i], sc0, s_par[c0 + 8 + i]) - mean[1];
            ss[0] = fmaf(d0, d0, ss[0]);
            ss[1] = fmaf(d1, d1, ss[1]);
          }
        }
      }
      float* s2 = s_stat + kParts * 2 * 128;
      s2[(part * 2 + 0) * 128 + row_local] = ss[0];
      s2[(part * 2 + 1) * 128 + row_local] = ss[1];
      epi_barrier();
      {
        float t0 = 0.0f, t1 = 0.0f;
#pragma unroll
        for (int q = 0; q < kParts; ++q) { t0 += s2[(q * 2 + 0) * 128 + row_local]; t1 += s2[(q * 2 + 1) * 128 + row_local]; }
        rstd[0] = rsqrtf(t0 * inv_n + 1e-5f);
        rstd[1] = rsqrtf(t1 * inv_n + 1e-5f);
      }
    }
    // Results are staged in shared memory (the operand stages are free once the accumulator is
    // complete) and written out cooperatively so every store instruction covers whole sectors:
    // a thread-per-row store pattern would touch 32 different 128-byte lines per instruction.
    // Staging per 16-column unit: [128 rows][UC x 16 B] hi, same for lo, [128 rows][64 B] residual.
    const int kch_out = a.cout >> E::kShift;
    const int et = threadIdx.x - 64;               // index among the epilogue threads
    const int plain_stride = N + 1;                // odd row stride: conflict-free column writes
    uint8_t* stg_hi = a_smem;
    uint8_t* stg_lo = stg_hi + (size_t)a.epi_units * (128 * UC * 16);
    float* stg_res = reinterpret_cast<float*>(stg_lo + (size_t)a.epi_units * (128 * UC * 16));
    float* stg_plain = reinterpret_cast<float*>(a_smem);
    const int sw_out = (UC == 4) ? ((row_local >> 1) & 3) : ((row_local >> 2) & 1);
    for (int u0 = 0; u0 < n_units; u0 += a.epi_units) {
      const int u1 = min(n_units, u0 + a.epi_units);
      if (a.mode == TC_GN_RES_ID) {
        // out + x (blocks.py:164, identity residual): x = hi + lo of the tiled block input, fetched
        // with sector-coalesced loads into the staging area as fp32
        const int items = (u1 - u0) * 128 * UC;
        const int kch_res = a.res.C >> E::kShift;
        for (int idx = et; idx < items; idx += kTcEpiThreads) {
          const int m = idx % UC, r = (idx / UC) & 127, uu = idx / (UC * 128);
          const int u = u0 + uu;
          const int lo = (u * 16) / a.ct;
          const int k = lo * a.res.C + nt * a.ct + (u * 16) % a.ct;
          const int chunk = (((k & (E::kCpc - 1)) * (BF16 ? 2 : 4)) >> 4) + m;
          const size_t src = ((size_t)rt * (a.lout * kch_res) + (k >> E::kShift)) * kTcBlockBytes + tc_swz_bytes(r, chunk);
          const uint4 h = *reinterpret_cast<const uint4*>((const uint8_t*)a.res.hi + src);
          uint4 l = make_uint4(0, 0, 0, 0);
          if (a.res.lo) l = *reinterpret_cast<const uint4*>((const uint8_t*)a.res.lo + src);
          float x[8];
          tc_chunk_sum<EL>(h, l, a.res.lo != nullptr, x);
          // residual staging row: 16 floats = 4 x 16 B, chunk q at position q ^ ((r>>1)&3)
          float* dst = stg_res + (size_t)(uu * 128 + r) * 16;
          const int rs = (r >> 1) & 3;
          if (BF16) {
            *reinterpret_cast<float4*>(dst + (((2 * m) ^ rs) << 2)) = make_float4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<float4*>(dst + (((2 * m + 1) ^ rs) << 2)) = make_float4(x[4], x[5], x[6], x[7]);
          } else {
            *reinterpret_cast<float4*>(dst + ((m ^ rs) << 2)) = make_float4(x[0], x[1], x[2], x[3]);
          }
        }
        epi_barrier();
      }
      for (int u = u0 + part; u < u1; u += kParts) {
        float v[16];
        umma::tmem_ld16(t_lane + a.ph[0].d_col + u * 16, v);
        const int lo = (u * 16) / a.ct;
        const int c0 = (u * 16) % a.ct;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float y = fmaf(v[i], sc0, s_par[c0 + i]);
          if (a.mode != TC_BIAS) {
            const bool g1 = (gpt == 2 && i >= 8);
            y = (y - (g1 ? mean[1] : mean[0])) * (g1 ? rstd[1] : rstd[0]) * s_par[64 + c0 + i] + s_par[128 + c0 + i];
            y = mish_fast(y) + s_par[192 + c0 + i];
          }
          v[i] = y;
        }
        if (a.mode == TC_GN_RES_PW) {
          float r[16];
          umma::tmem_ld16(t_lane + a.ph[1].d_col + u * 16, r);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += fmaf(r[i], sc1, s_par[256 + c0 + i]);
        } else if (a.mode == TC_GN_RES_ID) {
          const float* sr = stg_res + ((size_t)(u - u0) * 128 + row_local) * 16;
          const int rs = (row_local >> 1) & 3;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const float4 h = *reinterpret_cast<const float4*>(sr + ((m ^ rs) << 2));
            v[m * 4 + 0] += h.x;
            v[m * 4 + 1] += h.y;
            v[m * 4 + 2] += h.z;
            v[m * 4 + 3] += h.w;
          }
        }
        if (a.out_hi) {
          // staging tile of this unit: [128 rows][UC x 16 B], chunk m of row r at position m ^ sw_out
          uint4 h[4], l[4];
          tc_split_store<EL>(v, a.out_lo != nullptr, h, l);
          uint4* sh = reinterpret_cast<uint4*>(stg_hi + ((size_t)(u - u0) * 128 + row_local) * (UC * 16));
          uint4* sl = reinterpret_cast<uint4*>(stg_lo + ((size_t)(u - u0) * 128 + row_local) * (UC * 16));
#pragma unroll
          for (int m = 0; m < UC; ++m) {
            sh[m ^ sw_out] = h[m];
            if (a.out_lo) sl[m ^ sw_out] = l[m];
          }
        } else {
          // plain [row][c][l] staging: this CTA's channels are one contiguous run per row
          float* sp = stg_plain + (size_t)row_local * plain_stride + (size_t)c0 * a.lout + lo;
#pragma unroll
          for (int i = 0; i < 16; ++i) sp[(size_t)i * a.lout] = v[i];
        }
      }
      epi_barrier();
      if (a.out_hi) {
        const int items = (u1 - u0) * 128 * UC;          // (unit, row, 16-byte chunk)
        for (int idx = et; idx < items; idx += kTcEpiThreads) {
          const int m = idx % UC, r = (idx / UC) & 127, uu = idx / (UC * 128);
          if (rt * kTcRows + r >= a.rows) continue;
          const int u = u0 + uu;
          const int lo = (u * 16) / a.ct;
          const int k = lo * a.cout + nt * a.ct + (u * 16) % a.ct;
          const int chunk = (((k & (E::kCpc - 1)) * (BF16 ? 2 : 4)) >> 4) + m;
          size_t dst = ((size_t)rt * (a.lout * kch_out) + (k >> E::kShift)) * kTcBlockBytes + tc_swz_bytes(r, chunk);
          if (BF16 && a.out_pm) {
            // hand-over to the position-major levels: [row block of 8][lo + 2][8 rows][cout halves]
            const int grow = rt * kTcRows + r;
            const int c = nt * a.ct + (u * 16) % a.ct;
            dst = (size_t)(grow >> 3) * pm_img_bytes(a.lout, a.cout) + pm_act_off(a.lout, lo, grow & 7, (c >> 3) + m);
          }
          const int sw = (UC == 4) ? ((r >> 1) & 3) : ((r >> 2) & 1);
          const size_t src = ((size_t)(uu * 128 + r) * UC + (m ^ sw)) * 16;
          *reinterpret_cast<uint4*>((uint8_t*)a.out_hi + dst) = *reinterpret_cast<const uint4*>(stg_hi + src);
          if (a.out_lo) *reinterpret_cast<uint4*>((uint8_t*)a.out_lo + dst) = *reinterpret_cast<const uint4*>(stg_lo + src);
        }
      } else {
        // all units of a plain-output layer fit one round (host guarantees it)
        const int ew = et >> 5;
        for (int r = ew; r < kTcRows; r += kTcEpiWarps) {
          const int grow = rt * kTcRows + r;
          if (grow >= a.rows) break;
          float* o = a.out_plain + ((size_t)grow * a.cout + (size_t)nt * a.ct) * a.lout;
          const float* sp = stg_plain + (size_t)r * plain_stride;
          for (int e = lane; e < N; e += 32) o[e] = sp[e];
        }
      }
      epi_barrier();
    }
    if (dbg && threadIdx.x == 64) dbg[6] = clock64();
    umma::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    umma::tc_fence_after();
    umma::tmem_dealloc<512>(tmem_base);
  }
  if (dbg && threadIdx.x == 0) dbg[7] = clock64();
}

// plain [rows][C][L] float32  ->  tiled hi/lo operand blocks (K order (l, c)); one thread per
// (row, l, 16-byte chunk of channels)
template <int EL>
__global__ void tc_pack_kernel(const float* __restrict__ x, int rows, int C, int L, void* __restrict__ hi,
                               void* __restrict__ lo) {
  using E = TcElem<EL>;
  constexpr bool BF16 = E::k16;
  constexpr int EPC = BF16 ? 8 : 4;   // elements per 16-byte chunk
  pdl_launch_dependents();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int cn = C / EPC;
  const size_t total = (size_t)rows * L * cn;
  if (i >= total) return;
  const int cq = (int)(i % cn);
  const int l = (int)((i / cn) % L);
  const int row = (int)(i / ((size_t)cn * L));
  const int rt = row / kTcRows, rl = row % kTcRows;
  float v[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = 0.0f;
#pragma unroll
  for (int e = 0; e < EPC; ++e) v[e] = x[((size_t)row * C + cq * EPC + e) * L + l];
  const int k = l * C + cq * EPC;
  const size_t blk = ((size_t)rt * (L * (C >> E::kShift)) + (k >> E::kShift)) * kTcBlockBytes;
  const int off = tc_swz_bytes(rl, (k & (E::kCpc - 1)) / EPC);
  uint4 h[4], r[4];
  tc_split_store<EL>(v, lo != nullptr, h, r);
  *reinterpret_cast<uint4*>((uint8_t*)hi + blk + off) = h[0];
  if (lo) *reinterpret_cast<uint4*>((uint8_t*)lo + blk + off) = r[0];
}

// tiled hi/lo -> plain [rows][C][L] (debug read-back)
template <int EL>
__global__ void tc_unpack_kernel(const void* __restrict__ hi, const void* __restrict__ lo, int rows, int C, int L,
                                 float* __restrict__ x) {
  using E = TcElem<EL>;
  constexpr bool BF16 = E::k16;
  constexpr int EPC = BF16 ? 8 : 4;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)rows * C * L;
  if (i >= total) return;
  const int l = (int)(i % L);
  const int c = (int)((i / L) % C);
  const int row = (int)(i / ((size_t)C * L));
  const int rt = row / kTcRows, rl = row % kTcRows;
  const int k = l * C + c;
  const size_t blk = ((size_t)rt * (L * (C >> E::kShift)) + (k >> E::kShift)) * kTcBlockBytes;
  const int off = tc_swz_bytes(rl, (k & (E::kCpc - 1)) / EPC);
  const int e = k % EPC;
  if (BF16) {
    const uint16_t h = *reinterpret_cast<const uint16_t*>((const uint8_t*)hi + blk + off + e * 2);
    float v = unpack16x2<EL>((uint32_t)h).x;
    if (lo) v += unpack16x2<EL>((uint32_t)*reinterpret_cast<const uint16_t*>((const uint8_t*)lo + blk + off + e * 2)).x;
    x[i] = v;
  } else {
    float v = *reinterpret_cast<const float*>((const uint8_t*)hi + blk + off + e * 4);
    if (lo) v += *reinterpret_cast<const float*>((const uint8_t*)lo + blk + off + e * 4);
    x[i] = v;
  }
}

}  // namespace edmp
