// Persistent tensor-core (tcgen05 + TMEM) convolution kernel for the TemporalUNet levels with horizon <= 13,
// second generation of conv_tc.cuh (same GEMM view, same operand layouts in HBM, 16-bit operand elements):
//
//   D[row, (l_out, c_out)] = sum_{(l_in, c_in)} X[row, (l_in, c_in)] * W      ("rows as M", windowed implicit GEMM)
//
// What changed, and the measurement behind each change (profiles/micro/r1_*.txt, tools/micro/*.cu):
//   * tcgen05.mma issue is effectively synchronous with execution (queue depth ~1, >= 40 cycles per M=128
//     instruction): every cycle the issuing thread spends on waits / index arithmetic between MMAs is a cycle the
//     tensor pipe idles.  The issue loop is therefore lean (stage counters instead of div/mod, commit inside the
//     elected region, at most two MMA runs per step for the first K chunk) and the first-chunk accumulate flags
//     are precomputed on the host.
//   * one CTA per SM stays resident and walks tiles (row tile, column tile) of the layer: no per-tile TMEM
//     allocation / barrier setup, no wave quantisation, and -- with two TMEM accumulator buffers -- the epilogue
//     of tile i runs under the mainloop of tile i+1.
//   * the operand producer is one dedicated thread that polls both streams (activations, weights) and runs ahead
//     across tile boundaries; a bulk copy has ~1.3k cycles of latency, so throughput is set by bytes in flight,
//     not by the number of issuing threads.
//   * the epilogue keeps a thread's accumulator values in registers: ONE TMEM read (then the buffer is handed back
//     to the MMA warp at once), GroupNorm statistics two-pass on registers with per-unit partials meeting in shared
//     memory among the four warps of a TMEM lane quarter, results stored straight to the tiled layout the next
//     layer consumes -- no shared-memory staging, which would compete with the MMA operand fetch for the
//     128 B/cycle of shared-memory bandwidth (the binding resource of SS-mode MMAs at N <= 128).
//   * CG = 2: cta_group::2 -- a CTA pair (two row tiles, same column tile) shares each weight tile: every CTA
//     stages only half of it and the pair's MMA (M = 256) reads A from both CTAs, halving the weight traffic
//     and the shared-memory operand bandwidth per CTA.
//
// Reference ops covered: Conv1dBlock (blocks.py:13-34), ResidualConvolutionBlock (:137-166) incl. the 1x1
// residual conv (second accumulator), stride-2 Conv1d (:211), ConvTranspose1d (:249).
#pragma once
#include "conv_tc.cuh"

namespace edmp {

constexpr int kT2EpiWarp0 = 3;                       // warp 0: operand producer, 1 (+TMEM) and 2: MMA issue
constexpr int kT2EpiWarps = 16;
constexpr int kT2EpiThreads = kT2EpiWarps * 32;
constexpr int kT2ProdWarp2 = kT2EpiWarp0 + kT2EpiWarps;        // warp 19: the weight-stream producer (EDMP_PRODUCERS != 1)
constexpr int kT2Threads = (kT2EpiWarp0 + kT2EpiWarps + 1) * 32;   // 640 (registers are granted for 20 warps anyway)
constexpr int kT2MaxAStages = 4, kT2MaxBStages = 3;
constexpr int kT2MaxUnits = 16;                      // accumulator columns per tile <= 256

struct Tc2Sched {       // what one K chunk at input position l_in contributes to
  int8_t slot_begin;    // first weight slot
  int8_t n_slots;       // consecutive output positions touched (0: position unused)
  int8_t lo_begin;      // first output position
  int8_t n_acc;         // first K chunk only: leading positions of the window that already hold a partial sum
};

struct Tc2Phase {
  TcOperand a, b;       // b.C == 0 unless the input is a skip concat (blocks.py:253)
  const void* w_hi;     // packed weights [n_tile][c_chunk][CG halves][slots * ct / CG rows][128 B] swizzled
  const void* w_lo;
  float acc_scale;      // exact inverse of the power-of-two weight scale
  int lin, slots;
  int d_col;            // accumulator column base inside a buffer
  int col_step;         // accumulator columns per output position in the MMA's D address (ct, or ct/2 for the
                        // [half][position][ct/2] layout of a cta_group::2 full-window phase)
  Tc2Sched sched[kTcMaxLin];
};

struct Tc2Args {
  Tc2Phase ph[2];
  int n_phases;
  int rows, lout, ct, cout, cg, mode, split;
  int a_stages, b_stages;
  int acc_bufs, acc_stride;       // TMEM accumulator buffers (1 or 2) and their column stride
  int half_layout;                // 1: phase-0 accumulator columns are [half][position][ct/2] (cta_group::2)
  int mma_warps;                  // 1 or 2 MMA-issuing warps (2: alternate steps, see the kernel)
  // Row-tile chaining between consecutive conv_tc2 launches (everything a tile reads besides the weights was computed
  // from the same trajectory rows): `done[rt]` counts this layer's finished (row tile, column tile) tiles; a tile's
  // activation loads wait until the producing layer's counter of that row tile reaches dep_target.  With programmatic
  // dependent launch the next layer then starts on every SM the moment that SM runs out of tiles -- no grid-wide
  // drain between layers.  dep == null: plain griddepcontrol.wait (the producer is another kernel).
  const int* dep;
  int dep_target;
  int* done;
  int b_pad;                      // 1: weight stages are [zero slot][real slots][zero slot], neighbouring stages share a zero slot
  // Column split of a GroupNorm group over a thread-block cluster (small batches: more, narrower tiles so that one wave
  // covers the machine): ct = cg / nsplit, the nsplit CTAs of a cluster hold the column tiles of ONE group of the same row
  // tile and exchange their partial (mean, M2) through distributed shared memory.  1: no cluster.
  int nsplit;
  int n_row_tiles, n_col_tiles;   // n_row_tiles counts 128-row tiles (even for CG = 2)
  int ct_log2, cg_log2, nct_log2; // ct, cg and n_col_tiles are powers of two
  const float *bias, *gamma, *beta, *temb, *bres;
  TcOperand res;                  // identity residual source (tiled)
  void *out_hi, *out_lo;          // tiled output (or position-major images when out_pm)
  int out_pm;
  long long* dbg;                 // optional [ctas][16] clock64 stamps / counters, normally null
  unsigned* range_flag;           // set when a stored IEEE-half hi part is infinite (conv_tc.cuh range_track), may be null
};

// One persistent launch = a RUN of consecutive layers.  Every role (operand producer, MMA issue,
// epilogue) walks the layers of the run in order and keeps its ring / accumulator state across layer boundaries, so a
// CTA's last epilogue of layer i runs under its first mainloop of layer i+1, the weight and activation stages never
// drain between layers, and barrier set-up / TMEM allocation happen once per run.  Layers of a run share ONE
// shared-memory carve-up (the most demanding layer decides), the cta_group, the zero-slot layout and the number of
// issuing warps; single-buffered layers (1x1-residual phases: 512 accumulator columns) take both TMEM buffers under
// a both-buffers handshake.  Row tiles are handed from layer to layer through the global `done` / `dep` counters as
// between launches.  Work assignment: the tiles of all layers of a run form one round-robin sequence over the
// walkers (rot[j] = tiles of layers 0..j-1 mod walkers), so that the odd tiles of non-divisible layers spread evenly.
constexpr int kT2MaxRun = 14;
// (NR = capacity: single-layer launches use Tc2RunT<1> -- a 0.5 KB kernel parameter instead of 7 KB)
template <int NR>
struct Tc2RunT {
  int n;                                   // layers
  int a_stages, b_stages;                  // operand stage rings of the run
  int b_stage_stride, b_lo_off, b_real_off, b_total_bytes;   // weight stage geometry (bytes)
  int b_pad, slot_bytes;                   // zero tap slots ([Z][real][Z]... layout) of slot_bytes each
  int mma_warps, split;
  int lean;                                // 1: lean issue path (one thread, tabulated schedule; mma_warps == 1)
  int producers;                           // operand producer threads: 1 (warp 0: both streams), 2 (+ warp 19: the weight stream), 3 (lean
                                           // only: + warp 2: the lo parts of the activation stream)
  int rot[NR];
  Tc2Args l[NR];
};
typedef Tc2RunT<kT2MaxRun> Tc2Run;

namespace t2 {
// elect.sync: true in exactly one lane of the converged warp; `leader` = that lane's id in every lane
__device__ __forceinline__ bool elect_leader(uint32_t& leader) {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred px;\n\t"
      "elect.sync %0|px, 0xffffffff;\n\t"
      "selp.u32 %1, 1, 0, px;\n\t}"
      : "=r"(leader), "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mma_f16_cg2(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in both CTAs of the pair
__device__ __forceinline__ void commit_cg2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   umma::smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// arrive on a barrier of CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(umma::smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// store one float into the shared memory of CTA `rank` of the cluster (same offset as `p` has here)
__device__ __forceinline__ void st_remote_f32(float* p, uint32_t rank, float v) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(umma::smem_u32(p)), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}
// wait on a local barrier that CTAs of the cluster arrive on (acquire at cluster scope: their remote stores are visible)
__device__ __forceinline__ void wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(umma::smem_u32(bar)), "r"(parity), "r"(umma::kMbarSuspendNs)
        : "memory");
    if (!ok && ++spins > (1u << 22)) __trap();
  }
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread, WITHOUT waiting: several loads can be in
// flight before one tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait without the diagnostic printf of umma::mbar_wait in the hot loops (still bounded: traps instead of hanging)
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!umma::mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}
// Mish = x * n / (n + 2), n = e^x (e^x + 2), with the single-instruction ex2 / rcp approximations (relative error
// ~1e-7 each, below the split-MMA accumulation error).  Clamping x at 20 makes n / (n + 2) round to exactly 1.
__device__ __forceinline__ float mish(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(x, 20.0f) * 1.4426950408889634f));
  const float n = e * (e + 2.0f);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n + 2.0f));
  return x * (n * r);
}
// acquire-load of a global progress counter
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// non-blocking phase test
__device__ __forceinline__ bool test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(umma::smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bar_quarter(int quarter) {   // the four epilogue warps sharing a TMEM lane quarter
  asm volatile("bar.sync %0, 128;" ::"r"(quarter + 1) : "memory");
}
// token passing between the two MMA-issuing warps (named barriers 6 and 7, 64 threads: one warp arrives, the other syncs)
__device__ __forceinline__ void token_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(6 + id) : "memory"); }
__device__ __forceinline__ void token_pass(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(6 + id) : "memory"); }
__device__ __forceinline__ void bar_epilogue() { asm volatile("bar.sync 5, %0;" ::"n"(kT2EpiThreads) : "memory"); }
}  // namespace t2

template <int EL, int CG, int NR>
__global__ void __launch_bounds__(kT2Threads, 1) conv_tc2_kernel(const __grid_constant__ Tc2RunT<NR> r) {
  static_assert(EL != TC_EL_TF32, "conv_tc2 uses 16-bit operand elements");
  using E = TcElem<EL>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // (offset arithmetic on the array keeps the shared address space known to the compiler: LDS/STS, not generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t a_full[kT2MaxAStages], a_empty[kT2MaxAStages], b_full[kT2MaxBStages], b_empty[kT2MaxBStages];
  __shared__ uint64_t pa_full[kT2MaxAStages], pb_full[kT2MaxBStages];   // CG = 2, leader: "the peer's stage is full"
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint64_t xg_full[2];   // nsplit > 1: "the partial group statistics of every cluster peer have arrived" (by tile parity)
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_par2[2][5 * 128];        // bias | gamma | beta | temb | bres of a column tile (by tile parity)
  // lean issue path: a layer's MMA schedule, tabulated once per layer as [phase][first chunk | later chunks][input position]
  // (three packed words per step: the static shared memory must stay within its 7 KB -- the stage rings fill the rest):
  //   w = n0 | n1 << 8 | acc0 << 16 | n_runs << 24   window widths (slots) of the (at most two) MMA groups of the step, the
  //                                                  accumulate flag of group 0's first 32-byte K step, 0 runs = position unused
  //   d = d0 | d1 << 16                              accumulator column offsets;   b = b0 | b1 << 16: weight offsets in 128-byte rows
  __shared__ uint32_t s_st_w[2][2][kTcMaxLin], s_st_d[2][2][kTcMaxLin], s_st_b[2][2][kTcMaxLin];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? t2::cluster_ctarank() : 0u;
  const uint32_t nsplit = CG == 1 ? (uint32_t)r.l[0].nsplit : 1u;          // column split of a GroupNorm group over a cluster (single-layer launches)
  const uint32_t crank = nsplit > 1 ? t2::cluster_ctarank() : 0u;
  const int nparts = r.split ? 2 : 1;
  const int a_stage_bytes = kTcBlockBytes * nparts;
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + r.a_stages * a_stage_bytes;
  // Weight stages.  Plain: [stage][hi | lo][slots x rows].  Padded (full-window CTA-pair layers whose first
  // and last tap slot are all zero): per part [Z][real slots of stage 0][Z][real slots of stage 1][Z]... -- the zero
  // slots are written once, never copied, and shared by neighbouring stages; a stage's window view starts at its
  // leading zero slot.  The geometry is the run's (Tc2Run), not a layer's.
  const int b_stage_stride = r.b_stage_stride, b_lo_off = r.b_lo_off, b_real_off = r.b_real_off;
  float* s_part = reinterpret_cast<float*>(b_smem + r.b_total_bytes);   // GroupNorm pieces [piece][mean | M2][128 rows]
  const int unit0 = blockIdx.x / CG, n_walkers = gridDim.x / CG;
  // this walker's first tile of layer li, and that layer's tile count
  // (NR == 1: the layer index is the constant 0, so that every a.<field> below is a fixed constant-bank operand as in a
  // single-layer kernel; NR > 1 pays an indexed constant load per access)
  auto LX = [](int li) { return NR == 1 ? 0 : li; };
  auto first_tile = [&](int li) { if (NR == 1) return unit0; const int t = unit0 - r.rot[li]; return t < 0 ? t + n_walkers : t; };
  auto layer_tiles = [&](int li) { return (r.l[LX(li)].n_row_tiles / CG) * r.l[LX(li)].n_col_tiles; };

  long long* dbg = r.l[0].dbg ? r.l[0].dbg + (size_t)blockIdx.x * 16 : nullptr;
  const int ablate = dbg ? g_edmp_ablate : 0;   // trace-mode ablation (conv_pm.cuh)
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    for (int i = 0; i < r.a_stages; ++i) { umma::mbar_init(a_full + i, 1); umma::mbar_init(a_empty + i, 1); umma::mbar_init(pa_full + i, 1); }
    // (with two issuing warps each commits the weight stage / the accumulator after its own last step)
    for (int i = 0; i < r.b_stages; ++i) { umma::mbar_init(b_full + i, 1); umma::mbar_init(b_empty + i, r.mma_warps); umma::mbar_init(pb_full + i, 1); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(acc_full + i, r.mma_warps); umma::mbar_init(acc_empty + i, kT2EpiWarps * CG); }
    for (int i = 0; i < 2; ++i) umma::mbar_init(xg_full + i, nsplit > 1 ? (nsplit - 1) * 4 : 1);   // one arrive per peer and lane quarter
    umma::fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma::smem_u32(&tmem_slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      umma::tmem_alloc<512>(&tmem_slot);
    }
  }
  if (r.b_pad) {
    // zero slots: one before every stage's real slots and one after the last stage, in both parts
    const int nz = r.b_stages + 1, words = r.slot_bytes >> 4;
    for (int i = threadIdx.x; i < nparts * nz * words; i += blockDim.x) {
      const int w = i % words, z = (i / words) % nz, part_i = i / (words * nz);
      reinterpret_cast<uint4*>(b_smem + part_i * b_lo_off + z * b_stage_stride)[w] = make_uint4(0, 0, 0, 0);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA's operand reads
  }
  umma::tc_fence_before();
  __syncthreads();
  if (CG == 2 || nsplit > 1) t2::cluster_sync_all();
  umma::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // everything above overlapped the previous kernel; activations are read below.  Chained layers synchronise per row
  // tile instead (the producer thread, before a tile's first activation load)
  if (r.l[0].dep == nullptr) pdl_wait();
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();

  if (r.producers > 1 && (warp == 0 || warp == kT2ProdWarp2 || (warp == 2 && r.producers == 3))) {
    // ===== operand producers, one THREAD per stream (default EDMP_PRODUCERS=2; 3: activation lo parts on warp 2 as well).  A thread sustains about one 1-D bulk copy per
    // 1000 cycles however many stages it keeps in flight, and the rate adds up over issuing threads (tools/micro/l2_feed.cu,
    // profiles/micro/r2_l2_feed.txt): the six copies of a pair layer's K chunk (2 positions x hi / lo, weights hi / lo) take
    // one thread ~6k cycles against 3072 of MMA time.  Warp 0: activation blocks (the hi parts only when warp 2 carries the
    // lo parts: lean issuer), warp 19: weight tiles.  Every thread walks its stream's (layer, tile, phase, chunk[, position])
    // sequence on its own with blocking waits; the hi thread arms the "stage full" barrier with the byte count of both
    // parts (a lo copy that lands first only drives the transaction count negative until then). =====
    if (lane == 0) {
      if (warp == kT2ProdWarp2) {
        uint32_t bs = 0, bph = 0;
        int lb = 0, tb = first_tile(0), ntb = layer_tiles(0), pb = 0, ccb = 0;
        while (lb < r.n && tb >= ntb) { if (++lb < r.n) { tb = first_tile(lb); ntb = layer_tiles(lb); } }   // layers without a tile for this walker
        while (lb < r.n) {
          t2::wait(b_empty + bs, bph ^ 1);
          const Tc2Args& a = r.l[LX(lb)];
          const Tc2Phase& ph = a.ph[pb];
          const int kc = (ph.a.C + ph.b.C) >> E::kShift;
          const int nt = tb & (a.n_col_tiles - 1);
          const uint32_t wtile_bytes = (uint32_t)(ph.slots * (a.ct / CG) * 128);
          const size_t woff = (((size_t)nt * kc + ccb) * CG + rank) * wtile_bytes;
          uint8_t* dst = b_smem + bs * b_stage_stride + b_real_off;
          if (ablate & 4) {
            umma::mbar_arrive(b_full + bs);
          } else {
            umma::mbar_arrive_expect_tx(b_full + bs, wtile_bytes * (uint32_t)nparts);
            umma::bulk_g2s(dst, (const uint8_t*)ph.w_hi + woff, wtile_bytes, b_full + bs);
            if (r.split) umma::bulk_g2s(dst + b_lo_off, (const uint8_t*)ph.w_lo + woff, wtile_bytes, b_full + bs);
          }
          if (++bs == (uint32_t)r.b_stages) { bs = 0; bph ^= 1; }
          if (++ccb == kc) {
            ccb = 0;
            if (++pb == a.n_phases) {
              pb = 0;
              tb += n_walkers;
              while (lb < r.n && tb >= ntb) { if (++lb < r.n) { tb = first_tile(lb); ntb = layer_tiles(lb); } }
            }
          }
        }
      } else {
        const bool do_hi = warp == 0, do_lo = r.split && (warp == 2 || r.producers == 2);
        uint32_t as = 0, aph = 0;
        int la = 0, ta = first_tile(0), nta = layer_tiles(0), pa = 0, cca = 0, lia = 0;
        while (la < r.n && ta >= nta) { if (++la < r.n) { ta = first_tile(la); nta = layer_tiles(la); } }
        if (la < r.n) while (r.l[LX(la)].ph[0].sched[lia].n_slots == 0) ++lia;                                    // leading unused positions
        if (!do_hi && !do_lo) la = r.n;
        int ta_ready = -1;                            // (layer, tile) whose row-tile dependency has been observed
        while (la < r.n) {
          t2::wait(a_empty + as, aph ^ 1);
          const Tc2Args& a = r.l[LX(la)];
          if (a.dep != nullptr && ((la << 20) | ta) != ta_ready) {
            // first activation load of this tile: the producing layer must have finished this row tile
            const int rt_dep = (ta >> a.nct_log2) * CG + (int)rank;
            // (a pair's padding row tile past the batch has no producer: its operand blocks are the zero-filled tail)
            if (rt_dep * kTcRows < a.rows) {
              const long long dep_t0 = clock64();
              while (t2::ld_acquire(a.dep + rt_dep) < a.dep_target) {
                if (clock64() - dep_t0 > (1ll << 32)) {   // ~2 s: a protocol bug traps instead of hanging the GPU
                  printf("edmp: conv_tc2 row-tile dependency timed out (block %d, layer %d of the run, row tile %d: %d of %d)\n",
                         blockIdx.x, la, rt_dep, t2::ld_acquire(a.dep + rt_dep), a.dep_target);
                  __trap();
                }
              }
            }
            asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy acquire -> async-proxy (bulk copy) reads
            ta_ready = (la << 20) | ta;
          }
          {
            const Tc2Phase& ph = a.ph[pa];
            const int ka = ph.a.C >> E::kShift, kb = ph.b.C >> E::kShift;
            const bool first = cca < ka;
            const TcOperand& op = first ? ph.a : ph.b;
            const int c2 = first ? cca : cca - ka;
            const int kop = first ? ka : kb;
            const int rt = (ta >> a.nct_log2) * CG + (int)rank;
            const size_t blk = ((size_t)rt * (ph.lin * kop) + (size_t)lia * kop + c2) * kTcBlockBytes;
            uint8_t* dst = a_smem + as * a_stage_bytes;
            if (ablate & 4) {
              if (do_hi) umma::mbar_arrive(a_full + as);
            } else {
              if (do_hi) {
                umma::mbar_arrive_expect_tx(a_full + as, (uint32_t)a_stage_bytes);
                umma::bulk_g2s(dst, (const uint8_t*)op.hi + blk, kTcBlockBytes, a_full + as);
              }
              if (do_lo) umma::bulk_g2s(dst + kTcBlockBytes, (const uint8_t*)op.lo + blk, kTcBlockBytes, a_full + as);
            }
            if (++as == (uint32_t)r.a_stages) { as = 0; aph ^= 1; }
          }
          // advance (layer, tile, phase, chunk, position), skipping unused positions and layers without a tile
          for (;;) {
            const Tc2Args& c = r.l[LX(la)];
            const Tc2Phase& cp = c.ph[pa];
            if (++lia == cp.lin) {
              lia = 0;
              if (++cca == ((cp.a.C + cp.b.C) >> E::kShift)) {
                cca = 0;
                if (++pa == c.n_phases) {
                  pa = 0;
                  ta += n_walkers;
                  while (la < r.n && ta >= nta) { if (++la < r.n) { ta = first_tile(la); nta = layer_tiles(la); } }
                }
              }
            }
            if (la >= r.n || r.l[LX(la)].ph[pa].sched[lia].n_slots != 0) break;
          }
        }
      }
    }
  } else if (warp == kT2ProdWarp2) {
    // (one producer thread: this warp has no role)
  } else if (warp == 0) {
    // (Tried in round 2: FOUR issuing lanes -- activations hi / lo, weights hi / lo.  In isolation a thread sustains only
    // ~19 B/clk of bulk copies however many stages it keeps in flight and the rate adds up over issuing lanes
    // (tools/micro/l2_feed.cu, profiles/micro/r2_l2_feed.txt: 19 / 35 / 62 B/clk/SM for 1 / 2 / 4 lanes) -- but inside this
    // kernel the four-lane producer changed nothing (pair layers 1.69 ms either way, same stage waits): with <= 5 stages the
    // queue of one thread is never the limit here.  profiles/r2_s2_experiments.md.)
    // ===== operand producer: ONE thread polls the "stage empty" barriers of both streams (activations, weights)
    // without blocking and issues bulk copies of ready-made operand blocks; both streams run ahead across tile AND
    // layer boundaries as far as the stages allow =====
    if (lane == 0) {
      uint32_t as = 0, aph = 0, bs = 0, bph = 0;
      // stream positions: (layer, tile, phase, chunk, input position) of the activations, (layer, tile, phase, chunk) of the weights
      int la = 0, ta = first_tile(0), nta = layer_tiles(0), pa = 0, cca = 0, lia = 0;
      int lb = 0, tb = first_tile(0), ntb = layer_tiles(0), pb = 0, ccb = 0;
      while (la < r.n && ta >= nta) { if (++la < r.n) { ta = first_tile(la); nta = layer_tiles(la); } }   // layers without a tile for this walker
      while (lb < r.n && tb >= ntb) { if (++lb < r.n) { tb = first_tile(lb); ntb = layer_tiles(lb); } }
      if (la < r.n) while (r.l[LX(la)].ph[0].sched[lia].n_slots == 0) ++lia;                                    // leading unused positions
      int ta_ready = -1;                            // (layer, tile) whose row-tile dependency has been observed
      long long dep_t0 = 0;
      while (la < r.n || lb < r.n) {
        bool a_go = la < r.n && t2::test(a_empty + as, aph ^ 1);
        if (a_go && r.l[LX(la)].dep != nullptr && ((la << 20) | ta) != ta_ready) {
          // first activation load of this tile: the producing layer must have finished this row tile (the weight
          // stream below keeps being served meanwhile)
          const Tc2Args& a = r.l[LX(la)];
          const int rt_dep = (ta >> a.nct_log2) * CG + (int)rank;
          // (a pair's padding row tile past the batch has no producer: its operand blocks are the zero-filled tail)
          if (rt_dep * kTcRows < a.rows && t2::ld_acquire(a.dep + rt_dep) < a.dep_target) {
            a_go = false;
            if (dep_t0 == 0) dep_t0 = clock64();
            else if (clock64() - dep_t0 > (1ll << 32)) {   // ~2 s: a protocol bug traps instead of hanging the GPU
              printf("edmp: conv_tc2 row-tile dependency timed out (block %d, layer %d of the run, row tile %d: %d of %d)\n",
                     blockIdx.x, la, rt_dep, t2::ld_acquire(a.dep + rt_dep), a.dep_target);
              __trap();
            }
          } else {
            dep_t0 = 0;
            asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy acquire -> async-proxy (bulk copy) reads
            ta_ready = (la << 20) | ta;
          }
        }
        if (a_go) {
          const Tc2Args& a = r.l[LX(la)];
          const Tc2Phase& ph = a.ph[pa];
          const int ka = ph.a.C >> E::kShift, kb = ph.b.C >> E::kShift;
          const bool first = cca < ka;
          const TcOperand& op = first ? ph.a : ph.b;
          const int c2 = first ? cca : cca - ka;
          const int kop = first ? ka : kb;
          const int rt = (ta >> a.nct_log2) * CG + (int)rank;
          const size_t blk = ((size_t)rt * (ph.lin * kop) + (size_t)lia * kop + c2) * kTcBlockBytes;
          uint8_t* dst = a_smem + as * a_stage_bytes;
          if (ablate & 4) {   // (trace-mode ablation: no copies, the stage is "full" at once)
            umma::mbar_arrive(a_full + as);
          } else {
            umma::mbar_arrive_expect_tx(a_full + as, (uint32_t)a_stage_bytes);
            umma::bulk_g2s(dst, (const uint8_t*)op.hi + blk, kTcBlockBytes, a_full + as);
            if (r.split) umma::bulk_g2s(dst + kTcBlockBytes, (const uint8_t*)op.lo + blk, kTcBlockBytes, a_full + as);
          }
          if (++as == (uint32_t)r.a_stages) { as = 0; aph ^= 1; }
          // advance (layer, tile, phase, chunk, position), skipping unused positions and layers without a tile
          for (;;) {
            const Tc2Args& c = r.l[LX(la)];
            const Tc2Phase& cp = c.ph[pa];
            if (++lia == cp.lin) {
              lia = 0;
              if (++cca == ((cp.a.C + cp.b.C) >> E::kShift)) {
                cca = 0;
                if (++pa == c.n_phases) {
                  pa = 0;
                  ta += n_walkers;
                  while (la < r.n && ta >= nta) { if (++la < r.n) { ta = first_tile(la); nta = layer_tiles(la); } }
                }
              }
            }
            if (la >= r.n || r.l[LX(la)].ph[pa].sched[lia].n_slots != 0) break;
          }
        }
        if (lb < r.n && t2::test(b_empty + bs, bph ^ 1)) {
          const Tc2Args& a = r.l[LX(lb)];
          const Tc2Phase& ph = a.ph[pb];
          const int kc = (ph.a.C + ph.b.C) >> E::kShift;
          const int nt = tb & (a.n_col_tiles - 1);
          const uint32_t wtile_bytes = (uint32_t)(ph.slots * (a.ct / CG) * 128);
          const size_t woff = (((size_t)nt * kc + ccb) * CG + rank) * wtile_bytes;
          uint8_t* dst = b_smem + bs * b_stage_stride + b_real_off;
          if (ablate & 4) {
            umma::mbar_arrive(b_full + bs);
          } else {
            umma::mbar_arrive_expect_tx(b_full + bs, wtile_bytes * (uint32_t)nparts);
            umma::bulk_g2s(dst, (const uint8_t*)ph.w_hi + woff, wtile_bytes, b_full + bs);
            if (r.split) umma::bulk_g2s(dst + b_lo_off, (const uint8_t*)ph.w_lo + woff, wtile_bytes, b_full + bs);
          }
          if (++bs == (uint32_t)r.b_stages) { bs = 0; bph ^= 1; }
          if (++ccb == kc) {
            ccb = 0;
            if (++pb == a.n_phases) {
              pb = 0;
              tb += n_walkers;
              while (lb < r.n && tb >= ntb) { if (++lb < r.n) { tb = first_tile(lb); ntb = layer_tiles(lb); } }
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    uint32_t as = 0, aph = 0, bs = 0, bph = 0;
    const int mw = warp - 1, nw = r.mma_warps;   // issuing warp index, number of issuing warps
    if (CG == 2 && rank == 1) {
      // ===== peer of a CTA pair: relay "my stage is full" to the leader, which issues the MMAs for both =====
      if (mw == 0)
      for (int li = 0; li < r.n; ++li) {
        const Tc2Args& a = r.l[LX(li)];
        const int n_tiles = layer_tiles(li);
        for (int t = first_tile(li); t < n_tiles; t += n_walkers) {
          for (int p = 0; p < a.n_phases; ++p) {
            const Tc2Phase& ph = a.ph[p];
            const int kc = (ph.a.C + ph.b.C) >> E::kShift;
            for (int cc = 0; cc < kc; ++cc) {
              t2::wait(b_full + bs, bph);
              if (lane == 0) t2::mbar_arrive_remote(pb_full + bs, 0);
              __syncwarp();
              for (int li2 = 0; li2 < ph.lin; ++li2) {
                if (ph.sched[li2].n_slots == 0) continue;
                t2::wait(a_full + as, aph);
                if (lane == 0) t2::mbar_arrive_remote(pa_full + as, 0);
                __syncwarp();
                if (++as == (uint32_t)r.a_stages) { as = 0; aph ^= 1; }
              }
              if (++bs == (uint32_t)r.b_stages) { bs = 0; bph ^= 1; }
            }
          }
        }
      }
    } else if (r.lean) {
      // ===== lean MMA issuer (default; EDMP_MMA_LEAN=0 selects the first form below): ONE thread of warp 1 runs the whole issue walk of a layer.
      // Ablation traces (EDMP_ABLATE=2, profiles/r2_s2_ablation_*.txt) showed the warp-uniform walk below costing 500-700
      // cycles per step (one activation block = 12 MMAs) with NO MMA issued -- ~250 dynamic instructions of schedule decoding,
      // descriptor arithmetic, elect / reconvergence and token passing per step.  Here the per-step schedule (window widths,
      // accumulator / weight offsets, accumulate flags of the first K chunk) is tabulated in shared memory once per layer by the
      // whole warp, and the elected thread only waits for stages, reads a table entry and issues.  Measured: one lean thread
      // equals the two alternating warps (1020 rows: 860 against 960 cycles per step in the K loop, the forward within 1-2 %
      // either way): what remains per step is the shared-memory operand fetch of the MMAs themselves (N = 80: 52 cycles each);
      // together with the second producer thread it is 1 % (8190 rows) to 3 % (1020 rows) ahead. =====
      if (mw == 0) {
      const uint32_t a0 = umma::smem_u32(a_smem), b0 = umma::smem_u32(b_smem);
      const uint64_t desc0 = umma::make_desc_sw128(0);     // weight tiles: SWIZZLE_128B rows
      const uint64_t desc_a0 = tc_act_desc0();             // activation blocks: chunk-major, no swizzle (conv_tc.cuh)
      uint32_t as = 0, aph = 0, bs = 0, bph = 0, buf = 0, ne0 = 0, ne1 = 0;
      for (int li = 0; li < r.n; ++li) {
        const Tc2Args& a = r.l[LX(li)];
        const int n_tiles = layer_tiles(li);
        const int ctl = a.ct / CG;                   // weight rows per slot staged by this CTA
        // ---- this layer's step table (all lanes) ----
        __syncwarp();
        for (int e = lane; e < a.n_phases * 2 * kTcMaxLin; e += 32) {
          const int li2 = e % kTcMaxLin, v = (e / kTcMaxLin) & 1, p = e / (2 * kTcMaxLin);
          const Tc2Phase& ph = a.ph[p];
          uint32_t w = 0, dd = 0, bb = 0;
          if (li2 < ph.lin && ph.sched[li2].n_slots > 0) {
            const Tc2Sched sc = ph.sched[li2];
            // first K chunk: the leading n_acc positions of the window already hold a partial sum, the rest are written
            // for the first time -> two groups with their own accumulate flag; afterwards one group over the window
            const int n_first = v == 0 ? sc.n_acc : sc.n_slots;
            int nr = 0;
            for (int run = 0; run < 2; ++run) {
              const int sl0 = run == 0 ? 0 : n_first;
              const int n = run == 0 ? n_first : sc.n_slots - n_first;
              if (n <= 0) continue;
              const uint32_t d = (uint32_t)(ph.d_col + (sc.lo_begin + sl0) * ph.col_step);
              const uint32_t bo = (uint32_t)((sc.slot_begin + sl0) * ctl);   // 128-byte weight rows
              if (nr == 0) { w |= (uint32_t)n | ((run == 0 ? 1u : 0u) << 16); dd |= d; bb |= bo; }
              else { w |= (uint32_t)n << 8; dd |= d << 16; bb |= bo << 16; }
              ++nr;
            }
            w |= (uint32_t)nr << 24;
          }
          s_st_w[p][v][li2] = w; s_st_d[p][v][li2] = dd; s_st_b[p][v][li2] = bb;
        }
        __syncwarp();
        uint32_t leader;
        if (t2::elect_leader(leader)) {
          long long w_acc = 0, w_a = 0, w_b = 0;
          const long long t_begin = dbg ? clock64() : 0;
          if (dbg && li == 0) dbg[2] = dbg[3] = dbg[4] = dbg[5] = 0;
          const uint32_t idesc_n0 = umma::make_idesc(E::kFmt, kTcRows * CG, 0), ct8 = (uint32_t)a.ct >> 3;
          auto issue_group = [&](uint32_t idesc, uint32_t d, uint32_t b_off, uint32_t accf, uint64_t da_hi, uint64_t da_lo) {
            const uint64_t db_hi = desc0 | (uint64_t)((b_off & 0x3FFFF) >> 4);
            const uint64_t db_lo = desc0 | (uint64_t)(((b_off + (uint32_t)b_lo_off) & 0x3FFFF) >> 4);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {   // 32-byte K steps inside the 128-byte swizzle atom
              const uint32_t acc = accf | (uint32_t)(ks > 0);
              if (CG == 2) {
                if (r.split) {
                  t2::mma_f16_cg2(d, da_lo + kTcActKStep * ks, db_hi + 2 * ks, idesc, acc);
                  t2::mma_f16_cg2(d, da_hi + kTcActKStep * ks, db_lo + 2 * ks, idesc, 1u);
                  t2::mma_f16_cg2(d, da_hi + kTcActKStep * ks, db_hi + 2 * ks, idesc, 1u);
                } else {
                  t2::mma_f16_cg2(d, da_hi + kTcActKStep * ks, db_hi + 2 * ks, idesc, acc);
                }
              } else {
                if (r.split) {
                  umma::mma_bf16(d, da_lo + kTcActKStep * ks, db_hi + 2 * ks, idesc, acc);
                  umma::mma_bf16(d, da_hi + kTcActKStep * ks, db_lo + 2 * ks, idesc, 1u);
                  umma::mma_bf16(d, da_hi + kTcActKStep * ks, db_hi + 2 * ks, idesc, 1u);
                } else {
                  umma::mma_bf16(d, da_hi + kTcActKStep * ks, db_hi + 2 * ks, idesc, acc);
                }
              }
            }
          };
          for (int t = first_tile(li); t < n_tiles; t += n_walkers) {
            uint32_t ab;                               // accumulator buffer of this tile
            {
              const long long tw = dbg ? clock64() : 0;
              if (a.acc_bufs == 2) {
                ab = buf;
                t2::wait(acc_empty + ab, ((ab ? ne1 : ne0) & 1u) ^ 1u);
                if (ab) ++ne1; else ++ne0;
                buf ^= 1;
              } else {
                t2::wait(acc_empty + 0, (ne0 & 1u) ^ 1u);
                t2::wait(acc_empty + 1, (ne1 & 1u) ^ 1u);
                ++ne0; ++ne1;
                ab = 0; buf = 0;
              }
              if (dbg) w_acc += clock64() - tw;
            }
            umma::tc_fence_after();
            const uint32_t acc0 = tmem_base + ab * (uint32_t)a.acc_stride;
            for (int p = 0; p < a.n_phases; ++p) {
              const Tc2Phase& ph = a.ph[p];
              const int kc = (ph.a.C + ph.b.C) >> E::kShift;
              const int ph_lin = ph.lin;
              for (int cc = 0; cc < kc; ++cc) {
                {
                  const long long tw = dbg ? clock64() : 0;
                  t2::wait(b_full + bs, bph);
                  if (CG == 2) t2::wait(pb_full + bs, bph);
                  if (dbg) w_b += clock64() - tw;
                }
                const uint32_t b_base = b0 + bs * (uint32_t)b_stage_stride;
                const int v = cc == 0 ? 0 : 1;
                for (int li2 = 0; li2 < ph_lin; ++li2) {
                  const uint32_t w = s_st_w[p][v][li2], dd = s_st_d[p][v][li2], bb = s_st_b[p][v][li2];
                  if ((w >> 24) == 0) continue;        // position unused by this layer
                  {
                    const long long tw = dbg ? clock64() : 0;
                    t2::wait(a_full + as, aph);
                    if (CG == 2) t2::wait(pa_full + as, aph);
                    if (dbg) w_a += clock64() - tw;
                  }
                  if (!(ablate & 2)) {
                    const uint32_t a_base = a0 + as * (uint32_t)a_stage_bytes;
                    const uint64_t da_hi = desc_a0 | (uint64_t)((a_base & 0x3FFFF) >> 4);
                    const uint64_t da_lo = desc_a0 | (uint64_t)(((a_base + kTcBlockBytes) & 0x3FFFF) >> 4);
                    // instruction descriptor: N = window x ct (n_dim field = N >> 3 at bit 17)
                    issue_group(idesc_n0 + (((w & 0xFFu) * ct8) << 17), acc0 + (dd & 0xFFFFu), b_base + ((bb & 0xFFFFu) << 7),
                                (w >> 16) & 1u, da_hi, da_lo);
                    if ((w >> 24) == 2)
                      issue_group(idesc_n0 + ((((w >> 8) & 0xFFu) * ct8) << 17), acc0 + (dd >> 16), b_base + ((bb >> 16) << 7), 0u, da_hi, da_lo);
                  }
                  // frees the A stage (in both CTAs of a pair) once these MMAs have read it
                  if (CG == 2) t2::commit_cg2(a_empty + as); else umma::mma_commit(a_empty + as);
                  if (++as == (uint32_t)r.a_stages) { as = 0; aph ^= 1; }
                }
                // ... and the weight stage after the chunk / the accumulator after the tile
                if (CG == 2) t2::commit_cg2(b_empty + bs); else umma::mma_commit(b_empty + bs);
                if (p == a.n_phases - 1 && cc == kc - 1) { if (CG == 2) t2::commit_cg2(acc_full + ab); else umma::mma_commit(acc_full + ab); }
                if (++bs == (uint32_t)r.b_stages) { bs = 0; bph ^= 1; }
              }
            }
          }
          if (dbg) { dbg[2] += w_acc; dbg[3] += w_b; dbg[4] += w_a; dbg[5] += clock64() - t_begin; }
        }
        __syncwarp();
        // the ring / accumulator state lives in the elected thread: hand it to every lane for the next layer's election
        as = __shfl_sync(0xffffffffu, as, leader); aph = __shfl_sync(0xffffffffu, aph, leader);
        bs = __shfl_sync(0xffffffffu, bs, leader); bph = __shfl_sync(0xffffffffu, bph, leader);
        buf = __shfl_sync(0xffffffffu, buf, leader); ne0 = __shfl_sync(0xffffffffu, ne0, leader); ne1 = __shfl_sync(0xffffffffu, ne1, leader);
      }
      }
    } else if (mw < nw) {
      // ===== MMA issuer, first form (EDMP_MMA_LEAN=0): warp-uniform walk, one elected lane issues.  With two issuing warps the steps (one
      // activation block = 12 MMAs) alternate between them: while one warp sits in its (blocking) MMA issue the
      // other already waits for the next stage and builds its descriptors, so that per-step overhead leaves the
      // tensor pipe's critical path (profiles/micro/r1_mma_pipe.txt).  A token (named barrier + tcgen05 fences)
      // keeps the issue order identical to the single-warp order. =====
      const uint32_t a0 = umma::smem_u32(a_smem), b0 = umma::smem_u32(b_smem);
      const uint64_t desc0 = umma::make_desc_sw128(0);     // weight tiles: SWIZZLE_128B rows
      const uint64_t desc_a0 = tc_act_desc0();             // activation blocks: chunk-major, no swizzle (conv_tc.cuh)
      // accumulator buffers: `buf` = next buffer of a double-buffered tile; ne[b] = how often acc_empty[b] has been waited
      // for.  A single-buffered tile (512 columns) waits for BOTH buffers and leaves buf = 0 (the epilogue mirrors this).
      uint32_t buf = 0, ne0 = 0, ne1 = 0;
      uint32_t step = 0;
      long long w_acc = 0, w_a = 0, w_b = 0, t_begin = dbg ? clock64() : 0;
      if (nw == 2 && mw == 1) t2::token_pass(0);   // the first token
      for (int li = 0; li < r.n; ++li) {
      const Tc2Args& a = r.l[LX(li)];
      const int n_tiles = layer_tiles(li);
      const int ctl = a.ct / CG;                   // weight rows per slot staged by this CTA
      for (int t = first_tile(li); t < n_tiles; t += n_walkers) {
        uint32_t ab;                               // accumulator buffer of this tile
        {
          const long long tw = dbg ? clock64() : 0;
          if (a.acc_bufs == 2) {
            ab = buf;
            t2::wait(acc_empty + ab, ((ab ? ne1 : ne0) & 1u) ^ 1u);
            if (ab) ++ne1; else ++ne0;
            buf ^= 1;
          } else {
            t2::wait(acc_empty + 0, (ne0 & 1u) ^ 1u);
            t2::wait(acc_empty + 1, (ne1 & 1u) ^ 1u);
            ++ne0; ++ne1;
            ab = 0; buf = 0;
          }
          if (dbg) w_acc += clock64() - tw;
        }
        umma::tc_fence_after();
        const uint32_t acc0 = tmem_base + ab * (uint32_t)a.acc_stride;
        for (int p = 0; p < a.n_phases; ++p) {
          const Tc2Phase& ph = a.ph[p];
          const int kc = (ph.a.C + ph.b.C) >> E::kShift;
          const int ph_lin = ph.lin, ph_dcol = ph.d_col, ph_cstep = ph.col_step, l_ct = a.ct;   // (registers: the issuing warps have plenty)
          for (int cc = 0; cc < kc; ++cc) {
            {
              const long long tw = dbg ? clock64() : 0;
              t2::wait(b_full + bs, bph);
              if (CG == 2) t2::wait(pb_full + bs, bph);
              if (dbg) w_b += clock64() - tw;
            }
            const uint32_t b_base = b0 + bs * (uint32_t)b_stage_stride;
            const bool last_chunk = (p == a.n_phases - 1) && (cc == kc - 1);
            for (int li2 = 0; li2 < ph_lin; ++li2, ++step) {
              const Tc2Sched s = ph.sched[li2];
              if (s.n_slots == 0) continue;   // (never with two issuing warps, the host checks)
              const bool mine = nw == 1 || (int)(step & 1) == mw;
              if (mine) {
                {
                  const long long tw = dbg ? clock64() : 0;
                  t2::wait(a_full + as, aph);
                  if (CG == 2) t2::wait(pa_full + as, aph);
                  if (dbg) w_a += clock64() - tw;
                }
                const uint32_t a_base = a0 + as * (uint32_t)a_stage_bytes;
                const uint64_t da_hi = desc_a0 | (uint64_t)((a_base & 0x3FFFF) >> 4);
                const uint64_t da_lo = desc_a0 | (uint64_t)(((a_base + kTcBlockBytes) & 0x3FFFF) >> 4);
                // first K chunk: the leading n_acc positions of the window already hold a partial sum, the rest are
                // written for the first time -> two runs with their own accumulate flag; afterwards one run
                const int n_first = (cc == 0) ? s.n_acc : s.n_slots;
                const bool my_last_in_chunk = nw == 1 ? (li2 == ph_lin - 1) : (li2 >= ph_lin - 2);
                if (nw == 2) { t2::token_wait(mw); umma::tc_fence_after(); }
#pragma unroll 1
                for (int run = 0; run < 2; ++run) {
                  const int sl0 = run == 0 ? 0 : n_first;
                  const int n = run == 0 ? n_first : s.n_slots - n_first;
                  if (n <= 0) continue;
                  const uint32_t idesc = umma::make_idesc(E::kFmt, kTcRows * CG, n * l_ct);
                  const uint32_t d = acc0 + (uint32_t)(ph_dcol + (s.lo_begin + sl0) * ph_cstep);
                  const uint32_t b_off = b_base + (uint32_t)((s.slot_begin + sl0) * ctl * 128);
                  const uint64_t db_hi = desc0 | (uint64_t)((b_off & 0x3FFFF) >> 4);
                  const uint64_t db_lo = desc0 | (uint64_t)(((b_off + b_lo_off) & 0x3FFFF) >> 4);
                  const uint32_t accf = run == 0 ? 1u : 0u;
                  if (!(ablate & 2) && umma::elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {   // 32-byte K steps inside the 128-byte swizzle atom
                      const uint32_t acc = accf | (uint32_t)(ks > 0);
                      if (CG == 2) {
                        if (r.split) {
                          t2::mma_f16_cg2(d, da_lo + kTcActKStep * ks, db_hi + 2 * ks, idesc, acc);
                          t2::mma_f16_cg2(d, da_hi + kTcActKStep * ks, db_lo + 2 * ks, idesc, 1u);
                          t2::mma_f16_cg2(d, da_hi + kTcActKStep * ks, db_hi + 2 * ks, idesc, 1u);
                        } else {
                          t2::mma_f16_cg2(d, da_hi + kTcActKStep * ks, db_hi + 2 * ks, idesc, acc);
                        }
                      } else {
                        if (r.split) {
                          umma::mma_bf16(d, da_lo + kTcActKStep * ks, db_hi + 2 * ks, idesc, acc);
                          umma::mma_bf16(d, da_hi + kTcActKStep * ks, db_lo + 2 * ks, idesc, 1u);
                          umma::mma_bf16(d, da_hi + kTcActKStep * ks, db_hi + 2 * ks, idesc, 1u);
                        } else {
                          umma::mma_bf16(d, da_hi + kTcActKStep * ks, db_hi + 2 * ks, idesc, acc);
                        }
                      }
                    }
                  }
                  __syncwarp();
                }
                if (umma::elect_one()) {
                  // frees the A stage (in both CTAs of a pair) once these MMAs have read it; after this warp's last
                  // step of the chunk / the tile also the weight stage / the accumulator
                  if (CG == 2) t2::commit_cg2(a_empty + as); else umma::mma_commit(a_empty + as);
                  if (my_last_in_chunk) { if (CG == 2) t2::commit_cg2(b_empty + bs); else umma::mma_commit(b_empty + bs); }
                  if (my_last_in_chunk && last_chunk) { if (CG == 2) t2::commit_cg2(acc_full + ab); else umma::mma_commit(acc_full + ab); }
                }
                __syncwarp();
                if (nw == 2) { umma::tc_fence_before(); t2::token_pass(mw ^ 1); }
              }
              if (++as == (uint32_t)r.a_stages) { as = 0; aph ^= 1; }
            }
            if (++bs == (uint32_t)r.b_stages) { bs = 0; bph ^= 1; }
          }
        }
      }
      }
      if (nw == 2 && (int)(step & 1) == mw) t2::token_wait(mw);   // consume the last token
      if (dbg && lane == 0 && mw == 0) { dbg[2] = w_acc; dbg[3] = w_b; dbg[4] = w_a; dbg[5] = clock64() - t_begin; }
    }
  } else {
    // ===== epilogue: 16 warps; a thread owns one accumulator lane (trajectory row) and every 4th 16-column unit =====
    const int quarter = warp & 3;                         // TMEM lane quarter this warp may access
    const int part = (warp - kT2EpiWarp0) >> 2;           // which share of the column units (0..3)
    const int et = threadIdx.x - kT2EpiWarp0 * 32;
    const int row_local = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float* my_part = s_part + row_local;   // [piece][mean | M2][128 rows]
    uint32_t xg_par = 0, xg_ph = 0;
    uint32_t hmax = 0;                                    // largest |hi| half pattern stored (operand range check)
    // accumulator buffers, mirroring the MMA warps: `buf` = next buffer of a double-buffered tile, nf[b] = how often
    // acc_full[b] has been waited for; a single-buffered tile uses buffer 0's barrier and releases BOTH buffers
    uint32_t buf = 0, nf0 = 0, nf1 = 0;
    long long w_full = 0, t_busy = 0, t_stats = 0, t_bar = 0, t_fin = 0, t_par = 0;
    int tile_par = 0;
    for (int li = 0; li < r.n; ++li) {
    const Tc2Args& a = r.l[LX(li)];
    const int n_tiles = layer_tiles(li);
    const int L = a.lout, ct = a.ct, cg = a.cg;
    const int n_units = (L * ct) >> 4;
    const bool two = cg == 8;                             // a 16-column unit spans two GroupNorm groups
    const float inv_n = 1.0f / (float)(cg * L);
    const float sc0 = a.ph[0].acc_scale, sc1 = a.ph[1].acc_scale;
    const int kch_out = a.cout >> E::kShift;
    const int half_cols = (L * ct) >> 1;                  // half_layout: columns of one half
    const int ct_log2 = a.ct_log2, cg_log2 = a.cg_log2;
    const int n_pieces_alloc = n_units * (two ? 2 : 1);
    float* my_stat = my_part + n_pieces_alloc * 256;   // [group][mean | rstd][128 rows]
    const int gw = min(cg, ct);                           // channels of a group inside this column tile (ct < cg: column split)
    const int n_groups = two ? (ct >> 3) : max(1, ct >> cg_log2);
    const float inv_pieces = 1.0f / (float)(L * (two ? 1 : (gw >> 4)));
    float* s_xg = my_part + (n_pieces_alloc + n_groups) * 256;   // nsplit > 1: [tile parity][source rank][mean | M2][128 rows]
    for (int t = first_tile(li); t < n_tiles; t += n_walkers) {
      const int nt = t & (a.n_col_tiles - 1);
      const int rt = (t >> a.nct_log2) * CG + (int)rank;
      const int grow = rt * kTcRows + row_local;
      // ---- per-channel parameters of this column tile (double-buffered by tile parity: one barrier per tile) ----
      const long long tp0 = dbg ? clock64() : 0;
      float* s_par = s_par2[tile_par];
      tile_par ^= 1;
      if (et < ct) {
        const int c = nt * ct + et;
        s_par[et] = a.bias[c];
        s_par[128 + et] = a.gamma ? a.gamma[c] : 1.0f;
        s_par[256 + et] = a.beta ? a.beta[c] : 0.0f;
        s_par[384 + et] = a.temb ? a.temb[c] : 0.0f;
        s_par[512 + et] = a.bres ? a.bres[c] : 0.0f;
      }
      t2::bar_epilogue();   // also: every thread is past the previous tile's reads of the GroupNorm pieces
      if (dbg) t_par += clock64() - tp0;
      uint32_t ab;                               // accumulator buffer of this tile
      {
        const long long tw = dbg ? clock64() : 0;
        if (a.acc_bufs == 2) { ab = buf; buf ^= 1; } else { ab = 0; buf = 0; }
        t2::wait(acc_full + ab, (ab ? nf1 : nf0) & 1u);
        if (ab) ++nf1; else ++nf0;
        if (dbg) w_full += clock64() - tw;
      }
      const long long t_start = dbg ? clock64() : 0;
      __syncwarp();
      umma::tc_fence_after();
      const uint32_t t_acc = t_lane + ab * (uint32_t)a.acc_stride;

      // ---- my units: unit u = part + 4k covers accumulator columns [16u, 16u+16); two units per batch ----
      // (position, first channel within the tile) of the unit at accumulator column `col`
      auto unit_pos = [&](int col, int& lo, int& c0) {
        if (a.half_layout) {
          const int h = col >= half_cols ? 1 : 0, rem = col - h * half_cols;
          lo = rem >> (ct_log2 - 1);
          c0 = h * (ct >> 1) + (rem & ((ct >> 1) - 1));
        } else {
          lo = col >> ct_log2;
          c0 = col & (ct - 1);
        }
      };
      long long tf0 = 0;
      if (!(ablate & 1)) {
      if (a.mode != TC_BIAS) {
        // GroupNorm(8) over (cg channels x L positions) of this row (blocks.py:24-26).  Every 16-column unit (8-column
        // meet in shared memory among the four warps of this lane quarter and are combined with
        // M2 = sum M2_i + n_i * sum (mean_i - mean)^2 -- no E[x^2] - mean^2 cancellation anywhere.
#pragma unroll 1
        for (int kb = 0; kb < 4; kb += 2) {
          if (part + 4 * kb >= n_units) break;
          uint32_t raw[2][16];
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (part + 4 * (kb + j) < n_units)
              t2::tmem_ld16_nowait(t_acc + (uint32_t)(a.ph[0].d_col + ((part + 4 * (kb + j)) << 4)), raw[j]);
          t2::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int u = part + 4 * (kb + j);
            if (u < n_units) {
              int lo, c0;
              unit_pos(u << 4, lo, c0);
              float b[16];
              pm_ld_par16(s_par + c0, b);
              const f2::f32x2 sc2 = f2::dup(sc0);
              f2::f32x2 v2[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                v2[i] = f2::fma(f2::pku(raw[j][2 * i], raw[j][2 * i + 1]), sc2, f2::pk(b[2 * i], b[2 * i + 1]));
              const float s0 = f2::hsum(f2::add(f2::add(v2[0], v2[1]), f2::add(v2[2], v2[3])));
              const float s1 = f2::hsum(f2::add(f2::add(v2[4], v2[5]), f2::add(v2[6], v2[7])));
              const float m0 = two ? s0 * 0.125f : (s0 + s1) * 0.0625f, m1 = two ? s1 * 0.125f : m0;
              const f2::f32x2 mm0 = f2::dup(m0), mm1 = f2::dup(m1);
              f2::f32x2 qa = f2::dup(0.0f), qb = f2::dup(0.0f);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const f2::f32x2 da = f2::sub(v2[i], mm0), db = f2::sub(v2[4 + i], mm1);
                qa = f2::fma(da, da, qa);
                qb = f2::fma(db, db, qb);
              }
              const float q0 = f2::hsum(qa), q1 = f2::hsum(qb);
              if (two) {
                my_part[(2 * u) * 256] = m0; my_part[(2 * u) * 256 + 128] = q0;
                my_part[(2 * u + 1) * 256] = m1; my_part[(2 * u + 1) * 256 + 128] = q1;
              } else {
                my_part[u * 256] = m0; my_part[u * 256 + 128] = q0 + q1;
              }
            }
          }
        }
        const long long tb0 = dbg ? clock64() : 0;
        if (dbg) t_stats += tb0 - t_start;
        t2::bar_quarter(quarter);
        // one thread per (row, group) combines the pieces; (mean, rstd) go back through shared memory
        for (int g = part; g < n_groups; g += 4) {
          // the pieces of group g: for every position, `cnt` consecutive pieces from `p0` on, per segment (a group
          // may straddle the two halves of the [half][position][ct/2] accumulator layout of a CTA pair)
          int p0[2], cnt[2], pstep, nseg = 1;
          float npiece;             // elements per piece
          if (two) {
            p0[0] = g; cnt[0] = 1; pstep = 2 * (ct >> 4); npiece = 8.0f;
          } else if (a.half_layout) {
            const int ch = ct >> 1, g0 = g << cg_log2, g1 = g0 + cg;
            nseg = 0;
            for (int h = 0; h < 2; ++h) {
              const int cs = max(g0, h * ch), ce = min(g1, (h + 1) * ch);
              if (cs < ce) { p0[nseg] = (h * half_cols + (cs - h * ch)) >> 4; cnt[nseg] = (ce - cs) >> 4; ++nseg; }
            }
            pstep = ch >> 4; npiece = 16.0f;
          } else {
            p0[0] = ((g << cg_log2) & (ct - 1)) >> 4; cnt[0] = gw >> 4; pstep = ct >> 4; npiece = 16.0f;
          }
          // (four independent shared-memory loads per step: the serial load -> add chain over 13..16 pieces was 10-19 %
          // of the horizon 7 / 13 epilogues; cnt is a power of two)
          float sm4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
          for (int sg = 0; sg < nseg; ++sg) {
            const int cshift = 31 - __clz(cnt[sg]), n_i = L << cshift;
            const float* base = my_part + p0[sg] * 256;
            for (int i0 = 0; i0 < n_i; i0 += 4) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int ii = i0 + j;
                if (ii < n_i) sm4[j] += base[((ii >> cshift) * pstep + (ii & (cnt[sg] - 1))) * 256];
              }
            }
          }
          const float mu = ((sm4[0] + sm4[1]) + (sm4[2] + sm4[3])) * inv_pieces;
          float sq4[4] = {0.0f, 0.0f, 0.0f, 0.0f}, sd4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
          for (int sg = 0; sg < nseg; ++sg) {
            const int cshift = 31 - __clz(cnt[sg]), n_i = L << cshift;
            const float* base = my_part + p0[sg] * 256;
            for (int i0 = 0; i0 < n_i; i0 += 4) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int ii = i0 + j;
                if (ii < n_i) {
                  const float* pp = base + ((ii >> cshift) * pstep + (ii & (cnt[sg] - 1))) * 256;
                  const float d = pp[0] - mu;
                  sq4[j] += pp[128];
                  sd4[j] = fmaf(d, d, sd4[j]);
                }
              }
            }
          }
          const float sq = (sq4[0] + sq4[1]) + (sq4[2] + sq4[3]), sd = (sd4[0] + sd4[1]) + (sd4[2] + sd4[3]);
          float mean = mu, m2 = fmaf(npiece, sd, sq);
          if (nsplit > 1) {
            // this tile holds ct of the group's cg channels: exchange (mean, M2) with the cluster peers that hold the rest
            // (same row tile, neighbouring column tiles, same walk step) and combine in rank order, so that every CTA of
            // the cluster gets bit-identical statistics.  Only part 0 comes here (one group per tile): a warp per lane quarter.
            float* xg = s_xg + xg_par * (4 * 256);
            for (uint32_t pr = 0; pr < nsplit; ++pr)
              if (pr != crank) {
                t2::st_remote_f32(xg + crank * 256, pr, mean);
                t2::st_remote_f32(xg + crank * 256 + 128, pr, m2);
              }
            __syncwarp();
            if (lane == 0)
              for (uint32_t pr = 0; pr < nsplit; ++pr)
                if (pr != crank) t2::mbar_arrive_remote(xg_full + xg_par, pr);
            t2::wait_cluster(xg_full + xg_par, xg_ph);
            float ms[4], qs[4], msum = 0.0f;
            for (uint32_t pr = 0; pr < nsplit; ++pr) {
              ms[pr] = pr == crank ? mean : xg[pr * 256];
              qs[pr] = pr == crank ? m2 : xg[pr * 256 + 128];
              msum += ms[pr];
            }
            const float mall = msum / (float)nsplit;
            float qall = 0.0f, dall = 0.0f;
            for (uint32_t pr = 0; pr < nsplit; ++pr) { const float d = ms[pr] - mall; qall += qs[pr]; dall = fmaf(d, d, dall); }
            mean = mall;
            m2 = fmaf((float)(gw * L), dall, qall);
          }
          my_stat[g * 256] = mean;
          my_stat[g * 256 + 128] = rsqrtf(m2 * inv_n + 1e-5f);
        }
        if (nsplit > 1) { xg_par ^= 1; if (xg_par == 0) xg_ph ^= 1; }
        t2::bar_quarter(quarter);
        if (dbg) t_bar += clock64() - tb0;
      }
      tf0 = dbg ? clock64() : 0;

      // ---- normalise, Mish, time embedding, residual, hi/lo split, store.  Every load of a batch (accumulator,
      // residual accumulator, identity-residual operand blocks) is issued first; the group statistics are combined
      // from the shared-memory pieces while those loads are in flight ----
#pragma unroll 1
      for (int kb = 0; kb < 8; ++kb) {   // (up to 16 units per tile with GroupNorm, up to 32 for a bias-only layer)
        if (part + 4 * kb >= n_units) break;
        uint32_t yr[1][16], rr[1][16];   // rr: residual accumulator, or the identity residual's hi (0..7) / lo (8..15) chunks
        float mean[1][2], rstd[1][2];
#pragma unroll
        for (int j = 0; j < 1; ++j) {
          const int u = part + 4 * (kb + j);
          if (u < n_units) {
            int lo, c0;
            unit_pos(u << 4, lo, c0);
            t2::tmem_ld16_nowait(t_acc + (uint32_t)(a.ph[0].d_col + (u << 4)), yr[j]);
            if (a.mode == TC_GN_RES_PW) t2::tmem_ld16_nowait(t_acc + (uint32_t)(a.ph[1].d_col + lo * ct + c0), rr[j]);
            if (a.mode == TC_GN_RES_ID) {
              // out + x (blocks.py:164, identity residual): x = hi + lo of the tiled block input
              const int kr = lo * a.res.C + nt * ct + c0;
              const size_t rsrc = ((size_t)rt * (L * (a.res.C >> E::kShift)) + (kr >> E::kShift)) * kTcBlockBytes;
              const int chr = (kr & (E::kCpc - 1)) >> 3;
#pragma unroll
              for (int m = 0; m < 2; ++m) {
                const size_t off = rsrc + tc_swz_bytes(row_local, chr + m);
                const uint4 qh = *reinterpret_cast<const uint4*>((const uint8_t*)a.res.hi + off);
                const uint4 ql = a.res.lo ? *reinterpret_cast<const uint4*>((const uint8_t*)a.res.lo + off) : make_uint4(0, 0, 0, 0);
                rr[j][4 * m] = qh.x; rr[j][4 * m + 1] = qh.y; rr[j][4 * m + 2] = qh.z; rr[j][4 * m + 3] = qh.w;
                rr[j][8 + 4 * m] = ql.x; rr[j][8 + 4 * m + 1] = ql.y; rr[j][8 + 4 * m + 2] = ql.z; rr[j][8 + 4 * m + 3] = ql.w;
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 1; ++j) {
          mean[j][0] = mean[j][1] = 0.0f;
          rstd[j][0] = rstd[j][1] = 1.0f;
          const int u = part + 4 * (kb + j);
          if (a.mode != TC_BIAS && u < n_units) {
            int lo, c0;
            unit_pos(u << 4, lo, c0);
            const int g = two ? (c0 >> 3) : (c0 >> cg_log2);
            mean[j][0] = my_stat[g * 256]; rstd[j][0] = my_stat[g * 256 + 128];
            if (two) { mean[j][1] = my_stat[(g + 1) * 256]; rstd[j][1] = my_stat[(g + 1) * 256 + 128]; }
          }
        }
        t2::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 1; ++j) {
          const int u = part + 4 * (kb + j);
          if (u < n_units) {
            int lo, c0;
            unit_pos(u << 4, lo, c0);
            const int kk = lo * a.cout + nt * ct + c0;            // K index (position-major) of the unit's first channel
            const int chunk = (kk & (E::kCpc - 1)) >> 3;          // 16-byte chunk inside the 128-byte row
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              // (all dense fp32 math in packed f32x2 form, see umma.cuh f2::)
              f2::f32x2 v2[4];
              {
                const float* pb0 = s_par + c0 + m * 8;
                const float4 z0 = *reinterpret_cast<const float4*>(pb0), z1 = *reinterpret_cast<const float4*>(pb0 + 4);
                const f2::f32x2 sc2 = f2::dup(sc0);
                const f2::f32x2 bi[4] = {f2::pk(z0.x, z0.y), f2::pk(z0.z, z0.w), f2::pk(z1.x, z1.y), f2::pk(z1.z, z1.w)};
#pragma unroll
                for (int e = 0; e < 4; ++e) v2[e] = f2::fma(f2::pku(yr[j][m * 8 + 2 * e], yr[j][m * 8 + 2 * e + 1]), sc2, bi[e]);
              }
              if (a.mode != TC_BIAS) {
                const float m_ = (two && m) ? mean[j][1] : mean[j][0], r_ = (two && m) ? rstd[j][1] : rstd[j][0];
                const float* pg = s_par + 128 + c0 + m * 8;
                const float4 g0 = *reinterpret_cast<const float4*>(pg), g1 = *reinterpret_cast<const float4*>(pg + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(pg + 128), b1 = *reinterpret_cast<const float4*>(pg + 132);
                const float4 e0 = *reinterpret_cast<const float4*>(pg + 256), e1 = *reinterpret_cast<const float4*>(pg + 260);
                const f2::f32x2 ga[4] = {f2::pk(g0.x, g0.y), f2::pk(g0.z, g0.w), f2::pk(g1.x, g1.y), f2::pk(g1.z, g1.w)};
                const f2::f32x2 be[4] = {f2::pk(b0.x, b0.y), f2::pk(b0.z, b0.w), f2::pk(b1.x, b1.y), f2::pk(b1.z, b1.w)};
                const f2::f32x2 te[4] = {f2::pk(e0.x, e0.y), f2::pk(e0.z, e0.w), f2::pk(e1.x, e1.y), f2::pk(e1.z, e1.w)};
                const f2::f32x2 mm = f2::dup(m_), rr2 = f2::dup(r_);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  v2[e] = f2::mish_add(f2::fma(f2::sub(v2[e], mm), f2::mul(rr2, ga[e]), be[e]), te[e]);
              }
              if (a.mode == TC_GN_RES_PW) {
                const float* pb = s_par + 512 + c0 + m * 8;
                const float4 q0 = *reinterpret_cast<const float4*>(pb), q1 = *reinterpret_cast<const float4*>(pb + 4);
                const f2::f32x2 br[4] = {f2::pk(q0.x, q0.y), f2::pk(q0.z, q0.w), f2::pk(q1.x, q1.y), f2::pk(q1.z, q1.w)};
                const f2::f32x2 sc12 = f2::dup(sc1);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  v2[e] = f2::add(v2[e], f2::fma(f2::pku(rr[j][m * 8 + 2 * e], rr[j][m * 8 + 2 * e + 1]), sc12, br[e]));
              } else if (a.mode == TC_GN_RES_ID) {
                float x[8];
                tc_chunk_sum<EL>(make_uint4(rr[j][4 * m], rr[j][4 * m + 1], rr[j][4 * m + 2], rr[j][4 * m + 3]),
                                 make_uint4(rr[j][8 + 4 * m], rr[j][8 + 4 * m + 1], rr[j][8 + 4 * m + 2], rr[j][8 + 4 * m + 3]),
                                 a.res.lo != nullptr, x);
#pragma unroll
                for (int e = 0; e < 4; ++e) v2[e] = f2::add(v2[e], f2::pk(x[2 * e], x[2 * e + 1]));
              }
              // hi/lo split of 8 values -> one 16-byte chunk each
              uint32_t hh[4], ll[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float x0, x1, d0, d1;
                f2::upk(v2[e], x0, x1);
                hh[e] = pack16x2<EL>(x0, x1);
                range_track<EL>(hmax, hh[e]);
                const float2 hf = unpack16x2<EL>(hh[e]);
                f2::upk(f2::sub(v2[e], f2::pk(hf.x, hf.y)), d0, d1);
                ll[e] = pack16x2<EL>(d0, d1);
              }
              if (grow < a.rows) {
                size_t dst;
                if (a.out_pm) {
                  // hand-over to the position-major levels: [row block of 8][lo + 2][8 rows][cout halves]
                  const int c = nt * ct + c0;
                  dst = (size_t)(grow >> 3) * pm_img_bytes(L, a.cout) + pm_act_off(L, lo, grow & 7, (c >> 3) + m);
                } else {
                  dst = ((size_t)rt * (L * kch_out) + (kk >> E::kShift)) * kTcBlockBytes + tc_swz_bytes(row_local, chunk + m);
                }
                *reinterpret_cast<uint4*>((uint8_t*)a.out_hi + dst) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                if (a.out_lo) *reinterpret_cast<uint4*>((uint8_t*)a.out_lo + dst) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
              }
            }
          }
        }
      }
      }   // !(ablate & 1)
      {
        // every accumulator read of this tile is complete: hand the TMEM buffer back to the MMA warp
        umma::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2 && rank == 1) t2::mbar_arrive_remote(acc_empty + ab, 0); else umma::mbar_arrive(acc_empty + ab);
          if (a.acc_bufs != 2) {   // a single-buffered tile held both buffers
            if (CG == 2 && rank == 1) t2::mbar_arrive_remote(acc_empty + 1, 0); else umma::mbar_arrive(acc_empty + 1);
          }
        }
      }
      if (a.done != nullptr) {
        // publish the tile: the barrier orders every epilogue thread's stores before thread 0's GPU-scope fence (fences
        // are cumulative over what happens-before them), then that thread bumps the row tile's counter.  (A fence in
        // every thread cost ~7 % of the epilogue's stall samples: MEMBAR + CCTL.IVALL per thread and tile.)
        t2::bar_epilogue();
        if (et == 0) { __threadfence(); atomicAdd(a.done + rt, 1); }
      }
      if (dbg) { const long long te = clock64(); t_busy += te - t_start; t_fin += te - tf0; }
    }
    }
    if (dbg && et == 0) { dbg[6] = w_full; dbg[8] = t_busy; dbg[9] = t_stats; dbg[10] = t_bar; dbg[11] = t_fin; dbg[12] = t_par; }
    range_report<EL>(hmax, r.l[0].range_flag);
    umma::tc_fence_before();
  }
  __syncthreads();
  if (CG == 2 || nsplit > 1) t2::cluster_sync_all();   // (no CTA leaves while a peer may still write its shared memory)
  if (warp == 1) {
    umma::tc_fence_after();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    else umma::tmem_dealloc<512>(tmem_base);
  }
  if (dbg && threadIdx.x == 0) dbg[7] = clock64();
}

}  // namespace edmp
