// Persistent position-major tensor-core kernel for the horizon 25 / 50 levels: second generation of conv_pm.cuh
// (same GEMM view -- M lanes = (position, row) of a block of 8 trajectory rows, a filter tap = a descriptor start
// offset -- and the same operand images and epilogue arithmetic).
//
// conv_pm_kernel is latency-bound: a CTA loads its operands (~3k cycles), issues its MMAs (~5-10k), waits for them and
// then runs a ~20k-cycle epilogue whose warps sit idle before that; with two CTAs per SM only ~1 epilogue is active on
// an SM at any time and its issue slots are 14 % used (profiles/r1_f16x3_8190rows_cta_phases_v3.txt, ncu source
// counters).  Here ONE CTA per SM stays resident and walks row blocks:
//   * the layer's weights (all output channels: the MMA's N is the layer's C_out, 32 or 64) are loaded once per CTA;
//   * the activation image of block k+1 is fetched as soon as the MMAs of block k have read theirs;
//   * two TMEM accumulator buffers and TWO epilogue groups (8 warps each, blocks alternate between them): two
//     epilogues are always in flight next to the load + MMA of the following block.
// Reference ops covered: as conv_pm.cuh (Conv1dBlock blocks.py:13-34, ResidualConvolutionBlock :137-166, the stride-2
// Conv1d :211, ConvTranspose1d :249, final_conv temporalunet.py:35-36).
//
// PAIR = 1 (round 2, opt-in with EDMP_PM_PAIR=1): cta_group::2 -- a CTA pair walks DOUBLE row blocks with M = 256 MMAs
// (half the MMAs per row block), and the two CTAs ALTERNATE as the issuer (unit k is led by CTA k & 1, which is also the
// accumulator buffer / epilogue group k & 1; the odd CTA of a pair may issue cta_group::2 MMAs).  Each CTA stages its own
// row block's image and half of the layer's weight rows, the non-leading CTA relays its "image full" barrier, the
// leader's commits multicast "image empty" / "accumulator full" to both CTAs, both CTAs' epilogue groups release buffer g
// on CTA g's barrier.  Hypotheses tested: the kernel is bound (a) by the NUMBER of MMAs it issues, (b) by the issue rate
// of one warp per SM.  Measured: neither -- every layer is 4-6 us SLOWER with pairs, with and without alternating
// leadership (8190 rows: 64.5 -> 68.3 us for down_samplers.0.down.0.blocks.0).  What binds the MMA warp (busy ~90k of a
// layer's 110k cycles) is the shared-memory fetch of the A operand: 4 KB per M = 128 MMA (+1-2 KB of B at N = 32 / 64) is
// 40-48 cycles at 128 B/clk, paid by each of the three split products, and a pair does not share A.  Kept as a tested variant.
#pragma once
#include "conv_pm.cuh"
#include "conv_tc2.cuh"   // t2:: cluster / cta_group::2 helpers

namespace edmp {

constexpr int kPm2Groups = 2;
constexpr int kPm2Threads = 64 + kPm2Groups * kPmEpiThreads;   // producer, MMA, 2 x 8 epilogue warps = 576

// One precomputed MMA step: the shared-memory addresses and the TMEM offset do not depend on the row block, so the
// whole issue schedule of a block is tabulated once per CTA; the issuing warp then only streams table entries (the
// descriptor arithmetic of conv_pm_kernel's loop cost ~100 cycles per MMA against the 40-cycle issue floor).
struct __align__(16) PmIssue {   // low words of the four UMMA descriptors (the high words are per-kernel constants)
  uint32_t da_hi, da_lo, db_hi, db_lo;
  uint32_t d_off, acc, pad0, pad1;
};
constexpr int kPm2MaxIssue = 128;

__device__ __forceinline__ void pm2_group_barrier(int g) { asm volatile("bar.sync %0, %1;" ::"r"(2 + g), "n"(kPmEpiThreads) : "memory"); }


// The issue loop of one row block, run by ONE elected thread (no per-entry elect / branch / reconvergence, the arithmetic
// mode is a compile-time constant, the next table entry is fetched while the current MMAs issue).  Ablation traces
// (tools/gpu_ablate.sh, profiles/r2_ablation_8190rows.txt): with the epilogue idle the warp-uniform loop with a per-entry
// elect took ~200 cycles per table entry against ~90 of MMA time (N = 64 + N = 32) -- the issue loop, not the tensor
// pipe, bound the MMA warp.  MODE 0: one product, 1: three split products, 2: two-MMA form (PmArgs::fuse_b).
template <int MODE, int PAIR>
__device__ __forceinline__ void pm2_issue_block(const PmIssue* tab, int n_issue, uint32_t acc0, uint64_t hi_a, uint64_t hi_b,
                                                uint32_t idesc, uint32_t idesc2, uint32_t cout) {
  uint4 c0 = *reinterpret_cast<const uint4*>(tab);                 // da_hi, da_lo, db_hi, db_lo
  uint2 c1 = *reinterpret_cast<const uint2*>(&tab->d_off);         // d_off, acc
#pragma unroll 2
  for (int e = 0; e < n_issue; ++e) {
    const PmIssue* nx = tab + (e + 1 < n_issue ? e + 1 : e);
    const uint4 n0 = *reinterpret_cast<const uint4*>(nx);
    const uint2 n1 = *reinterpret_cast<const uint2*>(&nx->d_off);
    const uint32_t d = acc0 + c1.x;
    const uint64_t a_hi = hi_a | c0.x, a_lo = hi_a | c0.y, b_hi = hi_b | c0.z, b_lo = hi_b | c0.w;
    if (PAIR) {
      if (MODE == 1) {
        t2::mma_f16_cg2(d, a_lo, b_hi, idesc, c1.y);
        t2::mma_f16_cg2(d, a_hi, b_lo, idesc, 1u);
        t2::mma_f16_cg2(d, a_hi, b_hi, idesc, 1u);
      } else {
        t2::mma_f16_cg2(d, a_hi, b_hi, idesc, c1.y);
      }
    } else if (MODE == 2) {
      // [D_a | D_b] (+)= x_hi * [w_lo ; w_hi] first (one accumulate flag for both halves), then D_b += x_lo * w_hi
      umma::mma_bf16(d, a_hi, b_lo, idesc2, c1.y);
      umma::mma_bf16(d + cout, a_lo, b_hi, idesc, 1u);
    } else if (MODE == 1) {
      umma::mma_bf16(d, a_lo, b_hi, idesc, c1.y);
      umma::mma_bf16(d, a_hi, b_lo, idesc, 1u);
      umma::mma_bf16(d, a_hi, b_hi, idesc, 1u);
    } else {
      umma::mma_bf16(d, a_hi, b_hi, idesc, c1.y);
    }
    c0 = n0; c1 = n1;
  }
}

// CG = channels per GroupNorm group (4: C_out = 32, 8: C_out = 64); the CTA computes all C_out = 8 * CG channels.
template <int EL, int CG, int PAIR>
__global__ void __launch_bounds__(kPm2Threads, 1) conv_pm2_kernel(const __grid_constant__ PmArgs a) {
  static_assert(EL != TC_EL_TF32, "position-major kernels use 16-bit operand elements");
  constexpr int COUT = 8 * CG;
  constexpr int UNITS = COUT / 16;            // 16-column accumulator units per M tile
  constexpr int NG = 8;                       // GroupNorm groups
  constexpr int GPU_ = 16 / CG;               // groups per unit (4 or 2)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_w, bar_a_full, bar_a_empty, bar_acc_full[kPm2Groups], bar_acc_empty[kPm2Groups];
  __shared__ uint64_t pw_full, pa_full;      // PAIR, leader: "the peer's weights / image are in"
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_par[5 * 64];       // bias | gamma | beta | temb | aux bias
  __shared__ float s_fw[7 * 64 + 8];                  // final 1x1 conv weights + bias
  __shared__ float s_red[kPm2Groups][kPmEpiWarps][8][8];   // [group][warp][row][gn group] partial sums
  __shared__ float s_mr[kPm2Groups][2][8][8];              // [group][mean | rstd][row][gn group]
  __shared__ PmIssue s_issue[kPm2MaxIssue];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = a.a.C;
  const int rby = 2 * C;                      // bytes per (position, row) line
  const int atom = 8 * rby;
  const int nkc = a.b.C ? 2 : 1;
  const int nparts = a.split ? 2 : 1;
  const int ntiles = (a.n_m + 15) >> 4;
  const int n_blocks = (a.rows + kPmRows - 1) / kPmRows;
  // walk units: row blocks, or double blocks of a CTA pair (rank r of pair u holds row block 2u + r; the odd tail block's
  // partner re-reads the last block and stores nothing)
  constexpr int NC = PAIR ? 2 : 1;
  const uint32_t rank = PAIR ? t2::cluster_ctarank() : 0u;
  const int n_units = (n_blocks + NC - 1) / NC;
  const int unit0 = blockIdx.x / NC, n_walkers = gridDim.x / NC;
  const int w_part_cta = a.w_bytes_part / NC;                    // this CTA's share of one weight part (half of the rows of every slot)
  const int FB = (a.fuse_b && !PAIR) ? 2 : 1;                    // accumulator columns per output channel ([D_a | D_b], PmArgs::fuse_b)
  const int acc_cols = (a.n_groups + a.aux) * ntiles * COUT * FB;   // accumulator columns of one block (<= 256)
  uint8_t* a_smem = smem;                                        // [part][source] images
  uint8_t* w_smem = smem + ((a.a_bytes_total + 1023) & ~1023);   // [part][slot][kc][COUT rows][rby]

  long long* dbg = a.dbg ? a.dbg + (size_t)blockIdx.x * 16 : nullptr;
  const int ablate = dbg ? g_edmp_ablate : 0;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    umma::mbar_init(&bar_w, 1);
    umma::mbar_init(&bar_a_full, 1);
    umma::mbar_init(&bar_a_empty, 1);
    for (int g = 0; g < kPm2Groups; ++g) { umma::mbar_init(bar_acc_full + g, 1); umma::mbar_init(bar_acc_empty + g, kPmEpiWarps * NC); }
    umma::mbar_init(&pw_full, 1);
    umma::mbar_init(&pa_full, 1);
    umma::fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma::smem_u32(&tmem_slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      umma::tmem_alloc<512>(&tmem_slot);
    }
    // issue table: entry e = (M tile, term, K chunk, 32-byte K step)
    const uint32_t a_base = umma::smem_u32(a_smem), w_base = umma::smem_u32(w_smem);
    const int ksteps = C >> 4;
    const int n_issue = ntiles * a.n_terms * nkc * ksteps;
    for (int e = lane; e < n_issue; e += 32) {
      const int ks = e % ksteps, kc = (e / ksteps) % nkc, ti = (e / (ksteps * nkc)) % a.n_terms, mt = e / (ksteps * nkc * a.n_terms);
      const PmTerm t = a.terms[ti];
      uint32_t seen = 0;
      for (int tj = 0; tj < ti; ++tj) seen |= (a.terms[tj].acc == t.acc) ? 1u : 0u;
      const uint32_t a_addr = a_base + (uint32_t)(kc * a.a_bytes_img + (a.stride * 16 * mt + t.off) * 128);
      // weight rows of this (slot, K chunk): COUT / NC rows per part; fuse_b: [lo rows | hi rows] back to back
      const uint32_t w_addr = w_base + (uint32_t)((t.slot * nkc + kc) * (COUT / NC) * rby * FB);
      const uint32_t lbo = (uint32_t)(a.lin + 4) * 128u, ka = (uint32_t)ks * (lbo >> 3);   // K step = two 16-byte chunks
      PmIssue it;
      it.da_hi = (uint32_t)(umma::make_desc_interleaved(a_addr, lbo, a.stride * 128) + ka);
      it.da_lo = (uint32_t)(umma::make_desc_interleaved(a_addr + (uint32_t)(nkc * a.a_bytes_img), lbo, a.stride * 128) + ka);
      if (FB == 2) {
        it.db_lo = (uint32_t)(pm_desc(w_addr, atom, rby) + 2 * ks);                          // [lo | hi]: the N = 2 C_out operand starts at lo
        it.db_hi = (uint32_t)(pm_desc(w_addr + (uint32_t)(COUT * rby), atom, rby) + 2 * ks);  // the hi rows alone
      } else {
        it.db_hi = (uint32_t)(pm_desc(w_addr, atom, rby) + 2 * ks);
        it.db_lo = (uint32_t)(pm_desc(w_addr + (uint32_t)w_part_cta, atom, rby) + 2 * ks);
      }
      it.d_off = (uint32_t)((t.acc * ntiles + mt) * COUT * FB);
      it.acc = (seen | (uint32_t)(kc > 0) | (uint32_t)(ks > 0)) ? 1u : 0u;
      it.pad0 = it.pad1 = 0;
      s_issue[e] = it;
    }
  }
  if (warp >= 2) {
    const int e = threadIdx.x - 64;
    if (e < COUT) {
      s_par[e] = a.bias ? a.bias[e] : 0.0f;
      s_par[64 + e] = a.gamma ? a.gamma[e] : 1.0f;
      s_par[128 + e] = a.beta ? a.beta[e] : 0.0f;
      s_par[192 + e] = a.temb ? a.temb[e] : 0.0f;
      s_par[256 + e] = a.aux_bias ? a.aux_bias[e] : 0.0f;
    }
    if (a.eps) {
      for (int i = e; i < 7 * COUT + 7; i += kPm2Groups * kPmEpiThreads) s_fw[i] = i < 7 * COUT ? a.fw[i] : a.fb[i - 7 * COUT];
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (PAIR) t2::cluster_sync_all();
  umma::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== producer: weights once (they do not depend on the previous kernel), then one activation image per block =====
    if (lane == 0) {
      umma::mbar_arrive_expect_tx(&bar_w, (uint32_t)(nparts * w_part_cta));
      if (PAIR) {
        // this CTA's half of the rows (output channels) of every (slot, K chunk) block of the weight image
        const int blk = COUT * rby, n_sub = a.w_bytes_part / blk;
        for (int p = 0; p < nparts; ++p)
          for (int sb = 0; sb < n_sub; ++sb)
            umma::bulk_g2s(w_smem + (size_t)p * w_part_cta + (size_t)sb * (blk / 2),
                           (const uint8_t*)(p ? a.w_lo : a.w_hi) + (size_t)sb * blk + (size_t)rank * (blk / 2), (uint32_t)(blk / 2), &bar_w);
      } else if (FB == 2) {
        // [lo rows | hi rows] per (slot, K chunk) block
        const int blk = COUT * rby, n_sub = a.w_bytes_part / blk;
        for (int sb = 0; sb < n_sub; ++sb) {
          umma::bulk_g2s(w_smem + (size_t)sb * 2 * blk, (const uint8_t*)a.w_lo + (size_t)sb * blk, (uint32_t)blk, &bar_w);
          umma::bulk_g2s(w_smem + (size_t)sb * 2 * blk + blk, (const uint8_t*)a.w_hi + (size_t)sb * blk, (uint32_t)blk, &bar_w);
        }
      } else {
        for (int p = 0; p < nparts; ++p)
          umma::bulk_g2s(w_smem + (size_t)p * a.w_bytes_part, (const uint8_t*)(p ? a.w_lo : a.w_hi), (uint32_t)a.w_bytes_part, &bar_w);
      }
      pdl_wait();   // activations of the previous kernel are read below
      uint32_t ph = 0;
      for (int un = unit0; un < n_units; un += n_walkers) {
        const int rb = min(un * NC + (int)rank, n_blocks - 1);
        umma::mbar_wait(&bar_a_empty, ph ^ 1);     // the MMAs of the previous block have read the image
        umma::mbar_arrive_expect_tx(&bar_a_full, (uint32_t)(nparts * nkc * a.a_bytes_img));
        for (int p = 0; p < nparts; ++p)
          for (int s = 0; s < nkc; ++s) {
            const PmAct& src = s ? a.b : a.a;
            const uint8_t* g = (const uint8_t*)(p ? src.lo : src.hi) + (size_t)rb * a.a_bytes_img;
            umma::bulk_g2s(a_smem + (size_t)(p * nkc + s) * a.a_bytes_img, g, (uint32_t)a.a_bytes_img, &bar_a_full);
          }
        ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform walk, one elected lane issues) =====
    // PAIR: the two CTAs ALTERNATE as the issuer of the pair's M = 256 MMAs (unit k is led by CTA k & 1 -- which is also
    // the accumulator buffer / epilogue group k & 1): the kernel is bound by the issue rate of ONE warp per SM
    // (~27 cycles per MMA at N = 32), so the pair must issue from both SMs to halve the time, not only the count.
    umma::mbar_wait(&bar_w, 0);
    if (PAIR) {
      if (lane == 0) t2::mbar_arrive_remote(&pw_full, rank ^ 1u);   // "my weights are in"
      __syncwarp();
      umma::mbar_wait(&pw_full, 0);
    }
    const uint32_t idesc = umma::make_idesc(TcElem<EL>::kFmt, 128 * NC, COUT);
    const uint32_t idesc2 = umma::make_idesc(TcElem<EL>::kFmt, 128, 2 * COUT);   // fuse_b: x_hi * [w_lo ; w_hi]
    const int n_issue = ntiles * a.n_terms * nkc * (C >> 4);
    const uint64_t hi_a = umma::make_desc_interleaved(0, 0, a.stride * 128) & 0xFFFFFFFF00000000ull;   // SBO, version, no swizzle
    const uint64_t hi_b = pm_desc(0, atom, rby) & 0xFFFFFFFF00000000ull;
    uint32_t ph = 0, k = 0;
    long long w_acc = 0, w_a = 0, t_begin = dbg ? clock64() : 0;
    for (int un = unit0; un < n_units; un += n_walkers, ++k) {
      const uint32_t g = k & 1, use = k >> 1;
      if (PAIR && g != rank) {
        // the peer leads this unit: tell it that my image is in
        umma::mbar_wait(&bar_a_full, ph);
        if (lane == 0) t2::mbar_arrive_remote(&pa_full, rank ^ 1u);
        __syncwarp();
        ph ^= 1;
        continue;
      }
      long long tw = dbg ? clock64() : 0;
      umma::mbar_wait(bar_acc_empty + g, (use & 1) ^ 1);      // the group(s) have drained this accumulator buffer
      if (dbg) { const long long t1 = clock64(); w_acc += t1 - tw; tw = t1; }
      umma::mbar_wait(&bar_a_full, ph);
      if (PAIR) umma::mbar_wait(&pa_full, use & 1);           // (one relay per unit I lead)
      if (dbg) w_a += clock64() - tw;
      ph ^= 1;
      umma::tc_fence_after();
      const uint32_t acc0 = tmem_base + g * 256u;
      if (umma::elect_one()) {
        const int n_e = (ablate & 2) ? 0 : n_issue;
        if (FB == 2) pm2_issue_block<2, PAIR>(s_issue, n_e, acc0, hi_a, hi_b, idesc, idesc2, (uint32_t)COUT);
        else if (a.split) pm2_issue_block<1, PAIR>(s_issue, n_e, acc0, hi_a, hi_b, idesc, idesc2, (uint32_t)COUT);
        else pm2_issue_block<0, PAIR>(s_issue, n_e, acc0, hi_a, hi_b, idesc, idesc2, (uint32_t)COUT);
        // the image may be overwritten once these MMAs have read it (PAIR: in both CTAs)
        if (PAIR) { t2::commit_cg2(&bar_a_empty); t2::commit_cg2(bar_acc_full + g); }
        else { umma::mma_commit(&bar_a_empty); umma::mma_commit(bar_acc_full + g); }
      }
      __syncwarp();
    }
    if (dbg && lane == 0) { dbg[2] = w_acc; dbg[3] = 0; dbg[4] = w_a; dbg[5] = clock64() - t_begin; }
  } else {
    // ===== epilogue: two groups of 8 warps, blocks alternate between them =====
    const int grp = (warp - 2) >> 3;
    const int ew = (warp - 2) & 7;
    const int et = threadIdx.x - 64 - grp * kPmEpiThreads;
    const int quarter = warp & 3;               // TMEM lane quarter this warp may access (warp id mod 4)
    const int half = ew >> 2;                   // which M tiles (mt & 1 == half)
    const int row = lane & 7;
    const int pos_in_tile = quarter * 4 + (lane >> 3);
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)grp * 256u;
    float (*red)[8][8] = s_red[grp];
    float (*mr)[8][8] = s_mr[grp];
    (void)acc_cols;
    uint32_t k = (uint32_t)grp;
    long long w_full = 0, t_busy = 0, t_stats = 0;
    uint32_t hmax = 0;   // largest |hi| half pattern stored (operand range check, conv_tc.cuh)
    for (int un = unit0 + grp * n_walkers; un < n_units; un += kPm2Groups * n_walkers, k += kPm2Groups) {
      const uint32_t use = k >> 1;
      const int rb = un * NC + (int)rank;         // (past the batch for the odd tail's partner: nothing is stored)
      const int grow = rb * kPmRows + row;        // global trajectory row
      const long long tw0 = dbg ? clock64() : 0;
      umma::mbar_wait(bar_acc_full + grp, use & 1);
      __syncwarp();
      umma::tc_fence_after();
      const long long t_start = dbg ? clock64() : 0;
      if (dbg) w_full += t_start - tw0;

      if (ablate & 1) {
        umma::tc_fence_before();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(bar_acc_empty + grp);
        continue;
      }
      // M tiles of this warp: mt = half + 2 * ti; an (M tile, lane quarter) whose 4 positions all lie past n_m is skipped by
      // the whole warp (horizon 50: 3 of 16, horizon 25: 1 of 8 -- the MMAs computed zero-image positions there)
      auto tile_live = [&](int mt) { return mt < ntiles && 16 * mt + quarter * 4 < a.n_m; };
      // one accumulator unit (16 columns of M tile mt, accumulator group g_acc) -> raw fp32 values; fuse_b: D_a + D_b.
      // Two units (or the two halves of a fuse_b unit) are in flight per wait.
      auto ld_unit_pair = [&](int g_acc, int mt, int u, uint32_t (&x0)[16], uint32_t (&x1)[16]) {
        const uint32_t base = t_lane + (uint32_t)((g_acc * ntiles + mt) * COUT * FB + u * 16);
        t2::tmem_ld16_nowait(base, x0);
        t2::tmem_ld16_nowait(base + (uint32_t)(FB == 2 ? COUT : 16), x1);
      };

      // fuse_b: x0 = D_a + D_b (packed adds, in place)
      auto add_halves = [&](uint32_t (&x0)[16], const uint32_t (&x1)[16]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float lo, hi;
          f2::upk(f2::add(f2::pku(x0[2 * i], x0[2 * i + 1]), f2::pku(x1[2 * i], x1[2 * i + 1])), lo, hi);
          x0[2 * i] = __float_as_uint(lo); x0[2 * i + 1] = __float_as_uint(hi);
        }
      };

      if (a.mode != PM_BIAS) {
        // GroupNorm(8, C) over (CG channels x lout positions) of a row (blocks.py:24-26).  ONE pass over the accumulator:
        // a thread reduces the CG channels of a group at its position to an exact two-pass (sum, M2) piece in registers;
        // the row's mean comes from the sums, then M2 = sum_pieces M2_p + CG * (mean_p - mean)^2 (no E[x^2] - mean^2
        // cancellation, the same decomposition as conv_tc2's pieces) -- two small reductions, no second accumulator read.
        const float inv_n = 1.0f / (float)(CG * a.n_m);
        constexpr float kInvCg = 1.0f / (float)CG;
        float S[2][NG], Q[2][NG], nv[2];
#pragma unroll
        for (int ti = 0; ti < 2; ++ti) {
          nv[ti] = 0.0f;
#pragma unroll
          for (int g = 0; g < NG; ++g) { S[ti][g] = 0.0f; Q[ti][g] = 0.0f; }
        }
        const f2::f32x2 sc2 = f2::dup(a.acc_scale);
        auto stats_unit = [&](int ti, int u, const uint32_t (&x)[16]) {
          float b[16];
          pm_ld_par16(s_par + u * 16, b);
          f2::f32x2 y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = f2::fma(f2::pku(x[2 * i], x[2 * i + 1]), sc2, f2::pk(b[2 * i], b[2 * i + 1]));
#pragma unroll
          for (int k = 0; k < GPU_; ++k) {
            constexpr int PP = CG / 2;             // packed pairs per group
            f2::f32x2 t = y[k * PP];
#pragma unroll
            for (int i = 1; i < PP; ++i) t = f2::add(t, y[k * PP + i]);
            const float sum = f2::hsum(t);
            const f2::f32x2 m2 = f2::dup(sum * kInvCg);
            f2::f32x2 q = f2::dup(0.0f);
#pragma unroll
            for (int i = 0; i < PP; ++i) { const f2::f32x2 d = f2::sub(y[k * PP + i], m2); q = f2::fma(d, d, q); }
            S[ti][u * GPU_ + k] = sum;
            Q[ti][u * GPU_ + k] = f2::hsum(q);
          }
        };
#pragma unroll
        for (int ti = 0; ti < 2; ++ti) {
          const int mt = half + 2 * ti;
          if (!tile_live(mt)) continue;
          uint32_t x0[16], x1[16];
          if (FB == 2) {
#pragma unroll
            for (int u = 0; u < UNITS; ++u) {
              ld_unit_pair(0, mt, u, x0, x1);
              t2::tmem_ld_wait();
              add_halves(x0, x1);
              stats_unit(ti, u, x0);
            }
          } else {
#pragma unroll
            for (int u = 0; u < UNITS; u += 2) {
              ld_unit_pair(0, mt, u, x0, x1);
              t2::tmem_ld_wait();
              stats_unit(ti, u, x0);
              stats_unit(ti, u + 1, x1);
            }
          }
          if ((16 * mt + pos_in_tile) < a.n_m) nv[ti] = (float)CG;
          else {
#pragma unroll
            for (int g = 0; g < NG; ++g) { S[ti][g] = 0.0f; Q[ti][g] = 0.0f; }
          }
        }
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          float r8[NG];
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            if (pass == 0) r8[g] = S[0][g] + S[1][g];
            else {
              const float mean = mr[0][row][g];
              const float d0 = S[0][g] * kInvCg - mean, d1 = S[1][g] * kInvCg - mean;
              r8[g] = fmaf(nv[0] * d0, d0, Q[0][g]) + fmaf(nv[1] * d1, d1, Q[1][g]);
            }
            r8[g] += __shfl_xor_sync(0xffffffffu, r8[g], 8);
            r8[g] += __shfl_xor_sync(0xffffffffu, r8[g], 16);
          }
          if (lane < 8) {
#pragma unroll
            for (int g = 0; g < NG; ++g) red[ew][row][g] = r8[g];
          }
          pm2_group_barrier(grp);
          if (et < 8 * NG) {
            const int r = et / NG, g = et % NG;
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < kPmEpiWarps; ++w) t += red[w][r][g];
            mr[pass][r][g] = pass == 0 ? t * inv_n : rsqrtf(t * inv_n + 1e-5f);
          }
          pm2_group_barrier(grp);
        }
      }

      if (dbg) t_stats += clock64() - t_start;
      const int n_acc = a.n_groups + a.aux;
      for (int g_acc = 0; g_acc < n_acc; ++g_acc) {
        const bool is_aux = a.aux && g_acc == a.n_groups;
        const bool gn = !is_aux && a.mode != PM_BIAS;
        const float sc = is_aux ? a.aux_scale : a.acc_scale;
        const float* pb = is_aux ? s_par + 256 : s_par;
        void* o_hi = is_aux ? a.aux_hi : a.out_hi;
        void* o_lo = is_aux ? a.aux_lo : a.out_lo;
        for (int mt = half; mt < ntiles; mt += 2) {
          if (!tile_live(mt)) continue;
          const int idx = 16 * mt + pos_in_tile;
          const bool valid = idx < a.n_m && grow < a.rows;
          const int lo = is_aux ? idx : a.out_step * idx + a.out_off[g_acc];     // output position
          const size_t img = (size_t)rb * pm_img_bytes(a.lout, a.cout);                   // this row block's image
          const bool res = valid && !is_aux && a.mode == PM_GN_RES;
          f2::f32x2 fin2[7];
#pragma unroll
          for (int j = 0; j < 7; ++j) fin2[j] = f2::dup(0.0f);
          // one 16-channel unit: raw accumulator values (+ the second fuse_b half) -> output chunks
          auto final_unit = [&](int u, const uint32_t (&x)[16], const uint4 (&rh)[2], const uint4 (&rl)[2]) {
            // (all 16 channels at once: two sequential 8-channel halves, 24 fewer live registers, measured 10 % SLOWER on this
            // phase -- the instruction-level parallelism of eight independent pairs matters more than the spills)
            f2::f32x2 v2[8];
            {
              float b[16];
              pm_ld_par16(pb + u * 16, b);
              const f2::f32x2 sc2 = f2::dup(sc);
#pragma unroll
              for (int i = 0; i < 8; ++i) v2[i] = f2::fma(f2::pku(x[2 * i], x[2 * i + 1]), sc2, f2::pk(b[2 * i], b[2 * i + 1]));
            }
            if (gn) {
              float ga[16], be[16], te[16];
              pm_ld_par16(s_par + 64 + u * 16, ga);
              pm_ld_par16(s_par + 128 + u * 16, be);
              pm_ld_par16(s_par + 192 + u * 16, te);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int g = u * GPU_ + (2 * i) / CG;
                const f2::f32x2 A = f2::mul(f2::dup(mr[1][row][g]), f2::pk(ga[2 * i], ga[2 * i + 1]));
                const f2::f32x2 t = f2::fma(f2::sub(v2[i], f2::dup(mr[0][row][g])), A, f2::pk(be[2 * i], be[2 * i + 1]));
                v2[i] = f2::mish_add(t, f2::pk(te[2 * i], te[2 * i + 1]));
              }
            }
            if (!valid) return;
            if (res) {
#pragma unroll
              for (int m = 0; m < 2; ++m) {
                float xr[8];
                tc_chunk_sum<EL>(rh[m], rl[m], a.res.lo != nullptr, xr);
#pragma unroll
                for (int e = 0; e < 4; ++e) v2[m * 4 + e] = f2::add(v2[m * 4 + e], f2::pk(xr[2 * e], xr[2 * e + 1]));
              }
            }
            if (o_hi || (!is_aux && a.tc_hi)) {
              uint4 h[4], l[4];
              tc_split_store2<EL>(v2, (is_aux ? a.aux_lo : (a.tc_hi ? a.tc_lo : a.out_lo)) != nullptr, h, l);
#pragma unroll
              for (int m = 0; m < 2; ++m) {
                range_track<EL>(hmax, h[m].x); range_track<EL>(hmax, h[m].y); range_track<EL>(hmax, h[m].z); range_track<EL>(hmax, h[m].w);
              }
              if (o_hi) {
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                  const size_t off = img + pm_act_off(a.lout, lo, row, 2 * u + m);
                  *reinterpret_cast<uint4*>((uint8_t*)o_hi + off) = h[m];
                  if (o_lo) *reinterpret_cast<uint4*>((uint8_t*)o_lo + off) = l[m];
                }
              } else {
                // rows-as-M tiles of conv_tc.cuh: [row tile][l][c chunk of 64][8 chunks][128 rows x 16 B] (tc_swz_bytes)
                const int rt = grow / kTcRows, rl_ = grow % kTcRows;
                const int kk = lo * a.cout + u * 16;
                const size_t blk = ((size_t)rt * (a.lout * (a.cout >> 6)) + (kk >> 6)) * kTcBlockBytes;
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                  const size_t off = blk + tc_swz_bytes(rl_, ((kk & 63) >> 3) + m);
                  *reinterpret_cast<uint4*>((uint8_t*)a.tc_hi + off) = h[m];
                  if (a.tc_lo) *reinterpret_cast<uint4*>((uint8_t*)a.tc_lo + off) = l[m];
                }
              }
            }
            if (!is_aux && a.eps) {
              // fused final nn.Conv1d(32, 7, 1) (temporalunet.py:36): partial sums over this unit's channels
#pragma unroll
              for (int j = 0; j < 7; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  fin2[j] = f2::fma(f2::pk(s_fw[j * COUT + u * 16 + 2 * i], s_fw[j * COUT + u * 16 + 2 * i + 1]), v2[i], fin2[j]);
            }
          };
          // identity residual (blocks.py:164) of one unit: the loads go out before the accumulator wait that hides their latency
          auto ld_res = [&](int u, uint4 (&rh)[2], uint4 (&rl)[2]) {
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              rh[m] = make_uint4(0, 0, 0, 0); rl[m] = make_uint4(0, 0, 0, 0);
              if (res) {
                const size_t off = img + pm_act_off(a.lout, lo, row, 2 * u + m);
                rh[m] = *reinterpret_cast<const uint4*>((const uint8_t*)a.res.hi + off);
                if (a.res.lo) rl[m] = *reinterpret_cast<const uint4*>((const uint8_t*)a.res.lo + off);
              }
            }
          };
          uint32_t x0[16], x1[16];
          if (FB == 2) {
#pragma unroll
            for (int u = 0; u < UNITS; ++u) {
              uint4 rh[2], rl[2];
              ld_unit_pair(g_acc, mt, u, x0, x1);
              ld_res(u, rh, rl);
              t2::tmem_ld_wait();
              add_halves(x0, x1);
              final_unit(u, x0, rh, rl);
            }
          } else {
            // (one unit per wait here: a second unit's residual chunks in flight would cost 32 more registers than the
            // 96 the 576-thread CTA allows -- measured as spills in the hot loop, +16 % on this phase)
#pragma unroll
            for (int u = 0; u < UNITS; ++u) {
              uint4 rh[2], rl[2];
              t2::tmem_ld16_nowait(t_lane + (uint32_t)((g_acc * ntiles + mt) * COUT + u * 16), x0);
              ld_res(u, rh, rl);
              t2::tmem_ld_wait();
              final_unit(u, x0, rh, rl);
            }
          }
          if (valid && !is_aux && a.eps) {
#pragma unroll
            for (int j = 0; j < 7; ++j) a.eps[((size_t)grow * 7 + j) * a.lout + lo] = f2::hsum(fin2[j]) + s_fw[7 * COUT + j];
          }
        }
      }
      // every accumulator read of this block is complete: hand the buffer back to the MMA warp
      umma::tc_fence_before();
      __syncwarp();
      // (PAIR: buffer / group g is led -- and waited for -- by CTA g of the pair)
      if (lane == 0) { if (PAIR && rank != (uint32_t)grp) t2::mbar_arrive_remote(bar_acc_empty + grp, (uint32_t)grp); else umma::mbar_arrive(bar_acc_empty + grp); }
      if (dbg) t_busy += clock64() - t_start;
    }
    range_report<EL>(hmax, a.range_flag);
    if (dbg && et == 0 && grp == 0) { dbg[6] = w_full; dbg[8] = t_busy; dbg[9] = t_stats; dbg[10] = 0; dbg[11] = t_busy - t_stats; dbg[12] = 0; }
    umma::tc_fence_before();
  }
  __syncthreads();
  if (PAIR) t2::cluster_sync_all();
  if (warp == 1) {
    umma::tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    else umma::tmem_dealloc<512>(tmem_base);
  }
  if (dbg && threadIdx.x == 0) { dbg[1] = dbg[0]; dbg[7] = clock64(); }
}

}  // namespace edmp
