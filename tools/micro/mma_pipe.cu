// Microbenchmark: the conv_tc mainloop in isolation -- bulk-copy producers (A-hi, A-lo, W-hi, W-lo threads) feeding
// a single MMA-issuing warp through mbarrier stages, hi/lo split (3 MMAs per 32-byte K step), no epilogue.
// Bisects what keeps the real mainloop below the tensor-pipe rate: data values, spinning epilogue warps,
// stage counts, wait/commit overhead.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../edmp_b200/csrc -o mma_pipe mma_pipe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace edmp::umma;

struct Params {
  int N;          // window columns per MMA
  int chunks;     // K chunks (B stages consumed)
  int lin;        // A blocks per chunk
  int a_stages, b_stages;
  int spinners;   // extra warps spinning on the accumulator barrier (like the idle epilogue warps)
  int feed;       // 1: real bulk-copy producers, 0: operands assumed resident (barriers pre-armed by never waiting)
  int backoff;    // spinners use nanosleep backoff
  int opt;        // 1: lean issue loop (counters instead of modulo, no fence, commit inside the elected region)
  int two;        // 1: two MMA-issuing warps alternate A steps (warp 1: even, warp 5: odd)
  int fake;       // resident mode extras: bit0 = commit per step to a dummy barrier, bit1 = wait on an already complete barrier per step
};

__global__ void __launch_bounds__(576, 1) pipe_kernel(Params p, const uint8_t* src, size_t src_bytes, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t a_full[8], a_empty[8], b_full[4], b_empty[4], acc_full;
  __shared__ uint64_t dummy_done, dummy_sink;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_stage_bytes = 32768, b_part = p.N * 128, b_stage_bytes = 2 * b_part;
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + p.a_stages * a_stage_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(a_full + i, 2); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(b_full + i, 2); mbar_init(b_empty + i, (p.opt && p.two) ? 2 : 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&dummy_done, 1); mbar_init(&dummy_sink, 1);
    mbar_arrive(&dummy_done);   // phase 0 complete
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
  if (!p.feed) {
    const int op_bytes = p.a_stages * a_stage_bytes + p.b_stages * b_stage_bytes;
    for (int i = threadIdx.x; i < op_bytes / 16; i += blockDim.x) ((uint4*)smem)[i] = ((const uint4*)src)[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint8_t* my_src = src + ((size_t)blockIdx.x * (4u << 20)) % (src_bytes - (8u << 20));

  int prod_kind = -1;
  if (lane == 0 && p.feed) {
    if (warp == 0) prod_kind = 0;
    else if (warp == 2) prod_kind = 1;
    else if (warp == 3) prod_kind = 2;
    else if (warp == 4) prod_kind = 3;
  }
  if (prod_kind >= 0) {
    const bool is_lo = prod_kind & 1;
    int a_it = 0, b_it = 0;
    size_t off = (size_t)prod_kind * (1u << 20);
    for (int cc = 0; cc < p.chunks; ++cc) {
      if (prod_kind >= 2) {
        const int bs = b_it % p.b_stages;
        mbar_wait(b_empty + bs, ((b_it / p.b_stages) & 1) ^ 1);
        mbar_arrive_expect_tx(b_full + bs, (uint32_t)b_part);
        bulk_g2s(b_smem + bs * b_stage_bytes + (is_lo ? b_part : 0), my_src + off, (uint32_t)b_part, b_full + bs);
        off = (off + b_part) % (4u << 20);
        ++b_it;
      } else {
        for (int li = 0; li < p.lin; ++li) {
          const int as = a_it % p.a_stages;
          mbar_wait(a_empty + as, ((a_it / p.a_stages) & 1) ^ 1);
          mbar_arrive_expect_tx(a_full + as, 16384u);
          bulk_g2s(a_smem + as * a_stage_bytes + (is_lo ? 16384 : 0), my_src + off, 16384u, a_full + as);
          off = (off + 16384) % (4u << 20);
          ++a_it;
        }
      }
    }
  }

  if (p.opt && (warp == 1 || (p.two && warp == 5))) {
    const int me = warp == 1 ? 0 : 1, nw = p.two ? 2 : 1;
    uint32_t as = 0, aph = 0, bs = 0, bph = 0;
    const uint64_t desc0 = make_desc_sw128(0);
    const uint32_t idesc = make_idesc(0, 128, p.N);
    const uint32_t a0 = smem_u32(a_smem), b0 = smem_u32(b_smem);
    long long t0 = clock64();
    int step = 0;
    long long spins = 0, spins_b = 0;
    for (int cc = 0; cc < p.chunks; ++cc) {
      if (p.feed) { while (!mbar_try_wait(b_full + bs, bph)) ++spins_b; }
      const uint32_t b_base = b0 + bs * b_stage_bytes;
      const uint64_t db_hi = desc0 | (uint64_t)((b_base & 0x3FFFF) >> 4);
      const uint64_t db_lo = desc0 | (uint64_t)(((b_base + b_part) & 0x3FFFF) >> 4);
      for (int li = 0; li < p.lin; ++li, ++step) {
        const bool mine = p.two ? ((step & 1) == me) : true;
        if (mine) {
          if (p.feed) { while (!mbar_try_wait(a_full + as, aph)) ++spins; }
          if (p.fake & 2) { while (!mbar_try_wait(&dummy_done, 0)) ++spins; }
          const uint32_t a_base = a0 + as * a_stage_bytes;
          const uint64_t da_hi = desc0 | (uint64_t)((a_base & 0x3FFFF) >> 4);
          const uint64_t da_lo = desc0 | (uint64_t)(((a_base + 16384) & 0x3FFFF) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              mma_bf16(tmem, da_lo + 2 * ks, db_hi + 2 * ks, idesc, 1u);
              mma_bf16(tmem, da_hi + 2 * ks, db_lo + 2 * ks, idesc, 1u);
              mma_bf16(tmem, da_hi + 2 * ks, db_hi + 2 * ks, idesc, 1u);
            }
            if (p.fake & 1) mma_commit(&dummy_sink);
            if (p.feed) {
              mma_commit(a_empty + as);
              if (li >= p.lin - nw) mma_commit(b_empty + bs);   // the last step(s) of the chunk release the weight stage
            }
          }
          __syncwarp();
        }
        if (++as == (uint32_t)p.a_stages) { as = 0; aph ^= 1; }
      }
      if (++bs == (uint32_t)p.b_stages) { bs = 0; bph ^= 1; }
    }
    long long t1 = clock64();
    if (warp == 1) {
      if (p.two) { asm volatile("bar.sync 3, 64;"); }
      if (elect_one()) mma_commit(&acc_full);
      __syncwarp();
      mbar_wait(&acc_full, 0);
      long long t2 = clock64();
      if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = spins; out[3] = spins_b; }
    } else {
      asm volatile("bar.sync 3, 64;");
    }
  } else
  if (warp == 1) {
    int a_it = 0, b_it = 0;
    const uint64_t desc0 = make_desc_sw128(0);
    const uint32_t idesc = make_idesc(0, 128, p.N);
    long long t0 = clock64(), wa = 0, wb = 0;
    for (int cc = 0; cc < p.chunks; ++cc) {
      const int bs = b_it % p.b_stages;
      if (p.feed) { long long tw = clock64(); mbar_wait(b_full + bs, (b_it / p.b_stages) & 1); wb += clock64() - tw; }
      const uint32_t b_base = smem_u32(b_smem + bs * b_stage_bytes);
      for (int li = 0; li < p.lin; ++li) {
        const int as = a_it % p.a_stages;
        if (p.feed) { long long tw = clock64(); mbar_wait(a_full + as, (a_it / p.a_stages) & 1); wa += clock64() - tw; }
        tc_fence_after();
        const uint32_t a_base = smem_u32(a_smem + as * a_stage_bytes);
        const uint64_t da_hi = desc0 | (uint64_t)((a_base & 0x3FFFF) >> 4);
        const uint64_t da_lo = desc0 | (uint64_t)(((a_base + 16384) & 0x3FFFF) >> 4);
        const uint64_t db_hi = desc0 | (uint64_t)((b_base & 0x3FFFF) >> 4);
        const uint64_t db_lo = desc0 | (uint64_t)(((b_base + b_part) & 0x3FFFF) >> 4);
        const uint32_t acc0 = (cc | li) ? 1u : 0u;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            mma_bf16(tmem, da_lo + 2 * ks, db_hi + 2 * ks, idesc, acc0 | (uint32_t)(ks > 0));
            mma_bf16(tmem, da_hi + 2 * ks, db_lo + 2 * ks, idesc, 1u);
            mma_bf16(tmem, da_hi + 2 * ks, db_hi + 2 * ks, idesc, 1u);
          }
        }
        __syncwarp();
        if (p.feed) { if (elect_one()) mma_commit(a_empty + as); __syncwarp(); }
        ++a_it;
      }
      if (p.feed) { if (elect_one()) mma_commit(b_empty + bs); __syncwarp(); }
      ++b_it;
    }
    long long t1 = clock64();
    if (elect_one()) mma_commit(&acc_full);
    __syncwarp();
    mbar_wait(&acc_full, 0);
    long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = wa; out[3] = wb; }
  } else if (warp >= 6 && warp < 6 + p.spinners) {
    if (p.backoff) {
      while (!mbar_try_wait(&acc_full, 0)) __nanosleep(200);
    } else {
      mbar_wait(&acc_full, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main(int argc, char** argv) {
  const size_t src_bytes = 256u << 20;
  uint8_t* src; cudaMalloc(&src, src_bytes);
  {
    std::vector<__half> h(src_bytes / 2);
    srand(1);
    for (auto& v : h) v = __float2half((float)(rand() % 2001 - 1000) / 500.0f);
    cudaMemcpy(src, h.data(), src_bytes, cudaMemcpyHostToDevice);
  }
  uint8_t* zsrc; cudaMalloc(&zsrc, src_bytes); cudaMemset(zsrc, 0, src_bytes);
  long long* out; cudaMalloc(&out, 64);
  cudaFuncSetAttribute(pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  const int grid = argc > 1 ? atoi(argv[1]) : 64;
  struct Case { const char* name; Params p; int zero; };
  std::vector<Case> cases = {
      {"lean resident N128", {128, 32, 2, 3, 2, 0, 0, 0, 1, 0, 0}, 0},
      {"lean resident N128 +commit", {128, 32, 2, 3, 2, 0, 0, 0, 1, 0, 1}, 0},
      {"lean resident N128 +wait", {128, 32, 2, 3, 2, 0, 0, 0, 1, 0, 2}, 0},
      {"lean resident N128 +commit+wait", {128, 32, 2, 3, 2, 0, 0, 0, 1, 0, 3}, 0},
      {"lean resident N128 +commit+wait two", {128, 32, 2, 3, 2, 0, 0, 0, 1, 1, 3}, 0},
      {"lean feed N128 a3 b2", {128, 32, 2, 3, 2, 0, 1, 0, 1, 0, 0}, 0},
      {"lean feed N128 a3 b2 two warps", {128, 32, 2, 3, 2, 0, 1, 0, 1, 1, 0}, 0},
      {"lean feed N128 a5 b2 two warps", {128, 32, 2, 5, 2, 0, 1, 0, 1, 1, 0}, 0},
      {"lean feed N256 lin4 a2 b2", {256, 32, 4, 2, 2, 0, 1, 0, 1, 0, 0}, 0},
      {"lean feed N256 lin4 a2 b2 two warps", {256, 32, 4, 2, 2, 0, 1, 0, 1, 1, 0}, 0},
      {"lean feed N256 lin2 a2 b2 two warps", {256, 32, 2, 2, 2, 0, 1, 0, 1, 1, 0}, 0},
  };
  for (auto& c : cases) {
    const Params& p = c.p;
    const size_t smem = 1024 + (size_t)p.a_stages * 32768 + (size_t)p.b_stages * 2 * p.N * 128 + 1024;
    if (smem > 225 * 1024) { printf("%s: smem too large\n", c.name); continue; }
    for (int rep = 0; rep < 2; ++rep) pipe_kernel<<<grid, 576, smem>>>(p, c.zero ? zsrc : src, src_bytes, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[4]; cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
    const int steps = p.chunks * p.lin;
    printf("%-44s: %7.0f cyc/step (12 MMAs, ideal %4d)  issue-only %7.0f  spinsA %6.2f spinsB %6.2f per step  %s\n", c.name,
           (double)h[1] / steps, 12 * p.N / 2, (double)h[0] / steps, (double)h[2] / steps, (double)h[3] / steps,
           cudaGetErrorString(e));
  }
  return 0;
}
