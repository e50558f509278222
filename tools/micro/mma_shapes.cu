// Microbenchmark: tcgen05.mma kind::f16 throughput for the operand-access patterns of the conv kernels.
//   * cta_group::1 (M=128) vs cta_group::2 (M=256, B split over the CTA pair) at several N;
//   * the hi/lo split pattern (lo*hi + hi*lo + hi*hi per 32-byte K step, 4 K steps per 128-byte chunk);
//   * operands streamed through several shared-memory stages (distinct addresses, like the real mainloop),
//     optionally with concurrent bulk-copy (TMA) traffic into shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../edmp_b200/csrc -o mma_shapes mma_shapes.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace edmp::umma;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mma_f16_cg2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit_cg2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

struct Params {
  int cg;         // 1 or 2
  int N;          // MMA N (total, both CTAs)
  int steps;      // number of (K chunk) steps; each = 4 K steps x 3 MMAs
  int a_stages;   // A stages (32 KB each: hi + lo blocks)
  int b_stages;   // B stages ((N / cg) x 128 B x 2 parts each)
  int copies;     // background bulk copies of 32 KB (0 = none)
  int split;      // 1: three MMAs per K step, 0: one
};

template <int CG>
__global__ void __launch_bounds__(128) mma_kernel(Params p, const uint8_t* src, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t done, cp_full[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) { mbar_init(&done, 1); mbar_init(cp_full, 1); mbar_init(cp_full + 1, 1); fence_barrier_init(); }
  if (warp == 0) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      tmem_alloc<512>(&tmem_slot);
    }
  }
  const int a_bytes = 32768, b_part = (p.N / p.cg) * 128, b_bytes = 2 * b_part;
  const int op_bytes = p.a_stages * a_bytes + p.b_stages * b_bytes;
  for (int i = threadIdx.x; i < op_bytes / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && rank == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = a0 + p.a_stages * a_bytes;
    const uint64_t d0 = make_desc_sw128(0);
    const uint32_t idesc = make_idesc(0, CG == 2 ? 256 : 128, p.N);
    long long t0 = clock64();
    int sa = 0, sb = 0;
    for (int s = 0; s < p.steps; ++s) {
      const uint32_t a_base = a0 + sa * a_bytes, b_base = b0 + sb * b_bytes;
      if (++sa == p.a_stages) sa = 0;
      if (++sb == p.b_stages) sb = 0;
      const uint64_t da_hi = d0 | ((a_base & 0x3FFFF) >> 4), da_lo = d0 | (((a_base + 16384) & 0x3FFFF) >> 4);
      const uint64_t db_hi = d0 | ((b_base & 0x3FFFF) >> 4), db_lo = d0 | (((b_base + b_part) & 0x3FFFF) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (CG == 2) {
            if (p.split) {
              mma_f16_cg2(tmem, da_lo + 2 * ks, db_hi + 2 * ks, idesc, 1u);
              mma_f16_cg2(tmem, da_hi + 2 * ks, db_lo + 2 * ks, idesc, 1u);
            }
            mma_f16_cg2(tmem, da_hi + 2 * ks, db_hi + 2 * ks, idesc, 1u);
          } else {
            if (p.split) {
              mma_bf16(tmem, da_lo + 2 * ks, db_hi + 2 * ks, idesc, 1u);
              mma_bf16(tmem, da_hi + 2 * ks, db_lo + 2 * ks, idesc, 1u);
            }
            mma_bf16(tmem, da_hi + 2 * ks, db_hi + 2 * ks, idesc, 1u);
          }
        }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (elect_one()) { if (CG == 2) commit_cg2(&done, 3); else mma_commit(&done); }
    __syncwarp();
    mbar_wait(&done, 0);
    long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp == 0 && rank == 1) {
    mbar_wait(&done, 0);   // multicast commit from the leader
  } else if (warp == 1 && lane == 0 && p.copies) {
    uint8_t* dst = smem + ((op_bytes + 1023) & ~1023);
    for (int i = 0; i < p.copies; ++i) {
      int s = i & 1;
      if (i >= 2) mbar_wait(cp_full + s, ((i >> 1) - 1) & 1);
      mbar_arrive_expect_tx(cp_full + s, 32768);
      bulk_g2s(dst + s * 32768, src + ((size_t)(blockIdx.x * 64 + i) * 32768) % (64u << 20), 32768, cp_full + s);
    }
    if (p.copies >= 2) mbar_wait(cp_full + ((p.copies - 2) & 1), (((p.copies - 2) >> 1)) & 1);
    mbar_wait(cp_full + ((p.copies - 1) & 1), (((p.copies - 1) >> 1)) & 1);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    else tmem_dealloc<512>(tmem);
  }
}

int main(int argc, char** argv) {
  uint8_t* src; cudaMalloc(&src, 64u << 20); cudaMemset(src, 0, 64u << 20);
  long long* out; cudaMalloc(&out, 16);
  cudaFuncSetAttribute(mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  cudaFuncSetAttribute(mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  const int grid = argc > 1 ? atoi(argv[1]) : 128;
  for (int cg : {1})
    for (int N : {64, 128, 160, 192, 256})
      for (int split : {1, 0})
        for (int stages : {1, 3})
          for (int copies : {0, 160}) {
            Params p{cg, N, 256, stages, stages == 1 ? 1 : 2, copies, split};
            const int b_bytes = 2 * (N / cg) * 128;
            const size_t smem = 1024 + (size_t)p.a_stages * 32768 + (size_t)p.b_stages * b_bytes + 1024 + (copies ? 65536 : 0);
            if (smem > 225 * 1024) continue;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cg; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = cg == 2 ? 1 : 0;
            cudaError_t e = cudaSuccess;
            for (int rep = 0; rep < 2; ++rep) e = cg == 2 ? cudaLaunchKernelEx(&cfg, mma_kernel<2>, p, (const uint8_t*)src, out) : cudaLaunchKernelEx(&cfg, mma_kernel<1>, p, (const uint8_t*)src, out);
            cudaError_t e2 = cudaDeviceSynchronize();
            long long h[2] = {0, 0}; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            const int n_mma = p.steps * 4 * (split ? 3 : 1);
            const double cyc = (double)h[1] / n_mma;
            const double macs = (double)(cg == 2 ? 256 : 128) * N * 16 / cyc / cg;   // per SM
            printf("cg%d N=%3d split=%d stages=%d copies=%2d: issue %6.1f  complete %6.1f cyc/MMA  -> %6.0f MAC/cyc/SM (%4.1f%% of 4096)  %s %s\n",
                   cg, N, split, stages, copies, (double)h[0] / n_mma, cyc, macs, 100.0 * macs / 4096.0,
                   cudaGetErrorString(e), cudaGetErrorString(e2));
          }
  return 0;
}
