// Microbenchmark: tcgen05.mma issue/execute rate (cta_group::1, M=128, operands in shared memory,
// SWIZZLE_128B K-major) for kind::tf32 (K=8) and kind::f16/bf16 (K=16) at several N, optionally
// with concurrent bulk-copy traffic into shared memory.
#include <cstdio>
#include <vector>
#include "umma.cuh"
using namespace edmp::umma;

__global__ void __launch_bounds__(128) mma_kernel(int fmt, int N, int n_mma, int with_copy, const uint8_t* src,
                                                  long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t done, cp_full[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&done, 1); mbar_init(cp_full, 1); mbar_init(cp_full + 1, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  // zero operands (values do not matter for timing)
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 16384);
    const uint64_t d0 = make_desc_sw128(0);
    const uint32_t idesc = make_idesc(fmt, 128, N);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 4) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t da = (d0 | ((a_base & 0x3FFFF) >> 4)) + 2 * ks, db = (d0 | ((b_base & 0x3FFFF) >> 4)) + 2 * ks;
          if (fmt == 2) mma_tf32(tmem, da, db, idesc, 1u); else mma_bf16(tmem, da, db, idesc, 1u);
        }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (elect_one()) mma_commit(&done);
    __syncwarp();
    mbar_wait(&done, 0);
    long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp == 1 && lane == 0 && with_copy) {
    // background bulk copies into a separate smem region while the MMAs run
    for (int i = 0; i < with_copy; ++i) {
      int s = i & 1;
      if (i >= 2) mbar_wait(cp_full + s, ((i >> 1) - 1) & 1);
      mbar_arrive_expect_tx(cp_full + s, 32768);
      bulk_g2s(smem + 65536 + s * 32768, src + ((size_t)(blockIdx.x * 64 + i) * 32768) % (64u << 20), 32768, cp_full + s);
    }
    mbar_wait(cp_full + 0, ((with_copy - 1 - ((with_copy - 1) & 1)) >> 1) & 1);
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main() {
  uint8_t* src; cudaMalloc(&src, 64u << 20); cudaMemset(src, 0, 64u << 20);
  long long* out; cudaMalloc(&out, 16);
  cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int Ns[] = {16, 64, 80, 128, 256};
  for (int fmt : {2, 1})
    for (int N : Ns)
      for (int wc : {0, 48}) {
        const int n = 2048;
        for (int rep = 0; rep < 2; ++rep) mma_kernel<<<64, 128, 180 * 1024>>>(fmt, N, n, wc, src, out);
        cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("%s M=128 N=%3d K=%2d  copies=%2d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA  (%s)\n", fmt == 2 ? "tf32" : "bf16", N,
               fmt == 2 ? 8 : 16, wc, (double)h[0] / n, (double)h[1] / n, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
