// Microbenchmark: raw tcgen05.mma issue rate of one thread (cta_group::1, M=128, small N so that the tensor pipe is
// never the limit): identical descriptors vs descriptors advancing by constants, elect-guarded vs single-thread warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../edmp_b200/csrc -o mma_issue mma_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace edmp::umma;

template <int MODE>
__global__ void __launch_bounds__(128) issue_kernel(int N, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 32768;
    const uint64_t d0 = make_desc_sw128(0);
    const uint64_t da = d0 | ((a_base & 0x3FFFF) >> 4), db = d0 | ((b_base & 0x3FFFF) >> 4);
    const uint32_t idesc = make_idesc(0, 128, N);
    long long t0 = clock64();
    if (MODE == 0) {          // identical MMAs, elect-guarded, 16 per loop iteration
      for (int i = 0; i < iters; ++i) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 16; ++k) mma_bf16(tmem, da, db, idesc, 1u);
        }
        __syncwarp();
      }
    } else if (MODE == 1) {   // descriptors advance by constants
      for (int i = 0; i < iters; ++i) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 16; ++k) mma_bf16(tmem, da + 2 * (k & 3) + 1024 * (k >> 2), db + 2 * (k & 3), idesc, 1u);
        }
        __syncwarp();
      }
    } else if (MODE == 2) {   // only lane 0 runs the loop (divergent single thread)
      if (lane == 0) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
          for (int k = 0; k < 16; ++k) mma_bf16(tmem, da + 2 * (k & 3) + 1024 * (k >> 2), db + 2 * (k & 3), idesc, 1u);
        }
      }
      __syncwarp();
    } else {                  // hi/lo pattern: 3 MMAs per K step, different D per group of 12
      for (int i = 0; i < iters; ++i) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            mma_bf16(tmem + (i & 1) * 64, da + 1024 + 2 * k, db + 2 * k, idesc, 1u);
            mma_bf16(tmem + (i & 1) * 64, da + 2 * k, db + 512 + 2 * k, idesc, 1u);
            mma_bf16(tmem + (i & 1) * 64, da + 2 * k, db + 2 * k, idesc, 1u);
          }
        }
        __syncwarp();
      }
    }
    long long t1 = clock64();
    if (elect_one()) mma_commit(&done);
    __syncwarp();
    mbar_wait(&done, 0);
    long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int MODE> void run(const char* name, int per_iter, long long* out) {
  cudaFuncSetAttribute(issue_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int N : {16, 32, 64, 128}) {
    const int iters = 512;
    for (int rep = 0; rep < 2; ++rep) issue_kernel<MODE><<<32, 128, 70 * 1024>>>(N, iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-34s N=%3d: issue %6.1f  complete %6.1f cyc/MMA (tensor floor %3d)  %s\n", name, N, (double)h[0] / (iters * per_iter),
           (double)h[1] / (iters * per_iter), N / 2, cudaGetErrorString(e));
  }
}

int main() {
  long long* out; cudaMalloc(&out, 16);
  run<0>("identical, elect", 16, out);
  run<1>("advancing desc, elect", 16, out);
  run<2>("advancing desc, lane0 divergent", 16, out);
  run<3>("hi/lo x3 pattern, elect", 12, out);
  return 0;
}
