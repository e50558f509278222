// Microbenchmark: per-SM ingest rate of 1-D bulk async copies (global/L2 -> shared) as a function
// of copy size, copies in flight and number of CTAs.  nvcc -arch=sm_100a -O3 -I../../edmp_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "umma.cuh"
using namespace edmp::umma;

__global__ void bulk_kernel(const uint8_t* src, size_t src_bytes, int copy_bytes, int stages, int n_copies,
                            int shared_mode, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_base[];
  uint8_t* smem = smem_base;
  __shared__ uint64_t full_all[64];
  const int issuer = threadIdx.x >> 5;              // one issuing thread per warp
  const int n_issuers = blockDim.x >> 5;
  uint64_t* full = full_all + issuer * 16;
  if ((threadIdx.x & 31) == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(full + i, 1);
    fence_barrier_init();
  }
  __syncthreads();
  smem += (size_t)issuer * stages * copy_bytes;
  src_bytes /= n_issuers;
  src += (size_t)issuer * src_bytes;
  if ((threadIdx.x & 31) == 0) {
    // shared_mode 0: every CTA streams its own region; 1: groups of 8 CTAs read the same addresses
    size_t base = (size_t)(shared_mode ? blockIdx.x / 8 : blockIdx.x) * (size_t)n_copies * copy_bytes;
    long long t0 = clock64();
    for (int i = 0; i < n_copies + stages; ++i) {
      int s = i % stages;
      if (i >= stages) mbar_wait(full + s, ((i / stages) - 1) & 1);
      if (i < n_copies) {
        mbar_arrive_expect_tx(full + s, copy_bytes);
        bulk_g2s(smem + (size_t)s * copy_bytes, src + (base + (size_t)i * copy_bytes) % src_bytes, copy_bytes, full + s);
      }
    }
    if (issuer == 0) cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  size_t src_bytes = 96u << 20;
  uint8_t* src;
  cudaMalloc(&src, src_bytes);
  cudaMemset(src, 1, src_bytes);
  long long* cyc;
  cudaMalloc(&cyc, 1024 * sizeof(long long));
  cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int grids[] = {64, 148};
  int sizes[] = {16384, 32768, 65536, 98304};
  int stages_l[] = {1, 2};
  int issuers_l[] = {1, 2, 4};
  for (int shared_mode = 0; shared_mode < 1; ++shared_mode)
    for (int g : grids)
      for (int sz : sizes)
        for (int st : stages_l)
         for (int iss : issuers_l) {
          if ((size_t)sz * st * iss > 190 * 1024) continue;
          int n = (1 << 20) / sz * 2;   // 2 MB per issuer
          for (int rep = 0; rep < 2; ++rep)
            bulk_kernel<<<g, 32 * iss, (size_t)sz * st * iss, 0>>>(src, src_bytes, sz, st, n, shared_mode, cyc);
          cudaDeviceSynchronize();
          std::vector<long long> h(g);
          cudaMemcpy(h.data(), cyc, g * sizeof(long long), cudaMemcpyDeviceToHost);
          double avg = 0;
          for (auto v : h) avg += v;
          avg /= g;
          printf("ctas=%3d copy=%6d B stages=%d issuers=%d: %6.1f B/cycle/SM  (%5.2f TB/s aggregate @1.9GHz)  %7.0f cycles per copy-slot\n",
                 g, sz, st, iss, (double)n * sz * iss / avg, (double)n * sz * iss / avg * g * 1.9e9 / 1e12, avg / n);
        }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
