// Microbenchmark: what the L2 -> shared-memory path delivers to 148 persistent CTAs that stream operand blocks with 1-D
// bulk async copies the way conv_tc2's producer does, and whether cluster multicast raises it.
//   * "A" stream: 16 KB blocks, private to a CTA (its own region of an L2-resident buffer, re-walked several times);
//   * "W" stream: 24 KB blocks that EVERY CTA reads from the same small region at (nearly) the same time, like the weight
//     tiles of a column tile;
//   * stages = blocks in flight per stream; the consumer frees a stage the moment it is full (no MMAs: pure feed rate).
// Variants: cluster size 1 / 2 / 4 with the W block split over the CTAs of a cluster and multicast to all of them
// (every CTA issues 1/cs of the block with a cs-wide destination mask) against plain unicast of the whole block.
// Also prints cudaOccupancyMaxActiveClusters for the cluster sizes (how many SMs a cluster-of-4 launch can use).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../edmp_b200/csrc l2_feed.cu -o l2_feed
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace edmp::umma;

constexpr int kABlock = 16384, kWBlock = 24576;
constexpr int kMaxStages = 8;

__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ bool test_cta(uint64_t* bar, uint32_t parity) {   // non-blocking (try_wait may suspend the thread)
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// non-blocking phase test (acquire at cluster scope): a blocking wait here could deadlock two CTAs that each wait for the
// other to free a block of the OTHER stream
__device__ __forceinline__ bool test_cl(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// mode 0: unicast; 1: W multicast over the cluster; 2: A AND W multicast (all CTAs of a cluster stream the same A blocks)
__global__ void feed_kernel(const uint8_t* a_src, size_t a_region, const uint8_t* w_src, size_t w_region, int a_stages, int w_stages,
                            int n_a, int n_w, int cs, int mode, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t a_full[kMaxStages], w_full[kMaxStages], a_free[kMaxStages], w_free[kMaxStages];
  const uint32_t rank = cs > 1 ? ctarank() : 0u;
  uint8_t* a_smem = smem;
  uint8_t* w_smem = smem + (size_t)a_stages * kABlock;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxStages; ++i) { mbar_init(a_full + i, 1); mbar_init(w_full + i, 1); mbar_init(a_free + i, cs); mbar_init(w_free + i, cs); }
    fence_barrier_init();
  }
  __syncthreads();
  if (cs > 1) cluster_sync();
  const bool mc_w = mode >= 1 && cs > 1, mc_a = mode >= 2 && cs > 1;
  const uint16_t mask = (uint16_t)((1u << cs) - 1);
  // private A region per CTA (per cluster when A is multicast); one shared W region
  const size_t a_base = (size_t)(mc_a ? blockIdx.x / cs : blockIdx.x) * a_region;
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    int ia = 0, iw = 0, ca = 0, cw = 0;           // issued / consumed counts
    while (ca < n_a || cw < n_w) {
      // issue whatever has a free stage (a stage written by multicast peers needs every CTA of the cluster to have freed it)
      if (ia < n_a && ia - ca < a_stages) {
        const int s = ia % a_stages;
        const bool ok = !(mc_a && ia >= a_stages) || test_cl(a_free + s, ((ia / a_stages) - 1) & 1);
        if (ok) {
          mbar_arrive_expect_tx(a_full + s, kABlock);
          const uint8_t* src = a_src + a_base + ((size_t)ia * kABlock) % a_region;
          if (mc_a) bulk_g2s_mc(a_smem + (size_t)s * kABlock + rank * (kABlock / cs), src + rank * (kABlock / cs), kABlock / cs, a_full + s, mask);
          else bulk_g2s(a_smem + (size_t)s * kABlock, src, kABlock, a_full + s);
          ++ia;
        }
      }
      if (iw < n_w && iw - cw < w_stages) {
        const int s = iw % w_stages;
        if (!(mc_w && iw >= w_stages) || test_cl(w_free + s, ((iw / w_stages) - 1) & 1)) {
          mbar_arrive_expect_tx(w_full + s, kWBlock);
          const uint8_t* src = w_src + ((size_t)iw * kWBlock) % w_region;
          if (mc_w) bulk_g2s_mc(w_smem + (size_t)s * kWBlock + rank * (kWBlock / cs), src + rank * (kWBlock / cs), kWBlock / cs, w_full + s, mask);
          else bulk_g2s(w_smem + (size_t)s * kWBlock, src, kWBlock, w_full + s);
          ++iw;
        }
      }
      // consume: the oldest outstanding block of either stream, if it has landed
      if (ca < ia && test_cta(a_full + ca % a_stages, (ca / a_stages) & 1)) {
        if (mc_a) for (int r2 = 0; r2 < cs; ++r2) arrive_remote(a_free + ca % a_stages, r2);
        ++ca;
      }
      if (cw < iw && test_cta(w_full + cw % w_stages, (cw / w_stages) & 1)) {
        if (mc_w) for (int r2 = 0; r2 < cs; ++r2) arrive_remote(w_free + cw % w_stages, r2);
        ++cw;
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (cs > 1) cluster_sync();
}

static void run(const uint8_t* a_src, size_t a_region, const uint8_t* w_src, size_t w_region, int a_st, int w_st, int n_a, int n_w,
                int cs, int mode, long long* cyc, const char* label, int grid_req = 148) {
  const int grid = grid_req / cs * cs;
  const size_t smem = (size_t)a_st * kABlock + (size_t)w_st * kWBlock;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  float best = 1e30f;
  double avg_c = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, feed_kernel, a_src, a_region, w_src, w_region, a_st, w_st, n_a, n_w, cs, mode, cyc);
    cudaEventRecord(e1);
    if (err != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("%s: %s\n", label, cudaGetErrorString(cudaGetLastError())); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) {
      best = ms;
      std::vector<long long> h(grid);
      cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      avg_c = 0; for (auto v : h) avg_c += v; avg_c /= grid;
    }
  }
  const double bytes = (double)n_a * kABlock + (double)n_w * kWBlock;   // delivered into EACH CTA's shared memory
  printf("%-44s ctas %3d  A stages %d  W stages %d : %6.1f B/clk/SM delivered (%5.2f TB/s over the chip, kernel %7.1f us, %8.0f cycles)\n",
         label, grid, a_st, w_st, bytes / avg_c, bytes * grid / (best * 1e-3) / 1e12, best * 1e3, avg_c);
}

// Independent issuers: issuer k streams every (n_issuers/2)-th block of the A (k even) or W (k odd) stream through its OWN
// ring of `stages` blocks.  same_warp = 1: the issuers are lanes 0.. of warp 0; 0: lane 0 of warps 0..
__global__ void feed_multi_kernel(const uint8_t* a_src, size_t a_region, const uint8_t* w_src, size_t w_region, int stages, int n_a, int n_w,
                                  int n_issuers, int same_warp, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[16][kMaxStages];
  const int k = same_warp ? (int)threadIdx.x : ((threadIdx.x & 31) == 0 ? (int)(threadIdx.x >> 5) : -1);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) for (int j = 0; j < kMaxStages; ++j) mbar_init(&full[i][j], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (k >= 0 && k < n_issuers) {
    const bool is_w = k & 1;
    const int sub = k >> 1, n_sub = n_issuers >> 1;
    const int blk = is_w ? kWBlock : kABlock;
    const int n = ((is_w ? n_w : n_a) - sub + n_sub - 1) / n_sub;          // my blocks: sub, sub + n_sub, ...
    uint8_t* ring = smem + (size_t)k * stages * kWBlock;
    const uint8_t* src = is_w ? w_src : a_src + (size_t)blockIdx.x * a_region;
    const size_t region = is_w ? w_region : a_region;
    const long long t0 = clock64();
    int issued = 0, done = 0;
    while (done < n) {
      if (issued < n && issued - done < stages) {
        const int s = issued % stages;
        mbar_arrive_expect_tx(&full[k][s], blk);
        bulk_g2s(ring + (size_t)s * blk, src + ((size_t)(sub + issued * n_sub) * blk) % region, blk, &full[k][s]);
        ++issued;
      }
      if (done < issued && test_cta(&full[k][done % stages], (done / stages) & 1)) ++done;
    }
    if (k == 0) cycles[blockIdx.x] = clock64() - t0;
    if (k == 1) cycles[512 + blockIdx.x] = clock64() - t0;
  }
}

static void run_multi(const uint8_t* a_src, size_t a_region, const uint8_t* w_src, size_t w_region, int stages, int n_a, int n_w, int n_issuers,
                      int same_warp, long long* cyc, const char* label) {
  const int grid = 148;
  const size_t smem = (size_t)n_issuers * stages * kWBlock;
  float best = 1e30f;
  double avg_c = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    feed_multi_kernel<<<grid, same_warp ? 32 : 32 * n_issuers, smem>>>(a_src, a_region, w_src, w_region, stages, n_a, n_w, n_issuers, same_warp, cyc);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: %s\n", label, cudaGetErrorString(cudaGetLastError())); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) {
      best = ms;
      std::vector<long long> h(1024);
      cudaMemcpy(h.data(), cyc, 1024 * sizeof(long long), cudaMemcpyDeviceToHost);
      avg_c = 0; for (int i = 0; i < grid; ++i) avg_c += (double)(h[i] > h[512 + i] ? h[i] : h[512 + i]); avg_c /= grid;
    }
  }
  const double bytes = (double)n_a * kABlock + (double)n_w * kWBlock;
  printf("%-44s ctas %3d  %d issuers x %d stages     : %6.1f B/clk/SM delivered (%5.2f TB/s over the chip, kernel %7.1f us, %8.0f cycles)\n",
         label, grid, n_issuers, stages, bytes / avg_c, bytes * grid / (best * 1e-3) / 1e12, best * 1e3, avg_c);
}

int main() {
  const size_t a_region = 448 << 10;                // 148 x 448 KB = 65 MB of private activation blocks: L2-resident after the first walk
  const size_t w_region = 3 << 20;                  // one layer's weights
  uint8_t *a_src, *w_src;
  cudaMalloc(&a_src, a_region * 148); cudaMalloc(&w_src, w_region);
  cudaMemset(a_src, 1, a_region * 148); cudaMemset(w_src, 2, w_region);
  long long* cyc; cudaMalloc(&cyc, 1024 * sizeof(long long));
  cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 / cs * cs); cfg.blockDim = dim3(608); cfg.dynamicSmemBytes = 216 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, feed_kernel, &cfg);
    printf("cluster size %d, 216 KB shared memory per CTA: max active clusters %d = %d SMs (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  // the pair kernel's per-chunk traffic: 4 activation blocks (2 positions x hi/lo) + 2 weight blocks (hi/lo halves) = 112 KB
  const int n_a = 4 * 160, n_w = 2 * 160;
  run(a_src, a_region, w_src, w_region, 6, 4, n_a, n_w, 1, 0, cyc, "unicast  (A 96 KB + W 96 KB in flight)");
  run(a_src, a_region, w_src, w_region, 4, 2, n_a, n_w, 1, 0, cyc, "unicast  (A 64 KB + W 48 KB in flight)");
  run(a_src, a_region, w_src, w_region, 2, 2, n_a, n_w, 1, 0, cyc, "unicast  (A 32 KB + W 48 KB in flight)");
  run(a_src, a_region, w_src, w_region, 6, 4, n_a, n_w, 1, 0, cyc, "unicast  74 CTAs", 74);
  run(a_src, a_region, w_src, w_region, 6, 4, n_a, n_w, 1, 0, cyc, "unicast  37 CTAs", 37);
  run(a_src, a_region, w_src, w_region, 1, 1, n_a, n_w, 1, 0, cyc, "unicast  (A 16 KB + W 24 KB in flight)");
  run(a_src, a_region, w_src, w_region, 6, 4, n_a, 0, 1, 0, cyc, "unicast  A only (96 KB in flight)");
  run(a_src, a_region, w_src, w_region, 6, 4, 0, n_w, 1, 0, cyc, "unicast  W only (96 KB in flight)");
  for (int cs : {2, 4}) {
    char l[96];
    snprintf(l, sizeof l, "cluster %d unicast", cs);
    run(a_src, a_region, w_src, w_region, 6, 4, n_a, n_w, cs, 0, cyc, l);
    snprintf(l, sizeof l, "cluster %d W multicast", cs);
    run(a_src, a_region, w_src, w_region, 6, 4, n_a, n_w, cs, 1, cyc, l);
    snprintf(l, sizeof l, "cluster %d A + W multicast", cs);
    run(a_src, a_region, w_src, w_region, 6, 4, n_a, n_w, cs, 2, cyc, l);
  }
  cudaFuncSetAttribute(feed_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  run_multi(a_src, a_region, w_src, w_region, 3, n_a, n_w, 2, 0, cyc, "2 issuers in 2 warps (A | W)");
  run_multi(a_src, a_region, w_src, w_region, 3, n_a, n_w, 2, 1, cyc, "2 issuers, lanes of ONE warp (A | W)");
  run_multi(a_src, a_region, w_src, w_region, 2, n_a, n_w, 4, 0, cyc, "4 issuers in 4 warps (A A | W W)");
  run_multi(a_src, a_region, w_src, w_region, 2, n_a, n_w, 4, 1, cyc, "4 issuers, lanes of ONE warp");
  run_multi(a_src, a_region, w_src, w_region, 1, n_a, n_w, 8, 0, cyc, "8 issuers in 8 warps, 1 stage each");
  run_multi(a_src, a_region, w_src, w_region, 1, n_a, n_w, 8, 1, cyc, "8 issuers, lanes of ONE warp, 1 stage each");
  run_multi(a_src, a_region, w_src, w_region, 4, n_a, n_w, 2, 0, cyc, "2 issuers in 2 warps, 4 stages each");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
