// Shared helpers for the edmp_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: a no-op unless a profiler (nsys / ncu --nvtx) is attached

namespace edmp {

void set_error(const std::string& msg);

// NVTX range for the lifetime of the object: the sampler marks every pass and every reverse step, the UNet every
// forward (SURVEY.md section 5 "tracing"); `ncu --nvtx --nvtx-include "step t=254/"` then profiles one step.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

#define EDMP_CK(call)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::edmp::set_error(std::string(#call) + " failed: " + cudaGetErrorString(_e) + " (" +  \
                        __FILE__ + ":" + std::to_string(__LINE__) + ")");                   \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

#define EDMP_REQUIRE(cond, msg)                                              \
  do {                                                                       \
    if (!(cond)) {                                                           \
      ::edmp::set_error(std::string("edmp: ") + (msg) + " [" #cond "]");     \
      return 2;                                                              \
    }                                                                        \
  } while (0)

constexpr int kHorizon = 50;
constexpr int kDof = 7;
constexpr int kRowElems = kHorizon * kDof;  // 350
constexpr int kTSteps = 255;
constexpr int kMaxObs = 64;

// Programmatic dependent launch: every kernel of the per-step chain lets its successor start
// launching at once (the successor's prologue -- barrier init, TMEM allocation, parameter staging,
// launch latency -- overlaps this kernel's execution on the idle SMs) and waits for its predecessor
// to complete before touching activations.  Both are no-ops for a normal launch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// One-time device setup (constant tables, kernel attributes) is PER DEVICE: a process that builds engines on cuda:0 and
// then on cuda:1 must repeat it there.  once_per_device(mask) returns true the first time it is called with the current
// device for this flag word (thread-safe; devices 0..63).
bool once_per_device(unsigned long long* mask);
void once_per_device_failed(unsigned long long* mask);   // the setup did not complete: try again next time

bool pdl_enabled();
bool cluster_pdl_enabled();   // EDMP_NO_PAIR_PDL=1 switches it off (A/B)

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// launch with a thread-block cluster of `cluster_x` CTAs along x (CTA pairs of the cta_group::2 kernels)
template <class... KArgs, class... Args>
inline cudaError_t launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                                  Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  // programmatic dependent launch for the cluster kernels as well: without it a CTA-pair layer starts only after the
  // previous kernel has drained grid-wide, and the row-tile chaining of consecutive layers never comes into play
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = (pdl_enabled() && cluster_pdl_enabled()) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// x * tanh(softplus(x)) (reference blocks.py:27 nn.Mish), via tanh(log1p(e^x)) = n/(n+2), n = e^x(e^x+2).
__device__ __forceinline__ float mish_f(float x) {
  if (x > 20.0f) return x;
  float e = expf(x);
  float n = e * (e + 2.0f);
  return x * (n / (n + 2.0f));
}

}  // namespace edmp
