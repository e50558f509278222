// Guided reverse-diffusion loop: schedule, posterior update, endpoint conditioning, guidance cadence.
//
// Replaces reference diffusion/diffusion.py: schedule :10-20,:37-49; p_sample_using_posterior
// :116-135; clip_joints :280-298 (inside guide.cu's loader); denoise_guided :300-356.  The state is
// float64 on the device like the reference's numpy state; the network sees a float32 copy
// (diffusion.py:319).  Nothing returns to the host between steps.
#include "common.cuh"
#include "philox.cuh"
#include "guide.h"
#include "sampler.h"
#include "unet.h"

#include <cmath>
#include <cstdlib>
#include <vector>

namespace edmp {

struct Sampler {
  int T = 0;
  int max_rows = 0;
  std::vector<double> beta, alpha, alpha_bar;
  float* xf = nullptr;    // float32 copy of the state fed to the UNet
  float* eps = nullptr;   // UNet output
  double* x_dev = nullptr;  // staging for the host-buffer entry point
  float* cost_dev = nullptr;
  unsigned* bar = nullptr;  // grid-barrier counter of the fused per-step tail kernel (guide.cu)
  long long last_launches = 0;
  bool condition = true;    // overwrite the first / last waypoint with start / goal every step (diffusion.py:306-307,:347-349)
};

struct StepCoef {
  double c1;          // (1 - alpha) / sqrt(1 - alpha_bar)
  double sqrt_alpha;  // sqrt(alpha)
  double beta;
  double start[7], goal[7];
};

// X[:, :, 0] = start, X[:, :, -1] = goal (diffusion.py:306-307), and the float32 copy
__global__ void condition_kernel(double* __restrict__ x, float* __restrict__ xf, StepCoef sc, bool condition, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int e = (int)(i % kRowElems);
  const int j = e / kHorizon, l = e % kHorizon;
  double v = x[i];
  if (condition && l == 0) v = sc.start[j];
  if (condition && l == kHorizon - 1) v = sc.goal[j];
  x[i] = v;
  xf[i] = (float)v;
}

// x_{t-1} = (x_t - c1 * eps) / sqrt(alpha) + beta * z   (diffusion.py:133: beta, not sqrt(beta));
// at t == 1 row 0 of every ensemble gets z = 0 and the other rows keep their noise (:127 under
// numpy-1.x semantics); endpoints re-conditioned (:347-349).
__global__ void posterior_kernel(double* __restrict__ x, float* __restrict__ xf,
                                 const float* __restrict__ eps, const double* __restrict__ noise,
                                 uint64_t seed, int t, int ensemble_rows, StepCoef sc, bool condition, size_t n) {
  pdl_launch_dependents();
  pdl_wait();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int row = (int)(i / kRowElems);
  const int e = (int)(i % kRowElems);
  const int j = e / kHorizon, l = e % kHorizon;
  double z = noise ? noise[i] : philox_normal(seed, (uint32_t)t, i);
  if (t == 1 && (row % ensemble_rows) == 0) z = 0.0;
  double v = (x[i] - sc.c1 * (double)eps[i]) / sc.sqrt_alpha + sc.beta * z;
  if (condition && l == 0) v = sc.start[j];
  if (condition && l == kHorizon - 1) v = sc.goal[j];
  x[i] = v;
  xf[i] = (float)v;
}

int sampler_create(int T, double thresh, int max_rows, Sampler** out) {
  EDMP_REQUIRE(T == kTSteps, "only T=255 is supported (guide tables are [rows,255])");
  EDMP_REQUIRE(max_rows > 0, "max_rows must be positive");
  Sampler* s = new Sampler();
  s->T = T;
  s->max_rows = max_rows;
  s->beta.resize(T); s->alpha.resize(T); s->alpha_bar.resize(T);
  // np.linspace(0, thresh, T+1)[1:]  (diffusion.py:47): start + i*step with step = thresh/T
  const double step = (thresh - 0.0) / T;
  double prod = 1.0;
  for (int i = 0; i < T; ++i) {
    s->beta[i] = (i + 1 == T) ? thresh : 0.0 + (i + 1) * step;
    s->alpha[i] = 1 - s->beta[i];
    prod *= s->alpha[i];           // np.prod(alpha[:t]) multiplies left to right
    s->alpha_bar[i] = prod;
  }
  size_t n = (size_t)max_rows * kRowElems;
  if (cudaMalloc(&s->xf, n * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&s->eps, n * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&s->x_dev, n * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&s->cost_dev, max_rows * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&s->bar, sizeof(unsigned)) != cudaSuccess) {
    set_error("sampler_create: device allocation failed");
    sampler_destroy(s);
    return 1;
  }
  *out = s;
  return 0;
}

void sampler_destroy(Sampler* s) {
  if (!s) return;
  cudaFree(s->xf); cudaFree(s->eps); cudaFree(s->x_dev); cudaFree(s->cost_dev); cudaFree(s->bar);
  delete s;
}

int sampler_schedule(const Sampler* s, double* b, double* a, double* ab) {
  for (int i = 0; i < s->T; ++i) { b[i] = s->beta[i]; a[i] = s->alpha[i]; ab[i] = s->alpha_bar[i]; }
  return 0;
}

long long sampler_last_launches(const Sampler* s) { return s->last_launches; }
void sampler_set_condition(Sampler* s, bool condition) { s->condition = condition; }

int sample_guided(Sampler* s, UNet* u, Scene* scene, double* x, const double* start, const double* goal,
                  const double* noise, uint64_t seed, int rows, int t_start, int t_stop, float* final_cost,
                  cudaStream_t st) {
  EDMP_REQUIRE(rows > 0 && rows <= s->max_rows, "rows exceeds the sampler's max_rows");
  EDMP_REQUIRE(t_start <= s->T && t_stop >= 0 && t_start > t_stop, "bad step range");
  EDMP_REQUIRE(start && goal, "start/goal are required (the guide's swept volumes use them also when condition=False)");
  if (scene) EDMP_REQUIRE(scene->rows == rows, "guide tables were set for a different row count");
  const int ens = scene ? scene->ensemble_rows : rows;
  const size_t n = (size_t)rows * kRowElems;
  const int threads = 256;
  const int blocks = (int)((n + threads - 1) / threads);
  StepCoef sc;
  for (int j = 0; j < 7; ++j) { sc.start[j] = start[j]; sc.goal[j] = goal[j]; }
  sc.c1 = sc.sqrt_alpha = sc.beta = 0.0;
  long long launches = 0;
  char label[96];
  snprintf(label, sizeof(label), "edmp_sample_guided rows=%d t=%d..%d%s", rows, t_start, t_stop + 1, scene ? " guided" : "");
  NvtxRange pass_range(label);
  // one fused launch per step after the UNet (posterior, gradient, norm mix, guided update); EDMP_SPLIT_TAIL=1 keeps the
  // three separate kernels (A/B, and the unguided path always uses posterior_kernel)
  static const bool split_tail = getenv("EDMP_SPLIT_TAIL") != nullptr;
  const bool fused = scene != nullptr && !split_tail;
  unsigned bar_epoch = 0;
  if (fused) EDMP_CK(cudaMemsetAsync(s->bar, 0, sizeof(unsigned), st));
  // the fused tail also writes the network's input image of the next step (no separate pack launch from step 2 on)
  void *pack_hi = nullptr, *pack_lo = nullptr;
  int pack_el = 0;
  unsigned* range_flag = nullptr;
  const bool fold = fused && unet_input_image(u, &pack_hi, &pack_lo, &pack_el, &range_flag);
  bool packed = false;   // the image holds the current state (written by the previous step's tail)
  condition_kernel<<<blocks, threads, 0, st>>>(x, s->xf, sc, s->condition, n);
  ++launches;
  for (int t = t_start; t > t_stop; --t) {
    snprintf(label, sizeof(label), "step t=%d%s", t, (scene && (t % 2) == 0 && t >= 5) ? " (guided)" : "");
    NvtxRange step_range(label);
    if (packed ? unet_forward_packed(u, t, rows, s->eps, st) : unet_forward(u, s->xf, t, rows, s->eps, st)) return 1;
    launches += unet_launches(u) - (packed ? 1 : 0);
    const double a = s->alpha[t - 1], ab = s->alpha_bar[t - 1];
    sc.c1 = (1 - a) / std::sqrt(1 - ab);
    sc.sqrt_alpha = std::sqrt(a);
    sc.beta = s->beta[t - 1];
    const double* z = noise ? noise + (size_t)(t_start - t) * n : nullptr;
    if (fused) {
      // guidance cadence: (t % 2) < 1 and t >= 5  (diffusion.py:326-327)
      if (guide_step_tail_launch(scene, x, s->xf, s->eps, z, seed, t, sc.c1, sc.sqrt_alpha, sc.beta, start, goal, rows,
                                 (t % 2) == 0 && t >= 5, s->condition, s->bar, &bar_epoch, fold ? pack_hi : nullptr,
                                 pack_lo, pack_el, range_flag, st))
        return 1;
      ++launches;
      packed = fold;
      continue;
    }
    launch_pdl(posterior_kernel, dim3(blocks), dim3(threads), 0, st, x, s->xf, (const float*)s->eps, z, seed, t, ens, sc, s->condition, n);
    ++launches;
    // guidance cadence: (t % 2) < 1 and t >= 5  (diffusion.py:326-327)
    if (scene && (t % 2) == 0 && t >= 5) {
      if (guide_gradient_launch(scene, x, kHorizon, 1, kHorizon - 2, /*clip=*/true, start, goal, t, rows,
                                nullptr, nullptr, x, s->xf, st))
        return 1;
      launches += 2;
    }
  }
  if (final_cost && t_stop == 0 && scene) {
    if (guide_final_cost_launch(scene, x, start, goal, rows, final_cost, st)) return 1;
    ++launches;
  }
  EDMP_CK(cudaGetLastError());
  s->last_launches = launches;
  return 0;
}

int sample_guided_host(Sampler* s, UNet* u, Scene* scene, double* x_h, const double* start,
                       const double* goal, uint64_t seed, int rows, float* cost_h, cudaStream_t st) {
  EDMP_REQUIRE(rows > 0 && rows <= s->max_rows, "rows exceeds the sampler's max_rows");
  const size_t nb = (size_t)rows * kRowElems * sizeof(double);
  EDMP_CK(cudaMemcpyAsync(s->x_dev, x_h, nb, cudaMemcpyHostToDevice, st));
  if (sample_guided(s, u, scene, s->x_dev, start, goal, nullptr, seed, rows, s->T, 0,
                    (cost_h && scene) ? s->cost_dev : nullptr, st))
    return 1;
  EDMP_CK(cudaMemcpyAsync(x_h, s->x_dev, nb, cudaMemcpyDeviceToHost, st));
  if (cost_h && scene)
    EDMP_CK(cudaMemcpyAsync(cost_h, s->cost_dev, rows * sizeof(float), cudaMemcpyDeviceToHost, st));
  EDMP_CK(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace edmp
