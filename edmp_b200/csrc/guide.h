// Scene / guide state shared between guide.cu, sampler.cu and api.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

#include "common.cuh"

namespace edmp {

struct SceneDev {
  int n_obs;
  float R[kMaxObs][9];      // obstacle rotation (float32 like the reference's obstacle_transform)
  float C[kMaxObs][3];      // obstacle centre
  double dims[kMaxObs][3];  // obstacle extents (float64 until expansion/clearance are applied)
  float link_half[9][3];    // link box half extents
};

struct Scene {
  SceneDev host;
  SceneDev* dev = nullptr;
  // per-row guide tables (infer_serial.py:59-91)
  double *clearance = nullptr, *expansion = nullptr, *schedule = nullptr;
  unsigned char *method = nullptr, *grad_norm = nullptr;
  int rows = 0, ensemble_rows = 0;
  // work buffers
  float* raw = nullptr;
  size_t raw_cap = 0;
  double* rowsq = nullptr;
  size_t rowsq_cap = 0;
  double* zero_tables = nullptr;
  int tail_grid_cap = 0;    // CTAs of step_tail_kernel that are resident at once (its grid-wide barrier needs all of them)
};

int guide_init_constants();
int scene_create(const double* cfg, int n_obs, const double* link_dims, Scene** out);
void scene_destroy(Scene* s);
int scene_set_tables(Scene* s, const double* clr, const double* exp_, const double* sched,
                     const double* method, const double* gnorm, int rows, int ensemble_rows);
// x: float64 [rows,7,ld]; interior waypoint w lives at column off+w.  Either grad_out (standalone
// get_gradient) or x_state/xf_state (guided update of the sampler state, in place) is given.
int guide_gradient_launch(Scene* s, const double* x, int ld, int off, int n_inner, bool clip,
                          const double* start_h, const double* goal_h, int t, int rows,
                          double* grad_out, float* raw_out, double* x_state, float* xf_state,
                          cudaStream_t st);
// posterior + (guided steps) gradient, norm mix, guided update, endpoints, float32 copy: one launch per reverse step
int guide_step_tail_launch(Scene* s, double* x, float* xf, const float* eps, const double* noise, uint64_t seed, int t,
                           double c1, double sqrt_alpha, double beta, const double* start_h, const double* goal_h,
                           int rows, bool guided, bool condition, unsigned* bar, unsigned* bar_epoch, void* pack_hi,
                           void* pack_lo, int pack_el, unsigned* range_flag, cudaStream_t st);
int guide_volumes_launch(Scene* s, const float* q, const double* start_h, const double* goal_h, int t,
                         int mode, int rows, int n, float* vol, cudaStream_t st);
int guide_final_cost_launch(Scene* s, const double* traj, const double* start_h, const double* goal_h,
                            int rows, float* cost, cudaStream_t st);

}  // namespace edmp
