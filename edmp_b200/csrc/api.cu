// extern "C" boundary of libedmp_b200.so -- see include/edmp_b200.h for the contract.
#include "edmp_b200.h"

#include <cstdlib>
#include <mutex>
#include <string>

#include "common.cuh"
#include "guide.h"
#include "sampler.h"
#include "sdf_guide.h"
#include "metrics.h"
#include "unet.h"

namespace edmp {
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
bool pdl_enabled() {
  // Programmatic dependent launch along the per-step kernel chain: +7 % at 1020 rows, +0.4 % at 8190 rows with the
  // persistent kernels (profiles/r1_f16x3_pdl_ab.txt); EDMP_NO_PDL=1 falls back to plain stream order.
  static const bool on = std::getenv("EDMP_NO_PDL") == nullptr;
  return on;
}
static std::mutex g_once_mutex;
bool once_per_device(unsigned long long* mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;   // unknown device: just redo the setup
  std::lock_guard<std::mutex> lock(g_once_mutex);
  const unsigned long long bit = 1ull << dev;
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}
void once_per_device_failed(unsigned long long* mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return;
  std::lock_guard<std::mutex> lock(g_once_mutex);
  *mask &= ~(1ull << dev);
}
bool cluster_pdl_enabled() {
  static const bool on = std::getenv("EDMP_NO_PAIR_PDL") == nullptr;
  return on;
}
}  // namespace edmp

using namespace edmp;

struct edmp_unet { UNet* impl; };
struct edmp_scene { Scene* impl; };
struct edmp_sampler { Sampler* impl; };
struct edmp_sdf_scene { SdfScene* impl; };

#define EDMP_TRY(expr)                                                     \
  try {                                                                    \
    return (expr);                                                         \
  } catch (const std::exception& e) {                                      \
    set_error(std::string("edmp: exception: ") + e.what());                \
    return 3;                                                              \
  } catch (...) {                                                          \
    set_error("edmp: unknown exception");                                  \
    return 3;                                                              \
  }

extern "C" {

const char* edmp_last_error(void) { return g_error.c_str(); }
int edmp_version(void) { return 100; }

size_t edmp_unet_param_count(const int* dims, int n_dims) {
  try { return unet_param_count(dims, n_dims); } catch (...) { return 0; }
}

int edmp_unet_create(const float* params_h, size_t n_params, const int* dims, int n_dims, int precision,
                     int max_rows, edmp_unet** out) {
  if (!params_h || !dims || !out) { set_error("edmp_unet_create: null argument"); return 2; }
  try {
    UNet* u = nullptr;
    int rc = unet_create(params_h, n_params, dims, n_dims, precision, max_rows, &u);
    if (rc) return rc;
    *out = new edmp_unet{u};
    return 0;
  } catch (const std::exception& e) { set_error(std::string("edmp: exception: ") + e.what()); return 3; }
}
void edmp_unet_destroy(edmp_unet* u) { if (u) { unet_destroy(u->impl); delete u; } }
int edmp_unet_pack(const float* params_h, size_t n_params, const int* dims, int n_dims, int precision, int max_rows,
                   edmp_unet** out) {
  if (!params_h || !dims || !out) { set_error("edmp_unet_pack: null argument"); return 2; }
  try {
    UNet* u = nullptr;
    int rc = unet_pack(params_h, n_params, dims, n_dims, precision, max_rows, &u);
    if (rc) return rc;
    *out = new edmp_unet{u};
    return 0;
  } catch (const std::exception& e) { set_error(std::string("edmp: exception: ") + e.what()); return 3; }
}
size_t edmp_unet_blob_bytes(const edmp_unet* u) { return u ? unet_blob_bytes(u->impl) : 0; }
int edmp_unet_blob_read(edmp_unet* u, void* dst_h, size_t capacity) {
  if (!u || !dst_h) { set_error("edmp_unet_blob_read: null argument"); return 2; }
  EDMP_TRY(unet_blob_read(u->impl, dst_h, capacity));
}
int edmp_unet_create_from_blob(const void* blob_h, size_t nbytes, int max_rows, edmp_unet** out) {
  if (!blob_h || !out) { set_error("edmp_unet_create_from_blob: null argument"); return 2; }
  try {
    UNet* u = nullptr;
    int rc = unet_create_from_blob(blob_h, nbytes, max_rows, &u);
    if (rc) return rc;
    *out = new edmp_unet{u};
    return 0;
  } catch (const std::exception& e) { set_error(std::string("edmp: exception: ") + e.what()); return 3; }
}
int edmp_unet_blob_layout_version(void) { return unet_blob_layout_version(); }
int edmp_unet_forward(edmp_unet* u, const float* x_d, int t, int rows, float* eps_d, void* stream) {
  if (!u || !x_d || !eps_d) { set_error("edmp_unet_forward: null argument"); return 2; }
  EDMP_TRY(unet_forward(u->impl, x_d, t, rows, eps_d, (cudaStream_t)stream));
}
int edmp_unet_read_activation(edmp_unet* u, const char* name, int rows, float* out_d, int* C, int* L,
                              void* stream) {
  if (!u || !name) { set_error("edmp_unet_read_activation: null argument"); return 2; }
  EDMP_TRY(unet_read_activation(u->impl, name, rows, out_d, C, L, (cudaStream_t)stream));
}
int edmp_unet_profile(edmp_unet* u, const float* x_d, int t, int rows, int iters, float* ms_h, double* macs_h,
                      float* eps_d, void* stream) {
  if (!u || !x_d || !ms_h || !macs_h || !eps_d) { set_error("edmp_unet_profile: null argument"); return 2; }
  EDMP_TRY(unet_profile(u->impl, x_d, t, rows, iters, ms_h, macs_h, eps_d, (cudaStream_t)stream));
}
int edmp_unet_tc_trace(edmp_unet* u, int op, int rows, long long* out_h, int max_ctas, int* n_ctas, void* stream) {
  if (!u || !out_h || !n_ctas) { set_error("edmp_unet_tc_trace: null argument"); return 2; }
  EDMP_TRY(unet_tc_trace(u->impl, op, rows, out_h, max_ctas, n_ctas, (cudaStream_t)stream));
}
const char* edmp_unet_op_name(const edmp_unet* u, int i) { return u ? unet_op_name(u->impl, i) : nullptr; }
const char* edmp_unet_op_kernel(const edmp_unet* u, int i) { return u ? unet_op_kernel(u->impl, i) : nullptr; }
int edmp_unet_precision(const edmp_unet* u) { return u ? unet_precision(u->impl) : -1; }
int edmp_unet_range_status(edmp_unet* u, int* overflow, void* stream) {
  if (!u || !overflow) { set_error("edmp_unet_range_status: null argument"); return 2; }
  EDMP_TRY(unet_range_status(u->impl, overflow, static_cast<cudaStream_t>(stream)));
}
int edmp_unet_launches_per_forward(const edmp_unet* u) { return u ? unet_launches(u->impl) : 0; }

int edmp_scene_create(const double* obstacle_cfg_h, int n_obs, const double* link_dims_h, edmp_scene** out) {
  if (!obstacle_cfg_h || !out) { set_error("edmp_scene_create: null argument"); return 2; }
  try {
    Scene* s = nullptr;
    int rc = scene_create(obstacle_cfg_h, n_obs, link_dims_h, &s);
    if (rc) return rc;
    *out = new edmp_scene{s};
    return 0;
  } catch (const std::exception& e) { set_error(std::string("edmp: exception: ") + e.what()); return 3; }
}
void edmp_scene_destroy(edmp_scene* s) { if (s) { scene_destroy(s->impl); delete s; } }
int edmp_scene_set_guide_tables(edmp_scene* s, const double* clearance_h, const double* expansion_h,
                                const double* schedule_h, const double* method_h, const double* grad_norm_h,
                                int rows, int ensemble_rows) {
  if (!s || !clearance_h || !expansion_h || !schedule_h || !method_h || !grad_norm_h) {
    set_error("edmp_scene_set_guide_tables: null argument");
    return 2;
  }
  EDMP_TRY(scene_set_tables(s->impl, clearance_h, expansion_h, schedule_h, method_h, grad_norm_h, rows,
                            ensemble_rows));
}
int edmp_guide_gradient(edmp_scene* s, const double* q_d, const double* start_h, const double* goal_h, int t,
                        int rows, double* grad_d, float* raw_d, void* stream) {
  if (!s || !q_d || !start_h || !goal_h || !grad_d) { set_error("edmp_guide_gradient: null argument"); return 2; }
  // the caller has already clipped (diffusion.py:328), so no clipping here
  EDMP_TRY(guide_gradient_launch(s->impl, q_d, kHorizon - 2, 0, kHorizon - 2, false, start_h, goal_h, t, rows,
                                 grad_d, raw_d, nullptr, nullptr, (cudaStream_t)stream));
}
int edmp_guide_volumes(edmp_scene* s, const float* q_d, const double* start_h, const double* goal_h, int t,
                       int mode, int rows, int n, float* volumes_d, void* stream) {
  if (!s || !q_d || !volumes_d) { set_error("edmp_guide_volumes: null argument"); return 2; }
  if (mode && (!start_h || !goal_h)) { set_error("edmp_guide_volumes: sv mode needs start/goal"); return 2; }
  EDMP_TRY(guide_volumes_launch(s->impl, q_d, start_h, goal_h, t, mode, rows, n, volumes_d,
                                (cudaStream_t)stream));
}
int edmp_guide_final_cost(edmp_scene* s, const double* traj_d, const double* start_h, const double* goal_h,
                          int rows, float* cost_d, void* stream) {
  if (!s || !traj_d || !start_h || !goal_h || !cost_d) { set_error("edmp_guide_final_cost: null argument"); return 2; }
  EDMP_TRY(guide_final_cost_launch(s->impl, traj_d, start_h, goal_h, rows, cost_d, (cudaStream_t)stream));
}

int edmp_sampler_create(int T, double variance_thresh, int max_rows, edmp_sampler** out) {
  if (!out) { set_error("edmp_sampler_create: null argument"); return 2; }
  try {
    Sampler* s = nullptr;
    int rc = sampler_create(T, variance_thresh, max_rows, &s);
    if (rc) return rc;
    *out = new edmp_sampler{s};
    return 0;
  } catch (const std::exception& e) { set_error(std::string("edmp: exception: ") + e.what()); return 3; }
}
void edmp_sampler_destroy(edmp_sampler* s) { if (s) { sampler_destroy(s->impl); delete s; } }
int edmp_sample_guided(edmp_sampler* s, edmp_unet* u, edmp_scene* scene, double* x_d, const double* start_h,
                       const double* goal_h, const double* noise_d, uint64_t seed, int rows, int t_start,
                       int t_stop, float* final_cost_d, void* stream) {
  if (!s || !u || !x_d) { set_error("edmp_sample_guided: null argument"); return 2; }
  EDMP_TRY(sample_guided(s->impl, u->impl, scene ? scene->impl : nullptr, x_d, start_h, goal_h, noise_d, seed,
                         rows, t_start, t_stop, final_cost_d, (cudaStream_t)stream));
}
int edmp_sample_guided_host(edmp_sampler* s, edmp_unet* u, edmp_scene* scene, double* x_h,
                            const double* start_h, const double* goal_h, uint64_t seed, int rows, float* cost_h,
                            void* stream) {
  if (!s || !u || !x_h) { set_error("edmp_sample_guided_host: null argument"); return 2; }
  EDMP_TRY(sample_guided_host(s->impl, u->impl, scene ? scene->impl : nullptr, x_h, start_h, goal_h, seed, rows,
                              cost_h, (cudaStream_t)stream));
}
int edmp_sampler_schedule(const edmp_sampler* s, double* beta_h, double* alpha_h, double* alpha_bar_h) {
  if (!s || !beta_h || !alpha_h || !alpha_bar_h) { set_error("edmp_sampler_schedule: null argument"); return 2; }
  return sampler_schedule(s->impl, beta_h, alpha_h, alpha_bar_h);
}
long long edmp_sampler_last_launches(const edmp_sampler* s) { return s ? sampler_last_launches(s->impl) : 0; }
int edmp_sampler_set_condition(edmp_sampler* s, int condition) {
  if (!s) { set_error("edmp_sampler_set_condition: null argument"); return 2; }
  sampler_set_condition(s->impl, condition != 0);
  return 0;
}


/* ---- sphere / signed-distance guide family (SURVEY.md section 8 a-S) ---------------------------------- */
int edmp_sdf_scene_create(const double* boxes_h, int n_boxes, const double* cylinders_h, int n_cylinders,
                          edmp_sdf_scene** out) {
  if (!out) { set_error("edmp_sdf_scene_create: null argument"); return 2; }
  try {
    SdfScene* s = nullptr;
    int rc = sdf_scene_create(boxes_h, n_boxes, cylinders_h, n_cylinders, &s);
    if (rc) return rc;
    *out = new edmp_sdf_scene{s};
    return 0;
  } catch (const std::exception& e) { set_error(std::string("edmp: exception: ") + e.what()); return 3; }
}
void edmp_sdf_scene_destroy(edmp_sdf_scene* s) { if (s) { sdf_scene_destroy(s->impl); delete s; } }
int edmp_sdf_guide(edmp_sdf_scene* s, const float* q_d, int n, int rows, float margin, float* cost_d, float* grad_d,
                   float* clearance_d, void* stream) {
  if (!s || !q_d) { set_error("edmp_sdf_guide: null argument"); return 2; }
  EDMP_TRY(sdf_guide_launch(s->impl, q_d, n, rows, margin, cost_d, grad_d, clearance_d, (cudaStream_t)stream));
}
int edmp_sdf_cloud_clearance(const float* q_d, int n, int rows, const float* points_d, int n_points,
                             float* clearance_d, void* stream) {
  if (!q_d || !points_d || !clearance_d) { set_error("edmp_sdf_cloud_clearance: null argument"); return 2; }
  EDMP_TRY(sdf_cloud_launch(q_d, n, rows, points_d, n_points, clearance_d, (cudaStream_t)stream));
}

/* ---- trajectory metrics ----------------------------------------------------- */
int edmp_metrics_nfft(int m, int padlevel) { return metrics_nfft(m, padlevel); }
int edmp_ee_transform(const float* q_d, int rows, int n, float* T_d, void* stream) {
  if (!q_d || !T_d) { set_error("edmp_ee_transform: null argument"); return 2; }
  EDMP_TRY(ee_transform_launch(q_d, rows, n, T_d, (cudaStream_t)stream));
}
int edmp_trajectory_metrics(const double* traj_d, int rows, int n, double dt, int padlevel, double fc, double amp_th,
                            double* out_d, double* spectrum_d, int* selected_d, void* stream) {
  if (!traj_d || !out_d) { set_error("edmp_trajectory_metrics: null argument"); return 2; }
  EDMP_TRY(trajectory_metrics_launch(traj_d, rows, n, dt, padlevel, fc, amp_th, out_d, spectrum_d, selected_d,
                                     (cudaStream_t)stream));
}
int edmp_sparc(const double* movement_d, int rows, int m, double fs, int padlevel, double fc, double amp_th,
               double* sal_d, double* spectrum_d, int* selected_d, void* stream) {
  if (!movement_d || !sal_d) { set_error("edmp_sparc: null argument"); return 2; }
  EDMP_TRY(sparc_launch(movement_d, rows, m, fs, padlevel, fc, amp_th, sal_d, spectrum_d, selected_d,
                        (cudaStream_t)stream));
}
}  // extern "C"
