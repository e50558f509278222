/*
 * edmp_b200.h -- C ABI of the B200-native guided-diffusion trajectory sampler.
 *
 * Drop-in boundary for the hot path of vishal-2000/EDMP (`infer_serial.py`):
 *   Diffusion.denoise_guided            reference diffusion/diffusion.py:300-356
 *   TemporalUNet.forward                reference diffusion/models/temporalunet.py:47-76
 *   IntersectionVolumeGuide.{cost, swept_volume_cost, get_gradient, choose_best_trajectory}
 *                                       reference lib/guide.py:354-395, :473-537, :597-635, :637-653
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; edmp_last_error() gives the text.
 *     No C++ exceptions cross this boundary.
 *   - pointers suffixed _h are HOST pointers, _d are DEVICE pointers (caller owned, on the device
 *     that was current when the handle was created).  `stream` is a cudaStream_t passed as void*.
 *   - handles are opaque, one per device, not thread safe (the reference is single threaded and
 *     keeps per-scene state in the guide object, lib/guide.py:155-156).
 *   - trajectories are laid out exactly like the reference: [rows, 7, 50] (channels first,
 *     diffusion/diffusion.py:303).  Sampler state is float64 like the reference's numpy state;
 *     network input/output is float32 (diffusion.py:319-324).
 *   - "ensemble" = the rows of one reference denoise_guided call (n_guides x batch_size_per_guide,
 *     infer_serial.py:56-58).  A batch may hold several ensembles back to back; the whole-batch
 *     gradient norm (lib/guide.py:629) and the t==1 "row 0 gets no noise" quirk
 *     (diffusion.py:127) are applied per ensemble.
 */
#ifndef EDMP_B200_H
#define EDMP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDMP_T_STEPS 255   /* diffusion steps the tables are sized for (benchmark/cfgs/cfg1.yaml:15) */
#define EDMP_HORIZON 50    /* waypoints (cfg1.yaml:16)            */
#define EDMP_DOF 7         /* channels  (cfg1.yaml:17)            */
#define EDMP_MAX_OBSTACLES 64

typedef struct edmp_unet edmp_unet;       /* packed TemporalUNet weights + activation workspace */
typedef struct edmp_scene edmp_scene;     /* obstacle set + link boxes + per-row guide tables  */
typedef struct edmp_sampler edmp_sampler; /* schedule + per-batch state for denoise_guided      */

/* arithmetic used for the convolution contractions */
enum {
  EDMP_PRECISION_FP32 = 0,    /* fp32 FMA on CUDA cores (parity mode, always available)          */
  EDMP_PRECISION_TF32X3 = 1,  /* tcgen05 kind::tf32 with hi/lo operand split (3 MMAs, ~fp32)     */
  EDMP_PRECISION_TF32 = 2,    /* tcgen05 kind::tf32 single pass (fast, ~1e-3 relative)           */
  EDMP_PRECISION_BF16X3 = 3,  /* tcgen05 kind::f16 bf16 hi/lo split (3 MMAs, ~2^-16 relative)    */
  EDMP_PRECISION_BF16 = 4,    /* tcgen05 kind::f16 bf16 single pass (fastest, ~1e-2 relative)    */
  EDMP_PRECISION_F16X3 = 5,   /* tcgen05 kind::f16 IEEE-half hi/lo split with power-of-two weight
                                 scaling (3 MMAs at twice the tf32 rate, ~fp32; parity grade)     */
  EDMP_PRECISION_F16 = 6      /* tcgen05 kind::f16 IEEE-half single pass (~1e-3 relative)          */
};

const char* edmp_last_error(void);
int edmp_version(void);
/* number of float32 parameters a TemporalUNet(input_dim=7, time_dim=32, dims) state_dict holds */
size_t edmp_unet_param_count(const int* dims, int n_dims);

/* ---- TemporalUNet (replaces temporalunet.py:11-45 ctor/load and :47-76 forward) ------------- */
/* params_h: every tensor of the reference state_dict, flattened and concatenated in state_dict
 * order (temporalunet.py:78-92 checkpoint layout; 290 tensors / 29,938,471 floats for the
 * shipped dims (32,64,128,256,512,512)).  max_rows sizes the activation workspace. */
int edmp_unet_create(const float* params_h, size_t n_params, const int* dims, int n_dims,
                     int precision, int max_rows, edmp_unet** out);
void edmp_unet_destroy(edmp_unet* u);
/* ---- packed-weight blob (the checkpoint side is temporalunet.py:78-92 save / load) -------
 * edmp_unet_pack = edmp_unet_create that also RECORDS every device image it uploads (repacked hi / lo weight tiles in
 * their UMMA layouts, per-channel vectors, the 255-row time-embedding table) as one host blob; edmp_unet_blob_bytes /
 * edmp_unet_blob_read hand it out (the read releases the engine's host copy); edmp_unet_create_from_blob builds an
 * engine from such a blob with no state_dict and no repacking.  The blob carries the layout version
 * (edmp_unet_blob_layout_version), the precision, dims, and per item the tile geometry it was packed for: a stale or
 * mismatching blob (other layout version, other batch-size class / kernel generation) is refused with rc 4 and the
 * caller repacks from the state_dict.  The host class keeps the blobs as a versioned on-disk cache next to the
 * checkpoint (edmp_b200/diffusion/temporalunet.py). */
int edmp_unet_pack(const float* params_h, size_t n_params, const int* dims, int n_dims, int precision, int max_rows,
                   edmp_unet** out);
size_t edmp_unet_blob_bytes(const edmp_unet* u);
int edmp_unet_blob_read(edmp_unet* u, void* dst_h, size_t capacity);
int edmp_unet_create_from_blob(const void* blob_h, size_t nbytes, int max_rows, edmp_unet** out);
int edmp_unet_blob_layout_version(void);
/* eps[rows,7,50] = model(x[rows,7,50], t), t in 1..255 (one t for the whole batch,
 * diffusion.py:320).  x_d/eps_d float32 device pointers. */
int edmp_unet_forward(edmp_unet* u, const float* x_d, int t, int rows, float* eps_d, void* stream);
/* debug / per-layer parity: copies the activation named like the reference module path
 * (e.g. "down_samplers.2.down.1", "up_samplers.0.up.3", "final_conv.0") of the LAST forward into
 * out_d as [rows, C, L] float32; returns C and L. */
int edmp_unet_read_activation(edmp_unet* u, const char* name, int rows, float* out_d, int* C, int* L,
                              void* stream);
/* measurement: per-kernel device time of one forward (CUDA events on `stream` around every launch,
 * the median over iters forwards) and the non-padding multiply-accumulates each kernel performs.
 * ms_h / macs_h hold edmp_unet_launches_per_forward() entries; edmp_unet_op_name(u, i) names op i
 * after the reference module it implements. */
int edmp_unet_profile(edmp_unet* u, const float* x_d, int t, int rows, int iters, float* ms_h,
                      double* macs_h, float* eps_d, void* stream);
const char* edmp_unet_op_name(const edmp_unet* u, int i);
/* which kernel runs op i: "conv_tc2" (persistent tcgen05 kernel, single CTAs), "conv_tc2_pair" (cta_group::2 CTA
 * pairs), "conv_tc", "conv_pm", "conv_simt", "pack", "final_pw" */
const char* edmp_unet_op_kernel(const edmp_unet* u, int i);
/* debug: SM-clock stamps / counters ([ctas][16], see tools/tc_trace.py) of tensor-core op `op` run alone */
int edmp_unet_tc_trace(edmp_unet* u, int op, int rows, long long* out_h, int max_ctas, int* n_ctas,
                       void* stream);
int edmp_unet_precision(const edmp_unet* u);
/* Operand-range check of the IEEE-half modes (f16x3, f16): activations are split hi + lo un-scaled, so a value beyond
 * +-65504 would silently turn the forward into NaN.  Every epilogue raises a device flag instead; this call
 * synchronises `stream`, reads and clears the flag: *overflow = 1 when a forward since the last call left the range
 * (re-create the engine with tf32x3 / bf16x3 / fp32).  Always 0 for the other modes. */
int edmp_unet_range_status(edmp_unet* u, int* overflow, void* stream);
/* number of kernels one forward launches (for bench.py's gpu_launches claim) */
int edmp_unet_launches_per_forward(const edmp_unet* u);

/* ---- scene + guide (replaces lib/guide.py:13-43 ctor, :118-158 define_obstacles) ------------ */
/* obstacle_cfg_h [n_obs,10] = (xyz, quaternion xyzw, dims) float64 (load_test_dataset.py:189);
 * link_dims_h [9,3] float64 link box extents (lib/guide.py:243-281), NULL = built-in table. */
int edmp_scene_create(const double* obstacle_cfg_h, int n_obs, const double* link_dims_h,
                      edmp_scene** out);
void edmp_scene_destroy(edmp_scene* s);
/* Per-row guide tables as infer_serial.py:59-91 builds them (float64 host arrays):
 * clearance/expansion/schedule [rows,255], method/grad_norm [rows] (0/1).  ensemble_rows divides
 * rows. */
int edmp_scene_set_guide_tables(edmp_scene* s, const double* clearance_h, const double* expansion_h,
                                const double* schedule_h, const double* method_h,
                                const double* grad_norm_h, int rows, int ensemble_rows);
/* get_gradient (lib/guide.py:597-635): q_d [rows,7,48] float64 (already clipped by the caller, as
 * diffusion.py:328 does), start_h/goal_h [7] float64, t in 1..255.  grad_d [rows,7,48] float64
 * receives (1-g)*G + g*G/||G||_F(ensemble).  raw_d (optional, may be NULL) receives the float32
 * autograd-equivalent G before mixing. */
int edmp_guide_gradient(edmp_scene* s, const double* q_d, const double* start_h, const double* goal_h,
                        int t, int rows, double* grad_d, float* raw_d, void* stream);
/* cost (:354-395, mode 0) / swept_volume_cost (:473-537, mode 1): q_d [rows,7,n] float32, returns
 * volumes_d [rows, n (iv) | n+1 (sv), 9*n_obs] float32 with index link*n_obs+obs.  t == 0 means no
 * expansion/clearance; with t == 0 the guide tables need not be set (goal filter,
 * infer_serial.py:119). */
int edmp_guide_volumes(edmp_scene* s, const float* q_d, const double* start_h, const double* goal_h,
                       int t, int mode, int rows, int n, float* volumes_d, void* stream);
/* choose_best_trajectory (:637-653): traj_d [rows,7,50] float64 -> cost_d[rows] float32 =
 * sum of swept volumes at t=0. */
int edmp_guide_final_cost(edmp_scene* s, const double* traj_d, const double* start_h,
                          const double* goal_h, int rows, float* cost_d, void* stream);

/* ---- sampler (replaces diffusion.py:10-20 schedule, :116-135 posterior, :300-356 loop) ------- */
int edmp_sampler_create(int T, double variance_thresh, int max_rows, edmp_sampler** out);
void edmp_sampler_destroy(edmp_sampler* s);
/* Runs reverse steps t = t_start, t_start-1, ..., t_stop+1 of denoise_guided on x_d
 * [rows,7,50] float64 in place (full run: t_start=255, t_stop=0; a single teacher-forced step t:
 * t_start=t, t_stop=t-1).  Endpoints are conditioned on start/goal before the first step
 * (diffusion.py:306-307) and after every step (:347-349).
 *   noise_d : float64 [t_start-t_stop, rows, 7, 50], noise_d[k] is z for step t_start-k
 *             (the reference's np.random draws, diffusion.py:126), or NULL to draw N(0,1) on the
 *             device from Philox4x32-10 keyed by `seed`.
 *   scene   : NULL disables guidance (Diffusion.denoise, diffusion.py:253-278).
 *   final_cost_d : optional [rows] float32, filled with the t=0 swept volume when t_stop == 0.
 * Guide tables must have been set for exactly `rows` rows. */
int edmp_sample_guided(edmp_sampler* s, edmp_unet* u, edmp_scene* scene, double* x_d,
                       const double* start_h, const double* goal_h, const double* noise_d,
                       uint64_t seed, int rows, int t_start, int t_stop, float* final_cost_d,
                       void* stream);
/* Same, end to end from HOST buffers (the call bench.py's e2e leg times): copies x_T in, runs all
 * steps with device-side Philox noise, copies trajectories [rows,7,50] float64 and costs [rows]
 * float32 back.  x_h/cost_h should be pinned for full-speed copies. */
int edmp_sample_guided_host(edmp_sampler* s, edmp_unet* u, edmp_scene* scene, double* x_h,
                            const double* start_h, const double* goal_h, uint64_t seed, int rows,
                            float* cost_h, void* stream);
/* schedule read-back (beta, alpha, alpha_bar, each [T] float64) for host-side checks */
int edmp_sampler_schedule(const edmp_sampler* s, double* beta_h, double* alpha_h, double* alpha_bar_h);
/* kernels launched by the last edmp_sample_guided[_host] call */
long long edmp_sampler_last_launches(const edmp_sampler* s);
/* condition != 0 (the default): the first / last waypoint of every row is overwritten with start / goal before the
 * first step and after every step (diffusion.py:306-307,:347-349 `if condition:`); 0: the endpoints diffuse freely
 * (start / goal are then only the padding waypoints of the guide's swept volumes, lib/guide.py:484-492). */
int edmp_sampler_set_condition(edmp_sampler* s, int condition);

/* ---- sphere / signed-distance guide family (SURVEY.md section 8 a-S; BASELINE.json configs[4]) ---------------
 * Not on the reference's infer_serial.py path (its guide is the AABB-volume one above); specified by code the
 * reference vendors: robofin/robofin/robots.py:58-174 (59 collision spheres), the URDF chain
 * robofin/robofin/urdf/franka_panda/panda.urdf:47-235 (FK), robofin/robofin/pointcloud/torch.py:340-365
 * (compute_spheres), mpinets/geometry.py:238-288,:456-505 (cuboid / cylinder SDF), mpinets/loss.py:88-94 (hinge).
 * boxes_h [n_boxes,10] = (xyz, quaternion xyzw, dims) like obstacle_config; cylinders_h [n_cylinders,9] =
 * (xyz, quaternion xyzw, radius, height); at most 64 primitives. */
typedef struct edmp_sdf_scene edmp_sdf_scene;
int edmp_sdf_scene_create(const double* boxes_h, int n_boxes, const double* cylinders_h, int n_cylinders,
                          edmp_sdf_scene** out);
void edmp_sdf_scene_destroy(edmp_sdf_scene* s);
/* q_d float32 [rows,7,n] (n <= 64 waypoints).  cost_d [rows] = sum over (waypoint, sphere) of
 * max(0, radius + margin - sdf(centre)); grad_d [rows,7,n] = d cost / d q (analytic); clearance_d [rows,n] =
 * min over spheres of sdf(centre) - radius.  Any output may be null. */
int edmp_sdf_guide(edmp_sdf_scene* s, const float* q_d, int n, int rows, float margin, float* cost_d, float* grad_d,
                   float* clearance_d, void* stream);
/* point-cloud variant: points_d float32 [n_points,4] (xyz + pad); clearance_d [rows,n] = min over (sphere, point) of
 * |centre - p| - radius (n <= 50).  No reference implementation exists for this one (SURVEY.md section 8 a-S). */
int edmp_sdf_cloud_clearance(const float* q_d, int n, int rows, const float* points_d, int n_points,
                             float* clearance_d, void* stream);

/* ---- trajectory metrics for whole ensembles ----------------------------------------
 * Replaces the per-trajectory host code of the reference's lib/metrics.py (MetricsCalculator).  All pointers are
 * device pointers. */
/* nfft = 2^(ceil(log2 m) + padlevel) of an m-sample profile (lib/metrics.py:90); 0 if out of range */
int edmp_metrics_nfft(int m, int padlevel);
/* lib/guide.py:100-116 get_end_effector_transform: q_d float32 [rows,n,7] -> T_d float32 [rows,n,4,4] = product of
 * the 10 DH matrices (get_tf_mat :45-72, static table :29-38) */
int edmp_ee_transform(const float* q_d, int rows, int n, float* T_d, void* stream);
/* lib/metrics.py:11-45: traj_d float64 [rows,7,n] (n <= 64) -> out_d float64 [rows,4] = joint path length,
 * end-effector path length, joint SPARC, end-effector SPARC (speed profiles norm(diff / dt), fs = 1 / dt).
 * spectrum_d (may be null) float64 [rows,2,nfft]: normalised magnitude spectra Mf of the two profiles;
 * selected_d (may be null) int [rows,2,2]: first / last bin of the arc (-1, -1: empty selection). */
int edmp_trajectory_metrics(const double* traj_d, int rows, int n, double dt, int padlevel, double fc, double amp_th,
                            double* out_d, double* spectrum_d, int* selected_d, void* stream);
/* lib/metrics.py:47-130 sparc(movement, fs, padlevel, fc, amp_th) for a batch of raw profiles: movement_d float64
 * [rows,m] -> sal_d [rows]; spectrum_d [rows,nfft] and selected_d [rows,2] may be null. */
int edmp_sparc(const double* movement_d, int rows, int m, double fs, int padlevel, double fc, double amp_th,
               double* sal_d, double* spectrum_d, int* selected_d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EDMP_B200_H */
