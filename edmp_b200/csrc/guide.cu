// AABB intersection-volume / swept-volume guide: cost and closed-form gradient kernels.
//
// Replaces reference lib/guide.py: forward_kinematics :74-98 (modified-DH chain), link boxes
// :243-375, define_obstacles :118-158, cost :354-395, swept_volume_cost :473-537, get_gradient
// :597-635 (autograd there, analytic here -- SURVEY.md section 8 a-G), choose_best_trajectory
// :637-653.  One CTA per trajectory row, one thread per waypoint (or segment); the scene's
// obstacle boxes are staged once per CTA into shared memory.
#include "common.cuh"
#include "conv_pm.cuh"   // layout of the network's input image (pm_act_off, pm_img_bytes, kPmRows) and the hi / lo split
#include "guide.h"
#include "philox.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace edmp {

// ---- constant geometry ---------------------------------------------------------------------
// modified-DH rows (a, d, cos(alpha), sin(alpha)) evaluated in float32 like the reference does
// (lib/guide.py:29-38 + get_tf_mat :45-72; cos(float32(pi/2)) is -4.37e-8, not 0).
__constant__ float c_dh[7][4];
// link-box centre frames relative to the driving joint frame, 3x4 row major (lib/guide.py:289-340)
__constant__ float c_frame[9][12];
__constant__ double c_joint_lo[7];
__constant__ double c_joint_hi[7];

static const double kDhHost[7][3] = {  // a, d, alpha
    {0, 0.333, 0},          {0, 0, -M_PI / 2},    {0, 0.316, M_PI / 2}, {0.0825, 0, M_PI / 2},
    {-0.0825, 0.384, -M_PI / 2}, {0, 0, M_PI / 2}, {0.088, 0, M_PI / 2}};
static const double kFrameT[9][3] = {{8.71e-05, -3.709035e-02, -6.851545e-02},
                                     {-8.425e-05, -6.93950016e-02, 3.71961970e-02},
                                     {0.0414576, 0.0281429, -0.03293086},
                                     {-4.12337575e-02, 3.44296512e-02, 2.79226985e-02},
                                     {3.3450000e-05, 3.7388050e-02, -1.0619285e-01},
                                     {4.21935000e-02, 1.52195003e-02, 6.07699933e-03},
                                     {1.86357500e-02, 1.85788569e-02, 7.94137484e-02},
                                     {-1.26717073e-03, -1.25294673e-03, 1.27018693e-01},
                                     {9.29352476e-03, 9.28272434e-03, 1.92390375e-01}};
// Built-in link box extents: AABBs of robofin's hd_meshes/collision/{link1..7,hand,finger}.obj
// (finger y x4, lib/guide.py:278-279) -- stand-in for pybullet_data's meshes, SURVEY.md 8c.
static const double kLinkDims[9][3] = {
    {0.110016, 0.184406, 0.247002}, {0.110033, 0.249024, 0.184393}, {0.192511, 0.166063, 0.176002},
    {0.192507, 0.179, 0.166053},    {0.109996, 0.18493, 0.311199},  {0.179925, 0.132863, 0.100244},
    {0.125333, 0.125297, 0.0548},   {0.063045, 0.204516, 0.091946}, {0.021003, 0.105716, 0.053767}};
// clip_joints limits in degrees (diffusion/diffusion.py:280-298)
static const double kJointLoDeg[7] = {-166, -101, -166, -176, -166, -1, -166};
static const double kJointHiDeg[7] = {166, 101, 166, -4, 166, 215, 166};

static unsigned long long g_constants_ready = 0;   // per-device flags (__constant__ symbols are per device)

static int guide_upload_constants();
int guide_init_constants() {
  if (!once_per_device(&g_constants_ready)) return 0;
  const int rc = guide_upload_constants();
  if (rc) once_per_device_failed(&g_constants_ready);
  return rc;
}

static int guide_upload_constants() {
  float dh[7][4];
  for (int i = 0; i < 7; ++i) {
    float al = (float)kDhHost[i][2];
    dh[i][0] = (float)kDhHost[i][0];
    dh[i][1] = (float)kDhHost[i][1];
    dh[i][2] = cosf(al);
    dh[i][3] = sinf(al);
  }
  float fr[9][12];
  for (int l = 0; l < 9; ++l) {
    float r[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (l >= 7) {  // hand / finger boxes are yawed -45 deg (lib/guide.py:328-340)
      r[0] = 7.07106767e-01f; r[1] = 7.07106795e-01f;
      r[3] = -7.07106795e-01f; r[4] = 7.07106767e-01f;
    }
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) fr[l][i * 4 + j] = r[i * 3 + j];
      fr[l][i * 4 + 3] = (float)kFrameT[l][i];
    }
  }
  double lo[7], hi[7];
  for (int j = 0; j < 7; ++j) {
    lo[j] = kJointLoDeg[j] * (M_PI / 180);
    hi[j] = kJointHiDeg[j] * (M_PI / 180);
  }
  EDMP_CK(cudaMemcpyToSymbol(c_dh, dh, sizeof(dh)));
  EDMP_CK(cudaMemcpyToSymbol(c_frame, fr, sizeof(fr)));
  EDMP_CK(cudaMemcpyToSymbol(c_joint_lo, lo, sizeof(lo)));
  EDMP_CK(cudaMemcpyToSymbol(c_joint_hi, hi, sizeof(hi)));
  return 0;
}

// ---- device geometry -----------------------------------------------------------------------
struct Box {
  float mn[3], mx[3];
};

// joint frames T[i] (3x4 row major), i = 0..6: T_i = T_{i-1} * DH_i(q_i)
__device__ void fk_frames(const float q[7], float T[7][12]) {
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    float s, c;
    sincosf(q[i], &s, &c);
    const float a = c_dh[i][0], d = c_dh[i][1], ca = c_dh[i][2], sa = c_dh[i][3];
    float M[12] = {c,      -s,     0.0f, a,
                   s * ca, c * ca, -sa,  -sa * d,
                   s * sa, c * sa, ca,   ca * d};
    if (i == 0) {
#pragma unroll
      for (int k = 0; k < 12; ++k) T[0][k] = M[k];
    } else {
      const float* P = T[i - 1];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int cidx = 0; cidx < 4; ++cidx) {
          float v = P[r * 4 + 0] * M[cidx];
          v = fmaf(P[r * 4 + 1], M[4 + cidx], v);
          v = fmaf(P[r * 4 + 2], M[8 + cidx], v);
          if (cidx == 3) v = fmaf(P[r * 4 + 3], 1.0f, v);
          T[i][r * 4 + cidx] = v;
        }
      }
    }
  }
}

// World AABB of link box l; optionally the world positions of the arg-min / arg-max vertex of
// every axis (pmin[k] / pmax[k], k = axis) for the Jacobian pull.
template <bool WITH_POS>
__device__ __forceinline__ void link_box(const float T[7][12], int l, const float* __restrict__ half,
                                         Box& b, float pmin[3][3], float pmax[3][3]) {
  const int j = l < 6 ? l : 6;
  const float* P = T[j];
  const float* F = c_frame[l];
  float TL[12];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v = P[r * 4 + 0] * F[c];
      v = fmaf(P[r * 4 + 1], F[4 + c], v);
      v = fmaf(P[r * 4 + 2], F[8 + c], v);
      if (c == 3) v = fmaf(P[r * 4 + 3], 1.0f, v);
      TL[r * 4 + c] = v;
    }
  }
  const float hx = half[l * 3 + 0], hy = half[l * 3 + 1], hz = half[l * 3 + 2];
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    // vertex order of the reference's get_link_vertices (lib/guide.py:203-241)
    const float vx = ((v & 3) == 1 || (v & 3) == 2) ? hx : -hx;
    const float vy = (v & 2) ? hy : -hy;
    const float vz = (v & 4) ? hz : -hz;
    float p[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float w = TL[r * 4 + 0] * vx;
      w = fmaf(TL[r * 4 + 1], vy, w);
      w = fmaf(TL[r * 4 + 2], vz, w);
      w = fmaf(TL[r * 4 + 3], 1.0f, w);
      p[r] = w;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (v == 0 || p[k] < b.mn[k]) {
        b.mn[k] = p[k];
        if (WITH_POS) { pmin[k][0] = p[0]; pmin[k][1] = p[1]; pmin[k][2] = p[2]; }
      }
      if (v == 0 || p[k] > b.mx[k]) {
        b.mx[k] = p[k];
        if (WITH_POS) { pmax[k][0] = p[0]; pmax[k][1] = p[1]; pmax[k][2] = p[2]; }
      }
    }
  }
}

// Obstacle AABBs of one row into shared memory (define_obstacles, lib/guide.py:118-158):
// size = max(dims, expansion) + clearance (float64), half extents in float32, 8 vertices through
// the float32 obstacle transform, min/max.
__device__ void stage_obstacles(const SceneDev* __restrict__ sc, bool use_tables, double expansion,
                                double clearance, float* s_omin, float* s_omax) {
  for (int o = threadIdx.x; o < sc->n_obs; o += blockDim.x) {
    float h[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double sz = sc->dims[o][k];
      if (use_tables) sz = fmax(sz, expansion) + clearance;
      h[k] = (float)sz / 2.0f;
    }
    const float* R = sc->R[o];
    float mn[3], mx[3];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const float vx = ((v & 3) == 1 || (v & 3) == 2) ? h[0] : -h[0];
      const float vy = (v & 2) ? h[1] : -h[1];
      const float vz = (v & 4) ? h[2] : -h[2];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        float w = R[r * 3 + 0] * vx;
        w = fmaf(R[r * 3 + 1], vy, w);
        w = fmaf(R[r * 3 + 2], vz, w);
        w = fmaf(sc->C[o][r], 1.0f, w);
        if (v == 0 || w < mn[r]) mn[r] = w;
        if (v == 0 || w > mx[r]) mx[r] = w;
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      s_omin[o * 3 + k] = mn[k];
      s_omax[o * 3 + k] = mx[k];
    }
  }
}

// sum over obstacles of dV/d(bmax_k) (cmax) and -dV/d(bmin_k) (cmin) for V = prod clamp(len, 0)
__device__ __forceinline__ void face_coefficients(const Box& b, int n_obs, const float* s_omin,
                                                  const float* s_omax, float cmax[3], float cmin[3]) {
  cmax[0] = cmax[1] = cmax[2] = 0.0f;
  cmin[0] = cmin[1] = cmin[2] = 0.0f;
  for (int o = 0; o < n_obs; ++o) {
    const float l0 = fminf(b.mx[0], s_omax[o * 3 + 0]) - fmaxf(b.mn[0], s_omin[o * 3 + 0]);
    const float l1 = fminf(b.mx[1], s_omax[o * 3 + 1]) - fmaxf(b.mn[1], s_omin[o * 3 + 1]);
    const float l2 = fminf(b.mx[2], s_omax[o * 3 + 2]) - fmaxf(b.mn[2], s_omin[o * 3 + 2]);
    if (l0 > 0.0f && l1 > 0.0f && l2 > 0.0f) {
      const float o0 = l1 * l2, o1 = l0 * l2, o2 = l0 * l1;
      if (b.mx[0] < s_omax[o * 3 + 0]) cmax[0] += o0;
      if (b.mx[1] < s_omax[o * 3 + 1]) cmax[1] += o1;
      if (b.mx[2] < s_omax[o * 3 + 2]) cmax[2] += o2;
      if (b.mn[0] > s_omin[o * 3 + 0]) cmin[0] += o0;
      if (b.mn[1] > s_omin[o * 3 + 1]) cmin[1] += o1;
      if (b.mn[2] > s_omin[o * 3 + 2]) cmin[2] += o2;
    }
  }
}

// torch.max(a, b) / torch.min(a, b) backward: the strict winner takes all, exact ties split
// 1/2 - 1/2.  Ties are common here: clip_joints pins neighbouring waypoints to the same limit.
__device__ __forceinline__ float share_max(float mine, float other) {
  return mine > other ? 1.0f : (mine == other ? 0.5f : 0.0f);
}
__device__ __forceinline__ float share_min(float mine, float other) {
  return mine < other ? 1.0f : (mine == other ? 0.5f : 0.0f);
}

__device__ __forceinline__ void load_waypoint(const double* __restrict__ x, int ld, int off, int wi,
                                              int n_inner, const float* start, const float* goal,
                                              bool clip, float q[7]) {
  // wi in [-1, n_inner]: -1 = start, n_inner = goal (swept_volume_cost pads them, :484-492)
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    if (wi < 0) q[j] = start[j];
    else if (wi >= n_inner) q[j] = goal[j];
    else {
      double v = x[j * ld + off + wi];
      if (clip) v = fmin(fmax(v, c_joint_lo[j]), c_joint_hi[j]);
      q[j] = (float)v;
    }
  }
}

struct EndPoints {
  float start[7], goal[7];
};

// Gradient of one interior waypoint w: g[j] = d cost / d q_j(w); xr = the row's [7][ld] float64 trajectory (global
// or shared memory), obstacle boxes of the row staged in s_omin / s_omax.
__device__ __forceinline__ void waypoint_gradient(const double* __restrict__ xr, int ld, int off, int n_inner, bool clip,
                                                  const EndPoints& ep, int w, bool sv, int n_obs,
                                                  const float* s_omin, const float* s_omax, const float* s_half,
                                                  float g[7]) {
  float q[7], T[7][12];
  load_waypoint(xr, ld, off, w, n_inner, ep.start, ep.goal, clip, q);
  fk_frames(q, T);
#pragma unroll
  for (int j = 0; j < 7; ++j) g[j] = 0.0f;
  float Tp[7][12], Tn[7][12];
  if (sv) {
    float qq[7];
    load_waypoint(xr, ld, off, w - 1, n_inner, ep.start, ep.goal, clip, qq);
    fk_frames(qq, Tp);
    load_waypoint(xr, ld, off, w + 1, n_inner, ep.start, ep.goal, clip, qq);
    fk_frames(qq, Tn);
  }
  for (int l = 0; l < 9; ++l) {
    Box b;
    float pmin[3][3], pmax[3][3];
    link_box<true>(T, l, s_half, b, pmin, pmax);
    float cmax[3], cmin[3];
    if (!sv) {
      face_coefficients(b, n_obs, s_omin, s_omax, cmax, cmin);
    } else {
      Box bp, bn, u;
      link_box<false>(Tp, l, s_half, bp, nullptr, nullptr);
      link_box<false>(Tn, l, s_half, bn, nullptr, nullptr);
      float amax[3], amin[3], bmax_[3], bmin_[3];
      // segment (w, w+1)
#pragma unroll
      for (int k = 0; k < 3; ++k) { u.mn[k] = fminf(b.mn[k], bn.mn[k]); u.mx[k] = fmaxf(b.mx[k], bn.mx[k]); }
      face_coefficients(u, n_obs, s_omin, s_omax, amax, amin);
      // segment (w-1, w)
#pragma unroll
      for (int k = 0; k < 3; ++k) { u.mn[k] = fminf(bp.mn[k], b.mn[k]); u.mx[k] = fmaxf(bp.mx[k], b.mx[k]); }
      face_coefficients(u, n_obs, s_omin, s_omax, bmax_, bmin_);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        cmax[k] = amax[k] * share_max(b.mx[k], bn.mx[k]) + bmax_[k] * share_max(b.mx[k], bp.mx[k]);
        cmin[k] = amin[k] * share_min(b.mn[k], bn.mn[k]) + bmin_[k] * share_min(b.mn[k], bp.mn[k]);
      }
    }
    // Jacobian pull: d p_k / d q_i = (z_i x (p - o_i))_k for joints i <= j(l)
    const int jl = l < 6 ? l : 6;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (cmax[k] == 0.0f && cmin[k] == 0.0f) continue;
      const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
      for (int i = 0; i <= jl; ++i) {
        const float z1 = T[i][k1 * 4 + 2], z2 = T[i][k2 * 4 + 2];
        const float o1 = T[i][k1 * 4 + 3], o2 = T[i][k2 * 4 + 3];
        const float dmax = z1 * (pmax[k][k2] - o2) - z2 * (pmax[k][k1] - o1);
        const float dmin = z1 * (pmin[k][k2] - o2) - z2 * (pmin[k][k1] - o1);
        g[i] += cmax[k] * dmax - cmin[k] * dmin;
      }
    }
  }
}

// ---- gradient: raw G (float32) + per-row sum of squares --------------------------------------
// grid = rows, block = 64 (thread w < n_inner handles interior waypoint w)
__global__ void __launch_bounds__(64) guide_grad_kernel(const SceneDev* __restrict__ sc,
                                                        const double* __restrict__ x, int ld, int off,
                                                        int n_inner, bool clip, EndPoints ep, int t,
                                                        const double* __restrict__ clearance,
                                                        const double* __restrict__ expansion,
                                                        const unsigned char* __restrict__ method,
                                                        float* __restrict__ raw, double* __restrict__ rowsq) {
  __shared__ float s_omin[kMaxObs * 3], s_omax[kMaxObs * 3];
  __shared__ float s_half[27];
  __shared__ double s_red[2];
  const int row = blockIdx.x;
  const int w = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  stage_obstacles(sc, t != 0, t != 0 ? expansion[(size_t)row * kTSteps + t - 1] : 0.0,
                  t != 0 ? clearance[(size_t)row * kTSteps + t - 1] : 0.0, s_omin, s_omax);
  if (threadIdx.x < 27) s_half[threadIdx.x] = sc->link_half[threadIdx.x / 3][threadIdx.x % 3];
  __syncthreads();
  const bool sv = method[row] != 0;
  const double* xr = x + (size_t)row * 7 * ld;
  const int n_obs = sc->n_obs;
  double sq = 0.0;
  if (w < n_inner) {
    float g[7];
    waypoint_gradient(xr, ld, off, n_inner, clip, ep, w, sv, n_obs, s_omin, s_omax, s_half, g);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      raw[((size_t)row * 7 + j) * n_inner + w] = g[j];
      sq += (double)g[j] * (double)g[j];
    }
  }
  // deterministic block reduction of the row's sum of squares
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) rowsq[row] = s_red[0] + s_red[1];
}

// ---- mix with the ensemble Frobenius norm and (optionally) apply the guided update -----------
// get_gradient :627-629 and diffusion.py:341.  grid = rows, block = 128.
__global__ void __launch_bounds__(128) guide_apply_kernel(const float* __restrict__ raw,
                                                          const double* __restrict__ rowsq, int n_inner,
                                                          int ensemble_rows,
                                                          const unsigned char* __restrict__ grad_norm,
                                                          const double* __restrict__ schedule, int t,
                                                          double* __restrict__ grad_out,
                                                          double* __restrict__ x, float* __restrict__ xf) {
  __shared__ double s_red[4];
  __shared__ float s_norm;
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  const int e0 = (row / ensemble_rows) * ensemble_rows;
  double s = 0.0;
  for (int r = threadIdx.x; r < ensemble_rows; r += blockDim.x) s += rowsq[e0 + r];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) s_norm = (float)sqrt(s_red[0] + s_red[1] + s_red[2] + s_red[3]);
  __syncthreads();
  const float norm = s_norm;  // np.linalg.norm of a float32 array is float32
  const double gn = grad_norm[row] ? 1.0 : 0.0;
  const double scale = x ? schedule[(size_t)row * kTSteps + t - 1] : 0.0;
  for (int e = threadIdx.x; e < 7 * n_inner; e += blockDim.x) {
    const float G = raw[(size_t)row * 7 * n_inner + e];
    // float32 division, float64 mix -- 0/0 = NaN poisons every row of the ensemble like the
    // reference (SURVEY.md D5 / section 8b "Errors")
    const double mixed = (1.0 - gn) * (double)G + gn * (double)(G / norm);
    if (grad_out) grad_out[(size_t)row * 7 * n_inner + e] = mixed;
    if (x) {
      const int j = e / n_inner, w = e % n_inner;
      const size_t idx = (size_t)row * kRowElems + j * kHorizon + 1 + w;
      const double v = x[idx] - scale * mixed;
      x[idx] = v;
      xf[idx] = (float)v;
    }
  }
}

// ---- the whole per-step tail in ONE launch -----------------------------------------------------------
// posterior update (diffusion.py:116-135) -> clip copy (:328) -> gradient (lib/guide.py:597-623) -> whole-ensemble norm
// mix (:627-629) -> guided update (diffusion.py:341) -> endpoints (:347-349) -> float32 copy for the network.
// A persistent grid (every CTA resident, a row per CTA and walk step, a thread per waypoint) in two phases around one
// grid-wide barrier: the norm couples every row of an ensemble (and a zero norm poisons all of them with NaN like the
// reference), so no row can be updated before every row's gradient exists.  Unguided steps are phase 1 only.
struct TailArgs {
  double* x; float* xf; const float* eps; const double* noise;
  unsigned long long seed;
  int t, ensemble_rows, rows, guided, condition;
  double c1, sqrt_alpha, beta;
  double start[7], goal[7];
  EndPoints ep;
  const SceneDev* sc;
  const double *clearance, *expansion, *schedule;
  const unsigned char *method, *grad_norm;
  float* raw; double* rowsq;
  unsigned* bar; unsigned bar_target;
  // network input image of the NEXT step (unet_input_image): written here instead of by a separate pack launch
  void *pack_hi, *pack_lo;
  int pack_el;                 // 0: no image (write the float32 copy xf), else TcEl of the engine
  unsigned* range_flag;
};

// position l of one row into the position-major input image: 16 channels (7 joints + zeros), hi / lo halves -- the same
// bits pm_pack_input_kernel (conv_pm.cuh) makes from the float32 copy
template <int EL>
__device__ __forceinline__ void pack_input_position(const TailArgs& a, int row, int l, const double* __restrict__ s_x, uint32_t& hmax) {
  float w[8];
#pragma unroll
  for (int j = 0; j < 7; ++j) w[j] = (float)s_x[j * kHorizon + l];
  w[7] = 0.0f;
  uint32_t h[4], r[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    h[e] = pack16x2<EL>(w[2 * e], w[2 * e + 1]);
    range_track<EL>(hmax, h[e]);
    const float2 hf = unpack16x2<EL>(h[e]);
    r[e] = pack16x2<EL>(w[2 * e] - hf.x, w[2 * e + 1] - hf.y);
  }
  const size_t off = (size_t)(row / kPmRows) * pm_img_bytes(kHorizon, 16) + pm_act_off(kHorizon, l, row % kPmRows, 0);
  *reinterpret_cast<uint4*>((uint8_t*)a.pack_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  if (a.pack_lo) *reinterpret_cast<uint4*>((uint8_t*)a.pack_lo + off) = make_uint4(r[0], r[1], r[2], r[3]);
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(64) step_tail_kernel(const __grid_constant__ TailArgs a) {
  __shared__ float s_omin[kMaxObs * 3], s_omax[kMaxObs * 3];
  __shared__ float s_half[27];
  __shared__ double s_red[2];
  __shared__ double s_x[kRowElems];
  __shared__ float s_norm;
  __shared__ int s_norm_ens;
  constexpr int n_inner = kHorizon - 2;
  uint32_t hmax = 0;   // largest |hi| half pattern written to the input image (operand range check, conv_tc.cuh)
  // (guided steps release the next kernel only after the barrier: its CTAs must not take the place of CTAs of this
  // grid that are not resident yet)
  if (!a.guided) pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x < 27 && a.guided) s_half[threadIdx.x] = a.sc->link_half[threadIdx.x / 3][threadIdx.x % 3];
  for (int row = blockIdx.x; row < a.rows; row += gridDim.x) {
    // x_{t-1} = (x_t - c1 * eps) / sqrt(alpha) + beta * z; row 0 of an ensemble gets z = 0 at t == 1
    for (int e = threadIdx.x; e < kRowElems; e += 64) {
      const size_t i = (size_t)row * kRowElems + e;
      const int j = e / kHorizon, l = e % kHorizon;
      double z = a.noise ? a.noise[i] : philox_normal(a.seed, (uint32_t)a.t, i);
      if (a.t == 1 && (row % a.ensemble_rows) == 0) z = 0.0;
      double v = (a.x[i] - a.c1 * (double)a.eps[i]) / a.sqrt_alpha + a.beta * z;
      if (a.condition && l == 0) v = a.start[j];
      if (a.condition && l == kHorizon - 1) v = a.goal[j];
      s_x[e] = v;
      a.x[i] = v;
      if (!a.pack_el) a.xf[i] = (float)v;
    }
    if (!a.guided) {
      if (a.pack_el) {
        __syncthreads();
        if (threadIdx.x < kHorizon) {
          if (a.pack_el == TC_EL_F16) pack_input_position<TC_EL_F16>(a, row, threadIdx.x, s_x, hmax);
          else pack_input_position<TC_EL_BF16>(a, row, threadIdx.x, s_x, hmax);
        }
        __syncthreads();   // the next row overwrites s_x
      }
      continue;
    }
    stage_obstacles(a.sc, true, a.expansion[(size_t)row * kTSteps + a.t - 1], a.clearance[(size_t)row * kTSteps + a.t - 1],
                    s_omin, s_omax);
    __syncthreads();
    double sq = 0.0;
    const int w = threadIdx.x;
    if (w < n_inner) {
      float g[7];
      waypoint_gradient(s_x, kHorizon, 1, n_inner, true, a.ep, w, a.method[row] != 0, a.sc->n_obs, s_omin, s_omax, s_half, g);
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        a.raw[((size_t)row * 7 + j) * n_inner + w] = g[j];
        sq += (double)g[j] * (double)g[j];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sq;
    __syncthreads();   // (also: every thread is done with s_x and the obstacle boxes of this row)
    if (threadIdx.x == 0) a.rowsq[row] = s_red[0] + s_red[1];
  }
  if (!a.guided) {
    if (a.pack_el == TC_EL_F16) range_report<TC_EL_F16>(hmax, a.range_flag);
    return;
  }

  // ---- grid-wide barrier: every row's raw gradient and sum of squares is written ----
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(a.bar, 1u);
    const long long t0 = clock64();
    while (ld_acquire_u32(a.bar) < a.bar_target) {
      __nanosleep(64);
      if (clock64() - t0 > (1ll << 32)) {   // ~2 s: a protocol bug traps instead of hanging the GPU
        printf("edmp: step_tail grid barrier timed out (block %d: %u of %u)\n", blockIdx.x, ld_acquire_u32(a.bar), a.bar_target);
        __trap();
      }
    }
    s_norm_ens = -1;
  }
  __syncthreads();
  pdl_launch_dependents();

  for (int row = blockIdx.x; row < a.rows; row += gridDim.x) {
    const int ens = row / a.ensemble_rows;
    if (ens != s_norm_ens) {
      // Frobenius norm of the ensemble's whole gradient (lib/guide.py:629), summed in a fixed order
      const double* rs = a.rowsq + (size_t)ens * a.ensemble_rows;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int r = threadIdx.x;
      for (; r + 192 < a.ensemble_rows; r += 256) { s0 += rs[r]; s1 += rs[r + 64]; s2 += rs[r + 128]; s3 += rs[r + 192]; }
      for (; r < a.ensemble_rows; r += 64) s0 += rs[r];
      double sum = (s0 + s1) + (s2 + s3);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      __syncthreads();   // every thread has read s_norm_ens
      if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sum;
      __syncthreads();
      if (threadIdx.x == 0) { s_norm = (float)sqrt(s_red[0] + s_red[1]); s_norm_ens = ens; }
      __syncthreads();
    }
    const float norm = s_norm;  // np.linalg.norm of a float32 array is float32
    const double gn = a.grad_norm[row] ? 1.0 : 0.0;
    const double scale = a.schedule[(size_t)row * kTSteps + a.t - 1];
    for (int e = threadIdx.x; e < kRowElems; e += 64) {
      const int j = e / kHorizon, l = e % kHorizon;
      const size_t idx = (size_t)row * kRowElems + e;
      double v = a.x[idx];
      if (l >= 1 && l <= n_inner) {
        const float G = a.raw[((size_t)row * 7 + j) * n_inner + (l - 1)];
        // float32 division, float64 mix -- 0/0 = NaN poisons every row of the ensemble like the reference
        const double mixed = (1.0 - gn) * (double)G + gn * (double)(G / norm);
        v -= scale * mixed;
        a.x[idx] = v;
        if (!a.pack_el) a.xf[idx] = (float)v;
      }
      s_x[e] = v;
    }
    if (a.pack_el) {
      __syncthreads();
      if (threadIdx.x < kHorizon) {
        if (a.pack_el == TC_EL_F16) pack_input_position<TC_EL_F16>(a, row, threadIdx.x, s_x, hmax);
        else pack_input_position<TC_EL_BF16>(a, row, threadIdx.x, s_x, hmax);
      }
      __syncthreads();   // the next row overwrites s_x
    }
  }
  if (a.pack_el == TC_EL_F16) range_report<TC_EL_F16>(hmax, a.range_flag);
}

// ---- full volume tensors for the cost()/swept_volume_cost() API --------------------------------
// grid = (rows), block = 64; thread = waypoint (iv) or segment (sv)
__global__ void __launch_bounds__(64) guide_volumes_kernel(const SceneDev* __restrict__ sc,
                                                           const float* __restrict__ q, int n, int mode,
                                                           EndPoints ep, int t,
                                                           const double* __restrict__ clearance,
                                                           const double* __restrict__ expansion,
                                                           float* __restrict__ vol) {
  __shared__ float s_omin[kMaxObs * 3], s_omax[kMaxObs * 3];
  __shared__ float s_half[27];
  const int row = blockIdx.x;
  stage_obstacles(sc, t != 0, t != 0 ? expansion[(size_t)row * kTSteps + t - 1] : 0.0,
                  t != 0 ? clearance[(size_t)row * kTSteps + t - 1] : 0.0, s_omin, s_omax);
  if (threadIdx.x < 27) s_half[threadIdx.x] = sc->link_half[threadIdx.x / 3][threadIdx.x % 3];
  __syncthreads();
  const int n_obs = sc->n_obs;
  const int n_out = mode ? n + 1 : n;
  const float* qr = q + (size_t)row * 7 * n;
  for (int w = threadIdx.x; w < n_out; w += blockDim.x) {
    float qa[7], qb[7], Ta[7][12], Tb[7][12];
    // iv: waypoint w.  sv: segment w joins padded waypoints w-1 and w (start | q | goal).
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      if (!mode) qa[j] = qr[j * n + w];
      else {
        qa[j] = (w == 0) ? ep.start[j] : qr[j * n + w - 1];
        qb[j] = (w == n) ? ep.goal[j] : qr[j * n + w];
      }
    }
    fk_frames(qa, Ta);
    if (mode) fk_frames(qb, Tb);
    for (int l = 0; l < 9; ++l) {
      Box b;
      link_box<false>(Ta, l, s_half, b, nullptr, nullptr);
      if (mode) {
        Box c;
        link_box<false>(Tb, l, s_half, c, nullptr, nullptr);
#pragma unroll
        for (int k = 0; k < 3; ++k) { b.mn[k] = fminf(b.mn[k], c.mn[k]); b.mx[k] = fmaxf(b.mx[k], c.mx[k]); }
      }
      for (int o = 0; o < n_obs; ++o) {
        float v = 1.0f;
#pragma unroll
        for (int k = 0; k < 3; ++k)
          v *= fmaxf(fminf(b.mx[k], s_omax[o * 3 + k]) - fmaxf(b.mn[k], s_omin[o * 3 + k]), 0.0f);
        vol[((size_t)row * n_out + w) * (9 * n_obs) + l * n_obs + o] = v;
      }
    }
  }
}

// ---- best-of-ensemble cost: sum of swept volumes at t = 0 (lib/guide.py:637-653) ---------------
__global__ void __launch_bounds__(64) guide_final_cost_kernel(const SceneDev* __restrict__ sc,
                                                              const double* __restrict__ traj, EndPoints ep,
                                                              float* __restrict__ cost) {
  __shared__ float s_omin[kMaxObs * 3], s_omax[kMaxObs * 3];
  __shared__ float s_half[27];
  __shared__ double s_red[2];
  const int row = blockIdx.x;
  stage_obstacles(sc, false, 0.0, 0.0, s_omin, s_omax);
  if (threadIdx.x < 27) s_half[threadIdx.x] = sc->link_half[threadIdx.x / 3][threadIdx.x % 3];
  __syncthreads();
  const int n_obs = sc->n_obs;
  const double* xr = traj + (size_t)row * kRowElems;
  double total = 0.0;
  const int s = threadIdx.x;  // segment s joins waypoints s and s+1 of [start | interior | goal]
  if (s < kHorizon - 1) {
    float qa[7], qb[7], Ta[7][12], Tb[7][12];
    load_waypoint(xr, kHorizon, 1, s - 1, kHorizon - 2, ep.start, ep.goal, false, qa);
    load_waypoint(xr, kHorizon, 1, s, kHorizon - 2, ep.start, ep.goal, false, qb);
    fk_frames(qa, Ta);
    fk_frames(qb, Tb);
    float acc = 0.0f;
    for (int l = 0; l < 9; ++l) {
      Box b, c;
      link_box<false>(Ta, l, s_half, b, nullptr, nullptr);
      link_box<false>(Tb, l, s_half, c, nullptr, nullptr);
#pragma unroll
      for (int k = 0; k < 3; ++k) { b.mn[k] = fminf(b.mn[k], c.mn[k]); b.mx[k] = fmaxf(b.mx[k], c.mx[k]); }
      for (int o = 0; o < n_obs; ++o) {
        float v = 1.0f;
#pragma unroll
        for (int k = 0; k < 3; ++k)
          v *= fmaxf(fminf(b.mx[k], s_omax[o * 3 + k]) - fmaxf(b.mn[k], s_omin[o * 3 + k]), 0.0f);
        acc += v;
      }
    }
    total = (double)acc;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = total;
  __syncthreads();
  if (threadIdx.x == 0) cost[row] = (float)(s_red[0] + s_red[1]);
}

// ---- host side ---------------------------------------------------------------------------------
static void quat_xyzw_to_matrix(const double* qin, double R[9]) {
  // scipy Rotation.from_quat(...).as_matrix() (scalar last, normalised) -- lib/guide.py:143
  double n = std::sqrt(qin[0] * qin[0] + qin[1] * qin[1] + qin[2] * qin[2] + qin[3] * qin[3]);
  double x = qin[0] / n, y = qin[1] / n, z = qin[2] / n, w = qin[3] / n;
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

int scene_create(const double* cfg, int n_obs, const double* link_dims, Scene** out) {
  EDMP_REQUIRE(n_obs >= 1 && n_obs <= kMaxObs, "n_obs must be in 1..64");
  if (guide_init_constants()) return 1;
  Scene* s = new Scene();
  std::memset(&s->host, 0, sizeof(SceneDev));
  s->host.n_obs = n_obs;
  for (int o = 0; o < n_obs; ++o) {
    const double* c = cfg + o * 10;
    double R[9];
    quat_xyzw_to_matrix(c + 3, R);
    for (int k = 0; k < 9; ++k) s->host.R[o][k] = (float)R[k];
    for (int k = 0; k < 3; ++k) {
      s->host.C[o][k] = (float)c[k];
      s->host.dims[o][k] = c[7 + k];
    }
  }
  for (int l = 0; l < 9; ++l)
    for (int k = 0; k < 3; ++k)
      s->host.link_half[l][k] = (float)(link_dims ? link_dims[l * 3 + k] : kLinkDims[l][k]) / 2.0f;
  if (cudaMalloc(&s->dev, sizeof(SceneDev)) != cudaSuccess ||
      cudaMemcpy(s->dev, &s->host, sizeof(SceneDev), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("scene_create: device allocation/copy failed");
    delete s;
    return 1;
  }
  *out = s;
  return 0;
}

void scene_destroy(Scene* s) {
  if (!s) return;
  cudaFree(s->dev);
  cudaFree(s->clearance); cudaFree(s->expansion); cudaFree(s->schedule);
  cudaFree(s->method); cudaFree(s->grad_norm);
  cudaFree(s->raw); cudaFree(s->rowsq); cudaFree(s->zero_tables);
  delete s;
}

static int ensure_work(Scene* s, int rows, int n_inner) {
  size_t need = (size_t)rows * 7 * n_inner;
  if (need > s->raw_cap) {
    cudaFree(s->raw);
    s->raw = nullptr;
    EDMP_CK(cudaMalloc(&s->raw, need * sizeof(float)));
    s->raw_cap = need;
  }
  if ((size_t)rows > s->rowsq_cap) {
    cudaFree(s->rowsq);
    s->rowsq = nullptr;
    EDMP_CK(cudaMalloc(&s->rowsq, rows * sizeof(double)));
    s->rowsq_cap = rows;
  }
  return 0;
}

int scene_set_tables(Scene* s, const double* clr, const double* exp_, const double* sched,
                     const double* method, const double* gnorm, int rows, int ensemble_rows) {
  EDMP_REQUIRE(rows > 0 && ensemble_rows > 0 && rows % ensemble_rows == 0,
               "ensemble_rows must divide rows");
  cudaFree(s->clearance); cudaFree(s->expansion); cudaFree(s->schedule);
  cudaFree(s->method); cudaFree(s->grad_norm);
  s->clearance = s->expansion = s->schedule = nullptr;
  s->method = s->grad_norm = nullptr;
  size_t nb = (size_t)rows * kTSteps * sizeof(double);
  EDMP_CK(cudaMalloc(&s->clearance, nb));
  EDMP_CK(cudaMalloc(&s->expansion, nb));
  EDMP_CK(cudaMalloc(&s->schedule, nb));
  EDMP_CK(cudaMalloc(&s->method, rows));
  EDMP_CK(cudaMalloc(&s->grad_norm, rows));
  EDMP_CK(cudaMemcpy(s->clearance, clr, nb, cudaMemcpyHostToDevice));
  EDMP_CK(cudaMemcpy(s->expansion, exp_, nb, cudaMemcpyHostToDevice));
  EDMP_CK(cudaMemcpy(s->schedule, sched, nb, cudaMemcpyHostToDevice));
  std::vector<unsigned char> m(rows), g(rows);
  for (int r = 0; r < rows; ++r) {
    m[r] = method[r] != 0.0;
    g[r] = gnorm[r] != 0.0;
  }
  EDMP_CK(cudaMemcpy(s->method, m.data(), rows, cudaMemcpyHostToDevice));
  EDMP_CK(cudaMemcpy(s->grad_norm, g.data(), rows, cudaMemcpyHostToDevice));
  s->rows = rows;
  s->ensemble_rows = ensemble_rows;
  return 0;
}

static EndPoints make_endpoints(const double* start, const double* goal) {
  EndPoints ep;
  for (int j = 0; j < 7; ++j) {
    ep.start[j] = start ? (float)start[j] : 0.0f;
    ep.goal[j] = goal ? (float)goal[j] : 0.0f;
  }
  return ep;
}

int guide_gradient_launch(Scene* s, const double* x, int ld, int off, int n_inner, bool clip,
                          const double* start_h, const double* goal_h, int t, int rows,
                          double* grad_out, float* raw_out, double* x_state, float* xf_state,
                          cudaStream_t st) {
  EDMP_REQUIRE(s->rows == rows, "guide tables were set for a different row count");
  EDMP_REQUIRE(t >= 1 && t <= kTSteps, "t out of range");
  EDMP_REQUIRE(n_inner >= 1 && n_inner <= 64, "n_inner must be in 1..64");
  if (ensure_work(s, rows, n_inner)) return 1;
  EndPoints ep = make_endpoints(start_h, goal_h);
  launch_pdl(guide_grad_kernel, dim3(rows), dim3(64), 0, st, (const SceneDev*)s->dev, x, ld, off, n_inner, clip, ep, t,
             (const double*)s->clearance, (const double*)s->expansion, (const unsigned char*)s->method, s->raw, s->rowsq);
  launch_pdl(guide_apply_kernel, dim3(rows), dim3(128), 0, st, (const float*)s->raw, (const double*)s->rowsq, n_inner,
             s->ensemble_rows, (const unsigned char*)s->grad_norm, (const double*)s->schedule, t, grad_out, x_state,
             xf_state);
  EDMP_CK(cudaGetLastError());
  if (raw_out)
    EDMP_CK(cudaMemcpyAsync(raw_out, s->raw, (size_t)rows * 7 * n_inner * sizeof(float),
                            cudaMemcpyDeviceToDevice, st));
  return 0;
}

// One launch per reverse step after the UNet (see step_tail_kernel).  bar: a device counter zeroed by the caller at the
// start of a pass; *bar_epoch counts the guided steps since then.
int guide_step_tail_launch(Scene* s, double* x, float* xf, const float* eps, const double* noise, uint64_t seed, int t,
                           double c1, double sqrt_alpha, double beta, const double* start_h, const double* goal_h,
                           int rows, bool guided, bool condition, unsigned* bar, unsigned* bar_epoch, void* pack_hi,
                           void* pack_lo, int pack_el, unsigned* range_flag, cudaStream_t st) {
  EDMP_REQUIRE(s->rows == rows, "guide tables were set for a different row count");
  EDMP_REQUIRE(t >= 1 && t <= kTSteps, "t out of range");
  if (guided && ensure_work(s, rows, kHorizon - 2)) return 1;
  if (s->tail_grid_cap == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    EDMP_CK(cudaGetDevice(&dev));
    EDMP_CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    EDMP_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_tail_kernel, 64, 0));
    EDMP_REQUIRE(per_sm > 0 && sms > 0, "step_tail_kernel does not fit an SM");
    s->tail_grid_cap = per_sm * sms;
  }
  const int grid = std::min(rows, s->tail_grid_cap);
  TailArgs a;
  std::memset(&a, 0, sizeof(a));
  a.x = x; a.xf = xf; a.eps = eps; a.noise = noise; a.seed = seed;
  a.t = t; a.ensemble_rows = s->ensemble_rows; a.rows = rows; a.guided = guided ? 1 : 0; a.condition = condition ? 1 : 0;
  a.c1 = c1; a.sqrt_alpha = sqrt_alpha; a.beta = beta;
  for (int j = 0; j < 7; ++j) { a.start[j] = start_h[j]; a.goal[j] = goal_h[j]; }
  a.ep = make_endpoints(start_h, goal_h);
  a.sc = s->dev;
  a.clearance = s->clearance; a.expansion = s->expansion; a.schedule = s->schedule;
  a.method = s->method; a.grad_norm = s->grad_norm;
  a.raw = s->raw; a.rowsq = s->rowsq;
  a.bar = bar;
  a.pack_hi = pack_hi; a.pack_lo = pack_lo; a.pack_el = pack_hi ? pack_el : 0; a.range_flag = range_flag;
  if (guided) { *bar_epoch += 1; a.bar_target = *bar_epoch * (unsigned)grid; }
  launch_pdl(step_tail_kernel, dim3(grid), dim3(64), 0, st, a);
  EDMP_CK(cudaGetLastError());
  return 0;
}

int guide_volumes_launch(Scene* s, const float* q, const double* start_h, const double* goal_h, int t,
                         int mode, int rows, int n, float* vol, cudaStream_t st) {
  EDMP_REQUIRE(t >= 0 && t <= kTSteps, "t out of range");
  EDMP_REQUIRE(n >= 1, "n must be positive");
  if (t != 0) EDMP_REQUIRE(s->rows == rows, "guide tables were set for a different row count");
  EndPoints ep = make_endpoints(start_h, goal_h);
  guide_volumes_kernel<<<rows, 64, 0, st>>>(s->dev, q, n, mode, ep, t, s->clearance, s->expansion, vol);
  EDMP_CK(cudaGetLastError());
  return 0;
}

int guide_final_cost_launch(Scene* s, const double* traj, const double* start_h, const double* goal_h,
                            int rows, float* cost, cudaStream_t st) {
  EndPoints ep = make_endpoints(start_h, goal_h);
  guide_final_cost_kernel<<<rows, 64, 0, st>>>(s->dev, traj, ep, cost);
  EDMP_CK(cudaGetLastError());
  return 0;
}

}  // namespace edmp
