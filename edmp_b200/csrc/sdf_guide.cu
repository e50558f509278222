// Sphere / signed-distance guide family (SURVEY.md section 8 a-S, BASELINE.json configs[4]): analytic Franka FK ->
// 59 link collision spheres -> signed distance to the scene's primitives (or nearest-point distance to a point
// cloud) -> hinge cost and its analytic gradient with respect to the joint angles.  One launch, no autograd.
//
// Specification sources in the reference tree (vendored there, never called by its infer_serial.py):
//   * spheres: robofin/robofin/robots.py:58-174 (59 spheres, 11 radius groups, centres in link frames);
//   * FK: the URDF chain robofin/robofin/urdf/franka_panda/panda.urdf:47-235,:310-324 (joint origins xyz / rpy,
//     revolute about the child frame's z; hand yawed -pi/4; fingers prismatic at 0.025,
//     robofin/robofin/pointcloud/torch.py:343-350) evaluated like torch_urdf.py:466-522;
//   * sphere centres = link frame x centre: pointcloud/torch.py:340-365 (compute_spheres);
//   * primitive SDFs: mpinets/geometry.py:238-288 (cuboid: |p| - dims/2, outside norm + inside max) and :456-505
//     (cylinder: the same in (radial, axial)); scene SDF = min over primitives;
//   * collision predicate sdf <= radius (mpinets/model.py:301-312), hinge margin 0.03 (mpinets/loss.py:88-94).
// The cost is the builder's (the reference has no sphere guide): sum over (waypoint, sphere) of
// max(0, radius + margin - sdf(centre)); its gradient uses d centre / d q_i = z_i x (centre - o_i) for the joints
// upstream of the sphere's link and grad sdf of the nearest primitive.  The point-cloud variant has no reference
// implementation at all: clearance = min_p |centre - p| - radius.
#include "sdf_guide.h"

#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace edmp {

constexpr int kSdfSpheres = 59;
constexpr int kSdfMaxPrims = 64;
constexpr int kSdfThreads = 256;
constexpr int kSdfMaxWp = 64;

struct SdfSphere { float x, y, z, r; int link; };

// robofin/robofin/robots.py:58-174, in list order
static const SdfSphere kSphereTable[kSdfSpheres] = {
    {0.0f, 0.0f, 0.05f, 0.08f, 0},
    {0.0f, -0.08f, 0.0f, 0.06f, 1}, {0.0f, -0.03f, 0.0f, 0.06f, 1}, {0.0f, 0.0f, -0.12f, 0.06f, 1}, {0.0f, 0.0f, -0.17f, 0.06f, 1},
    {0.0f, 0.0f, 0.03f, 0.06f, 2}, {0.0f, 0.0f, 0.08f, 0.06f, 2}, {0.0f, -0.12f, 0.0f, 0.06f, 2}, {0.0f, -0.17f, 0.0f, 0.06f, 2},
    {0.0f, 0.0f, -0.1f, 0.06f, 3},
    {-0.08f, 0.095f, 0.0f, 0.06f, 4},
    {0.0f, 0.055f, 0.0f, 0.06f, 5}, {0.0f, 0.075f, 0.0f, 0.06f, 5}, {0.0f, 0.0f, -0.22f, 0.06f, 5},
    {0.0f, 0.0f, -0.06f, 0.05f, 3},
    {0.0f, 0.05f, -0.18f, 0.05f, 5},
    {0.0f, 0.0f, 0.0f, 0.05f, 6}, {0.08f, -0.01f, 0.0f, 0.05f, 6},
    {0.0f, 0.0f, 0.07f, 0.05f, 7},
    {0.08f, 0.06f, 0.0f, 0.055f, 3}, {0.08f, 0.02f, 0.0f, 0.055f, 3},
    {0.0f, 0.0f, 0.02f, 0.055f, 4}, {0.0f, 0.0f, 0.06f, 0.055f, 4}, {-0.08f, 0.06f, 0.0f, 0.055f, 4},
    {0.01f, 0.08f, -0.14f, 0.025f, 5}, {0.01f, 0.085f, -0.11f, 0.025f, 5}, {0.01f, 0.09f, -0.08f, 0.025f, 5}, {0.01f, 0.095f, -0.05f, 0.025f, 5},
    {-0.01f, 0.08f, -0.14f, 0.025f, 5}, {-0.01f, 0.085f, -0.11f, 0.025f, 5}, {-0.01f, 0.09f, -0.08f, 0.025f, 5}, {-0.01f, 0.095f, -0.05f, 0.025f, 5},
    {0.02f, 0.04f, 0.08f, 0.025f, 7}, {0.04f, 0.02f, 0.08f, 0.025f, 7},
    {0.08f, 0.035f, 0.0f, 0.052f, 6},
    {0.04f, 0.06f, 0.085f, 0.02f, 7}, {0.06f, 0.04f, 0.085f, 0.02f, 7},
    {0.0f, -0.075f, 0.01f, 0.028f, 8}, {0.0f, -0.045f, 0.01f, 0.028f, 8}, {0.0f, -0.015f, 0.01f, 0.028f, 8},
    {0.0f, 0.015f, 0.01f, 0.028f, 8}, {0.0f, 0.045f, 0.01f, 0.028f, 8}, {0.0f, 0.075f, 0.01f, 0.028f, 8},
    {0.0f, -0.075f, 0.03f, 0.026f, 8}, {0.0f, -0.045f, 0.03f, 0.026f, 8}, {0.0f, -0.015f, 0.03f, 0.026f, 8},
    {0.0f, 0.015f, 0.03f, 0.026f, 8}, {0.0f, 0.045f, 0.03f, 0.026f, 8}, {0.0f, 0.075f, 0.03f, 0.026f, 8},
    {0.0f, -0.075f, 0.05f, 0.024f, 8}, {0.0f, -0.045f, 0.05f, 0.024f, 8}, {0.0f, -0.015f, 0.05f, 0.024f, 8},
    {0.0f, 0.015f, 0.05f, 0.024f, 8}, {0.0f, 0.045f, 0.05f, 0.024f, 8}, {0.0f, 0.075f, 0.05f, 0.024f, 8},
    {0.0f, 0.015f, 0.022f, 0.012f, 9}, {0.0f, 0.008f, 0.044f, 0.012f, 9},
    {0.0f, -0.015f, 0.022f, 0.012f, 10}, {0.0f, -0.008f, 0.044f, 0.012f, 10},
};

__constant__ SdfSphere c_spheres[kSdfSpheres];

struct SdfPrim {        // world -> local: p = Rinv (x - c)
  float rinv[9];
  float c[3];
  float h[3];           // box half extents; cylinder: (radius, half height, -)
  float kind;           // 0 box, 1 cylinder
};

struct SdfScene {
  SdfPrim* prims = nullptr;
  int n_prims = 0;
};

// ---- FK: link frames as 3x4 row-major (R | t) ---------------------------------------------------------------
struct Frame { float m[12]; };

// child = parent * Trans(x, y, z) * Rx(roll) * Rz(q)      (URDF joint origin, then the revolute joint about z)
__device__ __forceinline__ Frame joint_frame(const Frame& p, float x, float y, float z, float roll_sin, float roll_cos, float q) {
  float s, c;
  sincosf(q, &s, &c);
  // L = Rx(roll) * Rz(q)
  const float l00 = c, l01 = -s, l02 = 0.0f;
  const float l10 = roll_cos * s, l11 = roll_cos * c, l12 = -roll_sin;
  const float l20 = roll_sin * s, l21 = roll_sin * c, l22 = roll_cos;
  Frame f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float a = p.m[r * 4], b = p.m[r * 4 + 1], d = p.m[r * 4 + 2];
    f.m[r * 4] = a * l00 + b * l10 + d * l20;
    f.m[r * 4 + 1] = a * l01 + b * l11 + d * l21;
    f.m[r * 4 + 2] = a * l02 + b * l12 + d * l22;
    f.m[r * 4 + 3] = a * x + b * y + d * z + p.m[r * 4 + 3];
  }
  return f;
}

// frames of link1..7, hand, left finger, right finger for one configuration (panda.urdf:47-235,:310-324)
__device__ void franka_fk(const float q[7], Frame out[10]) {
  Frame f;
#pragma unroll
  for (int i = 0; i < 12; ++i) f.m[i] = (i == 0 || i == 5 || i == 10) ? 1.0f : 0.0f;
  f = joint_frame(f, 0.0f, 0.0f, 0.333f, 0.0f, 1.0f, q[0]);            out[0] = f;   // roll 0
  f = joint_frame(f, 0.0f, 0.0f, 0.0f, -1.0f, 0.0f, q[1]);             out[1] = f;   // roll -pi/2
  f = joint_frame(f, 0.0f, -0.316f, 0.0f, 1.0f, 0.0f, q[2]);           out[2] = f;   // roll +pi/2
  f = joint_frame(f, 0.0825f, 0.0f, 0.0f, 1.0f, 0.0f, q[3]);           out[3] = f;
  f = joint_frame(f, -0.0825f, 0.384f, 0.0f, -1.0f, 0.0f, q[4]);       out[4] = f;
  f = joint_frame(f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, q[5]);              out[5] = f;
  f = joint_frame(f, 0.088f, 0.0f, 0.0f, 1.0f, 0.0f, q[6]);            out[6] = f;
  // link8 = Trans(0, 0, 0.107); hand = Rz(-pi/4)
  Frame h = joint_frame(f, 0.0f, 0.0f, 0.107f, 0.0f, 1.0f, -0.78539816339744831f);
  out[7] = h;
  Frame lf = h, rf = h;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    lf.m[r * 4 + 3] = h.m[r * 4 + 3] + h.m[r * 4 + 1] * 0.025f + h.m[r * 4 + 2] * 0.0584f;
    rf.m[r * 4 + 3] = h.m[r * 4 + 3] - h.m[r * 4 + 1] * 0.025f + h.m[r * 4 + 2] * 0.0584f;
  }
  out[8] = lf;
  out[9] = rf;
}

// signed distance of world point x to primitive pr and its world-space gradient
__device__ __forceinline__ float prim_sdf(const SdfPrim& pr, float x, float y, float z, float& gx, float& gy, float& gz) {
  const float dx = x - pr.c[0], dy = y - pr.c[1], dz = z - pr.c[2];
  const float px = pr.rinv[0] * dx + pr.rinv[1] * dy + pr.rinv[2] * dz;
  const float py = pr.rinv[3] * dx + pr.rinv[4] * dy + pr.rinv[5] * dz;
  const float pz = pr.rinv[6] * dx + pr.rinv[7] * dy + pr.rinv[8] * dz;
  float lx, ly, lz, sdf;   // local gradient
  if (pr.kind < 0.5f) {
    const float ax = fabsf(px) - pr.h[0], ay = fabsf(py) - pr.h[1], az = fabsf(pz) - pr.h[2];
    const float ox = fmaxf(ax, 0.0f), oy = fmaxf(ay, 0.0f), oz = fmaxf(az, 0.0f);
    const float outside = sqrtf(ox * ox + oy * oy + oz * oz);
    const float inner = fmaxf(ax, fmaxf(ay, az));
    sdf = outside + fminf(inner, 0.0f);
    if (outside > 0.0f) {
      const float inv = 1.0f / outside;
      lx = copysignf(ox * inv, px); ly = copysignf(oy * inv, py); lz = copysignf(oz * inv, pz);
    } else {
      lx = (ax >= ay && ax >= az) ? copysignf(1.0f, px) : 0.0f;
      ly = (lx == 0.0f && ay >= az) ? copysignf(1.0f, py) : 0.0f;
      lz = (lx == 0.0f && ly == 0.0f) ? copysignf(1.0f, pz) : 0.0f;
    }
  } else {
    const float rho = sqrtf(px * px + py * py);
    const float ar = rho - pr.h[0], az = fabsf(pz) - pr.h[1];
    const float orr = fmaxf(ar, 0.0f), oz = fmaxf(az, 0.0f);
    const float outside = sqrtf(orr * orr + oz * oz);
    sdf = outside + fminf(fmaxf(ar, az), 0.0f);
    const float irho = rho > 0.0f ? 1.0f / rho : 0.0f;
    float wr, wz;   // weights of the radial / axial unit vectors
    if (outside > 0.0f) { wr = orr / outside; wz = oz / outside; }
    else { wr = ar >= az ? 1.0f : 0.0f; wz = 1.0f - wr; }
    lx = wr * px * irho; ly = wr * py * irho; lz = copysignf(wz, pz);
  }
  // world gradient = R * local = Rinv^T * local
  gx = pr.rinv[0] * lx + pr.rinv[3] * ly + pr.rinv[6] * lz;
  gy = pr.rinv[1] * lx + pr.rinv[4] * ly + pr.rinv[7] * lz;
  gz = pr.rinv[2] * lx + pr.rinv[5] * ly + pr.rinv[8] * lz;
  return sdf;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One CTA per trajectory row.  Phase 1: a thread per waypoint runs the FK (link frames to shared memory; the
// primitives are staged once per CTA).  Phase 2: 64 lanes (two warps) per waypoint, a lane per sphere: centre,
// nearest primitive (serial loop over the staged primitives), hinge, analytic gradient; the waypoint's nearest
// clearance and its 7 gradient components are warp-shuffle reductions.
__global__ void __launch_bounds__(kSdfThreads) sdf_guide_kernel(const float* __restrict__ q, int n, int rows,
                                                                const SdfPrim* __restrict__ prims, int n_prims, float margin,
                                                                float* __restrict__ cost, float* __restrict__ grad,
                                                                float* __restrict__ clearance) {
  __shared__ Frame s_frames[kSdfMaxWp][10];
  __shared__ SdfPrim s_prims[kSdfMaxPrims];
  __shared__ float s_cost[kSdfThreads / 32];
  __shared__ float s_clear[kSdfMaxWp][2];
  __shared__ float s_grad[kSdfMaxWp][2][7];
  const int row = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < n_prims * (int)(sizeof(SdfPrim) / 4); i += kSdfThreads)
    reinterpret_cast<float*>(s_prims)[i] = reinterpret_cast<const float*>(prims)[i];
  if (tid < n) {
    float qq[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) qq[j] = q[((size_t)row * 7 + j) * n + tid];
    franka_fk(qq, s_frames[tid]);
  }
  __syncthreads();
  float my_cost = 0.0f;
  const int slot = tid >> 6, s = tid & 63;          // 4 waypoints in flight, sphere index
  for (int w0 = 0; w0 < n; w0 += 4) {
    const int w = w0 + slot;
    float clear = 3.0e38f, g[7] = {0, 0, 0, 0, 0, 0, 0};
    if (w < n && s < kSdfSpheres) {
      const SdfSphere sp = c_spheres[s];
      float cx = sp.x, cy = sp.y, cz = sp.z;
      if (sp.link > 0) {
        const Frame& f = s_frames[w][sp.link - 1];
        cx = f.m[0] * sp.x + f.m[1] * sp.y + f.m[2] * sp.z + f.m[3];
        cy = f.m[4] * sp.x + f.m[5] * sp.y + f.m[6] * sp.z + f.m[7];
        cz = f.m[8] * sp.x + f.m[9] * sp.y + f.m[10] * sp.z + f.m[11];
      }
      float best = 3.0e38f, nx = 0.0f, ny = 0.0f, nz = 0.0f;
      for (int m = 0; m < n_prims; ++m) {
        float gx, gy, gz;
        const float d = prim_sdf(s_prims[m], cx, cy, cz, gx, gy, gz);
        if (d < best) { best = d; nx = gx; ny = gy; nz = gz; }
      }
      clear = best - sp.r;
      const float pen = sp.r + margin - best;
      if (pen > 0.0f && n_prims > 0) {
        my_cost += pen;
        // d cost / d q_i = -n . (z_i x (c - o_i)) for the joints upstream of the sphere's link
        const int nj = sp.link >= 7 ? 7 : sp.link;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          if (i < nj) {
            const Frame& f = s_frames[w][i];
            const float zx = f.m[2], zy = f.m[6], zz = f.m[10];
            const float rx = cx - f.m[3], ry = cy - f.m[7], rz = cz - f.m[11];
            g[i] = -(nx * (zy * rz - zz * ry) + ny * (zz * rx - zx * rz) + nz * (zx * ry - zy * rx));
          }
        }
      }
    }
    // the two warps of a waypoint: shuffle reductions, then meet in shared memory
    const float cmin = warp_min(clear);
    float gs[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) gs[i] = warp_sum(g[i]);
    if (lane == 0 && w < n) {
      s_clear[w][warp & 1] = cmin;
#pragma unroll
      for (int i = 0; i < 7; ++i) s_grad[w][warp & 1][i] = gs[i];
    }
  }
  my_cost = warp_sum(my_cost);
  if (lane == 0) s_cost[warp] = my_cost;
  __syncthreads();
  if (tid == 0 && cost) {
    float c = 0.0f;
#pragma unroll
    for (int i = 0; i < kSdfThreads / 32; ++i) c += s_cost[i];
    cost[row] = c;
  }
  if (tid < n) {
    if (clearance) clearance[(size_t)row * n + tid] = fminf(s_clear[tid][0], s_clear[tid][1]);
    if (grad) {
#pragma unroll
      for (int j = 0; j < 7; ++j) grad[((size_t)row * 7 + j) * n + tid] = s_grad[tid][0][j] + s_grad[tid][1][j];
    }
  }
}

// Point-cloud variant: clearance[row, w] = min over (sphere, point) of |centre - p| - radius.  One CTA per row; the
// sphere centres of the row live in registers (a thread owns every 256th (waypoint, sphere) item), the cloud streams
// through shared memory in tiles that every thread scans (broadcast reads): ~8 FLOP per (item, point).
constexpr int kCloudTile = 512;
constexpr int kCloudMaxWp = 50;
constexpr int kCloudItems = 12;      // ceil(50 * 59 / 256)

__global__ void __launch_bounds__(kSdfThreads) sdf_cloud_kernel(const float* __restrict__ q, int n, int rows,
                                                                const float4* __restrict__ pts, int n_pts,
                                                                float* __restrict__ clearance) {
  __shared__ Frame s_frames[kCloudMaxWp][10];
  __shared__ float4 s_pts[kCloudTile];
  __shared__ float s_item[kCloudMaxWp * kSdfSpheres];
  const int row = blockIdx.x, tid = threadIdx.x;
  if (tid < n) {
    float qq[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) qq[j] = q[((size_t)row * 7 + j) * n + tid];
    franka_fk(qq, s_frames[tid]);
  }
  __syncthreads();
  const int n_items = n * kSdfSpheres;
  float cx[kCloudItems], cy[kCloudItems], cz[kCloudItems], best[kCloudItems];
#pragma unroll
  for (int k = 0; k < kCloudItems; ++k) {
    const int it = tid + k * kSdfThreads;
    best[k] = 3.0e38f;
    cx[k] = cy[k] = cz[k] = 0.0f;
    if (it < n_items) {
      const int w = it / kSdfSpheres, s = it - w * kSdfSpheres;
      const SdfSphere sp = c_spheres[s];
      cx[k] = sp.x; cy[k] = sp.y; cz[k] = sp.z;
      if (sp.link > 0) {
        const Frame& f = s_frames[w][sp.link - 1];
        cx[k] = f.m[0] * sp.x + f.m[1] * sp.y + f.m[2] * sp.z + f.m[3];
        cy[k] = f.m[4] * sp.x + f.m[5] * sp.y + f.m[6] * sp.z + f.m[7];
        cz[k] = f.m[8] * sp.x + f.m[9] * sp.y + f.m[10] * sp.z + f.m[11];
      }
    }
  }
  for (int p0 = 0; p0 < n_pts; p0 += kCloudTile) {
    const int np = min(kCloudTile, n_pts - p0);
    __syncthreads();
    for (int i = tid; i < np; i += kSdfThreads) s_pts[i] = pts[p0 + i];
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < np; ++i) {
      const float4 p = s_pts[i];
#pragma unroll
      for (int k = 0; k < kCloudItems; ++k) {
        const float dx = cx[k] - p.x, dy = cy[k] - p.y, dz = cz[k] - p.z;
        best[k] = fminf(best[k], fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kCloudItems; ++k) {
    const int it = tid + k * kSdfThreads;
    if (it < n_items) s_item[it] = sqrtf(best[k]) - c_spheres[it % kSdfSpheres].r;
  }
  __syncthreads();
  // nearest clearance per waypoint: a warp per waypoint, shuffle min over its 59 spheres
  const int lane = tid & 31, warp = tid >> 5;
  for (int w = warp; w < n; w += kSdfThreads / 32) {
    float v = 3.0e38f;
    for (int s = lane; s < kSdfSpheres; s += 32) v = fminf(v, s_item[w * kSdfSpheres + s]);
    v = warp_min(v);
    if (lane == 0) clearance[(size_t)row * n + w] = v;
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
// World -> primitive rotation exactly as mpinets/geometry.py:185-214 builds it from the conjugate quaternion --
// including its (2,1) element `yz - wx` (a proper rotation has `yz + wx` there, geometry.py:209-210), so that results
// are the reference's for every quaternion; yaw-only obstacles (x = y = 0) are unaffected by the quirk.  The kernel's
// gradient uses the transpose of this matrix, which is the chain rule for any linear map.
static void quat_xyzw_to_rinv(const double* qv, float* rinv) {
  double x = qv[0], y = qv[1], z = qv[2], w = qv[3];
  const double nrm = std::sqrt(x * x + y * y + z * z + w * w);
  x = -x / nrm; y = -y / nrm; z = -z / nrm; w /= nrm;
  const double xx = 2 * x * x, yy = 2 * y * y, zz = 2 * z * z;
  const double wx = 2 * w * x, wy = 2 * w * y, wz = 2 * w * z, xy = 2 * x * y, xz = 2 * x * z, yz = 2 * y * z;
  const double R[9] = {1 - yy - zz, xy - wz, xz + wy,
                       xy + wz, 1 - xx - zz, yz - wx,
                       xz - wy, yz - wx, 1 - xx - yy};
  for (int i = 0; i < 9; ++i) rinv[i] = (float)R[i];
}

// the sphere table is a __constant__ symbol: per device
static unsigned long long g_sdf_tables_ready = 0;
static int sdf_upload_tables() {
  if (!once_per_device(&g_sdf_tables_ready)) return 0;
  if (cudaMemcpyToSymbol(c_spheres, kSphereTable, sizeof(kSphereTable)) != cudaSuccess) {
    once_per_device_failed(&g_sdf_tables_ready);
    set_error("sdf: uploading the collision-sphere table failed");
    return 1;
  }
  return 0;
}

int sdf_scene_create(const double* boxes_h, int n_boxes, const double* cyls_h, int n_cyls, SdfScene** out) {
  EDMP_REQUIRE(n_boxes >= 0 && n_cyls >= 0 && n_boxes + n_cyls <= kSdfMaxPrims, "at most 64 primitives");
  EDMP_REQUIRE((n_boxes == 0 || boxes_h) && (n_cyls == 0 || cyls_h), "null primitive array");
  if (sdf_upload_tables()) return 1;
  std::vector<SdfPrim> prims((size_t)n_boxes + n_cyls);
  for (int i = 0; i < n_boxes; ++i) {          // [xyz, quaternion xyzw, dims]  (the reference's obstacle_config rows)
    const double* b = boxes_h + (size_t)i * 10;
    SdfPrim& p = prims[i];
    quat_xyzw_to_rinv(b + 3, p.rinv);
    for (int k = 0; k < 3; ++k) { p.c[k] = (float)b[k]; p.h[k] = (float)(0.5 * b[7 + k]); }
    p.kind = 0.0f;
  }
  for (int i = 0; i < n_cyls; ++i) {           // [xyz, quaternion xyzw, radius, height]
    const double* c = cyls_h + (size_t)i * 9;
    SdfPrim& p = prims[n_boxes + i];
    quat_xyzw_to_rinv(c + 3, p.rinv);
    for (int k = 0; k < 3; ++k) p.c[k] = (float)c[k];
    p.h[0] = (float)c[7]; p.h[1] = (float)(0.5 * c[8]); p.h[2] = 0.0f;
    p.kind = 1.0f;
  }
  SdfScene* s = new SdfScene();
  s->n_prims = n_boxes + n_cyls;
  if (s->n_prims > 0) {
    if (cudaMalloc(&s->prims, prims.size() * sizeof(SdfPrim)) != cudaSuccess) { delete s; set_error("sdf_scene_create: allocation failed"); return 1; }
    if (cudaMemcpy(s->prims, prims.data(), prims.size() * sizeof(SdfPrim), cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaFree(s->prims); delete s; set_error("sdf_scene_create: upload failed"); return 1;
    }
  }
  *out = s;
  return 0;
}

void sdf_scene_destroy(SdfScene* s) {
  if (!s) return;
  if (s->prims) cudaFree(s->prims);
  delete s;
}

int sdf_guide_launch(SdfScene* s, const float* q_d, int n, int rows, float margin, float* cost_d, float* grad_d,
                     float* clearance_d, cudaStream_t st) {
  EDMP_REQUIRE(rows > 0 && n > 0 && n <= kSdfMaxWp, "1..64 waypoints per row");
  sdf_guide_kernel<<<rows, kSdfThreads, 0, st>>>(q_d, n, rows, s->prims, s->n_prims, margin, cost_d, grad_d, clearance_d);
  EDMP_CK(cudaGetLastError());
  return 0;
}

int sdf_cloud_launch(const float* q_d, int n, int rows, const float* points_d, int n_points, float* clearance_d,
                     cudaStream_t st) {
  EDMP_REQUIRE(rows > 0 && n > 0 && n <= 50 && n_points > 0, "1..50 waypoints per row and a non-empty cloud");
  if (sdf_upload_tables()) return 1;
  sdf_cloud_kernel<<<rows, kSdfThreads, 0, st>>>(q_d, n, rows, reinterpret_cast<const float4*>(points_d), n_points, clearance_d);
  EDMP_CK(cudaGetLastError());
  return 0;
}

}  // namespace edmp
