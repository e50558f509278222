// Trajectory metrics for whole ensembles on the device:
//   * end-effector transform = product of the 10 DH matrices (reference lib/guide.py:100-116, get_tf_mat :45-72),
//     float32 like the reference;
//   * joint-space / end-effector path length (reference lib/metrics.py:33-45);
//   * SPARC smoothness (spectral arc length, reference lib/metrics.py:47-130) of the joint-space and end-effector
//     speed profiles (lib/metrics.py:11-31).
// One CTA per trajectory row.  The spectrum is a direct DFT of the (short: horizon - 1 = 49 samples) speed profile at
// the nfft = 2^(ceil(log2 m) + padlevel) zero-padded frequencies -- 49 x 513 complex MACs per profile from a twiddle
// table in shared memory; a radix-2 FFT of a 95 %-zero signal would do more work.  float64 throughout (the metrics
// are reported numbers, not hot-path arithmetic); the end-effector FK stays float32 because the reference's is.
#include "metrics.h"

#include <cmath>

#include "common.cuh"

namespace edmp {

namespace {

constexpr int kMThreads = 128;
constexpr int kMaxWp = 64;         // waypoints per trajectory
constexpr int kMaxNfft = 8192;     // 24 B of shared memory per bin

// static DH table (a, d, alpha, theta offset) of the reference, lib/guide.py:29-38; rows 0..6 take the joint angle
__constant__ float c_mdh[10][4] = {
    {0.0f, 0.333f, 0.0f, 0.0f},
    {0.0f, 0.0f, -1.57079632679489661923f, 0.0f},
    {0.0f, 0.316f, 1.57079632679489661923f, 0.0f},
    {0.0825f, 0.0f, 1.57079632679489661923f, 0.0f},
    {-0.0825f, 0.384f, -1.57079632679489661923f, 0.0f},
    {0.0f, 0.0f, 1.57079632679489661923f, 0.0f},
    {0.088f, 0.0f, 1.57079632679489661923f, 0.0f},
    {0.0f, 0.107f, 0.0f, 0.0f},
    {0.0f, 0.0f, 0.0f, -0.78539816339744830962f},
    {0.0f, 0.1034f, 0.0f, 0.0f}};

// T <- T * DH(a, d, alpha, q), 3x4 row-major (the bottom row stays 0 0 0 1)
__device__ __forceinline__ void dh_mul(float T[12], float a, float d, float alpha, float q) {
  float sq, cq, sa, ca;
  sincosf(q, &sq, &cq);
  sincosf(alpha, &sa, &ca);
  const float M[12] = {cq, -sq, 0.0f, a, sq * ca, cq * ca, -sa, -sa * d, sq * sa, cq * sa, ca, ca * d};
  float R[12];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float s = T[r * 4 + 0] * M[c];
      s = fmaf(T[r * 4 + 1], M[4 + c], s);
      s = fmaf(T[r * 4 + 2], M[8 + c], s);
      if (c == 3) s += T[r * 4 + 3];
      R[r * 4 + c] = s;
    }
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) T[i] = R[i];
}

__device__ void ee_transform(const float q[7], float T[12]) {
#pragma unroll
  for (int i = 0; i < 12; ++i) T[i] = (i == 0 || i == 5 || i == 10) ? 1.0f : 0.0f;
#pragma unroll 1
  for (int i = 0; i < 10; ++i) dh_mul(T, c_mdh[i][0], c_mdh[i][1], c_mdh[i][2], i < 7 ? q[i] : c_mdh[i][3]);
}

__global__ void ee_transform_kernel(const float* __restrict__ q, int total, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (row, waypoint)
  if (i >= total) return;
  float qq[7], T[12];
#pragma unroll
  for (int j = 0; j < 7; ++j) qq[j] = q[(size_t)i * 7 + j];
  ee_transform(qq, T);
  float* o = out + (size_t)i * 16;
#pragma unroll
  for (int k = 0; k < 12; ++k) o[k] = T[k];
  o[12] = 0.0f; o[13] = 0.0f; o[14] = 0.0f; o[15] = 1.0f;
}

__device__ __forceinline__ double block_reduce(double v, int op, double* s_red) {   // op 0: sum, 1: max, 2: min
  for (int o = 16; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = op == 0 ? v + w : (op == 1 ? fmax(v, w) : fmin(v, w));
  }
  __syncthreads();   // s_red may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = s_red[0];
  for (int w = 1; w < kMThreads / 32; ++w) r = op == 0 ? r + s_red[w] : (op == 1 ? fmax(r, s_red[w]) : fmin(r, s_red[w]));
  return r;
}

// Spectral arc length of the m-sample profile v (shared memory) sampled at fs; lib/metrics.py:86-130.
//   s_tw: [nfft] twiddles e^{-2 pi i j / nfft}; s_mf: [nfft] scratch for the normalised magnitude spectrum.
// Returns the metric; sel[0..1] = first / last selected bin (or -1).  All threads of the CTA call it.
__device__ double sparc_block(const double* v, int m, double fs, int nfft, double fc, double amp_th, const double2* s_tw,
                              double* s_mf, double* s_red, int* sel) {
  // np.allclose(movement, 0): |x| <= 1e-8 everywhere -> 0 (lib/metrics.py:86-88)
  double amax = 0.0;
  for (int i = threadIdx.x; i < m; i += kMThreads) amax = fmax(amax, fabs(v[i]));
  amax = block_reduce(amax, 1, s_red);
  if (!(amax > 1e-8)) {
    if (sel) { sel[0] = -1; sel[1] = -1; }
    for (int k = threadIdx.x; k < nfft; k += kMThreads) s_mf[k] = 0.0;
    __syncthreads();
    return 0.0;
  }
  // magnitude spectrum of the zero-padded signal; real input: |X[nfft - k]| = |X[k]|
  double mx = 0.0;
  for (int k = threadIdx.x; k <= nfft / 2; k += kMThreads) {
    double re = 0.0, im = 0.0;
    int idx = 0;
    for (int i = 0; i < m; ++i) {
      const double2 w = s_tw[idx];
      re = fma(v[i], w.x, re);
      im = fma(v[i], w.y, im);
      idx = (idx + k) & (nfft - 1);
    }
    const double mag = sqrt(re * re + im * im);
    s_mf[k] = mag;
    if (k > 0 && k < nfft / 2) s_mf[nfft - k] = mag;
    mx = fmax(mx, mag);
  }
  mx = block_reduce(mx, 1, s_red);
  const double step = fs / (double)nfft;            // f = np.arange(0, fs, fs / nfft): f[k] = k * step
  // low-pass selection f <= fc, then the first / last bin at or above the amplitude threshold (:100-113)
  int first = nfft, last = -1;
  for (int k = threadIdx.x; k < nfft; k += kMThreads) {
    const double mf = s_mf[k] / mx;
    s_mf[k] = mf;
    if ((double)k * step <= fc && mf >= amp_th) { first = min(first, k); last = max(last, k); }
  }
  first = (int)block_reduce((double)first, 2, s_red);
  last = (int)block_reduce((double)last, 1, s_red);
  if (sel) { sel[0] = last >= 0 ? first : -1; sel[1] = last; }
  double acc = 0.0;
  if (last > first) {
    // the reference slices range(first, last + 1) out of the low-passed arrays: every bin in between has f <= fc too
    const double span = (double)last * step - (double)first * step;
    for (int k = first + threadIdx.x; k < last; k += kMThreads) {
      const double df = ((double)(k + 1) * step - (double)k * step) / span;
      const double dm = s_mf[k + 1] - s_mf[k];
      acc += sqrt(df * df + dm * dm);
    }
  }
  acc = block_reduce(acc, 0, s_red);
  return -acc;
}

__device__ void fill_twiddles(double2* s_tw, int nfft) {
  for (int j = threadIdx.x; j < nfft; j += kMThreads) {
    double s, c;
    sincospi(-2.0 * (double)j / (double)nfft, &s, &c);
    s_tw[j] = make_double2(c, s);
  }
}

// traj: float64 [rows][7][n] (the sampler's output layout).  out: [rows][4] = joint path length, end-effector path
// length, joint SPARC, end-effector SPARC.  spec (optional): [rows][2][nfft] normalised magnitude spectra;
// sel (optional): [rows][2][2] first / last selected bins.
__global__ void __launch_bounds__(kMThreads) metrics_kernel(const double* __restrict__ traj, int n, double dt, int nfft,
                                                            double fc, double amp_th, double* __restrict__ out,
                                                            double* __restrict__ spec, int* __restrict__ sel) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  double2* s_tw = reinterpret_cast<double2*>(sm_raw);          // [nfft]
  double* s_mf = reinterpret_cast<double*>(s_tw + nfft);        // [nfft]
  __shared__ double s_q[7][kMaxWp];
  __shared__ float s_p[kMaxWp][3];
  __shared__ double s_vj[kMaxWp], s_ve[kMaxWp];
  __shared__ double s_red[kMThreads / 32];
  __shared__ int s_sel[2];
  const int row = blockIdx.x;
  const double* x = traj + (size_t)row * 7 * n;
  for (int i = threadIdx.x; i < 7 * n; i += kMThreads) s_q[i / n][i % n] = x[i];
  fill_twiddles(s_tw, nfft);
  __syncthreads();
  if (threadIdx.x < n) {
    // lib/metrics.py:17-18, :35-36: the joints go through float32 before the FK
    float q[7], T[12];
#pragma unroll
    for (int j = 0; j < 7; ++j) q[j] = (float)s_q[j][threadIdx.x];
    ee_transform(q, T);
    s_p[threadIdx.x][0] = T[3]; s_p[threadIdx.x][1] = T[7]; s_p[threadIdx.x][2] = T[11];
  }
  __syncthreads();
  const int m = n - 1;
  double lj = 0.0, le = 0.0;
  if (threadIdx.x < m) {
    const int i = threadIdx.x;
    double sj = 0.0;
#pragma unroll
    for (int j = 0; j < 7; ++j) { const double d = s_q[j][i + 1] - s_q[j][i]; sj = fma(d, d, sj); }
    lj = sqrt(sj);
    // end-effector positions are float32 in the reference (np.diff / norm of a float32 array)
    const float dx = s_p[i + 1][0] - s_p[i][0], dy = s_p[i + 1][1] - s_p[i][1], dz = s_p[i + 1][2] - s_p[i][2];
    le = (double)sqrtf(dx * dx + dy * dy + dz * dz);
    // speed profiles (lib/metrics.py:24-28): norm(diff / dt)
    double vj = 0.0;
#pragma unroll
    for (int j = 0; j < 7; ++j) { const double d = (s_q[j][i + 1] - s_q[j][i]) / dt; vj = fma(d, d, vj); }
    s_vj[i] = sqrt(vj);
    const float fdt = (float)dt;
    const float ex = dx / fdt, ey = dy / fdt, ez = dz / fdt;
    s_ve[i] = (double)sqrtf(ex * ex + ey * ey + ez * ez);
  }
  lj = block_reduce(lj, 0, s_red);
  le = block_reduce(le, 0, s_red);
  __syncthreads();
  const double fs = 1.0 / dt;
  double sal[2];
  for (int which = 0; which < 2; ++which) {
    sal[which] = sparc_block(which ? s_ve : s_vj, m, fs, nfft, fc, amp_th, s_tw, s_mf, s_red, threadIdx.x == 0 ? s_sel : nullptr);
    __syncthreads();
    if (spec) for (int k = threadIdx.x; k < nfft; k += kMThreads) spec[((size_t)row * 2 + which) * nfft + k] = s_mf[k];
    if (sel && threadIdx.x == 0) { sel[(row * 2 + which) * 2] = s_sel[0]; sel[(row * 2 + which) * 2 + 1] = s_sel[1]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[(size_t)row * 4 + 0] = lj;
    out[(size_t)row * 4 + 1] = le;
    out[(size_t)row * 4 + 2] = sal[0];
    out[(size_t)row * 4 + 3] = sal[1];
  }
}

// SPARC of raw speed profiles: movement float64 [rows][m]
__global__ void __launch_bounds__(kMThreads) sparc_kernel(const double* __restrict__ movement, int m, double fs, int nfft,
                                                          double fc, double amp_th, double* __restrict__ sal,
                                                          double* __restrict__ spec, int* __restrict__ sel) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  double2* s_tw = reinterpret_cast<double2*>(sm_raw);
  double* s_mf = reinterpret_cast<double*>(s_tw + nfft);
  double* s_v = s_mf + nfft;                                    // [m]
  __shared__ double s_red[kMThreads / 32];
  __shared__ int s_sel[2];
  const int row = blockIdx.x;
  for (int i = threadIdx.x; i < m; i += kMThreads) s_v[i] = movement[(size_t)row * m + i];
  fill_twiddles(s_tw, nfft);
  __syncthreads();
  const double r = sparc_block(s_v, m, fs, nfft, fc, amp_th, s_tw, s_mf, s_red, threadIdx.x == 0 ? s_sel : nullptr);
  __syncthreads();
  if (spec) for (int k = threadIdx.x; k < nfft; k += kMThreads) spec[(size_t)row * nfft + k] = s_mf[k];
  if (threadIdx.x == 0) {
    sal[row] = r;
    if (sel) { sel[row * 2] = s_sel[0]; sel[row * 2 + 1] = s_sel[1]; }
  }
}

}  // namespace

int metrics_nfft(int m, int padlevel) {
  // int(pow(2, ceil(log2(len)) + padlevel)), lib/metrics.py:90
  if (m < 1 || padlevel < 0) return 0;
  int e = 0;
  while ((1 << e) < m) ++e;
  if (e + padlevel > 20) return 0;
  return 1 << (e + padlevel);
}

int ee_transform_launch(const float* q_d, int rows, int n, float* T_d, cudaStream_t st) {
  EDMP_REQUIRE(rows > 0 && n > 0, "rows and n must be positive");
  const int total = rows * n;
  ee_transform_kernel<<<(total + 127) / 128, 128, 0, st>>>(q_d, total, T_d);
  EDMP_CK(cudaGetLastError());
  return 0;
}

int trajectory_metrics_launch(const double* traj_d, int rows, int n, double dt, int padlevel, double fc, double amp_th,
                              double* out_d, double* spec_d, int* sel_d, cudaStream_t st) {
  EDMP_REQUIRE(rows > 0, "rows must be positive");
  EDMP_REQUIRE(n >= 2 && n <= kMaxWp, "n (waypoints) must be in 2..64");
  EDMP_REQUIRE(dt > 0.0, "dt must be positive");
  const int nfft = metrics_nfft(n - 1, padlevel);
  EDMP_REQUIRE(nfft >= 2 && nfft <= kMaxNfft, "2^(ceil(log2(n-1)) + padlevel) must be in 2..8192");
  const size_t smem = (size_t)nfft * (sizeof(double2) + sizeof(double));
  EDMP_CK(cudaFuncSetAttribute(metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  metrics_kernel<<<rows, kMThreads, smem, st>>>(traj_d, n, dt, nfft, fc, amp_th, out_d, spec_d, sel_d);
  EDMP_CK(cudaGetLastError());
  return 0;
}

int sparc_launch(const double* movement_d, int rows, int m, double fs, int padlevel, double fc, double amp_th,
                 double* sal_d, double* spec_d, int* sel_d, cudaStream_t st) {
  EDMP_REQUIRE(rows > 0, "rows must be positive");
  EDMP_REQUIRE(m >= 1 && m <= 1024, "profile length must be in 1..1024");
  EDMP_REQUIRE(fs > 0.0, "fs must be positive");
  const int nfft = metrics_nfft(m, padlevel);
  EDMP_REQUIRE(nfft >= 2 && nfft <= kMaxNfft, "2^(ceil(log2(m)) + padlevel) must be in 2..8192");
  const size_t smem = (size_t)nfft * (sizeof(double2) + sizeof(double)) + (size_t)m * sizeof(double);
  EDMP_CK(cudaFuncSetAttribute(sparc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sparc_kernel<<<rows, kMThreads, smem, st>>>(movement_d, m, fs, nfft, fc, amp_th, sal_d, spec_d, sel_d);
  EDMP_CK(cudaGetLastError());
  return 0;
}

}  // namespace edmp
