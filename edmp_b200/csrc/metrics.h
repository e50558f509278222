// Trajectory metrics on the device (metrics.cu) used by api.cu.
#pragma once
#include <cuda_runtime.h>

namespace edmp {
int metrics_nfft(int m, int padlevel);
int ee_transform_launch(const float* q_d, int rows, int n, float* T_d, cudaStream_t st);
int trajectory_metrics_launch(const double* traj_d, int rows, int n, double dt, int padlevel, double fc, double amp_th,
                              double* out_d, double* spec_d, int* sel_d, cudaStream_t st);
int sparc_launch(const double* movement_d, int rows, int m, double fs, int padlevel, double fc, double amp_th,
                 double* sal_d, double* spec_d, int* sel_d, cudaStream_t st);
}  // namespace edmp
