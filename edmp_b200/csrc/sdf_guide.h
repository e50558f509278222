// Sphere / signed-distance guide family (sdf_guide.cu) used by api.cu.
#pragma once
#include <cuda_runtime.h>

namespace edmp {
struct SdfScene;
int sdf_scene_create(const double* boxes_h, int n_boxes, const double* cyls_h, int n_cyls, SdfScene** out);
void sdf_scene_destroy(SdfScene* s);
int sdf_guide_launch(SdfScene* s, const float* q_d, int n, int rows, float margin, float* cost_d, float* grad_d,
                     float* clearance_d, cudaStream_t st);
int sdf_cloud_launch(const float* q_d, int n, int rows, const float* points_d, int n_points, float* clearance_d,
                     cudaStream_t st);
}  // namespace edmp
