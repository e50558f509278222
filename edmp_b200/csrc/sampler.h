#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace edmp {
struct Sampler;
struct UNet;
struct Scene;
int sampler_create(int T, double thresh, int max_rows, Sampler** out);
void sampler_destroy(Sampler* s);
int sampler_schedule(const Sampler* s, double* b, double* a, double* ab);
long long sampler_last_launches(const Sampler* s);
void sampler_set_condition(Sampler* s, bool condition);
int sample_guided(Sampler* s, UNet* u, Scene* scene, double* x, const double* start, const double* goal,
                  const double* noise, uint64_t seed, int rows, int t_start, int t_stop, float* final_cost,
                  cudaStream_t st);
int sample_guided_host(Sampler* s, UNet* u, Scene* scene, double* x_h, const double* start,
                       const double* goal, uint64_t seed, int rows, float* cost_h, cudaStream_t st);
}  // namespace edmp
