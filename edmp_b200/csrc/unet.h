// TemporalUNet engine interface (unet.cu) used by sampler.cu / api.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace edmp {
struct UNet;
size_t unet_param_count(const int* dims, int n_dims);
int unet_create(const float* params, size_t n_params, const int* dims, int n_dims, int precision,
                int max_rows, UNet** out);
void unet_destroy(UNet* u);
// packed-weight blob: engine from a state_dict while recording every packed device image;
// read the blob out; engine straight from a blob (no repacking; rc 4 = stale / mismatching blob -> repack)
int unet_pack(const float* params, size_t n_params, const int* dims, int n_dims, int precision, int max_rows, UNet** out);
int unet_blob_layout_version();
size_t unet_blob_bytes(const UNet* u);
int unet_blob_read(UNet* u, void* dst, size_t cap);
int unet_create_from_blob(const void* blob, size_t bytes, int max_rows, UNet** out);
// eps[rows,7,50] = model(x[rows,7,50], t)
int unet_forward(UNet* u, const float* x, int t, int rows, float* eps, cudaStream_t st);
// The sampler's fused per-step tail writes the network input of the NEXT step straight into the first layer's operand
// image (position-major, 16 channels, hi / lo halves): unet_input_image hands out that image (el: 1 = BF16, 2 = IEEE half;
// returns false when the engine has no such image, e.g. the fp32 mode), unet_forward_packed runs a forward whose input
// is already there (the pack launch is skipped).
bool unet_input_image(UNet* u, void** hi, void** lo, int* el, unsigned** range_flag);
int unet_forward_packed(UNet* u, int t, int rows, float* eps, cudaStream_t st);
int unet_read_activation(UNet* u, const char* name, int rows, float* out, int* C, int* L, cudaStream_t st);
int unet_profile(UNet* u, const float* x, int t, int rows, int iters, float* ms, double* macs, float* eps,
                 cudaStream_t st);
const char* unet_op_name(const UNet* u, int i);
const char* unet_op_kernel(const UNet* u, int i);
int unet_tc_trace(UNet* u, int op, int rows, long long* out_h, int max_ctas, int* n_ctas, cudaStream_t st);
int unet_precision(const UNet* u);
int unet_range_status(UNet* u, int* overflow, cudaStream_t st);
int unet_launches(const UNet* u);
}  // namespace edmp
