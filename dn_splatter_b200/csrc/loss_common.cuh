// Per-pixel pieces of the DNRegularization terms shared by the loss kernels (csrc/image_ops.cu) and by the loss-gradient
// prologue of dnr_raster_bwd (csrc/raster.cu).  Reference: dn_splatter/losses.py:197-224 (EdgeAwareLogL1 weights),
// dn_splatter/dn_model.py:633 (gt image clamped below at 10/255), nerfstudio get_gt_img (uint8 / 255).
#pragma once
#include "common.cuh"

__device__ __forceinline__ float sgnf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

// exp(-mean_c |I[p] - I[q]|) of the EdgeAwareLogL1 weights (losses.py:197-224); the image is the fp32 gt_rgb map or, with
// DNR_LOSS_EDGE_FROM_IMAGE, the uint8 image scaled by 1/255 and clamped below at 10/255 (dn_model.py:633).
__device__ __forceinline__ float edge_weight(const DnrArgs& a, int p, int q) {
  float g = 0.f;
  if (a.loss_flags & DNR_LOSS_EDGE_FROM_IMAGE) {
    const uint8_t* im = (const uint8_t*)a.gt_image;
    const float inv = 1.0f / 255.0f, lo_ = 10.0f / 255.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      g += fabsf(fmaxf(__fmul_rn((float)im[p * 3 + c], inv), lo_) - fmaxf(__fmul_rn((float)im[q * 3 + c], inv), lo_));
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) g += fabsf(a.gt_rgb[p * 3 + c] - a.gt_rgb[q * 3 + c]);
  }
  return expf(-(g / 3.0f));
}

__device__ __forceinline__ float gt_normal_at(const DnrArgs& a, int idx) {
  if (a.loss_flags & DNR_LOSS_NORMAL_U8) return __fmul_rn((float)((const uint8_t*)a.gt_normal)[idx], 1.0f / 255.0f);
  return a.gt_normal[idx];
}

