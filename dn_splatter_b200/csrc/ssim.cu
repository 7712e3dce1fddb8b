// Fused SSIM (value + gradient) for the photometric term of SplatfactoModel.get_loss_dict [EXT nerfstudio 1.1.3], which
// dn-splatter evaluates with torchmetrics' StructuralSimilarityIndexMeasure(data_range=1.0, kernel_size=11)
// (/root/reference/dn_splatter/dn_model.py:180, loss assembled at :624-628).  SURVEY.md §8f-3 ("next" row).
//
// STATUS: written in round 1 after the GPU budget was spent — compiled, NOT yet validated on a GPU, and therefore not
// used unless DNSplatterModelConfig.fused_ssim is set (the torch implementation in dn_model.ssim() stays the default and
// is the reference this kernel must match).
//
// torchmetrics pads by reflection, filters with an 11x11 Gaussian (sigma 1.5) and then CROPS the padding away before
// taking the mean, so only windows that lie fully inside the image contribute: mean over the (H-10)x(W-10) interior of
//   S = ((2 mx my + C1)(2 sxy + C2)) / ((mx^2 + my^2 + C1)(sx + sy + C2)),   C1 = 0.01^2, C2 = 0.03^2.
// Forward: one CTA per 16x16 output tile and channel; 26x26 halo tile of pred / gt in shared memory, separable 11-tap
// filter of {x, y, x^2, y^2, xy}; writes the three partial-derivative maps dS/dmx, dS/dExx, dS/dExy and block-reduces
// the SSIM sum.  Backward: the same separable filter applied to those maps (the transposed correlation of a symmetric
// kernel), v_x = v * (F[dS/dmx] + 2 x F[dS/dExx] + y F[dS/dExy]) / count.
#include "common.cuh"

namespace {

constexpr int SS_R = 5;            // window radius
constexpr int SS_T = 16;           // output tile
constexpr int SS_H = SS_T + 2 * SS_R;  // halo tile edge (26)

// normalised 1-D Gaussian, sigma 1.5: exp(-d^2 / 4.5) / sum, evaluated in fp32 exactly as torchmetrics' _gaussian does
// (statically initialised: no host state, valid on every device of the process, nothing to do under graph capture)
__constant__ float c_win[11] = {1.028380357e-03f, 7.598758209e-03f, 3.600077331e-02f, 1.093606874e-01f, 2.130055279e-01f,
                                2.660117149e-01f, 2.130055279e-01f, 1.093606874e-01f, 3.600077331e-02f, 7.598758209e-03f,
                                1.028380357e-03f};

__device__ __forceinline__ float ld_img(const float* __restrict__ img, int H, int W, int C, int i, int j, int c) {
  return (i >= 0 && i < H && j >= 0 && j < W) ? img[((size_t)i * W + j) * C + c] : 0.f;
}

// Separable filter of NQ quantities held in s_in[q][SS_H][SS_H]; result for thread (ty,tx) in out[q].
template <int NQ>
__device__ __forceinline__ void separable(float (*s_in)[SS_H][SS_H], float (*s_mid)[SS_H][SS_T], int tid, int ty, int tx,
                                          float out[NQ]) {
  for (int e = tid; e < SS_H * SS_T; e += SS_T * SS_T) {  // horizontal pass: rows 0..25, cols 0..15
    const int r = e / SS_T, c = e % SS_T;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) acc += c_win[k] * s_in[q][r][c + k];
      s_mid[q][r][c] = acc;
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NQ; ++q) {  // vertical pass
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) acc += c_win[k] * s_mid[q][ty + k][tx];
    out[q] = acc;
  }
}

__global__ void __launch_bounds__(SS_T* SS_T) ssim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W,
                                                              int C, float* __restrict__ dmaps, float* __restrict__ sum_out) {
  __shared__ float s_in[5][SS_H][SS_H];
  __shared__ float s_mid[5][SS_H][SS_T];
  __shared__ float red[8];
  const int tid = threadIdx.y * SS_T + threadIdx.x, c = blockIdx.z;
  const int i0 = blockIdx.y * SS_T - SS_R, j0 = blockIdx.x * SS_T - SS_R;
  for (int e = tid; e < SS_H * SS_H; e += SS_T * SS_T) {
    const int r = e / SS_H, q = e % SS_H;
    const float xv = ld_img(x, H, W, C, i0 + r, j0 + q, c), yv = ld_img(y, H, W, C, i0 + r, j0 + q, c);
    s_in[0][r][q] = xv; s_in[1][r][q] = yv; s_in[2][r][q] = xv * xv; s_in[3][r][q] = yv * yv; s_in[4][r][q] = xv * yv;
  }
  __syncthreads();
  float f[5];
  separable<5>(s_in, s_mid, tid, threadIdx.y, threadIdx.x, f);
  const int i = blockIdx.y * SS_T + threadIdx.y, j = blockIdx.x * SS_T + threadIdx.x;
  const bool interior = (i >= SS_R) && (i < H - SS_R) && (j >= SS_R) && (j < W - SS_R);
  float s = 0.f, d_mu = 0.f, d_xx = 0.f, d_xy = 0.f;
  if (interior) {
    const float C1 = 0.0001f, C2 = 0.0009f;
    const float mx = f[0], my = f[1];
    const float sx = f[2] - mx * mx, sy = f[3] - my * my, sxy = f[4] - mx * my;
    const float A1 = 2.f * mx * my + C1, A2 = 2.f * sxy + C2, B1 = mx * mx + my * my + C1, B2 = sx + sy + C2;
    const float inv = 1.0f / (B1 * B2);
    s = A1 * A2 * inv;
    d_xx = -s / B2;
    d_xy = 2.f * A1 * inv;
    d_mu = 2.f * my * (A2 - A1) * inv - 2.f * mx * s / B1 + 2.f * mx * s / B2;
  }
  if (i < H && j < W) {
    const size_t p = ((size_t)i * W + j) * C + c, n = (size_t)H * W * C;
    dmaps[p] = d_mu; dmaps[n + p] = d_xx; dmaps[2 * n + p] = d_xy;
  }
  s = warp_sum(s);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    if (t != 0.f) atomicAdd(sum_out, t);
  }
}

__global__ void __launch_bounds__(SS_T* SS_T) ssim_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W,
                                                              int C, const float* __restrict__ dmaps, const float* v_mean,
                                                              float* __restrict__ v_x) {
  __shared__ float s_in[3][SS_H][SS_H];
  __shared__ float s_mid[3][SS_H][SS_T];
  const int tid = threadIdx.y * SS_T + threadIdx.x, c = blockIdx.z;
  const int i0 = blockIdx.y * SS_T - SS_R, j0 = blockIdx.x * SS_T - SS_R;
  const size_t n = (size_t)H * W * C;
  for (int e = tid; e < SS_H * SS_H; e += SS_T * SS_T) {
    const int r = e / SS_H, q = e % SS_H;
    const int i = i0 + r, j = j0 + q;
    // the maps are zero outside the interior by construction (forward wrote zeros there); outside the image: zero
    s_in[0][r][q] = ld_img(dmaps, H, W, C, i, j, c);
    s_in[1][r][q] = ld_img(dmaps + n, H, W, C, i, j, c);
    s_in[2][r][q] = ld_img(dmaps + 2 * n, H, W, C, i, j, c);
  }
  __syncthreads();
  float f[3];
  separable<3>(s_in, s_mid, tid, threadIdx.y, threadIdx.x, f);
  const int i = blockIdx.y * SS_T + threadIdx.y, j = blockIdx.x * SS_T + threadIdx.x;
  if (i < H && j < W) {
    const size_t p = ((size_t)i * W + j) * C + c;
    const float count = (float)(H - 2 * SS_R) * (float)(W - 2 * SS_R) * (float)C;
    const float g = (v_mean ? __ldg(v_mean) : 1.0f) / count;
    v_x[p] = g * (f[0] + 2.f * x[p] * f[1] + y[p] * f[2]);
  }
}

}  // namespace

// pred / gt: [H,W,C] fp32.  dmaps: [3,H,W,C] scratch kept for the backward.  *mean_out (zeroed by the call) receives the
// SUM of the SSIM map over the interior; the caller divides by (H-10)(W-10)C.
extern "C" int dnr_ssim_fwd(const float* pred, const float* gt, int32_t H, int32_t W, int32_t C, float* dmaps, float* sum_out,
                            void* stream) {
  if (!pred || !gt || !dmaps || !sum_out) return DNR_E_NULL;
  if (H <= 2 * SS_R || W <= 2 * SS_R || C <= 0) return DNR_E_SIZE;
  cudaStream_t s = (cudaStream_t)stream;
  DNR_CUDA(cudaMemsetAsync(sum_out, 0, sizeof(float), s));
  const dim3 grid((W + SS_T - 1) / SS_T, (H + SS_T - 1) / SS_T, C), block(SS_T, SS_T);
  ssim_fwd_kernel<<<grid, block, 0, s>>>(pred, gt, H, W, C, dmaps, sum_out);
  DNR_CHECK_LAUNCH();
  return 0;
}

// v_pred[H,W,C] = (*v_mean or 1) * d(mean SSIM)/d(pred).
extern "C" int dnr_ssim_bwd(const float* pred, const float* gt, int32_t H, int32_t W, int32_t C, const float* dmaps,
                            const float* v_mean, float* v_pred, void* stream) {
  if (!pred || !gt || !dmaps || !v_pred) return DNR_E_NULL;
  if (H <= 2 * SS_R || W <= 2 * SS_R || C <= 0) return DNR_E_SIZE;
  const dim3 grid((W + SS_T - 1) / SS_T, (H + SS_T - 1) / SS_T, C), block(SS_T, SS_T);
  ssim_bwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(pred, gt, H, W, C, dmaps, v_mean, v_pred);
  DNR_CHECK_LAUNCH();
  return 0;
}
