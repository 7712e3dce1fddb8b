// Shared device helpers for libdnr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dnr.h"

#define DNR_TILE 16
#define DNR_ALPHA_MIN (1.0f / 255.0f)
#define DNR_ALPHA_MAX 0.999f
#define DNR_T_STOP 1e-4f

#define DNR_CHECK_LAUNCH()                         \
  do {                                             \
    cudaError_t e__ = cudaGetLastError();          \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

#define DNR_CUDA(expr)                             \
  do {                                             \
    cudaError_t e__ = (expr);                      \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

static inline int dnr_tiles_x(const DnrArgs* a) { return (a->width + DNR_TILE - 1) / DNR_TILE; }
static inline int dnr_tiles_y(const DnrArgs* a) { return (a->height + DNR_TILE - 1) / DNR_TILE; }

// Tile box of a projected Gaussian: tile_min inclusive, tile_max exclusive (gsplat isect_tiles).
// Shared by the count (project_fwd) and emit (bin_sort) kernels so both see the same integers.
__device__ __forceinline__ void dnr_tile_box(float mx, float my, int radius, int tiles_x, int tiles_y,
                                             int& x0, int& y0, int& x1, int& y1) {
  const float r = (float)radius * (1.0f / DNR_TILE);
  const float tcx = mx * (1.0f / DNR_TILE);
  const float tcy = my * (1.0f / DNR_TILE);
  // (uint32_t)floor(negative) saturates to 0 on the GPU
  x0 = min(max((int)floorf(tcx - r), 0), tiles_x);
  y0 = min(max((int)floorf(tcy - r), 0), tiles_y);
  x1 = min(max((int)ceilf(tcx + r), 0), tiles_x);
  y1 = min(max((int)ceilf(tcy + r), 0), tiles_y);
}

// log2-domain exponent of a splat at offset (dx,dy): records hold a' = -0.5*log2(e)*A, b' = -log2(e)*B,
// c' = -0.5*log2(e)*C, so alpha = opac * 2^power.  Fixed rounding order, shared by the forward and backward
// kernels so that both take the same skip/stop branches for every (pixel, Gaussian) pair.
__device__ __forceinline__ float dnr_power2(float a, float b, float c, float dx, float dy) {
  const float t1 = __fmaf_rn(a, dx, __fmul_rn(b, dy));
  const float t2 = __fmul_rn(c, dy);
  return __fmaf_rn(dx, t1, __fmul_rn(dy, t2));
}
__device__ __forceinline__ float dnr_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#define DNR_LOG2E 1.4426950408889634f
#define DNR_LN2 0.6931471805599453f
#define DNR_CULL_MARGIN 0.1f /* in sigma units: a tile is dropped only if alpha_max < e^-0.1 / 255 */

// ---- mbarrier / bulk-copy (TMA, 1-D) primitives -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk async copy (SASS: UBLKCP), completes `bytes` on the mbarrier's tx-count.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
