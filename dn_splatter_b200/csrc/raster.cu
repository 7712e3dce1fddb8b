// Per-tile front-to-back alpha compositing of RGB + expected depth + per-Gaussian normal in ONE pass,
// forward and backward.
//
// Replaces (reference /root/reference/dn_splatter/dn_model.py):
//   :495-516  gsplat rasterize_to_pixels fwd (RGB+ED)  and its autograd backward      [EXT gsplat 1.0.0]
//   :564-575  gsplat legacy rasterize_gaussians on normals (white background) + bwd   [EXT]
//   :526-537  rgb = clamp(render + (1-alpha) bg), depth = where(alpha>0, ED, max)     (max: finalize)
//   :577-578  normal = (n/|n| + 1)/2
// The two reference passes share alpha/T exactly (same means2d, conics, opacities, order), so one
// compositing loop with 7 channels reproduces both; only the gradient routing differs (the normal pass
// sees detached xys: its alpha-gradient reaches conics/opacity but not means2d — quirk B3).
//
// Data movement: each CTA (one 16x16 tile) walks its slice of the sorted id list in chunks of 128.
// Records are gathered by id from the packed per-Gaussian record array straight into shared memory by
// per-thread 1-D bulk async copies (cp.async.bulk -> UBLKCP, TMA engine) completing on an mbarrier,
// double-buffered so the next chunk lands while the current one is composited.  Only the chunks a tile
// actually consumes before all of its pixels saturate are ever fetched.
#include "common.cuh"

namespace {

#ifndef DNR_BWD_PPT
#define DNR_BWD_PPT 2      // pixels per thread in raster_bwd (1 or 2)
#endif
constexpr int CH = 128;   // records per chunk
constexpr int STAGES = 2;

__device__ __forceinline__ void pixel_of_thread(int tid, int& lx, int& ly) {
  // warp -> 8x4 pixel patch (better alpha-test coherence than a 16x2 strip); 8 warps = 2x4 patches
  const int w = tid >> 5, l = tid & 31;
  lx = ((w & 1) << 3) + (l & 7);
  ly = ((w >> 1) << 2) + (l >> 3);
}

template <int REC>
__device__ __forceinline__ void issue_chunk(const float* __restrict__ records, const int32_t* __restrict__ ids,
                                            int n_c, float* stage_smem, uint64_t* bar, int tid) {
  if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(n_c * REC * 4));
  if (tid < n_c) {
    const int g = ids[tid];
    bulk_g2s(stage_smem + tid * REC, records + (size_t)g * REC, REC * 4, bar);
  }
}

template <bool NORMALS>
__global__ void __launch_bounds__(256) raster_fwd_kernel(const DnrArgs a, int tiles_x) {
  constexpr int REC = NORMALS ? DNR_REC_FLOATS_N : DNR_REC_FLOATS;
  __shared__ __align__(128) float recs[STAGES][CH * REC];
  __shared__ __align__(8) uint64_t bars[STAGES];
  __shared__ float red_max[8];

  const int tid = threadIdx.x;
  const int tile = blockIdx.y * tiles_x + blockIdx.x;
  int lx, ly;
  pixel_of_thread(tid, lx, ly);
  const int j = blockIdx.x * DNR_TILE + lx, i = blockIdx.y * DNR_TILE + ly;
  const bool inside = (i < a.height) && (j < a.width);
  const float px = (float)j + 0.5f, py = (float)i + 0.5f;
  const int start = a.tile_offsets[tile], end = a.tile_offsets[tile + 1];
  const int n = end - start;
  const int nchunks = (n + CH - 1) / CH;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  float T = 1.0f;
  float C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f, N0 = 0.f, N1 = 0.f, N2 = 0.f;
  int last = 0;
  bool done = !inside;

  int issued = 0, consumed = 0;
  if (nchunks > 0) {
    issue_chunk<REC>(a.records, a.flatten_ids + start, min(CH, n), recs[0], &bars[0], tid);
    issued = 1;
  }
  for (int c = 0; c < nchunks; ++c) {
    const int stage = c & 1;
    if (c + 1 < nchunks) {  // stage (c+1)&1 was released by the barrier that closed iteration c-1
      const int n_next = min(CH, n - (c + 1) * CH);
      issue_chunk<REC>(a.records, a.flatten_ids + start + (c + 1) * CH, n_next, recs[stage ^ 1], &bars[stage ^ 1], tid);
      issued = c + 2;
    }
    mbar_wait(&bars[stage], (uint32_t)((c >> 1) & 1));
    consumed = c + 1;
    const int n_c = min(CH, n - c * CH);
    const float4* r4 = reinterpret_cast<const float4*>(recs[stage]);
    const int base = start + c * CH;
    // Warp-uniform loop: every lane walks the records in lockstep (predicated), so the warp issues each
    // record once.  (A per-lane `continue`/`break` loop lets lanes drift apart under independent thread
    // scheduling: measured 3/32 active lanes and 8x the instruction count.)
    for (int t0 = 0; t0 < n_c; t0 += 8) {
      if (__all_sync(0xffffffffu, done)) break;
      const int t1 = min(t0 + 8, n_c);
#pragma unroll 8
      for (int t = t0; t < t1; ++t) {
        if (!done) {
          const float4 q0 = r4[t * (REC / 4) + 0];  // x, y, a', b'
          const float4 q1 = r4[t * (REC / 4) + 1];  // c', opacity, -log2(255 opacity) - slack, -
          const float dx = q0.x - px, dy = q0.y - py;
          const float pw = dnr_power2(q0.z, q0.w, q1.x, dx, dy);  // = -sigma * log2(e)
          if (!(pw > 0.f || pw < q1.z)) {                          // else: sigma < 0, or alpha certainly < 1/255
            const float alpha = fminf(DNR_ALPHA_MAX, __fmul_rn(q1.y, dnr_ex2(pw)));
            if (!(alpha < DNR_ALPHA_MIN)) {
              const float next_T = T * (1.0f - alpha);
              if (next_T <= DNR_T_STOP) {
                done = true;
              } else {
                const float4 q2 = r4[t * (REC / 4) + 2];  // r, g, b, depth
                const float vis = alpha * T;
                C0 += q2.x * vis; C1 += q2.y * vis; C2 += q2.z * vis; D += q2.w * vis;
                if (NORMALS) {
                  const float4 q3 = r4[t * (REC / 4) + 3];  // camera-space normal
                  N0 += q3.x * vis; N1 += q3.y * vis; N2 += q3.z * vis;
                }
                last = base + t;
                T = next_T;
              }
            }
          }
        }
      }
    }
    if (__syncthreads_count(done) == 256) break;
  }
  // never leave the CTA with a bulk copy still in flight into its shared memory
  if (issued > consumed) mbar_wait(&bars[(issued - 1) & 1], (uint32_t)(((issued - 1) >> 1) & 1));

  float ed_for_max = 0.f;
  if (inside) {
    const int pix = i * a.width + j;
    const float alpha = 1.0f - T;
    const float om = 1.0f - alpha;
    const float pre[3] = {C0 + om * a.background[0], C1 + om * a.background[1], C2 + om * a.background[2]};
    uint8_t mask = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (pre[k] >= 0.f && pre[k] <= 1.f) mask |= (uint8_t)(1u << k);
      a.out_rgb[pix * 3 + k] = fminf(fmaxf(pre[k], 0.f), 1.f);
    }
    a.clamp_mask[pix] = mask;
    const float ed = D / fmaxf(alpha, 1e-10f);
    a.out_depth[pix] = ed;
    a.out_alpha[pix] = alpha;
    a.last_ids[pix] = last;
    ed_for_max = ed;
    if (NORMALS) {
      const float n0 = N0 + T, n1 = N1 + T, n2 = N2 + T;  // white background (quirk B1)
      const float nn = sqrtf(n0 * n0 + n1 * n1 + n2 * n2);
      a.normal_norm[pix] = nn;
      a.out_normal[pix * 3 + 0] = (n0 / nn + 1.0f) * 0.5f;
      a.out_normal[pix * 3 + 1] = (n1 / nn + 1.0f) * 0.5f;
      a.out_normal[pix * 3 + 2] = (n2 / nn + 1.0f) * 0.5f;
    }
  }
  // global max of the expected depth (dn_model.py:536: depth_im.detach().max())
  const float wm = warp_max(ed_for_max);
  if ((tid & 31) == 0) red_max[tid >> 5] = wm;
  __syncthreads();
  if (tid == 0) {
    float m = red_max[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red_max[w]);
    if (m > 0.f) atomicMax(a.depth_max, __float_as_int(m));
  }
}

// 16 per-lane values -> per-value warp totals in 16 shuffles (transposing butterfly); on return the lane
// holds in v[0] the total of value index (lane>>1)&15.
__device__ __forceinline__ void butterfly16(float (&v)[16], int lane) {
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float send = up ? v[k] : v[k + 8];
      const float keep = up ? v[k + 8] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float send = up ? v[k] : v[k + 4];
      const float keep = up ? v[k + 4] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = (lane & 4) != 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float send = up ? v[k] : v[k + 2];
      const float keep = up ? v[k + 2] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool up = (lane & 2) != 0;
    const float send = up ? v[0] : v[1];
    const float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// PPT pixels per thread: 256/PPT threads per tile.  With PPT = 2 a warp covers an 8x8 patch and the 16-value
// butterfly + RED (the largest fixed cost per record) is paid once per 64 pixels instead of once per 32.
template <bool NORMALS, int PPT>
__global__ void __launch_bounds__(256 / PPT) raster_bwd_kernel(const DnrArgs a, int tiles_x) {
  constexpr int REC = NORMALS ? DNR_REC_FLOATS_N : DNR_REC_FLOATS;
  constexpr int NT = 256 / PPT;   // threads per tile
  constexpr int NW = NT / 32;     // warps per tile
  __shared__ __align__(128) float recs[STAGES][CH * REC];
  __shared__ int ids_s[STAGES][CH];
  __shared__ __align__(8) uint64_t bars[STAGES];
  __shared__ int red_last[NW];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.y * tiles_x + blockIdx.x;
  const int start = a.tile_offsets[tile], end = a.tile_offsets[tile + 1];

  // ---- per-pixel state and the gradient of the glue (P1/P3 backward) ----
  float px[PPT], py[PPT], T_final[PPT], vC0[PPT], vC1[PPT], vC2[PPT], vD[PPT], vN0[PPT], vN1[PPT], vN2[PPT];
  float Tf_va_cd[PPT], Tf_va_n[PPT];
  int last_id[PPT];
  bool inside[PPT];
  int wl = -1;
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    int lx, ly;
    if (PPT == 1) {
      pixel_of_thread(tid, lx, ly);
    } else {  // warp -> 8 wide x (4*PPT) tall patch; the thread's pixels are 4 rows apart
      lx = ((warp & 1) << 3) + (lane & 7);
      ly = (warp >> 1) * (4 * PPT) + (lane >> 3) + 4 * p;
    }
    const int j = blockIdx.x * DNR_TILE + lx, i = blockIdx.y * DNR_TILE + ly;
    inside[p] = (i < a.height) && (j < a.width);
    px[p] = (float)j + 0.5f; py[p] = (float)i + 0.5f;
    const int pix = inside[p] ? i * a.width + j : 0;
    last_id[p] = inside[p] ? a.last_ids[pix] : -1;
    wl = max(wl, last_id[p]);
    T_final[p] = 1.f;
    vC0[p] = vC1[p] = vC2[p] = vD[p] = vN0[p] = vN1[p] = vN2[p] = 0.f;
    float va_cd = 0.f, va_n = 0.f;
    if (inside[p]) {
      const float alpha = a.out_alpha[pix];
      T_final[p] = 1.0f - alpha;
      if (a.v_rgb) {
        const uint8_t m = a.clamp_mask[pix];
        vC0[p] = (m & 1) ? a.v_rgb[pix * 3 + 0] : 0.f;
        vC1[p] = (m & 2) ? a.v_rgb[pix * 3 + 1] : 0.f;
        vC2[p] = (m & 4) ? a.v_rgb[pix * 3 + 2] : 0.f;
        va_cd -= a.background[0] * vC0[p] + a.background[1] * vC1[p] + a.background[2] * vC2[p];
      }
      if (a.v_alpha) va_cd += a.v_alpha[pix];
      if (a.v_depth && alpha > 0.f) {
        const float v_ed = a.v_depth[pix];
        const float ac = fmaxf(alpha, 1e-10f);
        vD[p] = v_ed / ac;
        if (alpha >= 1e-10f) va_cd -= v_ed * a.out_depth[pix] / ac;
      }
      if (NORMALS && a.v_normal) {
        const float nn = a.normal_norm[pix];
        const float n0 = 2.0f * a.out_normal[pix * 3 + 0] - 1.0f, n1 = 2.0f * a.out_normal[pix * 3 + 1] - 1.0f,
                    n2 = 2.0f * a.out_normal[pix * 3 + 2] - 1.0f;
        const float g0 = 0.5f * a.v_normal[pix * 3 + 0], g1 = 0.5f * a.v_normal[pix * 3 + 1], g2 = 0.5f * a.v_normal[pix * 3 + 2];
        const float dp = n0 * g0 + n1 * g1 + n2 * g2;
        vN0[p] = (g0 - n0 * dp) / nn; vN1[p] = (g1 - n1 * dp) / nn; vN2[p] = (g2 - n2 * dp) / nn;
        va_n = -(vN0[p] + vN1[p] + vN2[p]);
      }
    }
    Tf_va_cd[p] = T_final[p] * va_cd;
    Tf_va_n[p] = T_final[p] * va_n;
  }

  // ---- range actually composited by this tile ----
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wl = max(wl, __shfl_xor_sync(0xffffffffu, wl, o));
  if (lane == 0) red_last[warp] = wl;  // wl: deepest record any pixel of this warp composited
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  int hi = red_last[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) hi = max(hi, red_last[w]);
  hi = min(hi + 1, end);  // exclusive
  const int n = hi - start;
  if (n <= 0) return;
  const int nchunks = (n + CH - 1) / CH;

  float T[PPT], S_cd[PPT], S_n[PPT];
#pragma unroll
  for (int p = 0; p < PPT; ++p) { T[p] = T_final[p]; S_cd[p] = 0.f; S_n[p] = 0.f; }

  // chunk c covers absolute indices [chi - n_c, chi), chi = hi - c*CH; slot t <-> index chi-1-t
  auto issue = [&](int c) {
    const int chi = hi - c * CH;
    const int n_c = min(CH, chi - start);
    const int stage = c & 1;
    if (tid == 0) mbar_arrive_expect_tx(&bars[stage], (uint32_t)(n_c * REC * 4));
    for (int t = tid; t < n_c; t += NT) {
      const int g = a.flatten_ids[chi - 1 - t];
      ids_s[stage][t] = g;
      bulk_g2s(recs[stage] + t * REC, a.records + (size_t)g * REC, REC * 4, &bars[stage]);
    }
  };
  issue(0);
  for (int c = 0; c < nchunks; ++c) {
    const int stage = c & 1;
    if (c + 1 < nchunks) issue(c + 1);
    mbar_wait(&bars[stage], (uint32_t)((c >> 1) & 1));
    __syncthreads();  // ids_s[stage] written by other threads
    const int chi = hi - c * CH;
    const int n_c = min(CH, chi - start);
    const float4* r4 = reinterpret_cast<const float4*>(recs[stage]);
    // records deeper than anything this warp composited are skipped without touching them (warp-uniform)
    for (int t = max(0, chi - 1 - wl); t < n_c; ++t) {
      const int idx = chi - 1 - t;
      const float4 q0 = r4[t * (REC / 4) + 0];
      const float4 q1 = r4[t * (REC / 4) + 1];
      bool valid[PPT];
      float dx[PPT], dy[PPT], vis[PPT], alpha[PPT];
      bool any = false;
#pragma unroll
      for (int p = 0; p < PPT; ++p) {
        valid[p] = inside[p] && (idx <= last_id[p]);
        dx[p] = q0.x - px[p]; dy[p] = q0.y - py[p];
        vis[p] = 0.f; alpha[p] = 0.f;
        if (valid[p]) {
          const float pw = dnr_power2(q0.z, q0.w, q1.x, dx[p], dy[p]);
          valid[p] = !(pw > 0.f || pw < q1.z);
          if (valid[p]) {
            vis[p] = dnr_ex2(pw);
            alpha[p] = fminf(DNR_ALPHA_MAX, __fmul_rn(q1.y, vis[p]));
            valid[p] = !(alpha[p] < DNR_ALPHA_MIN);
          }
        }
        any = any || valid[p];
      }
      if (!__any_sync(0xffffffffu, any)) continue;
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = 0.f;
      if (any) {
        const float4 q2 = r4[t * (REC / 4) + 2];
        float4 q3 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (NORMALS) q3 = r4[t * (REC / 4) + 3];
        const float opac = q1.y;
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
          if (valid[p]) {
            float ra;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(1.0f - alpha[p]));
            T[p] *= ra;
            const float fac = alpha[p] * T[p];
            // d(out)/d(alpha_i) = sum_k (c_k T - B_k ra) v_k + T_final ra v_a, B_k = sum_{j>i} c_jk fac_j.  Only the
            // contraction S = sum_k B_k v_k is needed: one running scalar per gradient route replaces 7 buffers.
            const float dot_cd = q2.x * vC0[p] + q2.y * vC1[p] + q2.z * vC2[p] + q2.w * vD[p];
            const float v_alpha_cd = fmaf(T[p], dot_cd, ra * (Tf_va_cd[p] - S_cd[p]));
            S_cd[p] = fmaf(fac, dot_cd, S_cd[p]);
            v[8] += fac * vC0[p]; v[9] += fac * vC1[p]; v[10] += fac * vC2[p]; v[11] += fac * vD[p];
            float v_alpha_n = 0.f;
            if (NORMALS) {
              const float dot_n = q3.x * vN0[p] + q3.y * vN1[p] + q3.z * vN2[p];
              v_alpha_n = fmaf(T[p], dot_n, ra * (Tf_va_n[p] - S_n[p]));
              S_n[p] = fmaf(fac, dot_n, S_n[p]);
              v[12] += fac * vN0[p]; v[13] += fac * vN1[p]; v[14] += fac * vN2[p];
            }
            const float ov = opac * vis[p];
            if (ov <= DNR_ALPHA_MAX) {
              const float v_alpha_all = v_alpha_cd + v_alpha_n;
              const float vs_cd = -ov * v_alpha_cd;   // d/d sigma
              const float vs_all = -ov * v_alpha_all;
              v[4] += 0.5f * vs_all * dx[p] * dx[p];
              v[5] += vs_all * dx[p] * dy[p];
              v[6] += 0.5f * vs_all * dy[p] * dy[p];
              // d sigma / d mean2d = (A dx + B dy, B dx + C dy) with A = -2 ln2 a', B = -ln2 b', C = -2 ln2 c'
              const float k = -DNR_LN2 * vs_cd;
              const float gx = k * (2.0f * q0.z * dx[p] + q0.w * dy[p]);
              const float gy = k * (q0.w * dx[p] + 2.0f * q1.x * dy[p]);
              v[0] += gx; v[1] += gy; v[2] += fabsf(gx); v[3] += fabsf(gy);
              v[7] += vis[p] * v_alpha_all;
            }
          }
        }
      }
      butterfly16(v, lane);
      if ((lane & 1) == 0 && v[0] != 0.f) {
        const int g = ids_s[stage][t];
        atomicAdd(a.grad_records + (size_t)g * DNR_GRAD_FLOATS + ((lane >> 1) & 15), v[0]);
      }
    }
    __syncthreads();  // stage free for the chunk after next
  }
}

}  // namespace

extern "C" int dnr_raster_fwd(const DnrArgs* a, void* stream) {
  if (!a) return DNR_E_NULL;
  if (a->width <= 0 || a->height <= 0 || a->n_isects < 0) return DNR_E_SIZE;
  if (a->tile_size != DNR_TILE) return DNR_E_OPTION;
  if (!a->records || !a->tile_offsets || !a->out_rgb || !a->out_depth || !a->out_alpha || !a->last_ids ||
      !a->clamp_mask || !a->depth_max)
    return DNR_E_NULL;
  if (a->n_isects > 0 && !a->flatten_ids) return DNR_E_NULL;
  const bool normals = (a->flags & DNR_FLAG_NORMALS) != 0;
  if (normals && (!a->out_normal || !a->normal_norm)) return DNR_E_NULL;
  cudaStream_t s = (cudaStream_t)stream;
  DNR_CUDA(cudaMemsetAsync(a->depth_max, 0, sizeof(int32_t), s));
  const dim3 grid(dnr_tiles_x(a), dnr_tiles_y(a));
  if (normals) raster_fwd_kernel<true><<<grid, 256, 0, s>>>(*a, grid.x);
  else raster_fwd_kernel<false><<<grid, 256, 0, s>>>(*a, grid.x);
  DNR_CHECK_LAUNCH();
  return 0;
}

extern "C" int dnr_raster_bwd(const DnrArgs* a, void* stream) {
  if (!a) return DNR_E_NULL;
  if (a->width <= 0 || a->height <= 0 || a->n_gauss <= 0 || a->n_isects < 0) return DNR_E_SIZE;
  if (!a->records || !a->tile_offsets || !a->out_depth || !a->out_alpha || !a->last_ids || !a->clamp_mask ||
      !a->grad_records)
    return DNR_E_NULL;
  if (a->n_isects > 0 && !a->flatten_ids) return DNR_E_NULL;
  const bool normals = (a->flags & DNR_FLAG_NORMALS) != 0;
  if (normals && (!a->out_normal || !a->normal_norm)) return DNR_E_NULL;
  cudaStream_t s = (cudaStream_t)stream;
  DNR_CUDA(cudaMemsetAsync(a->grad_records, 0, (size_t)a->n_gauss * DNR_GRAD_FLOATS * sizeof(float), s));
  if (a->n_isects == 0) return 0;
  const dim3 grid(dnr_tiles_x(a), dnr_tiles_y(a));
  constexpr int PPT = DNR_BWD_PPT;
  if (normals) raster_bwd_kernel<true, PPT><<<grid, 256 / PPT, 0, s>>>(*a, grid.x);
  else raster_bwd_kernel<false, PPT><<<grid, 256 / PPT, 0, s>>>(*a, grid.x);
  DNR_CHECK_LAUNCH();
  return 0;
}
