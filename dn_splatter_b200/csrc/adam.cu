// One-launch Adam over every Gaussian parameter group ("next" row).  The reference builds one
// torch.optim.Adam per group (/root/reference/dn_splatter/dn_config.py:29-68: lr per group, eps 1e-15, no weight decay,
// no amsgrad) and nerfstudio steps them one after the other [EXT Optimizers.optimizer_step]; at 1M Gaussians that is
// 62M parameters x 28 B = 1.7 GB of HBM traffic, i.e. ~0.25 ms at speed of light, but 7 optimizers x several
// elementwise passes each in torch.  Here: one kernel, blockIdx.y = group, float4 grid-stride over the group.
//
// Update rule — torch.optim.Adam [EXT torch 2.x, _single_tensor_adam], dense (zero-gradient rows still decay):
//   m = m + (1-b1)(g - m)  [lerp];  v = b2 v + (1-b2) g g;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// with bc1 = 1 - b1^t, bc2 = 1 - b2^t computed by the host in double precision and passed per group.
//
// STATUS: written in round 1 after the GPU budget was spent — compiled, formulas pinned against torch.optim.Adam on the
// CPU (tests/test_fused_adam_cpu.py), NOT yet run on a GPU; opt-in (optim.FusedAdam).
#include "common.cuh"

namespace {

struct AdamSegDev {
  float* p; const float* g; float* m; float* v;
  int64_t n;
  float step_size, bc2_sqrt, eps;  // rounded from the host's doubles, as torch rounds its python scalars
};

struct AdamLaunch {
  AdamSegDev seg[DNR_ADAM_MAX_SEGS];
  float beta2, w1, w2;  // w = 1 - beta, rounded from double as torch does
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float w1, float b2, float w2, float step_size,
                                         float bc2_sqrt, float eps) {
  m = m + w1 * (g - m);
  v = b2 * v + w2 * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + eps;  // division, as torch does
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_kernel(const AdamLaunch L) {
  const AdamSegDev& s = L.seg[blockIdx.y];
  const float w1 = L.w1, b2 = L.beta2, w2 = L.w2;
  const float step_size = s.step_size, bc2_sqrt = s.bc2_sqrt, eps = s.eps;
  const int64_t n = s.n;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  const bool vec = ((((uintptr_t)s.p | (uintptr_t)s.g | (uintptr_t)s.m | (uintptr_t)s.v) & 15) == 0);
  const int64_t n4 = vec ? (n >> 2) : 0;
  float4* p4 = reinterpret_cast<float4*>(s.p);
  const float4* g4 = reinterpret_cast<const float4*>(s.g);
  float4* m4 = reinterpret_cast<float4*>(s.m);
  float4* v4 = reinterpret_cast<float4*>(s.v);
  for (int64_t i = tid; i < n4; i += stride) {
    float4 p = p4[i], m = m4[i], v = v4[i];
    const float4 g = g4[i];
    adam_one(p.x, g.x, m.x, v.x, w1, b2, w2, step_size, bc2_sqrt, eps);
    adam_one(p.y, g.y, m.y, v.y, w1, b2, w2, step_size, bc2_sqrt, eps);
    adam_one(p.z, g.z, m.z, v.z, w1, b2, w2, step_size, bc2_sqrt, eps);
    adam_one(p.w, g.w, m.w, v.w, w1, b2, w2, step_size, bc2_sqrt, eps);
    p4[i] = p; m4[i] = m; v4[i] = v;
  }
  for (int64_t i = (n4 << 2) + tid; i < n; i += stride) {  // tail (or everything when a pointer is unaligned)
    float p = s.p[i], m = s.m[i], v = s.v[i];
    adam_one(p, s.g[i], m, v, w1, b2, w2, step_size, bc2_sqrt, eps);
    s.p[i] = p; s.m[i] = m; s.v[i] = v;
  }
}

}  // namespace

extern "C" int dnr_adam_step(const DnrAdamSeg* segs, int32_t n_segs, double beta1, double beta2, void* stream) {
  if (!segs) return DNR_E_NULL;
  if (n_segs <= 0 || n_segs > DNR_ADAM_MAX_SEGS) return DNR_E_SIZE;
  AdamLaunch L;
  int64_t longest = 0;
  for (int i = 0; i < n_segs; ++i) {
    const DnrAdamSeg& s = segs[i];
    if (!s.p || !s.g || !s.m || !s.v) return DNR_E_NULL;
    if (s.n <= 0 || !(s.bc1 > 0.0) || !(s.bc2_sqrt > 0.0)) return DNR_E_SIZE;
    L.seg[i] = AdamSegDev{s.p, s.g, s.m, s.v, s.n, (float)(s.lr / s.bc1), (float)s.bc2_sqrt, (float)s.eps};
    longest = s.n > longest ? s.n : longest;
  }
  L.beta2 = (float)beta2;
  L.w1 = (float)(1.0 - beta1);
  L.w2 = (float)(1.0 - beta2);
  // 148 SMs x 8 resident CTAs of 256 threads; short groups leave their extra CTAs idle after one bounds check
  int64_t blocks = (longest / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<dim3((unsigned)blocks, (unsigned)n_segs), 256, 0, (cudaStream_t)stream>>>(L);
  DNR_CHECK_LAUNCH();
  return 0;
}
