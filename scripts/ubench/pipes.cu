// Issue-rate microbenchmarks for the instruction classes raster_fwd/raster_bwd are made of (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/pipes scripts/ubench/pipes.cu && ./pipes
// Prints warp-instructions per clock per SM sub-partition (SMSP) for each class, measured with clock64() inside
// resident-everywhere launches (8 warps per SMSP, 8 independent dependency chains per thread).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define ITERS 2048

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float lo2(unsigned long long v) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  return a + b;
}

template <int MODE>
__global__ void __launch_bounds__(1024) bench(float* out, long long* cycles, float seed) {
  __shared__ float4 sm[256];
  if (threadIdx.x < 256) sm[threadIdx.x] = make_float4(seed, seed * 2, seed * 3, seed * 4);
  __syncthreads();
  float x[8];
  unsigned long long xp[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { x[k] = seed * (k + 1) + threadIdx.x * 1e-6f; xp[k] = pack2(x[k], x[k] * 0.5f); }
  const float a = 1.0f + seed * 1e-7f, b = seed * 1e-8f;
  const unsigned long long ap = pack2(a, a), bp = pack2(b, b);
  const int lane = threadIdx.x & 31;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE == 0) x[k] = fmaf(x[k], a, b);                                        // FFMA 3-reg
      if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(xp[k]) : "l"(ap), "l"(bp));  // FFMA2
      if (MODE == 2) x[k] = fminf(x[k], a) + 0.f * b;                                 // FMNMX (alu) (+ folded)
      if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[k]));          // MUFU.EX2
      if (MODE == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[k]));          // MUFU.RCP
      if (MODE == 5) x[k] = __shfl_xor_sync(0xffffffffu, x[k], 1 + (k & 3));           // SHFL
      if (MODE == 6) x[k] = (lane & (1 << (k & 3))) ? x[k] : x[(k + 1) & 7];           // SEL
      if (MODE == 7) { const float4 q = sm[(it + k) & 255]; x[k] += q.x; }             // LDS.128 broadcast + FADD
      if (MODE == 8) { x[k] = fmaf(x[k], a, b); x[k] = fminf(x[k], 3.0e38f); }        // FFMA + FMNMX interleaved
      if (MODE == 9) {                                                                 // FFMA2 + FMNMX interleaved
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(xp[k]) : "l"(ap), "l"(bp));
        x[k] = fminf(x[k], a);
      }
      if (MODE == 10) { x[k] = fmaf(x[k], a, b); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[(k + 4) & 7])); }  // 1:1 FFMA:MUFU
      if (MODE == 11) { x[k] = x[k] * a; }                                            // FMUL
      if (MODE == 12) { x[k] = x[k] + a; }                                            // FADD
      if (MODE == 13) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(xp[k]) : "l"(ap)); // FADD2
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k] + lo2(xp[k]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter_instr) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int threads = 1024;  // 32 warps / SM = 8 per SMSP, 1 block per SM
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * threads);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  bench<MODE><<<sms, threads>>>(out, cyc, 1.0f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<MODE><<<sms, threads>>>(out, cyc, 1.0f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (auto c : h) avg += (double)c;
  avg /= sms;
  const double instr_per_warp = (double)ITERS * 8 * per_iter_instr;
  const double rate = instr_per_warp * 8 /*warps per SMSP*/ / avg;
  printf("{\"class\": \"%s\", \"warp_instr_per_clk_per_smsp\": %.3f, \"cycles\": %.0f, \"ms\": %.4f, \"err\": \"%s\"}\n", name, rate, avg, ms,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FFMA", 1);
  run<11>("FMUL", 1);
  run<12>("FADD", 1);
  run<1>("FFMA2 (f32x2)", 1);
  run<13>("FADD2 (f32x2)", 1);
  run<2>("FMNMX", 1);
  run<3>("MUFU.EX2", 1);
  run<4>("MUFU.RCP", 1);
  run<5>("SHFL.BFLY", 1);
  run<6>("SEL", 1);
  run<7>("LDS.128 bcast + FADD", 2);
  run<8>("FFMA + FMNMX", 2);
  run<9>("FFMA2 + FMNMX", 2);
  run<10>("FFMA + MUFU.EX2", 2);
  return 0;
}
