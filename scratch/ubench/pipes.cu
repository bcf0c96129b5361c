// Pipe-overlap microbenchmark: per-SM throughput of MUFU.EX2, FADD2/FFMA2 (packed fp32), FMNMX3, F2FP and mixes.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
// MODE bits: 1 = 8 MUFU per iteration, 2 = 8 FFMA2, 4 = 8 scalar FFMA, 8 = 8 FMNMX(3-input), 16 = 8 cvt pack, 32 = 16 FFMA2
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float m[8]; float2 f[16]; float s[8]; float x[8]; unsigned pk[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; s[i] = seed * (i + 1); x[i] = seed + i; pk[i] = i; }
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = make_float2(seed + i, seed - i);
  const float2 c1 = make_float2(0.999f, 1.001f), c2 = make_float2(1e-3f, -1e-3f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE & 1) m[i] = ex2(m[i]);
      if (MODE & 2) f[i] = fma2(f[i], c1, c2);
      if (MODE & 32) { f[i] = fma2(f[i], c1, c2); f[8 + i] = fma2(f[8 + i], c1, c2); }
      if (MODE & 4) s[i] = fmaf(s[i], 0.999f, 1e-3f);
      if (MODE & 8) x[i] = fmaxf(fmaxf(x[i], s[(i + 1) & 7]), m[(i + 3) & 7]);
      if (MODE & 16) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(s[i] + pk[i]), "f"(x[i]));
    }
  }
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += m[i] + s[i] + x[i] + pk[i] + f[i].x + f[i].y + f[8 + i].x + f[8 + i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE>
void run(const char* name, int warps_per_smsp) {
  int sms = 148, threads = warps_per_smsp * 4 * 32, iters = 20000;
  float* out; cudaMalloc(&out, sms * threads * 4);
  k<MODE><<<sms, threads>>>(out, 100, 0.5f);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<sms, threads>>>(out, iters, 0.5f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  // warp-instruction groups of 8 per iteration per warp; report ns per (iteration x warp) per SMSP
  double per = ms * 1e6 / ((double)iters * warps_per_smsp);
  printf("%-34s warps/SMSP=%d  %.2f ms  %.2f ns per warp-iteration per SMSP (x1.9 GHz = %.1f clk)\n", name, warps_per_smsp, ms, per, per * 1.9);
  cudaFree(out);
}
int main() {
  for (int w : {1, 2, 4}) {
    run<1>("8 MUFU", w); run<2>("8 FFMA2", w); run<32>("16 FFMA2", w); run<4>("8 FFMA", w); run<8>("8 FMNMX3", w); run<16>("8 F2FP(+FADD)", w);
    run<3>("8 MUFU + 8 FFMA2", w); run<33>("8 MUFU + 16 FFMA2", w); run<5>("8 MUFU + 8 FFMA", w); run<9>("8 MUFU + 8 FMNMX3", w);
    run<35>("8 MUFU + 24 FFMA2", w); run<43>("8 MUFU + 24 FFMA2 + 8 FMNMX3", w); run<10>("8 FFMA2 + 8 FMNMX3", w);
  }
  return 0;
}
