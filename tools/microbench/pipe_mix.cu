// Do the FMA (packed FFMA2), ALU (FMNMX / LOP3 / F2FP) and XU (MUFU.EX2) pipes of an sm_100 scheduler overlap, or do
// their busy cycles add up?  Four kernels with the FFN epilogue's per-pair instruction mix (8 FFMA2, 9 ALU, 2 MUFU),
// 8 independent chains per thread, W warps per scheduler:   nvcc -arch=sm_100a -O3 -o pipe_mix pipe_mix.cu && ./pipe_mix
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float ex2(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fmn(float a, float b) { float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
template <int NF, int NA, int NX>
__global__ void k(float* out, int iters, float seed) {
  u64 p[8]; float a[8], x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { p[i] = (u64)__float_as_uint(seed + i) | ((u64)__float_as_uint(seed - i) << 32); a[i] = seed * i; x[i] = seed + 0.1f * i; }
  const u64 c = (u64)__float_as_uint(0.999f) | ((u64)__float_as_uint(1.001f) << 32);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int f = 0; f < NF; ++f) p[i] = fma2(p[i], c, c);
      if (NA) {   // 9 dependent half-rate ALU ops of the epilogue's kinds: 4 FMNMX, 2 F2FP, SHF, 2 LOP3
        float v = a[i];
        unsigned u;
        v = fmn(v, 10.0f);
        asm volatile("and.b32 %0, %1, 0xffff0000;" : "=r"(u) : "r"(__float_as_uint(v)));
        asm volatile("max.f32 %0, %1, 0f00000000;" : "=f"(v) : "f"(__uint_as_float(u)));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(v), "f"(v));
        asm volatile("shl.b32 %0, %1, 16;" : "=r"(u) : "r"(u));
        v = fmn(__uint_as_float(u), 9.0f);
        asm volatile("and.b32 %0, %1, 0xfffffff0;" : "=r"(u) : "r"(__float_as_uint(v)));
        asm volatile("max.f32 %0, %1, 0f3f000000;" : "=f"(v) : "f"(__uint_as_float(u)));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(v), "f"(v));
        a[i] = __uint_as_float(u);
      }
#pragma unroll
      for (int f = 0; f < NX; ++f) x[i] = ex2(x[i]);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)p[i]) + a[i] + x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (float)(t1 - t0);
}
template <int NF, int NA, int NX>
void run(const char* name, int warps_per_sched) {
  const int threads = warps_per_sched * 4 * 32, iters = 2000;
  float* out; cudaMalloc(&out, (148 * threads + 1) * sizeof(float));
  k<NF, NA, NX><<<148, threads>>>(out, iters, 1.0f); cudaDeviceSynchronize();
  k<NF, NA, NX><<<148, threads>>>(out, iters, 1.0f); cudaDeviceSynchronize();
  float cyc; cudaMemcpy(&cyc, out + 148 * threads, 4, cudaMemcpyDeviceToHost);
  // per scheduler and per "pair" (one chain step of one warp): cycles
  printf("%-28s warps/sched %d: %.2f cycles per warp-pair-step per scheduler\n", name, warps_per_sched, cyc / iters / 8.0 / warps_per_sched);
  cudaFree(out);
}
int main() {
  for (int w : {2, 4}) {
    if (w == 2) { run<8, 0, 0>("8 FFMA2", 2); run<0, 9, 0>("9 FMNMX(+FADD)", 2); run<0, 0, 2>("2 MUFU.EX2", 2); run<8, 9, 2>("8 FFMA2 + 9 ALU + 2 MUFU", 2); }
    else        { run<8, 0, 0>("8 FFMA2", 4); run<0, 9, 0>("9 FMNMX(+FADD)", 4); run<0, 0, 2>("2 MUFU.EX2", 4); run<8, 9, 2>("8 FFMA2 + 9 ALU + 2 MUFU", 4); }
  }
  return 0;
}
