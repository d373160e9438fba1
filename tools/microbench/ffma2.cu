// Microbenchmark: scalar FFMA vs packed FFMA2 throughput on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int ILP>
__global__ void k_scalar(float* out, int iters, float a, float b) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_packed(float* out, int iters, float a, float b) {
  u64 acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = pk2(threadIdx.x * 0.001f + i, i);
  const u64 a2 = pk2(a, a), b2 = pk2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma2(acc[i], a2, b2);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i])); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4 * 4);
  const int iters = 20000;
  for (int warps = 1; warps <= 8; warps *= 2) {
    const int threads = warps * 32 * 4;  // warps per SMSP
    const int grid = 148;
    float ms_s = timeit([&] { k_scalar<8><<<grid, threads>>>(out, iters, 1.0001f, 0.5f); });
    float ms_p = timeit([&] { k_packed<8><<<grid, threads>>>(out, iters, 1.0001f, 0.5f); });
    float ms_s1 = timeit([&] { k_scalar<1><<<grid, threads>>>(out, iters * 8, 1.0001f, 0.5f); });
    float ms_p1 = timeit([&] { k_packed<1><<<grid, threads>>>(out, iters * 8, 1.0001f, 0.5f); });
    const double fl = (double)grid * threads * iters * 8;
    printf("warps/SMSP %d: scalar ILP8 %.2f TFMA/s | packed ILP8 %.2f TFMA/s (=%.2f packed-instr T/s) | dep-chain latency: scalar %.2f cyc, packed %.2f cyc (at 1.9 GHz)\n",
           warps, fl / ms_s / 1e9, 2 * fl / ms_p / 1e9, fl / ms_p / 1e9,
           ms_s1 * 1e-3 * 1.9e9 / (iters * 8.0), ms_p1 * 1e-3 * 1.9e9 / (iters * 8.0));
  }
  return 0;
}
