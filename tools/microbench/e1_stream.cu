// The FFN epilogue's real per-chunk instruction stream (gelu_fast2 + split2 of pf_ffn_ws.cuh on 16 accumulator values),
// alone on an SM: how many cycles does a chunk cost with W epilogue-like warps per scheduler when nothing else runs
// (no TMEM traffic, no producer, no tensor pipe)?  In k_colapply_ffn_ws a chunk takes ~750 cycles per warp with two
// epilogue warps per scheduler; the pipes' busy cycles of a chunk are FMA 144 + ALU 128 + XU 128.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../phyloformer_b200/csrc -o e1_stream e1_stream.cu && ./e1_stream
#include <cstdio>
#include <cuda_runtime.h>
#include "pf_ffn_ws.cuh"

// MODE 0: the kernel's loop (pair by pair, the compiler interleaves)   MODE 1: phase by phase (all polynomials, all
// MUFU, all finals, all splits)   MODE 2: GELU only (no split)   MODE 3: split only
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_e1(float* out, int iters, float seed) {
  // the chunk's 16 accumulator values come from shared memory (4 x LDS.128, standing in for the tcgen05.ld) and the 16
  // operand words go back (4 x STS.128, standing in for the tcgen05.st): volatile, so nothing is hoisted out of the loop
  __shared__ uint4 stage[512 * 4];
  uint32_t v[16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    stage[i * 512 + threadIdx.x] = make_uint4(__float_as_uint(seed * (4 * i - 7) * 0.37f), __float_as_uint(seed * (4 * i - 6) * 0.37f),
                                            __float_as_uint(seed * (4 * i - 5) * 0.37f + 0.01f * (threadIdx.x & 31)), __float_as_uint(seed * (4 * i - 4) * 0.37f));
  __syncthreads();
  volatile uint4* my = stage + threadIdx.x;   // element i of this thread at my[i * 512]: conflict-free LDS.128
  uint32_t acc = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 q;
      asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                   : "r"((uint32_t)__cvta_generic_to_shared((const void*)(stage + i * 512 + threadIdx.x))));
      v[4 * i] = q.x ^ (uint32_t)it; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
    uint32_t hi[8], lo[8];
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) cvt2<WS_FMT_BF16X3>(gelu_fast2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), hi[i], lo[i]);
    } else if (MODE == 1) {
      u64 g[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = gelu_fast2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("" : "+l"(g[i]));   // keep the two phases apart
#pragma unroll
      for (int i = 0; i < 8; ++i) split2(g[i], hi[i], lo[i]);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const u64 g = gelu_fast2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])); hi[i] = (uint32_t)g; lo[i] = (uint32_t)(g >> 32); }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) split2(pk2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), hi[i], lo[i]);
    }
    if (hi[0] == 0x7fc00001u) {   // (practically) never true: the results are live without being stored every chunk
#pragma unroll
      for (int i = 0; i < 2; ++i) { my[i * 512].x = hi[4 * i]; my[i * 512].y = hi[4 * i + 1]; my[i * 512].z = hi[4 * i + 2]; my[i * 512].w = hi[4 * i + 3]; }
#pragma unroll
      for (int i = 0; i < 2; ++i) { my[(2 + i) * 512].x = lo[4 * i]; my[(2 + i) * 512].y = lo[4 * i + 1]; my[(2 + i) * 512].z = lo[4 * i + 2]; my[(2 + i) * 512].w = lo[4 * i + 3]; }
    }
    acc ^= hi[1] ^ lo[2] ^ hi[3] ^ lo[4] ^ hi[5] ^ lo[6] ^ hi[7] ^ lo[0] ^ hi[2] ^ lo[1] ^ hi[4] ^ lo[3] ^ hi[6] ^ lo[5] ^ lo[7];
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (float)(t1 - t0);
}

template <int MODE>
void run(const char* name, int warps_per_sched) {
  const int threads = warps_per_sched * 4 * 32, iters = 4000;
  float* out;
  cudaMalloc(&out, (148 * threads + 1) * sizeof(float));
  for (int rep = 0; rep < 2; ++rep) { k_e1<MODE><<<148, threads>>>(out, iters, 1.0f); cudaDeviceSynchronize(); }
  float cyc;
  cudaMemcpy(&cyc, out + 148 * threads, 4, cudaMemcpyDeviceToHost);
  printf("%-44s %d warps/scheduler: %7.1f cycles per chunk and warp  (%6.1f per chunk and scheduler-slot)\n", name, warps_per_sched,
         cyc / iters, cyc / iters / warps_per_sched);
  cudaFree(out);
}

int main() {
  for (int w : {1, 2, 3, 4}) {
    run<0>("chunk as in the kernel (gelu + split)", w);
    run<1>("gelu phase, then split phase", w);
    run<2>("gelu only", w);
    run<3>("split only", w);
  }
  return 0;
}
