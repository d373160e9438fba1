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
// Diagnostic variants of gelu_fast2 (wrong values by construction, same dependency shape): which pipe binds?
template <int DROP>   // 1: no MUFU (FMUL instead)   2: no FMNMX (FADD instead)   3: neither
__device__ __forceinline__ u64 gelu_diag2(float h0, float h1) {
  float t0, t1, m0, m1;
  if (DROP & 2) { t0 = h0 + 10.0f; t1 = h1 + 10.0f; m0 = h0 + 0.5f; m1 = h1 + 0.5f; }
  else { t0 = fminf(fabsf(h0), 10.0f); t1 = fminf(fabsf(h1), 10.0f); m0 = fmaxf(h0, 0.f); m1 = fmaxf(h1, 0.f); }
  const u64 t = pk2(t0, t1);
  u64 p = pk2(3.2904327396e-05f, 3.2904327396e-05f);
  p = fma2(p, t, pk2(-7.6214972445e-04f, -7.6214972445e-04f));
  p = fma2(p, t, pk2(8.0388012506e-03f, 8.0388012506e-03f));
  p = fma2(p, t, pk2(-5.3315325260e-02f, -5.3315325260e-02f));
  p = fma2(p, t, pk2(-4.5887145465e-01f, -4.5887145465e-01f));
  p = fma2(p, t, pk2(-1.1511568274e+00f, -1.1511568274e+00f));
  p = fma2(p, t, pk2(-9.9999958869e-01f, -9.9999958869e-01f));
  float p0, p1, e0, e1;
  up2(p, p0, p1);
  if (DROP & 1) { e0 = p0 * 1.0001f; e1 = p1 * 1.0001f; }
  else { asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(p0)); asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(p1)); }
  return fma2(pk2(-t0, -t1), pk2(e0, e1), pk2(m0, m1));
}
// truncating split: hi = the floats' upper halves (one PRMT), lo = rn(g - hi): one F2FP per pair instead of two
__device__ __forceinline__ void split2_trunc(u64 g, uint32_t& hi, uint32_t& lo) {
  float g0, g1;
  up2(g, g0, g1);
  const uint32_t u0 = __float_as_uint(g0), u1 = __float_as_uint(g1);
  asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(hi) : "r"(u0), "r"(u1));
  const u64 hf = pk2(__uint_as_float(u0 & 0xffff0000u), __uint_as_float(u1 & 0xffff0000u));
  float l0, l1;
  up2(fma2(hf, pk2(-1.f, -1.f), g), l0, l1);
  const __nv_bfloat162 ll = __floats2bfloat162_rn(l0, l1);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}

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
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) split2(pk2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), hi[i], lo[i]);
    } else if (MODE >= 4 && MODE <= 6) {   // full chunk with a diagnostic GELU
#pragma unroll
      for (int i = 0; i < 8; ++i) split2(gelu_diag2<MODE - 3>(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), hi[i], lo[i]);
    } else {                               // 7: real GELU, truncating split
#pragma unroll
      for (int i = 0; i < 8; ++i) split2_trunc(gelu_fast2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), hi[i], lo[i]);
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
  for (int w : {2, 4}) {
    run<0>("chunk as in the kernel (gelu + split)", w);
    run<1>("gelu phase, then split phase", w);
    run<2>("gelu only", w);
    run<3>("split only", w);
    run<4>("chunk, MUFU.EX2 replaced by FMUL", w);
    run<5>("chunk, FMNMX replaced by FADD", w);
    run<6>("chunk, neither MUFU nor FMNMX", w);
    run<7>("chunk, truncating split (1 F2FP per pair)", w);
  }
  return 0;
}
