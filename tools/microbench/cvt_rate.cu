// Issue cost of the conversion / packing instructions of the bf16 hi/lo split on sm_100 (per scheduler, many independent
// chains): F2FP.BF16.F32.PACK_AB (cvt.rn.bf16x2.f32), PRMT, LOP3, SHL, FADD2 and the whole split2 sequence.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cvt_rate cvt_rate.cu && ./cvt_rate
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int OP>
__global__ void k(float* out, int iters, float seed) {
  float a[8], b[8];
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + i; b[i] = seed * 0.5f - i; u[i] = i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) {        // 4 x F2FP (dependent through a)
#pragma unroll
        for (int r = 0; r < 4; ++r) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(b[i])); a[i] = __uint_as_float(u[i]); }
      } else if (OP == 1) { // 4 x PRMT
#pragma unroll
        for (int r = 0; r < 4; ++r) { asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u[i]) : "r"(u[i]), "r"(__float_as_uint(b[i]))); }
      } else if (OP == 2) { // 4 x LOP3 (and)
#pragma unroll
        for (int r = 0; r < 4; ++r) { asm volatile("and.b32 %0, %1, %2;" : "=r"(u[i]) : "r"(u[i]), "r"(0xffff0000u + r)); }
      } else if (OP == 3) { // the split2 sequence: F2FP, SHL, LOP3, FADD2 (as fma x2 scalar here), F2FP
        unsigned hi, lo;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b[i]), "f"(a[i]));
        float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
        float l0 = a[i] - h0, l1 = b[i] - h1;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(l1), "f"(l0));
        a[i] = __uint_as_float(lo) + l0; b[i] = __uint_as_float(hi) + l1;
      } else if (OP == 4) { // truncation split: 2 LOP3, PRMT (hi), 2 FADD, PRMT (lo, truncated)
        unsigned m0 = __float_as_uint(a[i]) & 0xffff0000u, m1 = __float_as_uint(b[i]) & 0xffff0000u, hi, lo;
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(hi) : "r"(__float_as_uint(a[i])), "r"(__float_as_uint(b[i])));
        float l0 = a[i] - __uint_as_float(m0), l1 = b[i] - __uint_as_float(m1);
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(lo) : "r"(__float_as_uint(l0)), "r"(__float_as_uint(l1)));
        a[i] = __uint_as_float(lo) + l0; b[i] = __uint_as_float(hi) + l1;
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + b[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (float)(t1 - t0);
}
template <int OP>
void run(const char* name, int w, int n_per_step) {
  const int threads = w * 4 * 32, iters = 2000;
  float* out; cudaMalloc(&out, (148 * threads + 1) * sizeof(float));
  k<OP><<<148, threads>>>(out, iters, 1.0f); cudaDeviceSynchronize();
  k<OP><<<148, threads>>>(out, iters, 1.0f); cudaDeviceSynchronize();
  float cyc; cudaMemcpy(&cyc, out + 148 * threads, 4, cudaMemcpyDeviceToHost);
  printf("%-44s warps/sched %d: %.2f cycles per chain-step per scheduler (%d listed ops)\n", name, w, cyc / iters / 8.0 / w, n_per_step);
  cudaFree(out);
}
int main() {
  for (int w : {2, 4}) {
    if (w == 2) { run<0>("4 F2FP.BF16.PACK_AB", 2, 4); run<1>("4 PRMT", 2, 4); run<2>("4 LOP3", 2, 4); run<3>("split2 (rounded hi, rounded lo)", 2, 7); run<4>("truncation split (PRMT)", 2, 6); }
    else        { run<0>("4 F2FP.BF16.PACK_AB", 4, 4); run<1>("4 PRMT", 4, 4); run<2>("4 LOP3", 4, 4); run<3>("split2 (rounded hi, rounded lo)", 4, 7); run<4>("truncation split (PRMT)", 4, 6); }
  }
  return 0;
}
