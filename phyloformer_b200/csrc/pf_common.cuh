// pf_common.cuh -- packed-weight layout and the per-token device helpers shared by every
// kernel of the Phyloformer forward path (sm_100a).
//
// Thread <-> data mapping used by all CUDA-core token work ("8 lanes per token"):
//   a warp holds 4 token slots; lane j (0..7) of a slot owns channels 4j..4j+3 and
//   32+4j..32+4j+3 of the token's 64-vector.  A token load is therefore two LDG.128 whose 8
//   lanes cover one full 128-byte line each, and every reduction over the 64 channels is a
//   3-step xor butterfly inside the 8-lane group.  The reduction tree is the same for every
//   token, so a token's result does not depend on where it sits in a tile or which pair it
//   belongs to: identical sequences give bit-identical rows.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Suspend-time hint (ns) of every bounded mbarrier.try_wait in this library; PF_WAIT_HINT=0 builds the waits without a
// hint (the hardware's own time limit per try).  A/B at the end of round 2: see profiles/r02_experiments.md.
#ifndef PF_WAIT_HINT
#define PF_WAIT_HINT 1000
#endif
#define PF_STR2(x) #x
#define PF_STR(x) PF_STR2(x)
#if PF_WAIT_HINT > 0
#define PF_WAIT_HINT_STR ", " PF_STR(PF_WAIT_HINT)
#else
#define PF_WAIT_HINT_STR ""
#endif

#define PF_D 64
#define PF_H 4
#define PF_DH 16
#define PF_HID 256
#define PF_NCHAR 22
#define PF_COLSUM 72   // exchange layout: sum k~ (4) | sum q~ (4) | sum k~ v (64)
#define PF_PART 264    // local partial:   sum k~ (4) | sum q~ (4) | sum k~ n (4 x 64), n = LN without affine
#define PF_FS 8        // sites per CTA in the column reduce / finalize kernels
#define PF_MROW 260    // applied form:    M[64][4] | qinv[4]
#define PF_FULL 0xffffffffu

// One attention module together with the LayerNorm in front of it.
// q/k projections have the LN affine folded in:  w.(g*n + b) + c = (w*g).n + (w.b + c).
struct PfAttnW {
  float wqk[8][PF_D];    // rows 0..3: k heads, 4..7: q heads (attention.py:163-172), gamma folded
  float bqk[8];
  float gamma[PF_D], beta[PF_D];
  float wvT[PF_D][PF_D]; // wvT[k][o] = v_proj.weight[o][k]
  float bv[PF_D];
  float wo[PF_D][PF_D];  // out_proj.weight[c][i]
  float bo[PF_D];
};
struct PfFfnW {
  float w1T[PF_D][PF_HID]; // w1T[k][j] = ffn.0.weight[j][k] * ffn_norm.weight[k]
  float b1[PF_HID];        // ffn.0.bias + ffn.0.weight . ffn_norm.bias
  float w2T[PF_HID][PF_D]; // w2T[j][c] = ffn.3.weight[c][j]
  float b2[PF_D];
};
struct PfBlockW {
  PfAttnW row, col;
  PfFfnW ffn;
};
struct PfHeadW {
  float table[PF_NCHAR][PF_D]; // relu(W_e[:,a] + b_e): embedding of a one-hot residue (model.py:138-143)
  float weT[PF_NCHAR][PF_D];   // weT[a][c] = embedding_block.0.weight[c][a]   (soft inputs)
  float be[PF_D];
  float whead[PF_D];           // pwFNN.0.weight
  float bhead;
  float pad[3];
};

struct Tok {
  float v[8];
};

__device__ __forceinline__ float grp_sum(float x) {
  x += __shfl_xor_sync(PF_FULL, x, 1);
  x += __shfl_xor_sync(PF_FULL, x, 2);
  x += __shfl_xor_sync(PF_FULL, x, 4);
  return x;
}

// Sum 8 per-lane partials over the 8-lane group; lane j ends up with the total of value j.
__device__ __forceinline__ float grp_reduce8(const float (&p)[8], int j) {
  const bool hi = (j & 4) != 0, mid = (j & 2) != 0, lo = (j & 1) != 0;
  float a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = hi ? p[i] : p[i + 4];
    const float keep = hi ? p[i + 4] : p[i];
    a[i] = keep + __shfl_xor_sync(PF_FULL, send, 4);
  }
  float b[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = mid ? a[i] : a[i + 2];
    const float keep = mid ? a[i + 2] : a[i];
    b[i] = keep + __shfl_xor_sync(PF_FULL, send, 2);
  }
  const float send = lo ? b[0] : b[1];
  const float keep = lo ? b[1] : b[0];
  return keep + __shfl_xor_sync(PF_FULL, send, 1);
}

// Sum 4 per-lane partials over the group; lanes j and j^1 end up with the total of value j>>1.
__device__ __forceinline__ float grp_reduce4(const float (&p)[4], int j) {
  const bool hi = (j & 4) != 0, mid = (j & 2) != 0;
  float a[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = hi ? p[i] : p[i + 2];
    const float keep = hi ? p[i + 2] : p[i];
    a[i] = keep + __shfl_xor_sync(PF_FULL, send, 4);
  }
  const float send = mid ? a[0] : a[1];
  const float keep = mid ? a[1] : a[0];
  float r = keep + __shfl_xor_sync(PF_FULL, send, 2);
  r += __shfl_xor_sync(PF_FULL, r, 1);
  return r;
}

// 1/sqrt(v): MUFU.RSQ + one Newton step (~1 ulp), instead of the IEEE div/sqrt slow path
__device__ __forceinline__ float rsqrt_nr(float v) {
  const float y = rsqrtf(v);
  return y * fmaf(-0.5f * v * y, y, 1.5f);
}

// LayerNorm without the affine part: n = (x - mean) * rstd, biased variance, eps = 1e-5
// (nn.LayerNorm, model.py:64-66).
// FAST_RSQRT picks rsqrt_nr over the IEEE 1/sqrt; which one is faster is kernel-dependent
// (measured A/B on B200: row kernel -14 % with rsqrt_nr, column-partial kernel +25 %), both are
// accurate to ~1 ulp.
template <bool FAST_RSQRT = false>
__device__ __forceinline__ void ln_normalize(const float (&x)[8], float (&n)[8]) {
  float s = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
  s = grp_sum(s);
  const float mean = s * (1.0f / PF_D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    n[i] = x[i] - mean;
    q = fmaf(n[i], n[i], q);
  }
  q = grp_sum(q);
  const float var = fmaf(q, 1.0f / PF_D, 1e-5f);
  const float rstd = FAST_RSQRT ? rsqrt_nr(var) : 1.0f / sqrtf(var);
#pragma unroll
  for (int i = 0; i < 8; ++i) n[i] *= rstd;
}

// phi(z) = elu(z) + 1  (attention.py:179-180)
__device__ __forceinline__ float phi_elu1(float z) { return z > 0.f ? z + 1.0f : expf(z); }

// nn.Softplus(beta=1, threshold=20)  (model.py:163)
__device__ __forceinline__ float softplus20(float z) { return z > 20.f ? z : log1pf(expf(z)); }

// nn.GELU() exact form
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ void load_tok(const float* __restrict__ row, int j, float (&x)[8]) {
  const float4 a = reinterpret_cast<const float4*>(row)[j];
  const float4 b = reinterpret_cast<const float4*>(row)[8 + j];
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
  x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void store_tok(float* __restrict__ row, int j, const float (&x)[8]) {
  reinterpret_cast<float4*>(row)[j] = make_float4(x[0], x[1], x[2], x[3]);
  reinterpret_cast<float4*>(row)[8 + j] = make_float4(x[4], x[5], x[6], x[7]);
}
// channel index of register i of lane j
__device__ __forceinline__ int chan_of(int j, int i) { return (i < 4) ? (4 * j + i) : (32 + 4 * j + (i - 4)); }

// ---------------------------------------------------------------------------------------------
// Packed-fp32 helpers (FFMA2/FMUL2/FADD2, sm_100): two fp32 values in one 64-bit register.  Used by
// the FFN kernel's GELU and LayerNorm statistics; the streaming kernels measured slower with them.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long pf_u64;
__device__ __forceinline__ pf_u64 pk2(float a, float b) { pf_u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(pf_u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ pf_u64 fma2(pf_u64 a, pf_u64 b, pf_u64 c) { pf_u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ pf_u64 mul2(pf_u64 a, pf_u64 b) { pf_u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ pf_u64 add2(pf_u64 a, pf_u64 b) { pf_u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float hsum2(pf_u64 v) { float a, b; up2(v, a, b); return a + b; }

// Lexicographic pair index -> (i, j), i < j < n   (model.py:13-17 order)
__host__ __device__ inline void pair_to_ij(long long p, int n, int* pi, int* pj) {
  const double b = 2.0 * n - 1.0;
  long long i = (long long)((b - sqrt(b * b - 8.0 * (double)p)) * 0.5);
  if (i < 0) i = 0;
  // first pair index of row i: i*n - i(i+1)/2
  while (i > 0 && (i * n - i * (i + 1) / 2) > p) --i;
  while (((i + 1) * n - (i + 1) * (i + 2) / 2) <= p) ++i;
  const long long start = i * n - i * (i + 1) / 2;
  *pi = (int)i;
  *pj = (int)(i + 1 + (p - start));
}
