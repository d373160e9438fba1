// pf_ffn_tc.cuh -- column apply + LayerNorm + FFN + residual on the 5th-gen tensor cores.
//
//   x2 = x1 + M_l qhat + bo ;  x3 = x2 + W2 gelu(W1 LN(x2) + b1) + b2       (model.py:97-104)
//
// The FFN (64 -> 256 -> 64 per token) is the only dense contraction on the path.  It runs as
// two tcgen05.mma GEMMs per 128-token tile with fp32 accumulators in TMEM:
//   GEMM1  D1[128x256] = A1[128x64]  . W1^T      A1 = LN(x2) from shared memory   (SS form)
//   GEMM2  D2[128x64]  = H [128x256] . W2^T      H  = gelu(D1 + b1) from TMEM     (TS form)
// Parity mode ("bf16x3") splits every fp32 operand into bf16 hi + lo and issues the three
// products hi.hi + hi.lo + lo.hi (9e-5 max-rel end to end); fast mode
// ("bf16") issues hi.hi only.
//
// Data movement: both weight matrices (hi and lo, 128 KB) are staged once per CTA into
// shared memory as pre-swizzled (SWIZZLE_128B, K-major) UMMA images built on the host; the
// kernel is persistent (one CTA per SM) so they are read from L2 once per SM.  The hidden
// activations never leave the SM: the epilogue of GEMM1 rewrites D1's TMEM columns in place
// with packed bf16 hi/lo pairs, which GEMM2 consumes as its A operand straight from TMEM.
// HBM traffic is one fp32 read and one fp32 write of the token (512 B).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "pf_common.cuh"

#define TC_TILE 128
#define TC_THREADS 512

struct PfFfnTcW {
  // UMMA K-major SWIZZLE_128B images (see umma_off_*): bf16 bit patterns
  uint16_t w1hi[PF_HID * PF_D];  // B of GEMM1: [n = hidden 256][k = 64]
  uint16_t w1lo[PF_HID * PF_D];
  uint16_t w2hi[PF_D * PF_HID];  // B of GEMM2: [n = out 64][k = hidden 256], 4 K-atoms of 64
  uint16_t w2lo[PF_D * PF_HID];
  float b1[PF_HID];
  float b2[PF_D];
};

// Byte offset of element (row, k) in a K-major SWIZZLE_128B operand whose K extent is 64
// (one 128-byte swizzle atom per row): 8-row groups of 1024 B, 16-byte chunk index XORed with
// the row index inside the group (Swizzle<3,4,3>).
__host__ __device__ inline uint32_t umma_off_k64(int row, int k) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ row) & 7) << 4) + (k & 7) * 2);
}
// Same for K = 256 with `rows` rows: four K-atoms, each a (rows x 128 B) block.
__host__ __device__ inline uint32_t umma_off_k256(int row, int k, int rows) {
  return (uint32_t)((k >> 6) * rows * 128) + umma_off_k64(row, k & 63);
}

inline uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

inline uint16_t f32_to_f16_rn(float f) {   // host: cuda_fp16.h conversions are host-callable
  const __half h = __float2half_rn(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
inline float f16_to_f32(uint16_t u) {
  __half h;
  memcpy(&h, &u, 2);
  return __half2float(h);
}

// Weight images for the tensor-core FFN: 16-bit hi/lo parts (bf16, or fp16 when `f16`), already in
// the UMMA shared-memory layout.
// fold63: the LayerNorm output sums to zero over its 64 channels, so channel 63 is redundant:
//   sum_c W[n][c] u_c = sum_{c<63} (W[n][c] - W[n][63]) u_c.
// The image then holds W[n][c] - W[n][63] in columns 0..62 and the bias b1[n] in column 63; the kernel writes the
// constant 1 into operand column 63, the GEMM delivers W1 LN(x) + b1, and the epilogue has no bias to add (b1 = 0 here).
inline void pf_pack_ffn_tc(const PfFfnW& f, PfFfnTcW* o, bool f16 = false, bool fold63 = false) {
  auto enc = [&](float w) { return f16 ? f32_to_f16_rn(w) : f32_to_bf16_rn(w); };
  auto dec = [&](uint16_t u) { return f16 ? f16_to_f32(u) : bf16_to_f32(u); };
  for (int n = 0; n < PF_HID; ++n)
    for (int k = 0; k < PF_D; ++k) {
      double wd = f.w1T[k][n];
      if (fold63) wd = (k == PF_D - 1) ? (double)f.b1[n] : wd - (double)f.w1T[PF_D - 1][n];
      const float w = (float)wd;
      const uint16_t hi = enc(w);
      const uint16_t lo = enc((float)(wd - (double)dec(hi)));
      const uint32_t off = umma_off_k64(n, k) / 2;
      o->w1hi[off] = hi;
      o->w1lo[off] = lo;
    }
  for (int n = 0; n < PF_D; ++n)
    for (int k = 0; k < PF_HID; ++k) {
      const float w = f.w2T[k][n];
      const uint16_t hi = enc(w);
      const uint16_t lo = enc(w - dec(hi));
      const uint32_t off = umma_off_k256(n, k, PF_D) / 2;
      o->w2hi[off] = hi;
      o->w2lo[off] = lo;
    }
  memcpy(o->b1, f.b1, sizeof(o->b1));
  if (fold63) memset(o->b1, 0, sizeof(o->b1));
  memcpy(o->b2, f.b2, sizeof(o->b2));
}

// ---- shared memory carve-up (offsets from a 1024-byte aligned base) -------------------------
#define TC_OFF_W1HI 0
#define TC_OFF_W1LO 32768
#define TC_OFF_W2HI 65536
#define TC_OFF_W2LO 98304
#define TC_OFF_A1HI 131072
#define TC_OFF_A1LO 147456
#define TC_OFF_XS 163840   // 2 x [128][64] fp32 residual / output staging, chunk-swizzled
#define TC_OFF_B1 229376
#define TC_OFF_B2 230400
#define TC_OFF_BAR 230656  // 2 mbarriers
#define TC_OFF_TMEM 230672
#define TC_SMEM_BYTES (230720 + 1024)
#define TC_TMEM_COLS 512
#define TC_COL_D2 256

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Bounded wait: returns false if the barrier did not complete (the caller raises an error flag
// instead of hanging the GPU).
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2" PF_WAIT_HINT_STR ";\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor: start>>4 | LBO(unused)=1 | SBO=1024>>4 |
// version=1 (bit 46) | layout_type=SWIZZLE_128B (2 at bits 61..63)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D=f32 (bit 4), A and B formats at bits 7 and 10 (0 = fp16,
// 1 = bf16), K-major A and B, N>>3 at bit 17, M>>4 at bit 24.
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, bool f16 = false) {
  return (1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// GELU(h) = max(h,0) - t E(t), t = |h| (clamped at 10), E(t) = 0.5 erfc(t/sqrt2) = exp2(-Q(t)) with a
// degree-6 minimax polynomial Q (tools/gelu_fit.py).  fp32 evaluation: max abs error 4.8e-7,
// rms 1.4e-7 over [-60,60] -- an order of magnitude below the bf16x3 product error it feeds.
__device__ __forceinline__ float gelu_fast(float h) {
  const float t = fminf(fabsf(h), 10.0f);
  float p = 3.2904327396e-05f;
  p = fmaf(p, t, -7.6214972445e-04f);
  p = fmaf(p, t, 8.0388012506e-03f);
  p = fmaf(p, t, -5.3315325260e-02f);
  p = fmaf(p, t, -4.5887145465e-01f);
  p = fmaf(p, t, -1.1511568274e+00f);
  p = fmaf(p, t, -9.9999958869e-01f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p));
  return fmaf(-t, e, fmaxf(h, 0.f));
}
#ifdef PF_TC_EXACT_GELU
#define TC_GELU gelu_erf
#else
#define TC_GELU gelu_fast
#endif

// position (in float4 chunks) of logical chunk c of row r in the XS staging buffer
__device__ __forceinline__ int xs_chunk(int r, int c) { return (c & 8) | ((c ^ r) & 7); }

// n_terms: 3 = bf16x3 (hi.hi + hi.lo + lo.hi), 1 = bf16 (hi.hi)
//
// Software pipeline (all 16 warps share every role; one elected thread issues the MMAs):
//   E1(i) -> issue G2(i) -> P(i+1) [hides G2(i)] -> issue G1(i+1) -> E2(i), store(i) [hide G1(i+1)]
// P = prologue (column apply, LN, split -> A1 smem), G1/G2 = the two tcgen05 GEMMs,
// E1 = GELU epilogue in TMEM, E2 = residual epilogue through the XS staging buffer.
__global__ void __launch_bounds__(TC_THREADS, 1)
k_colapply_ffn_tc(const PfAttnW* __restrict__ Wc, const PfFfnTcW* __restrict__ Wt, float* __restrict__ x,
                  const float* __restrict__ colM, int L, int Pl, long long n_tok, int n_terms,
                  int* __restrict__ err_flag, float* __restrict__ dump) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // keep the pointer derived from the __shared__ array so that accesses compile to LDS/STS
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  float* sb1 = reinterpret_cast<float*>(sm + TC_OFF_B1);
  float* sb2 = reinterpret_cast<float*>(sm + TC_OFF_B2);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + TC_OFF_TMEM);
  const uint32_t bar1 = sbase + TC_OFF_BAR, bar2 = sbase + TC_OFF_BAR + 8;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = lane & 7, slot = tid >> 3;  // 64 token slots

  // ---- one-time setup: weights -> smem, barriers, TMEM ----
  {
    const int4* src = reinterpret_cast<const int4*>(Wt->w1hi);  // w1hi,w1lo,w2hi,w2lo are contiguous
    int4* dst = reinterpret_cast<int4*>(sm + TC_OFF_W1HI);
    for (int i = tid; i < 131072 / 16; i += TC_THREADS) dst[i] = src[i];
    if (tid < PF_HID) sb1[tid] = Wt->b1[tid];
    if (tid < PF_D) sb2[tid] = Wt->b2[tid];
  }
  if (tid == 0) {
    mbar_init(bar1, 1);
    mbar_init(bar2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + TC_OFF_TMEM), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // column-attention q weights for this lane's channels (same mapping as the fp32 kernel)
  float wq[4][8];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const float4 a = reinterpret_cast<const float4*>(Wc->wqk[4 + v])[j];
    const float4 c = reinterpret_cast<const float4*>(Wc->wqk[4 + v])[8 + j];
    wq[v][0] = a.x; wq[v][1] = a.y; wq[v][2] = a.z; wq[v][3] = a.w;
    wq[v][4] = c.x; wq[v][5] = c.y; wq[v][6] = c.z; wq[v][7] = c.w;
  }
  const float bq = Wc->bqk[4 + (j >> 1)];

  const uint32_t idesc1 = umma_idesc(128, 256), idesc2 = umma_idesc(128, 64);
  const long long n_tiles = (n_tok + TC_TILE - 1) / TC_TILE;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int row_t = (warp & 3) * 32 + lane;  // this thread's row in the TMEM epilogues
  const int cgrp = warp >> 2;                // this warp's column group (0..3)

  // ---- prologue: column apply, LN, bf16 hi/lo split -> A1 (UMMA layout), x2 -> XS[buf] ----
  auto prologue = [&](long long tile, int buf) {
    float* XS = reinterpret_cast<float*>(sm + TC_OFF_XS + buf * 32768);
#pragma unroll 1
    for (int pass = 0; pass < TC_TILE / 64; ++pass) {
      const int r = pass * 64 + slot;
      const long long tok = tile * TC_TILE + r;
      const bool act = tok < n_tok;
      float x2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x2[i] = 0.f;
      int l = 0, b = 0;
      if (act) {
        l = (int)(tok % L);
        b = (int)(tok / ((long long)L * Pl));
        load_tok(x + (size_t)tok * PF_D, j, x2);
      }
      float nv[8];
      ln_normalize(x2, nv);
      float pr[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(wq[v][i], nv[i], s);
        pr[v] = s;
      }
      const float* cm = colM + ((size_t)b * L + l) * PF_MROW;
      float mine = phi_elu1(grp_reduce4(pr, j) + bq);
      mine *= cm[256 + (j >> 1)];
      float qh[PF_H];
#pragma unroll
      for (int h = 0; h < PF_H; ++h) qh[h] = __shfl_sync(PF_FULL, mine, (lane & 24) | (2 * h));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = chan_of(j, i);
        const float4 m = *reinterpret_cast<const float4*>(cm + c * 4);
        float a = Wc->bo[c];
        a = fmaf(m.x, qh[0], a);
        a = fmaf(m.y, qh[1], a);
        a = fmaf(m.z, qh[2], a);
        a = fmaf(m.w, qh[3], a);
        x2[i] += a;
      }
      float4* xr = reinterpret_cast<float4*>(XS + r * PF_D);
      xr[xs_chunk(r, j)] = make_float4(x2[0], x2[1], x2[2], x2[3]);
      xr[xs_chunk(r, 8 + j)] = make_float4(x2[4], x2[5], x2[6], x2[7]);
      ln_normalize(x2, nv);
      // channels 4j..4j+3 -> 16B chunk j>>1 (half j&1); 32+4j.. -> chunk 4+(j>>1)
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 hh = __floats2bfloat162_rn(nv[2 * i], nv[2 * i + 1]);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(nv[2 * i] - __low2float(hh), nv[2 * i + 1] - __high2float(hh));
        hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
        lo[i] = *reinterpret_cast<const uint32_t*>(&ll);
      }
      const uint32_t rowoff = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
      const uint32_t o0 = rowoff + ((((j >> 1)) ^ (r & 7)) << 4) + (j & 1) * 8;
      const uint32_t o1 = rowoff + (((4 + (j >> 1)) ^ (r & 7)) << 4) + (j & 1) * 8;
      *reinterpret_cast<uint2*>(sm + TC_OFF_A1HI + o0) = make_uint2(hi[0], hi[1]);
      *reinterpret_cast<uint2*>(sm + TC_OFF_A1HI + o1) = make_uint2(hi[2], hi[3]);
      *reinterpret_cast<uint2*>(sm + TC_OFF_A1LO + o0) = make_uint2(lo[0], lo[1]);
      *reinterpret_cast<uint2*>(sm + TC_OFF_A1LO + o1) = make_uint2(lo[2], lo[3]);
    }
  };
  // ---- GEMM1: D1 = A1 . W1^T  (SS, N=256, K=64 in 4 steps), one thread ----
  auto issue_g1 = [&]() {
    tc_fence_after();
    uint32_t acc = 0;
    for (int t = 0; t < n_terms; ++t) {
      const uint32_t a_base = sbase + ((t == 2) ? TC_OFF_A1LO : TC_OFF_A1HI);
      const uint32_t b_base = sbase + ((t == 1) ? TC_OFF_W1LO : TC_OFF_W1HI);
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        umma_ss(tmem, umma_desc(a_base + s * 32), umma_desc(b_base + s * 32), idesc1, acc);
        acc = 1;
      }
    }
    tc_commit(bar1);
  };
  // ---- GEMM2: D2 = H . W2^T  (TS, N=64, K=256 in 16 steps); H hi at column 16s, lo at 16s+8 ----
  auto issue_g2 = [&]() {
    tc_fence_after();
    uint32_t acc = 0;
    for (int t = 0; t < n_terms; ++t) {
      const uint32_t b_base = sbase + ((t == 2) ? TC_OFF_W2LO : TC_OFF_W2HI);
      const uint32_t a_sel = (t == 1) ? 8u : 0u;  // t=1: H_lo . W2_hi
#pragma unroll
      for (int s = 0; s < 16; ++s) {
        const uint32_t b_off = (uint32_t)((s >> 2) * (PF_D * 128) + (s & 3) * 32);
        umma_ts(tmem + TC_COL_D2, tmem + (uint32_t)(16 * s) + a_sel, umma_desc(b_base + b_off), idesc2, acc);
        acc = 1;
      }
    }
    tc_commit(bar2);
  };

  long long tile = blockIdx.x;
  uint32_t phase = 0;
  int buf = 0;
  bool ok = true;
  if (tile < n_tiles) {
    prologue(tile, 0);
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) issue_g1();
  }
  for (; tile < n_tiles; tile += gridDim.x) {
    const long long next = tile + gridDim.x;
    const bool has_next = next < n_tiles;
    const bool dump_this = (dump != nullptr) && (tile == 0);
    ok = mbar_wait(bar1, phase) && ok;
    tc_fence_after();
    // ============ E1: H = gelu(D1 + b1) as bf16 hi/lo, in place in TMEM ============
    {
      const int col0 = cgrp * 64;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int cc = col0 + c * 16;
        uint32_t v[16];
        tmem_ld16(tmem + lane_base + cc, v);
        tc_wait_ld();
        if (dump_this) {
#pragma unroll
          for (int i = 0; i < 16; ++i) dump[row_t * 320 + cc + i] = __uint_as_float(v[i]);
        }
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 bb = *reinterpret_cast<const float2*>(sb1 + cc + 2 * i);
          const float g0 = TC_GELU(__uint_as_float(v[2 * i]) + bb.x);
          const float g1 = TC_GELU(__uint_as_float(v[2 * i + 1]) + bb.y);
          const __nv_bfloat162 hh = __floats2bfloat162_rn(g0, g1);
          const __nv_bfloat162 ll = __floats2bfloat162_rn(g0 - __low2float(hh), g1 - __high2float(hh));
          hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
          lo[i] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        tmem_st8(tmem + lane_base + cc, hi);
        tmem_st8(tmem + lane_base + cc + 8, lo);
      }
      tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) issue_g2();
    // ============ P(i+1) while G2(i) runs ============
    if (has_next) prologue(next, buf ^ 1);
    fence_proxy_async_smem();
    ok = mbar_wait(bar2, phase) && ok;
    tc_fence_after();
    __syncthreads();
    if (has_next && tid == 0) issue_g1();
    // ============ E2: y = x2 + D2 + b2 -> XS[buf], while G1(i+1) runs ============
    float* XS = reinterpret_cast<float*>(sm + TC_OFF_XS + buf * 32768);
    {
      uint32_t v[16];
      tmem_ld16(tmem + lane_base + TC_COL_D2 + cgrp * 16, v);
      tc_wait_ld();
      if (dump_this) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dump[row_t * 320 + 256 + cgrp * 16 + i] = __uint_as_float(v[i]);
      }
      float4* xr = reinterpret_cast<float4*>(XS + row_t * PF_D);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = cgrp * 4 + i;
        const float4 bb = *reinterpret_cast<const float4*>(sb2 + 4 * c);
        float4 t = xr[xs_chunk(row_t, c)];
        t.x += __uint_as_float(v[4 * i + 0]) + bb.x;
        t.y += __uint_as_float(v[4 * i + 1]) + bb.y;
        t.z += __uint_as_float(v[4 * i + 2]) + bb.z;
        t.w += __uint_as_float(v[4 * i + 3]) + bb.w;
        xr[xs_chunk(row_t, c)] = t;
      }
    }
    tc_fence_before();
    __syncthreads();
    // ============ store: coalesced fp32 rows back to HBM ============
#pragma unroll
    for (int pass = 0; pass < TC_TILE / 64; ++pass) {
      const int r = pass * 64 + slot;
      const long long tok = tile * TC_TILE + r;
      if (tok < n_tok) {
        const float4* xr = reinterpret_cast<const float4*>(XS + r * PF_D);
        float4* g = reinterpret_cast<float4*>(x + (size_t)tok * PF_D);
        g[j] = xr[xs_chunk(r, j)];
        g[8 + j] = xr[xs_chunk(r, 8 + j)];
      }
    }
    phase ^= 1;
    buf ^= 1;
  }
  if (!ok && err_flag != nullptr) *err_flag = 1;
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
  }
}

inline int pf_ffn_tc_init() {
  return (int)cudaFuncSetAttribute(k_colapply_ffn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
}

inline int pf_ffn_tc_launch(const PfAttnW* Wc, const PfFfnTcW* Wt, float* x, const float* colM, int L, int Pl,
                            long long n_tok, int n_sm, int n_terms, int* err_flag, float* dump, cudaStream_t st) {
  const long long tiles = (n_tok + TC_TILE - 1) / TC_TILE;
  const int grid = (int)(tiles < n_sm ? tiles : n_sm);
  k_colapply_ffn_tc<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(Wc, Wt, x, colM, L, Pl, n_tok, n_terms, err_flag, dump);
  return (int)cudaGetLastError();
}
