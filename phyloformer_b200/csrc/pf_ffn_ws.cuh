// pf_ffn_ws.cuh -- warp-specialised tcgen05 kernel for column apply + LayerNorm + FFN + residual.
//
//   x2 = x1 + M_l qhat + bo ;  x3 = x2 + W2 gelu(W1 LN(x2) + b1) + b2       (model.py:97-104)
//
// Same math and the same UMMA operand images as pf_ffn_tc.cuh; what changes is who does what:
//
//   warps EW..EW+3   PRODUCER  one thread per token row (row r <-> TMEM lane r).  Loads the row
//               (16 x LDG.128), applies column attention with the per-site M_l window held in
//               shared memory, writes x2 + b2 straight into the GEMM2 accumulator in TMEM (the
//               residual add is then free: GEMM2 accumulates on top of it), LayerNorms, splits
//               into bf16 hi/lo and stores the GEMM1 A operand (SWIZZLE_128B) in shared memory.
//   warp  EW+4       MMA       one elected thread issues every tcgen05.mma / tcgen05.commit.
//   warps 0..EW-1    EPILOGUE (EW = 8 or 16)  GELU on the GEMM1 accumulator in TMEM, rewritten in place as packed bf16
//               hi/lo (FFMA2/FMUL2 packed-fp32 math); final rows TMEM -> HBM.
//
// A tile is WS_G pairs x WS_S consecutive sites (128 rows; 32 x 4 by default) so that the WS_S-site
// window of M_l stays resident in shared memory while a CTA walks down the pairs; CTAs own
// contiguous tile ranges, window-major.  The hidden layer is processed in two 128-unit halves
// with separate TMEM buffers and barriers, the A operand and the GEMM2 accumulator are double
// buffered, so the tensor pipe, the producer and the epilogue run concurrently:
//
//   TMEM columns   [0,128) D1 half a   [128,256) D1 half b   [256,448) D2, 3 slots of 64 columns
//   MMA order      G2a(i)  G1a(i+1)  G2b(i)  G1b(i+1)        (in-order issue covers the WAR on D1)
//   epilogue order E1a(i)  E2(i-1)  E1b(i)
#pragma once

#include "pf_ffn_tc.cuh"

// WS_EW = number of epilogue warps: 8 (512 threads, 128 registers each) or 16 (768 threads launched
// at 80 registers; setmaxnreg then gives the producer warpgroup 136 and shrinks the MMA warpgroup
// to 24 (the CTA pool is 768 x 80: 16x32x80 + 4x32x136 + 4x32x24) -- more epilogue warps to hide MUFU/LDS/dependency latency).
#ifndef WS_EW
#define WS_EW 8
#endif
// Variants that were built, measured and removed again (DESIGN.md section 3.1 has the numbers):
// final-row stores by the MMA warpgroup's idle warps, two producer threads per token row
// (8 producer warps), epilogue warps above the producer in warp-id order.
// 12 epilogue warps inside the 16-warp CTA (round 2: the MMA warpgroup's issuer and its three idle warps join the
// epilogue as a third column group, chunks split 3 / 3 / 2, the issuer issuing each GEMM as soon as its inputs are
// there and never blocking on the producer; no spills): parity green, 36.6 vs 32.5 ms -- the time per 16-column chunk
// rises by the same 50 % as the number of epilogue warps per scheduler, i.e. a scheduler's two epilogue warps already
// take everything it can issue for this instruction mix (packed FFMA2 / half-rate ALU / MUFU); more warps cannot help.
// (Re-measured after b1 moved to the constant bank: 36.4 vs 29.8 ms.  The issuer now also loses time whenever the
// producer is the late party: a dedicated MMA warp blocks on a1_full and issues at once, a working one finds out late.)
// Final rows (E2) on the MMA warpgroup instead of the epilogue warps (warps 13..15 for TMEM lane quadrants 1..3, the
// issuer for quadrant 0 right after a tile's last issue): parity green, 31.2 vs 30.1 ms -- E1 slows down by what E2
// took: the work per SM is conserved whichever warps do it (TMEM / MIO path, not warp latency).
// 16 epilogue warps at <= 80 registers with a row-chunked producer (x2 staged through TMEM, 12 extra TMEM
// instructions per row): built and measured in round 2, 34.7 vs 32.5 ms of FFN time per forward -> removed.
// Final-row stores (E2) after E1b instead of between E1a and E1b -- for column group 1 only (so that on every scheduler
// one epilogue warp waits on TMEM / stores while the other runs GELU) or for both groups (E2 off the E1a -> E1b -> G2b
// path): parity green, FFN time per forward 29.9 / 31.0 vs 28.97 ms (same box, back to back) -> removed.
// WS_B1_CONST = 1: the epilogue reads b1 through the constant bank (LDC) instead of shared memory, whose loads queue
// behind the previous chunk's tcgen05.st in the MIO queue
#ifdef WS_DIAG_NO_LDTM   // timing diagnostic only (wrong results): no TMEM reads in E1
#define E1_LD(addr, arr) do { for (int k_ = 0; k_ < 16; ++k_) (arr)[k_] = (uint32_t)(addr) + k_; } while (0)
#else
#define E1_LD(addr, arr) tmem_ld16(addr, arr)
#endif
#ifndef WS_ST16
#define WS_ST16 1   // one tcgen05.st.x16 per chunk (hi | lo words are adjacent columns) instead of two .x8
#endif
#ifndef WS_B1_CONST
#define WS_B1_CONST 1
#endif
// WS_FOLD63 = 1: the first bias rides on GEMM1.  LN(x) sums to zero over its 64 channels, so operand column 63 is
// redundant (pf_pack_ffn_tc folds its weight into the other 63 columns); the producer writes the constant 1 there and
// the weight image carries b1 in that column: D1 = W1 LN(x) + b1 comes out of the tensor core and the epilogue saves the
// bias load (LDC.64) and the packed add per pair of hidden units.  Needs the folded image (pf_handle::tcf_dev).
// Parity format (bf16 hi/lo) only: with single-rounded operands the larger folded weights and the rounded bias cost
// accuracy (fp16 mode 4.3e-3 -> 7.3e-3, bf16 mode 2.9e-2 -> 4.0e-2 max-rel on the 20 test alignments), so the fast
// formats keep the unfolded images and the epilogue's bias add.
#ifndef WS_FOLD63
#define WS_FOLD63 1
#endif
#define WS_NCG (WS_EW / 4)            // epilogue column groups per TMEM lane quadrant
#define WS_NPW 4                      // producer warps
#define WS_PW0 WS_EW                  // first producer warp
#define WS_EW0 0                      // first epilogue warp
#define WS_MW (WS_EW + WS_NPW)        // MMA warp (highest warp ids: highest issue priority)
#define WS_THREADS ((WS_EW + WS_NPW + 4) * 32)
#ifndef WS_S_LOG2
#define WS_S_LOG2 2                 // log2(sites per tile): 32 pairs x 4 sites.  Measured 16 / 8 / 4 / 2 / 1 sites:
                                    // 32.9 / 32.7 / 32.1 / 31.9 / 32.5 ms (fewer distinct M_l rows per warp = fewer
                                    // shared-memory wavefronts); 4 keeps 1 KB contiguous per pair and little tail waste
#endif
#define WS_S (1 << WS_S_LOG2)       // sites per tile
#define WS_G (128 / WS_S)           // pairs per tile (128 token rows = the MMA's M)
#define WS_OFF_A1 131072            // GEMM1 A operand: hi 16 KB + lo 16 KB (single buffer)
#define WS_XROW 272                 // staged row stride: 256 B + 16 B pad (conflict-free LDS.128 per row)
#define WS_OFF_XST 163840           // [128] rows prefetched with cp.async one tile ahead
#define WS_OFF_MWIN 198656          // [16][260] floats
#define WS_MWIN_BYTES (WS_S * PF_MROW * 4)
#define WS_OFF_XCH (WS_OFF_MWIN + 16896)  // 8 KB, unused (kept so the offsets below do not move)
#define WS_OFF_B1 (WS_OFF_XCH + 8192)     // [256]
#define WS_OFF_BAR (WS_OFF_B1 + 1024)     // 14 mbarriers
#define WS_OFF_TMEM (WS_OFF_BAR + 112)
#define WS_SMEM_BYTES (WS_OFF_TMEM + 32 + 1024)
#define WS_COL_D2 256

// Small per-block parameters the producer reads with compile-time indices: passed by value as a
// __grid_constant__ kernel parameter so that they sit in the constant bank and feed FFMA
// operands directly (no shared-memory loads, no registers).
struct PfFfnConst {
  float wq[PF_D][PF_H];  // folded column-attention q weights [channel][head]
  float bo[PF_D];        // column out_proj bias
  float b2[PF_D];        // ffn.3 bias
  float bq[PF_H];        // folded q bias
  float whead[PF_D];     // pwFNN.0.weight (HEAD instantiation: the last block's launch emits distances)
  float bhead;
  float pad[3];
  float b1[PF_HID];      // folded ffn.0 bias: read by the epilogue through the constant bank (WS_B1_CONST)
};

typedef pf_u64 u64;  // packed fp32 pair helpers (pk2, up2, fma2, mul2, add2) live in pf_common.cuh

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void ldg256(const float* p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// GELU on two values at once with packed fp32 math; returns (g0, g1).
//   GELU(h) = max(h,0) - t E(t),  t = |h|,  E(t) = 0.5 erfc(t/sqrt2) = exp2(-Q(t))
// with Q a degree-6 minimax polynomial on [0,6] (weighted by t E(t), tools/gelu_fit.py); t is
// clamped at 10 where t E(t) < 1e-20.  fp32 evaluation: max abs error 4.8e-7, rms 1.4e-7.
// Per pair: 6 FFMA2 + 2 FMNMX + 2 MUFU.EX2 + 1 FFMA2 + 2 FMNMX (vs 17 for the 1/p^16 form).
__device__ __forceinline__ u64 gelu_fast2(float h0, float h1) {
#ifdef WS_DIAG_NO_GELU   // timing diagnostic only (wrong results)
  return pk2(h0, h1);
#endif
  const float t0 = fminf(fabsf(h0), 10.0f), t1 = fminf(fabsf(h1), 10.0f);
  const u64 t = pk2(t0, t1);
  u64 p = pk2(3.2904327396e-05f, 3.2904327396e-05f);            // coefficients of -Q(t), high to low
  p = fma2(p, t, pk2(-7.6214972445e-04f, -7.6214972445e-04f));
  p = fma2(p, t, pk2(8.0388012506e-03f, 8.0388012506e-03f));
  p = fma2(p, t, pk2(-5.3315325260e-02f, -5.3315325260e-02f));
  p = fma2(p, t, pk2(-4.5887145465e-01f, -4.5887145465e-01f));
  p = fma2(p, t, pk2(-1.1511568274e+00f, -1.1511568274e+00f));
  p = fma2(p, t, pk2(-9.9999958869e-01f, -9.9999958869e-01f));
  float p0, p1, e0, e1;
  up2(p, p0, p1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(p0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(p1));
  return fma2(pk2(-t0, -t1), pk2(e0, e1), pk2(fmaxf(h0, 0.f), fmaxf(h1, 0.f)));
}

// Operand formats of the two GEMMs' A operands (template parameter FMT):
//   WS_FMT_BF16X3  bf16 hi + lo parts (3-term product with the hi/lo weight images)
//   WS_FMT_BF16    bf16, one rounding, no lo part (fast mode)
//   WS_FMT_F16     fp16, one rounding, no lo part; the weights keep hi + lo (2-term product)
#define WS_FMT_BF16X3 0
#define WS_FMT_BF16 1
#define WS_FMT_F16 2

// bf16 hi/lo split of a packed pair: hi = rn(g), lo = rn(g - hi); both as packed bf16x2 words.
__device__ __forceinline__ void split2(u64 g, uint32_t& hi, uint32_t& lo) {
#ifdef WS_DIAG_NO_SPLIT  // timing diagnostic only (wrong results)
  { float a, b; up2(g, a, b); hi = __float_as_uint(a); lo = __float_as_uint(b); return; }
#endif
  float g0, g1;
  up2(g, g0, g1);
  const __nv_bfloat162 hh = __floats2bfloat162_rn(g0, g1);
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  const u64 hf = pk2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u));
  float l0, l1;
  up2(fma2(hf, pk2(-1.f, -1.f), g), l0, l1);
  const __nv_bfloat162 ll = __floats2bfloat162_rn(l0, l1);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}
// One packed pair -> the 16-bit operand word(s) of format FMT (lo is written only for BF16X3).
template <int FMT>
__device__ __forceinline__ void cvt2(u64 g, uint32_t& hi, uint32_t& lo) {
  if (FMT == WS_FMT_BF16X3) {
    split2(g, hi, lo);
  } else {
    float g0, g1;
    up2(g, g0, g1);
    if (FMT == WS_FMT_F16) {
      const __half2 hh = __floats2half2_rn(g0, g1);
      hi = *reinterpret_cast<const uint32_t*>(&hh);
    } else {
      const __nv_bfloat162 hh = __floats2bfloat162_rn(g0, g1);
      hi = *reinterpret_cast<const uint32_t*>(&hh);
    }
  }
}

struct WsTileMap {  // tile index -> rows
  int L, Pl, nW, nPG;
  __device__ __forceinline__ void decode(unsigned t, int& b, int& w, int& pg) const {  // 32-bit: no emulated 64-bit division
    const unsigned bw = t / (unsigned)nPG;
    pg = (int)(t - bw * (unsigned)nPG);
    b = (int)(bw / (unsigned)nW);
    w = (int)(bw - (unsigned)b * (unsigned)nW);
  }
};

// QC: the column-attention q~ of every token comes from the cache written by k_col_partial_tc (16 B per
// token, staged into the pad of the row slot) instead of being recomputed (LN_col + 4 dots of length 64).
// HEAD: the launch of the last block does not write the activation back; the epilogue applies the distance head
// (model.py:182-185: dot with pwFNN.0.weight, softplus) to the final row straight out of TMEM and writes one
// partial site sum per (pair, 4-site window) into `dump` (there: headpart[b][window][pair]); k_head_reduce adds the
// windows in a fixed order.  Saves the last 256 B/token write and k_head's 256 B/token read.
template <bool PROF, int FMT, bool QC, bool HEAD = false>
__global__ void __launch_bounds__(WS_THREADS, 1)
k_colapply_ffn_ws(const __grid_constant__ PfFfnConst kc, const PfFfnTcW* __restrict__ Wt, float* __restrict__ x,
                  const float* __restrict__ colM, const float* __restrict__ qcache, int L, int Pl, int B, int n_terms,
                  int* __restrict__ err_flag, float* __restrict__ dump) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // keep the pointer derived from the __shared__ array so that accesses compile to LDS/STS
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  float* mwin = reinterpret_cast<float*>(sm + WS_OFF_MWIN);
  float* sb1 = reinterpret_cast<float*>(sm + WS_OFF_B1);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + WS_OFF_TMEM);
  // barriers: [0,1] a1_full  [2,3] a1_free  [4..6] d2_free  [7,8] g1_done  [9,10] h_full  [11..13] g2_done
  const uint32_t bars = sbase + WS_OFF_BAR;
  auto BAR = [&](int i) { return bars + 8u * (uint32_t)i; };

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- one-time setup ----
  {
    const int4* src = reinterpret_cast<const int4*>(Wt->w1hi);
    int4* dst = reinterpret_cast<int4*>(sm);
    for (int i = tid; i < 131072 / 16; i += WS_THREADS) dst[i] = src[i];
    if (tid < PF_HID) sb1[tid] = Wt->b1[tid];
  }
  if (tid == 0) {
    mbar_init(BAR(0), WS_NPW * 32); mbar_init(BAR(1), WS_NPW * 32);
    mbar_init(BAR(2), 1);   mbar_init(BAR(3), 1);
    mbar_init(BAR(4), WS_EW * 32); mbar_init(BAR(5), WS_EW * 32);
    mbar_init(BAR(6), WS_EW * 32);
    mbar_init(BAR(7), 1);   mbar_init(BAR(8), 1);
    mbar_init(BAR(9), WS_EW * 32); mbar_init(BAR(10), WS_EW * 32);
    mbar_init(BAR(11), 1);  mbar_init(BAR(12), 1);  mbar_init(BAR(13), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WS_MW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + WS_OFF_TMEM), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  WsTileMap tm;
  tm.L = L; tm.Pl = Pl; tm.nW = (L + WS_S - 1) / WS_S; tm.nPG = (Pl + WS_G - 1) / WS_G;
  const long long NT = (long long)B * tm.nW * tm.nPG;
  const long long t_begin = NT * blockIdx.x / gridDim.x;
  const long long t_end = NT * (blockIdx.x + 1) / gridDim.x;
  const int n_my = (int)(t_end - t_begin);
  bool ok = true;
  // optional role timing (test hook PF_WS_PROF=1): cycles spent in each wait, per role leader
  long long tw[7] = {0, 0, 0, 0, 0, 0, 0};
  const long long t_start = PROF ? clock64() : 0;
  auto TIC = [&]() -> long long { return PROF ? clock64() : 0; };
  auto TOC = [&](int slot, long long t0) { if (PROF) tw[slot] += clock64() - t0; };
  auto WAIT = [&](int slot, uint32_t bar, uint32_t parity) {
    if (PROF) {
      const long long t0 = clock64();
      ok = mbar_wait(bar, parity) && ok;
      tw[slot] += clock64() - t0;
    } else {
      ok = mbar_wait(bar, parity) && ok;
    }
  };

  if (warp >= WS_PW0 && warp < WS_PW0 + WS_NPW) {
#if WS_EW == 16
    asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");   // producer warpgroup takes the registers the MMA group returns
#elif WS_EW == 12
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
#endif
    // =============================== PRODUCER ===============================================
    const int ptid = tid - WS_PW0 * 32;  // 0..127
    const int r = ptid;                // row == TMEM lane
    const int g = r >> WS_S_LOG2, s = r & (WS_S - 1);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    int cur_b = -1, cur_w = -1;
    int b, w, pg;
    tm.decode((unsigned)t_begin, b, w, pg);
    const uint32_t xst_u32 = sbase + WS_OFF_XST + (uint32_t)r * WS_XROW;
    auto prefetch_row = [&](size_t tok) {  // 16 x 16-byte cp.async: global row -> this thread's smem row
      const float* src = x + tok * PF_D;
#pragma unroll
      for (int c = 0; c < 16; ++c)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(xst_u32 + 16 * c), "l"(src + 4 * c) : "memory");
      if (QC) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(xst_u32 + 256), "l"(qcache + tok * 4) : "memory");
    };
    if (n_my > 0) {
      const int pair0 = pg * WS_G + g, site0 = w * WS_S + s;
      if (pair0 < Pl && site0 < L) prefetch_row(((size_t)b * Pl + pair0) * L + site0);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int it = 0; it < n_my; ++it) {
      const uint32_t par = (uint32_t)(it & 1);
      if (it > 0 && ++pg == tm.nPG) { pg = 0; if (++w == tm.nW) { w = 0; ++b; } }   // next tile, no division
      // ---- this tile's row was prefetched into smem by cp.async one tile ago ----
      const int pair = pg * WS_G + g, site = w * WS_S + s;
      const bool valid = (pair < Pl) && (site < L);
      float xr[PF_D];
      float4 qc4 = make_float4(0.f, 0.f, 0.f, 0.f);
      {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        const float4* srow = reinterpret_cast<const float4*>(sm + WS_OFF_XST + r * WS_XROW);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) v = srow[c];
          xr[4 * c] = v.x; xr[4 * c + 1] = v.y; xr[4 * c + 2] = v.z; xr[4 * c + 3] = v.w;
        }
        if (QC) qc4 = valid ? srow[16] : make_float4(0.f, 0.f, 0.f, 0.f);
        // prefetch the next tile's row into the same slot (this thread is its only reader)
        if (it + 1 < n_my) {
          int nb = b, nw = w, npg = pg + 1;
          if (npg == tm.nPG) { npg = 0; if (++nw == tm.nW) { nw = 0; ++nb; } }
          const int npair = npg * WS_G + g, nsite = nw * WS_S + s;
          if (npair < Pl && nsite < L) prefetch_row(((size_t)nb * Pl + npair) * L + nsite);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      if (b != cur_b || w != cur_w) {  // new site window: reload M_l (producer warps only)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int n_sites = min(WS_S, L - w * WS_S);
        const float4* src = reinterpret_cast<const float4*>(colM + ((size_t)b * L + (size_t)w * WS_S) * PF_MROW);
        float4* dst = reinterpret_cast<float4*>(mwin);
        for (int i = ptid; i < n_sites * (PF_MROW / 4); i += 128) dst[i] = src[i];
        asm volatile("bar.sync 1, 128;" ::: "memory");
        cur_b = b; cur_w = w;
      }
      const float* mrow = mwin + (valid ? s : 0) * PF_MROW;
      const long long tp0 = TIC();
      // ---- column attention: q from LN_col(x1); the centred row is consumed on the fly ----
      float mean, rstd;
      float qh[PF_H];
      if (QC) {
        const float4 qi = *reinterpret_cast<const float4*>(mrow + 256);
        qh[0] = qc4.x * qi.x; qh[1] = qc4.y * qi.y; qh[2] = qc4.z * qi.z; qh[3] = qc4.w * qi.w;
        if (PROF) TOC(2, tp0);
      } else {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int c = 0; c < PF_D; c += 4) { s0 += xr[c]; s1 += xr[c + 1]; s2 += xr[c + 2]; s3 += xr[c + 3]; }
        mean = ((s0 + s1) + (s2 + s3)) * (1.0f / PF_D);
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
        for (int c = 0; c < PF_D; c += 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float d = xr[c + k] - mean;
            d0 = fmaf(kc.wq[c + k][0], d, d0); d1 = fmaf(kc.wq[c + k][1], d, d1);
            d2 = fmaf(kc.wq[c + k][2], d, d2); d3 = fmaf(kc.wq[c + k][3], d, d3);
            if (k == 0) q0 = fmaf(d, d, q0); else if (k == 1) q1 = fmaf(d, d, q1);
            else if (k == 2) q2 = fmaf(d, d, q2); else q3 = fmaf(d, d, q3);
          }
        }
        rstd = 1.0f / sqrtf(fmaf((q0 + q1) + (q2 + q3), 1.0f / PF_D, 1e-5f));
        if (PROF) { if (rstd > -1.f) TOC(2, tp0); }
        const float4 qi = *reinterpret_cast<const float4*>(mrow + 256);
        qh[0] = phi_elu1(fmaf(rstd, d0, kc.bq[0])) * qi.x;
        qh[1] = phi_elu1(fmaf(rstd, d1, kc.bq[1])) * qi.y;
        qh[2] = phi_elu1(fmaf(rstd, d2, kc.bq[2])) * qi.z;
        qh[3] = phi_elu1(fmaf(rstd, d3, kc.bq[3])) * qi.w;
      }
#pragma unroll
      for (int c = 0; c < PF_D; ++c) {
        const float4 m = *reinterpret_cast<const float4*>(mrow + 4 * c);
        float acc = kc.bo[c];
        acc = fmaf(m.x, qh[0], acc); acc = fmaf(m.y, qh[1], acc); acc = fmaf(m.z, qh[2], acc); acc = fmaf(m.w, qh[3], acc);
        xr[c] += acc;   // x2
      }
      // ---- wait for the slot, then seed the GEMM2 accumulator with x2 + b2 ----
      const int d = it % 3;                        // GEMM2 accumulator slot (3-deep ring)
      WAIT(0, BAR(2), par ^ 1);       // A1 free (both GEMM1 halves of the previous tile done)
      WAIT(1, BAR(4 + d), (uint32_t)(((it / 3) & 1) ^ 1));   // D2[d] free (E2 of tile it-3 done)
      tc_fence_after();
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(xr[16 * c4 + i] + kc.b2[16 * c4 + i]);
        tmem_st16(tmem + lane_base + WS_COL_D2 + 64 * d + 16 * c4, v);
      }
      // ---- LN_ffn (affine folded into W1/b1), bf16 hi/lo split -> A1[a] ----
      {   // packed pairs: half the issue slots for the two statistics passes (measured -3 % kernel time;
          // packing the q dots and the M q apply as well made the kernel 8 % slower)
        u64 sa = pk2(0.f, 0.f), sb = pk2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < PF_D; c += 4) { sa = add2(sa, pk2(xr[c], xr[c + 1])); sb = add2(sb, pk2(xr[c + 2], xr[c + 3])); }
        mean = hsum2(add2(sa, sb)) * (1.0f / PF_D);
        const u64 nm2 = pk2(-mean, -mean);
        u64 qa = pk2(0.f, 0.f), qb = pk2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < PF_D; c += 4) {
          const u64 da = add2(pk2(xr[c], xr[c + 1]), nm2), db = add2(pk2(xr[c + 2], xr[c + 3]), nm2);
          qa = fma2(da, da, qa); qb = fma2(db, db, qb);
        }
        rstd = 1.0f / sqrtf(fmaf(hsum2(add2(qa, qb)), 1.0f / PF_D, 1e-5f));
      }
      {
        unsigned char* a1hi = sm + WS_OFF_A1;
        unsigned char* a1lo = a1hi + 16384;
        const uint32_t rowoff = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
        const u64 nm = pk2(-mean, -mean), rs = pk2(rstd, rstd);
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {  // 16-byte chunk = 8 channels
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            u64 nv = mul2(add2(pk2(xr[8 * ch + 2 * i], xr[8 * ch + 2 * i + 1]), nm), rs);
            if (WS_FOLD63 && FMT == WS_FMT_BF16X3 && ch == 7 && i == 3) {   // operand column 63 carries the constant 1 (see WS_FOLD63)
              float n62, n63;
              up2(nv, n62, n63);
              nv = pk2(n62, 1.0f);
            }
            cvt2<FMT>(nv, hi[i], lo[i]);
          }
          const uint32_t off = rowoff + (uint32_t)(((ch ^ r) & 7) << 4);
          *reinterpret_cast<uint4*>(a1hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (FMT == WS_FMT_BF16X3) *reinterpret_cast<uint4*>(a1lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      { const long long t0 = TIC(); tc_wait_st(); TOC(3, t0); }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(BAR(0));
    }
  } else
  if (warp >= WS_MW) {
#if WS_EW == 16
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");   // whole MMA warpgroup (issuer + 3 idle warps)
#elif WS_EW == 12
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // 640 x 96 = 12x32x96 + 4x32x128 + 4x32x40 + slack
#endif
    if (warp == WS_MW) {
    // =============================== MMA ISSUER =============================================
    // The whole warp walks the loop (so descriptors live in uniform registers); one elected lane
    // issues.  Descriptors are built once: per k-step only the 16-byte-unit address field moves.
    if (n_my > 0) {
      const uint32_t idesc1 = umma_idesc(128, 128, FMT == WS_FMT_F16), idesc2 = umma_idesc(128, 64, FMT == WS_FMT_F16);
      const u64 dA00 = umma_desc(sbase + WS_OFF_A1), dA01 = umma_desc(sbase + WS_OFF_A1 + 16384);
      const u64 dW1h = umma_desc(sbase + TC_OFF_W1HI), dW1l = umma_desc(sbase + TC_OFF_W1LO);
      const u64 dW2h = umma_desc(sbase + TC_OFF_W2HI), dW2l = umma_desc(sbase + TC_OFF_W2LO);
      auto issue_g1 = [&](int a, int half) {  // D1[half] = A1[a] . W1[half*128 .. +128)^T
        const u64 ah = dA00, al = dA01;
        (void)a;
        const u64 hoff = (u64)(half * (16384 >> 4));
        const uint32_t dcol = tmem + 128 * half;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          if (t < n_terms) {
            const u64 da = (t == 2) ? al : ah;
            const u64 db = ((t == 1) ? dW1l : dW1h) + hoff;
#pragma unroll
            for (int s = 0; s < 4; ++s) umma_ss(dcol, da + 2 * s, db + 2 * s, idesc1, (t | s) ? 1u : 0u);
          }
        }
      };
      auto issue_g2 = [&](int d, int half) {  // D2[slot d] += H[half] . W2[:, half*128 .. +128)^T
        const uint32_t dcol = tmem + WS_COL_D2 + 64 * d;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          if (t < n_terms) {
            const u64 db = (t == 1) ? dW2l : dW2h;        // same term order as GEMM1: hi.hi, hi.lo(W), lo(A).hi
            const uint32_t a_sel = (t == 2) ? 8u : 0u;
#pragma unroll
            for (int s8 = 0; s8 < 8; ++s8) {
              const int s = half * 8 + s8;
              umma_ts(dcol, tmem + (uint32_t)(16 * s) + a_sel, db + (u64)((s >> 2) * 512 + (s & 3) * 2), idesc2, 1u);
            }
          }
        }
      };
      // Rolled step loop (one code instance of each issue function).  Order on the tensor pipe:
      //   G1a(0) G1b(0) | G2a(i) G1a(i+1) G2b(i) G1b(i+1) ...   (in-order issue covers the WAR on D1)
#pragma unroll 1
      for (int step = -1; step < 2 * n_my; ++step) {
        const int it = step >> 1, half = step & 1;      // step -1: preamble for tile 0 (it = -1, half = 1)
        const int d = (it < 0 ? 0 : it) % 3;
        const bool has_next = it + 1 < n_my;
        const long long t0 = TIC();
        if (step >= 0) {
          WAIT(1 + half, BAR(9 + half), (uint32_t)(it & 1));   // H half ready
          tc_fence_after();
          if (elect_one()) {
            issue_g2(d, half);
            if (half == 1) tc_commit(BAR(11 + d));
          }
          __syncwarp();
        }
        if ((half == 0 && has_next) || step == -1) {   // the next tile's A operand must have landed
          WAIT(0, BAR(0), (uint32_t)((it + 1) & 1));
          tc_fence_after();
        }
        if (has_next) {
          if (elect_one()) {
            if (step == -1) { issue_g1(0, 0); tc_commit(BAR(7)); }
            issue_g1(0, half);
            tc_commit(BAR(7 + half));
            if (half == 1) tc_commit(BAR(2));
          }
        }
        __syncwarp();
        TOC(3, t0);
      }
    }
    }
  } else if (warp >= WS_EW0 && warp < WS_EW0 + WS_EW) {
    // =============================== EPILOGUE ===============================================
    const int q = warp & 3, chf = (warp - WS_EW0) >> 2;   // TMEM lane quadrant, column group
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int r = q * 32 + lane, g = r >> WS_S_LOG2, s = r & (WS_S - 1);
    auto e1 = [&](int half, bool dump_this) {  // this warp's 16-column chunks of D1[half] -> gelu -> bf16 hi/lo in place
      // chunk j of the half (16 hidden units) belongs to column group j % WS_NCG
      constexpr int NCH = (8 - 1) / WS_NCG + 1;            // max chunks per warp and half
      uint32_t v[2][16];
      auto col_of = [&](int i) { return half * 128 + (chf + i * WS_NCG) * 16; };   // TMEM column == hidden unit
      auto chunk = [&](const uint32_t(&vc)[16], int cc) {
        if (PROF && dump_this) {   // debug builds only: predicated-off stores still cost issue slots
#pragma unroll
          for (int i = 0; i < 16; ++i) dump[r * 320 + cc + i] = __uint_as_float(vc[i]);
        }
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float h0, h1;   // one packed add for the pair's bias
#if defined(WS_DIAG_NO_B1)
          up2(pk2(__uint_as_float(vc[2 * i]), __uint_as_float(vc[2 * i + 1])), h0, h1);
#elif WS_B1_CONST
          if (WS_FOLD63 && FMT == WS_FMT_BF16X3)   // the accumulator already holds W1 LN(x) + b1
            up2(pk2(__uint_as_float(vc[2 * i]), __uint_as_float(vc[2 * i + 1])), h0, h1);
          else
            up2(add2(pk2(__uint_as_float(vc[2 * i]), __uint_as_float(vc[2 * i + 1])),
                     pk2(kc.b1[cc + 2 * i], kc.b1[cc + 2 * i + 1])), h0, h1);
#else
          up2(add2(pk2(__uint_as_float(vc[2 * i]), __uint_as_float(vc[2 * i + 1])),
                   *reinterpret_cast<const u64*>(sb1 + cc + 2 * i)), h0, h1);
#endif
          cvt2<FMT>(gelu_fast2(h0, h1), hi[i], lo[i]);
        }
#ifdef WS_DIAG_NO_STTM
        if (hi[0] == 0x12345678u && lo[3] == 0x9abcdef0u)   // timing diagnostic only: (practically) never stores
#endif
        if (WS_ST16 && FMT == WS_FMT_BF16X3) {   // hi words in columns cc..cc+7, lo words in cc+8..cc+15: one 16-column store
          uint32_t hl[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) { hl[i] = hi[i]; hl[8 + i] = lo[i]; }
          tmem_st16(tmem + lane_base + cc, hl);
        } else {
          tmem_st8(tmem + lane_base + cc, hi);
          if (FMT == WS_FMT_BF16X3) tmem_st8(tmem + lane_base + cc + 8, lo);
        }
      };
      const int n_mine = (8 - chf + WS_NCG - 1) / WS_NCG;   // chunks of this warp in a half
      E1_LD(tmem + lane_base + col_of(0), v[0]);
#pragma unroll 1
      for (int i = 0; i < NCH; i += 2) {  // ping-pong (rolled: keeps the epilogue inside the instruction cache)
        if (i < n_mine) {
          const long long t0 = TIC();
          tc_wait_ld();
          TOC(3, t0);
          if (i + 1 < n_mine) E1_LD(tmem + lane_base + col_of(i + 1), v[1]);
          chunk(v[0], col_of(i));
        }
        if (i + 1 < n_mine) {
          const long long t0 = TIC();
          tc_wait_ld();
          TOC(3, t0);
          if (i + 2 < n_mine) E1_LD(tmem + lane_base + col_of(i + 2), v[0]);
          chunk(v[1], col_of(i + 1));
        }
      }
      const long long t1 = TIC();
      tc_wait_st();
      TOC(4, t1);
      tc_fence_before();
    };
    bool first_e2 = true;
    int eb, ew, epg;                       // tile coordinates of the next tile e2 will store
    tm.decode((unsigned)t_begin, eb, ew, epg);
    int slot3 = 0;                         // it % 3 of that tile
    uint32_t par3 = 0;                     // (it / 3) & 1
    auto e2 = [&]() {  // D2[slot3] cols [WS_E2COLS chf, +WS_E2COLS) -> HBM, tiles in order
      const int a = slot3;
      WAIT(2, BAR(11 + a), par3);
      tc_fence_after();
      const int b = eb, w = ew, pg = epg;
      if (++epg == tm.nPG) { epg = 0; if (++ew == tm.nW) { ew = 0; ++eb; } }
      if (++slot3 == 3) { slot3 = 0; par3 ^= 1; }
      const int pair = pg * WS_G + g, site = w * WS_S + s;
      const bool valid = (pair < Pl) && (site < L);
      float* dst = x + (((size_t)b * Pl + (valid ? pair : 0)) * L + (valid ? site : 0)) * PF_D;
      if (HEAD) {   // distance head on the final row (column group 0 takes the whole row); nothing is stored to x
        static_assert(!HEAD || WS_S == 4, "the fused head sums a pair's sites with two lane shuffles");
        if (chf == 0) {
          float dot = kc.bhead;
#pragma unroll
          for (int jc = 0; jc < 4; ++jc) {
            uint32_t v[16];
            tmem_ld16(tmem + lane_base + WS_COL_D2 + 64 * a + 16 * jc, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i)
            dot = fmaf(__uint_as_float(v[i]), kc.whead[16 * jc + i], dot);
          }
          float sp = valid ? softplus20(dot) : 0.f;
          sp += __shfl_xor_sync(PF_FULL, sp, 1);      // the tile's WS_S = 4 sites of a pair sit in adjacent lanes
          sp += __shfl_xor_sync(PF_FULL, sp, 2);
          if ((lane & 3) == 0 && pair < Pl) dump[((size_t)b * tm.nW + w) * Pl + pair] = sp;
        }
      }
      // D2 has four 16-column chunks; chunk j goes to column group (WS_NCG-1-j) mod WS_NCG, which
      // gives the extra store chunk to the group with the fewest GELU chunks
#pragma unroll
      for (int jc = 0; jc < (HEAD ? 0 : 4); ++jc) {
        if ((WS_NCG - 1 - jc % WS_NCG) == chf) {
          uint32_t v[16];
          tmem_ld16(tmem + lane_base + WS_COL_D2 + 64 * a + 16 * jc, v);
          tc_wait_ld();
          if (PROF && dump != nullptr && blockIdx.x == 0 && first_e2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dump[r * 320 + 256 + 16 * jc + i] = __uint_as_float(v[i]);
          }
          if (valid) {
            stg256(dst + 16 * jc, v);
            stg256(dst + 16 * jc + 8, v + 8);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(BAR(4 + a));
      first_e2 = false;
    };
    // One code instance of e1 and e2 (rolled loops): the three roles run different code at the
    // same time and together must stay close to the instruction cache (measured: unrolling e1
    // 4x costs 17 % of the kernel).  Sequence per tile: E1a(it)  E2(it-1)  E1b(it).
#pragma unroll 1
    for (int step = 0; step < 2 * n_my + 1; ++step) {
      const int it = step >> 1, half = step & 1;
      if (step < 2 * n_my) {
        const uint32_t ph = (uint32_t)(it & 1);
        const bool dump_this = (dump != nullptr) && blockIdx.x == 0 && it == 0;
        WAIT(half, BAR(7 + half), ph);
        tc_fence_after();
        e1(half, dump_this);
        mbar_arrive(BAR(9 + half));
      }
      if (half == 0 && step > 0) {   // after E1a(it): store tile it-1 (also the final tile at step == 2 n_my)
        const long long t0 = TIC();
        e2();
        TOC(5, t0);
      }
    }
  }
  if (PROF && dump != nullptr && (tid == WS_PW0 * 32 || tid == WS_MW * 32 || tid == WS_EW0 * 32)) {
    // dump[cta][role 0..2][0..3]: total cycles, wait slot 0, 1, 2   (role 0 producer, 1 mma, 2 epilogue)
    float* o = dump + 128 * 320 + (blockIdx.x * 3 + (tid == WS_PW0 * 32 ? 0 : tid == WS_MW * 32 ? 1 : 2)) * 8;
    o[0] = (float)(clock64() - t_start);
    for (int k = 0; k < 7; ++k) o[1 + k] = (float)tw[k];
  }
  if (!ok && err_flag != nullptr) *err_flag = 2;
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == WS_MW) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

template <bool QC>
inline int pf_ffn_ws_init_qc() {
  int rc = (int)cudaFuncSetAttribute(k_colapply_ffn_ws<false, WS_FMT_BF16X3, QC>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES);
  if (rc == 0) rc = (int)cudaFuncSetAttribute(k_colapply_ffn_ws<true, WS_FMT_BF16X3, QC>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES);
  if (rc == 0) rc = (int)cudaFuncSetAttribute(k_colapply_ffn_ws<false, WS_FMT_BF16, QC>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES);
  if (rc == 0) rc = (int)cudaFuncSetAttribute(k_colapply_ffn_ws<false, WS_FMT_F16, QC>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES);
  if (rc == 0) rc = (int)cudaFuncSetAttribute(k_colapply_ffn_ws<false, WS_FMT_BF16X3, QC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES);
  if (rc == 0) rc = (int)cudaFuncSetAttribute(k_colapply_ffn_ws<false, WS_FMT_BF16, QC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES);
  if (rc == 0) rc = (int)cudaFuncSetAttribute(k_colapply_ffn_ws<false, WS_FMT_F16, QC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES);
  return rc;
}
inline int pf_ffn_ws_init() {
  const int rc = pf_ffn_ws_init_qc<false>();
  return rc ? rc : pf_ffn_ws_init_qc<true>();
}

// fmt: WS_FMT_*; n_terms: MMA passes per product (3 for BF16X3, 1 for BF16, 2 for F16 = hi and lo weights).
// Wt must hold the weight images of the matching 16-bit format.  qcache: the per-token q~ cache of
// k_col_partial_tc, or nullptr (the producer then recomputes q~ itself).
template <bool QC>
inline int pf_ffn_ws_launch_head(const PfFfnConst& kc, const PfFfnTcW* Wt, float* x, const float* colM, const float* qcache,
                                 int L, int Pl, int B, int grid, int fmt, int n_terms, int* err_flag, float* headpart, cudaStream_t st) {
  if (fmt == WS_FMT_F16)
    k_colapply_ffn_ws<false, WS_FMT_F16, QC, true><<<grid, WS_THREADS, WS_SMEM_BYTES, st>>>(kc, Wt, x, colM, qcache, L, Pl, B, n_terms, err_flag, headpart);
  else if (fmt == WS_FMT_BF16)
    k_colapply_ffn_ws<false, WS_FMT_BF16, QC, true><<<grid, WS_THREADS, WS_SMEM_BYTES, st>>>(kc, Wt, x, colM, qcache, L, Pl, B, n_terms, err_flag, headpart);
  else
    k_colapply_ffn_ws<false, WS_FMT_BF16X3, QC, true><<<grid, WS_THREADS, WS_SMEM_BYTES, st>>>(kc, Wt, x, colM, qcache, L, Pl, B, n_terms, err_flag, headpart);
  return (int)cudaGetLastError();
}
template <bool QC>
inline int pf_ffn_ws_launch_qc(const PfFfnConst& kc, const PfFfnTcW* Wt, float* x, const float* colM, const float* qcache,
                               int L, int Pl, int B, int grid, int fmt, int n_terms, int* err_flag, float* dump, int prof,
                               cudaStream_t st) {
  if (fmt == WS_FMT_F16)
    k_colapply_ffn_ws<false, WS_FMT_F16, QC><<<grid, WS_THREADS, WS_SMEM_BYTES, st>>>(kc, Wt, x, colM, qcache, L, Pl, B, n_terms, err_flag, nullptr);
  else if (fmt == WS_FMT_BF16)
    k_colapply_ffn_ws<false, WS_FMT_BF16, QC><<<grid, WS_THREADS, WS_SMEM_BYTES, st>>>(kc, Wt, x, colM, qcache, L, Pl, B, n_terms, err_flag, nullptr);
  else if (prof || dump != nullptr)   // the debug/profiling instantiation carries the dump and the role timers
    k_colapply_ffn_ws<true, WS_FMT_BF16X3, QC><<<grid, WS_THREADS, WS_SMEM_BYTES, st>>>(kc, Wt, x, colM, qcache, L, Pl, B, n_terms, err_flag, dump);
  else
    k_colapply_ffn_ws<false, WS_FMT_BF16X3, QC><<<grid, WS_THREADS, WS_SMEM_BYTES, st>>>(kc, Wt, x, colM, qcache, L, Pl, B, n_terms, err_flag, dump);
  return (int)cudaGetLastError();
}
inline int pf_ffn_ws_launch(const PfFfnConst& kc, const PfFfnTcW* Wt, float* x, const float* colM, const float* qcache, int L,
                            int Pl, int B, int n_sm, int fmt, int n_terms, int* err_flag, float* dump, int prof, cudaStream_t st,
                            float* headpart = nullptr) {
  const long long nt = (long long)B * ((L + WS_S - 1) / WS_S) * ((Pl + WS_G - 1) / WS_G);
  if (nt > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
  if (fmt != WS_FMT_BF16X3 && n_terms > 2) return (int)cudaErrorInvalidValue;   // no lo activations in these formats
  const int grid = (int)(nt < n_sm ? nt : n_sm);
  if (headpart != nullptr)
    return qcache != nullptr ? pf_ffn_ws_launch_head<true>(kc, Wt, x, colM, qcache, L, Pl, B, grid, fmt, n_terms, err_flag, headpart, st)
                             : pf_ffn_ws_launch_head<false>(kc, Wt, x, colM, qcache, L, Pl, B, grid, fmt, n_terms, err_flag, headpart, st);
  return qcache != nullptr
             ? pf_ffn_ws_launch_qc<true>(kc, Wt, x, colM, qcache, L, Pl, B, grid, fmt, n_terms, err_flag, dump, prof, st)
             : pf_ffn_ws_launch_qc<false>(kc, Wt, x, colM, qcache, L, Pl, B, grid, fmt, n_terms, err_flag, dump, prof, st);
}
