// pf_attn_tc.cuh -- the attention contractions on the 5th-gen tensor cores (tcgen05 / TMEM).
//
// Per token and attention module the reference evaluates (attention.py:163-190, collapsed as in
// SURVEY.md section 3.2, u = LayerNorm(x) without affine):
//     q, k  = Wqk u            8 dot products of length 64                 "QK"  [tokens x 64] . [64 x 8]
//     S    += k~ (x) u         4 x 64 outer-product accumulation           "S"   [64 x tokens] . [tokens x 4]
// Both are GEMMs over a 128-token tile.  The tile's LayerNorm output is split ONCE into bf16 hi/lo
// parts and stored as ONE shared-memory image [token][64 channels] (128-byte rows, SWIZZLE_128B); that
// image is read twice by the tensor core:
//   QK  as the K-major A operand   (M = 128 tokens, K = 64 channels), B = folded q/k weights [16 x 64]
//       3 passes hi.hi + hi.lo(W) + lo.hi  ->  D_qk[128 x 16] in TMEM (columns 0..3 k, 4..7 q)
//   S   as the MN-major A operand  (M = 128 = [64 channels of the hi image | 64 channels of the lo
//       image], K = tokens), B = k~ as [16 x tokens] K-major (rows 0..3 k~ hi, 4..7 k~ lo)
//       ->  D_S[128 x 16]:  S[h][c] = D[c][h] + D[c][4+h] + D[64+c][h] + D[64+c][4+h]
//       (all four hi/lo cross terms: the full fp32-split product, accumulated in fp32 over the tokens)
// tools/probe/umma_mn_probe.cu pins both operand forms bit-exactly on integer data.
// What is left on the CUDA cores per token: LayerNorm statistics, the bf16 split (shared by both
// GEMMs), phi on 8 values.  The fp32 ("exact") precision mode keeps the FFMA kernels of pf_kernels.cuh.
#pragma once
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "pf_ffn_ws.cuh"

struct PfAttnTcW {             // one attention module: B operand of the QK GEMM
  uint16_t wqk_hi[16 * PF_D];  // K-major SWIZZLE_128B image [16][64]: rows 0..3 k heads, 4..7 q heads, 8..15 zero
  uint16_t wqk_lo[16 * PF_D];
  float bqk[8];
  float pad[8];
};

inline void pf_pack_attn_tc(const PfAttnW& a, PfAttnTcW* o) {
  memset(o, 0, sizeof(*o));
  for (int n = 0; n < 8; ++n) {
    for (int k = 0; k < PF_D; ++k) {
      const float w = a.wqk[n][k];
      const uint16_t hi = f32_to_bf16_rn(w);
      const uint16_t lo = f32_to_bf16_rn(w - bf16_to_f32(hi));
      const uint32_t off = umma_off_k64(n, k) / 2;
      o->wqk_hi[off] = hi;
      o->wqk_lo[off] = lo;
    }
    o->bqk[n] = a.bqk[n];
  }
}

// MN-major SWIZZLE_128B shared-memory descriptor: LBO = byte distance between 64-element blocks along
// M (here: the hi and the lo image), SBO = byte distance between 8-row groups along K (1024).
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
#define UMMA_IDESC_A_MN (1u << 15)

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// ---- shared pieces of a 128-token attention tile (one thread per token row r) ----------------
#define AT_XROW 272          // staged fp32 row: 256 B + 16 B pad (conflict-free LDS.128 per thread)
#define AT_A1_BYTES 32768    // hi 16 KB | lo 16 KB
#define AT_KT_BYTES 4096     // [16][128] bf16, two 64-token K atoms of 2 KB

// LayerNorm (no affine) of the row held in xr, bf16 hi/lo split, store as row r of the A image.
__device__ __forceinline__ void at_ln_split_store(const float (&xr)[PF_D], unsigned char* a1, int r) {
  u64 sa = pk2(0.f, 0.f), sb = pk2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < PF_D; c += 4) { sa = add2(sa, pk2(xr[c], xr[c + 1])); sb = add2(sb, pk2(xr[c + 2], xr[c + 3])); }
  const float mean = hsum2(add2(sa, sb)) * (1.0f / PF_D);
  const u64 nm = pk2(-mean, -mean);
  u64 qa = pk2(0.f, 0.f), qb = pk2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < PF_D; c += 4) {
    const u64 da = add2(pk2(xr[c], xr[c + 1]), nm), db = add2(pk2(xr[c + 2], xr[c + 3]), nm);
    qa = fma2(da, da, qa); qb = fma2(db, db, qb);
  }
  const float rstd = 1.0f / sqrtf(fmaf(hsum2(add2(qa, qb)), 1.0f / PF_D, 1e-5f));
  const u64 rs = pk2(rstd, rstd);
  unsigned char* a1lo = a1 + 16384;
  const uint32_t rowoff = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {  // 16-byte chunk = 8 channels
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const u64 nv = mul2(add2(pk2(xr[8 * ch + 2 * i], xr[8 * ch + 2 * i + 1]), nm), rs);
      split2(nv, hi[i], lo[i]);
    }
    const uint32_t off = rowoff + (uint32_t)(((ch ^ r) & 7) << 4);
    *reinterpret_cast<uint4*>(a1 + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(a1lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// One elected thread: D_qk[128 x 16] = A1 . Wqk^T, three passes (hi.hi, hi.lo(W), lo.hi).
__device__ __forceinline__ void at_issue_qk(uint32_t a1_u32, uint32_t bq_u32, uint32_t d_tmem) {
  const uint32_t idesc = umma_idesc(128, 16);
  const u64 ah = umma_desc(a1_u32), al = umma_desc(a1_u32 + 16384);
  const u64 bh = umma_desc(bq_u32), bl = umma_desc(bq_u32 + 2048);
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const u64 da = (t == 2) ? al : ah, db = (t == 1) ? bl : bh;
#pragma unroll
    for (int s = 0; s < 4; ++s) umma_ss(d_tmem, da + 2 * s, db + 2 * s, idesc, (t | s) ? 1u : 0u);
  }
}
// One elected thread: D_S[128 x 16] (+)= A1^T[rows row0 .. row0 + 16 nk) . KT, contraction over tokens.
__device__ __forceinline__ void at_issue_s(uint32_t a1_u32, uint32_t kt_u32, uint32_t d_tmem, int row0, int nk, bool accumulate) {
  const uint32_t idesc = umma_idesc(128, 16) | UMMA_IDESC_A_MN;
  for (int k = 0; k < nk; ++k) {
    const int row = row0 + 16 * k;
    const u64 da = umma_desc_mn(a1_u32 + (uint32_t)row * 128u, 16384u, 1024u);
    const u64 db = umma_desc(kt_u32 + (uint32_t)((row >> 6) * 2048 + (row & 63) * 2));
    umma_ss(d_tmem, da, db, idesc, (accumulate || k > 0) ? 1u : 0u);
  }
}
// Token r's k~ values as bf16 hi/lo into column r of the K-major KT operand (rows 0..3 hi, 4..7 lo).
__device__ __forceinline__ void at_store_kt(unsigned char* kt, int r, const float (&kq)[8]) {
  const int kk = r & 63;
  unsigned char* base = kt + (r >> 6) * 2048 + (kk & 7) * 2;
#pragma unroll
  for (int h = 0; h < PF_H; ++h) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(kq[h]);
    const __nv_bfloat16 lo = __float2bfloat16_rn(kq[h] - __bfloat162float(hi));
    *reinterpret_cast<__nv_bfloat16*>(base + h * 128 + ((((kk >> 3) ^ h) & 7) << 4)) = hi;
    *reinterpret_cast<__nv_bfloat16*>(base + (4 + h) * 128 + ((((kk >> 3) ^ (4 + h)) & 7) << 4)) = lo;
  }
}

// ------------------------------------------------------------------------------------------
// Column attention, step 1 on the tensor cores: partial sums over a chunk of pairs at each site
// (same output as k_col_partial: part[chunk][b][l][264]) plus the per-token q~ cache that the FFN
// kernel's column apply reads (16 B per token) instead of recomputing LN_col and the q dots.
// A tile is 32 pairs x 4 consecutive sites (1 KB contiguous per pair): warp s owns site s, lane =
// pair, so a site's 32 tokens are 32 consecutive rows of the operand image (two K = 16 steps of the
// S GEMM) and the per-site sums of k~, q~ are plain warp reductions.  A work unit is (msa, 4-site
// window, pair chunk); 128-thread CTAs, two per SM, walk over units; the A image is double buffered so
// the S GEMM of a tile overlaps the next tile's LayerNorm/split.
// ------------------------------------------------------------------------------------------
#define CT_THREADS 128
#define CT_OFF_A1 0
#define CT_OFF_XST (2 * AT_A1_BYTES)
#define CT_OFF_KT (CT_OFF_XST + 128 * AT_XROW)
#define CT_OFF_BQ (CT_OFF_KT + 2 * AT_KT_BYTES)
#define CT_OFF_BAR (CT_OFF_BQ + 4096)
#define CT_OFF_TMEM (CT_OFF_BAR + 32)
#define CT_SMEM_BYTES (CT_OFF_TMEM + 32 + 1024)
#define CT_TM_COLS 128
#define CT_TM_S 32           // D_S of site s at TMEM column 32 + 16 s; D_qk at column 0

__global__ void __launch_bounds__(CT_THREADS, 2)
k_col_partial_tc(const PfAttnTcW* __restrict__ Wt, const float* __restrict__ x, float* __restrict__ part,
                 float* __restrict__ qcache, int B, int L, int Pl, int ppc, int n_chunks, int* __restrict__ err_flag) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar_qk = sbase + CT_OFF_BAR, bar_s0 = sbase + CT_OFF_BAR + 8;

  // ---- one-time setup: weights, zeroed KT (rows 8..15 stay zero), barriers, TMEM ----
  for (int i = tid; i < 4096 / 16; i += CT_THREADS)
    reinterpret_cast<int4*>(sm + CT_OFF_BQ)[i] = reinterpret_cast<const int4*>(Wt->wqk_hi)[i];
  for (int i = tid; i < 2 * AT_KT_BYTES / 16; i += CT_THREADS) reinterpret_cast<int4*>(sm + CT_OFF_KT)[i] = make_int4(0, 0, 0, 0);
  float bq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bq[i] = Wt->bqk[i];
  if (tid == 0) {
    mbar_init(bar_qk, 1); mbar_init(bar_s0, 1); mbar_init(bar_s0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + CT_OFF_TMEM), "r"(CT_TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + CT_OFF_TMEM);
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  const int nW = (L + 3) >> 2;
  const long long upc = (long long)B * nW;            // units per chunk
  const long long n_units = upc * n_chunks;
  const uint32_t xst_u32 = sbase + CT_OFF_XST + (uint32_t)tid * AT_XROW;
  const float4* srow = reinterpret_cast<const float4*>(sm + CT_OFF_XST + tid * AT_XROW);
  bool ok = true;

  // token of (unit coords, tile t) that this thread owns; returns nullptr when out of range
  auto tok_ptr = [&](int chunk, int b, int w, int t) -> const float* {
    const int pair = chunk * ppc + 32 * t + lane, site = 4 * w + warp;
    const int p1 = min(Pl, (chunk + 1) * ppc);
    if (pair >= p1 || site >= L) return nullptr;
    return x + (((size_t)b * Pl + pair) * L + site) * PF_D;
  };
  auto prefetch = [&](const float* src) {
    if (src != nullptr) {
#pragma unroll
      for (int c = 0; c < 16; ++c)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(xst_u32 + 16 * c), "l"(src + 4 * c) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto decode = [&](long long u, int& chunk, int& b, int& w) {
    chunk = (int)(u / upc);
    const int rem = (int)(u - (long long)chunk * upc);
    b = rem / nW;
    w = rem - b * nW;
  };

  int n_glob = 0;   // tiles processed by this CTA (barrier parities)
  if ((long long)blockIdx.x < n_units) {
    int c0, b0, w0;
    decode(blockIdx.x, c0, b0, w0);
    prefetch(tok_ptr(c0, b0, w0, 0));
  }
  for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
    int chunk, b, w, nchunk = 0, nb = 0, nw = 0;
    decode(u, chunk, b, w);
    const bool has_next = u + gridDim.x < n_units;
    if (has_next) decode(u + gridDim.x, nchunk, nb, nw);
    const int p0 = chunk * ppc, p1 = min(Pl, p0 + ppc);
    const int nt = (p1 - p0 + 31) >> 5;
    float ks[PF_H] = {0.f, 0.f, 0.f, 0.f}, qs[PF_H] = {0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < nt; ++t, ++n_glob) {
      const int buf = n_glob & 1;
      unsigned char* a1 = sm + CT_OFF_A1 + buf * AT_A1_BYTES;
      unsigned char* kt = sm + CT_OFF_KT + buf * AT_KT_BYTES;
      const float* mine = tok_ptr(chunk, b, w, t);
      const bool valid = mine != nullptr;
      // ---- phase 1: staged row -> registers, LayerNorm, split, operand image ----
      float xr[PF_D];
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) v = srow[c];
        xr[4 * c] = v.x; xr[4 * c + 1] = v.y; xr[4 * c + 2] = v.z; xr[4 * c + 3] = v.w;
      }
      prefetch(t + 1 < nt ? tok_ptr(chunk, b, w, t + 1) : (has_next ? tok_ptr(nchunk, nb, nw, 0) : nullptr));
      ok = mbar_wait(bar_s0 + 8 * buf, (uint32_t)(((n_glob >> 1) & 1) ^ 1)) && ok;   // S GEMM of tile n-2 has read this buffer
      at_ln_split_store(xr, a1, tid);
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        at_issue_qk(sbase + CT_OFF_A1 + buf * AT_A1_BYTES, sbase + CT_OFF_BQ, tmem);
        tc_commit(bar_qk);
      }
      // ---- phase 2: q, k from TMEM, phi, sums, q~ cache, k~ operand ----
      ok = mbar_wait(bar_qk, (uint32_t)(n_glob & 1)) && ok;
      tc_fence_after();
      uint32_t v[8];
      tmem_ld8(tmem + lane_base, v);
      tc_wait_ld();
      float kq[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) kq[i] = valid ? phi_elu1(__uint_as_float(v[i]) + bq[i]) : 0.f;
#pragma unroll
      for (int h = 0; h < PF_H; ++h) { ks[h] += kq[h]; qs[h] += kq[4 + h]; }
      if (valid) {
        const size_t tok = (size_t)(mine - x) / PF_D;
        *reinterpret_cast<float4*>(qcache + tok * 4) = make_float4(kq[4], kq[5], kq[6], kq[7]);
      }
      at_store_kt(kt, tid, kq);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < 4; ++s)
          at_issue_s(sbase + CT_OFF_A1 + buf * AT_A1_BYTES, sbase + CT_OFF_KT + buf * AT_KT_BYTES, tmem + CT_TM_S + 16 * s,
                     32 * s, 2, t > 0);
        tc_commit(bar_s0 + 8 * buf);
      }
    }
    // ---- unit end: D_S -> part ----
    {
      const int last = n_glob - 1;
      ok = mbar_wait(bar_s0 + 8 * (last & 1), (uint32_t)((last >> 1) & 1)) && ok;
      tc_fence_after();
      float* red = reinterpret_cast<float*>(sm + CT_OFF_A1);   // [4 sites][4 heads][64]: both images are idle here
      float val[4][PF_H];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        uint32_t d[8];
        tmem_ld8(tmem + lane_base + CT_TM_S + 16 * s, d);
        tc_wait_ld();
#pragma unroll
        for (int h = 0; h < PF_H; ++h) val[s][h] = __uint_as_float(d[h]) + __uint_as_float(d[4 + h]);
      }
      if (tid >= 64) {
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
          for (int h = 0; h < PF_H; ++h) red[(s * PF_H + h) * 64 + (tid - 64)] = val[s][h];
      }
      // per-site sums of k~ and q~: warp s owns site s
#pragma unroll
      for (int h = 0; h < PF_H; ++h) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ks[h] += __shfl_xor_sync(PF_FULL, ks[h], o);
          qs[h] += __shfl_xor_sync(PF_FULL, qs[h], o);
        }
      }
      tc_fence_before();
      __syncthreads();
      float* obase = part + (((size_t)chunk * B + b) * L + (size_t)4 * w) * PF_PART;
      if (tid < 64) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (4 * w + s < L) {
#pragma unroll
            for (int h = 0; h < PF_H; ++h)
              obase[(size_t)s * PF_PART + 8 + h * PF_D + tid] = val[s][h] + red[(s * PF_H + h) * 64 + tid];
          }
        }
      }
      if (lane == 0 && 4 * w + warp < L) {
        float* o = obase + (size_t)warp * PF_PART;
#pragma unroll
        for (int h = 0; h < PF_H; ++h) { o[h] = ks[h]; o[4 + h] = qs[h]; }
      }
      __syncthreads();   // red (aliasing the operand image) is read before the next unit overwrites it
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (!ok && err_flag != nullptr) *err_flag = 4;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(CT_TM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------
// k_col_partial_ws: the same computation, warp specialised and persistent (one CTA per SM):
//   warp 21        LOADER   per tile ONE tensor-map TMA copy (cp.async.bulk.tensor.3d, mbarrier completion) of the
//                           box [32 pairs][4 sites][64 channels] into a 3-stage shared-memory ring; sites past L
//                           and pairs past the tensor arrive as zeros (32 per-pair bulk copies serialise in the
//                           issuing warp, one uniform-register instruction per lane: measured loader-bound)
//   warps 0..15    P1       8 lanes per token (coalesced LDS.128, 3-step butterflies, ~50 registers):
//                           LayerNorm, bf16 hi/lo split, operand image (3-deep ring); warp w owns pairs 2w, 2w+1
//   warp 20        MMA      one elected lane issues QK(n+1) and S(n) (see the header of this file)
//   warps 16..19   P2       TMEM lane quadrant q = site q, lane = pair: q~, k~ from D_qk, phi, running sums,
//                           q~ cache, k~ operand; at the end of a work unit D_S -> part
// (k_col_partial_tc above, one thread per token and everything in program order, is kept as the
//  readable restatement and A/B partner: PF_COL_IMPL=tc1.)
// ------------------------------------------------------------------------------------------
#define C2_NP1 16
#define C2_THREADS (22 * 32)
#define C2_NS 3                                   // ring depth (staging, operand image, k~ operand, D_qk)
#define C2_OFF_A1 0
#define C2_OFF_XS (C2_NS * AT_A1_BYTES)           // [C2_NS][32 pairs][4 sites][64] fp32
#define C2_OFF_KT (C2_OFF_XS + C2_NS * 32768)
#define C2_OFF_BQ (C2_OFF_KT + C2_NS * AT_KT_BYTES)
#define C2_OFF_RED (C2_OFF_BQ + 4096)             // [4 sites][4 heads][64] floats
#define C2_OFF_BAR (C2_OFF_RED + 4096)
#define C2_OFF_TMEM (C2_OFF_BAR + 256)
#define C2_SMEM_BYTES (C2_OFF_TMEM + 32 + 1024)
#define C2_TM_COLS 256
#define C2_TM_S 128                               // D_S[parity][site] at column 128 + 64 parity + 16 site; D_qk[i] at 16 i
// barrier indices
#define C2_B_XFULL 0
#define C2_B_XFREE 3
#define C2_B_A1FULL 6
#define C2_B_A1FREE 9
#define C2_B_QKDONE 12
#define C2_B_KTFULL 15
#define C2_B_DSFULL 18
#define C2_B_DSFREE 20

struct C2Iter {   // the CTA's tile stream: units (chunk, msa, window) strided over the grid, tiles within a unit
  int B, L, Pl, ppc, nW;
  long long upc, n_units, u;
  int chunk, b, w, t, nt, np_last;
  __device__ __forceinline__ void load_unit() {
    chunk = (int)(u / upc);
    const int rem = (int)(u - (long long)chunk * upc);
    b = rem / nW;
    w = rem - b * nW;
    const int p0 = chunk * ppc, p1 = min(Pl, p0 + ppc);
    nt = (p1 - p0 + 31) >> 5;
    np_last = (p1 - p0) - 32 * (nt - 1);
    t = 0;
  }
  __device__ __forceinline__ bool init(int B_, int L_, int Pl_, int ppc_, int n_chunks) {
    B = B_; L = L_; Pl = Pl_; ppc = ppc_; nW = (L + 3) >> 2;
    upc = (long long)B * nW; n_units = upc * n_chunks; u = blockIdx.x;
    if (u >= n_units) return false;
    load_unit();
    return true;
  }
  __device__ __forceinline__ bool next() {     // false at the end of the stream
    if (++t < nt) return true;
    u += gridDim.x;
    if (u >= n_units) return false;
    load_unit();
    return true;
  }
  __device__ __forceinline__ int n_pairs() const { return t + 1 < nt ? 32 : np_last; }     // valid pairs of this tile
  __device__ __forceinline__ int n_sites() const { return min(4, L - 4 * w); }
  __device__ __forceinline__ size_t first_tok() const { return ((size_t)b * Pl + chunk * ppc + 32 * t) * L + 4 * w; }
};

__global__ void __launch_bounds__(C2_THREADS, 1)
k_col_partial_ws(const __grid_constant__ CUtensorMap tmap, const PfAttnTcW* __restrict__ Wt, float* __restrict__ part,
                 float* __restrict__ qcache, int B, int L, int Pl, int ppc, int n_chunks, int* __restrict__ err_flag) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bars = sbase + C2_OFF_BAR;
  auto BAR = [&](int i) { return bars + 8u * (uint32_t)i; };

  for (int i = tid; i < 4096 / 16; i += C2_THREADS)
    reinterpret_cast<int4*>(sm + C2_OFF_BQ)[i] = reinterpret_cast<const int4*>(Wt->wqk_hi)[i];
  for (int i = tid; i < C2_NS * AT_KT_BYTES / 16; i += C2_THREADS) reinterpret_cast<int4*>(sm + C2_OFF_KT)[i] = make_int4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < C2_NS; ++i) {
      mbar_init(BAR(C2_B_XFULL + i), 1);
      mbar_init(BAR(C2_B_XFREE + i), C2_NP1 * 32);
      mbar_init(BAR(C2_B_A1FULL + i), C2_NP1 * 32);
      mbar_init(BAR(C2_B_A1FREE + i), 1);
      mbar_init(BAR(C2_B_QKDONE + i), 1);
      mbar_init(BAR(C2_B_KTFULL + i), 128);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(C2_B_DSFULL + i), 1); mbar_init(BAR(C2_B_DSFREE + i), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 20) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + C2_OFF_TMEM), "r"(C2_TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + C2_OFF_TMEM);
  bool ok = true;
  C2Iter it;
  const bool any = it.init(B, L, Pl, ppc, n_chunks);

  if (warp == 21) {
    // =============================== LOADER ===============================================
    if (any) {
      int n = 0;
      do {
        const int st = n % C2_NS;
        ok = mbar_wait(BAR(C2_B_XFREE + st), (uint32_t)(((n / C2_NS) & 1) ^ 1)) && ok;
        if (elect_one()) {   // one 3-D tensor-map copy per tile: box [32 pairs][4 sites][64 channels], out-of-range elements arrive as zeros
          const uint32_t dst = sbase + C2_OFF_XS + st * 32768, bar = BAR(C2_B_XFULL + st);
          const int c1 = 4 * it.w, c2 = it.b * Pl + it.chunk * ppc + 32 * it.t;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(32768u) : "memory");
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                       ::"r"(dst), "l"(&tmap), "r"(0), "r"(c1), "r"(c2), "r"(bar) : "memory");
        }
        __syncwarp();
        ++n;
      } while (it.next());
    }
  } else if (warp < C2_NP1) {
    // =============================== P1: LayerNorm + split ================================
    const int j = lane & 7, k = lane >> 3;   // lane j of token slot k (= site k of the pair)
    if (any) {
      int n = 0;
      do {
        const int st = n % C2_NS;
        const uint32_t ph = (uint32_t)((n / C2_NS) & 1);
        const int np = it.n_pairs(), ns = it.n_sites();
        ok = mbar_wait(BAR(C2_B_XFULL + st), ph) && ok;
        ok = mbar_wait(BAR(C2_B_A1FREE + st), ph ^ 1) && ok;
        unsigned char* a1 = sm + C2_OFF_A1 + st * AT_A1_BYTES;
        const unsigned char* xs = sm + C2_OFF_XS + st * 32768;
#pragma unroll
        for (int i2 = 0; i2 < 2; ++i2) {
          const int g = (2 * warp + i2) ^ ((k & 1) << 2), r = 32 * k + g;   // odd slots take the pair 4 away: their 64-byte halves of the operand row land on the other banks
          float xv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) xv[i] = 0.f;
          if (g < np && k < ns) load_tok(reinterpret_cast<const float*>(xs + g * 1024 + k * 256), j, xv);
          float nv[8];
          ln_normalize<true>(xv, nv);
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) split2(pk2(nv[2 * i], nv[2 * i + 1]), hi[i], lo[i]);
          // channels 4j..4j+3 -> 16-byte chunk j>>1 (half j&1); 32+4j.. -> chunk 4+(j>>1)
          const uint32_t rowoff = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
          const uint32_t o0 = rowoff + ((((j >> 1)) ^ (r & 7)) << 4) + (j & 1) * 8;
          const uint32_t o1 = rowoff + (((4 + (j >> 1)) ^ (r & 7)) << 4) + (j & 1) * 8;
          *reinterpret_cast<uint2*>(a1 + o0) = make_uint2(hi[0], hi[1]);
          *reinterpret_cast<uint2*>(a1 + o1) = make_uint2(hi[2], hi[3]);
          *reinterpret_cast<uint2*>(a1 + 16384 + o0) = make_uint2(lo[0], lo[1]);
          *reinterpret_cast<uint2*>(a1 + 16384 + o1) = make_uint2(lo[2], lo[3]);
        }
        mbar_arrive(BAR(C2_B_XFREE + st));
        fence_proxy_async_smem();
        mbar_arrive(BAR(C2_B_A1FULL + st));
        ++n;
      } while (it.next());
    }
  } else if (warp == 20) {
    // =============================== MMA ISSUER ===========================================
    if (any) {
      C2Iter nx = it;                 // runs one tile ahead: QK(n+1) is issued before S(n)
      bool has_next = true;
      int n = 0, unit_count = 0;
      ok = mbar_wait(BAR(C2_B_A1FULL + 0), 0) && ok;
      tc_fence_after();
      if (elect_one()) { at_issue_qk(sbase + C2_OFF_A1, sbase + C2_OFF_BQ, tmem); tc_commit(BAR(C2_B_QKDONE + 0)); }
      __syncwarp();
      do {
        const int st = n % C2_NS, par = unit_count & 1;
        const bool first = it.t == 0, last = it.t + 1 == it.nt;
        if (has_next) has_next = nx.next();
        if (has_next) {
          const int s1 = (n + 1) % C2_NS;
          ok = mbar_wait(BAR(C2_B_A1FULL + s1), (uint32_t)((((n + 1) / C2_NS)) & 1)) && ok;
          tc_fence_after();
          if (elect_one()) {
            at_issue_qk(sbase + C2_OFF_A1 + s1 * AT_A1_BYTES, sbase + C2_OFF_BQ, tmem + 16 * s1);
            tc_commit(BAR(C2_B_QKDONE + s1));
          }
          __syncwarp();
        }
        ok = mbar_wait(BAR(C2_B_KTFULL + st), (uint32_t)((n / C2_NS) & 1)) && ok;
        if (first) ok = mbar_wait(BAR(C2_B_DSFREE + par), (uint32_t)(((unit_count >> 1) & 1) ^ 1)) && ok;
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < 4; ++s)
            at_issue_s(sbase + C2_OFF_A1 + st * AT_A1_BYTES, sbase + C2_OFF_KT + st * AT_KT_BYTES,
                       tmem + C2_TM_S + 64 * par + 16 * s, 32 * s, 2, !first);
          tc_commit(BAR(C2_B_A1FREE + st));
          if (last) tc_commit(BAR(C2_B_DSFULL + par));
        }
        __syncwarp();
        if (last) ++unit_count;
        ++n;
      } while (it.next());
    }
  } else if (warp >= 16 && warp < 20) {
    // =============================== P2: phi, sums, k~ operand, unit epilogue ==============
    const int q = warp - 16, ptid = tid - 16 * 32;   // site q, row = TMEM lane = 32 q + lane
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float bq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bq[i] = Wt->bqk[i];
    float* red = reinterpret_cast<float*>(sm + C2_OFF_RED);
    if (any) {
      int n = 0, unit_count = 0;
      float ks[PF_H] = {0.f, 0.f, 0.f, 0.f}, qs[PF_H] = {0.f, 0.f, 0.f, 0.f};
      do {
        const int st = n % C2_NS;
        const bool valid = lane < it.n_pairs() && q < it.n_sites();
        ok = mbar_wait(BAR(C2_B_QKDONE + st), (uint32_t)((n / C2_NS) & 1)) && ok;
        tc_fence_after();
        uint32_t v[8];
        tmem_ld8(tmem + lane_base + 16 * st, v);
        tc_wait_ld();
        float kq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) kq[i] = valid ? phi_elu1(__uint_as_float(v[i]) + bq[i]) : 0.f;
#pragma unroll
        for (int h = 0; h < PF_H; ++h) { ks[h] += kq[h]; qs[h] += kq[4 + h]; }
        if (valid) {
          const size_t tok = it.first_tok() + (size_t)lane * L + q;
          *reinterpret_cast<float4*>(qcache + tok * 4) = make_float4(kq[4], kq[5], kq[6], kq[7]);
        }
        at_store_kt(sm + C2_OFF_KT + st * AT_KT_BYTES, ptid, kq);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(BAR(C2_B_KTFULL + st));
        if (it.t + 1 == it.nt) {     // unit end: D_S -> part
          const int par = unit_count & 1;
          ok = mbar_wait(BAR(C2_B_DSFULL + par), (uint32_t)((unit_count >> 1) & 1)) && ok;
          tc_fence_after();
          float val[4][PF_H];
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            uint32_t d[8];
            tmem_ld8(tmem + lane_base + C2_TM_S + 64 * par + 16 * s, d);
            tc_wait_ld();
#pragma unroll
            for (int h = 0; h < PF_H; ++h) val[s][h] = __uint_as_float(d[h]) + __uint_as_float(d[4 + h]);
          }
          tc_fence_before();
          mbar_arrive(BAR(C2_B_DSFREE + par));
          if (ptid >= 64) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
              for (int h = 0; h < PF_H; ++h) red[(s * PF_H + h) * 64 + (ptid - 64)] = val[s][h];
          }
#pragma unroll
          for (int h = 0; h < PF_H; ++h) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              ks[h] += __shfl_xor_sync(PF_FULL, ks[h], o);
              qs[h] += __shfl_xor_sync(PF_FULL, qs[h], o);
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          float* obase = part + (((size_t)it.chunk * B + it.b) * L + (size_t)4 * it.w) * PF_PART;
          if (ptid < 64) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              if (4 * it.w + s < L) {
#pragma unroll
                for (int h = 0; h < PF_H; ++h)
                  obase[(size_t)s * PF_PART + 8 + h * PF_D + ptid] = val[s][h] + red[(s * PF_H + h) * 64 + ptid];
              }
            }
          }
          if (lane == 0 && 4 * it.w + q < L) {
            float* o = obase + (size_t)q * PF_PART;
#pragma unroll
            for (int h = 0; h < PF_H; ++h) { o[h] = ks[h]; o[4 + h] = qs[h]; }
          }
#pragma unroll
          for (int h = 0; h < PF_H; ++h) { ks[h] = 0.f; qs[h] = 0.f; }
          asm volatile("bar.sync 1, 128;" ::: "memory");   // red is free again
          ++unit_count;
        }
        ++n;
      } while (it.next());
    }
  }
  if (!ok && err_flag != nullptr) *err_flag = 5;
  tc_fence_before();
  __syncthreads();
  if (warp == 20) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C2_TM_COLS) : "memory");
}

inline int pf_attn_tc_init() {
  int rc = (int)cudaFuncSetAttribute(k_col_partial_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_BYTES);
  if (rc == 0) rc = (int)cudaFuncSetAttribute(k_col_partial_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM_BYTES);
  return rc;
}
