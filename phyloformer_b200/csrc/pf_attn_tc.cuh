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
  uint16_t wqk_hl[16 * PF_D];  // the same in one image: hi parts in rows 0..7, lo parts in rows 8..15 (two-pass QK)
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
      o->wqk_hl[off] = hi;
      o->wqk_hl[umma_off_k64(8 + n, k) / 2] = lo;
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
// k_col_partial_ws: the same computation, warp specialised and persistent (one CTA per SM).
//   warp 17        LOADER   per tile two tensor-map TMA copies (cp.async.bulk.tensor.3d, SWIZZLE_128B, box
//                           [32 pairs][4 sites][32 channels]; out-of-range elements arrive as zeros) into a
//                           3-stage ring: stage = [channel half][128 token rows][128 B], 16-byte chunks XOR-
//                           swizzled by the row, so a thread can read a whole row without bank conflicts
//   warps 0..7     P1       two threads per token (lanes l and l^16 own the two 32-channel halves): 8 conflict-free
//                           LDS.128, LayerNorm statistics with one shuffle each, bf16 hi/lo split, 8 STS.128
//                           into the operand image (3-deep ring)
//   warp 16        MMA      one elected lane issues QK(n+1) and S(n)
//   warps 8..11    P2-K     TMEM lane quadrant = site: k~ = phi(k), running sums, k~ operand (bf16 hi/lo)
//   warps 12..15   P2-Q     q~ = phi(q), running sums, q~ cache (16 B per token) for the FFN kernel's column apply;
//                           at the end of a work unit the K warps turn D_S into the chunk's partial sums
// Operand row of token (pair g, site k) of a tile: 32 k + ((g + 2k) & 31): a site's tokens are 32
// consecutive rows (two K = 16 steps of its S GEMM) and 8 consecutive staging rows (2 pairs x 4 sites)
// land on 8 different swizzle phases (conflict-free operand stores).
// The QK GEMM needs two passes only: its B operand holds the hi weights in rows 0..7 and the lo weights in
// rows 8..15, so D_qk[:, j] + D_qk[:, 8+j] = (u_hi + u_lo) . (w_hi + w_lo).
// (k_col_partial_tc above, one thread per token and everything in program order, is kept as the
//  readable restatement and A/B partner: PF_COL_IMPL=tc1.)
// ------------------------------------------------------------------------------------------
#define C2_NP1 8
#define C2_WK 8
#define C2_WQ 12
#define C2_WMMA 16
#define C2_WLD 17
#define C2_THREADS (18 * 32)
#define C2_NS 3                                   // ring depth (staging, operand image, k~ operand, D_qk)
#define C2_OFF_A1 0
#define C2_OFF_XS (C2_NS * AT_A1_BYTES)           // [C2_NS][2 halves][128 rows][128 B]
#define C2_OFF_KT (C2_OFF_XS + C2_NS * 32768)
#define C2_OFF_BQ (C2_OFF_KT + C2_NS * AT_KT_BYTES)   // [16][64] bf16: hi rows 0..7, lo rows 8..15
#define C2_OFF_RED (C2_OFF_BQ + 2048)             // [4 sites][4 heads][64] floats
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

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
// Bounded mbarrier wait of the attention kernels: ~65k suspended polls (tens of milliseconds; a legitimate
// wait is microseconds), and after the first time-out every later wait of the CTA returns at once (the flag
// lives in shared memory), so a broken pipeline drains in milliseconds instead of seconds per wait.
__device__ __forceinline__ bool at_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  if (*abort_flag) return false;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 16); ++spin) {
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
  *abort_flag = 1;
  return false;
}
__device__ __forceinline__ bool elect_one_converged() { __syncwarp(); return elect_one(); }   // elect.sync needs the whole warp
// phi(z) = elu(z) + 1 without the libm exp: ex2.approx of min(z,0) log2(e) (relative error < 1e-6, far
// below the bf16 split of the operands it feeds); branch free.
__device__ __forceinline__ float phi_fast(float z) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(z, 0.f) * 1.4426950408889634f));
  return z > 0.f ? z + 1.0f : e;
}
// 1/sqrt(v) for v >= 1e-5 (never denormal): MUFU.RSQ + one Newton step
__device__ __forceinline__ float rsqrt_nr_ftz(float v) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v));
  return y * fmaf(-0.5f * v * y, y, 1.5f);
}
// QK with the [hi | lo] weight image: two passes (A hi, A lo), 8 MMAs.
__device__ __forceinline__ void at_issue_qk2(uint32_t a1_u32, uint32_t bq_u32, uint32_t d_tmem) {
  const uint32_t idesc = umma_idesc(128, 16);
  const u64 ah = umma_desc(a1_u32), al = umma_desc(a1_u32 + 16384), b = umma_desc(bq_u32);
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const u64 da = t ? al : ah;
#pragma unroll
    for (int s = 0; s < 4; ++s) umma_ss(d_tmem, da + 2 * s, b + 2 * s, idesc, (t | s) ? 1u : 0u);
  }
}

// The P1 step shared by the column and the row kernel: this thread's 32-channel half (hh) of the token in
// staging row sr -> LayerNorm over the 64 channels (partner lane ^16 holds the other half) -> bf16 hi/lo
// -> operand row r.  xs: the stage ([half][row][128 B], TMA SWIZZLE_128B), a1: the operand image.
__device__ __forceinline__ void at_p1_half(const unsigned char* xs, unsigned char* a1, int sr, int r, int hh, bool valid) {
  float xr[32];
  const unsigned char* row = xs + hh * 16384 + sr * 128;
  const int sw = sr & 7;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) v = *reinterpret_cast<const float4*>(row + ((c ^ sw) << 4));
    xr[4 * c] = v.x; xr[4 * c + 1] = v.y; xr[4 * c + 2] = v.z; xr[4 * c + 3] = v.w;
  }
  u64 sa = pk2(0.f, 0.f), sb = pk2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < 32; c += 4) { sa = add2(sa, pk2(xr[c], xr[c + 1])); sb = add2(sb, pk2(xr[c + 2], xr[c + 3])); }
  float s = hsum2(add2(sa, sb));
  s += __shfl_xor_sync(PF_FULL, s, 16);
  const float mean = s * (1.0f / PF_D);
  const u64 nm = pk2(-mean, -mean);
  u64 d[16];
  u64 qa = pk2(0.f, 0.f), qb = pk2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    d[i] = add2(pk2(xr[2 * i], xr[2 * i + 1]), nm);
    d[i + 1] = add2(pk2(xr[2 * i + 2], xr[2 * i + 3]), nm);
    qa = fma2(d[i], d[i], qa); qb = fma2(d[i + 1], d[i + 1], qb);
  }
  float q = hsum2(add2(qa, qb));
  q += __shfl_xor_sync(PF_FULL, q, 16);
  const float rstd = rsqrt_nr_ftz(fmaf(q, 1.0f / PF_D, 1e-5f));
  const u64 rs = pk2(rstd, rstd);
  unsigned char* arow = a1 + (r >> 3) * 1024 + (r & 7) * 128;
  const int rw = r & 7;
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {   // 16-byte operand chunk = 8 channels
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(mul2(d[4 * cc + i], rs), hi[i], lo[i]);
    const uint32_t off = (uint32_t)((((4 * hh + cc) ^ rw) & 7) << 4);
    *reinterpret_cast<uint4*>(arow + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(arow + 16384 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__global__ void __launch_bounds__(C2_THREADS, 1)
k_col_partial_ws(const __grid_constant__ CUtensorMap tmap, const PfAttnTcW* __restrict__ Wt, float* __restrict__ part,
                 float* __restrict__ qcache, int B, int L, int Pl, int ppc, int n_chunks, int* __restrict__ err_flag) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bars = sbase + C2_OFF_BAR;
  auto BAR = [&](int i) { return bars + 8u * (uint32_t)i; };

  for (int i = tid; i < 2048 / 16; i += C2_THREADS)
    reinterpret_cast<int4*>(sm + C2_OFF_BQ)[i] = reinterpret_cast<const int4*>(Wt->wqk_hl)[i];
  for (int i = tid; i < C2_NS * AT_KT_BYTES / 16; i += C2_THREADS) reinterpret_cast<int4*>(sm + C2_OFF_KT)[i] = make_int4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < C2_NS; ++i) {
      mbar_init(BAR(C2_B_XFULL + i), 1);
      mbar_init(BAR(C2_B_XFREE + i), 256);
      mbar_init(BAR(C2_B_A1FULL + i), 256);
      mbar_init(BAR(C2_B_A1FREE + i), 1);
      mbar_init(BAR(C2_B_QKDONE + i), 1);
      mbar_init(BAR(C2_B_KTFULL + i), 256);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(C2_B_DSFULL + i), 1); mbar_init(BAR(C2_B_DSFREE + i), 128); }
    *reinterpret_cast<volatile int*>(sm + C2_OFF_TMEM + 16) = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == C2_WMMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + C2_OFF_TMEM), "r"(C2_TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + C2_OFF_TMEM);
  volatile int* abortf = reinterpret_cast<volatile int*>(sm + C2_OFF_TMEM + 16);
  bool ok = true;

  // The CTA's tile stream: units u = blockIdx.x, + gridDim.x, ... (unit = chunk, msa, 4-site window), tiles
  // of 32 pairs within a unit; n counts the tiles (ring stage n % 3, parity (n / 3) & 1).
  const int nW = (L + 3) >> 2, upc = B * nW, n_units = upc * n_chunks;
  struct Unit { int chunk, b, w, nt, np_last, ns; };
  auto decode = [&](int u) {
    Unit q;
    q.chunk = u / upc;
    const int rem = u - q.chunk * upc;
    q.b = rem / nW;
    q.w = rem - q.b * nW;
    const int p0 = q.chunk * ppc, p1 = min(Pl, p0 + ppc);
    q.nt = (p1 - p0 + 31) >> 5;
    q.np_last = (p1 - p0) - 32 * (q.nt - 1);
    q.ns = min(4, L - 4 * q.w);
    return q;
  };

  if (warp == C2_WLD) {
    // =============================== LOADER ===============================================
    int n = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const Unit un = decode(u);
      for (int t = 0; t < un.nt; ++t, ++n) {
        const int st = n % C2_NS;
        ok = at_wait(BAR(C2_B_XFREE + st), (uint32_t)(((n / C2_NS) & 1) ^ 1), abortf) && ok;
        if (elect_one_converged()) {
          const uint32_t dst = sbase + C2_OFF_XS + st * 32768, bar = BAR(C2_B_XFULL + st);
          const int c1 = 4 * un.w, c2 = un.b * Pl + un.chunk * ppc + 32 * t;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(32768u) : "memory");
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                       ::"r"(dst), "l"(&tmap), "r"(0), "r"(c1), "r"(c2), "r"(bar) : "memory");
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                       ::"r"(dst + 16384), "l"(&tmap), "r"(32), "r"(c1), "r"(c2), "r"(bar) : "memory");
        }
        __syncwarp();
      }
    }
  } else if (warp < C2_NP1) {
    // =============================== P1: LayerNorm + split ================================
    const int hh = lane >> 4, sr = 16 * warp + (lane & 15);
    const int g = sr >> 2, k = sr & 3, r = 32 * k + ((g + 2 * k) & 31);
    int n = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const Unit un = decode(u);
      for (int t = 0; t < un.nt; ++t, ++n) {
        const int st = n % C2_NS;
        const uint32_t ph = (uint32_t)((n / C2_NS) & 1);
        // (every P1 warp processes every tile: a parity wait must see every phase of a stage's barrier -- with two
        //  warp groups alternating tiles over a 3-stage ring a group skipped phases and could be released by the wrong one)
        ok = at_wait(BAR(C2_B_XFULL + st), ph, abortf) && ok;
        const bool valid = g < (t + 1 < un.nt ? 32 : un.np_last) && k < un.ns;
        ok = at_wait(BAR(C2_B_A1FREE + st), ph ^ 1, abortf) && ok;
        at_p1_half(sm + C2_OFF_XS + st * 32768, sm + C2_OFF_A1 + st * AT_A1_BYTES, sr, r, hh, valid);
        mbar_arrive(BAR(C2_B_XFREE + st));
        fence_proxy_async_smem();
        mbar_arrive(BAR(C2_B_A1FULL + st));
      }
    }
  } else if (warp == C2_WMMA) {
    // =============================== MMA ISSUER ===========================================
    if (n_units > (int)blockIdx.x) {
      int n = 0, unit_count = 0;
      ok = at_wait(BAR(C2_B_A1FULL + 0), 0, abortf) && ok;
      tc_fence_after();
      if (elect_one_converged()) { at_issue_qk2(sbase + C2_OFF_A1, sbase + C2_OFF_BQ, tmem); tc_commit(BAR(C2_B_QKDONE + 0)); }
      __syncwarp();
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int chunk = u / upc;
        const int p0 = chunk * ppc, p1 = min(Pl, p0 + ppc), nt = (p1 - p0 + 31) >> 5;
        const int par = unit_count & 1;
        for (int t = 0; t < nt; ++t, ++n) {
          const int st = n % C2_NS;
          const bool first = t == 0, last = t + 1 == nt;
          if (!last || u + (int)gridDim.x < n_units) {   // QK of the next tile of the stream
            const int s1 = (n + 1) % C2_NS;
            ok = at_wait(BAR(C2_B_A1FULL + s1), (uint32_t)(((n + 1) / C2_NS) & 1), abortf) && ok;
            tc_fence_after();
            if (elect_one_converged()) {
              at_issue_qk2(sbase + C2_OFF_A1 + s1 * AT_A1_BYTES, sbase + C2_OFF_BQ, tmem + 16 * s1);
              tc_commit(BAR(C2_B_QKDONE + s1));
            }
            __syncwarp();
          }
          ok = at_wait(BAR(C2_B_KTFULL + st), (uint32_t)((n / C2_NS) & 1), abortf) && ok;
          if (first) ok = at_wait(BAR(C2_B_DSFREE + par), (uint32_t)(((unit_count >> 1) & 1) ^ 1), abortf) && ok;
          tc_fence_after();
          if (elect_one_converged()) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
              at_issue_s(sbase + C2_OFF_A1 + st * AT_A1_BYTES, sbase + C2_OFF_KT + st * AT_KT_BYTES,
                         tmem + C2_TM_S + 64 * par + 16 * s, 32 * s, 2, !first);
            tc_commit(BAR(C2_B_A1FREE + st));
            if (last) tc_commit(BAR(C2_B_DSFULL + par));
          }
          __syncwarp();
        }
        ++unit_count;
      }
    }
  } else if (warp >= C2_WK && warp < C2_WMMA) {
    // =============================== P2: phi, sums, k~ operand / q~ cache, unit epilogue ==
    const bool is_k = warp < C2_WQ;
    const int q = (warp - C2_WK) & 3;                  // TMEM lane quadrant = site
    const int row = 32 * q + lane, g = (lane - 2 * q) & 31;   // operand row, pair within the tile
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int c0 = is_k ? 0 : 4;
    float bq[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) bq[i] = Wt->bqk[c0 + i];
    uint32_t kto[8];                                   // this row's byte offsets in the k~ operand (rows 0..3 hi, 4..7 lo)
    {
      const int kk = row & 63;
#pragma unroll
      for (int nn = 0; nn < 8; ++nn) kto[nn] = (uint32_t)((row >> 6) * 2048 + nn * 128 + ((((kk >> 3) ^ nn) & 7) << 4) + (kk & 7) * 2);
    }
    float* red = reinterpret_cast<float*>(sm + C2_OFF_RED);
    const int ptid = tid - C2_WK * 32;                 // K warps: 0..127
    int n = 0, unit_count = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const Unit un = decode(u);
      float sum[PF_H] = {0.f, 0.f, 0.f, 0.f};
      const size_t tok0 = ((size_t)un.b * Pl + un.chunk * ppc) * L + 4 * un.w + q;
      for (int t = 0; t < un.nt; ++t, ++n) {
        const int st = n % C2_NS;
        const bool valid = g < (t + 1 < un.nt ? 32 : un.np_last) && q < un.ns;
        ok = at_wait(BAR(C2_B_QKDONE + st), (uint32_t)((n / C2_NS) & 1), abortf) && ok;
        tc_fence_after();
        uint32_t va[4], vb[4];
        tmem_ld4(tmem + lane_base + 16 * st + c0, va);
        tmem_ld4(tmem + lane_base + 16 * st + 8 + c0, vb);
        tc_wait_ld();
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i] = valid ? phi_fast((__uint_as_float(va[i]) + __uint_as_float(vb[i])) + bq[i]) : 0.f;
          sum[i] += v[i];
        }
        if (is_k) {
          unsigned char* kt = sm + C2_OFF_KT + st * AT_KT_BYTES;
#pragma unroll
          for (int h = 0; h < PF_H; ++h) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(v[h]);
            const __nv_bfloat16 lo = __float2bfloat16_rn(v[h] - __bfloat162float(hi));
            *reinterpret_cast<__nv_bfloat16*>(kt + kto[h]) = hi;
            *reinterpret_cast<__nv_bfloat16*>(kt + kto[4 + h]) = lo;
          }
          fence_proxy_async_smem();
        } else if (valid) {
          *reinterpret_cast<float4*>(qcache + (tok0 + (size_t)(32 * t + g) * L) * 4) = make_float4(v[0], v[1], v[2], v[3]);
        }
        tc_fence_before();
        mbar_arrive(BAR(C2_B_KTFULL + st));
      }
      // ---- unit end: per-site sums (warp reductions) and, K warps, D_S -> part ----
      const int par = unit_count & 1;
#pragma unroll
      for (int h = 0; h < PF_H; ++h) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum[h] += __shfl_xor_sync(PF_FULL, sum[h], o);
      }
      float* obase = part + (((size_t)un.chunk * B + un.b) * L + (size_t)4 * un.w) * PF_PART;
      if (lane == 0 && q < un.ns) {
        float* o = obase + (size_t)q * PF_PART + c0;
#pragma unroll
        for (int h = 0; h < PF_H; ++h) o[h] = sum[h];
      }
      if (is_k) {
        ok = at_wait(BAR(C2_B_DSFULL + par), (uint32_t)((unit_count >> 1) & 1), abortf) && ok;
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
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (ptid < 64) {
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            if (s < un.ns) {
#pragma unroll
              for (int h = 0; h < PF_H; ++h)
                obase[(size_t)s * PF_PART + 8 + h * PF_D + ptid] = val[s][h] + red[(s * PF_H + h) * 64 + ptid];
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // red is free again
      }
      ++unit_count;
    }
  }
  if (!ok && err_flag != nullptr) *err_flag = 5;
  tc_fence_before();
  __syncthreads();
  if (warp == C2_WMMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C2_TM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------
// k_row_attn_ws: row attention (model.py:90-92) on the same tile engine, in place:
//     x[row, l, :] += M_row qhat_l + bo,      qhat = q~ L / sum_l q~,    M_row from S = sum_l k~ (x) LN(x)
// One persistent CTA per SM walks over its pair-rows (row = blockIdx.x + i gridDim.x).  A row is T tiles of
// 128 consecutive sites (32 KB contiguous, two 2-D tensor-map TMA copies into the swizzled stage).  Per row:
//   pass A  P1 (thread pair per token) -> operand image -> QK GEMM -> P2 (phi; k~ operand; q~ into a
//           shared-memory row cache) -> S GEMM accumulating the whole row in TMEM
//   finalize (the 4 P2-K warps) D_S, sums -> M_row, L / sum q~   (same arithmetic as row_finalize)
//   pass B  P1 warps, 8 lanes per token: y = x + M qhat + bo from the re-staged tile (L2 hit), streaming store
// The item stream of the loader and the P1 warps interleaves the next row's pass A with this row's pass
// B (lag RW_LAG tiles), so the tensor pipe, the finalize and the apply overlap; D_S, M_row and the q~ cache
// are double buffered by row parity.
// Measured and removed (round 2): the apply items on the 8 TMEM-reader warps instead of the P1 warps (which are busy
// 74 % of the kernel with both passes; the readers 19 %): parity green once both groups observe and release every
// ring stage, but no faster (15.9 vs 16.0 ms of row time per forward) -- the kernel moves 12.2 GB per launch at 4.5
// TB/s and is bound by that traffic (2.1 GB of it pass-B re-reads that miss L2), not by any warp role.
// ------------------------------------------------------------------------------------------
#define RW_THREADS C2_THREADS
#define RW_NS 3                                   // staging ring
#define RW_NA 2                                   // operand image / k~ operand / D_qk ring
#ifndef RW_LAG
#define RW_LAG 1                                  // pass B of row i-1 starts this many tiles into row i.  Measured row time per
#endif                                            // forward at lag 5 / 3 / 2 / 1 / 0: 17.1 / 16.4 / 15.9 / 15.3 / 15.5 ms: every tile
                                                  // between a tile's first read and its re-read costs L2 hits (0 exposes the finalize).
                                                  // Two pass-B items per pass-A item (re-read over after half a row): 17.3 ms.
// Order in which pass B walks the tiles of a row.  Reverse (last tile first): the tiles pass A staged last are
// the ones most likely to be L2 hits when the re-read starts, so the part of the row that does fall out of L2 is
// re-fetched once instead of pushing the still-unread part out ahead of the read pointer.
#ifndef RW_B_REVERSE
#define RW_B_REVERSE 1
#endif
#if RW_B_REVERSE
#define RW_B_TILE(j, T) ((T) - 1 - (j))
#else
#define RW_B_TILE(j, T) (j)
#endif
#define RW_OFF_A1 0
#define RW_OFF_XS (RW_NA * AT_A1_BYTES)
#define RW_OFF_KT (RW_OFF_XS + RW_NS * 32768)
#define RW_OFF_BQ (RW_OFF_KT + RW_NA * AT_KT_BYTES)
#define RW_OFF_FIN (RW_OFF_BQ + 2048)             // [2] { M[80][4] (row of channel c at c + c/4: conflict-free for the 8-lane
                                                  //       layout), bo[64], qinv[4], pad } = 2 x 2048 B
#define RW_FIN_BYTES 2048
#define RW_FIN_BO 320
#define RW_FIN_QINV 384
#define RW_OFF_SCR (RW_OFF_FIN + 2 * RW_FIN_BYTES)   // finalize scratch: tot[264] ubar[256] ctxp[256] ctx[64] red[256] wsum[32]
#define RW_SCR_BYTES ((264 + 256 + 256 + 64 + 256 + 32) * 4)
#define RW_OFF_BAR (RW_OFF_SCR + RW_SCR_BYTES)
#define RW_OFF_TMEM (RW_OFF_BAR + 256)
#define RW_OFF_QS (RW_OFF_TMEM + 64)              // [2][Lpad][4] floats, Lpad = T * 128
#define RW_TM_COLS 64
#define RW_TM_S 32                                // D_S[parity] at column 32 + 16 parity; D_qk[i] at 16 i
#define RW_B_XFULL 0
#define RW_B_XFREE 3
#define RW_B_A1FULL 6
#define RW_B_A1FREE 8
#define RW_B_QKDONE 10
#define RW_B_KTFULL 12
#define RW_B_DSFULL 14
#define RW_B_DSFREE 16
#define RW_B_QSUM 18
#define RW_B_FINDONE 20
#define RW_B_FINFREE 22
inline size_t rw_smem_bytes(int L) { return (size_t)RW_OFF_QS + 2 * (size_t)((L + 127) / 128) * 128 * 16 + 1024; }

__global__ void __launch_bounds__(RW_THREADS, 1)
k_row_attn_ws(const __grid_constant__ CUtensorMap tmap, const PfAttnW* __restrict__ W, const PfAttnTcW* __restrict__ Wt,
              float* __restrict__ x, int rows, int L, int* __restrict__ err_flag) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bars = sbase + RW_OFF_BAR;
  auto BAR = [&](int i) { return bars + 8u * (uint32_t)i; };
  const int T = (L + 127) >> 7, Lpad = T * 128;
  const int m = (rows - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // rows of this CTA
  const int lag = RW_LAG;

  for (int i = tid; i < 2048 / 16; i += RW_THREADS)
    reinterpret_cast<int4*>(sm + RW_OFF_BQ)[i] = reinterpret_cast<const int4*>(Wt->wqk_hl)[i];
  for (int i = tid; i < RW_NA * AT_KT_BYTES / 16; i += RW_THREADS) reinterpret_cast<int4*>(sm + RW_OFF_KT)[i] = make_int4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < RW_NS; ++i) { mbar_init(BAR(RW_B_XFULL + i), 1); mbar_init(BAR(RW_B_XFREE + i), 256); }
    for (int i = 0; i < RW_NA; ++i) {
      mbar_init(BAR(RW_B_A1FULL + i), 256); mbar_init(BAR(RW_B_A1FREE + i), 1);
      mbar_init(BAR(RW_B_QKDONE + i), 1);   mbar_init(BAR(RW_B_KTFULL + i), 256);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(RW_B_DSFULL + i), 1);    mbar_init(BAR(RW_B_DSFREE + i), 128);
      mbar_init(BAR(RW_B_QSUM + i), 128);    mbar_init(BAR(RW_B_FINDONE + i), 128);
      mbar_init(BAR(RW_B_FINFREE + i), 256 * T);
    }
    *reinterpret_cast<volatile int*>(sm + RW_OFF_TMEM + 16) = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == C2_WMMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + RW_OFF_TMEM), "r"(RW_TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + RW_OFF_TMEM);
  volatile int* abortf = reinterpret_cast<volatile int*>(sm + RW_OFF_TMEM + 16);
  float* qs = reinterpret_cast<float*>(sm + RW_OFF_QS);
  bool ok = true;

  if (warp == C2_WLD) {
    // =============================== LOADER ===============================================
    int nx = 0;
    // L2 policies: pass-A tiles are read again one row later (pass B) -> evict_last; pass-B tiles are dead after
    // the read -> evict_first (the result goes out with streaming stores), so the re-read stays an L2 hit
    unsigned long long pol_keep, pol_drop;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_drop));
    auto load = [&](int i, int t, unsigned long long pol) {
      const int st = nx % RW_NS;
      ok = at_wait(BAR(RW_B_XFREE + st), (uint32_t)(((nx / RW_NS) & 1) ^ 1), abortf) && ok;
      if (elect_one_converged()) {
        const uint32_t dst = sbase + RW_OFF_XS + st * 32768, bar = BAR(RW_B_XFULL + st);
        const long long tok0 = (long long)(blockIdx.x + (long long)i * gridDim.x) * L + 128 * t;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(32768u) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                     ::"r"(dst), "l"(&tmap), "r"(0), "r"((int)tok0), "r"(bar), "l"(pol) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                     ::"r"(dst + 16384), "l"(&tmap), "r"(32), "r"((int)tok0), "r"(bar), "l"(pol) : "memory");
      }
      __syncwarp();
      ++nx;
    };
    for (int i = 0; i <= m; ++i)
      for (int k = 0; k < T + lag; ++k) {
        if (i < m && k < T) load(i, k, pol_keep);
        if (i >= 1 && k >= lag) load(i - 1, RW_B_TILE(k - lag, T), pol_drop);
      }
  } else if (warp < C2_NP1) {
    // =============================== P1: pass A (LN + split) and pass B (apply) ============
    const int wi = warp, hh = lane >> 4, sr = 16 * wi + (lane & 15);
    const int j = lane & 7, slot = lane >> 3;
    int nx = 0, na = 0;
    for (int i = 0; i <= m; ++i)
      for (int k = 0; k < T + lag; ++k) {
        if (i < m && k < T) {            // ---- item A(i, k) ----
          {
            const int st = nx % RW_NS, ab = na % RW_NA;
            ok = at_wait(BAR(RW_B_XFULL + st), (uint32_t)((nx / RW_NS) & 1), abortf) && ok;
            ok = at_wait(BAR(RW_B_A1FREE + ab), (uint32_t)(((na / RW_NA) & 1) ^ 1), abortf) && ok;
            at_p1_half(sm + RW_OFF_XS + st * 32768, sm + RW_OFF_A1 + ab * AT_A1_BYTES, sr, sr, hh, 128 * k + sr < L);
            mbar_arrive(BAR(RW_B_XFREE + st));
            fence_proxy_async_smem();
            mbar_arrive(BAR(RW_B_A1FULL + ab));
          }
          ++nx; ++na;
        }
        if (i >= 1 && k >= lag) {        // ---- item B(i - 1, k - lag) ----
          {
            const int st = nx % RW_NS, ib = i - 1, t = RW_B_TILE(k - lag, T), par = ib & 1;
            ok = at_wait(BAR(RW_B_XFULL + st), (uint32_t)((nx / RW_NS) & 1), abortf) && ok;
            ok = at_wait(BAR(RW_B_FINDONE + par), (uint32_t)((ib >> 1) & 1), abortf) && ok;
            const float* fin = reinterpret_cast<const float*>(sm + RW_OFF_FIN + par * RW_FIN_BYTES);
            float4 Mr[8];
            float bo[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int ch = chan_of(j, c);
              Mr[c] = *reinterpret_cast<const float4*>(fin + 4 * (ch + (ch >> 2)));
              bo[c] = fin[RW_FIN_BO + ch];
            }
            const float4 qi = *reinterpret_cast<const float4*>(fin + RW_FIN_QINV);
            const unsigned char* xs = sm + RW_OFF_XS + st * 32768;
            float* xrow = x + (size_t)(blockIdx.x + (size_t)ib * gridDim.x) * L * PF_D;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int tk = 16 * wi + 4 * it + slot, site = 128 * t + tk;
              if (site < L) {
                const unsigned char* row = xs + tk * 128 + ((j ^ (tk & 7)) << 4);
                const float4 a = *reinterpret_cast<const float4*>(row);
                const float4 b = *reinterpret_cast<const float4*>(row + 16384);
                float4 q = *reinterpret_cast<const float4*>(qs + ((size_t)par * Lpad + site) * 4);
                q.x *= qi.x; q.y *= qi.y; q.z *= qi.z; q.w *= qi.w;
                float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  float acc = bo[c];
                  acc = fmaf(Mr[c].x, q.x, acc);
                  acc = fmaf(Mr[c].y, q.y, acc);
                  acc = fmaf(Mr[c].z, q.z, acc);
                  acc = fmaf(Mr[c].w, q.w, acc);
                  xv[c] += acc;
                }
                float4* dst = reinterpret_cast<float4*>(xrow + (size_t)site * PF_D);
                __stcs(dst + j, make_float4(xv[0], xv[1], xv[2], xv[3]));
                __stcs(dst + 8 + j, make_float4(xv[4], xv[5], xv[6], xv[7]));
              }
            }
            mbar_arrive(BAR(RW_B_XFREE + st));
            mbar_arrive(BAR(RW_B_FINFREE + par));
          }
          ++nx;
        }
      }
  } else if (warp == C2_WMMA) {
    // =============================== MMA ISSUER ===========================================
    if (m > 0) {
      int n = 0;
      ok = at_wait(BAR(RW_B_A1FULL + 0), 0, abortf) && ok;
      tc_fence_after();
      if (elect_one_converged()) { at_issue_qk2(sbase + RW_OFF_A1, sbase + RW_OFF_BQ, tmem); tc_commit(BAR(RW_B_QKDONE + 0)); }
      __syncwarp();
      for (int i = 0; i < m; ++i) {
        const int par = i & 1;
        for (int t = 0; t < T; ++t, ++n) {
          const int ab = n % RW_NA;
          if (t + 1 < T || i + 1 < m) {
            const int a1 = (n + 1) % RW_NA;
            ok = at_wait(BAR(RW_B_A1FULL + a1), (uint32_t)(((n + 1) / RW_NA) & 1), abortf) && ok;
            tc_fence_after();
            if (elect_one_converged()) {
              at_issue_qk2(sbase + RW_OFF_A1 + a1 * AT_A1_BYTES, sbase + RW_OFF_BQ, tmem + 16 * a1);
              tc_commit(BAR(RW_B_QKDONE + a1));
            }
            __syncwarp();
          }
          ok = at_wait(BAR(RW_B_KTFULL + ab), (uint32_t)((n / RW_NA) & 1), abortf) && ok;
          if (t == 0) ok = at_wait(BAR(RW_B_DSFREE + par), (uint32_t)(((i >> 1) & 1) ^ 1), abortf) && ok;
          tc_fence_after();
          if (elect_one_converged()) {
            at_issue_s(sbase + RW_OFF_A1 + ab * AT_A1_BYTES, sbase + RW_OFF_KT + ab * AT_KT_BYTES, tmem + RW_TM_S + 16 * par, 0, 8, t > 0);
            tc_commit(BAR(RW_B_A1FREE + ab));
            if (t + 1 == T) tc_commit(BAR(RW_B_DSFULL + par));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= C2_WK && warp < C2_WMMA) {
    // =============================== P2: phi, sums, k~ operand / q~ row cache, finalize ====
    const bool is_k = warp < C2_WQ;
    const int q = (warp - C2_WK) & 3, row = 32 * q + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int c0 = is_k ? 0 : 4;
    float bq[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) bq[i] = Wt->bqk[c0 + i];
    uint32_t kto[8];
    {
      const int kk = row & 63;
#pragma unroll
      for (int nn = 0; nn < 8; ++nn) kto[nn] = (uint32_t)((row >> 6) * 2048 + nn * 128 + ((((kk >> 3) ^ nn) & 7) << 4) + (kk & 7) * 2);
    }
    float* scr = reinterpret_cast<float*>(sm + RW_OFF_SCR);
    float* tot = scr;                 // [264]
    float* ubar = scr + 264;          // [4][64]
    float* ctxp = scr + 520;          // [4][64]
    float* ctx = scr + 776;           // [64]
    float* red = scr + 840;           // [4][64]
    float* wsum = scr + 1096;         // [8 warps][4]
    const int ptid = tid - C2_WK * 32;
    int n = 0;
    for (int i = 0; i < m; ++i) {
      const int par = i & 1;
      const uint32_t rph = (uint32_t)((i >> 1) & 1);
      float sum[PF_H] = {0.f, 0.f, 0.f, 0.f};
      if (!is_k) ok = at_wait(BAR(RW_B_FINFREE + par), rph ^ 1, abortf) && ok;    // pass B of row i-2 has read this q~ buffer
      for (int t = 0; t < T; ++t, ++n) {
        const int ab = n % RW_NA, site = 128 * t + row;
        const bool valid = site < L;
        ok = at_wait(BAR(RW_B_QKDONE + ab), (uint32_t)((n / RW_NA) & 1), abortf) && ok;
        tc_fence_after();
        uint32_t va[4], vb[4];
        tmem_ld4(tmem + lane_base + 16 * ab + c0, va);
        tmem_ld4(tmem + lane_base + 16 * ab + 8 + c0, vb);
        tc_wait_ld();
        float v[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          v[h] = valid ? phi_fast((__uint_as_float(va[h]) + __uint_as_float(vb[h])) + bq[h]) : 0.f;
          sum[h] += v[h];
        }
        if (is_k) {
          unsigned char* kt = sm + RW_OFF_KT + ab * AT_KT_BYTES;
#pragma unroll
          for (int h = 0; h < PF_H; ++h) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(v[h]);
            const __nv_bfloat16 lo = __float2bfloat16_rn(v[h] - __bfloat162float(hi));
            *reinterpret_cast<__nv_bfloat16*>(kt + kto[h]) = hi;
            *reinterpret_cast<__nv_bfloat16*>(kt + kto[4 + h]) = lo;
          }
          fence_proxy_async_smem();
        } else {
          *reinterpret_cast<float4*>(qs + ((size_t)par * Lpad + site) * 4) = make_float4(v[0], v[1], v[2], v[3]);
        }
        tc_fence_before();
        mbar_arrive(BAR(RW_B_KTFULL + ab));
      }
      // ---- row end: sums of k~ / q~ per warp, then (K warps) finalize ----
#pragma unroll
      for (int h = 0; h < PF_H; ++h) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum[h] += __shfl_xor_sync(PF_FULL, sum[h], o);
      }
      if (!is_k) {
        // wsum is rewritten per row: the K warps have read the previous row's values before they arrive on
        // FINDONE, and this row's Q sums are only complete after its last tile, i.e. after that finalize
        // (K and Q warps process tiles in lock step through KTFULL/QKDONE; one row of slack is guarded below)
        if (i > 0) ok = at_wait(BAR(RW_B_FINDONE + (par ^ 1)), (uint32_t)(((i - 1) >> 1) & 1), abortf) && ok;
        if (lane == 0) {
#pragma unroll
          for (int h = 0; h < PF_H; ++h) wsum[(4 + q) * 4 + h] = sum[h];
        }
        mbar_arrive(BAR(RW_B_QSUM + par));
      } else {
        ok = at_wait(BAR(RW_B_DSFULL + par), rph, abortf) && ok;
        tc_fence_after();
        uint32_t d[8];
        tmem_ld8(tmem + lane_base + RW_TM_S + 16 * par, d);
        tc_wait_ld();
        float val[PF_H];
#pragma unroll
        for (int h = 0; h < PF_H; ++h) val[h] = __uint_as_float(d[h]) + __uint_as_float(d[4 + h]);
        tc_fence_before();
        mbar_arrive(BAR(RW_B_DSFREE + par));
        if (lane == 0) {
#pragma unroll
          for (int h = 0; h < PF_H; ++h) wsum[q * 4 + h] = sum[h];
        }
        if (ptid >= 64) {
#pragma unroll
          for (int h = 0; h < PF_H; ++h) red[h * 64 + (ptid - 64)] = val[h];
        }
        ok = at_wait(BAR(RW_B_QSUM + par), rph, abortf) && ok;
        ok = at_wait(BAR(RW_B_FINFREE + par), rph ^ 1, abortf) && ok;     // pass B of row i-2 has read M / bo / qinv of this parity
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (ptid < 64) {
#pragma unroll
          for (int h = 0; h < PF_H; ++h) tot[8 + h * PF_D + ptid] = val[h] + red[h * 64 + ptid];
        }
        if (ptid < 8) {
          const int w0 = (ptid >> 2) * 4, h = ptid & 3;
          tot[ptid] = ((wsum[(w0 + 0) * 4 + h] + wsum[(w0 + 1) * 4 + h]) + wsum[(w0 + 2) * 4 + h]) + wsum[(w0 + 3) * 4 + h];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        float* fin = reinterpret_cast<float*>(sm + RW_OFF_FIN + par * RW_FIN_BYTES);   // M[80][4] | bo[64] | qinv[4]
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int idx = ptid + 128 * e, h = idx >> 6, c = idx & 63;
          ubar[h * PF_D + c] = fmaf(W->gamma[c], tot[8 + h * PF_D + c] / tot[h], W->beta[c]);
        }
        if (ptid < PF_H) fin[RW_FIN_QINV + ptid] = (float)L / tot[4 + ptid];
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int idx = ptid + 128 * e, o = idx & 63, prt = idx >> 6, h = o >> 4;
          float acc = 0.f;
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) acc = fmaf(W->wvT[16 * prt + kk][o], ubar[h * PF_D + 16 * prt + kk], acc);
          ctxp[prt * PF_D + o] = acc;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (ptid < PF_D) {
          ctx[ptid] = (((W->bv[ptid] + ctxp[ptid]) + ctxp[PF_D + ptid]) + ctxp[2 * PF_D + ptid]) + ctxp[3 * PF_D + ptid];
          fin[RW_FIN_BO + ptid] = W->bo[ptid];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int idx = ptid + 128 * e, c = idx >> 2, h = idx & 3;
          float acc = 0.f;
          const float4* wrow = reinterpret_cast<const float4*>(&W->wo[c][h * PF_DH]);   // 4 x LDG.128 instead of 16 strided LDG.32
#pragma unroll
          for (int e4 = 0; e4 < PF_DH / 4; ++e4) {
            const float4 w = wrow[e4];
            acc = fmaf(w.x, ctx[h * PF_DH + 4 * e4 + 0], acc);
            acc = fmaf(w.y, ctx[h * PF_DH + 4 * e4 + 1], acc);
            acc = fmaf(w.z, ctx[h * PF_DH + 4 * e4 + 2], acc);
            acc = fmaf(w.w, ctx[h * PF_DH + 4 * e4 + 3], acc);
          }
          fin[(c + (c >> 2)) * 4 + h] = acc;
        }
        mbar_arrive(BAR(RW_B_FINDONE + par));
      }
    }
  }
  if (!ok && err_flag != nullptr) *err_flag = 6;
  tc_fence_before();
  __syncthreads();
  if (warp == C2_WMMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(RW_TM_COLS) : "memory");
}

inline int pf_attn_tc_init() {
  int rc = (int)cudaFuncSetAttribute(k_col_partial_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_BYTES);
  if (rc == 0) rc = (int)cudaFuncSetAttribute(k_col_partial_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM_BYTES);
  if (rc == 0) rc = (int)cudaFuncSetAttribute(k_row_attn_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  return rc;
}
