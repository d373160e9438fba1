// pf_kernels.cuh -- CUDA-core kernels of the Phyloformer forward path (sm_100a):
//   k_onehot_to_idx, k_embed_sequences   input conversion            (data.py:28-29, model.py:173)
//   k_row_attn<MODE>                     pair embedding + row attention, in place (model.py:175, 90-92)
//   k_col_partial, k_col_reduce, k_col_finalize   column attention summaries      (model.py:96-97)
//   k_colapply_ffn_fp32                  column apply + FFN, fp32 FFMA            (model.py:97-104)
//   k_head                               pwFNN + softplus + site mean             (model.py:182-185)
// All of them are HBM-bound streaming kernels except the FFN contraction; see DESIGN.md.
#pragma once
#include "pf_common.cuh"

// ------------------------------------------------------------------------------------------
// (B,22,L,n) fp32 one-hot -> (B,n,L) uint8 codes; sets *flag if some column is not one-hot.
// ------------------------------------------------------------------------------------------
__global__ void k_onehot_to_idx(const float* __restrict__ x, int B, int L, int n,
                                uint8_t* __restrict__ idx, int* __restrict__ flag) {
  const long long total = (long long)B * L * n;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int i = (int)(t % n);
  const int l = (int)((t / n) % L);
  const int b = (int)(t / ((long long)n * L));
  const float* px = x + ((long long)b * PF_NCHAR * L + l) * n + i;
  int code = 0, ones = 0, other = 0;
#pragma unroll
  for (int c = 0; c < PF_NCHAR; ++c) {
    const float v = px[(long long)c * L * n];
    if (v == 1.0f) { code = c; ++ones; }
    else if (v != 0.0f) ++other;
  }
  idx[((long long)b * n + i) * L + l] = (uint8_t)code;
  if (ones != 1 || other != 0) *flag = 1;  // benign race: every writer stores 1
}

// Soft-input path: E[b][i][l][:] = relu(W_e x[b,:,l,i] + b_e)      (model.py:138-143,173)
// Skipped (early exit) when the input was one-hot.
__global__ void k_embed_sequences(const PfHeadW* __restrict__ hw, const float* __restrict__ x,
                                  const int* __restrict__ flag, int B, int L, int n,
                                  float* __restrict__ emb) {
  if (flag == nullptr || *flag == 0) return;
  const long long total = (long long)B * n * L * PF_D;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % PF_D);
  const int l = (int)((t / PF_D) % L);
  const int i = (int)((t / ((long long)PF_D * L)) % n);
  const int b = (int)(t / ((long long)PF_D * L * n));
  float acc = hw->be[c];
#pragma unroll
  for (int a = 0; a < PF_NCHAR; ++a)
    acc = fmaf(hw->weT[a][c], x[(((long long)b * PF_NCHAR + a) * L + l) * n + i], acc);
  emb[t] = fmaxf(acc, 0.f);
}

// ------------------------------------------------------------------------------------------
// Row attention over one pair-row (attends along the L sites of a pair), in place:
//   x[b,p,l,:] += out_proj( qhat_l * ctx )          attention.py:160-197 with R=P, N=L
// collapsed as in SURVEY.md section 3.2:  ctx_h = Wv[h](sum_l khat_l u_l) + bv[h],  y_l = M qhat_l + bo
// with M[:,h] = Wo[:,h-slice] ctx_h.   One CTA per (msa, pair); 32 token slots x 8 lanes.
// MODE 0: x is read from HBM.   MODE 1: x0 = pair embedding, computed on the fly from the
// uint8 MSA (table lookup) or, for soft inputs, from the per-sequence embedding E.
// Phase A: LN, q/k, per-row sums.  Finalize: M, qinv.  Phase B: apply + residual, write x.
// ------------------------------------------------------------------------------------------
struct RowSmem {
  float red[8][PF_PART];   // per-warp partial sums
  float tot[PF_PART];
  float ubar[PF_H][PF_D];
  float ctxp[4][PF_D];     // ctx partial sums over 16-wide slices of the contraction
  float ctx[PF_D];
  float M[PF_D][PF_H];
  float qinv[PF_H];
  float table[PF_NCHAR][PF_D];
};

// Per-row finalize (all 256 threads): totals -> ubar = LN-affine(sum k u / sum k) -> ctx = W_v ubar + b_v
// -> M[:,h] = W_o[:, h-slice] ctx_h, qinv = L / sum q.  The 64x64 contraction for ctx is split
// over 4 thread groups (16 terms each) and summed in a fixed order: with one 64-term chain per
// thread this section was ~40 % of a CTA's lifetime on short rows (L = 200).
__device__ __forceinline__ void row_finalize(const PfAttnW* __restrict__ W, RowSmem& sm, int tid, int L) {
  {
    const int h = tid >> 6, c = tid & 63;
    sm.ubar[h][c] = fmaf(W->gamma[c], sm.tot[8 + h * PF_D + c] / sm.tot[h], W->beta[c]);
  }
  if (tid < PF_H) sm.qinv[tid] = (float)L / sm.tot[4 + tid];
  __syncthreads();
  {
    const int o = tid & 63, part = tid >> 6, h = o >> 4;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(W->wvT[16 * part + k][o], sm.ubar[h][16 * part + k], acc);
    sm.ctxp[part][o] = acc;
  }
  __syncthreads();
  if (tid < PF_D) sm.ctx[tid] = (((W->bv[tid] + sm.ctxp[0][tid]) + sm.ctxp[1][tid]) + sm.ctxp[2][tid]) + sm.ctxp[3][tid];
  __syncthreads();
  {
    const int c = tid >> 2, h = tid & 3;
    float acc = 0.f;
    const float4* wrow = reinterpret_cast<const float4*>(&W->wo[c][h * PF_DH]);   // 4 x LDG.128 (a thread's 16 weights are 64 contiguous bytes)
#pragma unroll
    for (int e4 = 0; e4 < PF_DH / 4; ++e4) {
      const float4 w = wrow[e4];
      acc = fmaf(w.x, sm.ctx[h * PF_DH + 4 * e4 + 0], acc);
      acc = fmaf(w.y, sm.ctx[h * PF_DH + 4 * e4 + 1], acc);
      acc = fmaf(w.z, sm.ctx[h * PF_DH + 4 * e4 + 2], acc);
      acc = fmaf(w.w, sm.ctx[h * PF_DH + 4 * e4 + 3], acc);
    }
    sm.M[c][h] = acc;
  }
  __syncthreads();
}

template <int MODE>
__device__ __forceinline__ void row_fetch(const float* __restrict__ xrow, const RowSmem& sm,
                                          const uint8_t* __restrict__ si, const uint8_t* __restrict__ sj,
                                          const float* __restrict__ ei, const float* __restrict__ ej,
                                          bool soft, int l, int j, float (&x)[8]) {
  if (MODE == 0) {
    load_tok(xrow + (size_t)l * PF_D, j, x);
  } else if (!soft) {
    const int a = si[l], b = sj[l];
    const float4 a0 = reinterpret_cast<const float4*>(sm.table[a])[j];
    const float4 a1 = reinterpret_cast<const float4*>(sm.table[a])[8 + j];
    const float4 b0 = reinterpret_cast<const float4*>(sm.table[b])[j];
    const float4 b1 = reinterpret_cast<const float4*>(sm.table[b])[8 + j];
    x[0] = a0.x + b0.x; x[1] = a0.y + b0.y; x[2] = a0.z + b0.z; x[3] = a0.w + b0.w;
    x[4] = a1.x + b1.x; x[5] = a1.y + b1.y; x[6] = a1.z + b1.z; x[7] = a1.w + b1.w;
  } else {
    float u[8], w[8];
    load_tok(ei + (size_t)l * PF_D, j, u);
    load_tok(ej + (size_t)l * PF_D, j, w);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = u[i] + w[i];
  }
}

template <int MODE>
__global__ void __launch_bounds__(256, 2)
k_row_attn(const PfAttnW* __restrict__ W, const PfHeadW* __restrict__ hw, float* __restrict__ x,
           const uint8_t* __restrict__ msa, const float* __restrict__ semb,
           const int* __restrict__ soft_flag, int n, int L, long long pair_lo, int Pl, int embed_only) {
  if (MODE == 1 && embed_only == 2 && (soft_flag == nullptr || *soft_flag == 0)) return;   // residue codes: k_row_attn_combo did this row
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RowSmem& sm = *reinterpret_cast<RowSmem*>(smem_raw);
  float* qcache = reinterpret_cast<float*>(smem_raw + sizeof(RowSmem));  // [L][4]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = lane & 7, slot = tid >> 3;  // 32 token slots
  const int row = blockIdx.x;
  const int b = row / Pl, pl = row - b * Pl;
  float* xrow = x + (size_t)row * L * PF_D;

  const uint8_t *si = nullptr, *sj = nullptr;
  const float *ei = nullptr, *ej = nullptr;
  bool soft = false;
  if (MODE == 1) {
    int pi, pj;
    pair_to_ij(pair_lo + pl, n, &pi, &pj);
    soft = (soft_flag != nullptr) && (*soft_flag != 0);
    si = msa + ((size_t)b * n + pi) * L;
    sj = msa + ((size_t)b * n + pj) * L;
    if (soft) {
      ei = semb + ((size_t)b * n + pi) * L * PF_D;
      ej = semb + ((size_t)b * n + pj) * L * PF_D;
    } else {
      for (int t = tid; t < PF_NCHAR * PF_D; t += 256) (&sm.table[0][0])[t] = (&hw->table[0][0])[t];
    }
    __syncthreads();
  }

  // folded q/k weights for this lane's 8 channels
  float wqk[8][8];
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    const float4 a = reinterpret_cast<const float4*>(W->wqk[v])[j];
    const float4 c = reinterpret_cast<const float4*>(W->wqk[v])[8 + j];
    wqk[v][0] = a.x; wqk[v][1] = a.y; wqk[v][2] = a.z; wqk[v][3] = a.w;
    wqk[v][4] = c.x; wqk[v][5] = c.y; wqk[v][6] = c.z; wqk[v][7] = c.w;
  }
  const float bias_j = W->bqk[j];

  // ---------------- phase A ----------------
  float S[PF_H][8];
#pragma unroll
  for (int h = 0; h < PF_H; ++h)
#pragma unroll
    for (int i = 0; i < 8; ++i) S[h][i] = 0.f;
  float own = 0.f;  // lane j<4: sum k_j ; lane j>=4: sum q_{j-4}

  const int n_it = (L + 31) >> 5;
  float xn[8];
  {
    const int l0 = slot;
    if (l0 < L) row_fetch<MODE>(xrow, sm, si, sj, ei, ej, soft, l0, j, xn);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) xn[i] = 0.f;
    }
  }
  for (int it = 0; it < n_it; ++it) {
    const int l = it * 32 + slot;
    const bool act = l < L;
    float xc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xc[i] = xn[i];
    const int ln = l + 32;
    if (it + 1 < n_it) {
      if (ln < L) row_fetch<MODE>(xrow, sm, si, sj, ei, ej, soft, ln, j, xn);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) xn[i] = 0.f;
      }
    }
    float nv[8];
    ln_normalize<true>(xc, nv);
    float part[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(wqk[v][i], nv[i], s);
      part[v] = s;
    }
    const float mine = phi_elu1(grp_reduce8(part, j) + bias_j);
    float kh[PF_H];
#pragma unroll
    for (int h = 0; h < PF_H; ++h) kh[h] = __shfl_sync(PF_FULL, mine, (lane & 24) | h);
    if (act) {
      own += mine;
      if (j >= 4) qcache[(size_t)l * 4 + (j - 4)] = mine;
#pragma unroll
      for (int h = 0; h < PF_H; ++h)
#pragma unroll
        for (int i = 0; i < 8; ++i) S[h][i] = fmaf(kh[h], nv[i], S[h][i]);
    }
  }
  // reduce the 4 slots of a warp (fixed tree), then the 8 warps (fixed order)
#pragma unroll
  for (int h = 0; h < PF_H; ++h)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = S[h][i];
      v += __shfl_xor_sync(PF_FULL, v, 8);
      v += __shfl_xor_sync(PF_FULL, v, 16);
      S[h][i] = v;
    }
  own += __shfl_xor_sync(PF_FULL, own, 8);
  own += __shfl_xor_sync(PF_FULL, own, 16);
  if (lane < 8) {
    sm.red[warp][j] = own;
#pragma unroll
    for (int h = 0; h < PF_H; ++h)
#pragma unroll
      for (int i = 0; i < 8; ++i) sm.red[warp][8 + h * PF_D + chan_of(j, i)] = S[h][i];
  }
  __syncthreads();
  for (int t = tid; t < PF_PART; t += 256) {
    float s = sm.red[0][t];
#pragma unroll
    for (int w = 1; w < 8; ++w) s += sm.red[w][t];
    sm.tot[t] = s;
  }
  __syncthreads();
  row_finalize(W, sm, tid, L);
  // ---------------- phase B: y = x + M qhat + bo ----------------
  float4 Mr[8];
  float bo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = chan_of(j, i);
    Mr[i] = *reinterpret_cast<const float4*>(sm.M[c]);
    bo[i] = W->bo[c];
  }
  const float4 qi = *reinterpret_cast<const float4*>(sm.qinv);
  for (int l = slot; l < L; l += 32) {
    float xc[8];
    row_fetch<MODE>(xrow, sm, si, sj, ei, ej, soft, l, j, xc);
    float4 q = *reinterpret_cast<const float4*>(qcache + (size_t)l * 4);
    q.x *= qi.x; q.y *= qi.y; q.z *= qi.z; q.w *= qi.w;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = bo[i];
      a = fmaf(Mr[i].x, q.x, a);
      a = fmaf(Mr[i].y, q.y, a);
      a = fmaf(Mr[i].z, q.z, a);
      a = fmaf(Mr[i].w, q.w, a);
      if (embed_only != 1) xc[i] += a;
    }
    store_tok(xrow + (size_t)l * PF_D, j, xc);
  }
}

// ------------------------------------------------------------------------------------------
// k_row_attn_combo: block-0 row attention for residue-code inputs (the normal case), from residue-PAIR tables.
// x0[l] = T[a_l] + T[b_l] takes one of only 22 x 22 = 484 values ("combos") along a pair-row, so everything the
// row attention needs per token -- LN(x0), k~, q~ -- is a function of the combo and is tabulated once per
// checkpoint (pf_create, fp64 on the host: Row0Tab).  Per pair-row (one CTA):
//   1. histogram of the row's combos (integer shared-memory atomics: exact, order-free)
//   2. sums over the PRESENT combos in ascending combo order: S[h][c] = sum cnt K[h] N[c], sum k~, sum q~
//   3. row_finalize (as every row kernel): M_row, L / sum q~
//   4. Y[combo] = x0 + M qhat + bo for the present combos (R0_YCAP at a time), then the row is written by a pure
//      shared-memory gather + coalesced 16-byte stores: the only per-token work left is the 256-byte write
// against ~1.1 k instructions per token in k_row_attn<1> (3.1 ms for a write-only 5.1 GB pass at 200 x 1000).
// Identical sequence pairs give bit-identical rows by construction.
// ------------------------------------------------------------------------------------------
#define PF_NCOMBO (PF_NCHAR * PF_NCHAR)
#define R0_YCAP 128
#define R0_MAXP 256              // unordered residue pairs: 22 * 23 / 2 = 253 distinct combos at most
struct Row0Tab {
  float kq[PF_NCOMBO][8];      // k~[4] | q~[4] = phi(wqk . LN(x0) + bqk)
  float n[PF_NCOMBO][PF_D];    // LN(x0), no affine
};
struct Row0Smem {
  RowSmem r;                   // tot / ubar / ctx / M / qinv / table; r.red is the 4-way partial buffer of step 2
  float Y[R0_YCAP][PF_D];
  float pk[R0_MAXP][8];        // per present combo (list order): cnt k~[4] | cnt q~[4]
  float qh[R0_MAXP][4];        // per present combo: q~[4] (for Y)
  int cnt[PF_NCOMBO];
  short slot_of[PF_NCOMBO];
  short list[R0_MAXP];
  int n_present;
};
inline size_t row0_smem_bytes(int L) { return sizeof(Row0Smem) + (((size_t)L * 2 + 15) & ~(size_t)15); }

__global__ void __launch_bounds__(256, 3)
k_row_attn_combo(const PfAttnW* __restrict__ W, const PfHeadW* __restrict__ hw, const Row0Tab* __restrict__ tab,
                 float* __restrict__ x, const uint8_t* __restrict__ msa, const int* __restrict__ soft_flag, int n, int L,
                 long long pair_lo, int Pl) {
  if (soft_flag != nullptr && *soft_flag != 0) return;      // soft (non one-hot) input: k_row_attn<1> does this row
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Row0Smem& sm = *reinterpret_cast<Row0Smem*>(smem_raw);
  unsigned short* tslot = reinterpret_cast<unsigned short*>(smem_raw + sizeof(Row0Smem));   // [L]: combo, then its slot
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row = blockIdx.x;
  const int b = row / Pl, pl = row - b * Pl;
  float* xrow = x + (size_t)row * L * PF_D;
  int pi, pj;
  pair_to_ij(pair_lo + pl, n, &pi, &pj);
  const uint8_t* si = msa + ((size_t)b * n + pi) * L;
  const uint8_t* sj = msa + ((size_t)b * n + pj) * L;
  float treg[(PF_NCHAR * PF_D + 255) / 256];      // the embedding table travels through registers: its load latency
#pragma unroll                                    // hides behind the histogram instead of sitting in front of it
  for (int k = 0; k < (PF_NCHAR * PF_D + 255) / 256; ++k) {
    const int t = tid + 256 * k;
    treg[k] = t < PF_NCHAR * PF_D ? __ldg(&hw->table[0][0] + t) : 0.f;
  }
  for (int t = tid; t < PF_NCOMBO; t += 256) sm.cnt[t] = 0;
  __syncthreads();
  // ---- 1. histogram; a combo is the UNORDERED residue pair (x0 = T[a] + T[b] is symmetric), so that the pairs
  //         (i, k) and (k, j) of identical sequences i, j walk the same combos in the same order: bit-identical rows
  for (int l = tid; l < L; l += 256) {
    const int u = min((int)si[l], PF_NCHAR - 1), v = min((int)sj[l], PF_NCHAR - 1);
    const int c = min(u, v) * PF_NCHAR + max(u, v);
    tslot[l] = (unsigned short)c;
    atomicAdd(&sm.cnt[c], 1);
  }
#pragma unroll
  for (int k = 0; k < (PF_NCHAR * PF_D + 255) / 256; ++k) {
    const int t = tid + 256 * k;
    if (t < PF_NCHAR * PF_D) (&sm.r.table[0][0])[t] = treg[k];
  }
  __syncthreads();
  if (tid < 32) {   // present combos, ascending (ballot scan by one warp)
    int base = 0;
    for (int c0 = 0; c0 < PF_NCOMBO; c0 += 32) {
      const int c = c0 + lane;
      const bool on = c < PF_NCOMBO && sm.cnt[c] > 0;
      const unsigned bal = __ballot_sync(PF_FULL, on);
      const int pos = base + __popc(bal & ((1u << lane) - 1u));
      if (on) { sm.list[pos] = (short)c; sm.slot_of[c] = (short)pos; }
      base += __popc(bal);
    }
    if (lane == 0) sm.n_present = base;
  }
  __syncthreads();
  const int np = sm.n_present;
  for (int t = tid; t < np * 8; t += 256) {      // the present combos' k~ / q~ (independent loads, one L2 round trip)
    const int i = t >> 3, v = t & 7, c = sm.list[i];
    const float kq = __ldg(&tab->kq[c][v]);
    sm.pk[i][v] = (float)sm.cnt[c] * kq;
    if (v >= 4) sm.qh[i][v - 4] = kq;
  }
  for (int l = tid; l < L; l += 256) tslot[l] = (unsigned short)sm.slot_of[tslot[l]];
  __syncthreads();
  // ---- 2. row sums over the present combos: thread (g, c) takes the combos i = g (mod 4) for channel c and all
  //         four heads; the four partial sums are combined in a fixed order
  {
    const int g = tid >> 6, c = tid & 63;
    float acc[PF_H] = {0.f, 0.f, 0.f, 0.f};
    int i = g;
    for (; i + 12 < np; i += 16) {
      float nv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) nv[u] = __ldg(&tab->n[sm.list[i + 4 * u]][c]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 w = *reinterpret_cast<const float4*>(sm.pk[i + 4 * u]);
        acc[0] = fmaf(w.x, nv[u], acc[0]); acc[1] = fmaf(w.y, nv[u], acc[1]);
        acc[2] = fmaf(w.z, nv[u], acc[2]); acc[3] = fmaf(w.w, nv[u], acc[3]);
      }
    }
    for (; i < np; i += 4) {
      const float nv = __ldg(&tab->n[sm.list[i]][c]);
      const float4 w = *reinterpret_cast<const float4*>(sm.pk[i]);
      acc[0] = fmaf(w.x, nv, acc[0]); acc[1] = fmaf(w.y, nv, acc[1]);
      acc[2] = fmaf(w.z, nv, acc[2]); acc[3] = fmaf(w.w, nv, acc[3]);
    }
#pragma unroll
    for (int h = 0; h < PF_H; ++h) sm.r.red[g][8 + h * PF_D + c] = acc[h];
    // sum k~ / sum q~: warp v sums value v over the combos (lane-strided, then a fixed xor tree)
    float sv = 0.f;
    for (int k = lane; k < np; k += 32) sv += sm.pk[k][warp];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(PF_FULL, sv, o);
    if (lane == 0) sm.r.tot[warp] = sv;
  }
  __syncthreads();
  {
    const int e = 8 + tid;      // 256 entries of S
    sm.r.tot[e] = ((sm.r.red[0][e] + sm.r.red[1][e]) + sm.r.red[2][e]) + sm.r.red[3][e];
  }
  __syncthreads();
  // ---- 3. M_row, L / sum q~ ----
  row_finalize(W, sm.r, tid, L);
  // ---- 4. Y for R0_YCAP present combos at a time, then gather + store ----
  const int part = tid & 15;                       // this thread's 16-byte piece of a token: channels 4 part ..
  const float4 bo4 = *reinterpret_cast<const float4*>(&W->bo[4 * part]);
  const float4 M0 = *reinterpret_cast<const float4*>(sm.r.M[4 * part + 0]);
  const float4 M1 = *reinterpret_cast<const float4*>(sm.r.M[4 * part + 1]);
  const float4 M2 = *reinterpret_cast<const float4*>(sm.r.M[4 * part + 2]);
  const float4 M3 = *reinterpret_cast<const float4*>(sm.r.M[4 * part + 3]);
  const float4 qi = *reinterpret_cast<const float4*>(sm.r.qinv);
  for (int s0 = 0; s0 < np; s0 += R0_YCAP) {
    const int ns = min(R0_YCAP, np - s0);
    for (int idx = tid; idx < ns * 16; idx += 256) {
      const int sl = idx >> 4, c = sm.list[s0 + sl];
      const int a = c / PF_NCHAR, bb = c - a * PF_NCHAR;
      const float4 ta = *reinterpret_cast<const float4*>(&sm.r.table[a][4 * part]);
      const float4 tb = *reinterpret_cast<const float4*>(&sm.r.table[bb][4 * part]);
      float4 q = *reinterpret_cast<const float4*>(sm.qh[s0 + sl]);
      q.x *= qi.x; q.y *= qi.y; q.z *= qi.z; q.w *= qi.w;
      float4 y;
      y.x = (ta.x + tb.x) + fmaf(M0.w, q.w, fmaf(M0.z, q.z, fmaf(M0.y, q.y, fmaf(M0.x, q.x, bo4.x))));
      y.y = (ta.y + tb.y) + fmaf(M1.w, q.w, fmaf(M1.z, q.z, fmaf(M1.y, q.y, fmaf(M1.x, q.x, bo4.y))));
      y.z = (ta.z + tb.z) + fmaf(M2.w, q.w, fmaf(M2.z, q.z, fmaf(M2.y, q.y, fmaf(M2.x, q.x, bo4.z))));
      y.w = (ta.w + tb.w) + fmaf(M3.w, q.w, fmaf(M3.z, q.z, fmaf(M3.y, q.y, fmaf(M3.x, q.x, bo4.w))));
      *reinterpret_cast<float4*>(&sm.Y[sl][4 * part]) = y;
    }
    __syncthreads();
    // 16 threads per token (one 256-byte line per half-warp), four tokens in flight per thread
    int l = tid >> 4;
    for (; l + 48 < L; l += 64) {
      int sl[4];
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) sl[u] = (int)tslot[l + 16 * u] - s0;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (sl[u] >= 0 && sl[u] < ns) v[u] = *reinterpret_cast<const float4*>(&sm.Y[sl[u]][4 * part]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (sl[u] >= 0 && sl[u] < ns) *reinterpret_cast<float4*>(xrow + (size_t)(l + 16 * u) * PF_D + 4 * part) = v[u];
    }
    for (; l < L; l += 16) {
      const int sl = (int)tslot[l] - s0;
      if (sl >= 0 && sl < ns)
        *reinterpret_cast<float4*>(xrow + (size_t)l * PF_D + 4 * part) = *reinterpret_cast<const float4*>(&sm.Y[sl][4 * part]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Row attention, blocks 1..: same math as k_row_attn<0>, but the pair-row is streamed through a
// shared-memory ring by the TMA engine (cp.async.bulk, 1-D, mbarrier completion): 64-token
// (16 KB) chunks, 4 stages in flight per CTA, issued by one thread.  The copies use no
// registers, so memory-level parallelism no longer depends on how many loads each of the 128-
// register threads can keep in flight.  Both passes over the row (sums, then apply) run through
// the same ring; the second pass is an L2 hit.
// ------------------------------------------------------------------------------------------
#define RT_TOK 64
#define RT_STAGES 4
#define RT_STAGE_BYTES (RT_TOK * PF_D * 4)

__device__ __forceinline__ void rt_issue(uint32_t dst, const float* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// Bounded wait (never hangs the GPU).  On a timeout the device error flag is raised (pf_device_error reports it)
// instead of silently consuming a stage that never arrived.
__device__ __forceinline__ void rt_wait(uint32_t bar, uint32_t parity, int* err_flag) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok && spin < (1u << 22); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2" PF_WAIT_HINT_STR ";\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
  if (!ok && err_flag != nullptr) *err_flag = 7;
}

__global__ void __launch_bounds__(256, 2)
k_row_attn_tma(const PfAttnW* __restrict__ W, float* __restrict__ x, int L, int* __restrict__ err_flag) {
  extern __shared__ __align__(128) unsigned char smem_rt[];
  unsigned char* ring = smem_rt;                                               // RT_STAGES x 16 KB
  RowSmem& sm = *reinterpret_cast<RowSmem*>(smem_rt + RT_STAGES * RT_STAGE_BYTES);
  float* qcache = reinterpret_cast<float*>(smem_rt + RT_STAGES * RT_STAGE_BYTES + sizeof(RowSmem));  // [L][4]
  __shared__ __align__(8) unsigned long long bars[RT_STAGES];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = lane & 7, slot = tid >> 3;
  float* xrow = x + (size_t)blockIdx.x * L * PF_D;
  const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t bar_u32 = (uint32_t)__cvta_generic_to_shared(bars);

  const int n_chunks = (L + RT_TOK - 1) / RT_TOK;
  const bool resident = n_chunks <= RT_STAGES;             // short rows: the whole row stays in the ring
  const int n_total = resident ? n_chunks : 2 * n_chunks;  // otherwise pass B streams the row again (L2 hit)
  auto issue = [&](int c) {          // thread 0 only
    const int cr = c < n_chunks ? c : c - n_chunks;
    const int tok0 = cr * RT_TOK;
    const uint32_t bytes = (uint32_t)(min(RT_TOK, L - tok0) * PF_D * 4);
    const int st = c % RT_STAGES;
    rt_issue(ring_u32 + st * RT_STAGE_BYTES, xrow + (size_t)tok0 * PF_D, bytes, bar_u32 + 8 * st);
  };
  if (tid == 0) {
    for (int s = 0; s < RT_STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u32 + 8 * s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int c = 0; c < RT_STAGES && c < n_total; ++c) issue(c);
  }

  float wqk[8][8];
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    const float4 a = reinterpret_cast<const float4*>(W->wqk[v])[j];
    const float4 c = reinterpret_cast<const float4*>(W->wqk[v])[8 + j];
    wqk[v][0] = a.x; wqk[v][1] = a.y; wqk[v][2] = a.z; wqk[v][3] = a.w;
    wqk[v][4] = c.x; wqk[v][5] = c.y; wqk[v][6] = c.z; wqk[v][7] = c.w;
  }
  const float bias_j = W->bqk[j];
  __syncthreads();  // barriers initialised

  // ---------------- pass A ----------------
  float S[PF_H][8];
#pragma unroll
  for (int h = 0; h < PF_H; ++h)
#pragma unroll
    for (int i = 0; i < 8; ++i) S[h][i] = 0.f;
  float own = 0.f;
  for (int c = 0; c < n_chunks; ++c) {
    const int st = c % RT_STAGES;
    rt_wait(bar_u32 + 8 * st, (uint32_t)((c / RT_STAGES) & 1), err_flag);
    const float* stage = reinterpret_cast<const float*>(ring + st * RT_STAGE_BYTES);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int lt = half * 32 + slot;             // token within the chunk
      const int l = c * RT_TOK + lt;
      const bool act = l < L;
      float xc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) xc[i] = 0.f;
      if (act) load_tok(stage + lt * PF_D, j, xc);
      float nv[8];
      ln_normalize<true>(xc, nv);
      float part[8];
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(wqk[v][i], nv[i], s);
        part[v] = s;
      }
      const float mine = phi_elu1(grp_reduce8(part, j) + bias_j);
      float kh[PF_H];
#pragma unroll
      for (int h = 0; h < PF_H; ++h) kh[h] = __shfl_sync(PF_FULL, mine, (lane & 24) | h);
      if (act) {
        own += mine;
        if (j >= 4) qcache[(size_t)l * 4 + (j - 4)] = mine;
#pragma unroll
        for (int h = 0; h < PF_H; ++h)
#pragma unroll
          for (int i = 0; i < 8; ++i) S[h][i] = fmaf(kh[h], nv[i], S[h][i]);
      }
    }
    __syncthreads();  // everyone is done with this stage
    if (tid == 0 && c + RT_STAGES < n_total) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(c + RT_STAGES);
    }
  }
  // reduce the 4 slots of a warp (fixed tree), then the 8 warps (fixed order)
#pragma unroll
  for (int h = 0; h < PF_H; ++h)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = S[h][i];
      v += __shfl_xor_sync(PF_FULL, v, 8);
      v += __shfl_xor_sync(PF_FULL, v, 16);
      S[h][i] = v;
    }
  own += __shfl_xor_sync(PF_FULL, own, 8);
  own += __shfl_xor_sync(PF_FULL, own, 16);
  if (lane < 8) {
    sm.red[warp][j] = own;
#pragma unroll
    for (int h = 0; h < PF_H; ++h)
#pragma unroll
      for (int i = 0; i < 8; ++i) sm.red[warp][8 + h * PF_D + chan_of(j, i)] = S[h][i];
  }
  __syncthreads();
  for (int t = tid; t < PF_PART; t += 256) {
    float s = sm.red[0][t];
#pragma unroll
    for (int w = 1; w < 8; ++w) s += sm.red[w][t];
    sm.tot[t] = s;
  }
  __syncthreads();
  row_finalize(W, sm, tid, L);
  // ---------------- pass B: y = x + M qhat + bo ----------------
  float4 Mr[8];
  float bo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = chan_of(j, i);
    Mr[i] = *reinterpret_cast<const float4*>(sm.M[c]);
    bo[i] = W->bo[c];
  }
  const float4 qi = *reinterpret_cast<const float4*>(sm.qinv);
  for (int c = n_chunks; c < 2 * n_chunks; ++c) {
    const int st = resident ? (c - n_chunks) : (c % RT_STAGES);
    if (!resident) rt_wait(bar_u32 + 8 * st, (uint32_t)((c / RT_STAGES) & 1), err_flag);
    const float* stage = reinterpret_cast<const float*>(ring + st * RT_STAGE_BYTES);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int lt = half * 32 + slot;
      const int l = (c - n_chunks) * RT_TOK + lt;
      if (l < L) {
        float xc[8];
        load_tok(stage + lt * PF_D, j, xc);
        float4 q = *reinterpret_cast<const float4*>(qcache + (size_t)l * 4);
        q.x *= qi.x; q.y *= qi.y; q.z *= qi.z; q.w *= qi.w;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a = bo[i];
          a = fmaf(Mr[i].x, q.x, a);
          a = fmaf(Mr[i].y, q.y, a);
          a = fmaf(Mr[i].z, q.z, a);
          a = fmaf(Mr[i].w, q.w, a);
          xc[i] += a;
        }
        store_tok(xrow + (size_t)l * PF_D, j, xc);
      }
    }
    if (!resident) {
      __syncthreads();
      if (tid == 0 && c + RT_STAGES < n_total) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(c + RT_STAGES);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Column attention, step 1: partial sums over a chunk of pairs at each site.
//   part[chunk][b][l][0:4]  = sum_p k~_h      [4:8] = sum_p q~_h      [8+64h+c] = sum_p k~_h n_c
// (n = LN(x) without affine; k~, q~ un-normalised phi values, attention.py:179-180 with
//  R=L, N=P).  A token slot owns one site and walks down the pairs, so there is no cross-lane
//  reduction and the order of the sum is fixed.   grid = (ceil(L/32), n_chunks, B)
// ------------------------------------------------------------------------------------------
#ifndef PF_COL_MINB
#define PF_COL_MINB 2
#endif
__global__ void __launch_bounds__(256, PF_COL_MINB)
k_col_partial(const PfAttnW* __restrict__ W, const float* __restrict__ x, float* __restrict__ part,
              int L, int Pl, int pairs_per_chunk) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int j = lane & 7, slot = tid >> 3;
  const int l = blockIdx.x * 32 + slot;
  const int chunk = blockIdx.y, b = blockIdx.z, B = gridDim.z;
  const bool act = l < L;
  const int p0 = chunk * pairs_per_chunk;
  const int p1 = min(Pl, p0 + pairs_per_chunk);

  float wqk[8][8];
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    const float4 a = reinterpret_cast<const float4*>(W->wqk[v])[j];
    const float4 c = reinterpret_cast<const float4*>(W->wqk[v])[8 + j];
    wqk[v][0] = a.x; wqk[v][1] = a.y; wqk[v][2] = a.z; wqk[v][3] = a.w;
    wqk[v][4] = c.x; wqk[v][5] = c.y; wqk[v][6] = c.z; wqk[v][7] = c.w;
  }
  const float bias_j = W->bqk[j];
  float S[PF_H][8];
#pragma unroll
  for (int h = 0; h < PF_H; ++h)
#pragma unroll
    for (int i = 0; i < 8; ++i) S[h][i] = 0.f;
  float own = 0.f;

  const size_t pstride = (size_t)L * PF_D;
  const float* px = x + ((size_t)b * Pl + p0) * pstride + (size_t)(act ? l : 0) * PF_D;
  float xn[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) xn[i] = 0.f;
  if (act && p0 < p1) load_tok(px, j, xn);
  for (int p = p0; p < p1; ++p) {
    float xc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xc[i] = xn[i];
    if (act && p + 1 < p1) load_tok(px + pstride, j, xn);
    px += pstride;
    float nv[8];
    ln_normalize(xc, nv);
    float pr[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(wqk[v][i], nv[i], s);
      pr[v] = s;
    }
    const float mine = phi_elu1(grp_reduce8(pr, j) + bias_j);
    float kh[PF_H];
#pragma unroll
    for (int h = 0; h < PF_H; ++h) kh[h] = __shfl_sync(PF_FULL, mine, (lane & 24) | h);
    own += mine;
#pragma unroll
    for (int h = 0; h < PF_H; ++h)
#pragma unroll
      for (int i = 0; i < 8; ++i) S[h][i] = fmaf(kh[h], nv[i], S[h][i]);
  }
  if (act) {
    float* o = part + (((size_t)chunk * B + b) * L + l) * PF_PART;
    o[j] = own;
#pragma unroll
    for (int h = 0; h < PF_H; ++h) {
      reinterpret_cast<float4*>(o + 8 + h * PF_D)[j] = make_float4(S[h][0], S[h][1], S[h][2], S[h][3]);
      reinterpret_cast<float4*>(o + 8 + h * PF_D)[8 + j] = make_float4(S[h][4], S[h][5], S[h][6], S[h][7]);
    }
  }
}

// Column attention, step 2: fixed-order sum over the chunks, then the linear map to the
// 72-float exchange form:  sum_p k~ v = Wv[h] (g * sum_p k~ n + b sum_p k~) + bv[h] sum_p k~.
//   grid = ceil(B L / spc) CTAs of 256 threads, spc <= PF_FS sites per CTA (host: about two CTAs per SM): a thread keeps its column of
//   W_v in registers and reuses it for its sites (one 16 KB weight read per CTA instead of per
//   site; batches of small alignments have tens of thousands of sites).  Summation orders are
//   fixed and independent of the site's position in the CTA.
__device__ __forceinline__ void col_reduce_sites(const PfAttnW* __restrict__ W, const float* __restrict__ part, int n_chunks,
                                                 int n_sites, int s0, int ns, float* __restrict__ colsum,
                                                 float (*tot)[PF_PART], float (*ub)[PF_H][PF_D]) {
  const int t = threadIdx.x;
  for (int i = t; i < ns * PF_PART; i += 256) {
    const int sl = i / PF_PART, e = i - sl * PF_PART;
    // four interleaved chains (chunks ch = k mod 4), combined in a fixed order: the loads of a chain of n_chunks
    // dependent adds were the latency of this kernel
    const float* src = part + ((size_t)s0 + sl) * PF_PART + e;
    const size_t cs = (size_t)n_sites * PF_PART;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int ch = 0;
    for (; ch + 3 < n_chunks; ch += 4) {
      const float v0 = src[(size_t)ch * cs], v1 = src[(size_t)(ch + 1) * cs], v2 = src[(size_t)(ch + 2) * cs], v3 = src[(size_t)(ch + 3) * cs];
      a0 += v0; a1 += v1; a2 += v2; a3 += v3;
    }
    if (ch < n_chunks) a0 += src[(size_t)ch * cs];
    if (ch + 1 < n_chunks) a1 += src[(size_t)(ch + 1) * cs];
    if (ch + 2 < n_chunks) a2 += src[(size_t)(ch + 2) * cs];
    tot[sl][e] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  for (int i = t; i < ns * PF_H * PF_D; i += 256) {
    const int sl = i >> 8, h = (i >> 6) & 3, c = i & 63;
    ub[sl][h][c] = fmaf(W->gamma[c], tot[sl][8 + h * PF_D + c], W->beta[c] * tot[sl][h]);
  }
  for (int i = t; i < ns * 8; i += 256) colsum[(size_t)(s0 + (i >> 3)) * PF_COLSUM + (i & 7)] = tot[i >> 3][i & 7];
  __syncthreads();
  const int o = t & 63, h = o >> 4;
  float wv[PF_D];
#pragma unroll
  for (int k = 0; k < PF_D; ++k) wv[k] = W->wvT[k][o];
  const float bvo = W->bv[o];
  for (int sl = t >> 6; sl < ns; sl += 4) {
    float acc = bvo * tot[sl][h];
#pragma unroll
    for (int k = 0; k < PF_D; ++k) acc = fmaf(wv[k], ub[sl][h][k], acc);
    colsum[(size_t)(s0 + sl) * PF_COLSUM + 8 + o] = acc;
  }
}

__global__ void __launch_bounds__(256)
k_col_reduce(const PfAttnW* __restrict__ W, const float* __restrict__ part, int n_chunks, int n_sites, int spc,
             float* __restrict__ colsum) {
  __shared__ float tot[PF_FS][PF_PART];
  __shared__ float ub[PF_FS][PF_H][PF_D];
  const int s0 = blockIdx.x * spc;   // spc <= PF_FS sites per CTA
  col_reduce_sites(W, part, n_chunks, n_sites, s0, min(spc, n_sites - s0), colsum, tot, ub);
}

// Column attention, step 3 (after the cross-shard sum): ctx = kv / sum k,  M_l = Wo[:,h] ctx_h,
// qinv = P / sum q~.     grid = ceil(B L / PF_FS), 256 threads; thread (c, h) keeps its 16 W_o
// weights in registers for the CTA's PF_FS sites.
__device__ __forceinline__ void col_finalize_sites(const PfAttnW* __restrict__ W, const float (*tot)[PF_COLSUM], int ns,
                                                   int s0, float p_total, float* __restrict__ colM,
                                                   float (*ctx)[PF_D]) {
  const int t = threadIdx.x;
  for (int i = t; i < ns * PF_D; i += 256) {
    const int sl = i >> 6, e = i & 63;
    ctx[sl][e] = tot[sl][8 + e] / tot[sl][e >> 4];
  }
  for (int i = t; i < ns * PF_H; i += 256)
    colM[(size_t)(s0 + (i >> 2)) * PF_MROW + 256 + (i & 3)] = p_total / tot[i >> 2][4 + (i & 3)];
  __syncthreads();
  const int c = t >> 2, h = t & 3;
  float wo[PF_DH];
#pragma unroll
  for (int e = 0; e < PF_DH; ++e) wo[e] = W->wo[c][h * PF_DH + e];
  for (int sl = 0; sl < ns; ++sl) {
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < PF_DH; ++e) acc = fmaf(wo[e], ctx[sl][h * PF_DH + e], acc);
    colM[(size_t)(s0 + sl) * PF_MROW + c * 4 + h] = acc;
  }
}

__global__ void __launch_bounds__(256)
k_col_finalize(const PfAttnW* __restrict__ W, const float* __restrict__ colsum, float p_total,
               int n_sites, int spc, float* __restrict__ colM) {
  __shared__ float tot[PF_FS][PF_COLSUM];
  __shared__ float ctx[PF_FS][PF_D];
  const int t = threadIdx.x, s0 = blockIdx.x * spc;
  const int ns = min(spc, n_sites - s0);
  for (int i = t; i < ns * PF_COLSUM; i += 256) (&tot[0][0])[i] = colsum[(size_t)s0 * PF_COLSUM + i];
  __syncthreads();
  col_finalize_sites(W, tot, ns, s0, p_total, colM, ctx);
}

// ------------------------------------------------------------------------------------------
// Column apply + FFN, fp32 FFMA ("exact" mode and the check for the tcgen05 kernel):
//   x2 = x1 + M_l qhat + bo ;  x3 = x2 + W2 gelu(W1 LN(x2) + b1) + b2        model.py:97-104
// Persistent: one CTA per SM, 32-token tiles, both weight matrices resident in smem.
// ------------------------------------------------------------------------------------------
#define FFN32_T 32
#define FFN32_HS (PF_HID + 4)
struct Ffn32Smem {
  float w1T[PF_D][PF_HID];
  float w2T[PF_HID][PF_D];
  float A[FFN32_T][PF_D];      // LN(x2) without affine; reused for the GEMM2 output
  float Hid[FFN32_T][FFN32_HS];
  float b1[PF_HID];
  float b2[PF_D];
};

__global__ void __launch_bounds__(256, 1)
k_colapply_ffn_fp32(const PfAttnW* __restrict__ Wc, const PfFfnW* __restrict__ Wf,
                    float* __restrict__ x, const float* __restrict__ colM, int L, int Pl,
                    long long n_tok, int apply_only) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ffn32Smem& sm = *reinterpret_cast<Ffn32Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int j = lane & 7, slot = tid >> 3;

  for (int t = tid; t < PF_D * PF_HID / 4; t += 256) {
    reinterpret_cast<float4*>(&sm.w1T[0][0])[t] = reinterpret_cast<const float4*>(&Wf->w1T[0][0])[t];
    reinterpret_cast<float4*>(&sm.w2T[0][0])[t] = reinterpret_cast<const float4*>(&Wf->w2T[0][0])[t];
  }
  sm.b1[tid] = Wf->b1[tid];
  if (tid < PF_D) sm.b2[tid] = Wf->b2[tid];

  float wq[4][8];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const float4 a = reinterpret_cast<const float4*>(Wc->wqk[4 + v])[j];
    const float4 c = reinterpret_cast<const float4*>(Wc->wqk[4 + v])[8 + j];
    wq[v][0] = a.x; wq[v][1] = a.y; wq[v][2] = a.z; wq[v][3] = a.w;
    wq[v][4] = c.x; wq[v][5] = c.y; wq[v][6] = c.z; wq[v][7] = c.w;
  }
  const float bq = Wc->bqk[4 + (j >> 1)];
  float bo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bo[i] = Wc->bo[chan_of(j, i)];
  __syncthreads();

  const long long n_tiles = (n_tok + FFN32_T - 1) / FFN32_T;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- step 1: column apply, LN (one token per slot) ----
    const long long tok = tile * FFN32_T + slot;
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
    {
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
        const float4 m = *reinterpret_cast<const float4*>(cm + chan_of(j, i) * 4);
        float a = bo[i];
        a = fmaf(m.x, qh[0], a);
        a = fmaf(m.y, qh[1], a);
        a = fmaf(m.z, qh[2], a);
        a = fmaf(m.w, qh[3], a);
        x2[i] += a;
      }
      if (apply_only) {  // test hook: x2 only (uniform branch)
        if (act) store_tok(x + (size_t)tok * PF_D, j, x2);
        continue;
      }
      ln_normalize(x2, nv);
      reinterpret_cast<float4*>(sm.A[slot])[j] = make_float4(nv[0], nv[1], nv[2], nv[3]);
      reinterpret_cast<float4*>(sm.A[slot])[8 + j] = make_float4(nv[4], nv[5], nv[6], nv[7]);
    }
    __syncthreads();
    // ---- step 2: Hid = gelu(A w1T + b1): thread = 4 tokens x 8 hidden units ----
    {
      const int ty = tid >> 5, tx = tid & 31;
      float acc[4][8];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[r][i] = sm.b1[tx + 32 * i];
#pragma unroll 2
      for (int k = 0; k < PF_D; k += 4) {
        float4 a[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(&sm.A[4 * ty + r][k]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = sm.w1T[k + kk][tx + 32 * i];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[r][i] = fmaf(av, w[i], acc[r][i]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) sm.Hid[4 * ty + r][tx + 32 * i] = gelu_erf(acc[r][i]);
    }
    __syncthreads();
    // ---- step 3: O = Hid w2T + b2: thread = 2 tokens x 4 channels; O overwrites A ----
    {
      const int ty = tid >> 4, tx = tid & 15;
      float acc[2][4];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[r][i] = sm.b2[tx + 16 * i];
#pragma unroll 2
      for (int k = 0; k < PF_HID; k += 4) {
        float4 a[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) a[r] = *reinterpret_cast<const float4*>(&sm.Hid[2 * ty + r][k]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) w[i] = sm.w2T[k + kk][tx + 16 * i];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[r][i] = fmaf(av, w[i], acc[r][i]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) sm.A[2 * ty + r][tx + 16 * i] = acc[r][i];
    }
    __syncthreads();
    // ---- step 4: residual, store ----
    if (act) {
      const float4 o0 = reinterpret_cast<const float4*>(sm.A[slot])[j];
      const float4 o1 = reinterpret_cast<const float4*>(sm.A[slot])[8 + j];
      x2[0] += o0.x; x2[1] += o0.y; x2[2] += o0.z; x2[3] += o0.w;
      x2[4] += o1.x; x2[5] += o1.y; x2[6] += o1.z; x2[7] += o1.w;
      store_tok(x + (size_t)tok * PF_D, j, x2);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Distance head: dist[b,p] = mean_l softplus(w . x[b,p,l,:] + c)        model.py:158-164,182-185
// One CTA per pair-row; fixed-order reduction over the sites.
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// Pair-sharded column attention over NVLink peer memory (no NCCL on the data path).
// Every rank owns a symmetric exchange buffer, mapped into all peers:
//     [ flags: PF_PEER_MAX_WORLD x PF_PEER_MAX_CTAS x uint32 ][ slot 0: B*L*72 floats ][ slot 1: B*L*72 floats ]
// k_col_reduce writes this rank's 72-float summaries into slot (epoch & 1) of its own buffer;
// k_peer_sync publishes "epoch reached" into every peer's flag word [rank] (system-scope
// release after a system fence) and waits until all peers have published the same epoch
// (acquire); k_col_finalize_peer then reads the same slot of every rank with plain P2P loads
// and sums in rank order (identical on every rank, deterministic).  Two slots suffice: a slot is
// rewritten two exchanges later, and a rank only gets past the next exchange's sync after every
// peer has finished reading the previous one.
// ------------------------------------------------------------------------------------------
#define PF_PEER_MAX_CTAS 512                               // CTAs of k_col_exchange (one flag word per rank and CTA)
#define PF_PEER_MAX_WORLD 32
#define PF_PEER_FLAG_BYTES (PF_PEER_MAX_WORLD * PF_PEER_MAX_CTAS * 4)   // 64 KB

__global__ void k_peer_sync(unsigned char* const* __restrict__ peers, int rank, int world, unsigned epoch,
                            int* __restrict__ err_flag) {
  const int t = threadIdx.x;
  if (t >= world) return;
  __threadfence_system();  // this rank's summaries (written by the previous kernel) before the flag
  // three-launch form (PF_EXCH_IMPL=split): one flag per rank, kept in the last CTA column of the flag table
  volatile unsigned* remote = reinterpret_cast<volatile unsigned*>(peers[t]) + (size_t)rank * PF_PEER_MAX_CTAS + (PF_PEER_MAX_CTAS - 1);
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  const unsigned* mine = reinterpret_cast<const unsigned*>(peers[rank]) + (size_t)t * PF_PEER_MAX_CTAS + (PF_PEER_MAX_CTAS - 1);
  unsigned v = 0;
  for (unsigned spin = 0; spin < (1u << 26); ++spin) {   // bounded (~seconds): never hang the GPU
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int)(v - epoch) >= 0) return;
    __nanosleep(64);
  }
  if (err_flag != nullptr) *err_flag = 3;
}

__global__ void __launch_bounds__(256)
k_col_finalize_peer(const PfAttnW* __restrict__ W, unsigned char* const* __restrict__ peers, int world, int slot,
                    size_t slot_floats, float p_total, int n_sites, int spc, float* __restrict__ colM) {
  __shared__ float tot[PF_FS][PF_COLSUM];
  __shared__ float ctx[PF_FS][PF_D];
  const int t = threadIdx.x, s0 = blockIdx.x * spc;
  const int ns = min(spc, n_sites - s0);
  for (int i = t; i < ns * PF_COLSUM; i += 256) {
    float a = 0.f;
    for (int r = 0; r < world; ++r) {  // fixed rank order
      const float* src = reinterpret_cast<const float*>(peers[r] + PF_PEER_FLAG_BYTES) + (size_t)slot * slot_floats;
      a += src[(size_t)s0 * PF_COLSUM + i];
    }
    (&tot[0][0])[i] = a;
  }
  __syncthreads();
  col_finalize_sites(W, tot, ns, s0, p_total, colM, ctx);
}

// ------------------------------------------------------------------------------------------
// k_col_exchange: chunk reduce + cross-rank exchange + finalize of the column summaries in ONE launch per block
// (replaces k_col_reduce -> k_peer_sync -> k_col_finalize_peer, and k_col_reduce -> k_col_finalize on one GPU).
// The exchange is site-chunked: CTA c owns the site groups c, c + grid, ... (spc sites each; the same grid on every
// rank), and the flag that guards a peer's data is per (rank, CTA), not per rank:
//   phase 1  reduce the pair-chunk partials of my site groups into the 72-float exchange form, in my own slot of
//            the symmetric buffer (plain stores to local HBM)
//   publish  CTA barrier, system fence, st.release.sys of the epoch into flag [rank][c] of every peer
//   phase 2  ld.acquire.sys until flag [r][c] of every rank r shows the epoch, then read the same site groups out of
//            every rank's slot over NVLink (L1-bypassing loads), sum in rank order (bit-identical on all ranks),
//            finalize M_l / qinv
// No CTA waits before it has published everything it owns, so the kernel cannot deadlock whatever order the CTAs
// are scheduled in, and the NVLink latency of the first groups' flags and data hides behind the reduction of the
// other groups and CTAs instead of behind a device-wide kernel boundary (the three-launch form's two boundaries
// and its single-warp sync kernel were 5 % of a step at 8 ranks).  Slot reuse: as above, two slots.
// ------------------------------------------------------------------------------------------
struct ColExSmem {
  union {
    struct { float tot[PF_FS][PF_PART]; float ub[PF_FS][PF_H][PF_D]; } a;      // phase 1
    struct { float tot[PF_FS][PF_COLSUM]; float ctx[PF_FS][PF_D]; } b;         // phase 2
  };
};

__global__ void __launch_bounds__(256)
k_col_exchange(const PfAttnW* __restrict__ W, const float* __restrict__ part, int n_chunks, int n_sites, int spc,
               float* __restrict__ my_slot, unsigned char* const* __restrict__ peers, int rank, int world, int slot,
               size_t slot_floats, unsigned epoch, float p_total, float* __restrict__ colM, int* __restrict__ err_flag) {
  __shared__ ColExSmem sm;
  const int t = threadIdx.x;
  const int n_groups = (n_sites + spc - 1) / spc;
  for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int s0 = g * spc;
    col_reduce_sites(W, part, n_chunks, n_sites, s0, min(spc, n_sites - s0), my_slot, sm.a.tot, sm.a.ub);
    __syncthreads();
  }
  if (world > 1) {
    __syncthreads();
    if (t < world) {
      __threadfence_system();   // the CTA's summaries (ordered before this thread by the barrier) before the flag
      unsigned* remote = reinterpret_cast<unsigned*>(peers[t]) + (size_t)rank * PF_PEER_MAX_CTAS + blockIdx.x;
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
      const unsigned* mine = reinterpret_cast<const unsigned*>(peers[rank]) + (size_t)t * PF_PEER_MAX_CTAS + blockIdx.x;
      unsigned v = 0;
      bool ok = false;
      for (unsigned spin = 0; spin < (1u << 26); ++spin) {   // bounded (~seconds): never hang the GPU
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int)(v - epoch) >= 0) { ok = true; break; }
        __nanosleep(32);
      }
      if (!ok && err_flag != nullptr) *err_flag = 3;
    }
    __syncthreads();
  }
  for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int s0 = g * spc, ns = min(spc, n_sites - s0);
    for (int i = t; i < ns * PF_COLSUM; i += 256) {
      float a;
      if (world > 1) {
        a = 0.f;
        for (int r = 0; r < world; ++r) {  // fixed rank order
          const float* src = reinterpret_cast<const float*>(peers[r] + PF_PEER_FLAG_BYTES) + (size_t)slot * slot_floats;
          a += __ldcg(src + (size_t)s0 * PF_COLSUM + i);
        }
      } else {
        a = my_slot[(size_t)s0 * PF_COLSUM + i];
      }
      (&sm.b.tot[0][0])[i] = a;
    }
    __syncthreads();
    col_finalize_sites(W, sm.b.tot, ns, s0, p_total, colM, sm.b.ctx);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
k_head(const PfHeadW* __restrict__ hw, const float* __restrict__ x, int L, float* __restrict__ dist) {
  __shared__ float red[32];
  const int tid = threadIdx.x, lane = tid & 31;
  const int j = lane & 7, slot = tid >> 3;
  const float* xrow = x + (size_t)blockIdx.x * L * PF_D;
  float w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = hw->whead[chan_of(j, i)];
  const float c = hw->bhead;
  float acc = 0.f;
  const int n_it = (L + 31) >> 5;
  for (int it = 0; it < n_it; ++it) {
    const int l = it * 32 + slot;
    float xc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xc[i] = 0.f;
    if (l < L) load_tok(xrow + (size_t)l * PF_D, j, xc);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(w[i], xc[i], s);
    s = grp_sum(s) + c;
    if (l < L) acc += softplus20(s);
  }
  if (j == 0) red[slot] = acc;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += red[i];
    dist[blockIdx.x] = s / (float)L;
  }
}

// Fused head, step 2: dist[b,p] = (sum over the 4-site windows of the partial softplus sums) / L, fixed order.
// headpart[b][w][p] is written by the HEAD instantiation of k_colapply_ffn_ws.
__global__ void k_head_reduce(const float* __restrict__ headpart, int nW, int Pl, int L, float* __restrict__ dist) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (p >= Pl) return;
  const float* src = headpart + (size_t)b * nW * Pl + p;
  float s = 0.f;
  for (int w = 0; w < nW; ++w) s += src[(size_t)w * Pl];
  dist[(size_t)b * Pl + p] = s / (float)L;
}

// (B,P) upper-triangle vectors -> (B,n,n) symmetric matrices        infer_alns.py:14-25
__global__ void k_dist_to_matrix(const float* __restrict__ dist, int n, long long P,
                                 float* __restrict__ mat) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= (long long)n * n) return;
  const int r = (int)(t / n), c = (int)(t % n);
  float v = 0.f;
  if (r != c) {
    const long long i = r < c ? r : c, jj = r < c ? c : r;
    v = dist[(size_t)b * P + (i * n - i * (i + 1) / 2 + (jj - i - 1))];
  }
  mat[(size_t)b * n * n + t] = v;
}
