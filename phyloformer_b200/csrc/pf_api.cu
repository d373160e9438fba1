// pf_api.cu -- the C ABI of libpf_sm100.so (see include/pf_sm100.h) and the host-side
// orchestration of one Phyloformer forward: which kernels run, in which order, on which
// buffers.  No torch, no NCCL: device memory and the cross-shard sum come from the caller.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pf_sm100.h"
#include "pf_common.cuh"
#include "pf_kernels.cuh"
#include "pf_ffn_tc.cuh"
#include "pf_ffn_ws.cuh"
#include "pf_attn_tc.cuh"
#include "pf_bme.h"

namespace {

thread_local char g_err[512] = "";

// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
typedef CUresult (*pf_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
pf_encode_tiled_fn get_encode_tiled() {
  static pf_encode_tiled_fn fn = []() -> pf_encode_tiled_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (pf_encode_tiled_fn)p;
  }();
  return fn;
}
// Tensor map of the activation x[(B Pl) pairs][L sites][64] fp32: box [32 pairs][4 sites][32 channels],
// SWIZZLE_128B (the 16-byte chunks of every 128-byte row XOR-ed with the row index mod 8).
// 2-D view of the same buffer, x[tokens][64]: box [128 tokens][32 channels], SWIZZLE_128B (row kernel).
bool make_x_tensor_map_2d(CUtensorMap* tm, const float* x, long long n_tok) {
  pf_encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return false;
  const cuuint64_t dims[2] = {PF_D, (cuuint64_t)n_tok};
  const cuuint64_t strides[1] = {PF_D * sizeof(float)};
  const cuuint32_t box[2] = {32, 128}, estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
bool make_x_tensor_map(CUtensorMap* tm, const float* x, long long rows, int L) {
  pf_encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return false;
  const cuuint64_t dims[3] = {PF_D, (cuuint64_t)L, (cuuint64_t)rows};
  const cuuint64_t strides[2] = {PF_D * sizeof(float), (cuuint64_t)L * PF_D * sizeof(float)};
  const cuuint32_t box[3] = {32, 4, 32}, estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(PF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                  __LINE__);                                                                  \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Plan {  // workspace carve-up for one (B, n, L, pair range)
  long long Pl, P;
  int n_chunks, ppc;        // k_col_partial (FFMA): pair chunks
  int n_chunks_tc, ppc_tc;  // k_col_partial_tc (tcgen05): pair chunks, ppc_tc a multiple of 32
  int n_chunks_ws, ppc_ws;  // k_col_partial_ws (tcgen05, warp specialised, one CTA per SM)
  size_t off_x, off_part, off_colsum, off_colM, off_semb, off_qc, off_head, total;
};

}  // namespace

struct pf_ctx {
  pf_cfg cfg;
  int n_sm = 0;
  PfHeadW* head_dev = nullptr;
  PfBlockW* blk_dev = nullptr;  // [nb]
  PfFfnTcW* tc_dev = nullptr;   // [nb] bf16 hi/lo smem images for the tcgen05 FFN
  PfFfnTcW* tc16_dev = nullptr; // [nb] the same in fp16 (PF_PREC_FP16)
  PfFfnTcW* tcf_dev = nullptr;  // [nb] bf16 hi/lo images with the first bias folded into W1 (WS_FOLD63: k_colapply_ffn_ws, parity format)
  PfAttnTcW* atc_dev = nullptr; // [nb][2] q/k weight images for the tcgen05 attention kernels (0: row, 1: column)
  int col_impl = 2;             // 0: k_col_partial (FFMA), 1: k_col_partial_tc (tcgen05, one thread per token),
                                // 2: k_col_partial_ws (tcgen05, warp specialised); env PF_COL_IMPL=cc|tc1|tc
  std::vector<PfFfnConst> ffn_const;  // [nb] host copies passed as __grid_constant__ kernel parameters
  int launches = 0;
  int ffn_impl = 1;             // 0: pf_ffn_tc.cuh (phased), 1: pf_ffn_ws.cuh (warp-specialised); env PF_FFN_IMPL=tc|ws
  int row_impl = 2;             // 0: k_row_attn<0> (register loads), 1: k_row_attn_tma (bulk-copy ring), 2: k_row_attn_ws
                                // (tcgen05, warp specialised; tensor-core modes, blocks 1.., L <= 1536); env PF_ROW_IMPL=ld|tma|tc
  // peer-memory exchange (pf_set_peer_exchange): symmetric buffers of all ranks, mapped locally
  int peer_rank = 0, peer_world = 0;
  unsigned char** peers_dev = nullptr;   // [world] device array of buffer base pointers
  unsigned char* peer_self = nullptr;    // this rank's buffer
  size_t peer_slot_floats = 0;
  unsigned peer_epoch = 0;
  int ws_prof = 0;              // env PF_WS_PROF=1: role timing into the dump buffer (test hook)
  Row0Tab* row0_dev = nullptr;  // residue-pair tables of block 0's row attention (k_row_attn_combo)
  int row0_impl = 1;            // 1: k_row_attn_combo for residue-code inputs, 0: k_row_attn<1> always; env PF_ROW0_IMPL=combo|ffma
  int exch_impl = 1;            // 1: k_col_exchange (one fused, site-chunked launch per block), 0: reduce / sync / finalize launches; env PF_EXCH_IMPL=fused|split
  int head_impl = 1;            // 1: distance head fused into the last FFN launch (+ k_head_reduce), 0: k_head; env PF_HEAD_IMPL=fused|sep
  int* err_dev = nullptr;       // set by a kernel whose mbarrier wait timed out
  float* dump_dev = nullptr;    // test hook: raw accumulators of the first FFN tile
  // optional per-kernel timing (pf_profile_*)
  bool prof = false;
  struct Rec { int kc; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get_event() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
};

namespace {
// RAII bracket around one kernel launch
struct Timed {
  pf_ctx* h; cudaStream_t st; int idx = -1;
  Timed(pf_ctx* h_, int kc, cudaStream_t st_) : h(h_), st(st_) {
    ++h->launches;
    if (h->prof) {
      pf_ctx::Rec r{kc, h->get_event(), h->get_event()};
      cudaEventRecord(r.a, st);
      h->recs.push_back(r);
      idx = (int)h->recs.size() - 1;
    }
  }
  ~Timed() { if (idx >= 0) cudaEventRecord(h->recs[idx].b, st); }
};
}  // namespace

namespace {

Plan make_plan(const pf_ctx* h, int B, int n, int L, long long lo, long long hi) {
  Plan p;
  p.P = (long long)n * (n - 1) / 2;
  p.Pl = hi - lo;
  // Column-partial grid: site_tiles x n_chunks x B CTAs, 2 resident per SM.  Pick the chunk count
  // that minimises (waves x pairs per chunk), i.e. avoids a nearly empty last wave.
  const int site_tiles = (L + 31) / 32;
  const long long slots = (long long)h->n_sm * PF_COL_MINB;
  const long long per_chunk = (long long)site_tiles * B;
  long long best_nc = 1, best_cost = -1;
  const long long nc_max = p.Pl < 64 ? (p.Pl > 0 ? p.Pl : 1) : 64;
  for (long long nc = 1; nc <= nc_max; ++nc) {
    const long long ppc = (p.Pl + nc - 1) / nc;
    const long long chunks = (p.Pl + ppc - 1) / ppc;
    const long long waves = (per_chunk * chunks + slots - 1) / slots;
    // per-CTA start-up (weights into registers) ~ 16 pairs' worth; small penalty per partial buffer
    const long long cost = waves * (ppc + 16) * 64 + chunks;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_nc = nc; }
  }
  p.ppc = (int)((p.Pl + best_nc - 1) / best_nc);
  if (p.ppc < 1) p.ppc = 1;
  p.n_chunks = (int)((p.Pl + p.ppc - 1) / p.ppc);
  if (p.n_chunks < 1) p.n_chunks = 1;
  // tensor-core variants: units = (msa, 4-site window, chunk) walked by `slots` persistent CTAs;
  // cost = rounds x (tiles per unit + per-unit epilogue, about 3 tiles' worth)
  auto pick = [&](long long slots, int* ppc_out, int* nc_out) {
    const long long win = (long long)B * ((L + 3) / 4);
    long long bppc = 32, bcost = -1;
    for (long long nc = 1; nc <= 64; ++nc) {
      long long ppc = (p.Pl + nc - 1) / nc;
      ppc = (ppc + 31) / 32 * 32;
      if (ppc < 32) ppc = 32;
      const long long chunks = (p.Pl + ppc - 1) / ppc;
      const long long rounds = (win * chunks + slots - 1) / slots;
      const long long cost = rounds * (ppc / 32 + 3) * 64 + chunks;
      if (bcost < 0 || cost < bcost) { bcost = cost; bppc = ppc; }
    }
    *ppc_out = (int)bppc;
    *nc_out = (int)((p.Pl + bppc - 1) / bppc);
    if (*nc_out < 1) *nc_out = 1;
  };
  pick(2LL * h->n_sm, &p.ppc_tc, &p.n_chunks_tc);
  pick((long long)h->n_sm, &p.ppc_ws, &p.n_chunks_ws);
  const int max_chunks = std::max(p.n_chunks, std::max(p.n_chunks_tc, p.n_chunks_ws));
  size_t off = 0;
  p.off_x = off;       off = align_up(off + (size_t)B * p.Pl * L * PF_D * sizeof(float), 256);
  p.off_part = off;    off = align_up(off + (size_t)max_chunks * B * L * PF_PART * sizeof(float), 256);
  p.off_colsum = off;  off = align_up(off + (size_t)B * L * PF_COLSUM * sizeof(float), 256);
  p.off_colM = off;    off = align_up(off + (size_t)B * L * PF_MROW * sizeof(float), 256);
  p.off_semb = off;    off = align_up(off + (size_t)B * n * L * PF_D * sizeof(float), 256);
  p.off_qc = off;      off = align_up(off + (size_t)B * p.Pl * L * 4 * sizeof(float), 256);   // per-token q~ of column attention
  p.off_head = off;    off = align_up(off + (size_t)B * ((L + 3) / 4) * p.Pl * sizeof(float), 256);   // fused head: partial site sums
  p.total = off;
  return p;
}

// ---- host-side weight packing ---------------------------------------------------------------
struct HostW {
  std::vector<std::vector<float>> t;  // tensors in state-dict order
};

void pack_attn(const HostW& w, int base, int ln_base, PfAttnW* o) {
  const float* kw = w.t[base + 0].data(); const float* kb = w.t[base + 1].data();
  const float* qw = w.t[base + 2].data(); const float* qb = w.t[base + 3].data();
  const float* vw = w.t[base + 4].data(); const float* vb = w.t[base + 5].data();
  const float* ow = w.t[base + 6].data(); const float* ob = w.t[base + 7].data();
  const float* g = w.t[ln_base].data();   const float* be = w.t[ln_base + 1].data();
  for (int v = 0; v < 8; ++v) {
    const float* src = v < 4 ? kw + v * PF_D : qw + (v - 4) * PF_D;
    double acc = v < 4 ? kb[v] : qb[v - 4];
    for (int c = 0; c < PF_D; ++c) {
      o->wqk[v][c] = (float)((double)src[c] * (double)g[c]);
      acc += (double)src[c] * (double)be[c];
    }
    o->bqk[v] = (float)acc;
  }
  for (int c = 0; c < PF_D; ++c) {
    o->gamma[c] = g[c];
    o->beta[c] = be[c];
    o->bv[c] = vb[c];
    o->bo[c] = ob[c];
    for (int k = 0; k < PF_D; ++k) {
      o->wvT[k][c] = vw[c * PF_D + k];
      o->wo[c][k] = ow[c * PF_D + k];
    }
  }
}

void pack_ffn(const HostW& w, int base, int ln_base, PfFfnW* o) {
  const float* w1 = w.t[base + 0].data(); const float* b1 = w.t[base + 1].data();  // (256,64)
  const float* w2 = w.t[base + 2].data(); const float* b2 = w.t[base + 3].data();  // (64,256)
  const float* g = w.t[ln_base].data();   const float* be = w.t[ln_base + 1].data();
  for (int jj = 0; jj < PF_HID; ++jj) {
    double acc = b1[jj];
    for (int k = 0; k < PF_D; ++k) {
      o->w1T[k][jj] = (float)((double)w1[jj * PF_D + k] * (double)g[k]);
      acc += (double)w1[jj * PF_D + k] * (double)be[k];
    }
    o->b1[jj] = (float)acc;
    for (int c = 0; c < PF_D; ++c) o->w2T[jj][c] = w2[c * PF_HID + jj];
  }
  for (int c = 0; c < PF_D; ++c) o->b2[c] = b2[c];
}

}  // namespace

extern "C" {

int pf_abi_version(void) { return PF_ABI_VERSION; }
const char* pf_last_error(void) { return g_err; }

int pf_create(pf_handle* out, const pf_cfg* cfg, const float* const* weights_dev, int n_weights) {
  if (!out || !cfg || !weights_dev) return fail(PF_ERR_ARG, "pf_create: null argument");
  if (cfg->nb_heads != PF_H || cfg->embed_dim != PF_D || cfg->ffn_mult != 4 || cfg->nb_blocks < 1)
    return fail(PF_ERR_ARG, "pf_create: only nb_heads=4, embed_dim=64, ffn_mult=4 are built (got %d,%d,%d)",
                cfg->nb_heads, cfg->embed_dim, cfg->ffn_mult);
  const int nb = cfg->nb_blocks;
  if (n_weights != PF_N_WEIGHTS(nb))
    return fail(PF_ERR_ARG, "pf_create: expected %d weight tensors, got %d", PF_N_WEIGHTS(nb), n_weights);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(PF_ERR_NO_DEVICE, "pf_create: no CUDA device (there is no CPU fallback)");
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(PF_ERR_NO_DEVICE, "pf_create: device %s is sm_%d%d; this library is built for sm_100a only",
                prop.name, prop.major, prop.minor);

  // sizes in state-dict order
  std::vector<size_t> sz;
  sz.push_back((size_t)PF_D * PF_NCHAR); sz.push_back(PF_D);
  for (int b = 0; b < nb; ++b) {
    for (int a = 0; a < 2; ++a) {
      sz.push_back(PF_H * PF_D); sz.push_back(PF_H); sz.push_back(PF_H * PF_D); sz.push_back(PF_H);
      sz.push_back(PF_D * PF_D); sz.push_back(PF_D); sz.push_back(PF_D * PF_D); sz.push_back(PF_D);
    }
    for (int i = 0; i < 6; ++i) sz.push_back(PF_D);
    sz.push_back((size_t)PF_HID * PF_D); sz.push_back(PF_HID);
    sz.push_back((size_t)PF_D * PF_HID); sz.push_back(PF_D);
  }
  sz.push_back(PF_D); sz.push_back(1);
  HostW hw;
  hw.t.resize(sz.size());
  for (size_t i = 0; i < sz.size(); ++i) {
    if (!weights_dev[i]) return fail(PF_ERR_ARG, "pf_create: weight %zu is null", i);
    hw.t[i].resize(sz[i]);
    CUDA_TRY(cudaMemcpy(hw.t[i].data(), weights_dev[i], sz[i] * sizeof(float), cudaMemcpyDeviceToHost));
  }

  pf_ctx* h = new pf_ctx();
  h->cfg = *cfg;
  h->n_sm = prop.multiProcessorCount;

  std::vector<PfHeadW> head(1);
  memset(head.data(), 0, sizeof(PfHeadW));
  {
    const float* we = hw.t[0].data();  // (64,22)
    const float* be = hw.t[1].data();
    for (int a = 0; a < PF_NCHAR; ++a)
      for (int c = 0; c < PF_D; ++c) {
        const float v = we[c * PF_NCHAR + a] + be[c];
        head[0].table[a][c] = v > 0.f ? v : 0.f;
        head[0].weT[a][c] = we[c * PF_NCHAR + a];
      }
    for (int c = 0; c < PF_D; ++c) head[0].be[c] = be[c];
    const size_t hb = sz.size() - 2;
    for (int c = 0; c < PF_D; ++c) head[0].whead[c] = hw.t[hb][c];
    head[0].bhead = hw.t[hb + 1][0];
  }
  std::vector<PfBlockW> blk(nb);
  std::vector<PfFfnTcW> tc(nb), tc16(nb), tcf(nb);
  std::vector<PfAttnTcW> atc(2 * nb);
  for (int b = 0; b < nb; ++b) {
    const int base = 2 + 26 * b;
    pack_attn(hw, base + 0, base + 16, &blk[b].row);
    pack_attn(hw, base + 8, base + 18, &blk[b].col);
    pack_ffn(hw, base + 22, base + 20, &blk[b].ffn);
    pf_pack_ffn_tc(blk[b].ffn, &tc[b]);
    pf_pack_ffn_tc(blk[b].ffn, &tc16[b], true);
    pf_pack_ffn_tc(blk[b].ffn, &tcf[b], false, true);
    pf_pack_attn_tc(blk[b].row, &atc[2 * b]);
    pf_pack_attn_tc(blk[b].col, &atc[2 * b + 1]);
    PfFfnConst kc;
    for (int c = 0; c < PF_D; ++c) {
      for (int hh = 0; hh < PF_H; ++hh) kc.wq[c][hh] = blk[b].col.wqk[4 + hh][c];
      kc.bo[c] = blk[b].col.bo[c];
      kc.b2[c] = blk[b].ffn.b2[c];
      kc.whead[c] = head[0].whead[c];
    }
    kc.bhead = head[0].bhead;
    for (int jj = 0; jj < PF_HID; ++jj) kc.b1[jj] = blk[b].ffn.b1[jj];
    kc.pad[0] = kc.pad[1] = kc.pad[2] = 0.f;
    for (int hh = 0; hh < PF_H; ++hh) kc.bq[hh] = blk[b].col.bqk[4 + hh];
    h->ffn_const.push_back(kc);
  }
  // Residue-pair tables for block 0's row attention (k_row_attn_combo): LN(x0), k~ and q~ of the 22 x 22 possible
  // pair embeddings x0 = T[a] + T[b] (fp32 sum as on the device, everything after it in fp64)
  std::vector<Row0Tab> row0(1);
  for (int a = 0; a < PF_NCHAR; ++a)
    for (int bb = 0; bb < PF_NCHAR; ++bb) {
      const int c = a * PF_NCHAR + bb;
      double x0[PF_D], mean = 0.0, var = 0.0;
      for (int k = 0; k < PF_D; ++k) { x0[k] = (double)(head[0].table[a][k] + head[0].table[bb][k]); mean += x0[k]; }
      mean /= PF_D;
      for (int k = 0; k < PF_D; ++k) var += (x0[k] - mean) * (x0[k] - mean);
      const double rstd = 1.0 / sqrt(var / PF_D + 1e-5);
      double nn[PF_D];
      for (int k = 0; k < PF_D; ++k) { nn[k] = (x0[k] - mean) * rstd; row0[0].n[c][k] = (float)nn[k]; }
      for (int v = 0; v < 8; ++v) {
        double z = blk[0].row.bqk[v];
        for (int k = 0; k < PF_D; ++k) z += (double)blk[0].row.wqk[v][k] * nn[k];
        row0[0].kq[c][v] = (float)(z > 0.0 ? z + 1.0 : exp(z));
      }
    }
  auto cleanup = [&]() { pf_destroy(h); };
  cudaError_t e;
  if ((e = cudaMalloc(&h->head_dev, sizeof(PfHeadW))) != cudaSuccess ||
      (e = cudaMalloc(&h->blk_dev, sizeof(PfBlockW) * nb)) != cudaSuccess ||
      (e = cudaMalloc(&h->tc_dev, sizeof(PfFfnTcW) * nb)) != cudaSuccess ||
      (e = cudaMalloc(&h->tc16_dev, sizeof(PfFfnTcW) * nb)) != cudaSuccess ||
      (e = cudaMalloc(&h->tcf_dev, sizeof(PfFfnTcW) * nb)) != cudaSuccess ||
      (e = cudaMemcpy(h->tcf_dev, tcf.data(), sizeof(PfFfnTcW) * nb, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMalloc(&h->atc_dev, sizeof(PfAttnTcW) * 2 * nb)) != cudaSuccess ||
      (e = cudaMemcpy(h->atc_dev, atc.data(), sizeof(PfAttnTcW) * 2 * nb, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMalloc(&h->row0_dev, sizeof(Row0Tab))) != cudaSuccess ||
      (e = cudaMemcpy(h->row0_dev, row0.data(), sizeof(Row0Tab), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMalloc(&h->err_dev, sizeof(int))) != cudaSuccess ||
      (e = cudaMemset(h->err_dev, 0, sizeof(int))) != cudaSuccess ||
      (e = cudaMemcpy(h->head_dev, head.data(), sizeof(PfHeadW), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(h->blk_dev, blk.data(), sizeof(PfBlockW) * nb, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(h->tc_dev, tc.data(), sizeof(PfFfnTcW) * nb, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(h->tc16_dev, tc16.data(), sizeof(PfFfnTcW) * nb, cudaMemcpyHostToDevice)) != cudaSuccess) {
    cleanup();
    return fail(PF_ERR_CUDA, "pf_create: %s", cudaGetErrorString(e));
  }
  // opt in to large dynamic shared memory once
  CUDA_TRY(cudaFuncSetAttribute(k_colapply_ffn_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(Ffn32Smem)));
  CUDA_TRY(cudaFuncSetAttribute(k_row_attn<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUDA_TRY(cudaFuncSetAttribute(k_row_attn<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUDA_TRY(cudaFuncSetAttribute(k_row_attn_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CUDA_TRY(cudaFuncSetAttribute(k_row_attn_combo, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  if (const char* e_r0 = getenv("PF_ROW0_IMPL")) h->row0_impl = (strcmp(e_r0, "ffma") == 0) ? 0 : 1;
  if (const char* e_impl = getenv("PF_FFN_IMPL")) h->ffn_impl = (strcmp(e_impl, "tc") == 0) ? 0 : 1;
  if (const char* e_row = getenv("PF_ROW_IMPL")) h->row_impl = (strcmp(e_row, "ld") == 0) ? 0 : (strcmp(e_row, "tma") == 0) ? 1 : 2;
  if (const char* e_prof = getenv("PF_WS_PROF")) h->ws_prof = atoi(e_prof);
  if (const char* e_ex = getenv("PF_EXCH_IMPL")) h->exch_impl = (strcmp(e_ex, "split") == 0) ? 0 : 1;
  if (const char* e_hd = getenv("PF_HEAD_IMPL")) h->head_impl = (strcmp(e_hd, "sep") == 0) ? 0 : 1;
  if (const char* e_col = getenv("PF_COL_IMPL")) h->col_impl = (strcmp(e_col, "cc") == 0) ? 0 : (strcmp(e_col, "tc1") == 0) ? 1 : 2;
  // Load every kernel that spins on (or is waited for by) a peer NOW: with CUDA's lazy module loading a first launch can
  // block the host until the device is idle, i.e. until a peer-wait kernel already running on it has timed out
  // (seen with two ranks driven from one process: tests/test_gpu_peer_emulated.py).
  {
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_col_exchange));
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_col_reduce));
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_peer_sync));
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_col_finalize_peer));
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_col_finalize));
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_head_reduce));
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_head));
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_col_partial));
  }
  int rc = pf_ffn_tc_init();
  if (rc == 0) rc = pf_ffn_ws_init();
  if (rc == 0) rc = pf_attn_tc_init();
  if (rc != 0) { cleanup(); return fail(PF_ERR_CUDA, "pf_create: tcgen05 FFN kernel setup failed (%d)", rc); }
  *out = h;
  return PF_OK;
}

void pf_destroy(pf_handle h) {
  if (!h) return;
  if (h->head_dev) cudaFree(h->head_dev);
  if (h->blk_dev) cudaFree(h->blk_dev);
  if (h->tc_dev) cudaFree(h->tc_dev);
  if (h->tc16_dev) cudaFree(h->tc16_dev);
  if (h->tcf_dev) cudaFree(h->tcf_dev);
  if (h->atc_dev) cudaFree(h->atc_dev);
  if (h->err_dev) cudaFree(h->err_dev);
  if (h->row0_dev) cudaFree(h->row0_dev);
  if (h->peers_dev) cudaFree(h->peers_dev);
  for (auto& r : h->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : h->pool) cudaEventDestroy(e);
  delete h;
}

int pf_set_precision(pf_handle h, int precision) {
  if (!h) return fail(PF_ERR_ARG, "pf_set_precision: null handle");
  if (precision < PF_PREC_FP32 || precision > PF_PREC_FP16) return fail(PF_ERR_ARG, "pf_set_precision: bad mode %d", precision);
  h->cfg.precision = precision;
  return PF_OK;
}

size_t pf_workspace_bytes(pf_handle h, int B, int n, int L, int64_t pair_lo, int64_t pair_hi) {
  if (!h || B < 1 || n < 2 || L < 1) return 0;
  const long long P = (long long)n * (n - 1) / 2;
  if (pair_lo < 0 || pair_hi > P || pair_hi < pair_lo) return 0;
  return make_plan(h, B, n, L, pair_lo, pair_hi).total;
}

int pf_onehot_to_idx(const float* x_dev, int B, int L, int n, uint8_t* idx_dev, int32_t* not_onehot_dev,
                     void* stream) {
  if (!x_dev || !idx_dev || !not_onehot_dev || B < 1 || L < 1 || n < 1) return fail(PF_ERR_ARG, "pf_onehot_to_idx: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(not_onehot_dev, 0, sizeof(int32_t), st));
  const long long total = (long long)B * L * n;
  k_onehot_to_idx<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x_dev, B, L, n, idx_dev, not_onehot_dev);
  CUDA_TRY(cudaGetLastError());
  return PF_OK;
}

int pf_forward_debug(pf_handle h, const uint8_t* msa_idx_dev, const float* x_dev, const int32_t* not_onehot_dev,
                     int B, int n, int L, int64_t pair_lo, int64_t pair_hi, float* dist_dev, void* ws_dev,
                     size_t ws_bytes, void* stream, pf_reduce_fn reduce, void* reduce_user, int n_stages,
                     float* act_dev) {
  if (!h) return fail(PF_ERR_ARG, "pf_forward: null handle");
  if (!msa_idx_dev || !ws_dev) return fail(PF_ERR_ARG, "pf_forward: null buffer");
  if (B < 1 || n < 2 || L < 1) return fail(PF_ERR_ARG, "pf_forward: need B>=1, n>=2, L>=1 (got %d,%d,%d)", B, n, L);
  if (B > 65535) return fail(PF_ERR_ARG, "pf_forward: B=%d exceeds 65535 alignments per call (the batch index is a launch-grid y/z dimension); split the batch", B);
  const long long P = (long long)n * (n - 1) / 2;
  if (pair_lo < 0 || pair_hi > P || pair_hi <= pair_lo)
    return fail(PF_ERR_ARG, "pf_forward: bad pair range [%lld,%lld) of %lld", (long long)pair_lo, (long long)pair_hi, P);
  if ((pair_hi - pair_lo) != P && reduce == nullptr && h->peer_world <= 1)
    return fail(PF_ERR_ARG, "pf_forward: a partial pair range needs a reduce callback or a peer exchange");
  if (!dist_dev && n_stages < 0) return fail(PF_ERR_ARG, "pf_forward: null dist buffer");
  const Plan pl = make_plan(h, B, n, L, pair_lo, pair_hi);
  if (ws_bytes < pl.total) return fail(PF_ERR_WORKSPACE, "pf_forward: workspace %zu < %zu bytes", ws_bytes, pl.total);
  if ((long long)B * pl.Pl > 0x7fffffffLL) return fail(PF_ERR_ARG, "pf_forward: too many pair rows");
  const size_t row_smem = sizeof(RowSmem) + (size_t)L * 4 * sizeof(float);
  // implementation of the row kernel for THIS call (blocks 1..): tensor cores when the mode and the row length allow
  // it, else the bulk-copy ring, else (very long rows) plain register loads
  int row_impl = h->row_impl;
  if (row_impl == 2 && (h->cfg.precision == PF_PREC_FP32 || rw_smem_bytes(L) > 227 * 1024)) row_impl = 1;
  if (row_impl == 1 && row_smem + RT_STAGES * RT_STAGE_BYTES > 220 * 1024) row_impl = 0;
  if (row_smem > 200 * 1024) return fail(PF_ERR_ARG, "pf_forward: L=%d exceeds the row kernel's shared-memory budget", L);

  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* ws = (unsigned char*)ws_dev;
  float* x = (float*)(ws + pl.off_x);
  float* part = (float*)(ws + pl.off_part);
  float* colsum = (float*)(ws + pl.off_colsum);
  float* colM = (float*)(ws + pl.off_colM);
  float* semb = (float*)(ws + pl.off_semb);
  float* qcache = (float*)(ws + pl.off_qc);
  float* headpart = (float*)(ws + pl.off_head);
  bool head_fused = false;   // the last block's FFN launch emitted the head's partial sums (no k_head pass)
  const int rows = (int)(B * pl.Pl);
  const long long n_tok = (long long)rows * L;
  const int nb = h->cfg.nb_blocks;
  const bool dbg = n_stages >= 0;
  int stage = 0;
  h->launches = 0;
  auto done = [&]() -> int {
    if (dbg && act_dev) {
      cudaError_t e = cudaMemcpyAsync(act_dev, x, (size_t)n_tok * PF_D * sizeof(float), cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return fail(PF_ERR_CUDA, "debug copy: %s", cudaGetErrorString(e));
    }
    return PF_OK;
  };

  if (x_dev != nullptr && not_onehot_dev != nullptr) {
    const long long tot = (long long)B * n * L * PF_D;
    Timed t_(h, PF_KC_INPUT, st);
    k_embed_sequences<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(h->head_dev, x_dev, not_onehot_dev, B, L, n, semb);
  }
  const int* flag = (x_dev != nullptr) ? not_onehot_dev : nullptr;
  alignas(64) CUtensorMap x_tmap;
  memset(&x_tmap, 0, sizeof(x_tmap));
  alignas(64) CUtensorMap x_tmap2;
  memset(&x_tmap2, 0, sizeof(x_tmap2));
  if (h->cfg.precision != PF_PREC_FP32 &&
      (!make_x_tensor_map(&x_tmap, x, (long long)rows, L) || !make_x_tensor_map_2d(&x_tmap2, x, n_tok)))
    return fail(PF_ERR_CUDA, "pf_forward: cuTensorMapEncodeTiled failed (driver too old for TMA tensor maps?)");

  for (int b = 0; b < nb; ++b) {
    const PfBlockW* bw = h->blk_dev + b;
    // ---- row attention (block 0: fused with the pair embedding) ----
    if (b == 0) {
      const int embed_only = (dbg && n_stages == 0) ? 1 : 0;
      // residue codes (forward_idx, or forward(x) whose x turned out one-hot: device flag): tables + gather;
      // soft inputs: the FFMA kernel.  Both check the flag on the device; the one that does not apply exits at once.
      const bool combo = h->row0_impl == 1 && !embed_only && row0_smem_bytes(L) <= 200 * 1024;
      if (combo) {
        Timed t_(h, PF_KC_ROW, st);
        k_row_attn_combo<<<rows, 256, row0_smem_bytes(L), st>>>(&bw->row, h->head_dev, h->row0_dev, x, msa_idx_dev, flag, n, L,
                                                               pair_lo, (int)pl.Pl);
      }
      if (!combo || flag != nullptr) {
        Timed t_(h, PF_KC_ROW, st);
        k_row_attn<1><<<rows, 256, row_smem, st>>>(&bw->row, h->head_dev, x, msa_idx_dev, semb, flag, n, L, pair_lo,
                                                   (int)pl.Pl, combo ? 2 : embed_only);
      }
      CUDA_TRY(cudaGetLastError());
      if (embed_only) return done();
    } else {
      Timed t_(h, PF_KC_ROW, st);
      if (row_impl == 2)
        k_row_attn_ws<<<rows < h->n_sm ? rows : h->n_sm, RW_THREADS, rw_smem_bytes(L), st>>>(x_tmap2, &bw->row, h->atc_dev + 2 * b, x,
                                                                                          rows, L, h->err_dev);
      else if (row_impl == 1)
        k_row_attn_tma<<<rows, 256, row_smem + RT_STAGES * RT_STAGE_BYTES, st>>>(&bw->row, x, L, h->err_dev);
      else
        k_row_attn<0><<<rows, 256, row_smem, st>>>(&bw->row, h->head_dev, x, nullptr, nullptr, nullptr, n, L, pair_lo,
                                                   (int)pl.Pl, 0);
    }
    CUDA_TRY(cudaGetLastError());
    stage = 3 * b + 1;
    if (dbg && stage == n_stages) return done();
    // ---- column attention summaries ----
    const bool apply_only = dbg && (n_stages == 3 * b + 2);
    // tensor-core summaries (and the q~ cache the FFN kernel then reads) in every mode but the fp32 one
    const bool col_tc = h->col_impl >= 1 && h->cfg.precision != PF_PREC_FP32 && h->ffn_impl == 1;
    {
      if (col_tc && h->col_impl == 2) {
        const long long units = (long long)B * ((L + 3) / 4) * pl.n_chunks_ws;
        const int g = (int)(units < (long long)h->n_sm ? units : (long long)h->n_sm);
        Timed t_(h, PF_KC_COLSUM, st);
        k_col_partial_ws<<<g, C2_THREADS, C2_SMEM_BYTES, st>>>(x_tmap, h->atc_dev + 2 * b + 1, part, qcache, B, L, (int)pl.Pl,
                                                              pl.ppc_ws, pl.n_chunks_ws, h->err_dev);
      } else if (col_tc) {
        const long long units = (long long)B * ((L + 3) / 4) * pl.n_chunks_tc;
        const int g = (int)(units < 2LL * h->n_sm ? units : 2LL * h->n_sm);
        Timed t_(h, PF_KC_COLSUM, st);
        k_col_partial_tc<<<g, CT_THREADS, CT_SMEM_BYTES, st>>>(h->atc_dev + 2 * b + 1, x, part, qcache, B, L, (int)pl.Pl,
                                                              pl.ppc_tc, pl.n_chunks_tc, h->err_dev);
      } else {
        dim3 g((L + 31) / 32, pl.n_chunks, B);
        Timed t_(h, PF_KC_COLSUM, st);
        k_col_partial<<<g, 256, 0, st>>>(&bw->col, x, part, L, (int)pl.Pl, pl.ppc);
      }
      const int n_part = !col_tc ? pl.n_chunks : h->col_impl == 2 ? pl.n_chunks_ws : pl.n_chunks_tc;
      // column reduce/finalize: spc sites per CTA (weights stay in registers across them), about
      // two CTAs per SM when there are few sites, PF_FS per CTA for batches of small alignments
      const int n_sites = B * L;
      const int spc = std::max(1, std::min(PF_FS, n_sites / (2 * h->n_sm)));
      const unsigned gfs = (unsigned)((n_sites + spc - 1) / spc);
      const bool peer = (h->peer_world > 1) && ((pair_hi - pair_lo) != P);
      // fused form: at most PF_PEER_MAX_CTAS - 1 CTAs (the flag table has one word per rank and CTA; the last column
      // belongs to the three-launch form), each looping over its site groups; the same grid on every rank
      const unsigned gex = std::min(gfs, (unsigned)std::min(PF_PEER_MAX_CTAS - 1, 2 * h->n_sm));
      if (peer) {
        const size_t need = (size_t)B * L * PF_COLSUM;
        if (need > h->peer_slot_floats)
          return fail(PF_ERR_ARG, "pf_forward: peer exchange slot holds %zu floats, need %zu", h->peer_slot_floats, need);
        const unsigned epoch = ++h->peer_epoch;
        const int slot = (int)(epoch & 1u);
        float* my_slot = reinterpret_cast<float*>(h->peer_self + PF_PEER_FLAG_BYTES) + (size_t)slot * h->peer_slot_floats;
        if (h->exch_impl == 1) {
          Timed t_(h, PF_KC_COLFIN, st);
          k_col_exchange<<<gex, 256, 0, st>>>(&bw->col, part, n_part, n_sites, spc, my_slot, h->peers_dev, h->peer_rank,
                                              h->peer_world, slot, h->peer_slot_floats, epoch, (float)P, colM, h->err_dev);
        } else {
          {
            Timed t_(h, PF_KC_COLFIN, st);
            k_col_reduce<<<gfs, 256, 0, st>>>(&bw->col, part, n_part, n_sites, spc, my_slot);
          }
          {
            Timed t_(h, PF_KC_COLFIN, st);
            k_peer_sync<<<1, 32, 0, st>>>(h->peers_dev, h->peer_rank, h->peer_world, epoch, h->err_dev);
          }
          {
            Timed t_(h, PF_KC_COLFIN, st);
            k_col_finalize_peer<<<gfs, 256, 0, st>>>(&bw->col, h->peers_dev, h->peer_world, slot, h->peer_slot_floats,
                                                     (float)P, n_sites, spc, colM);
          }
        }
        CUDA_TRY(cudaGetLastError());
      } else if (!reduce && h->exch_impl == 1) {   // one rank holds every pair: same kernel, no flags, no peers, one site group per CTA
        Timed t_(h, PF_KC_COLFIN, st);
        k_col_exchange<<<gfs, 256, 0, st>>>(&bw->col, part, n_part, n_sites, spc, colsum, nullptr, 0, 1, 0, 0, 0u, (float)P,
                                            colM, h->err_dev);
        CUDA_TRY(cudaGetLastError());
      } else {
        {
          Timed t_(h, PF_KC_COLFIN, st);
          k_col_reduce<<<gfs, 256, 0, st>>>(&bw->col, part, n_part, n_sites, spc, colsum);
        }
        CUDA_TRY(cudaGetLastError());
        if (reduce) {
          const int rc = reduce(reduce_user, colsum, (size_t)B * L * PF_COLSUM, stream);
          if (rc != 0) return fail(PF_ERR_REDUCE, "pf_forward: reduce callback returned %d", rc);
        }
        {
          Timed t_(h, PF_KC_COLFIN, st);
          k_col_finalize<<<gfs, 256, 0, st>>>(&bw->col, colsum, (float)P, n_sites, spc, colM);
        }
        CUDA_TRY(cudaGetLastError());
      }
    }
    // ---- column apply + FFN ----
    if (h->cfg.precision == PF_PREC_FP32 || apply_only) {
      long long tiles = (n_tok + FFN32_T - 1) / FFN32_T;
      const int grid = (int)(tiles < h->n_sm ? tiles : h->n_sm);
      Timed t_(h, PF_KC_FFN, st);
      k_colapply_ffn_fp32<<<grid, 256, sizeof(Ffn32Smem), st>>>(&bw->col, &bw->ffn, x, colM, L, (int)pl.Pl, n_tok,
                                                               apply_only ? 1 : 0);
    } else {
      Timed t_(h, PF_KC_FFN, st);
      const int prec = h->cfg.precision;
      // last block, not a debug run, no accumulator dump / role timing requested: fuse the distance head
      const bool fuse_head = (b == nb - 1) && !dbg && h->ffn_impl == 1 && h->head_impl == 1 && h->dump_dev == nullptr && !h->ws_prof;
      head_fused = fuse_head;
      const int terms = prec == PF_PREC_BF16 ? 1 : prec == PF_PREC_FP16 ? 2 : 3;
      const int fmt = prec == PF_PREC_BF16 ? WS_FMT_BF16 : prec == PF_PREC_FP16 ? WS_FMT_F16 : WS_FMT_BF16X3;
      if (prec == PF_PREC_FP16 && h->ffn_impl != 1)
        return fail(PF_ERR_ARG, "pf_forward: PF_PREC_FP16 needs the warp-specialised FFN kernel (unset PF_FFN_IMPL)");
      const int rc = h->ffn_impl == 1
                         ? pf_ffn_ws_launch(h->ffn_const[b],
                                            (WS_FOLD63 && fmt == WS_FMT_BF16X3) ? h->tcf_dev + b
                                                      : (prec == PF_PREC_FP16 ? h->tc16_dev : h->tc_dev) + b, x, colM,
                                            col_tc ? qcache : nullptr, L,
                                            (int)pl.Pl, B, h->n_sm, fmt, terms, h->err_dev, h->dump_dev, h->ws_prof, st,
                                            fuse_head ? headpart : nullptr)
                         : pf_ffn_tc_launch(&bw->col, h->tc_dev + b, x, colM, L, (int)pl.Pl, n_tok, h->n_sm, terms,
                                            h->err_dev, h->dump_dev, st);
      if (rc != 0) return fail(PF_ERR_CUDA, "pf_forward: tcgen05 FFN launch failed (%d)", rc);
    }
    CUDA_TRY(cudaGetLastError());
    if (apply_only) return done();
    stage = 3 * b + 3;
    if (dbg && stage == n_stages) return done();
  }
  if (head_fused) {
    Timed t_(h, PF_KC_HEAD, st);
    dim3 g((unsigned)((pl.Pl + 255) / 256), B);
    k_head_reduce<<<g, 256, 0, st>>>(headpart, (L + 3) / 4, (int)pl.Pl, L, dist_dev);
  } else {
    Timed t_(h, PF_KC_HEAD, st);
    k_head<<<rows, 256, 0, st>>>(h->head_dev, x, L, dist_dev);
  }
  CUDA_TRY(cudaGetLastError());
  return done();
}

int pf_forward(pf_handle h, const uint8_t* msa_idx_dev, const float* x_dev, const int32_t* not_onehot_dev, int B,
               int n, int L, int64_t pair_lo, int64_t pair_hi, float* dist_dev, void* ws_dev, size_t ws_bytes,
               void* stream, pf_reduce_fn reduce, void* reduce_user) {
  if (!dist_dev) return fail(PF_ERR_ARG, "pf_forward: null dist buffer");
  return pf_forward_debug(h, msa_idx_dev, x_dev, not_onehot_dev, B, n, L, pair_lo, pair_hi, dist_dev, ws_dev,
                          ws_bytes, stream, reduce, reduce_user, -1, nullptr);
}

int pf_dist_to_matrix(const float* dist_dev, int B, int n, float* mat_dev, void* stream) {
  if (!dist_dev || !mat_dev || B < 1 || n < 2) return fail(PF_ERR_ARG, "pf_dist_to_matrix: bad argument");
  const long long P = (long long)n * (n - 1) / 2;
  dim3 g((unsigned)(((long long)n * n + 255) / 256), B);
  k_dist_to_matrix<<<g, 256, 0, (cudaStream_t)stream>>>(dist_dev, n, P, mat_dev);
  CUDA_TRY(cudaGetLastError());
  return PF_OK;
}

// ---- host-side FASTA parse (no device work) ---------------------------------------------------
static inline bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }   // bytes.strip() set

long long pf_parse_fasta(const char* text, long long len, uint8_t* codes, long long cap, int32_t* L_out,
                         int64_t* name_off, int32_t* name_len, int32_t max_names, int32_t* bad_char) {
  if (!text || len < 0 || !codes || !L_out || !name_off || !name_len)
    return (long long)fail(PF_ERR_ARG, "pf_parse_fasta: null argument");
  static const char alphabet[] = "ARNDCQEGHILKMFPSTWYVX-";   // reference data.py:7
  uint8_t lut[256];
  memset(lut, 255, sizeof(lut));
  for (int i = 0; alphabet[i]; ++i) lut[(unsigned char)alphabet[i]] = (uint8_t)i;
  long long n = 0, pos = 0, cur_len = 0, L = -1;
  bool ragged = false;
  auto close_record = [&]() {
    if (n == 0) return;
    if (L < 0) L = cur_len; else if (cur_len != L) ragged = true;
  };
  long long i = 0;
  while (i < len) {
    long long e = i;
    while (e < len && text[e] != '\n') ++e;
    long long a = i, b = e;                      // strip the line like bytes.strip()
    while (a < b && is_space((unsigned char)text[a])) ++a;
    while (b > a && is_space((unsigned char)text[b - 1])) --b;
    if (a < b && text[a] == '>') {
      close_record();
      if (n >= max_names) return (long long)fail(PF_ERR_ARG, "pf_parse_fasta: more than %d records", max_names);
      name_off[n] = a + 1;
      name_len[n] = (int32_t)(b - a - 1);
      ++n;
      cur_len = 0;
    } else if (a < b) {
      if (n == 0) return (long long)fail(PF_ERR_FASTA_NOHEADER, "pf_parse_fasta: sequence data before the first '>' header");
      for (long long k = a; k < b; ++k) {
        const uint8_t c = lut[(unsigned char)text[k]];
        if (c == 255) {
          if (bad_char) *bad_char = (unsigned char)text[k];
          return (long long)fail(PF_ERR_FASTA_RESIDUE, "pf_parse_fasta: residue 0x%02x outside the alphabet", (unsigned char)text[k]);
        }
        if (pos >= cap) return (long long)fail(PF_ERR_ARG, "pf_parse_fasta: code buffer too small");
        codes[pos++] = c;
      }
      cur_len += b - a;
    }
    i = e + 1;
  }
  close_record();
  if (ragged) return (long long)fail(PF_ERR_FASTA_RAGGED, "pf_parse_fasta: sequences have different lengths");
  *L_out = (int32_t)(L < 0 ? 0 : L);
  return n;
}

// ---- host-side PHYLIP text (no device work) --------------------------------------------------
// "%.10f" of a float, exactly as printf / Python format it (correctly rounded, ties to even),
// through integer arithmetic: f = mant * 2^e, so round(f * 10^10) = round(mant * 10^10 / 2^-e).
static inline int fmt_fixed10(float f, char* out) {
  uint32_t bits;
  memcpy(&bits, &f, 4);
  const uint32_t ex = (bits >> 23) & 0xffu;
  if ((bits >> 31) || ex == 0xffu || f >= 1.0e6f) {   // negative, inf/nan or huge: libc path
    const double d = (double)f;
    if (d != d) { memcpy(out, "nan", 3); return 3; }   // Python prints nan / inf / -inf
    return snprintf(out, 64, "%.10f", d);
  }
  const uint32_t mant = (bits & 0x7fffffu) | (ex ? 0x800000u : 0u);
  const int shift = 150 - (int)(ex ? ex : 1u);          // f = mant * 2^-shift, shift in [5, 149] here
  unsigned long long q = 0;
  if (shift < 100) {
    const unsigned __int128 num = (unsigned __int128)mant * 10000000000ULL;   // < 2^58
    const unsigned __int128 one = 1;
    q = (unsigned long long)(num >> shift);
    const unsigned __int128 rem = num & ((one << shift) - 1), half = one << (shift - 1);
    if (rem > half || (rem == half && (q & 1ULL))) ++q;
  }
  const unsigned long long ip = q / 10000000000ULL, fp = q % 10000000000ULL;
  char tmp[24];
  int ni = 0;
  unsigned long long t = ip;
  do { tmp[ni++] = (char)('0' + t % 10); t /= 10; } while (t);
  int k = 0;
  while (ni) out[k++] = tmp[--ni];
  out[k++] = '.';
  unsigned long long r = fp;
  for (int dgt = 9; dgt >= 0; --dgt) { out[k + dgt] = (char)('0' + r % 10); r /= 10; }
  return k + 10;
}

long long pf_format_phylip(const float* dm_host, int n, const char* const* names, char* out, long long cap) {
  if (!dm_host || !names || n < 1 || (cap > 0 && !out)) return (long long)fail(PF_ERR_ARG, "pf_format_phylip: bad argument");
  long long pos = 0;
  char num[80];
  auto put = [&](const char* src, size_t len) {
    if (pos + (long long)len <= cap) memcpy(out + pos, src, len);
    pos += (long long)len;
  };
  put(num, (size_t)snprintf(num, sizeof(num), "%d\n", n));
  for (int i = 0; i < n; ++i) {
    if (!names[i]) return (long long)fail(PF_ERR_ARG, "pf_format_phylip: null name %d", i);
    put(names[i], strlen(names[i]));
    for (int j = 0; j < n; ++j) {
      num[0] = ' ';
      put(num, (size_t)(1 + fmt_fixed10(dm_host[(size_t)i * n + j], num + 1)));
    }
    put("\n", 1);
  }
  return pos;   // bytes of text; the caller retries with a larger buffer if this exceeds cap
}

// Newick label: names holding blanks or Newick punctuation are single-quoted, embedded quotes doubled (nj.py: newick_label)
static std::string newick_label(const char* s) {
  std::string v(s);
  if (!v.empty() && v.find_first_of(" \t\r\n,:;()[]'") == std::string::npos) return v;
  std::string q = "'";
  for (char c : v) { q += c; if (c == '\'') q += c; }
  return q + "'";
}

// ---- host-side neighbour joining (no device work) ---------------------------------------------
// Saitou & Nei / Studier & Keppler, O(n^3) in double precision; same joins, tie-breaking (first
// minimum of Q in row-major order) and Newick layout as phyloformer_b200/nj.py.
long long pf_neighbor_joining(const float* dm_host, int n, const char* const* names, char* out, long long cap) {
  if (!dm_host || !names || n < 1 || (cap > 0 && !out)) return (long long)fail(PF_ERR_ARG, "pf_neighbor_joining: bad argument");
  for (int i = 0; i < n; ++i)
    if (!names[i]) return (long long)fail(PF_ERR_ARG, "pf_neighbor_joining: null name %d", i);
  auto fmt = [](double x) {
    char b[400];
    if (x < 0) x = 0.0;                      // clip negative branch lengths like nj.py
    snprintf(b, sizeof(b), "%.10f", x);
    return std::string(b);
  };
  auto label = [](const char* s) { return newick_label(s); };
  std::string tree;
  try {
  if (n == 1) {
    tree = label(names[0]) + ";";
  } else if (n == 2) {
    const double h = (double)dm_host[1] / 2;
    tree = "(" + label(names[0]) + ":" + fmt(h) + "," + label(names[1]) + ":" + fmt(h) + ");";
  } else {
    std::vector<double> d((size_t)n * n), r((size_t)n);
    for (size_t i = 0; i < (size_t)n * n; ++i) d[i] = (double)dm_host[i];
    std::vector<std::string> nodes((size_t)n);
    std::vector<int> active((size_t)n);
    for (int i = 0; i < n; ++i) { nodes[i] = label(names[i]); active[i] = i; }
    auto D = [&](int i, int j) -> double& { return d[(size_t)i * n + j]; };
    while (active.size() > 3) {
      const int m = (int)active.size();
      for (int a = 0; a < m; ++a) {
        double s = 0.0;
        for (int b = 0; b < m; ++b) s += D(active[a], active[b]);
        r[a] = s;
      }
      double best = 0.0;
      int ba = -1, bb = -1;
      for (int a = 0; a < m; ++a)
        for (int b = 0; b < m; ++b) {
          if (a == b) continue;
          const double q = (m - 2) * D(active[a], active[b]) - r[a] - r[b];
          if (ba < 0 || q < best) { best = q; ba = a; bb = b; }
        }
      if (ba > bb) { const int t = ba; ba = bb; bb = t; }
      const int ia = active[ba], ib = active[bb];
      const double dab = D(ia, ib);
      const double la = 0.5 * dab + (r[ba] - r[bb]) / (2.0 * (m - 2));
      const double lb = dab - la;
      nodes[ia] = "(" + nodes[ia] + ":" + fmt(la) + "," + nodes[ib] + ":" + fmt(lb) + ")";
      nodes[ib].clear();
      for (int k = 0; k < n; ++k) {            // the new node lives in ia's slot
        const double dn = 0.5 * (D(ia, k) + D(ib, k) - dab);
        D(ia, k) = dn;
      }
      for (int k = 0; k < n; ++k) D(k, ia) = D(ia, k);
      D(ia, ia) = 0.0;
      active.erase(active.begin() + bb);
    }
    const int i = active[0], j = active[1], k = active[2];
    const double li = 0.5 * (D(i, j) + D(i, k) - D(j, k));
    const double lj = 0.5 * (D(i, j) + D(j, k) - D(i, k));
    const double lk = 0.5 * (D(i, k) + D(j, k) - D(i, j));
    tree = "(" + nodes[i] + ":" + fmt(li) + "," + nodes[j] + ":" + fmt(lj) + "," + nodes[k] + ":" + fmt(lk) + ");";
  }
  } catch (...) {   // nothing may propagate across the C ABI
    return (long long)fail(PF_ERR_ARG, "pf_neighbor_joining: out of host memory (n = %d)", n);
  }
  const long long len = (long long)tree.size();
  if (len <= cap) memcpy(out, tree.data(), (size_t)len);
  return len;
}

// ---- host-side BIONJ + balanced NNI / SPR tree search (no device work; pf_bme.h) -----------------
long long pf_bme_tree(const double* dm_host, int n, const char* const* names, int flags, char* out, long long cap,
                      double* stats) {
  if (!dm_host || !names || n < 1 || (cap > 0 && !out)) return (long long)fail(PF_ERR_ARG, "pf_bme_tree: bad argument");
  if (n > PF_BME_MAX_TAXA)
    return (long long)fail(PF_ERR_ARG, "pf_bme_tree: %d taxa exceed the limit of %d (the table of subtree averages grows with n^2)", n, PF_BME_MAX_TAXA);
  for (int i = 0; i < n; ++i)
    if (!names[i]) return (long long)fail(PF_ERR_ARG, "pf_bme_tree: null name %d", i);
  for (size_t i = 0; i < (size_t)n * n; ++i)
    if (!std::isfinite(dm_host[i])) return (long long)fail(PF_ERR_ARG, "pf_bme_tree: non-finite distance at element %zu", i);
  std::string tree;
  try {
    std::vector<std::string> labels((size_t)n);
    for (int i = 0; i < n; ++i) labels[i] = newick_label(names[i]);
    const pfbme::Result r = pfbme::build(dm_host, n, labels, flags);
    tree = r.newick;
    if (stats) {
      stats[0] = r.length_own; stats[1] = r.length_start; stats[2] = r.length_nni; stats[3] = r.length_spr;
      stats[4] = r.n_nni; stats[5] = r.n_spr; stats[6] = r.kept;
    }
  } catch (...) {   // nothing may propagate across the C ABI
    return (long long)fail(PF_ERR_ARG, "pf_bme_tree: out of host memory (n = %d)", n);
  }
  const long long len = (long long)tree.size();
  if (len <= cap) memcpy(out, tree.data(), (size_t)len);
  return len;
}

int pf_last_launch_count(pf_handle h) { return h ? h->launches : 0; }

int pf_set_peer_exchange(pf_handle h, int rank, int world, void* const* peer_bufs_host, size_t slot_floats) {
  if (!h) return fail(PF_ERR_ARG, "pf_set_peer_exchange: null handle");
  if (h->peers_dev) { cudaFree(h->peers_dev); h->peers_dev = nullptr; }
  h->peer_world = 0;
  if (world <= 1 || peer_bufs_host == nullptr) return PF_OK;   // exchange disabled
  if (rank < 0 || rank >= world || world > 32) return fail(PF_ERR_ARG, "pf_set_peer_exchange: bad rank/world %d/%d", rank, world);
  CUDA_TRY(cudaMalloc(&h->peers_dev, sizeof(void*) * world));
  CUDA_TRY(cudaMemcpy(h->peers_dev, peer_bufs_host, sizeof(void*) * world, cudaMemcpyHostToDevice));
  h->peer_self = (unsigned char*)peer_bufs_host[rank];
  h->peer_rank = rank;
  h->peer_world = world;
  h->peer_slot_floats = slot_floats;
  h->peer_epoch = 0;
  return PF_OK;
}

size_t pf_peer_exchange_bytes(size_t slot_floats) { return PF_PEER_FLAG_BYTES + 2 * slot_floats * sizeof(float); }

int pf_device_error(pf_handle h) {
  if (!h) return fail(PF_ERR_ARG, "pf_device_error: null handle");
  int v = 0;
  CUDA_TRY(cudaMemcpy(&v, h->err_dev, sizeof(int), cudaMemcpyDeviceToHost));
  if (v != 0) {
    cudaMemset(h->err_dev, 0, sizeof(int));
    return fail(PF_ERR_CUDA, "a bounded device-side wait timed out (flag %d: 1/2 FFN pipeline, 3 peer exchange, 4-6 attention tile engine, 7 row staging ring); results of that call are invalid", v);
  }
  return PF_OK;
}

int pf_debug_set_dump(pf_handle h, float* dump_dev) {
  if (!h) return fail(PF_ERR_ARG, "pf_debug_set_dump: null handle");
  h->dump_dev = dump_dev;
  return PF_OK;
}

int pf_profile_enable(pf_handle h, int on) {
  if (!h) return fail(PF_ERR_ARG, "pf_profile_enable: null handle");
  h->prof = on != 0;
  return PF_OK;
}

int pf_profile_read(pf_handle h, float* ms_out, int32_t* launches_out) {
  if (!h || !ms_out || !launches_out) return fail(PF_ERR_ARG, "pf_profile_read: null argument");
  for (int i = 0; i < PF_KC_COUNT; ++i) { ms_out[i] = 0.f; launches_out[i] = 0; }
  for (auto& r : h->recs) {
    CUDA_TRY(cudaEventSynchronize(r.b));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_out[r.kc] += ms;
    ++launches_out[r.kc];
    h->pool.push_back(r.a);
    h->pool.push_back(r.b);
  }
  h->recs.clear();
  return PF_OK;
}

}  // extern "C"
