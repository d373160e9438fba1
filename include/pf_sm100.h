/* pf_sm100.h -- C ABI of libpf_sm100.so, the sm_100a implementation of the Phyloformer
 * inference hot path (MSA -> pairwise evolutionary distances).
 *
 * The reference (lucanest/Phyloformer) has no FFI layer: its boundary for this path is the
 * Python surface `phyloformer.model.Phyloformer.forward(x)` (reference phyloformer/model.py:
 * 166-187).  `phyloformer_b200/model.py` keeps that surface and calls the entry points below
 * through ctypes; INTEGRATION.md shows the binding.  Each entry point names the reference
 * code it replaces.
 *
 * Conventions
 *   - plain C: pointers + sizes only, no C++/torch types; nothing throws across the ABI.
 *   - every `*_dev` pointer is CALLER-OWNED DEVICE memory on the current CUDA device; the
 *     library allocates only its packed copy of the weights (pf_create / pf_destroy).
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no hidden
 *     synchronisation (pf_create is the one exception: it copies and repacks the weights).
 *   - return value 0 = success, negative = error; pf_last_error() gives the message
 *     (thread-local).
 *   - pair order is the reference's lexicographic (i<j) order (model.py:13-17); a shard
 *     owns the contiguous pair range [pair_lo, pair_hi).
 */
#ifndef PF_SM100_H
#define PF_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_ABI_VERSION 1

/* arithmetic of the FFN contraction (the only dense GEMM on the path) */
#define PF_PREC_FP32   0 /* fp32 FFMA everywhere ("exact" mode, ~1e-6 of the reference)      */
#define PF_PREC_BF16X3 1 /* tcgen05 bf16 MMAs, 3-term hi/lo split, fp32 accumulate in TMEM   */
#define PF_PREC_BF16   2 /* single bf16 tcgen05 pass (fast mode, reported separately)        */
#define PF_PREC_FP16   3 /* fp16 activations (one rounding) x fp16 hi+lo weights, 2 passes   */

#define PF_OK              0
#define PF_ERR_ARG        -1
#define PF_ERR_CUDA       -2
#define PF_ERR_WORKSPACE  -3
#define PF_ERR_NO_DEVICE  -4 /* no sm_100 device: there is no CPU fallback */
#define PF_ERR_REDUCE     -5
#define PF_ERR_FASTA_RESIDUE  -10 /* residue outside ALPHABET: the reference raises KeyError (data.py:26) */
#define PF_ERR_FASTA_RAGGED   -11 /* unequal sequence lengths: the reference's torch.tensor() raises ValueError */
#define PF_ERR_FASTA_NOHEADER -12 /* sequence data before the first '>' (reference: IndexError, data.py:26)  */

/* number of fp32 values exchanged per (MSA, site) per block between pair shards:
 * sum_p k~ (4), sum_p q~ (4), sum_p k~ * v (64)    (attention.py:179-190 before normalising) */
#define PF_COLSUM_FLOATS 72
/* number of weight tensors, in reference state-dict order */
#define PF_N_WEIGHTS(nb) (2 + 26 * (nb) + 2)

typedef struct pf_ctx* pf_handle;

typedef struct {
  int32_t nb_blocks; /* 6  */
  int32_t nb_heads;  /* 4  (only 4 is built)  */
  int32_t embed_dim; /* 64 (only 64 is built) */
  int32_t ffn_mult;  /* 4  */
  int32_t precision; /* PF_PREC_* */
} pf_cfg;

/* Sums `count` fp32 values at `buf_dev` in place over all pair shards, enqueued on `stream`.
 * Called once per block (column attention, reference model.py:97) when a forward covers only
 * part of the pair list.  Return 0 on success. */
typedef int (*pf_reduce_fn)(void* user, float* buf_dev, size_t count, void* stream);

/* Replaces Phyloformer.__init__ + load_state_dict (model.py:112-164, infer_alns.py:71-86):
 * takes the 160 fp32 weight tensors as device pointers in state-dict order
 *   embedding_block.0.{weight,bias};
 *   attention_blocks.i.{row_attention,col_attention}.{k_proj,q_proj,v_proj,out_proj}.{weight,bias},
 *   attention_blocks.i.{row_norm,col_norm,ffn_norm}.{weight,bias},
 *   attention_blocks.i.ffn.{0,3}.{weight,bias};   pwFNN.0.{weight,bias}
 * and builds the packed device copy the kernels read. Synchronous. */
int pf_create(pf_handle* out, const pf_cfg* cfg, const float* const* weights_dev, int n_weights);
void pf_destroy(pf_handle h);

/* Change the FFN arithmetic of an existing handle (PF_PREC_*). */
int pf_set_precision(pf_handle h, int precision);

/* Bytes of scratch pf_forward needs for B MSAs of n taxa x L sites, pairs [pair_lo,pair_hi). */
size_t pf_workspace_bytes(pf_handle h, int B, int n, int L, int64_t pair_lo, int64_t pair_hi);

/* Replaces data.py:28-29 + infer_alns.py:112 on the device side: (B,22,L,n) fp32 one-hot ->
 * (B,n,L) uint8 residue codes.  *not_onehot_dev is set to 1 if any column is not exactly
 * one-hot (the forward then uses the general embedding path on x itself). */
int pf_onehot_to_idx(const float* x_dev, int B, int L, int n, uint8_t* idx_dev,
                     int32_t* not_onehot_dev, void* stream);

/* Replaces Phyloformer.forward (model.py:166-187) for pairs [pair_lo, pair_hi):
 *   msa_idx_dev     (B,n,L) uint8 residue codes 0..21 (ALPHABET order, data.py:7)
 *   x_dev           NULL, or the (B,22,L,n) fp32 input; used instead of msa_idx_dev when
 *                   *not_onehot_dev != 0 (soft inputs keep forward(x)'s semantics)
 *   not_onehot_dev  NULL (treated as 0) or the flag written by pf_onehot_to_idx
 *   dist_dev        (B, pair_hi-pair_lo) fp32 out: mean_l softplus(pwFNN(.)) (model.py:182-185)
 *   ws_dev/ws_bytes scratch of at least pf_workspace_bytes()
 *   reduce          NULL when the shard holds all n(n-1)/2 pairs, else the cross-shard sum */
int pf_forward(pf_handle h, const uint8_t* msa_idx_dev, const float* x_dev,
               const int32_t* not_onehot_dev, int B, int n, int L, int64_t pair_lo,
               int64_t pair_hi, float* dist_dev, void* ws_dev, size_t ws_bytes, void* stream,
               pf_reduce_fn reduce, void* reduce_user);

/* Test hook: same as pf_forward but stops after `n_stages` sub-blocks (stage 0 = pair
 * embedding, then row, col, ffn per block) and copies the (B,Pl,L,64) activation to act_dev. */
int pf_forward_debug(pf_handle h, const uint8_t* msa_idx_dev, const float* x_dev,
                     const int32_t* not_onehot_dev, int B, int n, int L, int64_t pair_lo,
                     int64_t pair_hi, float* dist_dev, void* ws_dev, size_t ws_bytes,
                     void* stream, pf_reduce_fn reduce, void* reduce_user, int n_stages,
                     float* act_dev);

/* Replaces infer_alns.py:14-25 (vec_to_phylip's triu scatter + dm + dm.T):
 * (B,P) distances -> (B,n,n) symmetric matrices with a zero diagonal. */
int pf_dist_to_matrix(const float* dist_dev, int B, int n, float* mat_dev, void* stream);

/* Replaces the parsing half of phyloformer/data.py:11-27 (load_alignment): host-only, no device
 * work.  `text[0..len)` is the FASTA file.  Lines are stripped like bytes.strip(); a line that
 * starts with '>' opens a record, every other non-empty line is residues.  Writes the residue
 * codes (index in ALPHABET = "ARNDCQEGHILKMFPSTWYVX-") row-major (n, L) into codes[0..cap), the
 * alignment length into *L_out and the byte range of every record name (text after '>') into
 * name_off/name_len (at most max_names).  Returns n, or PF_ERR_FASTA_* / PF_ERR_ARG; on
 * PF_ERR_FASTA_RESIDUE *bad_char holds the offending byte. */
long long pf_parse_fasta(const char* text, long long len, uint8_t* codes, long long cap,
                         int32_t* L_out, int64_t* name_off, int32_t* name_len, int32_t max_names,
                         int32_t* bad_char);

/* Replaces the text half of infer_alns.py:19-23 (vec_to_phylip): host-only, no device work.
 * Writes "n\n" and one line per taxon, "name v v ... v\n" with every value as "%.10f" (same
 * digits as the reference's f"{x:.10f}"), into out[0..cap).  dm_host is the (n,n) fp32 matrix in
 * host memory, names are NUL-terminated UTF-8.  Returns the text length in bytes (call again
 * with a larger buffer if it exceeds cap; nothing is NUL-terminated), or a negative PF_ERR_*. */
long long pf_format_phylip(const float* dm_host, int n, const char* const* names, char* out,
                           long long cap);

/* Replaces skbio.tree.nj behind `infer_alns.py --trees` (infer_alns.py:120-123): host-only
 * neighbour joining on the (n,n) fp32 matrix, Newick text (trifurcating root, "%.10f" branch
 * lengths, negative lengths clipped to 0) into out[0..cap).  Same contract as pf_format_phylip:
 * returns the text length, or a negative PF_ERR_*. */
long long pf_neighbor_joining(const float* dm_host, int n, const char* const* names, char* out,
                              long long cap);

/* Replaces the tree step of the reference's workflow (README.md:85-92: `fastme -i x.phy -o x.nwk --nni --spr`,
 * FastME 2.1.6.4, shipped only as bin/bin_linux/fastme): host-only.  BIONJ start tree (PF_TREE_NJ_START: plain
 * NJ), then a balanced-minimum-evolution search by NNIs (PF_TREE_NNI) and one by SPRs (PF_TREE_SPR), both from
 * the start tree, best improvement first.  Output rule as observed on the binary: without a search the start
 * tree with its own (BIONJ) branch lengths; with a search the shortest of { start tree with its own BIONJ branch
 * lengths, NNI result, SPR result }.  dm_host is the symmetric (n,n) matrix in DOUBLE precision (what FastME
 * parses from the "%.10f" PHYLIP text); Newick with "%.8f" branch lengths (FastME's default digits, negative
 * lengths kept) into out[0..cap).  stats (optional, 7 doubles): length of the start tree with its own branch
 * lengths, balanced length of the start tree, after NNI, after SPR, number of NNIs, number of SPRs, which tree
 * was kept (0 start, 1 NNI, 2 SPR).  Returns the text length (call again with a larger buffer if it exceeds
 * cap), or a negative PF_ERR_*.  n <= PF_BME_MAX_TAXA. */
#define PF_TREE_NNI 1
#define PF_TREE_SPR 2
#define PF_TREE_NJ_START 4
#define PF_BME_MAX_TAXA 1000
long long pf_bme_tree(const double* dm_host, int n, const char* const* names, int flags, char* out,
                      long long cap, double* stats);

/* Kernel launches enqueued by the last pf_forward on this handle (for bench.py's
 * gpu_launches claim). */
int pf_last_launch_count(pf_handle h);

/* Pair-sharded column attention over NVLink peer memory instead of the reduce callback.
 * Every rank allocates one symmetric buffer of pf_peer_exchange_bytes(slot_floats) bytes, zero
 * filled, and maps all ranks' buffers (CUDA IPC / torch symmetric memory); peer_bufs_host[r] is
 * rank r's buffer as addressable from THIS process.  slot_floats >= B*L*72 of the largest
 * forward.  After this call pf_forward ignores `reduce` for partial pair ranges: the library's
 * own kernels publish, synchronise (system-scope flags) and read the (B,L,72) summaries with
 * plain P2P loads, summing in rank order.  world <= 1 or NULL disables it again.
 * Every rank must issue the same sequence of forwards (the exchange epochs are counted per
 * handle from this call on: re-zero the buffers, with a barrier on either side, before binding
 * them to a new handle).  Buffer layout: 64 KB of flags, one 32-bit word per (rank, exchange CTA)
 * -- the exchange is site-chunked: a CTA that has reduced its site groups publishes its own flag
 * and waits only for the same CTA of every peer -- followed by two slots of slot_floats floats.
 * world <= 32. */
int pf_set_peer_exchange(pf_handle h, int rank, int world, void* const* peer_bufs_host, size_t slot_floats);
size_t pf_peer_exchange_bytes(size_t slot_floats);

/* Synchronises and returns PF_ERR_CUDA if a kernel raised the device-side error flag (a
 * tensor-core pipeline wait that timed out instead of hanging the GPU); clears the flag. */
int pf_device_error(pf_handle h);

/* Test hook: when non-NULL, the tcgen05 FFN kernel copies the raw fp32 accumulators of its
 * first 128-token tile to dump_dev as [128][320] (columns 0..255 = GEMM1, 256..319 = GEMM2),
 * followed by the role timers of the profiling build: 3 roles x 8 floats per CTA (one CTA per SM).
 * dump_dev must hold 128*320 + 24 * (number of SMs) floats. */
int pf_debug_set_dump(pf_handle h, float* dump_dev);

/* Per-kernel device timing (bench.py's roofline leg): when enabled, pf_forward brackets every
 * kernel launch with CUDA events on `stream`.  pf_profile_read waits for them and returns, per
 * kernel class, the summed milliseconds and the number of launches since the last read. */
#define PF_KC_INPUT 0   /* one-hot conversion / sequence embedding            */
#define PF_KC_ROW 1     /* pair embedding + row attention                      */
#define PF_KC_COLSUM 2  /* column partial sums                                 */
#define PF_KC_COLFIN 3  /* column reduce + finalize (tiny)                     */
#define PF_KC_FFN 4     /* column apply + LayerNorm + FFN + residual           */
#define PF_KC_HEAD 5    /* distance head                                       */
#define PF_KC_COUNT 6
int pf_profile_enable(pf_handle h, int on);
int pf_profile_read(pf_handle h, float* ms_out /*[PF_KC_COUNT]*/, int32_t* launches_out /*[PF_KC_COUNT]*/);

int pf_abi_version(void);
const char* pf_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* PF_SM100_H */
