/* Plain-C caller of libpf_sm100.so: no Python, no torch -- only include/pf_sm100.h and the CUDA
 * runtime for device memory.  This is what a non-Python host (or the reference's maintainers,
 * through any FFI) links against.  Used by tests/test_gpu_cabi_c.py, which compares the output
 * with the Python path and the oracle, and by tests/test_host_cpu.py (compile-only: the header
 * must be valid C99).
 *
 *   cabi_demo weights.bin msa.bin out.bin [precision]
 *
 * weights.bin: int32 count, then per tensor { int64 numel, float[numel] } in pf_create's order.
 * msa.bin:     int32 B, n, L, then uint8[B*n*L] residue codes.
 * out.bin:     float[B * n(n-1)/2] distances.
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "pf_sm100.h"

#define CHECK_CUDA(x)                                                        \
  do {                                                                       \
    cudaError_t e_ = (x);                                                    \
    if (e_ != cudaSuccess) {                                                 \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));               \
      return 2;                                                              \
    }                                                                        \
  } while (0)
#define CHECK_PF(x)                                                          \
  do {                                                                       \
    int rc_ = (x);                                                           \
    if (rc_ != PF_OK) {                                                      \
      fprintf(stderr, "%s -> %d: %s\n", #x, rc_, pf_last_error());           \
      return 3;                                                              \
    }                                                                        \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s weights.bin msa.bin out.bin [precision]\n", argv[0]);
    return 1;
  }
  const int precision = argc > 4 ? atoi(argv[4]) : PF_PREC_BF16X3;
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  int32_t count = 0;
  if (fread(&count, 4, 1, f) != 1 || count < 1 || count > 4096) { fprintf(stderr, "bad weights file\n"); return 1; }
  const float** dev = (const float**)calloc((size_t)count, sizeof(float*));
  for (int i = 0; i < count; ++i) {
    int64_t numel = 0;
    if (fread(&numel, 8, 1, f) != 1 || numel < 1) { fprintf(stderr, "bad tensor %d\n", i); return 1; }
    float* host = (float*)malloc((size_t)numel * 4);
    if (fread(host, 4, (size_t)numel, f) != (size_t)numel) { fprintf(stderr, "short tensor %d\n", i); return 1; }
    float* d = NULL;
    CHECK_CUDA(cudaMalloc((void**)&d, (size_t)numel * 4));
    CHECK_CUDA(cudaMemcpy(d, host, (size_t)numel * 4, cudaMemcpyHostToDevice));
    dev[i] = d;
    free(host);
  }
  fclose(f);

  f = fopen(argv[2], "rb");
  if (!f) { perror(argv[2]); return 1; }
  int32_t dims[3];
  if (fread(dims, 4, 3, f) != 3) { fprintf(stderr, "bad msa file\n"); return 1; }
  const int B = dims[0], n = dims[1], L = dims[2];
  const size_t cells = (size_t)B * n * L;
  uint8_t* msa = (uint8_t*)malloc(cells);
  if (fread(msa, 1, cells, f) != cells) { fprintf(stderr, "short msa file\n"); return 1; }
  fclose(f);

  pf_cfg cfg;
  cfg.nb_blocks = (count - 4) / 26;   /* 2 embedding + 26 per block + 2 head tensors */
  cfg.nb_heads = 4;
  cfg.embed_dim = 64;
  cfg.ffn_mult = 4;
  cfg.precision = precision;
  pf_handle h = NULL;
  CHECK_PF(pf_create(&h, &cfg, dev, count));

  const int64_t P = (int64_t)n * (n - 1) / 2;
  const size_t ws_bytes = pf_workspace_bytes(h, B, n, L, 0, P);
  uint8_t* msa_dev = NULL;
  float* dist_dev = NULL;
  void* ws = NULL;
  CHECK_CUDA(cudaMalloc((void**)&msa_dev, cells));
  CHECK_CUDA(cudaMalloc((void**)&dist_dev, (size_t)B * (size_t)P * 4));
  CHECK_CUDA(cudaMalloc(&ws, ws_bytes));
  CHECK_CUDA(cudaMemcpy(msa_dev, msa, cells, cudaMemcpyHostToDevice));
  CHECK_PF(pf_forward(h, msa_dev, NULL, NULL, B, n, L, 0, P, dist_dev, ws, ws_bytes, NULL /* default stream */,
                      NULL, NULL));
  CHECK_PF(pf_device_error(h));
  float* dist = (float*)malloc((size_t)B * (size_t)P * 4);
  CHECK_CUDA(cudaMemcpy(dist, dist_dev, (size_t)B * (size_t)P * 4, cudaMemcpyDeviceToHost));
  f = fopen(argv[3], "wb");
  if (!f) { perror(argv[3]); return 1; }
  fwrite(dist, 4, (size_t)B * (size_t)P, f);
  fclose(f);
  printf("abi %d, %d blocks, B=%d n=%d L=%d, %d launches, d[0]=%.8f\n", pf_abi_version(), cfg.nb_blocks, B, n, L,
         pf_last_launch_count(h), dist[0]);
  pf_destroy(h);
  return 0;
}
