// umma_mn_probe.cu -- pins the tcgen05 operand forms the attention kernels rely on (run on a B200):
//   test 1  "QK":  D[128 x 16] = A[128 tok x 64 ch] . B[16 x 64]^T       A, B K-major SWIZZLE_128B, N = 16
//   test 2  "S":   D[128 x 16] = A^T . B^T with the contraction over TOKENS:
//                  A = the same token-major image read as an MN-major operand, M = 128 = [64 ch of image 0 | 64 ch
//                  of image 1] (LBO = distance of the two images), K = tokens (SBO = 1024 B per 8 tokens);
//                  B[16 x K tokens] K-major, two 64-token atoms.
// Integer-valued data: every product and sum is exact, so the check is bit-exact.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_mn_probe umma_mn_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../phyloformer_b200/csrc/pf_ffn_tc.cuh"

#define P_OFF_A0 0          // [128 tok][64 ch] bf16, SW128 K-major image (16 KB)
#define P_OFF_A1 16384      // second image
#define P_OFF_BQ 32768      // [16][64] bf16 K-major (2 KB)
#define P_OFF_BK 34816      // [16][128 tok] bf16 K-major, 2 atoms x 2 KB
#define P_OFF_BAR 38912
#define P_OFF_TM 38928
#define P_SMEM (39936 + 1024)

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// variant: 0 = (LBO = image distance, SBO = 1024), 1 = swapped
__global__ void __launch_bounds__(128, 1)
k_probe(const uint16_t* __restrict__ img, int test, int variant, float* __restrict__ out, int* __restrict__ err) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 38912 / 16; i += 128) reinterpret_cast<int4*>(sm)[i] = reinterpret_cast<const int4*>(img)[i];
  const uint32_t bar = sbase + P_OFF_BAR;
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + P_OFF_TM), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + P_OFF_TM);
  if (tid == 0) {
    if (test == 1) {
      const uint32_t idesc = umma_idesc(128, 16);
      for (int s = 0; s < 4; ++s)
        umma_ss(tmem, umma_desc(sbase + P_OFF_A0) + 2 * s, umma_desc(sbase + P_OFF_BQ) + 2 * s, idesc, s ? 1u : 0u);
    } else {
      const uint32_t idesc = umma_idesc(128, 16) | (1u << 15);   // A is MN-major
      const uint32_t lbo = variant == 0 ? 16384u : 1024u, sbo = variant == 0 ? 1024u : 16384u;
      for (int s = 0; s < 8; ++s) {   // 16 tokens per step
        const uint64_t da = desc_mn(sbase + P_OFF_A0 + 2048 * s, lbo, sbo);
        const uint64_t db = umma_desc(sbase + P_OFF_BK + (s >> 2) * 2048 + (s & 3) * 32);
        umma_ss(tmem, da, db, idesc, s ? 1u : 0u);
      }
    }
    tc_commit(bar);
  }
  const bool ok = mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t v[16];
  tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
  tc_wait_ld();
  for (int i = 0; i < 16; ++i) out[tid * 16 + i] = __uint_as_float(v[i]);
  if (!ok) *err = 1;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}

int main() {
  std::vector<uint16_t> img(38912 / 2, 0);
  std::vector<float> A0(128 * 64), A1(128 * 64), BQ(16 * 64), BK(16 * 128);
  srand(7);
  auto rnd = [] { return (float)((rand() % 9) - 4); };
  for (auto& v : A0) v = rnd();
  for (auto& v : A1) v = rnd();
  for (auto& v : BQ) v = rnd();
  for (auto& v : BK) v = rnd();
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 64; ++c) {
      img[(P_OFF_A0 + umma_off_k64(r, c)) / 2] = f32_to_bf16_rn(A0[r * 64 + c]);
      img[(P_OFF_A1 + umma_off_k64(r, c)) / 2] = f32_to_bf16_rn(A1[r * 64 + c]);
    }
  for (int n = 0; n < 16; ++n) {
    for (int k = 0; k < 64; ++k) img[(P_OFF_BQ + umma_off_k64(n, k)) / 2] = f32_to_bf16_rn(BQ[n * 64 + k]);
    for (int k = 0; k < 128; ++k) img[(P_OFF_BK + umma_off_k256(n, k, 16)) / 2] = f32_to_bf16_rn(BK[n * 128 + k]);
  }
  uint16_t* d_img; float* d_out; int* d_err;
  cudaMalloc(&d_img, 38912); cudaMalloc(&d_out, 128 * 16 * 4); cudaMalloc(&d_err, 4);
  cudaMemcpy(d_img, img.data(), 38912, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM);
  int fails = 0;
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int test = cfg == 0 ? 1 : 2, variant = cfg == 2 ? 1 : 0;
    cudaMemset(d_out, 0xff, 128 * 16 * 4); cudaMemset(d_err, 0, 4);
    k_probe<<<1, 128, P_SMEM>>>(d_img, test, variant, d_out, d_err);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(128 * 16); int err = 0;
    cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost);
    int bad = 0; float first_got = 0, first_want = 0; int first = -1;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 16; ++n) {
        double want = 0;
        if (test == 1) for (int c = 0; c < 64; ++c) want += (double)A0[m * 64 + c] * BQ[n * 64 + c];
        else for (int t = 0; t < 128; ++t) want += (double)(m < 64 ? A0[t * 64 + m] : A1[t * 64 + (m - 64)]) * BK[n * 128 + t];
        if ((double)out[m * 16 + n] != want) { if (first < 0) { first = m * 16 + n; first_got = out[m * 16 + n]; first_want = (float)want; } ++bad; }
      }
    printf("probe test=%d variant=%d: cuda=%s barrier_timeout=%d mismatches=%d/2048", test, variant, cudaGetErrorString(e), err, bad);
    if (bad) printf("  first at [%d][%d] got %g want %g", first / 16, first % 16, first_got, first_want);
    printf("  %s\n", bad == 0 && e == cudaSuccess && !err ? "PASS" : "FAIL");
    if (cfg < 2 && (bad || e != cudaSuccess)) ++fails;
    if (e != cudaSuccess) break;
  }
  return fails;
}
