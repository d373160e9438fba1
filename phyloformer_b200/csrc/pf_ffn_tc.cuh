// pf_ffn_tc.cuh -- tcgen05/TMEM implementation of column-apply + FFN (placeholder: filled in
// by the next milestone; the fp32 FFMA kernel in pf_kernels.cuh is the path until then).
#pragma once
#include "pf_common.cuh"

struct PfFfnTcW {
  float pad[4];
};
inline void pf_pack_ffn_tc(const PfFfnW&, PfFfnTcW*) {}
inline int pf_ffn_tc_init() { return 0; }
inline int pf_ffn_tc_launch(const PfAttnW*, const PfFfnTcW*, float*, const float*, int, int, long long, int, int,
                            cudaStream_t) {
  return -1;
}
