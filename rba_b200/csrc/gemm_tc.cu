// tcgen05 bf16x3 GEMM backend (placeholder until the TMA/TMEM kernel lands; the entry points fail loudly).
#include "common.cuh"

namespace rba {
int gemm_tc_launch(const rba_gemm_args& a, cudaStream_t st) {
  (void)a; (void)st;
  return fail(RBA_ERR_STATE, "gemm: RBA_GEMM_TC backend not built in this revision");
}
int conv3x3_tc_launch(const uint16_t*, const uint16_t*, const uint16_t*, const uint16_t*, int, int, int, int, int, float*,
                      cudaStream_t) {
  return fail(RBA_ERR_STATE, "conv3x3: RBA_GEMM_TC backend not built in this revision");
}
}  // namespace rba
