// matmul_tc.cu -- tcgen05 / TMEM / TMA GEMM (placeholder until the kernel lands).
#include "common.cuh"
#include "matmul.cuh"
namespace sk {
bool tc_supported(const GemmProblem &, int) { return false; }
bool tc_profitable(const GemmProblem &) { return false; }
int launch_gemm_tc(const GemmProblem &, int) {
  set_error("tcgen05 GEMM not built");
  return SK_ERR_UNSUPPORTED;
}
}  // namespace sk
