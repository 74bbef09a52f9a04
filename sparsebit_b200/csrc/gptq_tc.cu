// gptq_tc.cu -- tcgen05 tensor-core path of the GPTQ int4 dequant-matmul (placeholder until the
// kernel lands: reports "unsupported" so the dispatcher uses the SIMT path).
#include "common.cuh"
namespace sb200 {
bool gptq4_tc_supported(const float*, const int32_t*, const float*, long long, long long, long long, long long, int) {
  return false;
}
size_t gptq4_tc_workspace(long long, long long, long long, int) { return 0; }
int gptq4_tc(const float*, const int32_t*, float*, const float*, const float*, long long, long long, long long,
             long long, int, void*, size_t, cudaStream_t) {
  set_error("gptq4_tc: not built");
  return SB200_E_UNSUPPORTED;
}
}  // namespace sb200
