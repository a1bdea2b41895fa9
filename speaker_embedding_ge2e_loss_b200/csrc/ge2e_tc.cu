// tcgen05 / TMA / TMEM path (TF32 operands, fp32 accumulators in tensor memory).
#include "ge2e_common.cuh"

namespace ge2e {

bool tc_supported(int, int, int, int, int) { return false; }
size_t tc_workspace_bytes(int, int, int, int, int) { return 0; }
int tc_fwd_rows(const RowsArgs&, float*, int32_t*, float*, float*, void*, size_t, cudaStream_t) {
  return GE2E_ERR_UNSUPPORTED;
}
int tc_bwd_rows(const RowsArgs&, const float*, const int32_t*, const float*, float*, float*, float*,
                void*, size_t, cudaStream_t) {
  return GE2E_ERR_UNSUPPORTED;
}

}  // namespace ge2e
