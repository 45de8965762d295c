// placeholder: tcgen05 path (replaced by the real implementation)
#include "common.cuh"
namespace cpgb {
bool tc_eligible(const cpgb_conv_desc &, int) { return false; }
size_t tc_workspace_bytes(const cpgb_conv_desc &) { return 0; }
int tc_fprop(const cpgb_conv_desc &, const float *, const float *, const float *, const float *, float *, float, void *, size_t, cudaStream_t) { return CPGB_ENOTELIGIBLE; }
int tc_dgrad(const cpgb_conv_desc &, const float *, const float *, const float *, float *, float, void *, size_t, cudaStream_t) { return CPGB_ENOTELIGIBLE; }
int tc_wgrad_raw(const cpgb_conv_desc &, const float *, const float *, float *, void *, size_t, cudaStream_t) { return CPGB_ENOTELIGIBLE; }
}
