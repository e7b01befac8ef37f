// tcgen05 FC GEMM -- placeholder
#include "nnal_common.cuh"
int nnal_tc_prepare_layer(nnal_ctx*, Layer&) { return NNAL_OK; }
bool nnal_tc_fc_supported(const nnal_ctx*, const Layer&) { return false; }
int nnal_tc_fc(nnal_ctx* ctx, const Layer&, const float*, float*, int64_t) { NNAL_FAIL(ctx, NNAL_ERR_UNSUPPORTED, "tc fc not built"); }
int nnal_tc_release(nnal_ctx*) { return NNAL_OK; }
