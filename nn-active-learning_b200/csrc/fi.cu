// Fisher-information scoring kernels -- placeholder until fi.cu lands (entry points fail loudly).
#include "nnal_common.cuh"
#include "../../include/nnal_b200.h"
int nnal_fi_release(nnal_ctx*) { return NNAL_OK; }
#define FI_STUB(ctx) do { if (!(ctx)) return NNAL_ERR_INVALID; (ctx)->err = "FI path not built yet"; return NNAL_ERR_UNSUPPORTED; } while (0)
extern "C" int nnal_fi_set_candidates(nnal_ctx* ctx, const int64_t*, int64_t, int) { FI_STUB(ctx); }
extern "C" int nnal_fi_gram(nnal_ctx* ctx, const double*, float*) { FI_STUB(ctx); }
extern "C" void* nnal_fi_gram_ptr(nnal_ctx*, int64_t* n) { if (n) *n = 0; return nullptr; }
extern "C" int nnal_fi_greedy(nnal_ctx* ctx, int64_t, double, int64_t*, double*, double*) { FI_STUB(ctx); }
extern "C" int nnal_fi_step_local_best(nnal_ctx* ctx, int64_t, double, double*, int64_t*) { FI_STUB(ctx); }
extern "C" int nnal_fi_winner_factors(nnal_ctx* ctx, int64_t, float*, int64_t*) { FI_STUB(ctx); }
extern "C" int nnal_fi_step_apply(nnal_ctx* ctx, int64_t, const float*, int64_t, int, int64_t) { FI_STUB(ctx); }
