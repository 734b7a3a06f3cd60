// placeholder until the BA kernels land (next commit)
#include "ctx.h"
int flv_ba_free(flv_ctx* ctx) { if (ctx->ba_ws) cudaFree(ctx->ba_ws); ctx->ba_ws = nullptr; return FLV_OK; }
extern "C" {
int flv_ba_reserve(flv_ctx* ctx, int, int, int) { if (!ctx) return FLV_ERR_INVALID; FLV_FAIL(ctx, FLV_ERR_UNSUPPORTED, "BA not built yet"); }
int flv_ba_optimize(flv_ctx* ctx, int, const flv_ba_problem*, const flv_ba_params*, double*, double*, const int*, const int*, const double*, uint8_t*, flv_ba_stats*, flv_memspace) { if (!ctx) return FLV_ERR_INVALID; FLV_FAIL(ctx, FLV_ERR_UNSUPPORTED, "BA not built yet"); }
}
