// Entry points not implemented yet fail loudly (no fallback).
#include "common.cuh"
#define NOTYET(ctx) do { if (!(ctx)) return EPIVO_ERR_INVALID; EPV_FAIL(ctx, EPIVO_ERR_UNSUPPORTED, "%s: not implemented yet", __func__); } while (0)
extern "C" {
int epivo_find_essential(epivo_ctx* ctx, const float*, const float*, int, const double*, int, double, double, int, const int32_t*, int, double*, uint8_t*, int*, int*) { NOTYET(ctx); }
int epivo_five_point(epivo_ctx* ctx, const double*, const double*, int, double*, int32_t*) { NOTYET(ctx); }
int epivo_score_sampson(epivo_ctx* ctx, const double*, int, const float*, const float*, int, const double*, double, int32_t*, float*, int*, uint8_t*) { NOTYET(ctx); }
int epivo_recover_pose(epivo_ctx* ctx, const double*, const float*, const float*, int, const double*, double, const uint8_t*, double*, double*, uint8_t*, int*) { NOTYET(ctx); }
int epivo_lm_rt(epivo_ctx* ctx, int, double, const int32_t*, const double*, int, double, int, double, double*, const double*, const double*, int, epivo_lm_res*, int*) { NOTYET(ctx); }
int epivo_lm_rt_batch(epivo_ctx* ctx, int, int, double, const int32_t*, const double*, int, double, int, double, double*, const double*, const double*, int, epivo_lm_res*, int32_t*) { NOTYET(ctx); }
void epivo_pipeline_params_default(epivo_pipeline_params*) {}
int epivo_seq_create(epivo_ctx* ctx, epivo_seq**, int, int) { NOTYET(ctx); }
void epivo_seq_destroy(epivo_seq*) {}
int epivo_seq_upload(epivo_seq*, int, int, const float*, const uint8_t*) { return EPIVO_ERR_UNSUPPORTED; }
int epivo_seq_run(epivo_seq*, const epivo_pipeline_params*, int, int) { return EPIVO_ERR_UNSUPPORTED; }
int epivo_seq_download(epivo_seq*, epivo_pair_result*, int, int) { return EPIVO_ERR_UNSUPPORTED; }
int epivo_seq_stage_ms(epivo_seq*, float*, int) { return EPIVO_ERR_UNSUPPORTED; }
int epivo_seq_get_matches(epivo_seq*, int, int32_t*, int32_t*, int32_t*, int*) { return EPIVO_ERR_UNSUPPORTED; }
int epivo_seq_get_masks(epivo_seq*, int, uint8_t*, int*, uint8_t*, int*) { return EPIVO_ERR_UNSUPPORTED; }
}
