// todo.cu -- entry points declared in include/gavisunk_b200.h whose kernels are not written yet.
// They fail loudly (no CPU fallback).  Entries move out of this file as stages land.
#include "common.cuh"
#define NOTYET(name) return gvs_fail(ctx, GVS_E_STATE, name ": not implemented yet")
extern "C" {
int gvs_group_hist(gvs_ctx* ctx, int, int32_t**) { NOTYET("gvs_group_hist"); }
int gvs_hist_mode(gvs_ctx* ctx, int64_t*) { NOTYET("gvs_hist_mode"); }
int gvs_bad_groups(gvs_ctx* ctx, const int64_t*, uint64_t*) { NOTYET("gvs_bad_groups"); }
int gvs_bad_get(gvs_ctx* ctx, uint32_t*) { NOTYET("gvs_bad_get"); }
int gvs_validate(gvs_ctx* ctx, uint32_t, uint64_t*) { NOTYET("gvs_validate"); }
int gvs_pairs_get(gvs_ctx* ctx, uint32_t*, uint32_t*, uint32_t*, uint32_t*) { NOTYET("gvs_pairs_get"); }
int gvs_components_local(gvs_ctx* ctx, int, uint32_t**) { NOTYET("gvs_components_local"); }
int gvs_components_merge(gvs_ctx* ctx, const uint32_t*) { NOTYET("gvs_components_merge"); }
int gvs_intervals(gvs_ctx* ctx, uint64_t*) { NOTYET("gvs_intervals"); }
int gvs_intervals_get(gvs_ctx* ctx, uint32_t*, uint32_t*, uint32_t*) { NOTYET("gvs_intervals_get"); }
int gvs_gaps(gvs_ctx* ctx, const uint32_t*, uint64_t*, uint64_t*) { NOTYET("gvs_gaps"); }
int gvs_gaps_get(gvs_ctx* ctx, uint32_t*, uint32_t*, uint32_t*, uint32_t*) { NOTYET("gvs_gaps_get"); }
int gvs_covprob_table(gvs_ctx* ctx, const int64_t*, const int64_t*, uint32_t, double, double, double*) { NOTYET("gvs_covprob_table"); }
}
