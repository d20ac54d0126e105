// TEMPORARY during bring-up: entry points not implemented yet return KODB_ERR_UNSUPPORTED.
#include "kodb_common.cuh"
extern "C" int kodb_decode(const kodb_head_level*, const kodb_level_cfg*, int32_t, int32_t, float*, void*) { return KODB_ERR_UNSUPPORTED; }
extern "C" int64_t kodb_eval_workspace_bytes(int32_t, int64_t, int32_t) { return 0; }
extern "C" int kodb_nms(const float*, int32_t, int64_t, const kodb_nms_cfg*, const kodb_nms_out*, const kodb_targets*, const kodb_match_out*, void*, int64_t, void*) { return KODB_ERR_UNSUPPORTED; }
extern "C" int kodb_eval_tail(const kodb_head_level*, const kodb_level_cfg*, int32_t, const kodb_nms_cfg*, const kodb_nms_out*, const kodb_targets*, const kodb_match_out*, void*, int64_t, void*) { return KODB_ERR_UNSUPPORTED; }
extern "C" int kodb_map_match(const float*, const int32_t*, int32_t, int32_t, int32_t, const kodb_targets*, const kodb_match_out*, void*) { return KODB_ERR_UNSUPPORTED; }
extern "C" int64_t kodb_ap_workspace_bytes(int64_t, int32_t, int32_t) { return 0; }
extern "C" int kodb_ap_accumulate(const float*, const int32_t*, const uint8_t*, const int32_t*, int64_t, int32_t, int32_t, const int64_t*, double*, void*, int64_t, void*) { return KODB_ERR_UNSUPPORTED; }
