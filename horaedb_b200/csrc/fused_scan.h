// fused_scan.h — single-pass decode+filter+aggregate fast path (fused_scan.cu).
#pragma once
#include "engine_internal.h"

namespace horae {
namespace fused {
constexpr int NOT_APPLICABLE = -1000;
// Returns NOT_APPLICABLE when the inputs do not satisfy the fast path's preconditions (the general pipeline then
// runs), otherwise an hg_status.
int try_scan_aggregate(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* ssts, size_t n, const hg_predicate* preds,
                       size_t np, const hg_agg_spec* agg, AggBuffers* out);
}  // namespace fused
}  // namespace horae
