// fused_scan.cu — single-pass decode+filter+aggregate fast path.  (placeholder: always defers to the general pipeline)
#include "fused_scan.h"

namespace horae {
namespace fused {
int try_scan_aggregate(hg_engine*, const hg_schema_desc*, const hg_sst_desc*, size_t, const hg_predicate*, size_t,
                       const hg_agg_spec*, AggBuffers*) {
  return NOT_APPLICABLE;
}
}  // namespace fused
}  // namespace horae
