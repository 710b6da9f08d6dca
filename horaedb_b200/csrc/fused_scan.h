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

// Data-driven row-group pruning for transient (host-resident) SSTs: evaluates the conjunction of the predicates on ONE
// column over the PLAIN values of n row groups (already on the device) and writes flags[i] = 1 iff some row passes.
// The filter runs before merge/dedup (read.rs:459-480), so a row group without a passing row contributes nothing and
// its other columns never have to cross PCIe.  Launched on the engine's stream.
struct GateRg {
  const uint8_t* vals;       // PLAIN values of the gate column — or, with `prefixed`, the page body [u32 len][levels][values]
  uint32_t nrows, prefixed;  //   (a Snappy page decompressed on the device: the level length is only known there)
};
// first / last = the first and the last row (0-based) that pass; first > last: no row passes.  mask: bit b set = a row of
// block b passes, blocks of gate_block_rows(nrows) consecutive rows (32 blocks cover the row group)
struct GateOut { uint32_t first, last, mask; };
inline __host__ __device__ uint32_t gate_block_rows(uint32_t nrows) { const uint32_t b = (nrows + 31u) / 32u; return b < 32u ? 32u : b; }
int gate_row_groups(hg_engine* e, const GateRg* d_rgs, uint32_t n, uint32_t type, const hg_predicate* preds, size_t np, GateOut* d_out);
}  // namespace fused
}  // namespace horae
