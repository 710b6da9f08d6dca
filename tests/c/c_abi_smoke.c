/* c_abi_smoke.c — binds include/horae_gpu.h from plain C, the way the Rust FFI (bindgen / extern "C") would:
 * no C++ types, no Python.  Usage: c_abi_smoke <sst file> <num columns incl. builtin> <num pk>
 * Schema assumed: the metric schema of SURVEY 8 (series_id u64, ts i64, value f64, tag u32, __seq__ u64, __reserved__ u64).
 * Exit code 0 = every check passed.  Without a GPU the engine must fail loudly (HG_ERR_CUDA), never fall back. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "horae_gpu.h"

#define CHECK(cond, msg) do { if (!(cond)) { fprintf(stderr, "FAIL %s:%d %s (%s)\n", __FILE__, __LINE__, msg, hg_last_error()); return 1; } } while (0)

_Static_assert(sizeof(hg_predicate) == 48, "hg_predicate layout");
_Static_assert(sizeof(hg_sst_desc) == 64, "hg_sst_desc layout");
_Static_assert(sizeof(hg_agg_spec) == 24, "hg_agg_spec layout");
_Static_assert(sizeof(hg_schema_desc) == 32, "hg_schema_desc layout");

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s <sst>\n", argv[0]); return 2; }
  CHECK(hg_abi_version() == HG_ABI_VERSION, "ABI version");
  FILE* f = fopen(argv[1], "rb");
  CHECK(f != NULL, "open sst");
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  uint8_t* data = (uint8_t*)malloc((size_t)n);
  CHECK(fread(data, 1, (size_t)n, f) == (size_t)n, "read sst");
  fclose(f);

  /* host-only entry points */
  hg_parquet_summary sum;
  CHECK(hg_parquet_inspect(data, (uint64_t)n, &sum) == HG_OK, "inspect");
  CHECK(sum.num_columns == 6 && sum.num_rows > 0, "summary");
  uint32_t types[6] = {HG_U64, HG_I64, HG_F64, HG_U32, HG_U64, HG_U64};
  const char* names[6] = {"series_id", "ts", "value", "tag", "__seq__", "__reserved__"};
  hg_schema_desc schema = {6, 2, HG_UPDATE_OVERWRITE, 0, types, names};
  hg_predicate pred;
  memset(&pred, 0, sizeof pred);
  pred.column = 3; pred.op = HG_OP_EQ; pred.u64 = 3;
  uint8_t keep[4096];
  uint32_t nrg = 0;
  CHECK(hg_plan_row_groups(&schema, data, (uint64_t)n, &pred, 1, keep, 4096, &nrg) == HG_OK && nrg == sum.num_row_groups, "plan_row_groups");
  CHECK(hg_parquet_inspect(data, 10, &sum) == HG_ERR_FORMAT, "truncated file must be HG_ERR_FORMAT");
  CHECK(hg_parquet_inspect(NULL, 0, &sum) == HG_ERR_INVALID, "null argument");

  hg_config cfg = {0, 0, 0, 0, 0};
  hg_engine* e = NULL;
  int rc = hg_engine_create(&cfg, &e);
  if (rc != HG_OK) {
    CHECK(rc == HG_ERR_CUDA && strstr(hg_last_error(), "no CPU fallback") != NULL, "engine creation without a GPU must fail loudly");
    printf("c_abi_smoke: host-only checks ok (no GPU: %s)\n", hg_last_error());
    return 0;
  }
  /* GPU: resident load, aggregate through an Arrow C stream, compaction stream */
  hg_sst_desc sst;
  memset(&sst, 0, sizeof sst);
  sst.id = 7; sst.data = data; sst.size = (uint64_t)n; sst.num_rows = (uint32_t)sum.num_rows;
  CHECK(hg_sst_load(e, &schema, &sst) == HG_OK, "sst_load");
  hg_sst_desc res = sst;
  res.data = NULL; res.size = 0;
  hg_agg_spec agg = {0, -1, 0, 2, HG_AGG_RUNS};
  struct ArrowArrayStream st;
  CHECK(hg_scan_aggregate(e, &schema, &res, 1, &pred, 1, &agg, &st) == HG_OK, "scan_aggregate");
  struct ArrowSchema as;
  CHECK(st.get_schema(&st, &as) == 0 && as.n_children == 5, "aggregate schema (series_id, count, sum, min, max)");
  struct ArrowArray batch;
  int64_t groups = 0;
  for (;;) {
    CHECK(st.get_next(&st, &batch) == 0, "get_next");
    if (!batch.release) break;
    groups += batch.length;
    const uint64_t* cnt = (const uint64_t*)batch.children[1]->buffers[1] + batch.children[1]->offset;
    for (int64_t i = 0; i < batch.length; i++) CHECK(cnt[i] > 0, "empty group");
    batch.release(&batch);
  }
  as.release(&as);
  st.release(&st);
  hg_scan_stats stats;
  CHECK(hg_last_stats(e, &stats) == HG_OK && stats.groups_out == (uint64_t)groups && stats.rows_in_files == sum.num_rows, "stats");
  /* a request the reference rejects is an error code, not a fallback: Append merges Binary value columns only (operator.rs:66-73) */
  hg_schema_desc append = schema;
  append.update_mode = HG_UPDATE_APPEND;
  CHECK(hg_scan_open(e, &append, &res, 1, NULL, 0, NULL, 0, 0, &st) == HG_ERR_INVALID && strstr(hg_last_error(), "binary column") != NULL,
        "Append mode over non-Binary value columns must be HG_ERR_INVALID");
  CHECK(hg_compact_open(e, &schema, &res, 1, &st) == HG_OK, "compact_open");
  int64_t rows = 0;
  for (;;) {
    CHECK(st.get_next(&st, &batch) == 0, "get_next");
    if (!batch.release) break;
    CHECK(batch.n_children == 6, "compaction keeps the builtin columns");
    rows += batch.length;
    batch.release(&batch);
  }
  st.release(&st);
  CHECK(rows == (int64_t)sum.num_rows, "single-file compaction returns every row");
  CHECK(hg_sst_unload(e, 7) == HG_OK, "unload");
  hg_engine_destroy(e);
  free(data);
  printf("c_abi_smoke: ok, %lld groups, %lld rows\n", (long long)groups, (long long)rows);
  return 0;
}
