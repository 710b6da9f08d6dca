// fuzz_parquet_meta.cpp — TEST INFRASTRUCTURE.  Damaged SSTs through the library's footer / page-header reader
// (horaedb_b200/csrc/parquet_meta.cpp + inspect.cpp, compiled straight into this binary with AddressSanitizer and UBSan by
// tests/test_host_parquet_meta.py).  Every case is an exact-size heap copy of a valid SST with a few bytes overwritten / bits flipped
// (mostly in the footer, some anywhere: page headers) or cut to a prefix / suffix, so a read one byte past the end is a sanitizer
// report.  Pass = no report and every call returns (0 or an error code).   usage: fuzz_parquet_meta FILE ITERATIONS SEED
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../include/horae_gpu.h"

int set_error(int code, const std::string&) { return code; }      // (the library's own lives in engine.cu)

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<uint8_t> base(size_t(1) << 26);
  base.resize(fread(base.data(), 1, base.size(), f));
  fclose(f);
  if (base.size() < 12) return 2;
  const int iters = atoi(argv[2]);
  std::mt19937_64 rng(uint64_t(atoll(argv[3])));
  uint32_t flen;
  memcpy(&flen, base.data() + base.size() - 8, 4);
  long accepted = 0, rejected = 0;
  for (int it = 0; it < iters; it++) {
    size_t n = base.size();
    const bool cut = rng() % 8 == 0;
    if (cut) n = size_t(rng() % base.size()) + 1;
    uint8_t* buf = static_cast<uint8_t*>(malloc(n));
    memcpy(buf, base.data() + (base.size() - n) * size_t(rng() % 2), n);
    const int k = 1 + int(rng() % 4);
    for (int j = 0; j < k; j++) {
      size_t p;
      if (rng() % 10 < 6 && n > size_t(flen) + 8) p = n - 8 - flen + size_t(rng() % (flen + 8));
      else p = size_t(rng() % n);
      if (rng() % 2) buf[p] = uint8_t(rng()); else buf[p] ^= uint8_t(1u << (rng() % 8));
    }
    hg_parquet_summary s;
    if (hg_parquet_inspect(buf, n, &s) == 0) {
      accepted++;
      hg_parquet_chunk c;
      for (uint32_t g = 0; g < 5; g++)
        for (uint32_t col = 0; col < 8; col++) (void)hg_parquet_chunk_info(buf, n, g, col, &c);
    } else rejected++;
    free(buf);
  }
  printf("accepted %ld rejected %ld\n", accepted, rejected);
  return 0;
}
