// tools/model_ratio.cpp -- CPU-side size check of the host model (tests/model/enc_model.cpp) against libzstd, per data class.
// TEST/DEV TOOL ONLY (links the oracle's generators and libzstd).
//   g++ -O2 -std=c++17 tools/model_ratio.cpp tests/model/enc_model.cpp -o /tmp/model_ratio -L oracle -l:liboracle_zstd.so -l:libzstd.so.1 -Wl,-rpath,$PWD/oracle
//   /tmp/model_ratio <level> [chunk] [nchunks] [file ...]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {
void orc_gen_batch(uint8_t *dst, size_t U, uint64_t first_idx, size_t n, int kind, uint32_t P);
void orc_gen_textlike(uint8_t *dst, size_t size);
size_t ZSTD_compress(void *dst, size_t cap, const void *src, size_t n, int level);
size_t ZSTD_decompress(void *dst, size_t cap, const void *src, size_t n);
size_t ZSTD_compressBound(size_t n);
unsigned ZSTD_isError(size_t c);
size_t model_compress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, int level, int checksum);
}

int main(int argc, char **argv) {
  const int level = argc > 1 ? atoi(argv[1]) : 3;
  const size_t chunk = argc > 2 ? (size_t)atoi(argv[2]) : 65536;
  const int nchunks = argc > 3 ? atoi(argv[3]) : 24;
  struct Cls { const char *name; int kind; uint32_t P; };
  const Cls classes[] = {{"p0", 0, 0}, {"p25", 0, 16384}, {"p50", 0, 32768}, {"p75", 0, 49152}, {"p90", 0, 58982}, {"mixed", 2, 0}, {"text", -1, 0}};
  std::vector<uint8_t> data(chunk * nchunks), out(chunk + chunk / 200 + 1024), zo(ZSTD_compressBound(chunk)), back(chunk);
  auto run = [&](const char *name, const uint8_t *d, int n) {
    size_t mine = 0, theirs = 0;
    for (int i = 0; i < n; i++) {
      const uint8_t *src = d + (size_t)i * chunk;
      size_t m = model_compress(src, chunk, out.data(), out.size(), level, 0);
      size_t r = ZSTD_decompress(back.data(), chunk, out.data(), m);
      if (ZSTD_isError(r) || r != chunk || memcmp(back.data(), src, chunk)) { printf("ROUNDTRIP FAILED class %s chunk %d\n", name, i); exit(1); }
      mine += m;
      theirs += ZSTD_compress(zo.data(), zo.size(), src, chunk, level);
    }
    printf("%-10s mine %9zu  libzstd %9zu  ratio %.4f\n", name, mine, theirs, (double)mine / theirs);
  };
  printf("level %d chunk %zu\n", level, chunk);
  for (const Cls &c : classes) {
    if (c.kind >= 0) orc_gen_batch(data.data(), chunk, 0, nchunks, c.kind, c.P);
    else orc_gen_textlike(data.data(), data.size());
    run(c.name, data.data(), nchunks);
  }
  for (int a = 4; a < argc; a++) {
    FILE *f = fopen(argv[a], "rb");
    if (!f) continue;
    std::vector<uint8_t> fd(chunk * 64);
    size_t got = fread(fd.data(), 1, fd.size(), f);
    fclose(f);
    const int n = (int)(got / chunk);
    if (n) run(argv[a] + (strlen(argv[a]) > 10 ? strlen(argv[a]) - 10 : 0), fd.data(), n);
  }
  return 0;
}
