// tests/cpp/test_cpp_api.cu -- native test of the C++ drop-in surface, written the way the reference's
// own boundary tests are (standalone main(), returns 0/1; see tests/test_nvcomp_interface.cu:195-369,
// tests/test_c_api.cpp, tests/test_inference_api.cu:545-585, tests/test_c_api_edge_cases.cu in the
// reference).  Built against include/*.h + libcuda_zstd_b200.so only.  Input data is generated here
// (integer generator), results are checked by round trip and by frame-header inspection; the parity
// against libzstd / the oracle is the job of the Python tests.
#include "cuda_zstd_manager.h"
#include "cuda_zstd_nvcomp.h"
#include "cuda_zstd_batch_c.h"
#include "cuda_zstd_hybrid.h"

#include <cstdio>
#include <cstring>
#include <vector>

using namespace cuda_zstd;
using namespace cuda_zstd::nvcomp_v5;

static int g_fail = 0;
#define CHECK(cond)                                                                      \
  do {                                                                                   \
    if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); g_fail++; } \
  } while (0)
#define CUDA_OK(x) CHECK((x) == cudaSuccess)

static void fill(std::vector<unsigned char> &v, unsigned seed) {
  // repetitive text-like bytes with some noise: compresses well, exercises Huffman + FSE + repcodes
  unsigned long long s = seed * 0x9E3779B97F4A7C15ull + 1;
  const char *words[] = {"alpha ", "beta ", "gamma ", "delta ", "epsilon ", "0123456789 "};
  size_t pos = 0;
  while (pos < v.size()) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    const char *w = words[(s >> 33) % 6];
    size_t len = std::strlen(w);
    for (size_t k = 0; k < len && pos < v.size(); ++k) v[pos++] = (unsigned char)w[k];
    if (((s >> 40) & 15) == 0 && pos < v.size()) v[pos++] = (unsigned char)(s >> 48);
  }
}

static void test_batch_manager_roundtrip() {
  // shape of tests/test_nvcomp_interface.cu:195-369: 4 x 32 KB, host pointer tables, capacity = bound
  const size_t chunk = 32 * 1024, n = 4;
  std::vector<unsigned char> h(chunk * n);
  fill(h, 1);
  unsigned char *d_in, *d_comp, *d_back;
  auto mgr = create_batch_manager(5);
  CHECK(mgr != nullptr);
  CHECK(mgr->get_compression_level() == 5);
  const size_t bound = mgr->get_max_compressed_size(chunk);
  CHECK(bound == chunk + chunk / 255 + 3 + 512);
  CUDA_OK(cudaMalloc(&d_in, chunk * n));
  CUDA_OK(cudaMalloc(&d_comp, bound * n));
  CUDA_OK(cudaMalloc(&d_back, chunk * n));
  CUDA_OK(cudaMemcpy(d_in, h.data(), chunk * n, cudaMemcpyHostToDevice));
  std::vector<size_t> sizes(n, chunk);
  const size_t ws_bytes = mgr->get_batch_compress_temp_size(sizes);
  CHECK(ws_bytes > 0);
  void *ws;
  CUDA_OK(cudaMalloc(&ws, ws_bytes));
  std::vector<BatchItem> items(n);
  for (size_t i = 0; i < n; ++i) { items[i].input_ptr = d_in + i * chunk; items[i].input_size = chunk; items[i].output_ptr = d_comp + i * bound; items[i].output_size = bound; }
  CHECK(mgr->compress_batch(items, ws, ws_bytes) == Status::SUCCESS);
  size_t total = 0;
  for (auto &it : items) { CHECK(it.status == Status::SUCCESS); CHECK(it.output_size > 0 && it.output_size < chunk / 2); total += it.output_size; }
  CHECK(mgr->get_stats().input_bytes == chunk * n && mgr->get_stats().output_bytes == total);
  // frames carry the content size the reference's metadata parser expects (tests/test_nvcomp_interface.cu:599)
  NvcompV5Metadata md;
  CHECK(get_metadata(d_comp, items[0].output_size, md) == Status::SUCCESS);
  CHECK(md.uncompressed_size == chunk);
  size_t dsz = 0;
  CHECK(get_decompressed_size(d_comp, items[0].output_size, &dsz) == Status::SUCCESS && dsz == chunk);
  CHECK(is_nvcomp_zstd_format(d_comp, items[0].output_size));
  // decompress with the SAME workspace (tests/test_nvcomp_interface.cu:298-307)
  std::vector<BatchItem> ditems(n);
  for (size_t i = 0; i < n; ++i) { ditems[i].input_ptr = d_comp + i * bound; ditems[i].input_size = items[i].output_size; ditems[i].output_ptr = d_back + i * chunk; ditems[i].output_size = chunk; }
  CHECK(mgr->decompress_batch(ditems, ws, ws_bytes) == Status::SUCCESS);
  std::vector<unsigned char> back(chunk * n);
  CUDA_OK(cudaMemcpy(back.data(), d_back, chunk * n, cudaMemcpyDeviceToHost));
  CHECK(std::memcmp(back.data(), h.data(), chunk * n) == 0);
  for (auto &it : ditems) CHECK(it.status == Status::SUCCESS && it.output_size == chunk);
  // per-item failure -> ERROR_GENERIC overall, item status set (src/cuda_zstd_manager.cu:5770-5795)
  ditems[2].output_size = 100;
  CHECK(mgr->decompress_batch(ditems, ws, ws_bytes) == Status::ERROR_GENERIC);
  CHECK(ditems[2].status == Status::ERROR_BUFFER_TOO_SMALL && ditems[0].status == Status::SUCCESS);
  // empty batch, tiny workspace, level setter
  CHECK(mgr->compress_batch(std::vector<BatchItem>(), ws, ws_bytes) == Status::SUCCESS);
  CHECK(mgr->compress_batch(items, ws, 64) == Status::ERROR_BUFFER_TOO_SMALL);
  CHECK(mgr->set_compression_level(0) == Status::ERROR_INVALID_PARAMETER && mgr->get_compression_level() == 5);   // unchanged (manager.cu:1517-1522)
  CHECK(mgr->set_compression_level(9) == Status::SUCCESS && mgr->get_compression_level() == 9);
  cudaFree(d_in); cudaFree(d_comp); cudaFree(d_back); cudaFree(ws);
}

static void test_single_buffer_and_inference_api() {
  const size_t n = 128 * 1024;                                  // tests/test_c_api.cpp uses 128 KB
  std::vector<unsigned char> h(n);
  fill(h, 2);
  ZstdBatchManager mgr(CompressionConfig::from_level(3));
  unsigned char *d_in, *d_comp, *d_back;
  const size_t bound = mgr.get_max_compressed_size(n), ws_bytes = mgr.get_compress_temp_size(n);
  void *ws;
  CUDA_OK(cudaMalloc(&d_in, n)); CUDA_OK(cudaMalloc(&d_comp, bound)); CUDA_OK(cudaMalloc(&d_back, n)); CUDA_OK(cudaMalloc(&ws, ws_bytes));
  CUDA_OK(cudaMemcpy(d_in, h.data(), n, cudaMemcpyHostToDevice));
  size_t csz = bound;
  CHECK(mgr.compress(d_in, n, d_comp, &csz, ws, ws_bytes, nullptr, 0) == Status::SUCCESS);
  CHECK(csz > 0 && csz < n / 2);
  // null / zero-size / capacity errors (tests/test_c_api_edge_cases.cu:31-107, tests/test_inference_api.cu:545-585)
  size_t tmp = bound;
  CHECK(mgr.compress(nullptr, n, d_comp, &tmp, ws, ws_bytes, nullptr, 0) == Status::ERROR_INVALID_PARAMETER);
  CHECK(mgr.compress(d_in, 0, d_comp, &tmp, ws, ws_bytes, nullptr, 0) == Status::ERROR_INVALID_PARAMETER);
  size_t out_sz = n;
  CHECK(mgr.decompress(d_comp, csz, d_back, &out_sz, ws, ws_bytes) == Status::SUCCESS && out_sz == n);
  size_t actual = 0;
  CHECK(mgr.decompress_to_preallocated(d_comp, csz, d_back, n, &actual, ws, ws_bytes) == Status::SUCCESS && actual == n);
  CHECK(mgr.decompress_to_preallocated(nullptr, csz, d_back, n, &actual, ws, ws_bytes) == Status::ERROR_INVALID_PARAMETER);
  CHECK(mgr.decompress_to_preallocated(d_comp, csz, d_back, 0, &actual, ws, ws_bytes) == Status::ERROR_BUFFER_TOO_SMALL);
  // true no-sync: size lands in device memory, caller synchronises
  size_t *d_actual;
  CUDA_OK(cudaMalloc(&d_actual, sizeof(size_t)));
  CUDA_OK(cudaMemset(d_back, 0, n));
  cudaStream_t s;
  CUDA_OK(cudaStreamCreate(&s));
  CHECK(mgr.decompress_async_no_sync(d_comp, csz, d_back, n, d_actual, ws, ws_bytes, s) == Status::SUCCESS);
  CUDA_OK(cudaStreamSynchronize(s));
  size_t got = 0;
  CUDA_OK(cudaMemcpy(&got, d_actual, sizeof got, cudaMemcpyDeviceToHost));
  CHECK(got == n);
  std::vector<unsigned char> back(n);
  CUDA_OK(cudaMemcpy(back.data(), d_back, n, cudaMemcpyDeviceToHost));
  CHECK(std::memcmp(back.data(), h.data(), n) == 0);
  // pageable host buffers on either side are staged through the workspace tail (tests/test_two_phase_unit.cu:56 passes
  // a std::vector as the compress input): 256 KB = two blocks in one frame
  {
    const size_t m = 256 * 1024;
    std::vector<unsigned char> hi(m), hc(mgr.get_max_compressed_size(m)), hb(m);
    fill(hi, 9);
    const size_t wb = mgr.get_compress_temp_size(m);
    void *w2;
    CUDA_OK(cudaMalloc(&w2, wb));
    size_t hcs = hc.size();
    CHECK(mgr.compress(hi.data(), m, hc.data(), &hcs, w2, wb, nullptr, 0) == Status::SUCCESS && hcs > 0 && hcs < m);
    CHECK(hc[0] == 0x28 && hc[1] == 0xB5 && hc[2] == 0x2F && hc[3] == 0xFD);
    size_t hbs = m;
    CHECK(mgr.decompress(hc.data(), hcs, hb.data(), &hbs, w2, wb) == Status::SUCCESS && hbs == m);
    CHECK(std::memcmp(hb.data(), hi.data(), m) == 0);
    size_t small = 64;                                             // no room to stage the output: refused, not a fault
    CHECK(mgr.compress(hi.data(), m, hc.data(), &hcs, w2, small, nullptr, 0) == Status::ERROR_BUFFER_TOO_SMALL);
    cudaFree(w2);
  }
  // a checksummed multi-block frame (blocks encoded side by side, XXH64 over the whole content appended): the decoder
  // verifies it, and a flipped payload byte is reported
  {
    CompressionConfig cc = CompressionConfig::from_level(3);
    cc.checksum = ChecksumPolicy::COMPUTE_AND_VERIFY;
    ZstdBatchManager cm(cc);
    const size_t m = (1 << 20) + 777;
    std::vector<unsigned char> hi(m), hb(m);
    fill(hi, 11);
    unsigned char *di, *dc, *db; void *w3;
    const size_t cb = cm.get_max_compressed_size(m), wb = cm.get_compress_temp_size(m);
    CUDA_OK(cudaMalloc(&di, m)); CUDA_OK(cudaMalloc(&dc, cb)); CUDA_OK(cudaMalloc(&db, m)); CUDA_OK(cudaMalloc(&w3, wb));
    CUDA_OK(cudaMemcpy(di, hi.data(), m, cudaMemcpyHostToDevice));
    size_t cs = cb;
    CHECK(cm.compress(di, m, dc, &cs, w3, wb, nullptr, 0) == Status::SUCCESS && cs > 14 && cs < m);
    unsigned char fh[6];
    CUDA_OK(cudaMemcpy(fh, dc, 6, cudaMemcpyDeviceToHost));
    CHECK((fh[4] & 0x04) && (fh[4] >> 6) == 2 && fh[5] == 0x30);       // checksum flag, 4-byte content size, 64 KB window (buffers <= 4 MB: 64 KB blocks)
    size_t os = m;
    CHECK(cm.decompress(dc, cs, db, &os, w3, wb) == Status::SUCCESS && os == m);
    CUDA_OK(cudaMemcpy(hb.data(), db, m, cudaMemcpyDeviceToHost));
    CHECK(std::memcmp(hb.data(), hi.data(), m) == 0);
    unsigned char last4[4];
    CUDA_OK(cudaMemcpy(last4, dc + cs - 4, 4, cudaMemcpyDeviceToHost));
    last4[0] ^= 0x55;
    CUDA_OK(cudaMemcpy(dc + cs - 4, last4, 4, cudaMemcpyHostToDevice));
    os = m;
    CHECK(cm.decompress(dc, cs, db, &os, w3, wb) == Status::ERROR_CHECKSUM_FAILED);
    cudaFree(di); cudaFree(dc); cudaFree(db); cudaFree(w3);
  }
  // corrupt magic -> ERROR_INVALID_MAGIC
  unsigned char zero = 0;
  CUDA_OK(cudaMemcpy(d_comp, &zero, 1, cudaMemcpyHostToDevice));
  out_sz = n;
  CHECK(mgr.decompress(d_comp, csz, d_back, &out_sz, ws, ws_bytes) == Status::ERROR_INVALID_MAGIC);
  // convenience calls
  size_t c2 = bound;
  CHECK(compress_simple(d_in, n, d_comp, &c2, 1) == Status::SUCCESS);
  size_t d2 = n;
  CHECK(decompress_simple(d_comp, c2, d_back, &d2) == Status::SUCCESS && d2 == n);
  void *iws = nullptr; size_t iws_bytes = 0;
  CHECK(mgr.allocate_inference_workspace(bound, n, &iws, &iws_bytes) == Status::SUCCESS && iws && iws_bytes > 0);
  CHECK(mgr.free_inference_workspace(iws) == Status::SUCCESS);
  cudaStreamDestroy(s);
  cudaFree(d_in); cudaFree(d_comp); cudaFree(d_back); cudaFree(ws); cudaFree(d_actual);
}

static void test_nvcomp_facade_device_tables() {
  // shape of tests/test_nvcomp_batch.cu:132-150: 8 x 64 KB, level 5, device pointer tables, host sizes
  const size_t chunk = 64 * 1024, n = 8;
  std::vector<unsigned char> h(chunk * n);
  fill(h, 3);
  NvcompV5Options opts;
  opts.level = 5; opts.enable_checksum = true;
  NvcompV5BatchManager mgr(opts);
  const size_t bound = mgr.get_max_compressed_chunk_size(chunk);
  unsigned char *d_in, *d_comp, *d_back;
  CUDA_OK(cudaMalloc(&d_in, chunk * n)); CUDA_OK(cudaMalloc(&d_comp, bound * n)); CUDA_OK(cudaMalloc(&d_back, chunk * n));
  CUDA_OK(cudaMemcpy(d_in, h.data(), chunk * n, cudaMemcpyHostToDevice));
  std::vector<const void *> in(n); std::vector<void *> out(n), back(n);
  std::vector<size_t> in_sz(n, chunk), out_sz(n, bound), back_sz(n, chunk);
  for (size_t i = 0; i < n; ++i) { in[i] = d_in + i * chunk; out[i] = d_comp + i * bound; back[i] = d_back + i * chunk; }
  const void **d_inp; void **d_outp; size_t *d_outsz;
  CUDA_OK(cudaMalloc(&d_inp, n * 8)); CUDA_OK(cudaMalloc(&d_outp, n * 8)); CUDA_OK(cudaMalloc(&d_outsz, n * 8));
  CUDA_OK(cudaMemcpy(d_inp, in.data(), n * 8, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(d_outp, out.data(), n * 8, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(d_outsz, out_sz.data(), n * 8, cudaMemcpyHostToDevice));
  const size_t ws_bytes = mgr.get_compress_temp_size(in_sz.data(), n);
  void *ws;
  CUDA_OK(cudaMalloc(&ws, ws_bytes));
  CHECK(mgr.compress_async(d_inp, in_sz.data(), n, d_outp, d_outsz, ws, ws_bytes) == Status::SUCCESS);
  CUDA_OK(cudaMemcpy(out_sz.data(), d_outsz, n * 8, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i) CHECK(out_sz[i] > 0 && out_sz[i] < chunk / 2);
  // the frame announces its checksum (opts.enable_checksum) in the header descriptor
  unsigned char head[8];
  CUDA_OK(cudaMemcpy(head, d_comp, 8, cudaMemcpyDeviceToHost));
  CHECK(head[0] == 0x28 && head[1] == 0xB5 && head[2] == 0x2F && head[3] == 0xFD && (head[4] & 0x04));
  CHECK(mgr.decompress_async(out.data(), out_sz.data(), n, back.data(), back_sz.data(), ws, ws_bytes) == Status::SUCCESS);
  std::vector<unsigned char> hb(chunk * n);
  CUDA_OK(cudaMemcpy(hb.data(), d_back, chunk * n, cudaMemcpyDeviceToHost));
  CHECK(std::memcmp(hb.data(), h.data(), chunk * n) == 0);
  // a flipped payload bit is caught by the XXH64 trailer (COMPUTE_AND_VERIFY)
  unsigned char b;
  CUDA_OK(cudaMemcpy(&b, d_comp + out_sz[0] - 6, 1, cudaMemcpyDeviceToHost));
  b ^= 0x10;
  CUDA_OK(cudaMemcpy(d_comp + out_sz[0] - 6, &b, 1, cudaMemcpyHostToDevice));
  CHECK(mgr.decompress_async(out.data(), out_sz.data(), n, back.data(), back_sz.data(), ws, ws_bytes) == Status::ERROR_GENERIC);
  // null tables / empty batch (src/cuda_zstd_nvcomp.cpp:305-310)
  CHECK(mgr.compress_async(nullptr, in_sz.data(), n, d_outp, d_outsz, ws, ws_bytes) == Status::ERROR_INVALID_PARAMETER);
  CHECK(mgr.compress_async(d_inp, in_sz.data(), 0, d_outp, d_outsz, ws, ws_bytes) == Status::SUCCESS);
  cudaFree(d_in); cudaFree(d_comp); cudaFree(d_back); cudaFree(d_inp); cudaFree(d_outp); cudaFree(d_outsz); cudaFree(ws);
}

static void test_status_maps_and_config() {
  // tests/test_nvcomp_interface.cu:634-677
  for (int s = 0; s <= 28; ++s) {
    const int e = status_to_nvcomp_error(static_cast<Status>(s));
    const bool keep = s == 0 || s == 2 || s == 3 || s == 4 || s == 6 || s == 7 || s == 10 || s == 12;
    CHECK(e == (keep ? s : 1));
  }
  CHECK(nvcomp_error_to_status(999) == Status::ERROR_GENERIC);
  CHECK(std::strcmp(status_to_string(Status::ERROR_CHECKSUM_FAILED), "ERROR_CHECKSUM_FAILED") == 0);
  CompressionConfig c = CompressionConfig::from_level(1);
  CHECK(c.strategy == Strategy::FAST && c.validate() == Status::SUCCESS);
  CHECK(CompressionConfig::level_to_strategy(3) == Strategy::DFAST && CompressionConfig::level_to_strategy(9) == Strategy::LAZY);
  c.level = 99;
  CHECK(c.validate() == Status::ERROR_INVALID_PARAMETER);
  NvcompV5Options o = to_nvcomp_v5_opts(CompressionConfig::from_level(7));
  CHECK(o.level == 7 && from_nvcomp_v5_opts(o).level == 7);
  CHECK(estimate_compressed_size(65536, 3) == 66308 && estimate_compressed_size(131072, 3) == 132101);
  nvcompZstdManagerHandle hdl = nvcomp_zstd_create_manager_v5(3);
  CHECK(hdl != nullptr && nvcomp_zstd_get_compress_temp_size_v5(hdl, 65536) > 0 && nvcomp_zstd_get_compress_temp_size_v5(nullptr, 65536) == 0);
  nvcomp_zstd_destroy_manager_v5(hdl);
  clear_last_error();
  ZstdBatchManager m;
  size_t x = 10;
  CHECK(m.decompress(nullptr, 10, nullptr, &x, nullptr, 0) == Status::ERROR_INVALID_PARAMETER);
  CHECK(get_last_error().status == Status::ERROR_INVALID_PARAMETER);
}

static void test_hybrid_engine_cpp() {
  // HybridEngine C++ surface (reference include/cuda_zstd_hybrid.h:88-263; shape of tests/test_hybrid.cu's round trips), GPU-only here
  HybridConfig cfg;
  cfg.mode = HybridMode::FORCE_CPU;                         // accepted, stored, and still served by the GPU
  cfg.compression_level = 5;
  cfg.enable_profiling = true;
  HybridEngine eng(cfg);
  CHECK(eng.get_config().mode == HybridMode::FORCE_CPU && eng.get_config().compression_level == 5);
  CHECK(eng.query_routing(100, DataLocation::HOST, DataLocation::HOST, true) == ExecutionBackend::GPU_KERNELS);
  HybridConfig bad = cfg;
  bad.compression_level = 0;
  CHECK(eng.configure(bad) == Status::ERROR_INVALID_PARAMETER && eng.get_config().compression_level == 5);
  CHECK(eng.set_compression_level(23) == Status::ERROR_INVALID_PARAMETER);
  CHECK(eng.set_compression_level(3) == Status::SUCCESS && eng.get_config().compression_level == 3);
  // single buffer, host to host, 300 KB (a multi-block frame)
  std::vector<unsigned char> h(300 * 1024 + 17), c(eng.get_max_compressed_size(h.size())), back(h.size());
  fill(h, 21);
  size_t csz = c.size();
  HybridResult res;
  CHECK(eng.compress(h.data(), h.size(), c.data(), &csz, DataLocation::HOST, DataLocation::HOST, &res) == Status::SUCCESS);
  CHECK(csz > 0 && csz < h.size() / 2 && res.backend_used == ExecutionBackend::GPU_KERNELS && res.output_bytes == csz);
  CHECK(c[0] == 0x28 && c[1] == 0xB5 && c[2] == 0x2F && c[3] == 0xFD);
  size_t bsz = back.size();
  CHECK(eng.decompress(c.data(), csz, back.data(), &bsz, DataLocation::UNKNOWN, DataLocation::UNKNOWN, &res) == Status::SUCCESS);
  CHECK(bsz == h.size() && std::memcmp(back.data(), h.data(), h.size()) == 0);
  CHECK(res.input_location == DataLocation::HOST && res.output_location == DataLocation::HOST);
  CHECK(eng.get_observed_throughput(ExecutionBackend::GPU_KERNELS, true) > 0 && eng.get_observed_throughput(ExecutionBackend::CPU_LIBZSTD, true) == 0);
  size_t tiny = 10;
  CHECK(eng.decompress(c.data(), csz, back.data(), &tiny) == Status::ERROR_BUFFER_TOO_SMALL);
  CHECK(eng.compress(nullptr, 10, c.data(), &csz) == Status::ERROR_INVALID_PARAMETER);
  CHECK(eng.compress(h.data(), 0, c.data(), &csz) == Status::ERROR_INVALID_PARAMETER);
  // batch: 6 host items of different sizes in ONE launch, one of them with too little room
  const size_t n = 6;
  const size_t sizes[n] = {1, 100, 4096, 65536, 100000, 131072};
  std::vector<std::vector<unsigned char>> in(n), comp(n), out(n);
  std::vector<const void *> ip(n);
  std::vector<void *> cp(n), op(n);
  std::vector<size_t> isz(n), csz2(n), osz(n);
  for (size_t i = 0; i < n; ++i) {
    in[i].resize(sizes[i]); fill(in[i], 30 + (unsigned)i);
    comp[i].resize(eng.get_max_compressed_size(sizes[i])); out[i].resize(sizes[i]);
    ip[i] = in[i].data(); cp[i] = comp[i].data(); op[i] = out[i].data(); isz[i] = sizes[i]; csz2[i] = comp[i].size(); osz[i] = sizes[i];
  }
  std::vector<BatchRoutingResult> rr(n);
  CHECK(eng.compress_batch(ip.data(), isz.data(), cp.data(), csz2.data(), n, DataLocation::HOST, DataLocation::HOST, rr.data()) == Status::SUCCESS);
  for (size_t i = 0; i < n; ++i) CHECK(rr[i].status == Status::SUCCESS && rr[i].backend_used == ExecutionBackend::GPU_BATCH && csz2[i] > 0 && rr[i].output_bytes == csz2[i]);
  std::vector<const void *> cip(cp.begin(), cp.end());
  CHECK(eng.decompress_batch(cip.data(), csz2.data(), op.data(), osz.data(), n, DataLocation::HOST, DataLocation::HOST, rr.data()) == Status::SUCCESS);
  for (size_t i = 0; i < n; ++i) CHECK(osz[i] == sizes[i] && std::memcmp(out[i].data(), in[i].data(), sizes[i]) == 0);
  osz[3] = 1000;                                            // capacity too small for item 3 only
  CHECK(eng.decompress_batch(cip.data(), csz2.data(), op.data(), osz.data(), n, DataLocation::HOST, DataLocation::HOST, rr.data()) == Status::ERROR_DECOMPRESSION);
  CHECK(rr[3].status == Status::ERROR_BUFFER_TOO_SMALL && osz[3] == 0 && rr[2].status == Status::SUCCESS && osz[2] == sizes[2]);
  CHECK(eng.compress_batch(ip.data(), isz.data(), cp.data(), csz2.data(), 0) == Status::ERROR_INVALID_PARAMETER);   // hybrid.cu:920
  CompressionStats st = eng.get_stats();
  CHECK(st.input_bytes > 0 && st.output_bytes > 0 && st.bytes_decompressed > 0);
  eng.reset_stats();
  CHECK(eng.get_stats().input_bytes == 0);
  // free functions and factories
  csz = c.size();
  CHECK(hybrid_compress(h.data(), 5000, c.data(), &csz, DataLocation::HOST, DataLocation::HOST, 7) == Status::SUCCESS);
  bsz = back.size();
  CHECK(hybrid_decompress(c.data(), csz, back.data(), &bsz) == Status::SUCCESS && bsz == 5000 && std::memcmp(back.data(), h.data(), 5000) == 0);
  CHECK(create_hybrid_engine(4)->get_config().compression_level == 4);
  size_t vs = 0;
  CHECK(validate_compressed_data(c.data(), csz) == Status::SUCCESS && validate_compressed_data(h.data(), 100) == Status::ERROR_INVALID_MAGIC);
  CHECK(get_decompressed_size(c.data(), csz, &vs) == Status::SUCCESS && vs == 5000);
}

int main() {
  int dev = 0;
  if (cudaGetDeviceCount(&dev) != cudaSuccess || dev == 0) { std::printf("SKIP: no CUDA device\n"); return 77; }
  test_status_maps_and_config();
  test_batch_manager_roundtrip();
  test_single_buffer_and_inference_api();
  test_nvcomp_facade_device_tables();
  test_hybrid_engine_cpp();
  std::printf(g_fail ? "FAILED (%d)\n" : "ALL PASSED\n", g_fail);
  return g_fail ? 1 : 0;
}
