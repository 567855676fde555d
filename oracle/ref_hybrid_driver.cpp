// TEST INFRASTRUCTURE ONLY (oracle/): a thin C-callable driver around the UNMODIFIED reference
// library built from /root/reference (see oracle/build_ref.sh).  It runs the reference's own CPU
// implementation of the batch path -- cuda_zstd::HybridEngine{mode = FORCE_CPU} -- which is what
// BASELINE.json config 1 names (reference: src/cuda_zstd_hybrid.cu:779 compress, :838 decompress,
// :402-458 cpu_compress -> ZSTD_compress).  One engine per thread, chunks split contiguously
// (the reference documents "one engine per thread", docs/HYBRID-ENGINE.md:359-361).
// Nothing under custom-nvcomp-with-zstd_b200/ may link or load this.
#include "cuda_zstd_hybrid.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

using cuda_zstd::DataLocation;
using cuda_zstd::HybridConfig;
using cuda_zstd::HybridEngine;
using cuda_zstd::HybridMode;
using cuda_zstd::Status;

namespace {
template <class Fn> double run_threads(int threads, size_t n, Fn fn) {
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    size_t lo = n * t / threads, hi = n * (t + 1) / threads;
    pool.emplace_back([=] { fn(lo, hi); });
  }
  for (auto &th : pool) th.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
} // namespace

extern "C" {

// Compress n chunks (in + i*stride, sizes[i]) into out + i*out_stride; returns seconds, <0 on error.
double ref_hybrid_cpu_compress(const unsigned char *in, const size_t *sizes, size_t stride, size_t n,
                               unsigned char *out, size_t out_stride, size_t *out_sizes, int level,
                               int threads) {
  if (threads < 1) threads = 1;
  std::vector<int> bad(threads, 0);
  double s = run_threads(threads, n, [&](size_t lo, size_t hi) {
    HybridConfig cfg;
    cfg.mode = HybridMode::FORCE_CPU;
    cfg.compression_level = level;
    HybridEngine eng(cfg);
    for (size_t i = lo; i < hi; ++i) {
      size_t cap = out_stride;
      Status st = eng.compress(in + i * stride, sizes[i], out + i * out_stride, &cap,
                               DataLocation::HOST, DataLocation::HOST);
      if (st != Status::SUCCESS) { out_sizes[i] = 0; continue; }
      out_sizes[i] = cap;
    }
  });
  for (size_t i = 0; i < n; ++i) if (out_sizes[i] == 0) return -1.0;
  return s;
}

double ref_hybrid_cpu_decompress(const unsigned char *in, const size_t *sizes, size_t stride, size_t n,
                                 unsigned char *out, size_t out_stride, size_t *out_sizes,
                                 int threads) {
  if (threads < 1) threads = 1;
  double s = run_threads(threads, n, [&](size_t lo, size_t hi) {
    HybridConfig cfg;
    cfg.mode = HybridMode::FORCE_CPU;
    HybridEngine eng(cfg);
    for (size_t i = lo; i < hi; ++i) {
      size_t cap = out_stride;
      Status st = eng.decompress(in + i * stride, sizes[i], out + i * out_stride, &cap,
                                 DataLocation::HOST, DataLocation::HOST);
      out_sizes[i] = (st == Status::SUCCESS) ? cap : (size_t)-1;
    }
  });
  for (size_t i = 0; i < n; ++i) if (out_sizes[i] == (size_t)-1) return -1.0;
  return s;
}

int ref_hybrid_hw_threads(void) { return (int)std::max(1u, std::thread::hardware_concurrency()); }
}
