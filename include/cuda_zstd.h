// cuda_zstd.h -- umbrella header (reference include/cuda_zstd.h).  The reference's also pulls in its FSE / Huffman /
// dictionary / stream-pool / RAII-pointer headers; those are internals of its codec or out of scope here (SURVEY.md
// section 8), so this one gathers the public surface this build exports.  The reference's Python binding
// (python/src/binding.cpp:24-28) compiles against it unchanged.
#ifndef CUDA_ZSTD_H
#define CUDA_ZSTD_H
#include "cuda_zstd_types.h"
#include "cuda_zstd_manager.h"
#include "cuda_zstd_safe_alloc.h"
#endif // CUDA_ZSTD_H
