/* Minimal declaration shim for the system libzstd (1.5.5), whose development header is not
 * in this image.  TEST INFRASTRUCTURE ONLY: used (a) to compile the unmodified reference
 * sources from /root/reference into oracle/_ref/, and (b) by oracle/ref_hybrid_driver.cpp.
 * Only the identifiers the reference uses under src/ and include/ are declared
 * (SURVEY.md section 8c).  The prototypes are libzstd's public, stable ABI. */
#ifndef ORACLE_SHIM_ZSTD_H
#define ORACLE_SHIM_ZSTD_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
size_t ZSTD_compress(void *dst, size_t dstCapacity, const void *src, size_t srcSize, int level);
size_t ZSTD_decompress(void *dst, size_t dstCapacity, const void *src, size_t compressedSize);
size_t ZSTD_compressBound(size_t srcSize);
unsigned ZSTD_isError(size_t code);
const char *ZSTD_getErrorName(size_t code);
unsigned long long ZSTD_getFrameContentSize(const void *src, size_t srcSize);
#define ZSTD_BLOCKSIZELOG_MAX 17
#define ZSTD_BLOCKSIZE_MAX (1 << ZSTD_BLOCKSIZELOG_MAX)
#define ZSTD_CONTENTSIZE_UNKNOWN (0ULL - 1)
#define ZSTD_CONTENTSIZE_ERROR (0ULL - 2)
#define ZSTD_MAGIC_SKIPPABLE_START 0x184D2A50
#ifdef __cplusplus
}
#endif
#endif
