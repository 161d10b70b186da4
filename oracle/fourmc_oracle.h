/*
 * fourmc_oracle.h -- CPU restatement of the 4mc LZ4 block path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product (4mc_b200/, include/) may include, link or call this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 *
 * Pinning: checked against the reference built from its own sources (oracle/_ref, see
 * oracle/Makefile) and against the committed golden vectors in tests/golden/ harvested from it
 * (tests/golden/make_golden.py).  The reference's own test-suite holds no vectors for this path
 * (SURVEY.md section 4), so the reference build itself is the pin.
 */
#ifndef FOURMC_ORACLE_H
#define FOURMC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMO_BLOCKSIZE   (4 * 1024 * 1024)       /* native/4mc.c:116 */
#define FMO_MAGIC_4MC   0x344D4300u              /* native/4mc.c:111 */
#define FMO_MAGIC_4MZ   0x344D5A00u              /* native/4mc.c:112 */
#define FMO_VERSION     1u                       /* native/4mc.c:113 */

/* error codes of the container functions: the reference CLI's exit codes, negated
 * (native/4mc.c:135-161): -1 generic, -2 input (truncated), -3 output, -4 content */
#define FMO_ERR_GENERIC  (-1)
#define FMO_ERR_INPUT    (-2)
#define FMO_ERR_OUTPUT   (-3)
#define FMO_ERR_CONTENT  (-4)

/* XXH32 -- native/lz4/xxhash.c:263-416 */
uint32_t fmo_xxh32(const void *data, size_t len, uint32_t seed);

/* LZ4_decompress_safe -- native/lz4/lz4.c:1936-2350 (full block, no dictionary).
 * Returns decoded byte count, or -(input position of the failure)-1 like the reference. */
int fmo_lz4_decompress_safe(const uint8_t *src, uint8_t *dst, int src_size, int dst_capacity);

/* LZ4_COMPRESSBOUND -- native/lz4/lz4.h:212 */
int fmo_lz4_compress_bound(int n);

/* A greedy single-probe LZ4 block compressor obeying the end-of-block rules the reference decoder
 * enforces (native/lz4/lz4.c:243-247).  NOT byte-identical to LZ4_compress_default (the task
 * does not require it); it exists to make valid streams for tests.  Returns the compressed
 * size, or 0 when the result would exceed max_out (native/4mc.c:301 "stored" convention). */
int fmo_lz4_compress(const uint8_t *src, uint8_t *dst, int n, int max_out);

/* ---- container (native/4mc.c:220-386 writer, :560-707 reader, Appendix A of SURVEY.md) ---- */

/* worst-case size of a .4mc stream for n input bytes (all blocks stored) */
size_t fmo_4mc_bound(size_t n);

/* Whole-buffer writer: header, one block per 4 MiB, EOS, footer.  Returns stream size, <0 on error. */
long long fmo_4mc_compress(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap);

/* Whole-buffer reader; accepts concatenated streams (native/4mc.c:908-912).  Returns decoded size
 * or one of FMO_ERR_*. */
long long fmo_4mc_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap);

/* zstd_oracle.c: a strict Zstandard frame decoder (ZSTD_decompress for valid frames: native/4mc.c:810) and the
 * 4mz reader (native/4mc.c:709-857).  fmo_zstd_decompress returns the decoded size or -1. */
long long fmo_zstd_decompress(uint8_t *dst, long long cap, const uint8_t *src, long long n);
long long fmo_4mz_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap);

/* Footer/index reader restating FourMcInputStream.readIndex
 * (java/hadoop-4mc/src/main/java/com/fing/compression/fourmc/FourMcInputStream.java:163-239).
 * Returns number of blocks (0 when the file is too small to hold an index), writes up to
 * max_blocks absolute offsets; FMO_ERR_CONTENT for the IOException cases. */
long long fmo_4mc_read_index(const uint8_t *file, size_t file_size, uint64_t magic,
                             int64_t *offsets, size_t max_blocks);

/* FourMcBlockIndex search semantics (FourMcBlockIndex.java:92-173); NOT_FOUND = -1 */
int64_t fmo_index_find_next_position(const int64_t *offs, int n, int64_t pos);
int64_t fmo_index_find_belonging_block(const int64_t *offs, int n, int64_t pos);
int64_t fmo_index_align_slice_start(const int64_t *offs, int n, int64_t start, int64_t end);
int64_t fmo_index_align_slice_end(const int64_t *offs, int n, int64_t end, int64_t file_size);

#ifdef __cplusplus
}
#endif
#endif
