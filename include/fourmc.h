/*
 * fourmc.h -- C-ABI of lib4mcgpu.so, the B200 (sm_100a) implementation of the 4mc block hot path.
 *
 * Plain C: pointers, sizes, integer status codes.  No CUDA or torch types in any signature
 * (a CUDA stream is passed as void*, NULL = the context's own stream).  Every entry point names
 * the reference call site(s) it replaces; paths are relative to the reference repository root.
 *
 * Threading: a fourmc_ctx owns its stream and workspaces and must not be used from two threads at
 * once; distinct contexts are independent (the reference is re-entrant the same way: one
 * stack/heap context per call, native/lz4/lz4.c:1416-1432).  Nothing here ever falls back to a
 * CPU codec: without a usable CUDA device every call fails with FOURMC_E_CUDA.
 */
#ifndef FOURMC_H
#define FOURMC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FOURMC_BLOCKSIZE      (4 * 1024 * 1024)   /* native/4mc.c:116 FOURMC_BLOCKSIZE */
#define FOURMC_MAGIC_4MC      0x344D4300u          /* native/4mc.c:111 */
#define FOURMC_MAGIC_4MZ      0x344D5A00u          /* native/4mc.c:112 */
#define FOURMC_VERSION        1u                   /* native/4mc.c:113 */
#define FOURMC_HEADERSIZE     12                   /* native/4mc.c:115 */

/* status / error codes (all negative).  The container functions map the CLI's exit codes
 * (native/4mc.c:135-161): 1 generic, 2 input, 3 output, 4 content. */
#define FOURMC_OK             0
#define FOURMC_E_GENERIC     (-1)   /* exit(1): allocation, unreadable footer                       */
#define FOURMC_E_INPUT       (-2)   /* exit(2): truncated stream ("cannot read next block size")    */
#define FOURMC_E_OUTPUT      (-3)   /* exit(3): destination too small                               */
#define FOURMC_E_CONTENT     (-4)   /* exit(4): bad magic/version/checksum, corrupt LZ4 block       */
#define FOURMC_E_CUDA        (-10)  /* no device, or a CUDA call failed (see fourmc_last_error)     */
#define FOURMC_E_ARG         (-11)  /* invalid argument                                             */
#define FOURMC_E_UNSUPPORTED (-12)  /* valid request this build does not implement (e.g. 4mz)       */

/* per-block status written by the batch decoders */
#define FOURMC_BLOCK_OK        0
#define FOURMC_BLOCK_CHECKSUM  1    /* XXH32(payload) != header checksum   (native/4mc.c:637,645)  */
#define FOURMC_BLOCK_CORRUPT   2    /* LZ4_decompress_safe(...) < 0        (native/4mc.c:662)      */
#define FOURMC_BLOCK_TOOLARGE  3    /* csize or usize beyond 4 MiB         (native/4mc.c:618,651)  */

typedef struct fourmc_ctx fourmc_ctx;

/* ---- context ------------------------------------------------------------------------------ */

/* Creates a context on CUDA device `device` (-1 = current device).  Returns FOURMC_OK or
 * FOURMC_E_CUDA.  Workspaces grow on demand and are kept until fourmc_ctx_destroy. */
int  fourmc_ctx_create(fourmc_ctx **out, int device);
void fourmc_ctx_destroy(fourmc_ctx *ctx);
/* Text of the last failure on this context (never NULL). */
const char *fourmc_last_error(const fourmc_ctx *ctx);
/* Number of kernels this library has launched on the context since creation. */
uint64_t fourmc_kernel_launches(const fourmc_ctx *ctx);
/* Per-kernel device timing for benchmarks: while enabled, every launch is bracketed by a CUDA
 * event pair on its own stream (no synchronisation).  fourmc_timing_collect() synchronises the
 * device, writes one line "kernel_name launches total_ms\n" per kernel into buf, resets the
 * counters and returns the text length. */
int  fourmc_timing_enable(fourmc_ctx *ctx, int on);
long long fourmc_timing_collect(fourmc_ctx *ctx, char *buf, size_t cap);
/* Reproducible compressed bytes.  The match finders fill their tables with racing stores; with this on, ties are
 * settled by position (lowest / highest) and every run -- any context, any GPU of the type -- writes the same bytes,
 * at the price of a few extra barrier rounds per region (a third of the Fast encode kernel's time on text; nothing
 * where the call is bound by PCIe or the file system).  mode -1 (default): on for everything that hands bytes to the
 * host -- whole-stream host calls, per-block calls, files, the JNI library -- and off for the device-resident calls;
 * 0 / 1: off / on everywhere.  Environment FOURMC_REPRODUCIBLE=0|1 sets the default of new contexts. */
int  fourmc_ctx_set_reproducible(fourmc_ctx *ctx, int mode);
/* Blocks until everything queued on the context's stream (or `stream`) has finished. */
int  fourmc_sync(fourmc_ctx *ctx, void *stream);

/* ---- per-block calls, HOST pointers: what native/jni*.c and the native/4mc.c loops call ----- */

/* LZ4_compressBound: native/jniCompressor.c:171, native/lz4/lz4.h:212.  No device needed. */
int fourmc_lz4_compress_bound(int n);

/* One LZ4 block.  Replaces LZ4_compress_default (native/4mc.c:301 via :214, level <= 1),
 * LZ4_compress / LZ4_compressMC / LZ4_compressHC2 (native/jniCompressor.c:91,123,156).
 * Returns the compressed size (> 0), 0 when it does not fit in dst_capacity (the caller then
 * stores the block raw, native/4mc.c:318-329), or a negative FOURMC_E_*.
 * The bytes are a valid LZ4 block but not the reference's bytes.  level: 1 fast (first-occurrence
 * table, greedy) .. 2 medium / 3 high / 4 ultra (LZ4_compressMC / LZ4_compressHC at 4 / 8, native/4mc.c:243-253: exact
 * hash chains over the sliding 64 KiB history, 4 / 32 / 128 candidates, cost-optimal parse; ratios at or above the
 * reference's; 2 bytes of device scratch per input byte). */
int fourmc_lz4_compress(fourmc_ctx *ctx, int level, const void *src, int src_size,
                        void *dst, int dst_capacity);

/* LZ4_decompress_safe: native/4mc.c:661, native/jniDecompressor.c:88.  Same return convention
 * as the reference: decoded size >= 0, or -(input position)-1 on malformed input
 * (native/lz4/lz4.c:2337); FOURMC_E_CUDA (-10) cannot be confused with it only by asking
 * fourmc_last_error(), so device failures are also latched in the context. */
int fourmc_lz4_decompress_safe(fourmc_ctx *ctx, const void *src, int compressed_size,
                               void *dst, int dst_capacity);

/* XXH32: native/4mc.c:269,311,323,585,637,645,685; native/jniCompressor.c:188,
 * native/jniDecompressor.c:112.  *status (may be NULL) receives FOURMC_OK or FOURMC_E_CUDA. */
uint32_t fourmc_xxh32(fourmc_ctx *ctx, const void *data, size_t len, uint32_t seed, int *status);

/* ---- whole-stream calls, HOST buffers: the bodies of fourMCcompressFilename /
 *      fourMcDecompressFileName (native/4mc.c:220-386, :896-934) minus stdio ------------------- */

/* Upper bound of a .4mc stream for n input bytes (every block stored). */
size_t fourmc_4mc_bound(size_t n);

/* in[0..n) -> header, one block per 4 MiB, EOS, footer.  Copies the input
 * to the device in pipelined slices, runs the block kernels, copies the stream back.
 * Returns the stream size or a negative FOURMC_E_*. */
long long fourmc_4mc_compress_host(fourmc_ctx *ctx, int level, const void *in, size_t n,
                                   void *out, size_t out_capacity);

/* Decodes one or more concatenated .4mc streams (native/4mc.c:908-912).  Headers are walked on
 * the host exactly like decodeFourMC (:603-668); payload verification (XXH32) and LZ4 decoding
 * run on the device.  Returns the decoded size or a negative FOURMC_E_* with the reference's
 * precedence (the first failing block in stream order decides). */
long long fourmc_4mc_decompress_host(fourmc_ctx *ctx, const void *in, size_t n,
                                     void *out, size_t out_capacity);

/* Size a stream decodes to (sum of the block headers' usize), without decoding.  Negative on a
 * malformed container. */
long long fourmc_4mc_decoded_size_host(const void *in, size_t n);

/* ---- device-resident calls: inputs and outputs already in HBM ------------------------------- */
/* All d_* are device pointers on the context's device.  Calls are asynchronous on `stream`
 * unless stated; results land in device memory and are read after fourmc_sync(). */

/* Whole stream on device.  d_out receives the complete .4mc stream; *d_out_size (device u64)
 * its length.  d_block_lens (device u32[n_blocks], may be NULL) receives 12+csize per block --
 * the footer deltas a multi-GPU writer all-gathers. */
int fourmc_4mc_compress_device(fourmc_ctx *ctx, void *stream, int level,
                               const void *d_in, size_t n,
                               void *d_out, size_t out_capacity,
                               uint64_t *d_out_size, uint32_t *d_block_lens);

/* Block range only (no file header / EOS / footer): the per-rank span of a sharded writer.
 * d_span receives block records back to back; *d_span_size its length. */
int fourmc_4mc_compress_span_device(fourmc_ctx *ctx, void *stream, int level,
                                    const void *d_in, size_t n,
                                    void *d_span, size_t span_capacity,
                                    uint64_t *d_span_size, uint32_t *d_block_lens);

/* Footer (and file header / EOS) assembly from block lengths: native/4mc.c:263-274,335-362.
 * d_block_lens[i] = 12 + csize_i for ALL blocks of the file; writes the 12-byte header to
 * d_header (may be NULL), and EOS+footer (12 + 20 + 4*n_blocks bytes) to d_tail. */
int fourmc_4mc_build_index_device(fourmc_ctx *ctx, void *stream, const uint32_t *d_block_lens,
                                  uint32_t n_blocks, void *d_header, void *d_tail);

/* Whole single stream on device (located through its footer index, cross-checked against the
 * block headers and the EOS mark).  d_result (device i64[2]): [0] decoded size or FOURMC_E_*,
 * [1] index of the first failing block or -1. */
int fourmc_4mc_decompress_device(fourmc_ctx *ctx, void *stream, const void *d_in, size_t n,
                                 void *d_out, size_t out_capacity, long long *d_result);

/* A block range of one stream: blocks [first_block, first_block + n_blocks) of the .4mc stream at d_in, located
 * through the footer index like the call above (the whole index is validated); the range's first block decodes
 * to d_out.  This is how the ranks of a multi-GPU reader share ONE stream (contiguous block ranges,
 * no collective); n_blocks = 0xffffffff means "to the end".  d_result as above, sizes and block numbers counted
 * within the range. */
int fourmc_4mc_decompress_range_device(fourmc_ctx *ctx, void *stream, const void *d_in, size_t n,
                                       uint32_t first_block, uint32_t n_blocks,
                                       void *d_out, size_t out_capacity, long long *d_result);

/* Batch of independent blocks.  Block i: payload at d_src + src_off[i] (csize[i] bytes, raw when
 * csize[i] == usize[i]), expected checksum xxh[i] (ignored when check_xxh == 0), output at
 * d_dst + dst_off[i] with capacity usize[i].  The five tables are DEVICE arrays.
 * d_status[i] receives FOURMC_BLOCK_*; d_out_size[i] the decoded size (or the negative
 * LZ4_decompress_safe value). */
int fourmc_lz4_decompress_batch_device(fourmc_ctx *ctx, void *stream, uint32_t n_blocks,
                                       const void *d_src, const uint64_t *d_src_off,
                                       const uint32_t *d_csize, const uint32_t *d_usize,
                                       const uint32_t *d_xxh, int check_xxh,
                                       void *d_dst, const uint64_t *d_dst_off,
                                       int32_t *d_out_size, uint8_t *d_status);

/* XXH32 of n_items device ranges: d_out[i] = XXH32(d_base + d_off[i], d_len[i], seed). */
int fourmc_xxh32_batch_device(fourmc_ctx *ctx, void *stream, uint32_t n_items,
                              const void *d_base, const uint64_t *d_off, const uint32_t *d_len,
                              uint32_t seed, uint32_t *d_out);

/* ---- 4mz (zstd blocks) ------------------------------------------------------------------------ */
/* Same contracts as the 4mc calls above, for streams with the "4MZ\0" magic whose compressed blocks
 * are zstd frames.  Readers: native/4mc.c:709-857 (decodeFourMZ), :810 ZSTD_decompress.  Writers:
 * native/4mc.c:389-553 (fourMZcompressFilename), :467 ZSTD_compress with the stored fallback
 * :469-485.  The frames are valid zstd (ZSTD_decompress restores the input) but not the reference's
 * bytes; levels 2..4 use the hash-chain match finder of the LZ4 levels 2..4. */
long long fourmc_4mz_decompress_host(fourmc_ctx *ctx, const void *in, size_t n, void *out, size_t out_capacity);
long long fourmc_4mz_decoded_size_host(const void *in, size_t n);
int fourmc_4mz_decompress_device(fourmc_ctx *ctx, void *stream, const void *d_in, size_t n,
                                 void *d_out, size_t out_capacity, long long *d_result);
int fourmc_4mz_decompress_range_device(fourmc_ctx *ctx, void *stream, const void *d_in, size_t n,
                                       uint32_t first_block, uint32_t n_blocks,
                                       void *d_out, size_t out_capacity, long long *d_result);
/* ZSTD_decompress on one block (host pointers): native/4mc.c:810, native/jniZstdDecompressor.c.
 * Returns the decoded size, or a negative value where ZSTD_isError() is true for the reference. */
long long fourmc_zstd_decompress(fourmc_ctx *ctx, const void *src, size_t compressed_size,
                                 void *dst, size_t dst_capacity);

/* ZSTD_compressBound: native/jniZstdCompressor.c compressBound, native/zstd/zstd.h:204. */
size_t fourmc_zstd_compress_bound(size_t n);
/* ZSTD_compress on one block (host pointers): native/4mc.c:467, native/jniZstdCompressor.c:93,125,158.
 * Returns the frame size, or -70 (zstd's dstSize_tooSmall, for which ZSTD_isError() is true) when
 * the frame does not fit in dst_capacity -- the caller then stores the block raw (native/4mc.c:469). */
long long fourmc_zstd_compress(fourmc_ctx *ctx, int level, const void *src, size_t src_size,
                               void *dst, size_t dst_capacity);
/* fourmc_4mc_compress_host / _device / _span_device / fourmc_4mc_build_index_device for 4mz. */
long long fourmc_4mz_compress_host(fourmc_ctx *ctx, int level, const void *in, size_t n,
                                   void *out, size_t out_capacity);
int fourmc_4mz_compress_device(fourmc_ctx *ctx, void *stream, int level, const void *d_in, size_t n,
                               void *d_out, size_t out_capacity, uint64_t *d_out_size, uint32_t *d_block_lens);
int fourmc_4mz_compress_span_device(fourmc_ctx *ctx, void *stream, int level, const void *d_in, size_t n,
                                    void *d_span, size_t span_capacity, uint64_t *d_span_size,
                                    uint32_t *d_block_lens);
int fourmc_4mz_build_index_device(fourmc_ctx *ctx, void *stream, const uint32_t *d_block_lens,
                                  uint32_t n_blocks, void *d_header, void *d_tail);

/* ---- block index, splits, line records: the callers either side of the path --------------------
 * (BASELINE.json configs[4]: per-split decode of a .4mc read through the InputFormat) */
#define FOURMC_NOT_FOUND (-1)                         /* FourMcBlockIndex.NOT_FOUND */

/* FourMcInputStream.readIndex (FourMcInputStream.java:163-239): absolute offsets of the block headers
 * from the footer of a whole .4mc / .4mz file in host memory (footer XXH32 verified on the device).
 * Returns the block count (offsets[] receives min(count, cap)), 0 when the file cannot hold an
 * index, FOURMC_E_CONTENT on a damaged footer. */
long long fourmc_read_index_host(fourmc_ctx *ctx, const void *file, size_t file_size, int64_t *offsets, size_t cap);

/* FourMcBlockIndex.java:92-104, :111-124, :142-153, :163-173 -- same results, FOURMC_NOT_FOUND included. */
int64_t fourmc_index_find_next_position(const int64_t *offsets, int n, int64_t pos);
int64_t fourmc_index_find_belonging_block(const int64_t *offsets, int n, int64_t pos);
int64_t fourmc_index_align_slice_start(const int64_t *offsets, int n, int64_t start, int64_t end);
int64_t fourmc_index_align_slice_end(const int64_t *offsets, int n, int64_t end, int64_t file_size);

/* FourMcInputFormat.getSplits for one file (FourMcInputFormat.java:126-173) on top of Hadoop's default
 * byte-range splits of split_size bytes.  Returns the number of splits. */
int fourmc_plan_splits(const int64_t *offsets, int n, int64_t file_size, int64_t split_size,
                       int64_t *starts, int64_t *lengths, int cap);

/* FourMcLineRecordReader over one split (FourMcLineRecordReader.java:116-163): every record (line, with its
 * terminator) the reader returns for [start, start + length), concatenated into out.  The split's blocks
 * (and the block(s) that finish its last line) are decoded on the device.  Returns the byte count.
 * Lines end like Hadoop's LineReader ends them: at LF, at CR, or at CR LF.  Splits are expected as fourmc_plan_splits
 * makes them (both ends on block starts or the end of the file); for other ranges the blocks that START inside the range
 * are read.  One corner is deliberately not reproduced: when a block's last byte is a lone CR, Hadoop 1.x's reader
 * stops there and the next split skips the following line (which is then returned by nobody); here that line is
 * returned with the split that ends at the CR, as for LF. */
long long fourmc_read_split_lines_host(fourmc_ctx *ctx, const void *file, size_t file_size, int64_t start,
                                       int64_t length, void *out, size_t out_capacity);

/* Many splits of one file in one call: the same records per split, the blocks of all the splits decoded as ONE batch
 * (a split alone is two or three blocks -- a latency-bound call; a node that runs many map tasks over one file hands
 * their splits over together).  The records of split i land at out + out_offsets[i], back to back in split order;
 * out_offsets has n_splits + 1 entries, the last one being the total, which is also the return value. */
long long fourmc_read_splits_lines_host(fourmc_ctx *ctx, const void *file, size_t file_size, int n_splits,
                                        const int64_t *starts, const int64_t *lengths,
                                        void *out, size_t out_capacity, int64_t *out_offsets);

/* ---- synthetic inputs, bit-identical on host and device ---------------------- */

/* kind 0 = log-text (configs[0], [1]), 1 = JSON lines (configs[2]), 2 = silesia-like mix of block types
 * incl. incompressible blocks (configs[3]).  Fills pages [first_page, first_page + n_pages) of 4096
 * bytes each; a page is a pure function of (kind, seed, page index). */
int fourmc_gen_device(fourmc_ctx *ctx, void *stream, int kind, uint64_t seed,
                      uint64_t first_page, uint64_t n_pages, void *d_out);
int fourmc_gen_host(int kind, uint64_t seed, uint64_t first_page, uint64_t n_pages, void *out);

/* ---- raw codec streams (Hadoop block-stream framing): Lz4Codec / ZstdCodec and their level variants ----------
 * Lz4Codec.java:95-104 (createOutputStream -> BlockCompressorStream(out, compressor, 4 MiB, compressBound(4 MiB) -
 * 4 MiB)), :128-138 (createInputStream -> BlockDecompressorStream(in, decompressor, 4 MiB)); ZstdCodec.java:103-112,
 * :136-146.  Wire format and writer / reader rules: 4mc_b200/csrc/blockstream.h.  `zstd`: 0 = LZ4 chunks, 1 = zstd
 * frames; level 1..4 as in the per-block calls; write_size = the size of the application's write() calls, which
 * decides where blocks are cut (0 = one write with everything).  Host pointers.  Return the stream / decoded size
 * or a negative FOURMC_E_*. */
size_t fourmc_blockstream_bound(int zstd, size_t n, size_t write_size);
long long fourmc_blockstream_compress_host(fourmc_ctx *ctx, int zstd, int level, const void *in, size_t n,
                                           size_t write_size, void *out, size_t out_capacity);
long long fourmc_blockstream_decompress_host(fourmc_ctx *ctx, int zstd, const void *in, size_t n,
                                             void *out, size_t out_capacity);

/* ---- files: the reference's own library entry points (native/4mc.h:36-41, callers native/4mccli.c:342-356) -----
 * Same arguments, console messages per display level, "stdin" / "stdout" / "/dev/null" names and overwrite
 * prompt as native/4mc.c:163-211, :220-386, :388-556, :896-966 -- and, like the reference, a fatal error ends
 * the process with exit(1 generic | 2 input | 3 output | 4 content) (native/4mc.c:135-161); 0 on success.
 * Files are streamed through pinned bounce buffers in slices of many blocks (a reader thread, the calling
 * thread on the GPU, a writer thread): any size, pipes included.  The process-wide context they share is
 * created on first use.  The reference's native/4mccli.c links against these unchanged. */
int fourMCcompressFilename(int displayLevel, int overwrite, char *input_filename, char *output_filename, int compressionlevel);
int fourMcDecompressFileName(int displayLevel, int overwrite, char *input_filename, char *output_filename);
int fourMZcompressFilename(int displayLevel, int overwrite, char *input_filename, char *output_filename, int compressionlevel);
int fourMZDecompressFileName(int displayLevel, int overwrite, char *input_filename, char *output_filename);

/* The bodies of those four over open descriptors, with status codes instead of exit(): bytes written /
 * decoded, or FOURMC_E_INPUT (read error, truncated stream), FOURMC_E_OUTPUT (write error), FOURMC_E_CONTENT,
 * FOURMC_E_GENERIC.  Decoding writes every block that precedes a damaged one before it reports the damage,
 * like the serial reader (native/4mc.c:603-668).  *in_bytes (may be NULL) = bytes consumed from in_fd. */
long long fourmc_compress_fd(fourmc_ctx *ctx, int zstd, int level, int in_fd, int out_fd, uint64_t *in_bytes);
long long fourmc_decompress_fd(fourmc_ctx *ctx, int zstd, int in_fd, int out_fd, uint64_t *in_bytes);

#ifdef __cplusplus
}
#endif
#endif
