// blockstream.h -- the wire format of the reference's RAW codecs: Lz4Codec / Lz4MediumCodec /
// Lz4HighCodec / Lz4UltraCodec and ZstdCodec / ... wrap the same per-block natives as the 4mc container, but frame
// them with Hadoop's BlockCompressorStream / BlockDecompressorStream (Lz4Codec.java:95-104, :128-138;
// ZstdCodec.java:103-112, :136-146) instead of the 4mc header / footer: non-splittable `.lz4_fast`, `.zstd_fast` ... files.
//
// Hadoop (org.apache.hadoop:hadoop-core 1.1.2, java/hadoop-4mc/pom.xml:81-104) is NOT vendored in the reference
// repository; the framing below restates its published BlockCompressorStream / BlockDecompressorStream
// (org/apache/hadoop/io/compress/) as driven by the reference's Lz4Compressor / Lz4Decompressor state machines
// (Lz4Compressor.java:143-256, Lz4Decompressor.java).  PARITY UNPINNED by vectors: there is no JVM in this image and
// the reference's tests hold no file of this format; the pins are the call sites above and the restated rules.
//
//   STREAM := BLOCK* [00 00 00 00]
//   BLOCK  := BE32 rawLen  CHUNK+          rawLen = uncompressed bytes of the block (sum over its chunks)
//   CHUNK  := BE32 cLen    bytes[cLen]     one LZ4 block / one zstd frame, NEVER stored raw, no checksum
//
// Writer rules (BlockCompressorStream.write / finish / compress), bufferSize = 4 MiB, compressionOverhead =
// compressBound(4 MiB) - 4 MiB, MAX_INPUT_SIZE = bufferSize - compressionOverhead (4 177 840 LZ4, 4 177 920 zstd):
//   * write(len): if bytes already buffered and len + buffered > MAX_INPUT_SIZE -> finish() the block first;
//     len > MAX_INPUT_SIZE -> rawLen = len, then one chunk per MAX_INPUT_SIZE bytes; otherwise the bytes join the
//     compressor's 4 MiB direct buffer (Lz4Compressor.setInput, :143-161).
//   * finish(): if the compressor is not finished -> rawLen = buffered bytes (0 when nothing is buffered: an empty
//     stream, or a stream whose last write was a large one, ends with 00 00 00 00), then its chunk.
// Reader rules (BlockDecompressorStream.decompress / getCompressedData): read rawLen, then chunks until rawLen bytes
// came out; a chunk is decoded with capacity directBufferSize = 4 MiB; rawLen == 0 or end of input at a block
// boundary ends the stream; end of input anywhere else is an error (EOFException).
//
// Host-only, codec-agnostic (the codec is two callbacks): the C-ABI binds the GPU per-block / batch calls
// (capi.cu), tests/native/bs_emul.cpp binds the oracle's CPU codec to check the framing without a GPU.
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace fbs {

constexpr uint32_t BUFFER = 4u * 1024 * 1024;              // Lz4Codec.java:54 LZ4_BUFFER_SIZE, ZstdCodec.java:62
constexpr long long E_INPUT = -2, E_OUTPUT = -3, E_CONTENT = -4, E_ARG = -11;     // = FOURMC_E_*

struct Codec {
    void *user;
    uint32_t bound_of_buffer;                              // compressBound(4 MiB): Lz4Codec.java:102, ZstdCodec.java:110
    // one block through the per-block native: compressed size > 0, or <= 0 on failure (InternalError in Java)
    long long (*compress)(void *user, int level, const uint8_t *src, uint32_t n, uint8_t *dst, size_t cap);
    // decoded size >= 0, negative on failure
    long long (*decompress)(void *user, const uint8_t *src, uint32_t c, uint8_t *dst, uint32_t cap);
};

inline uint32_t max_input(const Codec &c) { return BUFFER - (c.bound_of_buffer - BUFFER); }

inline void put32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
inline uint32_t get32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// The blocks a writer makes when the application hands it `n` bytes in write() calls of `write_size` bytes
// (0 = one call with everything): {offset, rawLen} per block, in order.  `trailing_zero` = the stream ends with a
// zero rawLen (nothing buffered at close()).
struct Block { size_t off; uint32_t raw; };
inline void plan_blocks(const Codec &c, size_t n, size_t write_size, std::vector<Block> &blocks, bool *trailing_zero)
{
    const uint32_t MAX = max_input(c);
    if (write_size == 0 || write_size > n) write_size = n;
    size_t pos = 0, block_start = 0, buffered = 0;
    bool open = false;                       // bytes buffered in the compressor since its last reset
    while (pos < n) {
        const size_t len = std::min(write_size, n - pos);
        if (open && len + buffered > MAX) {                                     // finish(); compressor.reset()
            blocks.push_back(Block{block_start, (uint32_t)buffered});
            open = false; buffered = 0;
        }
        if (len > MAX) {                                                        // written out at once, in MAX pieces
            blocks.push_back(Block{pos, (uint32_t)len});
        } else {
            if (!open) { block_start = pos; open = true; }
            buffered += len;
        }
        pos += len;
    }
    if (open) blocks.push_back(Block{block_start, (uint32_t)buffered});
    *trailing_zero = !open;
}

// Worst-case size of the stream plan_blocks + compress produce.  A block is cut before the write that would overflow
// MAX_INPUT_SIZE, so it holds at least max(write_size, MAX_INPUT_SIZE - write_size + 1) >= MAX_INPUT_SIZE / 2 bytes
// (a large write is one block of ceil(len / MAX_INPUT_SIZE) chunks): at most 2n / MAX + 2 blocks whatever the write size.
inline size_t bound(const Codec &c, size_t n, size_t write_size)
{
    (void)write_size;
    const uint32_t MAX = max_input(c);
    const size_t nblocks = 2 * (n / MAX) + 3, nchunks = n / MAX + nblocks + 1;
    return n + n / 255 + (n >> 8) + 4 + nblocks * 4 + nchunks * (4 + 64 + 16);
}

inline long long compress(const Codec &c, int level, const uint8_t *in, size_t n, size_t write_size, uint8_t *out, size_t cap)
{
    if ((!in && n) || !out) return E_ARG;
    if (n > 0xffffffffu && (write_size == 0 || write_size > 0x7fffffffu)) return E_ARG;      // write(byte[], int, int)
    const uint32_t MAX = max_input(c);
    std::vector<Block> blocks;
    bool trailing_zero = false;
    plan_blocks(c, n, write_size, blocks, &trailing_zero);
    size_t op = 0;
    for (const Block &b : blocks) {
        if (cap - op < 4) return E_OUTPUT;
        put32(out + op, b.raw); op += 4;
        for (uint32_t done = 0; done < b.raw;) {
            const uint32_t piece = std::min(MAX, b.raw - done);
            if (cap - op < 4) return E_OUTPUT;
            const long long r = c.compress(c.user, level, in + b.off + done, piece, out + op + 4, cap - op - 4);
            if (r <= 0) return r < 0 ? r : E_OUTPUT;
            put32(out + op, (uint32_t)r);
            op += 4 + (size_t)r;
            done += piece;
        }
    }
    if (trailing_zero) {
        if (cap - op < 4) return E_OUTPUT;
        put32(out + op, 0); op += 4;
    }
    return (long long)op;
}

// Serial reader: exactly BlockDecompressorStream's loop.
inline long long decompress(const Codec &c, const uint8_t *in, size_t n, uint8_t *out, size_t cap)
{
    if ((!in && n) || (!out && cap)) return E_ARG;
    size_t ip = 0, op = 0;
    for (;;) {
        if (n - ip < 4) return (long long)op;                 // rawReadInt fails at a block boundary: end of stream
        const uint32_t raw = get32(in + ip); ip += 4;
        if (raw == 0) return (long long)op;                   // a zero-length block ends the stream
        for (uint32_t got = 0; got < raw;) {
            if (n - ip < 4) return E_INPUT;                   // EOFException inside a block
            const uint32_t clen = get32(in + ip); ip += 4;
            if (clen > BUFFER) return E_CONTENT;              // does not fit the decompressor's 4 MiB direct buffer
            if (n - ip < clen) return E_INPUT;
            if (clen == 0) continue;                          // getCompressedData with len 0: nothing to decode
            const uint32_t room = (uint32_t)std::min<size_t>(BUFFER, cap - op);
            const long long r = c.decompress(c.user, in + ip, clen, out + op, room);
            if (r < 0) {
                // the Java reader always offers 4 MiB; with less room here, find out whether the room was the problem
                if (room == BUFFER) return E_CONTENT;
                std::vector<uint8_t> full(BUFFER);
                return c.decompress(c.user, in + ip, clen, full.data(), BUFFER) >= 0 ? E_OUTPUT : E_CONTENT;
            }
            ip += clen; op += (size_t)r; got += (uint32_t)r;
            if (r == 0 && clen) return E_CONTENT;             // no progress: the Java loop would spin on needsInput
        }
    }
}

// What a reader can know without decoding, ASSUMING the writer above made the stream: chunk i of a block holds
// min(MAX_INPUT_SIZE, rawLen - i * MAX_INPUT_SIZE) bytes.  Lets a batch decoder run every chunk at once; a stream
// for which the prediction fails (return false, or a chunk that decodes to another size) goes through decompress().
struct Chunk { size_t src_off, dst_off; uint32_t clen, usize; };
inline bool predict_chunks(const Codec &c, const uint8_t *in, size_t n, std::vector<Chunk> &chunks, size_t *total)
{
    const uint32_t MAX = max_input(c);
    size_t ip = 0, op = 0;
    for (;;) {
        if (n - ip < 4) break;
        const uint32_t raw = get32(in + ip); ip += 4;
        if (raw == 0) break;
        for (uint32_t got = 0; got < raw;) {
            if (n - ip < 4) return false;
            const uint32_t clen = get32(in + ip); ip += 4;
            if (clen == 0 || clen > BUFFER || n - ip < clen) return false;
            const uint32_t u = std::min(MAX, raw - got);
            chunks.push_back(Chunk{ip, op, clen, u});
            ip += clen; op += u; got += u;
        }
    }
    *total = op;
    return true;
}

}  // namespace fbs
