// zstd_encode.h -- Zstandard block encoder for 4mz blocks ("4mz Fast"), host + device.
//
// Reference behaviour being replaced: ZSTD_compress(dst, cap, src, n, level) as called per 4 MiB
// block at native/4mc.c:467 and native/jniZstdCompressor.c:93 (zstd 1.5.3 as vendored):
//   frame            native/zstd/compress/zstd_compress.c:4065-4114 (header), :3983-4063 (block loop)
//   literals         native/zstd/compress/zstd_compress_literals.c:95-187, huf_compress.c
//   sequences        native/zstd/compress/zstd_compress_sequences.c:290-383 (bit order), :156-240 (modes)
//   FSE tables       native/zstd/common/fse_compress.c (normalise, header, encoding table)
// Compressed bytes need not match the reference (BASELINE.json north_star): only a valid frame
// that ZSTD_decompress turns back into the input.  So the encoder is shaped for the GPU and uses
// the subset of the format that parallelises (SURVEY.md Appendix C "Encoder freedom"):
//   * one zstd block per 64 KiB REGION of the 4 MiB 4mz block (64 per frame); the LZ sequences of
//     a region come from the same shared-memory match finder as the LZ4 path (lz4_encode.cuh);
//   * literals: raw, RLE, or Huffman with a fresh tree per block (direct 4-bit weights), 4 streams;
//   * sequences: per table predefined / RLE / FSE-described, never "repeat"; offsets are real
//     offsets or "same offset as the previous sequence of this block" (the one repeat code that
//     does not depend on earlier blocks), so blocks do not depend on each other;
//   * single-segment frame header with a 4-byte content size, no checksum.
// Everything that writes output bits works on a zeroed buffer with OR semantics, so any number
// of threads can assemble one bitstream from prefix-summed bit positions.
//
// The region encoder is written as a sequence of PHASES separated by CTA barriers
// (zenc_region<Exec>): on the GPU a phase is "every thread runs the body, then __syncthreads()";
// tests/native/zenc_emul.cpp runs the same source with a loop over thread ids, so the whole
// encoder is checked against ZSTD_decompress on a machine without a GPU.
#pragma once

#include <stdint.h>

#include "zstd_decode.h"

namespace fmz {

constexpr int ZE_REGION = 65536;               // input bytes per zstd block (32768 under the chain parse of levels 2..4)
constexpr int ZE_THREADS = 128;                // threads of the entropy CTA
constexpr int ZE_WARPS = ZE_THREADS / 32;
#ifndef FOURMC_ZE_TILE
#define FOURMC_ZE_TILE 512
#endif
constexpr int ZE_TILE = FOURMC_ZE_TILE;        // sequences per pass of the sequence encoder
constexpr int ZE_PER_THREAD = ZE_TILE / ZE_THREADS;
constexpr int ZE_HUF_MAXBITS = 11;             // zstd_compress_literals.c: LitHufLog
constexpr int ZE_MIN_HUF_LITS = 256;           // below: raw literals
constexpr int ZE_MIN_FSE_SEQ = 64;             // below: predefined tables
// Region scratch written by the match finder: u16 ll[n] | u16 ml[n] | u16 off[n] | literals,
// n rounded up to 8.  6 n + literals <= 65536 + 2 n <= 98304 (each sequence covers >= 4 bytes).
constexpr int ZE_IN_SLOT = 98304 + 64;
constexpr int ZE_OUT_SLOT = ZE_REGION + 64;    // a block body that does not fit is emitted raw
FZ_HD inline uint32_t ze_in_slot(uint32_t region) { return region + region / 2 + 64; }
FZ_HD inline uint32_t ze_out_slot(uint32_t region) { return region + 64; }

FZ_HD inline uint32_t ze_seq_stride(uint32_t nseq) { return (nseq + 7u) & ~7u; }

// ---- OR-writes into a zeroed little-endian bit array -------------------------------------------

FZ_HD inline void or32(uint32_t *p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}

// nb <= 32 bits of v (no bits above nb set) at absolute bit position g
FZ_HD inline void put_bits(uint32_t *base, uint64_t g, uint32_t v, int nb)
{
    if (nb == 0) return;
    const uint32_t w = (uint32_t)(g >> 5), s = (uint32_t)(g & 31);
    or32(base + w, v << s);
    if (s + (uint32_t)nb > 32) or32(base + w + 1, v >> (32 - s));
}

FZ_HD inline void put_byte(uint32_t *base, uint32_t byte_pos, uint32_t v) { put_bits(base, (uint64_t)byte_pos * 8, v & 255u, 8); }

// A run of consecutive bit fields written by one thread: the first and the last (partial) word go
// through or32 because neighbours share them, the words in between are owned and stored plainly.
struct BitRun {
    uint32_t *base;
    uint32_t w;
    uint64_t acc;
    int n;
    bool first;
    FZ_HD void start(uint32_t *b, uint64_t g) { base = b; w = (uint32_t)(g >> 5); n = (int)(g & 31); acc = 0; first = true; }
    FZ_HD void add(uint32_t v, int nb)
    {
        acc |= (uint64_t)v << n;
        n += nb;
        if (n >= 32) {
            if (first) { or32(base + w, (uint32_t)acc); first = false; } else base[w] = (uint32_t)acc;
            w++; acc >>= 32; n -= 32;
        }
    }
    FZ_HD void finish() { if (n > 0) or32(base + w, (uint32_t)acc); }
};

// ---- symbol codes (zstd_internal.h:121-165, zstd_compress_internal.h ZSTD_LLcode / ZSTD_MLcode) --

FZ_HD inline uint32_t ll_code(uint32_t ll)
{
    if (ll < 16) return ll;
    if (ll < 64) {
        // 16-17:16 18-19:17 20-21:18 22-23:19 24-27:20 28-31:21 32-39:22 40-47:23 48-63:24
        if (ll < 24) return 16 + ((ll - 16) >> 1);
        if (ll < 32) return 20 + ((ll - 24) >> 2);
        if (ll < 48) return 22 + ((ll - 32) >> 3);
        return 24;
    }
    return (uint32_t)highbit(ll) + 19;
}

FZ_HD inline uint32_t ml_code(uint32_t mlbase)          // mlbase = match length - 3
{
    if (mlbase < 32) return mlbase;
    if (mlbase < 128) {
        // 32-33:32 34-35:33 36-37:34 38-39:35 40-43:36 44-47:37 48-55:38 56-63:39 64-79:40 80-95:41 96-127:42
        if (mlbase < 40) return 32 + ((mlbase - 32) >> 1);
        if (mlbase < 48) return 36 + ((mlbase - 40) >> 2);
        if (mlbase < 64) return 38 + ((mlbase - 48) >> 3);
        if (mlbase < 96) return 40 + ((mlbase - 64) >> 4);
        return 42;
    }
    return (uint32_t)highbit(mlbase) + 36;
}

// ---- FSE: normalisation, table description, encoding table -----------------------------------------

struct FseCTable {                    // fse_compress.c FSE_buildCTable_wksp
    uint16_t state[512];              // next-state table, indexed by cumulative rank
    int32_t dnb[53];                  // deltaNbBits per symbol
    int16_t dfs[53];                  // deltaFindState per symbol
    int32_t log;                      // 0 = RLE (no state bits at all)
};

// Counts -> probabilities that sum to 1 << log, every present symbol >= 1.  (Any distribution with
// these two properties is a valid table description; this one is proportional with the rounding
// error given to / taken from the largest symbols.)
FZ_HD inline void fse_normalize(short *norm, int log, const uint32_t *count, uint32_t total, int max_sym)
{
    const int size = 1 << log;
    int sum = 0;
    for (int s = 0; s <= max_sym; s++) {
        int p = 0;
        if (count[s]) {
            p = (int)(((uint64_t)count[s] * (uint32_t)size + total / 2) / total);
            if (p < 1) p = 1;
        }
        norm[s] = (short)p;
        sum += p;
    }
    int diff = size - sum;
    while (diff != 0) {
        int big = 0;
        for (int s = 1; s <= max_sym; s++) if (norm[s] > norm[big]) big = s;
        if (diff > 0) { norm[big] = (short)(norm[big] + diff); diff = 0; }
        else {
            int take = norm[big] - 1;
            if (take > -diff) take = -diff;
            if (take > (norm[big] + 1) / 2) take = (norm[big] + 1) / 2;     // spread a large deficit
            if (take < 1) break;                                            // cannot happen: symbols <= size
            norm[big] = (short)(norm[big] - take);
            diff += take;
        }
    }
}

// FSE table description (entropy_common.c:43-213 reads it).  Returns the byte count.
FZ_HD inline int fse_write_ncount(uint8_t *out, const short *norm, int max_sym, int log)
{
    const int size = 1 << log;
    uint64_t bits = 0;
    int nbit = 0, pos = 0;
    auto flush16 = [&]() { while (nbit >= 16) { out[pos++] = (uint8_t)bits; out[pos++] = (uint8_t)(bits >> 8); bits >>= 16; nbit -= 16; } };
    bits |= (uint64_t)(log - 5);
    nbit = 4;
    int remaining = size + 1, threshold = size, nb = log + 1;
    int sym = 0;
    bool prev0 = false;
    const int alphabet = max_sym + 1;
    while (sym < alphabet && remaining > 1) {
        if (prev0) {
            int start = sym;
            while (sym < alphabet && norm[sym] == 0) sym++;
            if (sym == alphabet) break;
            while (sym >= start + 24) { start += 24; bits |= (uint64_t)0xFFFF << nbit; nbit += 16; flush16(); }
            while (sym >= start + 3) { start += 3; bits |= (uint64_t)3 << nbit; nbit += 2; }
            bits |= (uint64_t)(sym - start) << nbit;
            nbit += 2;
            flush16();
        }
        {
            int count = norm[sym++];
            const int max = (2 * threshold - 1) - remaining;
            remaining -= count < 0 ? -count : count;
            count++;
            if (count >= threshold) count += max;
            bits |= (uint64_t)(uint32_t)count << nbit;
            nbit += nb;
            nbit -= (count < max);
            prev0 = (count == 1);
            while (remaining < threshold) { nb--; threshold >>= 1; }
        }
        flush16();
    }
    while (nbit > 0) { out[pos++] = (uint8_t)bits; bits >>= 8; nbit -= 8; }
    return pos;
}

// cells: scratch of 1 << log bytes
FZ_HD inline void fse_build_ctable(FseCTable &ct, const short *norm, int max_sym, int log, uint8_t *cells)
{
    const int size = 1 << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    uint16_t cumul[54];
    int high = size - 1;
    cumul[0] = 0;
    for (int s = 0; s <= max_sym; s++) {
        if (norm[s] == -1) { cumul[s + 1] = (uint16_t)(cumul[s] + 1); cells[high--] = (uint8_t)s; }
        else cumul[s + 1] = (uint16_t)(cumul[s] + norm[s]);
    }
    int pos = 0;
    for (int s = 0; s <= max_sym; s++)
        for (int i = 0; i < norm[s]; i++) {
            cells[pos] = (uint8_t)s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    for (int u = 0; u < size; u++) { const int s = cells[u]; ct.state[cumul[s]++] = (uint16_t)(size + u); }
    int total = 0;
    for (int s = 0; s <= max_sym; s++) {
        const int n = norm[s];
        if (n == 0) { ct.dnb[s] = ((log + 1) << 16) - size; ct.dfs[s] = 0; }
        else if (n == -1 || n == 1) { ct.dnb[s] = (log << 16) - size; ct.dfs[s] = (int16_t)(total - 1); total++; }
        else {
            const int maxbits = log - highbit((uint32_t)(n - 1));
            ct.dnb[s] = (maxbits << 16) - (n << maxbits);
            ct.dfs[s] = (int16_t)(total - n);
            total += n;
        }
    }
    ct.log = log;
}

// ---- Huffman: code lengths (<= 11 bits), canonical codes, tree description ---------------------------

struct HufBuild {                     // scratch of the single thread that builds the tree
    uint32_t weight[512];
    uint16_t parent[512];
    uint8_t depth[512];
};

// sorted[0..n): symbols with count > 0 in ascending (count, symbol) order, n >= 2.
// Writes nbits[] (0 for absent symbols) and returns the table log (longest code).
FZ_HD inline int huf_build_lengths(HufBuild &hb, const uint32_t *count, const uint16_t *sorted, int n, uint8_t *nbits)
{
    for (int i = 0; i < n; i++) hb.weight[i] = count[sorted[i]];
    int li = 0, ni = n, nn = n;
    for (int k = 0; k < n - 1; k++) {
        int a, b;
        if (li < n && (ni >= nn || hb.weight[li] <= hb.weight[ni])) a = li++; else a = ni++;
        if (li < n && (ni >= nn || hb.weight[li] <= hb.weight[ni])) b = li++; else b = ni++;
        hb.weight[nn] = hb.weight[a] + hb.weight[b];
        hb.parent[a] = hb.parent[b] = (uint16_t)nn;
        nn++;
    }
    const int root = 2 * n - 2;
    hb.depth[root] = 0;
    uint32_t bl[64];
    for (int i = 0; i < 64; i++) bl[i] = 0;
    for (int i = root - 1; i >= 0; i--) {
        int d = hb.depth[hb.parent[i]] + 1;
        if (d > 63) d = 63;                                   // cannot happen below 2^40 symbols
        hb.depth[i] = (uint8_t)d;
        if (i < n) bl[d]++;
    }
    // length limit: move the deepest leaf pairs up, pushing one shorter leaf down (Kraft sum stays 1)
    for (int i = 63; i > ZE_HUF_MAXBITS; i--) {
        while (bl[i] > 0) {
            int j = i - 2;
            while (bl[j] == 0) j--;
            bl[i] -= 2; bl[i - 1] += 1; bl[j + 1] += 2; bl[j] -= 1;
        }
    }
    // smallest counts get the longest codes
    int idx = 0, log = 0;
    for (int len = ZE_HUF_MAXBITS; len >= 1; len--) {
        if (bl[len] && !log) log = len;
        for (uint32_t k = 0; k < bl[len]; k++) nbits[sorted[idx++]] = (uint8_t)len;
    }
    return log;
}

// Canonical codes in the decoder's order (huf_decompress.c:344-480, zstd_decode.h read_huffman):
// longest codes first, symbol order inside a length.  code[s] = value | nbits << 16.
FZ_HD inline void huf_assign_codes(uint32_t *code, const uint8_t *nbits, int max_sym, int log)
{
    uint32_t per_len[16], val[16];
    for (int i = 0; i < 16; i++) per_len[i] = 0;
    for (int s = 0; s <= max_sym; s++) per_len[nbits[s]]++;
    uint32_t min = 0;
    for (int n = log; n > 0; n--) { val[n] = min; min += per_len[n]; min >>= 1; }
    for (int s = 0; s <= max_sym; s++) {
        const int nb = nbits[s];
        code[s] = nb ? (val[nb]++ | ((uint32_t)nb << 16)) : 0u;
    }
    for (int s = max_sym + 1; s < 256; s++) code[s] = 0;
}

// Direct tree description (huf_compress.c HUF_writeCTable, raw 4-bit weights): only for
// max_sym <= 128.  Returns the byte count.
FZ_HD inline int huf_write_direct(uint8_t *out, const uint8_t *nbits, int max_sym, int log)
{
    out[0] = (uint8_t)(128 + (max_sym - 1));
    for (int n = 0; n < max_sym; n += 2) {
        const int w0 = nbits[n] ? log + 1 - nbits[n] : 0;
        const int w1 = (n + 1 < max_sym && nbits[n + 1]) ? log + 1 - nbits[n + 1] : 0;
        out[1 + n / 2] = (uint8_t)((w0 << 4) | w1);
    }
    return 1 + (max_sym + 1) / 2;
}

// FSE-compressed tree description (huf_compress.c HUF_compressWeights; read by
// entropy_common.c:259-275 / fse_decompress.c:232-300: table log <= 6, two interleaved states).
// weights[0..n) with n = max_sym (the last symbol's weight is implied).  Returns the byte count
// including the leading size byte, or 0 when this form is not usable (then: direct or raw).
FZ_HD inline int huf_write_fse(uint8_t *out, const uint8_t *weights, int n, FseCTable &ct, uint8_t *cells)
{
    if (n < 2) return 0;
    uint32_t count[16];
    for (int i = 0; i < 16; i++) count[i] = 0;
    int max_w = 0;
    for (int i = 0; i < n; i++) { count[weights[i]]++; if (weights[i] > max_w) max_w = weights[i]; }
    for (int i = 0; i <= max_w; i++) if (count[i] == (uint32_t)n) return 0;      // one weight only: RLE is not expressible here
    int log = 6;
    { const int mb = highbit((uint32_t)(n - 1)) - 2; if (mb < log) log = mb; }
    { int a = highbit((uint32_t)n) + 1, b = highbit((uint32_t)max_w) + 2; int minb = a < b ? a : b; if (log < minb) log = minb; }
    if (log < 5) log = 5;
    if (log > 6) log = 6;
    short norm[16];
    fse_normalize(norm, log, count, (uint32_t)n, max_w);
    int pos = 1 + fse_write_ncount(out + 1, norm, max_w, log);
    fse_build_ctable(ct, norm, max_w, log, cells);
    // two states, symbols taken from the end (fse_compress.c FSE_compress_usingCTable_generic)
    uint8_t *bs = out + pos;
    uint64_t acc = 0;
    int nacc = 0, bp = 0;
    auto add = [&](uint32_t v, int nb) {
        acc |= (uint64_t)(v & ((1u << nb) - 1)) << nacc;
        nacc += nb;
        while (nacc >= 8) { bs[bp++] = (uint8_t)acc; acc >>= 8; nacc -= 8; }
    };
    auto init = [&](uint32_t &st, int sym) {
        const uint32_t nbout = (uint32_t)(ct.dnb[sym] + (1 << 15)) >> 16;
        st = (nbout << 16) - (uint32_t)ct.dnb[sym];
        st = ct.state[(st >> nbout) + ct.dfs[sym]];
    };
    auto enc = [&](uint32_t &st, int sym) {
        const uint32_t nbout = (st + (uint32_t)ct.dnb[sym]) >> 16;
        add(st, (int)nbout);
        st = ct.state[(st >> nbout) + ct.dfs[sym]];
    };
    int ip = n;
    uint32_t s1, s2;
    if (n & 1) { init(s1, weights[--ip]); init(s2, weights[--ip]); enc(s1, weights[--ip]); }
    else { init(s2, weights[--ip]); init(s1, weights[--ip]); }
    while (ip > 0) { enc(s2, weights[--ip]); if (ip > 0) enc(s1, weights[--ip]); }
    // n even: pairs remain; the loop above alternates s2, s1 like the reference's main loop
    add(s2, log); add(s1, log);
    add(1, 1);
    if (nacc > 0) { bs[bp++] = (uint8_t)acc; }
    pos += bp;
    if (pos - 1 >= 128) return 0;
    out[0] = (uint8_t)(pos - 1);
    return pos;
}

// ---- per-region encoder --------------------------------------------------------------------------

struct alignas(16) Z16 { uint32_t a, b, c, d; };

struct ZRegionIn {
    const uint16_t *ll, *ml, *off;    // nseq entries each
    const uint8_t *lits;              // nlit bytes: literals of all sequences, then the tail literals
    uint32_t nseq, nlit;
    uint32_t rlen;                    // region length in input bytes
    uint32_t cap;                     // capacity of the output slot's body area (the region size: a larger body is useless)
};

// Offset as coded (zstd_compress_internal.h: offBase).  3 + offset for a real offset; 1 = "the
// previous sequence's offset again" (repeat code 0, needs a non-empty literal run).  Only that one
// repeat code is used: right after any sequence of the same block the decoder's newest history
// entry IS that sequence's offset, whatever the earlier blocks did -- so blocks stay independent.
FZ_HD inline uint32_t ze_off_base(const uint16_t *ll, const uint16_t *off, uint32_t i)
{
    return (i > 0 && ll[i] > 0 && off[i] == off[i - 1]) ? 1u : (uint32_t)off[i] + 3u;
}

struct ZRegionOut {
    uint32_t bytes;                   // body size of the compressed block (literals + sequences sections)
    uint32_t raw;                     // 1: emit the region as a raw block instead
};

struct ZShared {
    uint32_t hist[256];               // literal histogram, then Huffman codes (value | nbits << 16)
    uint16_t sorted[256];
    uint8_t nbits[256];
    uint32_t chist[3][64];            // LL, OF, ML code histograms
    short norm[3][64];
    FseCTable ct[3];
    union {                           // two working sets that never live at the same time
        struct {                      // P3..P5: building the Huffman code and the FSE tables, their descriptions
            HufBuild hb;
            uint8_t cells[3][512];
            uint8_t ncount[3][128];
            uint8_t hufdesc[160];
            uint8_t wts[256];         // Huffman weights of symbols 0 .. max_sym-1
            FseCTable wct;            // FSE table of the weights (tree description)
            uint8_t wcells[64];
        } b;
        struct {                      // the sequence tiles
            uint8_t code[2][3][ZE_TILE];      // double-buffered: the next tile's codes are made while this one is packed
            uint16_t fse[3][ZE_TILE];         // value | nbits << 10, in encoding order inside the tile
        } t;
    } u;
    uint32_t scan[ZE_THREADS];
    uint32_t scan_tmp[ZE_WARPS];
    uint32_t ll_base[36], ml_base[53];        // copies of the format constants (the global ones cost a cache miss per use)
    uint8_t ll_bits[36], ml_bits[53];
    uint32_t state[3];
    int32_t mode[3], log[3], max_code[3], ncount_len[3];
    int32_t nsym, max_sym, huf_log, hufdesc_len;
    int32_t lit_mode;                 // 0 raw, 1 RLE, 2 Huffman
    uint32_t lit_hdr, lit_section;    // header bytes, whole section bytes
    uint32_t stream_off[4], stream_bits[4];
    uint32_t seq_hdr_off, bits_base;  // byte offsets in the slot
    uint32_t bitpos;                  // bits of the sequence stream written so far
    uint32_t scan_total;
    int32_t fail;
    int32_t pending;                  // the last packed tile's bits are not yet added to bitpos
};

// Executors: the GPU one (zstd_encode.cuh) runs a phase on every thread and ends it with a CTA
// barrier; the emulation one loops over thread ids.  excl_scan turns arr[ZE_THREADS] into its
// exclusive prefix sums in place and stores the total.

template <class Exec>
FZ_HD inline void zenc_region(Exec &ex, ZShared &sh, const ZRegionIn &in, uint32_t *slot, ZRegionOut *out, const Tables &T)
{
    const uint32_t nseq = in.nseq, nlit = in.nlit;
    const uint32_t cap_bits = in.cap * 8;                       // a body beyond the region size is useless anyway

    // ---- P0: zero the slot and the histograms
    ex.phase([&](int tid) {
        Z16 *z = (Z16 *)slot;
        for (int i = tid; i < (int)((in.cap + 64) / 16); i += ZE_THREADS) z[i] = Z16{0, 0, 0, 0};
        for (int i = tid; i < 256; i += ZE_THREADS) { sh.hist[i] = 0; sh.nbits[i] = 0; }
        for (int i = tid; i < 3 * 64; i += ZE_THREADS) sh.chist[i / 64][i % 64] = 0;
        if (tid < 36) { sh.ll_base[tid] = T.ll_base[tid]; sh.ll_bits[tid] = T.ll_bits[tid]; }
        if (tid < 53) { sh.ml_base[tid] = T.ml_base[tid]; sh.ml_bits[tid] = T.ml_bits[tid]; }
        if (tid == 0) { sh.nsym = 0; sh.max_sym = 0; sh.fail = 0; sh.bitpos = 0; sh.max_code[0] = sh.max_code[1] = sh.max_code[2] = 0; }
    });

    // ---- P1: histograms of the literals and of the three sequence codes
    ex.phase([&](int tid) {
        for (uint32_t i = (uint32_t)tid; i < nlit; i += ZE_THREADS) ex.add32(&sh.hist[in.lits[i]], 1);
        for (uint32_t i = (uint32_t)tid; i < nseq; i += ZE_THREADS) {
            ex.add32(&sh.chist[0][ll_code(in.ll[i])], 1);
            ex.add32(&sh.chist[1][(uint32_t)highbit(ze_off_base(in.ll, in.off, i))], 1);
            ex.add32(&sh.chist[2][ml_code((uint32_t)in.ml[i] - 3u)], 1);
        }
    });

    // ---- P2: rank sort of the present literal symbols by (count, symbol)
    ex.phase([&](int tid) {
        for (int s = tid; s < 256; s += ZE_THREADS) {
            const uint32_t c = sh.hist[s];
            if (!c) continue;
            int rank = 0;
            for (int t = 0; t < 256; t++) {
                const uint32_t ct = sh.hist[t];
                rank += (ct != 0) && (ct < c || (ct == c && t < s));
            }
            sh.sorted[rank] = (uint16_t)s;
            ex.add32((uint32_t *)&sh.nsym, 1);
            ex.max32((uint32_t *)&sh.max_sym, (uint32_t)s);
        }
    });

    // ---- P3: thread 0 builds the Huffman code; the first threads of warps 1..3 build the three FSE
    // tables (different warps, so the four serial jobs run side by side)
    ex.phase([&](int tid) {
        if (tid == 0) {
            sh.lit_mode = 0;
            if (nlit > 0 && sh.nsym == 1) sh.lit_mode = 1;
            else if (nlit >= (uint32_t)ZE_MIN_HUF_LITS && sh.nsym >= 2) {
                const int log = huf_build_lengths(sh.u.b.hb, sh.hist, sh.sorted, sh.nsym, sh.nbits);
                // the tree description: FSE-compressed weights when that is smaller or the only form
                const int ms = sh.max_sym;
                for (int s = 0; s < ms; s++) sh.u.b.wts[s] = sh.nbits[s] ? (uint8_t)(log + 1 - sh.nbits[s]) : 0;
                int dl = 0;
                if (ms > 16) dl = huf_write_fse(sh.u.b.hufdesc, sh.u.b.wts, ms, sh.u.b.wct, sh.u.b.wcells);
                const int direct = ms <= 128 ? 1 + (ms + 1) / 2 : 0;
                if (dl == 0 || (direct && direct <= dl)) dl = direct ? huf_write_direct(sh.u.b.hufdesc, sh.nbits, ms, log) : 0;
                if (dl > 0) {
                    uint64_t bits = 0;
                    for (int s = 0; s <= ms; s++) bits += (uint64_t)sh.hist[s] * sh.nbits[s];
                    const uint64_t est = (bits >> 3) + 4 + 6 + (uint64_t)dl;
                    if (est + (nlit >> 6) < nlit) {                 // worth it (the exact size is checked in P5)
                        sh.lit_mode = 2; sh.huf_log = log; sh.hufdesc_len = dl;
                        huf_assign_codes(sh.hist, sh.nbits, ms, log);
                    }
                }
            }
        }
        if ((tid & 31) == 0 && tid >= 32 && nseq > 0) {
            const int k = (tid >> 5) - 1;                           // 0 LL, 1 OF, 2 ML
            const int nsymbols = k == 0 ? 36 : k == 1 ? 32 : 53;
            const int max_log = k == 1 ? 8 : 9;
            int used = 0, maxc = 0;
            for (int s = 0; s < nsymbols; s++) if (sh.chist[k][s]) { used++; maxc = s; }
            sh.max_code[k] = maxc;
            if (used == 1) {                                        // RLE: one symbol, no state bits
                sh.mode[k] = 1; sh.ct[k].log = 0; sh.log[k] = 0; sh.ncount_len[k] = 1;
                sh.u.b.ncount[k][0] = (uint8_t)maxc;
            } else if (nseq < (uint32_t)ZE_MIN_FSE_SEQ) {           // predefined distribution
                const short *dn = k == 0 ? T.ll_norm : k == 1 ? T.of_norm : T.ml_norm;
                const int dl = k == 1 ? 5 : 6, dmax = k == 0 ? 35 : k == 1 ? 28 : 52;
                fse_build_ctable(sh.ct[k], dn, dmax, dl, sh.u.b.cells[k]);
                sh.mode[k] = 0; sh.log[k] = dl; sh.ncount_len[k] = 0;
            } else {
                int log = max_log;
                { const int mb = highbit(nseq - 1) - 2; if (mb < log) log = mb; }
                { int a = highbit(nseq) + 1, b = highbit((uint32_t)maxc) + 2; const int minb = a < b ? a : b; if (log < minb) log = minb; }
                if (log < 5) log = 5;
                if (log > max_log) log = max_log;
                fse_normalize(sh.norm[k], log, sh.chist[k], nseq, maxc);
                sh.ncount_len[k] = fse_write_ncount(sh.u.b.ncount[k], sh.norm[k], maxc, log);
                fse_build_ctable(sh.ct[k], sh.norm[k], maxc, log, sh.u.b.cells[k]);
                sh.mode[k] = 2; sh.log[k] = log;
            }
        }
    });

    // ---- P4: Huffman bits of every chunk.  Warp w <-> stream w, lane l <-> l-th chunk in encoding
    // order (the stream is written from its last symbol to its first).
    const uint32_t seg = (nlit + 3) / 4;
    ex.phase([&](int tid) {
        uint32_t bits = 0;
        if (sh.lit_mode == 2) {
            const uint32_t w = (uint32_t)tid >> 5, l = (uint32_t)tid & 31;
            const uint32_t s_lo = w * seg < nlit ? w * seg : nlit, s_hi = (w + 1) * seg < nlit ? (w + 1) * seg : nlit;
            const uint32_t cnt = s_hi - s_lo, ch = (cnt + 31) / 32;
            const uint32_t hi = l * ch < cnt ? s_hi - l * ch : s_lo, lo = (l + 1) * ch < cnt ? s_hi - (l + 1) * ch : s_lo;
            for (uint32_t j = lo; j < hi; j++) bits += sh.hist[in.lits[j]] >> 16;
        }
        sh.scan[tid] = bits;
    });
    ex.excl_scan(sh.scan, sh.scan_tmp, &sh.scan_total);

    // ---- P5: layout of the literals section and of the sequence header (thread 0)
    ex.phase([&](int tid) {
        if (tid != 0) return;
        uint32_t pos = 0;
        if (sh.lit_mode == 2) {
            uint32_t total = 0;
            for (int w = 0; w < 4; w++) {
                const uint32_t b0 = sh.scan[w * 32], b1 = w < 3 ? sh.scan[(w + 1) * 32] : sh.scan_total;
                sh.stream_bits[w] = b1 - b0;
                total += (b1 - b0) / 8 + 1;
            }
            const uint32_t comp = (uint32_t)sh.hufdesc_len + 6 + total;
            if (comp >= nlit) sh.lit_mode = 0;                      // no gain: raw literals
            else {
                const uint32_t lh = 3 + (nlit >= 1024) + (nlit >= 16384);
                if (lh == 3) { const uint32_t v = 2u | (1u << 2) | (nlit << 4) | (comp << 14); put_bits(slot, 0, v, 24); }
                else if (lh == 4) { const uint32_t v = 2u | (2u << 2) | (nlit << 4) | (comp << 18); put_bits(slot, 0, v, 32); }
                else { const uint32_t v = 2u | (3u << 2) | (nlit << 4) | (comp << 22); put_bits(slot, 0, v, 32); put_byte(slot, 4, comp >> 10); }
                pos = lh;
                for (int i = 0; i < sh.hufdesc_len; i++) put_byte(slot, pos + i, sh.u.b.hufdesc[i]);
                pos += (uint32_t)sh.hufdesc_len;
                uint32_t so = pos + 6;
                for (int w = 0; w < 4; w++) {
                    const uint32_t sz = sh.stream_bits[w] / 8 + 1;
                    if (w < 3) { put_byte(slot, pos + 2 * w, sz & 255); put_byte(slot, pos + 2 * w + 1, sz >> 8); }
                    sh.stream_off[w] = so;
                    so += sz;
                }
                pos = so;
            }
        }
        if (sh.lit_mode != 2) {
            const uint32_t type = sh.lit_mode == 1 ? 1u : 0u;
            uint32_t lh;
            if (nlit < 32) { put_byte(slot, 0, type | (nlit << 3)); lh = 1; }
            else if (nlit < 4096) { put_bits(slot, 0, type | (1u << 2) | (nlit << 4), 16); lh = 2; }
            else { put_bits(slot, 0, type | (3u << 2) | (nlit << 4), 24); lh = 3; }
            sh.lit_hdr = lh;
            if (sh.lit_mode == 1) { put_byte(slot, lh, in.lits[0]); pos = lh + 1; }
            else pos = lh + nlit;
        }
        sh.lit_section = pos;
        // sequences section header (zstd_compress.c:2668-2682, zstd_decompress_block.c:656-750)
        sh.seq_hdr_off = pos;
        if (nseq < 128) put_byte(slot, pos++, nseq);
        else { put_byte(slot, pos++, (nseq >> 8) + 128); put_byte(slot, pos++, nseq & 255); }
        if (nseq > 0) {
            put_byte(slot, pos++, ((uint32_t)sh.mode[0] << 6) | ((uint32_t)sh.mode[1] << 4) | ((uint32_t)sh.mode[2] << 2));
            for (int k = 0; k < 3; k++)
                if (sh.mode[k]) { for (int i = 0; i < sh.ncount_len[k]; i++) put_byte(slot, pos + i, sh.u.b.ncount[k][i]); pos += (uint32_t)sh.ncount_len[k]; }
        }
        sh.bits_base = pos;
        if (pos >= in.rlen) sh.fail = 1;                            // already no smaller than a raw block
    });

    // ---- P6: the literals themselves
    ex.phase([&](int tid) {
        if (sh.fail) return;
        if (sh.lit_mode == 2) {
            const uint32_t w = (uint32_t)tid >> 5, l = (uint32_t)tid & 31;
            const uint32_t s_lo = w * seg < nlit ? w * seg : nlit, s_hi = (w + 1) * seg < nlit ? (w + 1) * seg : nlit;
            const uint32_t cnt = s_hi - s_lo, ch = (cnt + 31) / 32;
            const uint32_t hi = l * ch < cnt ? s_hi - l * ch : s_lo, lo = (l + 1) * ch < cnt ? s_hi - (l + 1) * ch : s_lo;
            const uint64_t g0 = (uint64_t)sh.stream_off[w] * 8;
            if (hi > lo) {
                BitRun br;
                br.start(slot, g0 + (sh.scan[tid] - sh.scan[w * 32]));
                for (uint32_t j = hi; j-- > lo;) { const uint32_t c = sh.hist[in.lits[j]]; br.add(c & 0xffffu, (int)(c >> 16)); }
                br.finish();
            }
            if (l == 0) put_bits(slot, g0 + sh.stream_bits[w], 1, 1);      // end mark (bitstream.h:213-222)
        } else if (sh.lit_mode == 0) {
            const uint32_t base = sh.lit_hdr;
            for (uint32_t i = (uint32_t)tid * 4; i < nlit; i += ZE_THREADS * 4) {
                uint32_t v = 0;
                const uint32_t m = nlit - i < 4 ? nlit - i : 4;
                for (uint32_t k = 0; k < m; k++) v |= (uint32_t)in.lits[i + k] << (8 * k);
                put_bits(slot, (uint64_t)(base + i) * 8, v, 32);
            }
        }
    });

    // ---- sequences: tiles of ZE_TILE from the LAST sequence backwards (zstd_compress_sequences.c:290-383).
    // Per tile: the three state chains (T2), per-thread bit totals (T3), scan, pack (T4).  The codes of the
    // next tile are made in the same phase as the pack of this one (double-buffered), and the running bit
    // position is brought up to date by thread 0 at the start of the next T2 / of the close phase.
    const uint32_t ntiles = (nseq + ZE_TILE - 1) / ZE_TILE;
    auto make_codes = [&](int tid, uint32_t tile) {
        const uint32_t hi = nseq - tile * ZE_TILE, cnt = hi < (uint32_t)ZE_TILE ? hi : (uint32_t)ZE_TILE;
        uint8_t(*code)[ZE_TILE] = sh.u.t.code[tile & 1];
        for (uint32_t p = (uint32_t)tid; p < cnt; p += ZE_THREADS) {                // position p of the tile <-> sequence hi - 1 - p
            const uint32_t i = hi - 1 - p;
            code[0][p] = (uint8_t)ll_code(in.ll[i]);
            code[1][p] = (uint8_t)highbit(ze_off_base(in.ll, in.off, i));
            code[2][p] = (uint8_t)ml_code((uint32_t)in.ml[i] - 3u);
        }
    };
    auto settle = [&]() {                                                           // thread 0 only
        if (!sh.pending) return;
        sh.pending = 0;
        if ((uint64_t)sh.bits_base * 8 + sh.bitpos + sh.scan_total + 64 > cap_bits) sh.fail = 1;
        else sh.bitpos += sh.scan_total;
    };
    ex.phase([&](int tid) {
        if (tid == 0) sh.pending = 0;
        if (ntiles > 0) make_codes(tid, 0);
    });
    for (uint32_t tile = 0; tile < ntiles; tile++) {
        const uint32_t hi = nseq - tile * ZE_TILE, cnt = hi < (uint32_t)ZE_TILE ? hi : (uint32_t)ZE_TILE;
        uint8_t(*code)[ZE_TILE] = sh.u.t.code[tile & 1];
        // T2: the three state chains, one thread each (lanes 0..2 of warp 0 run in lockstep)
        ex.phase([&](int tid) {
            if (tid == 0) settle();
            if (tid >= 3) return;
            const int k = tid;
            const FseCTable &ct = sh.ct[k];
            uint32_t p = 0;
            if (ct.log == 0) { for (; p < cnt; p++) sh.u.t.fse[k][p] = 0; return; }
            uint32_t st = sh.state[k];
            if (tile == 0) {                                         // FSE_initCState2 on the last sequence
                const int sym = code[k][0];
                const uint32_t nbout = (uint32_t)(ct.dnb[sym] + (1 << 15)) >> 16;
                st = (nbout << 16) - (uint32_t)ct.dnb[sym];
                st = ct.state[(st >> nbout) + ct.dfs[sym]];
                sh.u.t.fse[k][0] = 0;
                p = 1;
            }
            // The only loop-carried value is the state: the symbol two steps ahead and the per-symbol
            // constants one step ahead are fetched before they are needed, so a step costs one dependent
            // shared-memory load (the next state) instead of three.
            if (p < cnt) {
                const uint8_t *cd = code[k];
                int sym_next = p + 1 < cnt ? cd[p + 1] : 0;
                uint32_t dnb = (uint32_t)ct.dnb[cd[p]];
                int dfs = ct.dfs[cd[p]];
                for (; p < cnt; p++) {
                    const int sym_after = p + 2 < cnt ? cd[p + 2] : 0;
                    const uint32_t dnb_next = (uint32_t)ct.dnb[sym_next];
                    const int dfs_next = ct.dfs[sym_next];
                    const uint32_t nbout = (st + dnb) >> 16;
                    const uint32_t outv = (st & ((1u << nbout) - 1)) | (nbout << 10);
                    st = ct.state[(st >> nbout) + dfs];
                    sh.u.t.fse[k][p] = (uint16_t)outv;
                    dnb = dnb_next; dfs = dfs_next; sym_next = sym_after;
                }
            }
            sh.state[k] = st;
        });
        // T3: bits of every thread's run of ZE_PER_THREAD sequences
        ex.phase([&](int tid) {
            uint32_t bits = 0;
            const uint32_t p0 = (uint32_t)tid * ZE_PER_THREAD, p1 = p0 + ZE_PER_THREAD < cnt ? p0 + ZE_PER_THREAD : cnt;
            for (uint32_t p = p0; p < p1; p++) {
                const uint32_t lc = code[0][p], oc = code[1][p], mc = code[2][p];
                bits += (sh.u.t.fse[0][p] >> 10) + (sh.u.t.fse[1][p] >> 10) + (sh.u.t.fse[2][p] >> 10) + sh.ll_bits[lc] + sh.ml_bits[mc] + oc;
            }
            sh.scan[tid] = bits;
        });
        ex.excl_scan(sh.scan, sh.scan_tmp, &sh.scan_total);
        // T4: pack this tile; make the next tile's codes
        ex.phase([&](int tid) {
            if (tile + 1 < ntiles) make_codes(tid, tile + 1);
            if (tid == 0) sh.pending = 1;
            if (sh.fail) return;
            if ((uint64_t)sh.bits_base * 8 + sh.bitpos + sh.scan_total + 64 > cap_bits) return;      // settle() marks the failure
            const uint32_t p0 = (uint32_t)tid * ZE_PER_THREAD, p1 = p0 + ZE_PER_THREAD < cnt ? p0 + ZE_PER_THREAD : cnt;
            if (p0 >= p1) return;
            BitRun br;
            br.start(slot, (uint64_t)sh.bits_base * 8 + sh.bitpos + sh.scan[tid]);
            for (uint32_t p = p0; p < p1; p++) {
                const uint32_t i = hi - 1 - p;
                const uint32_t lc = code[0][p], oc = code[1][p], mc = code[2][p];
                const uint32_t fo = sh.u.t.fse[1][p], fm = sh.u.t.fse[2][p], fl = sh.u.t.fse[0][p];
                br.add(fo & 1023u, (int)(fo >> 10));
                br.add(fm & 1023u, (int)(fm >> 10));
                br.add(fl & 1023u, (int)(fl >> 10));
                br.add((uint32_t)in.ll[i] - sh.ll_base[lc], sh.ll_bits[lc]);
                br.add((uint32_t)in.ml[i] - sh.ml_base[mc], sh.ml_bits[mc]);
                br.add(ze_off_base(in.ll, in.off, i) - (1u << oc), (int)oc);
            }
            br.finish();
        });
    }

    // ---- close: final states (ML, OF, LL), end mark, sizes
    ex.phase([&](int tid) {
        if (tid != 0) return;
        settle();
        uint32_t total = sh.bits_base;
        if (!sh.fail && nseq > 0) {
            uint64_t g = (uint64_t)sh.bits_base * 8 + sh.bitpos;
            const int order[3] = {2, 1, 0};
            for (int q = 0; q < 3; q++) {
                const int k = order[q], lg = sh.ct[k].log;
                if (lg) { put_bits(slot, g, sh.state[k] & ((1u << lg) - 1), lg); g += (uint64_t)lg; }
            }
            put_bits(slot, g, 1, 1);
            g += 1;
            total = (uint32_t)((g + 7) >> 3);
        }
        out->bytes = total;
        out->raw = (sh.fail || total >= in.rlen) ? 1u : 0u;
    });
}

// ---- frame assembly (shared by the kernels and the emulation) ----------------------------------------

constexpr int ZE_FRAME_HDR = 9;       // magic, descriptor 0xA0 (single segment, 4-byte content size), size

FZ_HD inline void ze_write_frame_header(uint8_t *p, uint32_t usize)
{
    p[0] = 0x28; p[1] = 0xB5; p[2] = 0x2F; p[3] = 0xFD;          // zstd.h ZSTD_MAGICNUMBER, little endian
    p[4] = 0xA0;
    p[5] = (uint8_t)usize; p[6] = (uint8_t)(usize >> 8); p[7] = (uint8_t)(usize >> 16); p[8] = (uint8_t)(usize >> 24);
}

FZ_HD inline void ze_write_block_header(uint8_t *p, bool last, int type, uint32_t size)
{
    const uint32_t v = (last ? 1u : 0u) | ((uint32_t)type << 1) | (size << 3);
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16);
}

// ZSTD_compressBound (zstd.h:204): n + n/256 + margin for inputs below 128 KiB
FZ_HD inline size_t ze_compress_bound(size_t n)
{
    return n + (n >> 8) + (n < (128u << 10) ? ((128u << 10) - n) >> 11 : 0);
}

}  // namespace fmz
