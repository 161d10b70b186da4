// zstd_decode.h -- Zstandard frame decoder for 4mz blocks (host + device).
//
// Reference behaviour: ZSTD_decompress(out, usize, in, csize) as called per 4mz block at
// native/4mc.c:810 and native/jniZstdDecompressor.c (zstd 1.5.3 as vendored):
//   frames            native/zstd/decompress/zstd_decompress.c:443-551 (header), :901-987, :989-1100
//   blocks            native/zstd/decompress/zstd_decompress_block.c:57-71 (3-byte header), :2004
//   literals          :120-313 (raw / RLE / Huffman 1 or 4 streams / treeless, 3 size formats)
//   Huffman weights   native/zstd/common/entropy_common.c:244-312, decoding table huf_decompress.c:344-480
//   FSE descriptions  native/zstd/common/entropy_common.c:43-213, tables zstd_decompress_block.c:447-564
//   sequences         :656-750 (header, modes predefined / RLE / FSE / repeat), :1177-1295 (decode,
//                     repeat-offset rules), :956-1051 (execution), :1565-1650 (loop and end test)
// Written from the Zstandard format (RFC 8878) and the observable behaviour of the reference; this
// round's GPU mapping is one THREAD per frame (lz4mz kernel in zstd_decode.cuh): all frames of a
// batch decode concurrently, each one serially.  Intra-frame parallel entropy decoding is the
// next step.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define FZ_HD __host__ __device__
#else
#define FZ_HD
#endif

namespace fmz {

constexpr int ERR_CORRUPT = -20;        // corruption_detected and friends
constexpr int ERR_DSTSIZE = -70;        // dstSize_tooSmall
constexpr int ERR_SRCSIZE = -72;        // srcSize_wrong
constexpr int ERR_UNSUPPORTED = -14;    // dictionaries, window above 2^31
constexpr int BLOCK_MAX = 128 * 1024;   // zstd.h:132-133

struct HufEntry { uint8_t sym, nbits; };
struct SeqEntry { uint32_t base; uint16_t next; uint8_t nbits, extra; };

struct WtEntry { uint16_t next; uint8_t sym, nbits; };     // FSE table of the Huffman weights (log <= 6)

struct Work {                            // per frame, lives in global memory on the GPU
    static constexpr int HUF_MAX_LOG = 12;
    HufEntry huf[4096];
    SeqEntry ll[512], ml[512], of[256];
    int huf_log, ll_log, ml_log, of_log;
    int huf_ok, ll_ok, ml_ok, of_ok;     // a table exists (needed by treeless / repeat modes)
    int huf_x2;                          // the reference would hold a double-symbol table (see huf_stream)
    uint32_t rep[3];
    short norm[256];
    uint16_t symnext[256];
    uint8_t weights[256];
    uint32_t rank[16];
    WtEntry wt[64];
    uint8_t lit[BLOCK_MAX + 32];
};

FZ_HD inline int highbit(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);           // -1 for 0, like the loop below
#else
    int r = -1; while (v) { v >>= 1; r++; } return r;
#endif
}

// ---- bit readers ----------------------------------------------------------------------------

// Backward reader (bitstream.h:252-298,399-426): the last byte holds a stop bit; bits are taken
// from the top down.  Reading past the beginning yields zero bits and leaves `pos` negative,
// which is how the reference's "overflow" state is observed.
struct BackBits {
    const uint8_t *p;
    long long pos;                       // unread bits
    FZ_HD bool init(const uint8_t *src, long long n)
    {
        p = src;
        if (n < 1 || src[n - 1] == 0) { pos = 0; return false; }
        pos = 8 * (n - 1) + highbit(src[n - 1]);
        return true;
    }
    FZ_HD uint32_t read(int nb)         // nb <= 32
    {
        if (nb == 0) return 0;
        long long lo = pos - nb;
        pos = lo;
        uint64_t v = 0;
        int have = nb;
        if (lo < 0) { have = (int)(nb + lo); if (have <= 0) return 0; lo = 0; }
        const long long idx = lo >> 3;
        const int sh = (int)(lo & 7);
        const int nbytes = (sh + have + 7) >> 3;
        for (int i = 0; i < nbytes; i++) v |= (uint64_t)p[idx + i] << (8 * i);
        v = (v >> sh) & ((have >= 64) ? ~0ull : ((1ull << have) - 1));
        return (uint32_t)(v << (nb - have));
    }
};

// The sequence bitstream is read through a model of the reference's 64-bit container
// (bitstream.h:252-298 init, :334-372 look/read, :388-426 reload) instead of an abstract bit
// position: the reference accepts a sequence stream that ran past its beginning
// (zstd_decompress_block.c:1632 only rejects "not yet finished"), and what such reads return is
// whatever the container holds at the wrapped shift.  Reloads happen where the reference's do.
struct SeqBits {
    const uint8_t *p;
    long long at;                        // byte offset the container was loaded from
    uint32_t used;                       // bits consumed from the top of the container
    uint64_t c;
    long long len;                       // bytes of the stream
    FZ_HD void load()
    {
#if defined(__CUDA_ARCH__)
        if (at + 12 <= len) {            // three aligned words + funnel shifts (the extra bytes read stay inside the block)
            const uintptr_t a = (uintptr_t)(p + at);
            const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
            const uint32_t sh = (uint32_t)(a & 3) * 8;
            const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
            c = (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
            return;
        }
#endif
        uint64_t v = 0; for (int i = 0; i < 8; i++) v |= (uint64_t)p[at + i] << (8 * i); c = v;
    }
    FZ_HD bool init(const uint8_t *src, long long n)
    {
        p = src; len = n;
        if (n < 1 || src[n - 1] == 0) return false;
        used = 8 - highbit(src[n - 1]);
        if (n >= 8) { at = n - 8; load(); }
        else {
            at = 0; c = 0;
            for (int i = 0; i < n; i++) c |= (uint64_t)src[i] << (8 * i);
            used += (uint32_t)(8 - n) * 8;
        }
        return true;
    }
    FZ_HD uint32_t read(uint32_t nb)       // BIT_readBits: state initialisation and state updates
    {
        const uint64_t v = (c >> ((64u - used - nb) & 63u)) & ((1ull << nb) - 1);
        used += nb;
        return (uint32_t)v;
    }
    FZ_HD uint32_t read_fast(uint32_t nb)  // BIT_readBitsFast: extra bits, nb >= 1
    {
        const uint64_t v = (c << (used & 63u)) >> ((64u - nb) & 63u);
        used += nb;
        return (uint32_t)v;
    }
    // 0 unfinished, 1 end of buffer, 2 completed, 3 overflow (bitstream.h:50-54)
    FZ_HD int reload()
    {
        if (used > 64) return 3;
        if (at >= 8) { at -= used >> 3; used &= 7; load(); return 0; }
        if (at == 0) return used < 64 ? 1 : 2;
        uint32_t nbytes = used >> 3;
        int r = 0;
        if (at < (long long)nbytes) { nbytes = (uint32_t)at; r = 1; }
        at -= nbytes;
        used -= nbytes * 8;
        load();
        return r;
    }
};

// ---- FSE ------------------------------------------------------------------------------------

// FSE_readNCount (entropy_common.c:43-213).  Returns bytes consumed or <0.
FZ_HD inline int read_ncount(short *norm, int *max_sym, int *table_log, int max_log, const uint8_t *src, long long n)
{
    if (n < 1) return ERR_SRCSIZE;
    // forward LE bit cursor
    long long bitpos = 0;
    auto peek = [&](int nb) -> uint32_t {
        uint64_t v = 0;
        const long long idx = bitpos >> 3;
        for (int i = 0; i < 5; i++) if (idx + i < n) v |= (uint64_t)src[idx + i] << (8 * i);
        return (uint32_t)((v >> (bitpos & 7)) & ((1ull << nb) - 1));
    };
    int al = (int)peek(4) + 5;
    bitpos += 4;
    if (al > 15 || al > max_log) return ERR_CORRUPT;
    *table_log = al;
    int remaining = (1 << al) + 1, threshold = 1 << al, nbits = al + 1;
    int sym = 0;
    const int maxsv = *max_sym;
    bool prev0 = false;
    while (remaining > 1 && sym <= maxsv) {
        if (prev0) {
            int n0 = sym;
            while (peek(16) == 0xFFFF) { n0 += 24; bitpos += 16; if ((bitpos >> 3) > n) return ERR_CORRUPT; }
            while (peek(2) == 3) { n0 += 3; bitpos += 2; }
            n0 += (int)peek(2);
            bitpos += 2;
            if (n0 > maxsv + 1) return ERR_CORRUPT;
            while (sym < n0) norm[sym++] = 0;
            if (sym > maxsv) break;
        }
        {
            const int max = (2 * threshold - 1) - remaining;
            int count;
            const uint32_t lowv = peek(nbits - 1);
            if ((int)lowv < max) { count = (int)lowv; bitpos += nbits - 1; }
            else {
                count = (int)peek(nbits);
                if (count >= threshold) count -= max;
                bitpos += nbits;
            }
            count--;
            remaining -= count < 0 ? -count : count;
            norm[sym++] = (short)count;
            prev0 = count == 0;
            while (remaining < threshold) { nbits--; threshold >>= 1; }
        }
        if ((bitpos >> 3) > n) return ERR_CORRUPT;
    }
    if (remaining != 1) return ERR_CORRUPT;
    if (sym > maxsv + 1) return ERR_CORRUPT;
    *max_sym = sym - 1;
    const long long used = (bitpos + 7) >> 3;
    if (used > n) return ERR_SRCSIZE;
    return (int)used;
}

// Spread of symbols over the table (zstd_decompress_block.c:447-564, fse_decompress.c:68-175): both
// builders lay symbols out identically; `cells[u]` receives the symbol of table position u.
FZ_HD inline bool fse_spread(uint8_t *cells, uint16_t *symnext, const short *norm, int max_sym, int log)
{
    const int size = 1 << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    int high = size - 1;
    for (int s = 0; s <= max_sym; s++) {
        if (norm[s] == -1) { cells[high--] = (uint8_t)s; symnext[s] = 1; }
        else symnext[s] = (uint16_t)norm[s];
    }
    int pos = 0;
    for (int s = 0; s <= max_sym; s++) {
        for (int i = 0; i < norm[s]; i++) {
            cells[pos] = (uint8_t)s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    }
    return pos == 0;
}

FZ_HD inline bool build_seq_table(SeqEntry *t, uint16_t *symnext, const short *norm, int max_sym, int log,
                                  const uint32_t *base, const uint8_t *extra, uint8_t *cells)
{
    if (!fse_spread(cells, symnext, norm, max_sym, log)) return false;
    const int size = 1 << log;
    for (int u = 0; u < size; u++) {
        const int s = cells[u];
        const uint32_t nx = symnext[s]++;
        const int nb = log - highbit(nx);
        t[u].nbits = (uint8_t)nb;
        t[u].next = (uint16_t)((nx << nb) - size);
        t[u].extra = extra[s];
        t[u].base = base[s];
    }
    return true;
}

// ---- Huffman ----------------------------------------------------------------------------------

// HUF_readStats + HUF_readDTableX1 (entropy_common.c:244-312, huf_decompress.c:344-480).
// Returns bytes consumed or <0.
template <class W>
FZ_HD inline int read_huffman(W &w, const uint8_t *src, long long n)
{
    if (n < 1) return ERR_SRCSIZE;
    int isz = src[0], osz;
    uint8_t *wt = w.weights;
    if (isz >= 128) {
        osz = isz - 127;
        isz = (osz + 1) / 2;
        if (isz + 1 > n) return ERR_SRCSIZE;
        if (osz >= 256) return ERR_CORRUPT;
        for (int k = 0; k < osz; k += 2) { wt[k] = src[1 + k / 2] >> 4; if (k + 1 < 256) wt[k + 1] = src[1 + k / 2] & 15; }
    } else {
        if (isz + 1 > n) return ERR_SRCSIZE;
        // FSE-compressed weights (fse_decompress.c:232-300): table log <= 6, two interleaved states
        int max_sym = 255, log = 0;
        const int hdr = read_ncount(w.norm, &max_sym, &log, 6, src + 1, isz);
        if (hdr < 0) return hdr;
        uint8_t cells[64];
        if (!fse_spread(cells, w.symnext, w.norm, max_sym, log)) return ERR_CORRUPT;
        const int size = 1 << log;
        for (int u = 0; u < size; u++) {
            const int s = cells[u];
            const uint32_t nx = w.symnext[s]++;
            const int nb = log - highbit(nx);
            w.wt[u].sym = (uint8_t)s; w.wt[u].nbits = (uint8_t)nb; w.wt[u].next = (uint16_t)((nx << nb) - size);
        }
        BackBits b;
        if (!b.init(src + 1 + hdr, isz - hdr)) return ERR_CORRUPT;
        uint32_t s1 = b.read(log), s2 = b.read(log);
        osz = 0;
        for (;;) {
            if (osz > 253) return ERR_CORRUPT;
            wt[osz++] = w.wt[s1].sym;
            s1 = w.wt[s1].next + b.read(w.wt[s1].nbits);
            if (b.pos < 0) { wt[osz++] = w.wt[s2].sym; break; }
            if (osz > 253) return ERR_CORRUPT;
            wt[osz++] = w.wt[s2].sym;
            s2 = w.wt[s2].next + b.read(w.wt[s2].nbits);
            if (b.pos < 0) { wt[osz++] = w.wt[s1].sym; break; }
        }
    }
    for (int k = 0; k < 16; k++) w.rank[k] = 0;
    uint32_t total = 0;
    for (int k = 0; k < osz; k++) {
        if (wt[k] > 12) return ERR_CORRUPT;
        w.rank[wt[k]]++;
        total += (1u << wt[k]) >> 1;
    }
    if (total == 0) return ERR_CORRUPT;
    const int log = highbit(total) + 1;
    if (log > 12) return ERR_CORRUPT;
    if (log > W::HUF_MAX_LOG) return ERR_UNSUPPORTED;       // table does not fit this work area (the warp kernel's: retried serially)
    {
        const uint32_t rest = (1u << log) - total;
        if ((1u << highbit(rest)) != rest) return ERR_CORRUPT;
        wt[osz] = (uint8_t)(highbit(rest) + 1);
        w.rank[wt[osz]]++;
    }
    if (w.rank[1] < 2 || (w.rank[1] & 1)) return ERR_CORRUPT;
    const int nsym = osz + 1;
    // table: weight-1 symbols first (one cell each), then weight 2 (two cells), ... in symbol order
    uint32_t start[16];
    {
        uint32_t next = 0;
        for (int r = 1; r <= log; r++) { start[r] = next; next += w.rank[r] << (r - 1); }
    }
    for (int s = 0; s < nsym; s++) {
        const int r = wt[s];
        if (!r) continue;
        const uint32_t len = (1u << r) >> 1;
        const uint8_t nb = (uint8_t)(log + 1 - r);
        for (uint32_t u = 0; u < len; u++) { w.huf[start[r] + u].sym = (uint8_t)s; w.huf[start[r] + u].nbits = nb; }
        start[r] += len;
    }
    w.huf_log = log;
    w.huf_ok = 1;
    return isz + 1;
}

// HUF_selectDecoder (huf_decompress.c:1566-1617): the reference picks its single- or double-symbol
// decoder from a timing model.  Both decode valid streams identically; they differ in which
// corrupted streams they accept, so the choice is reproduced to keep accept/reject identical.
FZ_HD inline int huf_select_x2(long long dst_size, long long src_size)
{
    const uint16_t t0[16] = {0, 0, 150, 170, 177, 197, 221, 256, 359, 582, 688, 825, 976, 1180, 1377, 1412};
    const uint8_t d0[16] = {0, 0, 216, 205, 199, 194, 192, 189, 188, 187, 187, 186, 185, 186, 185, 185};
    const uint16_t t1[16] = {1, 1, 381, 514, 539, 644, 735, 881, 1167, 1570, 1712, 1965, 2131, 2070, 1731, 1695};
    const uint8_t d1[16] = {1, 1, 119, 112, 110, 107, 107, 106, 109, 114, 122, 136, 150, 175, 202, 202};
    const uint32_t q = src_size >= dst_size ? 15u : (uint32_t)(src_size * 16 / dst_size);
    const uint32_t d256 = (uint32_t)(dst_size >> 8);
    const uint32_t time0 = t0[q] + d0[q] * d256;
    uint32_t time1 = t1[q] + d1[q] * d256;
    time1 += time1 >> 5;
    return time1 < time0;
}

// One Huffman stream of `count` symbols.  The table is always the single-symbol one; when the
// reference would run its double-symbol decoder (w.huf_x2) its pairing is replayed on top of it:
// a lookup of dl = max(11, log) bits yields two symbols when both codes fit in dl bits
// (HUF_fillDTableX2, huf_decompress.c:983-1045), and the odd last symbol follows
// HUF_decodeLastSymbolX2 (:1150-1165): a pair entry there consumes at most the bits that are left,
// and on an exhausted stream it is looked up in the still-loaded first 8 bytes and consumes nothing.
FZ_HD inline int huf_stream(const Work &w, uint8_t *dst, int count, const uint8_t *src, long long n)
{
    BackBits b;
    if (!b.init(src, n)) return ERR_CORRUPT;
    const int log = w.huf_log;
    if (!w.huf_x2) {
        for (int i = 0; i < count; i++) {
            // peek `log` bits without consuming, then consume the code length
            BackBits t = b;
            const uint32_t idx = t.read(log);
            const HufEntry e = w.huf[idx];
            dst[i] = e.sym;
            b.pos -= e.nbits;
        }
        return b.pos == 0 ? 0 : ERR_CORRUPT;                  // BIT_endOfDStream (huf_decompress.c:643)
    }
    const int dl = log <= 11 ? 11 : 12;                       // :1083
    const uint32_t dmask = (1u << dl) - 1;
    int i = 0;
    while (count - i >= 2) {
        BackBits t = b;
        const uint32_t win = t.read(dl);
        const HufEntry e1 = w.huf[win >> (dl - log)];
        const HufEntry e2 = w.huf[((win << e1.nbits) & dmask) >> (dl - log)];
        dst[i++] = e1.sym;
        b.pos -= e1.nbits;
        if (e1.nbits + e2.nbits <= dl) { dst[i++] = e2.sym; b.pos -= e2.nbits; }
    }
    if (i < count) {
        if (b.pos < 0) return ERR_CORRUPT;
        uint32_t win;
        if (b.pos > 0) { BackBits t = b; win = t.read(dl); }
        else {
            uint64_t c = 0;
            for (int k = 0; k < 8 && k < n; k++) c |= (uint64_t)src[k] << (8 * k);
            win = (uint32_t)(c >> (64 - dl));
        }
        const HufEntry e1 = w.huf[win >> (dl - log)];
        const HufEntry e2 = w.huf[((win << e1.nbits) & dmask) >> (dl - log)];
        dst[i] = e1.sym;
        if (e1.nbits + e2.nbits > dl) b.pos -= e1.nbits;
        else if (b.pos > 0) { b.pos -= e1.nbits + e2.nbits; if (b.pos < 0) b.pos = 0; }
    }
    return b.pos == 0 ? 0 : ERR_CORRUPT;                      // :1244,1359
}

// ---- one compressed block -----------------------------------------------------------------------

struct Tables {                                             // constant tables, passed in (no statics in device code)
    uint32_t ll_base[36], ml_base[53], of_base[32];
    uint8_t ll_bits[36], ml_bits[53], of_bits[32];
    short ll_norm[36], ml_norm[53], of_norm[29];
};

FZ_HD inline void make_tables(Tables &T)
{
    const uint32_t llb[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40,
                              48, 64, 0x80, 0x100, 0x200, 0x400, 0x800, 0x1000, 0x2000, 0x4000, 0x8000, 0x10000};
    const uint8_t llx[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    const uint32_t mlb[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
                              35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 0x83, 0x103, 0x203, 0x403, 0x803, 0x1003, 0x2003, 0x4003, 0x8003, 0x10003};
    const uint8_t mlx[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                             1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    const short lln[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
    const short mln[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                           1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
    const short ofn[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
    for (int i = 0; i < 36; i++) { T.ll_base[i] = llb[i]; T.ll_bits[i] = llx[i]; T.ll_norm[i] = lln[i]; }
    for (int i = 0; i < 53; i++) { T.ml_base[i] = mlb[i]; T.ml_bits[i] = mlx[i]; T.ml_norm[i] = mln[i]; }
    for (int i = 0; i < 29; i++) T.of_norm[i] = ofn[i];
    for (int i = 0; i < 32; i++) { T.of_base[i] = (i == 0) ? 0u : ((1u << i) - 0u); T.of_bits[i] = (uint8_t)i; }
    // OF_base (zstd_decompress_block.c:411-416): 0, 1, 1, 5, 0xD, 0x1D, ... = (1 << code) - 3 for code >= 2
    T.of_base[0] = 0; T.of_base[1] = 1;
    for (int i = 2; i < 32; i++) T.of_base[i] = (1u << i) - 3u;
}

// one of the three sequence tables (zstd_decompress_block.c:608-653)
template <class W>
FZ_HD inline int build_mode(int mode, SeqEntry *t, int *log, int *ok, W &w, const Tables &T, int which,
                            const uint8_t *src, long long n)
{
    const int max_sym = which == 0 ? 35 : which == 1 ? 31 : 52;         // LL, OF, ML
    const int max_log = which == 0 ? 9 : which == 1 ? 8 : 9;
    const uint32_t *base = which == 0 ? T.ll_base : which == 1 ? T.of_base : T.ml_base;
    const uint8_t *extra = which == 0 ? T.ll_bits : which == 1 ? T.of_bits : T.ml_bits;
    uint8_t cells[512];
    if (mode == 0) {                                                     // predefined
        const short *norm = which == 0 ? T.ll_norm : which == 1 ? T.of_norm : T.ml_norm;
        const int dl = which == 1 ? 5 : 6, dmax = which == 0 ? 35 : which == 1 ? 28 : 52;
        if (!build_seq_table(t, w.symnext, norm, dmax, dl, base, extra, cells)) return ERR_CORRUPT;
        *log = dl; *ok = 1;
        return 0;
    }
    if (mode == 1) {                                                     // RLE
        if (n < 1) return ERR_SRCSIZE;
        const int s = src[0];
        if (s > max_sym) return ERR_CORRUPT;
        t[0].base = base[s]; t[0].extra = extra[s]; t[0].nbits = 0; t[0].next = 0;
        *log = 0; *ok = 1;
        return 1;
    }
    if (mode == 2) {                                                     // FSE description
        int ms = max_sym, l = 0;
        const int used = read_ncount(w.norm, &ms, &l, max_log, src, n);
        if (used < 0) return used;
        if (!build_seq_table(t, w.symnext, w.norm, ms, l, base, extra, cells)) return ERR_CORRUPT;
        *log = l; *ok = 1;
        return used;
    }
    return *ok ? 0 : ERR_CORRUPT;                                        // repeat
}

// Decodes one compressed block body (literals + sequences) into dst[op..]; `base` is the start of
// the frame's output (offsets may reach back to it).  Returns the new op or <0.
FZ_HD inline long long decode_block(Work &w, const Tables &T, uint8_t *dst, long long op, long long cap,
                                    const uint8_t *src, long long n)
{
    // ---- literals section (zstd_decompress_block.c:120-313)
    if (n < 1) return ERR_CORRUPT;
    const int ltype = src[0] & 3, sf = (src[0] >> 2) & 3;
    long long hdr, regen, comp = 0;
    const uint8_t *lit = nullptr;
    bool own_lit = true;                                   // literals were regenerated (not used in place)
    if (ltype < 2) {                                       // raw / RLE
        if (sf == 0 || sf == 2) { hdr = 1; regen = src[0] >> 3; }
        else if (sf == 1) { if (n < 2) return ERR_CORRUPT; hdr = 2; regen = (src[0] >> 4) + ((long long)src[1] << 4); }
        else { if (n < 3) return ERR_CORRUPT; hdr = 3; regen = (src[0] >> 4) + ((long long)src[1] << 4) + ((long long)src[2] << 12); }
        if (regen > BLOCK_MAX) return ERR_CORRUPT;
        if (ltype == 0) {
            if (hdr + regen > n) return ERR_CORRUPT;
            lit = src + hdr;                               // used in place
            own_lit = hdr + regen + 32 > n;                // :249 copied when a wild copy could over-read
            hdr += regen;
        } else {
            if (hdr + 1 > n) return ERR_CORRUPT;
            for (long long i = 0; i < regen; i++) w.lit[i] = src[hdr];
            lit = w.lit;
            hdr += 1;
        }
    } else {                                               // Huffman / treeless
        if (n < 5 && sf == 3) return ERR_CORRUPT;
        if (n < 3) return ERR_CORRUPT;
        int streams = 4;
        if (sf <= 1) {
            const uint32_t v = src[0] | (src[1] << 8) | (src[2] << 16);
            hdr = 3; regen = (v >> 4) & 0x3FF; comp = (v >> 14) & 0x3FF;
            streams = sf == 0 ? 1 : 4;
        } else if (sf == 2) {
            if (n < 4) return ERR_CORRUPT;
            const uint32_t v = src[0] | (src[1] << 8) | (src[2] << 16) | ((uint32_t)src[3] << 24);
            hdr = 4; regen = (v >> 4) & 0x3FFF; comp = v >> 18;
        } else {
            if (n < 5) return ERR_CORRUPT;
            const uint64_t v = (uint64_t)src[0] | ((uint64_t)src[1] << 8) | ((uint64_t)src[2] << 16) | ((uint64_t)src[3] << 24) | ((uint64_t)src[4] << 32);
            hdr = 5; regen = (long long)((v >> 4) & 0x3FFFF); comp = (long long)(v >> 22);
        }
        if (regen > BLOCK_MAX) return ERR_CORRUPT;
        if (hdr + comp > n) return ERR_CORRUPT;
        const uint8_t *cs = src + hdr;
        long long cn = comp;
        if (ltype == 2) {
            const int used = read_huffman(w, cs, cn);
            if (used < 0) return used;
            cs += used; cn -= used;
            w.huf_x2 = streams == 4 && huf_select_x2(regen, comp);   // zstd_decompress_block.c:189-206
        } else if (!w.huf_ok) return ERR_CORRUPT;         // treeless without a previous table
        if (streams == 1) {
            const int e = huf_stream(w, w.lit, (int)regen, cs, cn);
            if (e < 0) return e;
        } else {
            if (cn < 10) return ERR_CORRUPT;               // huf_decompress.c:573 (jump table + 1 byte per stream)
            const long long s1 = cs[0] | (cs[1] << 8), s2 = cs[2] | (cs[3] << 8), s3 = cs[4] | (cs[5] << 8);
            const long long s4 = cn - 6 - s1 - s2 - s3;
            if (s4 < 1 || s1 < 1 || s2 < 1 || s3 < 1) return ERR_CORRUPT;
            const int seg = (int)((regen + 3) / 4);
            if (3LL * seg > regen) return ERR_CORRUPT;
            const uint8_t *q = cs + 6;
            int e;
            if ((e = huf_stream(w, w.lit, seg, q, s1)) < 0) return e;
            if ((e = huf_stream(w, w.lit + seg, seg, q + s1, s2)) < 0) return e;
            if ((e = huf_stream(w, w.lit + 2 * seg, seg, q + s1 + s2, s3)) < 0) return e;
            if ((e = huf_stream(w, w.lit + 3 * seg, (int)(regen - 3 * seg), q + s1 + s2 + s3, s4)) < 0) return e;
        }
        lit = w.lit;
        hdr += comp;
    }
    // Where the reference keeps regenerated literals decides how far this block may write: with
    // room to spare they sit in dst 128 KiB + 32 past the block start and bound the output
    // (ZSTD_allocateLiteralsBuffer :77-83, oend :1574).  Only malformed blocks can tell.
    long long out_cap = cap;
    if (own_lit && cap - op > BLOCK_MAX + 32 + regen + 32) out_cap = op + BLOCK_MAX + 32;
    // ---- sequences section header (:656-750)
    const uint8_t *sp = src + hdr;
    long long sn = n - hdr;
    if (sn < 1) return ERR_SRCSIZE;
    long long nseq = sp[0];
    long long shdr = 1;
    if (nseq == 0 && sn != 1) return ERR_SRCSIZE;          // :667-671
    if (nseq >= 128) {
        if (nseq == 255) { if (sn < 3) return ERR_SRCSIZE; nseq = sp[1] + (sp[2] << 8) + 0x7F00; shdr = 3; }
        else { if (sn < 2) return ERR_SRCSIZE; nseq = ((nseq - 128) << 8) + sp[1]; shdr = 2; }
    }
    long long lit_pos = 0;
    if (nseq > 0) {
        if (shdr + 1 > sn) return ERR_SRCSIZE;
        const int modes = sp[shdr];
        // the two reserved bits are not checked by zstd 1.5.3 (:688-692)
        shdr += 1;
        int used;
        if ((used = build_mode((modes >> 6) & 3, w.ll, &w.ll_log, &w.ll_ok, w, T, 0, sp + shdr, sn - shdr)) < 0) return used;
        shdr += used;
        if ((used = build_mode((modes >> 4) & 3, w.of, &w.of_log, &w.of_ok, w, T, 1, sp + shdr, sn - shdr)) < 0) return used;
        shdr += used;
        if ((used = build_mode((modes >> 2) & 3, w.ml, &w.ml_log, &w.ml_ok, w, T, 2, sp + shdr, sn - shdr)) < 0) return used;
        shdr += used;
        // ---- sequence decoding and execution (:1565-1650, :1177-1295, :956-1051)
        SeqBits b;
        if (!b.init(sp + shdr, sn - shdr)) return ERR_CORRUPT;
        uint32_t sl = b.read(w.ll_log); b.reload();                          // ZSTD_initFseState, :1138-1148
        uint32_t so = b.read(w.of_log); b.reload();
        uint32_t sm = b.read(w.ml_log); b.reload();
        uint32_t r0 = w.rep[0], r1 = w.rep[1], r2 = w.rep[2];
        for (long long k = 0; k < nseq; k++) {
            const SeqEntry el = w.ll[sl], eo = w.of[so], em = w.ml[sm];
            uint32_t offset;
            if (eo.extra > 1) {
                offset = eo.base + b.read_fast(eo.extra);
                r2 = r1; r1 = r0; r0 = offset;
            } else {
                const uint32_t ll0 = el.base == 0;
                if (eo.extra == 0) {
                    offset = ll0 ? r1 : r0;
                    r1 = ll0 ? r0 : r1;
                    r0 = offset;
                } else {
                    const uint32_t code = eo.base + ll0 + b.read_fast(1);
                    uint32_t t = code == 3 ? r0 - 1 : (code == 1 ? r1 : code == 2 ? r2 : r0);
                    t += !t;
                    if (code != 1) r2 = r1;
                    r1 = r0;
                    r0 = offset = t;
                }
            }
            long long mlen = em.base, llen = el.base;
            if (em.extra) mlen += b.read_fast(em.extra);
            if (el.extra + em.extra + eo.extra >= 31) b.reload();           // :1270, 57 - (9 + 9 + 8)
            if (el.extra) llen += b.read_fast(el.extra);
            sl = el.next + b.read(el.nbits);
            sm = em.next + b.read(em.nbits);
            so = eo.next + b.read(eo.nbits);
            // execution (ZSTD_execSequence / ZSTD_execSequenceEnd, :956-1051, :862-905)
            if (llen + mlen > out_cap - op) return ERR_DSTSIZE;
            if (llen > regen - lit_pos) return ERR_CORRUPT;                 // literals overrun
            for (long long i = 0; i < llen; i++) dst[op + i] = lit[lit_pos + i];
            op += llen; lit_pos += llen;
            if ((long long)offset > op) return ERR_CORRUPT;                 // offset beyond the frame start
            for (long long i = 0; i < mlen; i++) dst[op + i] = dst[op - offset + i];
            op += mlen;
            if (k + 1 < nseq) b.reload();                                   // :1625-1628
        }
        if (b.reload() < 2) return ERR_CORRUPT;                             // :1632 stream not consumed
        w.rep[0] = r0; w.rep[1] = r1; w.rep[2] = r2;
    }
    // last literals (:1638-1646)
    {
        const long long rest = regen - lit_pos;
        if (rest > out_cap - op) return ERR_DSTSIZE;
        for (long long i = 0; i < rest; i++) dst[op + i] = lit[lit_pos + i];
        op += rest;
    }
    return op;
}

// ZSTD_decompress: one or more frames (zstd_decompress.c:989-1100).  Returns decoded size or <0.
FZ_HD inline long long decompress(uint8_t *dst, long long cap, const uint8_t *src, long long n, Work &w, const Tables &T)
{
    long long ip = 0, op = 0;
    bool any = false;
    while (ip < n) {
        if (n - ip < 4) return ERR_SRCSIZE;
        const uint32_t magic = src[ip] | (src[ip + 1] << 8) | (src[ip + 2] << 16) | ((uint32_t)src[ip + 3] << 24);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {       // skippable frame
            if (n - ip < 8) return ERR_SRCSIZE;
            const uint32_t sz = src[ip + 4] | (src[ip + 5] << 8) | (src[ip + 6] << 16) | ((uint32_t)src[ip + 7] << 24);
            if ((long long)sz + 8 > n - ip) return ERR_SRCSIZE;
            ip += 8 + sz;
            continue;
        }
        if (magic != 0xFD2FB528u) return any ? ERR_SRCSIZE : ERR_CORRUPT;   // :1023-1030 prefix_unknown / trailing garbage
        // frame header (zstd_decompress.c:443-551)
        if (n - ip < 5) return ERR_SRCSIZE;
        const int fhd = src[ip + 4];
        const int did = fhd & 3, cksum = (fhd >> 2) & 1, single = (fhd >> 5) & 1, fcsid = fhd >> 6;
        if (fhd & 8) return ERR_UNSUPPORTED;              // reserved bit
        long long h = ip + 5;
        uint64_t window = 0;
        if (!single) {
            if (h >= n) return ERR_SRCSIZE;
            const int wl = (src[h] >> 3) + 10;
            if (wl > 31) return ERR_UNSUPPORTED;          // zstd.h:1150 ZSTD_WINDOWLOG_MAX
            window = (1ull << wl) + ((1ull << wl) >> 3) * (src[h] & 7);
            h++;
        }
        const int dsz = did == 0 ? 0 : did == 1 ? 1 : did == 2 ? 2 : 4;
        if (h + dsz > n) return ERR_SRCSIZE;
        if (dsz) { uint32_t dict = 0; for (int i = 0; i < dsz; i++) dict |= (uint32_t)src[h + i] << (8 * i); if (dict) return ERR_UNSUPPORTED; }
        h += dsz;
        const int fsz = fcsid == 0 ? (single ? 1 : 0) : fcsid == 1 ? 2 : fcsid == 2 ? 4 : 8;
        if (h + fsz > n) return ERR_SRCSIZE;
        uint64_t fcs = 0;
        for (int i = 0; i < fsz; i++) fcs |= (uint64_t)src[h + i] << (8 * i);
        if (fsz == 2) fcs += 256;
        h += fsz;
        if (single) window = fcs;
        const long long block_max = (long long)(window < (uint64_t)BLOCK_MAX ? window : (uint64_t)BLOCK_MAX);
        // per-frame state (zstd_decompress.c: ZSTD_decompressBegin -> entropy reset, rep = {1,4,8})
        w.huf_ok = w.ll_ok = w.ml_ok = w.of_ok = w.huf_x2 = 0;
        w.rep[0] = 1; w.rep[1] = 4; w.rep[2] = 8;
        const long long frame_start = op;
        ip = h;
        for (;;) {
            if (n - ip < 3) return ERR_SRCSIZE;
            const uint32_t bh = src[ip] | (src[ip + 1] << 8) | (src[ip + 2] << 16);
            ip += 3;
            const int last = bh & 1, type = (bh >> 1) & 3;
            const long long bsz = bh >> 3;
            if (type == 3) return ERR_CORRUPT;
            if (type == 1) {                               // RLE
                if (n - ip < 1) return ERR_SRCSIZE;
                if (bsz > block_max) return ERR_CORRUPT;
                if (bsz > cap - op) return ERR_DSTSIZE;
                for (long long i = 0; i < bsz; i++) dst[op + i] = src[ip];
                op += bsz; ip += 1;
            } else {
                if (bsz > n - ip) return ERR_SRCSIZE;
                if (bsz > block_max) return ERR_CORRUPT;   // zstd_decompress.c:939
                if (type == 2 && bsz >= BLOCK_MAX) return ERR_SRCSIZE;      // zstd_decompress_block.c:2019
                if (type == 0) {
                    if (bsz > cap - op) return ERR_DSTSIZE;
                    for (long long i = 0; i < bsz; i++) dst[op + i] = src[ip + i];
                    op += bsz;
                } else {
                    // offsets are relative to this frame's output: decode with the frame as base
                    const long long r = decode_block(w, T, dst + frame_start, op - frame_start, cap - frame_start, src + ip, bsz);
                    if (r < 0) return r;
                    op = frame_start + r;
                }
                ip += bsz;
            }
            if (last) break;
        }
        if (fsz && (uint64_t)(op - frame_start) != fcs) return ERR_CORRUPT;   // zstd_decompress.c:967-970
        if (cksum) { if (n - ip < 4) return ERR_CORRUPT; ip += 4; }           // XXH64 of the content: not verified (4mz frames carry none)
        any = true;
    }
    return op;
}

}  // namespace fmz
