/*
 * zstd_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product): a plain-C Zstandard
 * frame decoder restated from the format (RFC 8878) and from the reference's decoder sources, used to
 * check the 4mz path independently of the product's own decoder:
 *   frames / blocks      native/zstd/decompress/zstd_decompress.c:443-551 (frame header), :901-987 (block loop)
 *   literals section     native/zstd/decompress/zstd_decompress_block.c:120-313
 *   Huffman tables       native/zstd/common/entropy_common.c:244-312, native/zstd/decompress/huf_decompress.c:344-480
 *   FSE tables           native/zstd/common/entropy_common.c:43-213, zstd_decompress_block.c:447-564
 *   sequences            zstd_decompress_block.c:656-750 (header, modes), :1177-1295 (decoding, repeat offsets),
 *                        :956-1051 (execution)
 * It is STRICT: anything the format does not allow is an error (-1).  The reference tolerates a few
 * malformed inputs that a strict reader rejects (and the product reproduces those quirks, pinned to
 * the reference's verdicts in tests/golden/zstd_decode.json); this oracle is the yardstick for VALID
 * frames: same bytes, same size.  Pinned by tests/test_oracle_zstd.py against the .4mz files the
 * reference CLI wrote and against the reference's own decode runs.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "fourmc_oracle.h"

#define ZBLOCK_MAX (128 * 1024)

typedef struct { const uint8_t *p; long long bits; } back_t;          /* backward bit reader: `bits` unread */
typedef struct { uint16_t next; uint8_t sym, nbits; } fse_cell;
typedef struct { fse_cell cell[512]; int log; int rle; } fse_table;
typedef struct { uint8_t sym[4096], len[4096]; int log; int ok; } huf_table;
typedef struct {
    huf_table huf;
    fse_table ll, of, ml;
    int ll_ok, of_ok, ml_ok;
    uint32_t rep[3];
    uint8_t lit[ZBLOCK_MAX + 8];
} zctx;

static int hibit(uint32_t v) { int r = -1; while (v) { v >>= 1; r++; } return r; }

static int back_init(back_t *b, const uint8_t *p, long long n)
{
    if (n < 1 || p[n - 1] == 0) return -1;
    b->p = p; b->bits = 8 * (n - 1) + hibit(p[n - 1]);
    return 0;
}
/* next nb bits (nb <= 32), most significant first; bits before the start of the stream read as zero */
static uint32_t back_read(back_t *b, int nb)
{
    uint32_t v = 0;
    int i;
    for (i = 0; i < nb; i++) {
        long long pos = b->bits - 1 - i;
        v <<= 1;
        if (pos >= 0) v |= (b->p[pos >> 3] >> (pos & 7)) & 1u;
    }
    b->bits -= nb;
    return v;
}

/* FSE table description -> normalized counts.  Returns bytes used or -1. */
static int read_norm(const uint8_t *src, long long n, short *norm, int *max_sym, int *log_out, int max_log)
{
    long long bit = 0;
    int al, remaining, sym = 0, i;
#define FWD(nb, out) do { uint32_t v_ = 0; for (i = 0; i < (nb); i++) { long long q = bit + i; if ((q >> 3) >= n) return -1; v_ |= (uint32_t)((src[q >> 3] >> (q & 7)) & 1u) << i; } bit += (nb); (out) = v_; } while (0)
    uint32_t v;
    FWD(4, v);
    al = (int)v + 5;
    if (al > max_log) return -1;
    remaining = 1 << al;
    while (remaining > 0 && sym <= *max_sym) {
        int nb = hibit((uint32_t)(remaining + 1)) + 1;            /* bits to hold values 0 .. remaining + 1 */
        int thresh = (1 << nb) - 1 - (remaining + 1);
        uint32_t lowv, val;
        long long save = bit;
        FWD(nb - 1, lowv);
        if ((int)lowv < thresh) val = lowv;
        else {
            bit = save;
            FWD(nb, val);
            if ((int)val >= (1 << (nb - 1))) val -= (uint32_t)thresh;
        }
        {
            int count = (int)val - 1;                             /* -1 = "less than one" */
            norm[sym++] = (short)count;
            remaining -= count < 0 ? 1 : count;
            if (count == 0) {
                for (;;) {
                    uint32_t rep;
                    FWD(2, rep);
                    for (i = 0; i < (int)rep; i++) { if (sym > *max_sym) return -1; norm[sym++] = 0; }
                    if (rep != 3) break;
                }
            }
        }
    }
#undef FWD
    if (remaining != 0) return -1;
    *max_sym = sym - 1;
    *log_out = al;
    return (int)((bit + 7) >> 3);
}

static int build_fse(fse_table *t, const short *norm, int max_sym, int log)
{
    const int size = 1 << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    uint16_t next[256];
    uint8_t spread[512];
    int high = size - 1, pos = 0, s, i;
    for (s = 0; s <= max_sym; s++) {
        if (norm[s] == -1) { spread[high--] = (uint8_t)s; next[s] = 1; } else next[s] = (uint16_t)norm[s];
    }
    for (s = 0; s <= max_sym; s++)
        for (i = 0; i < norm[s]; i++) {
            spread[pos] = (uint8_t)s;
            do pos = (pos + step) & mask; while (pos > high);
        }
    if (pos != 0) return -1;
    for (i = 0; i < size; i++) {
        const int sym = spread[i];
        const uint32_t x = next[sym]++;
        const int nb = log - hibit(x);
        t->cell[i].sym = (uint8_t)sym; t->cell[i].nbits = (uint8_t)nb; t->cell[i].next = (uint16_t)((x << nb) - size);
    }
    t->log = log; t->rle = 0;
    return 0;
}

/* Huffman tree description -> decoding table.  Returns bytes used or -1. */
static int read_huf(huf_table *h, const uint8_t *src, long long n)
{
    uint8_t w[256];
    int nw = 0, hb, used, i, s;
    uint32_t total = 0, rest;
    if (n < 1) return -1;
    hb = src[0];
    if (hb >= 128) {
        nw = hb - 127;
        used = 1 + (nw + 1) / 2;
        if (used > n) return -1;
        for (i = 0; i < nw; i++) w[i] = (i & 1) ? (src[1 + i / 2] & 15) : (src[1 + i / 2] >> 4);
    } else {
        short norm[16];
        int max_sym = 12, log, hdr;
        fse_table t;
        back_t b;
        uint32_t s1, s2;
        used = 1 + hb;
        if (used > n || hb < 2) return -1;
        hdr = read_norm(src + 1, hb, norm, &max_sym, &log, 6);
        if (hdr < 0 || build_fse(&t, norm, max_sym, log)) return -1;
        if (back_init(&b, src + 1 + hdr, hb - hdr)) return -1;
        s1 = back_read(&b, log); s2 = back_read(&b, log);
        for (;;) {                                                /* two interleaved states (fse_decompress.c:232-300) */
            if (nw > 254) return -1;
            w[nw++] = t.cell[s1].sym;
            s1 = t.cell[s1].next + back_read(&b, t.cell[s1].nbits);
            if (b.bits < 0) { w[nw++] = t.cell[s2].sym; break; }
            if (nw > 254) return -1;
            w[nw++] = t.cell[s2].sym;
            s2 = t.cell[s2].next + back_read(&b, t.cell[s2].nbits);
            if (b.bits < 0) { w[nw++] = t.cell[s1].sym; break; }
        }
    }
    for (i = 0; i < nw; i++) { if (w[i] > 12) return -1; total += w[i] ? 1u << (w[i] - 1) : 0; }
    if (total == 0) return -1;
    h->log = hibit(total) + 1;
    if (h->log > 12) return -1;
    rest = (1u << h->log) - total;
    if (rest & (rest - 1)) return -1;                             /* the implied last weight must be a power of two */
    w[nw++] = (uint8_t)(hibit(rest) + 1);
    {   /* canonical layout: weight 1 first, symbol order inside a weight */
        uint32_t pos = 0;
        int wt;
        for (wt = 1; wt <= h->log; wt++)
            for (s = 0; s < nw; s++)
                if (w[s] == wt) {
                    const uint32_t span = 1u << (wt - 1);
                    uint32_t k;
                    for (k = 0; k < span; k++) { h->sym[pos + k] = (uint8_t)s; h->len[pos + k] = (uint8_t)(h->log + 1 - wt); }
                    pos += span;
                }
        if (pos != (1u << h->log)) return -1;
    }
    h->ok = 1;
    return used;
}

static int huf_stream(const huf_table *h, uint8_t *dst, long long count, const uint8_t *src, long long n)
{
    back_t b;
    long long i;
    if (back_init(&b, src, n)) return -1;
    for (i = 0; i < count; i++) {
        back_t peek = b;
        const uint32_t idx = back_read(&peek, h->log);
        dst[i] = h->sym[idx];
        b.bits -= h->len[idx];
    }
    return b.bits == 0 ? 0 : -1;
}

static const uint32_t LL_BASE[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40,
                                     48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
static const uint8_t LL_BITS[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static const uint32_t ML_BASE[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
                                     35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
static const uint8_t ML_BITS[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                    1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static const short LL_DEF[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
static const short ML_DEF[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
static const short OF_DEF[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};

/* one of the three sequence tables; which: 0 LL, 1 OF, 2 ML.  Returns bytes used or -1. */
static int seq_table(int mode, fse_table *t, int *ok, int which, const uint8_t *src, long long n)
{
    const int max_sym = which == 0 ? 35 : which == 1 ? 31 : 52, max_log = which == 1 ? 8 : 9;
    short norm[64];
    if (mode == 0) {
        const short *d = which == 0 ? LL_DEF : which == 1 ? OF_DEF : ML_DEF;
        if (build_fse(t, d, which == 0 ? 35 : which == 1 ? 28 : 52, which == 1 ? 5 : 6)) return -1;
        *ok = 1;
        return 0;
    }
    if (mode == 1) {
        if (n < 1 || src[0] > max_sym) return -1;
        t->rle = 1; t->log = 0; t->cell[0].sym = src[0]; t->cell[0].nbits = 0; t->cell[0].next = 0;
        *ok = 1;
        return 1;
    }
    if (mode == 2) {
        int ms = max_sym, log, used = read_norm(src, n, norm, &ms, &log, max_log);
        if (used < 0 || build_fse(t, norm, ms, log)) return -1;
        *ok = 1;
        return used;
    }
    return *ok ? 0 : -1;
}

/* one compressed block: literals + sequences.  Returns the new output position or -1. */
static long long block(zctx *z, uint8_t *dst, long long op, long long cap, const uint8_t *src, long long n)
{
    int ltype, sf, streams = 1;
    long long hdr, regen, comp = 0, nseq, lit_pos = 0, k;
    const uint8_t *lit;
    if (n < 1) return -1;
    ltype = src[0] & 3; sf = (src[0] >> 2) & 3;
    if (ltype < 2) {
        if (sf == 0 || sf == 2) { hdr = 1; regen = src[0] >> 3; }
        else if (sf == 1) { if (n < 2) return -1; hdr = 2; regen = (src[0] >> 4) | ((long long)src[1] << 4); }
        else { if (n < 3) return -1; hdr = 3; regen = (src[0] >> 4) | ((long long)src[1] << 4) | ((long long)src[2] << 12); }
        if (regen > ZBLOCK_MAX) return -1;
        if (ltype == 0) { if (hdr + regen > n) return -1; lit = src + hdr; hdr += regen; }
        else { if (hdr + 1 > n) return -1; memset(z->lit, src[hdr], (size_t)regen); lit = z->lit; hdr += 1; }
    } else {
        uint64_t v = 0;
        int i;
        const int hl = sf <= 1 ? 3 : sf == 2 ? 4 : 5;
        if (n < hl) return -1;
        for (i = 0; i < hl; i++) v |= (uint64_t)src[i] << (8 * i);
        hdr = hl;
        if (sf <= 1) { regen = (v >> 4) & 0x3FF; comp = (v >> 14) & 0x3FF; streams = sf == 0 ? 1 : 4; }
        else if (sf == 2) { regen = (v >> 4) & 0x3FFF; comp = (v >> 18) & 0x3FFF; streams = 4; }
        else { regen = (v >> 4) & 0x3FFFF; comp = (v >> 22) & 0x3FFFF; streams = 4; }
        if (regen > ZBLOCK_MAX || hdr + comp > n) return -1;
        {
            const uint8_t *cs = src + hdr;
            long long cn = comp;
            if (ltype == 2) { int used = read_huf(&z->huf, cs, cn); if (used < 0) return -1; cs += used; cn -= used; }
            else if (!z->huf.ok) return -1;
            if (streams == 1) { if (huf_stream(&z->huf, z->lit, regen, cs, cn)) return -1; }
            else {
                long long s1, s2, s3, s4, seg = (regen + 3) / 4;
                if (cn < 10) return -1;
                s1 = cs[0] | (cs[1] << 8); s2 = cs[2] | (cs[3] << 8); s3 = cs[4] | (cs[5] << 8);
                s4 = cn - 6 - s1 - s2 - s3;
                if (s4 < 1 || 3 * seg > regen) return -1;
                if (huf_stream(&z->huf, z->lit, seg, cs + 6, s1) || huf_stream(&z->huf, z->lit + seg, seg, cs + 6 + s1, s2) ||
                    huf_stream(&z->huf, z->lit + 2 * seg, seg, cs + 6 + s1 + s2, s3) ||
                    huf_stream(&z->huf, z->lit + 3 * seg, regen - 3 * seg, cs + 6 + s1 + s2 + s3, s4))
                    return -1;
            }
        }
        lit = z->lit;
        hdr += comp;
    }
    /* sequences section */
    {
        const uint8_t *sp = src + hdr;
        long long sn = n - hdr, sh = 1;
        if (sn < 1) return -1;
        nseq = sp[0];
        if (nseq >= 128) {
            if (nseq == 255) { if (sn < 3) return -1; nseq = sp[1] + (sp[2] << 8) + 0x7F00; sh = 3; }
            else { if (sn < 2) return -1; nseq = ((nseq - 128) << 8) + sp[1]; sh = 2; }
        }
        if (nseq == 0) { if (sn != sh) return -1; }
        else {
            int modes, used;
            back_t b;
            uint32_t sl, so, sm;
            if (sh + 1 > sn) return -1;
            modes = sp[sh++];
            if (modes & 3) return -1;                             /* reserved bits */
            if ((used = seq_table((modes >> 6) & 3, &z->ll, &z->ll_ok, 0, sp + sh, sn - sh)) < 0) return -1;
            sh += used;
            if ((used = seq_table((modes >> 4) & 3, &z->of, &z->of_ok, 1, sp + sh, sn - sh)) < 0) return -1;
            sh += used;
            if ((used = seq_table((modes >> 2) & 3, &z->ml, &z->ml_ok, 2, sp + sh, sn - sh)) < 0) return -1;
            sh += used;
            if (back_init(&b, sp + sh, sn - sh)) return -1;
            sl = back_read(&b, z->ll.log); so = back_read(&b, z->of.log); sm = back_read(&b, z->ml.log);
            for (k = 0; k < nseq; k++) {
                const int lc = z->ll.cell[sl].sym, oc = z->of.cell[so].sym, mc = z->ml.cell[sm].sym;
                uint32_t offv, offset;
                long long mlen, llen, i;
                if (oc > 31) return -1;
                offv = (1u << oc) + back_read(&b, oc);            /* offset value: > 3 real offset + 3, 1..3 repeat codes */
                mlen = ML_BASE[mc] + back_read(&b, ML_BITS[mc]);
                llen = LL_BASE[lc] + back_read(&b, LL_BITS[lc]);
                if (offv > 3) { offset = offv - 3; z->rep[2] = z->rep[1]; z->rep[1] = z->rep[0]; z->rep[0] = offset; }
                else {
                    uint32_t idx = offv - 1 + (llen == 0);       /* 0, 1, 2, or 3 = rep[0] - 1 */
                    if (idx == 0) offset = z->rep[0];
                    else {
                        offset = idx < 3 ? z->rep[idx] : z->rep[0] - 1;
                        if (offset == 0) return -1;
                        if (idx > 1) z->rep[2] = z->rep[1];
                        z->rep[1] = z->rep[0];
                        z->rep[0] = offset;
                    }
                }
                if (k + 1 < nseq) {                               /* state updates: LL, ML, OF */
                    sl = z->ll.cell[sl].next + back_read(&b, z->ll.cell[sl].nbits);
                    sm = z->ml.cell[sm].next + back_read(&b, z->ml.cell[sm].nbits);
                    so = z->of.cell[so].next + back_read(&b, z->of.cell[so].nbits);
                }
                if (b.bits < 0) return -1;
                if (llen > regen - lit_pos || llen + mlen > cap - op) return -1;
                memcpy(dst + op, lit + lit_pos, (size_t)llen);
                op += llen; lit_pos += llen;
                if ((long long)offset > op) return -1;
                for (i = 0; i < mlen; i++) dst[op + i] = dst[op + i - offset];
                op += mlen;
            }
            if (b.bits != 0) return -1;                           /* the stream is consumed exactly */
        }
    }
    if (regen - lit_pos > cap - op) return -1;
    memcpy(dst + op, lit + lit_pos, (size_t)(regen - lit_pos));
    return op + (regen - lit_pos);
}

/* ZSTD_decompress restated: every frame of src into dst.  Returns the decoded size or -1. */
long long fmo_zstd_decompress(uint8_t *dst, long long cap, const uint8_t *src, long long n)
{
    long long ip = 0, op = 0;
    zctx *z = (zctx *)malloc(sizeof(zctx));
    if (!z) return -1;
#define FAIL do { free(z); return -1; } while (0)
    while (ip < n) {
        uint32_t magic;
        int fhd, single, fcs_id, did, ck, fsz, i;
        uint64_t fcs = 0, window = 0;
        long long frame_start = op, bmax;
        if (n - ip < 4) FAIL;
        magic = src[ip] | (src[ip + 1] << 8) | (src[ip + 2] << 16) | ((uint32_t)src[ip + 3] << 24);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {               /* skippable frame */
            uint32_t sz;
            if (n - ip < 8) FAIL;
            sz = src[ip + 4] | (src[ip + 5] << 8) | (src[ip + 6] << 16) | ((uint32_t)src[ip + 7] << 24);
            if ((long long)sz + 8 > n - ip) FAIL;
            ip += 8 + sz;
            continue;
        }
        if (magic != 0xFD2FB528u || n - ip < 5) FAIL;
        fhd = src[ip + 4]; ip += 5;
        did = fhd & 3; ck = (fhd >> 2) & 1; single = (fhd >> 5) & 1; fcs_id = fhd >> 6;
        if (fhd & 8) FAIL;
        if (!single) {
            int wl;
            if (ip >= n) FAIL;
            wl = (src[ip] >> 3) + 10;
            window = (1ull << wl) + ((1ull << wl) >> 3) * (src[ip] & 7);
            ip++;
        }
        if (did) FAIL;                                            /* dictionaries are not part of this path */
        fsz = fcs_id == 0 ? single : fcs_id == 1 ? 2 : fcs_id == 2 ? 4 : 8;
        if (n - ip < fsz) FAIL;
        for (i = 0; i < fsz; i++) fcs |= (uint64_t)src[ip + i] << (8 * i);
        if (fsz == 2) fcs += 256;
        ip += fsz;
        if (single) window = fcs;
        bmax = window < ZBLOCK_MAX ? (long long)window : ZBLOCK_MAX;
        z->huf.ok = 0; z->ll_ok = z->of_ok = z->ml_ok = 0;
        z->rep[0] = 1; z->rep[1] = 4; z->rep[2] = 8;
        for (;;) {
            uint32_t bh;
            int last, type;
            long long bsz;
            if (n - ip < 3) FAIL;
            bh = src[ip] | (src[ip + 1] << 8) | (src[ip + 2] << 16);
            ip += 3;
            last = bh & 1; type = (bh >> 1) & 3; bsz = bh >> 3;
            if (type == 3 || bsz > bmax) FAIL;
            if (type == 1) {
                if (n - ip < 1 || bsz > cap - op) FAIL;
                memset(dst + op, src[ip], (size_t)bsz);
                op += bsz; ip += 1;
            } else {
                if (bsz > n - ip) FAIL;
                if (type == 0) {
                    if (bsz > cap - op) FAIL;
                    memcpy(dst + op, src + ip, (size_t)bsz);
                    op += bsz;
                } else {
                    const long long r = block(z, dst + frame_start, op - frame_start, cap - frame_start, src + ip, bsz);
                    if (r < 0 || r - (op - frame_start) > bmax) FAIL;
                    op = frame_start + r;
                }
                ip += bsz;
            }
            if (last) break;
        }
        if (fsz && (uint64_t)(op - frame_start) != fcs) FAIL;
        if (ck) { if (n - ip < 4) FAIL; ip += 4; }              /* content checksum (XXH64): 4mz frames carry none; not verified */
    }
#undef FAIL
    free(z);
    return op;
}

/* decodeFourMZ restated (native/4mc.c:709-857): one or more concatenated 4mz streams.  Same return
 * convention as fmo_4mc_decompress. */
static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

long long fmo_4mz_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap)
{
    size_t pos = 0, op = 0;
    while (pos < n) {
        const size_t op_before = op;
        if (n - pos < 12) return FMO_ERR_CONTENT;
        if (be32(in + pos) != 0x344D5A00u || be32(in + pos + 4) != 1 || be32(in + pos + 8) != fmo_xxh32(in + pos, 8, 0)) return FMO_ERR_CONTENT;
        pos += 12;
        for (;;) {
            uint32_t u, c, ck;
            if (n - pos < 12) return FMO_ERR_INPUT;
            u = be32(in + pos); c = be32(in + pos + 4); ck = be32(in + pos + 8);
            pos += 12;
            if (u == 0 && c == 0 && ck == 0) break;
            if (c > (4u << 20) || (u != c && u > (4u << 20))) return FMO_ERR_CONTENT;
            if (n - pos < c) return FMO_ERR_INPUT;
            if (fmo_xxh32(in + pos, c, 0) != ck) return FMO_ERR_CONTENT;
            if (u > cap - op) return FMO_ERR_OUTPUT;
            if (u == c) { memcpy(out + op, in + pos, c); op += u; }
            else {
                /* native/4mc.c:810-815: whatever ZSTD_decompress returns is written (`filesize += decodedBytes`) */
                const long long dsz = fmo_zstd_decompress(out + op, u, in + pos, c);
                if (dsz < 0) return FMO_ERR_CONTENT;
                op += (size_t)dsz;
            }
            pos += c;
        }
        {
            uint32_t fsize;
            if (n - pos < 4) return FMO_ERR_GENERIC;
            fsize = be32(in + pos);
            if (fsize < 8 || n - pos < fsize) return FMO_ERR_INPUT;
            if (be32(in + pos + fsize - 4) != fmo_xxh32(in + pos, fsize - 4, 0)) return FMO_ERR_CONTENT;
            pos += fsize;
        }
        if (op == op_before) break;           /* native/4mc.c:943-947 `do {...} while (decodedSize)` */
    }
    return (long long)op;
}
