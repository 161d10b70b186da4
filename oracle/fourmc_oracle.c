/*
 * fourmc_oracle.c -- CPU restatement of the 4mc LZ4 block path.  TEST INFRASTRUCTURE ONLY
 * (see fourmc_oracle.h).  Plain C99, no dependencies; written from the format rules and the
 * reference's observable behaviour, each function citing the reference lines it follows.
 */
#include "fourmc_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ helpers */

static uint32_t rd_le32(const uint8_t *p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static uint32_t rd_be32(const uint8_t *p)
{
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}
static void wr_be32(uint8_t *p, uint32_t v)
{
    p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}
static uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

/* ------------------------------------------------------------------ XXH32 */
/* native/lz4/xxhash.c: primes :252-256 (names there PRIME32_1..5), round :269-275,
 * avalanche :278-286, tail :290-348, body :351-389, entry :392-416. */

#define XP1 0x9E3779B1u
#define XP2 0x85EBCA77u
#define XP3 0xC2B2AE3Du
#define XP4 0x27D4EB2Fu
#define XP5 0x165667B1u

static uint32_t xxh_round(uint32_t acc, uint32_t lane)
{
    acc += lane * XP2;
    acc = rotl32(acc, 13);
    return acc * XP1;
}

uint32_t fmo_xxh32(const void *data, size_t len, uint32_t seed)
{
    const uint8_t *p = (const uint8_t *)data;
    const uint8_t *end = p + len;
    uint32_t h;

    if (len >= 16) {
        uint32_t v1 = seed + XP1 + XP2, v2 = seed + XP2, v3 = seed, v4 = seed - XP1;
        const uint8_t *limit = end - 15;
        do {
            v1 = xxh_round(v1, rd_le32(p));
            v2 = xxh_round(v2, rd_le32(p + 4));
            v3 = xxh_round(v3, rd_le32(p + 8));
            v4 = xxh_round(v4, rd_le32(p + 12));
            p += 16;
        } while (p < limit);
        h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    } else {
        h = seed + XP5;
    }
    h += (uint32_t)len;
    while (p + 4 <= end) {
        h += rd_le32(p) * XP3;
        h = rotl32(h, 17) * XP4;
        p += 4;
    }
    while (p < end) {
        h += (uint32_t)(*p) * XP5;
        h = rotl32(h, 11) * XP1;
        p++;
    }
    h ^= h >> 15; h *= XP2;
    h ^= h >> 13; h *= XP3;
    h ^= h >> 16;
    return h;
}

/* ------------------------------------------------------------------ LZ4 decode */
/*
 * Restates LZ4_decompress_generic(decode_full_block, noDict) as instantiated by
 * LZ4_decompress_safe (native/lz4/lz4.c:1936-2350) on x86-64, where LZ4_FAST_DEC_LOOP is on.
 *
 * The reference has two loops (a "fast" loop used while at least 64 bytes of output space remain,
 * :1996-2115, and a "safe" loop, :2120-2328) whose accept/reject conditions differ slightly on
 * malformed input, so both are kept as two states of one machine over integer positions
 * (ip, op) -- every accept/reject decision depends on positions and lengths only, never on data.
 * All copies are expressed as their observable effect: literals are a plain copy, a match is a
 * byte-sequential copy from op-offset (so it self-overlaps), and offset 0 yields zero bytes
 * (:453-549 LZ4_memcpy_using_offset_base / :2300-2303 write32(op,0) before the self-copy).
 */

enum { ML_MINMATCH = 4, LASTLITERALS = 5, MFLIMIT = 12, FAST_DISTANCE = 64 };

/* read_variable_length, native/lz4/lz4.c:1903-1928: *ip advances; returns -1 on error */
static long long lz4_rvl(const uint8_t *src, long long *ip, long long ilimit, int initial_check)
{
    long long length = 0;
    unsigned s;
    if (initial_check && *ip >= ilimit) return -1;
    do {
        s = src[*ip];
        (*ip)++;
        length += s;
        if (*ip > ilimit) return -1;
    } while (s == 255);
    return length;
}

static void lz4_match_copy(uint8_t *dst, long long op, long long offset, long long len)
{
    long long i;
    if (offset == 0) { memset(dst + op, 0, (size_t)len); return; }
    for (i = 0; i < len; i++) dst[op + i] = dst[op - offset + i];
}

int fmo_lz4_decompress_safe(const uint8_t *src, uint8_t *dst, int src_size, int dst_capacity)
{
    long long ip = 0, op = 0;
    const long long iend = src_size, oend = dst_capacity;
    long long length, offset, match, cpy;
    unsigned token;
    int fast;

    if (src == NULL || dst_capacity < 0) return -1;                     /* :1951 */
    if (dst_capacity == 0) return (src_size == 1 && src[0] == 0) ? 0 : -1;   /* :1977-1981 */
    if (src_size == 0) return -1;                                        /* :1982 */

    fast = (oend - op) >= FAST_DISTANCE;                                 /* :1990 */

    for (;;) {
        if (fast) {
            /* ---- fast loop iteration, :1996-2115 ---- */
            token = src[ip++];
            length = token >> 4;
            if (length == 15) {
                long long addl = lz4_rvl(src, &ip, iend - 15, 1);
                if (addl < 0) goto error;
                length += addl;
                cpy = op + length;
                if (cpy > oend - 32 || ip + length > iend - 32) { fast = 0; goto safe_literal_copy; }
            } else {
                cpy = op + length;
                if (ip > iend - 17) { fast = 0; goto safe_literal_copy; }
            }
            memcpy(dst + op, src + ip, (size_t)length);
            ip += length; op = cpy;

            offset = (long long)src[ip] | ((long long)src[ip + 1] << 8); ip += 2;
            match = op - offset;
            length = token & 15;
            if (length == 15) {
                long long addl = lz4_rvl(src, &ip, iend - LASTLITERALS + 1, 0);
                if (addl < 0) goto error;
                length += addl + ML_MINMATCH;
                if (match < 0) goto error;                               /* :2041 */
                if (op + length >= oend - FAST_DISTANCE) { fast = 0; goto safe_match_copy; }
            } else {
                length += ML_MINMATCH;
                if (op + length >= oend - FAST_DISTANCE) { fast = 0; goto safe_match_copy; }
                /* :2051-2063 is a copy shortcut with the same effect as below */
            }
            if (match < 0) goto error;                                   /* :2065 */
            lz4_match_copy(dst, op, offset, length);
            op += length;
            continue;
        }

        /* ---- safe loop iteration, :2120-2328 ---- */
        token = src[ip++];
        length = token >> 4;

        /* two-stage shortcut :2132-2161 -- note it skips the end-of-input test below */
        if (length != 15 && ip < iend - 16 && op <= oend - 32) {
            memcpy(dst + op, src + ip, (size_t)length);
            op += length; ip += length;
            length = token & 15;
            offset = (long long)src[ip] | ((long long)src[ip + 1] << 8); ip += 2;
            match = op - offset;
            if (length != 15 && offset >= 8 && match >= 0) {
                lz4_match_copy(dst, op, offset, length + ML_MINMATCH);
                op += length + ML_MINMATCH;
                continue;
            }
            goto copy_match;
        }

        if (length == 15) {
            long long addl = lz4_rvl(src, &ip, iend - 15, 1);
            if (addl < 0) goto error;
            length += addl;
        }
        cpy = op + length;
safe_literal_copy:
        if (cpy > oend - MFLIMIT || ip + length > iend - (2 + 1 + LASTLITERALS)) {
            /* must be the last sequence, :2203-2213 */
            if (ip + length != iend || cpy > oend) goto error;
            memmove(dst + op, src + ip, (size_t)length);
            ip += length; op += length;
            break;
        }
        memcpy(dst + op, src + ip, (size_t)length);
        ip += length; op = cpy;

        offset = (long long)src[ip] | ((long long)src[ip + 1] << 8); ip += 2;
        match = op - offset;
        length = token & 15;
copy_match:
        if (length == 15) {
            long long addl = lz4_rvl(src, &ip, iend - LASTLITERALS + 1, 0);
            if (addl < 0) goto error;
            length += addl;
        }
        length += ML_MINMATCH;
safe_match_copy:
        if (match < 0) goto error;                                       /* :2250 */
        cpy = op + length;
        if (cpy > oend - MFLIMIT) {                                      /* :2315-2317 */
            if (cpy > oend - LASTLITERALS) {
                /* the reference has already written up to 8 bytes here; harmless, the call fails */
                goto error;
            }
        }
        lz4_match_copy(dst, op, offset, length);
        op = cpy;
    }
    return (int)op;

error:
    return (int)(-ip) - 1;                                               /* :2337 */
}

/* ------------------------------------------------------------------ LZ4 encode */

int fmo_lz4_compress_bound(int n)
{
    /* native/lz4/lz4.h:211-212 */
    return ((unsigned)n > 0x7E000000u) ? 0 : n + n / 255 + 16;
}

static int emit_len(uint8_t *dst, int op, int cap, int len)
{
    /* LZ4 length continuation bytes: 255,255,...,rest */
    while (len >= 255) { if (op >= cap) return -1; dst[op++] = 255; len -= 255; }
    if (op >= cap) return -1;
    dst[op++] = (uint8_t)len;
    return op;
}

int fmo_lz4_compress(const uint8_t *src, uint8_t *dst, int n, int max_out)
{
    /* greedy parse, 2^16-entry table of 4-byte hashes; end rules per native/lz4/lz4.c:243-247:
     * last match must start >= 12 bytes before the end, last 5 bytes are literals */
    enum { HLOG = 16 };
    int *table;
    int ip = 0, anchor = 0, op = 0;
    const int mflimit = n - MFLIMIT;       /* last position where a match may start */
    const int matchlimit = n - LASTLITERALS;

    if (n < 0 || max_out <= 0) return 0;
    table = (int *)malloc(sizeof(int) << HLOG);
    if (!table) return 0;
    memset(table, 0xFF, sizeof(int) << HLOG);

    while (n >= 13 && ip <= mflimit) {
        uint32_t seq = rd_le32(src + ip);
        uint32_t h = (seq * 2654435761u) >> (32 - HLOG);
        int cand = table[h];
        table[h] = ip;
        if (cand >= 0 && ip - cand <= 65535 && rd_le32(src + cand) == seq) {
            int mlen = 4, lit = ip - anchor, tok;
            while (ip + mlen < matchlimit && src[cand + mlen] == src[ip + mlen]) mlen++;
            /* token */
            if (op >= max_out) goto overflow;
            tok = op++;
            dst[tok] = (uint8_t)((lit >= 15 ? 15 : lit) << 4);
            if (lit >= 15) { op = emit_len(dst, op, max_out, lit - 15); if (op < 0) goto overflow; }
            if (op + lit + 2 > max_out) goto overflow;
            memcpy(dst + op, src + anchor, (size_t)lit); op += lit;
            dst[op++] = (uint8_t)(ip - cand); dst[op++] = (uint8_t)((ip - cand) >> 8);
            if (mlen - 4 >= 15) {
                dst[tok] |= 15;
                op = emit_len(dst, op, max_out, mlen - 4 - 15); if (op < 0) goto overflow;
            } else dst[tok] |= (uint8_t)(mlen - 4);
            ip += mlen; anchor = ip;
        } else {
            ip++;
        }
    }
    {   /* last literals */
        int lit = n - anchor;
        if (op >= max_out) goto overflow;
        dst[op++] = (uint8_t)((lit >= 15 ? 15 : lit) << 4);
        if (lit >= 15) { op = emit_len(dst, op, max_out, lit - 15); if (op < 0) goto overflow; }
        if (op + lit > max_out) goto overflow;
        memcpy(dst + op, src + anchor, (size_t)lit); op += lit;
    }
    free(table);
    return op;
overflow:
    free(table);
    return 0;
}

/* ------------------------------------------------------------------ container */

size_t fmo_4mc_bound(size_t n)
{
    size_t blocks = (n + FMO_BLOCKSIZE - 1) / FMO_BLOCKSIZE;
    return 12 + n + 12 * blocks + 12 + 20 + 4 * blocks;
}

long long fmo_4mc_compress(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap)
{
    /* native/4mc.c:263-362 */
    size_t pos = 0, off = 0, nblocks = (n + FMO_BLOCKSIZE - 1) / FMO_BLOCKSIZE, b;
    uint64_t *starts;
    size_t fsize;
    if (out_cap < fmo_4mc_bound(n)) return FMO_ERR_OUTPUT;
    starts = (uint64_t *)malloc(sizeof(uint64_t) * (nblocks ? nblocks : 1));
    if (!starts) return FMO_ERR_GENERIC;

    wr_be32(out, FMO_MAGIC_4MC); wr_be32(out + 4, FMO_VERSION);
    wr_be32(out + 8, fmo_xxh32(out, 8, 0));
    pos = 12;
    for (b = 0; b < nblocks; b++) {
        size_t u = n - off < FMO_BLOCKSIZE ? n - off : FMO_BLOCKSIZE;
        int c = fmo_lz4_compress(in + off, out + pos + 12, (int)u, (int)u - 1);   /* :301 */
        starts[b] = pos;
        wr_be32(out + pos, (uint32_t)u);
        if (c > 0) {
            wr_be32(out + pos + 4, (uint32_t)c);
            wr_be32(out + pos + 8, fmo_xxh32(out + pos + 12, (size_t)c, 0));
            pos += 12 + (size_t)c;
        } else {                                                                   /* :318-329 */
            memcpy(out + pos + 12, in + off, u);
            wr_be32(out + pos + 4, (uint32_t)u);
            wr_be32(out + pos + 8, fmo_xxh32(in + off, u, 0));
            pos += 12 + u;
        }
        off += u;
    }
    memset(out + pos, 0, 12); pos += 12;                                           /* :335-341 */
    fsize = 20 + 4 * nblocks;                                                      /* :117 */
    wr_be32(out + pos, (uint32_t)fsize); wr_be32(out + pos + 4, 1);
    for (b = 0; b < nblocks; b++)
        wr_be32(out + pos + 8 + 4 * b, (uint32_t)(b == 0 ? starts[0] : starts[b] - starts[b - 1]));
    wr_be32(out + pos + 8 + 4 * nblocks, (uint32_t)fsize);
    wr_be32(out + pos + 12 + 4 * nblocks, FMO_MAGIC_4MC);
    wr_be32(out + pos + 16 + 4 * nblocks, fmo_xxh32(out + pos, fsize - 4, 0));
    pos += fsize;
    free(starts);
    return (long long)pos;
}

long long fmo_4mc_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap)
{
    /* native/4mc.c:896-912 (loop over concatenated streams), :862-880, :560-707 */
    size_t pos = 0, opos = 0;
    while (pos < n) {
        uint32_t fsize;
        const size_t opos_before = opos;
        if (n - pos < 4) return FMO_ERR_CONTENT;                     /* magic unreadable :868 */
        if (rd_be32(in + pos) != FMO_MAGIC_4MC) return FMO_ERR_CONTENT;       /* :873 */
        if (n - pos < 12) return FMO_ERR_CONTENT;                    /* unreadable header :577 */
        if (rd_be32(in + pos + 4) != FMO_VERSION) return FMO_ERR_CONTENT;     /* :583 */
        if (rd_be32(in + pos + 8) != fmo_xxh32(in + pos, 8, 0)) return FMO_ERR_CONTENT; /* :584 */
        pos += 12;
        for (;;) {
            uint32_t u, c, ck;
            if (n - pos < 12) return FMO_ERR_INPUT;                  /* :610 */
            u = rd_be32(in + pos); c = rd_be32(in + pos + 4); ck = rd_be32(in + pos + 8);
            pos += 12;
            if (u == 0 && c == 0 && ck == 0) break;                  /* :616 */
            if (c > FMO_BLOCKSIZE) return FMO_ERR_CONTENT;           /* :618 */
            if (n - pos < c) return FMO_ERR_INPUT;                   /* :632 */
            if (fmo_xxh32(in + pos, c, 0) != ck) return FMO_ERR_CONTENT;      /* :637,:645 */
            if (u == c) {                                            /* :635 stored */
                if (out_cap - opos < u) return FMO_ERR_OUTPUT;
                memcpy(out + opos, in + pos, u);
                opos += u;
            } else {
                int d;
                if (u > FMO_BLOCKSIZE) return FMO_ERR_CONTENT;       /* :651 */
                if (out_cap - opos < u) return FMO_ERR_OUTPUT;
                d = fmo_lz4_decompress_safe(in + pos, out + opos, (int)c, (int)u);   /* :661 */
                if (d < 0) return FMO_ERR_CONTENT;                   /* :662 */
                opos += (size_t)d;
            }
            pos += c;
        }
        /* footer :670-688 */
        if (n - pos < 4) return FMO_ERR_GENERIC;                     /* :672 exit(1) */
        fsize = rd_be32(in + pos);
        if (fsize < 4 || n - pos < fsize) return FMO_ERR_INPUT;      /* :680 */
        if (fsize < 8) return FMO_ERR_CONTENT;
        if (fmo_xxh32(in + pos, fsize - 4, 0) != rd_be32(in + pos + fsize - 4)) return FMO_ERR_CONTENT; /* :685 */
        if (rd_be32(in + pos + 4) != 1) return FMO_ERR_CONTENT;      /* :687 */
        pos += fsize;
        if (opos == opos_before) break;       /* :909-913 `do {...} while (decodedSize)`: a stream that decodes to nothing ends the loop */
    }
    return (long long)opos;
}

long long fmo_4mc_read_index(const uint8_t *file, size_t file_size, uint64_t magic,
                             int64_t *offsets, size_t max_blocks)
{
    /* FourMcInputStream.java:163-239 */
    uint32_t fsize, ck;
    const uint8_t *f;
    size_t nb, i;
    int64_t cur = 0;
    if (file_size < 12 + 20) return 0;                                /* :166 empty index */
    fsize = rd_be32(file + file_size - 12);
    if (rd_be32(file + file_size - 8) != (uint32_t)magic) return FMO_ERR_CONTENT;   /* :196 */
    ck = rd_be32(file + file_size - 4);
    if (fsize >= file_size - 12) return FMO_ERR_CONTENT;              /* :199 */
    if (fsize < 20) return FMO_ERR_CONTENT;
    f = file + file_size - fsize;
    if (rd_be32(f) != fsize) return FMO_ERR_CONTENT;                  /* :217 */
    if (rd_be32(f + 4) != FMO_VERSION) return FMO_ERR_CONTENT;        /* :221 */
    if (ck != fmo_xxh32(f, fsize - 4, 0)) return FMO_ERR_CONTENT;     /* :225 */
    nb = (fsize - 20) / 4;
    for (i = 0; i < nb; i++) {
        cur += rd_be32(f + 8 + 4 * i);
        if (i < max_blocks) offsets[i] = cur;
    }
    return (long long)nb;
}

/* ------------------------------------------------------------------ block index */

/* java.util.Arrays.binarySearch semantics: index if found, else -(insertion point)-1 */
static int bsearch_i64(const int64_t *a, int n, int64_t key)
{
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1);
        if (a[mid] < key) lo = mid + 1;
        else if (a[mid] > key) hi = mid - 1;
        else return mid;
    }
    return -(lo + 1);
}

int64_t fmo_index_find_next_position(const int64_t *offs, int n, int64_t pos)
{
    /* FourMcBlockIndex.java:92-104 */
    int b = bsearch_i64(offs, n, pos);
    if (b >= 0) return offs[b];
    b = -b - 1;
    if (b > n - 1) return -1;
    return offs[b];
}

int64_t fmo_index_find_belonging_block(const int64_t *offs, int n, int64_t pos)
{
    /* FourMcBlockIndex.java:111-124 */
    int b = bsearch_i64(offs, n, pos);
    if (b >= 0) return b;
    b = -b - 1 - 1;
    if (b > n - 1 || b < 0) return -1;
    return b;
}

int64_t fmo_index_align_slice_start(const int64_t *offs, int n, int64_t start, int64_t end)
{
    /* FourMcBlockIndex.java:142-153 */
    if (start != 0) {
        int64_t ns = fmo_index_find_next_position(offs, n, start);
        if (ns == -1 || ns >= end) return -1;
        start = ns;
    }
    return start;
}

int64_t fmo_index_align_slice_end(const int64_t *offs, int n, int64_t end, int64_t file_size)
{
    /* FourMcBlockIndex.java:163-173 */
    int64_t ne = fmo_index_find_next_position(offs, n, end);
    return ne != -1 ? ne : file_size;
}
