/*
 * fourmc_gen.h -- deterministic synthetic inputs for the BASELINE.json configs.
 *
 * One source for host (gcc/g++) and device (nvcc): every 4 KiB page of the input is a pure
 * function of (seed, global page index), integer arithmetic only, so any 4 MiB block can be
 * regenerated on the CPU and on the GPU bit-identically.  A page is log lines concatenated and
 * hard-truncated at 4096 bytes.
 *
 *   kind 0, log-text line:  "<epoch> host-%03d svc[%d] <LEVEL> req=%08x path=/api/v1/<w>/<w> latency=%dms msg=\"<3-12 w>\"\n"
 *   kind 1, JSON: one object per line with 12 fixed keys (ts, id, user, level, service, region, latency_ms,
 *           status, bytes, ok, tags, msg) and int / float / enum / uuid / short-text values (configs[2])
 *   kind 2, silesia-like mix: the type of every 4 MiB block is drawn from {log-text, JSON, prose, 16-bit PCM
 *           random walk, random bytes (exercises the stored path), sparse binary with zero runs} (configs[3])
 *   <w> is drawn with a power-law skew (u^5 over a 5000-word vocabulary, Zipf-like) and the
 *   words themselves are derived from a hash of their rank, so no tables are needed.
 */
#ifndef FOURMC_GEN_H
#define FOURMC_GEN_H

#include <stdint.h>

#if defined(__CUDACC__)
#define FMG_HD __host__ __device__ __forceinline__
#else
#define FMG_HD static inline
#endif

#define FMG_PAGE 4096u
#define FMG_VOCAB 5000u

typedef struct { uint64_t s; } fmg_rng;

FMG_HD uint64_t fmg_next(fmg_rng *r)
{   /* splitmix64 */
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

FMG_HD uint32_t fmg_below(fmg_rng *r, uint32_t n)
{
    return (uint32_t)(((fmg_next(r) >> 32) * (uint64_t)n) >> 32);
}

FMG_HD uint32_t fmg_put_dec(uint8_t *out, uint32_t pos, uint32_t cap, uint64_t v, int min_digits)
{
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + (int)(v % 10)); v /= 10; } while (v);
    while (n < min_digits) tmp[n++] = '0';
    while (n > 0 && pos < cap) out[pos++] = (uint8_t)tmp[--n];
    return pos;
}

FMG_HD uint32_t fmg_put_str(uint8_t *out, uint32_t pos, uint32_t cap, const char *s)
{
    while (*s && pos < cap) out[pos++] = (uint8_t)*s++;
    return pos;
}

FMG_HD uint32_t fmg_put_word(uint8_t *out, uint32_t pos, uint32_t cap, fmg_rng *r)
{
    /* rank ~ N * u^5  (density ~ rank^(-4/5); the top word is ~18% of all words) */
    uint64_t u = fmg_next(r) >> 32;
    uint64_t y = (u * u) >> 32;
    y = (y * y) >> 32;
    y = (y * u) >> 32;
    uint32_t rank = (uint32_t)((y * FMG_VOCAB) >> 32);
    fmg_rng w; w.s = 0x4D43ull * 0x100000001B3ull + rank;
    uint64_t h = fmg_next(&w);
    uint32_t len = 3u + (uint32_t)(h & 7u);          /* 3..10 letters */
    h >>= 3;
    for (uint32_t i = 0; i < len && pos < cap; i++) {
        out[pos++] = (uint8_t)('a' + (uint32_t)(((h & 31u) * 26u) >> 5));
        h >>= 5;
    }
    return pos;
}

/* Generates page `page` (global index over the whole input) into out[0..4096). */
FMG_HD void fmg_logtext_page(uint64_t seed, uint64_t page, uint8_t *out)
{
    const uint32_t cap = FMG_PAGE;
    fmg_rng r; r.s = seed ^ (page * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull);
    uint64_t epoch = 1700000000ull + page * 37ull;
    uint32_t pos = 0;
    while (pos < cap) {
        epoch += fmg_below(&r, 3);
        pos = fmg_put_dec(out, pos, cap, epoch, 1);
        pos = fmg_put_str(out, pos, cap, " host-");
        pos = fmg_put_dec(out, pos, cap, fmg_below(&r, 200), 3);
        pos = fmg_put_str(out, pos, cap, " svc[");
        pos = fmg_put_dec(out, pos, cap, 1000 + fmg_below(&r, 64), 1);
        uint32_t lv = fmg_below(&r, 100);
        pos = fmg_put_str(out, pos, cap, lv < 70 ? "] INFO req=" : lv < 85 ? "] WARN req=" : lv < 95 ? "] ERROR req=" : "] DEBUG req=");
        uint32_t req = (uint32_t)fmg_next(&r);
        for (int i = 7; i >= 0 && pos < cap; i--) {
            uint32_t d = (req >> (4 * i)) & 15u;
            out[pos++] = (uint8_t)(d < 10 ? '0' + d : 'a' + d - 10);
        }
        pos = fmg_put_str(out, pos, cap, " path=/api/v1/");
        pos = fmg_put_word(out, pos, cap, &r);
        if (pos < cap) out[pos++] = '/';
        pos = fmg_put_word(out, pos, cap, &r);
        pos = fmg_put_str(out, pos, cap, " latency=");
        pos = fmg_put_dec(out, pos, cap, fmg_below(&r, 2000), 1);
        pos = fmg_put_str(out, pos, cap, "ms msg=\"");
        uint32_t nw = 3 + fmg_below(&r, 10);
        for (uint32_t k = 0; k < nw; k++) {
            if (k && pos < cap) out[pos++] = ' ';
            pos = fmg_put_word(out, pos, cap, &r);
        }
        pos = fmg_put_str(out, pos, cap, "\"\n");
    }
}


FMG_HD uint32_t fmg_put_hex(uint8_t *out, uint32_t pos, uint32_t cap, uint64_t v, int digits)
{
    for (int i = digits - 1; i >= 0 && pos < cap; i--) {
        uint32_t d = (uint32_t)(v >> (4 * i)) & 15u;
        out[pos++] = (uint8_t)(d < 10 ? '0' + d : 'a' + d - 10);
    }
    return pos;
}

FMG_HD void fmg_json_page(uint64_t seed, uint64_t page, uint8_t *out)
{
    const uint32_t cap = FMG_PAGE;
    fmg_rng r; r.s = seed ^ (page * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull);
    uint64_t ts = 1700000000000ull + page * 911ull;
    uint32_t pos = 0;
    while (pos < cap) {
        ts += fmg_below(&r, 40);
        pos = fmg_put_str(out, pos, cap, "{\"ts\":");
        pos = fmg_put_dec(out, pos, cap, ts, 1);
        pos = fmg_put_str(out, pos, cap, ",\"id\":\"");
        uint64_t a = fmg_next(&r), b = fmg_next(&r);
        pos = fmg_put_hex(out, pos, cap, a >> 32, 8); if (pos < cap) out[pos++] = '-';
        pos = fmg_put_hex(out, pos, cap, a >> 16, 4); if (pos < cap) out[pos++] = '-';
        pos = fmg_put_hex(out, pos, cap, a, 4); if (pos < cap) out[pos++] = '-';
        pos = fmg_put_hex(out, pos, cap, b >> 48, 4); if (pos < cap) out[pos++] = '-';
        pos = fmg_put_hex(out, pos, cap, b, 12);
        pos = fmg_put_str(out, pos, cap, "\",\"user\":");
        pos = fmg_put_dec(out, pos, cap, 10000 + fmg_below(&r, 50000), 1);
        uint32_t lv = fmg_below(&r, 100);
        pos = fmg_put_str(out, pos, cap, lv < 70 ? ",\"level\":\"info\"" : lv < 85 ? ",\"level\":\"warn\"" : lv < 95 ? ",\"level\":\"error\"" : ",\"level\":\"debug\"");
        pos = fmg_put_str(out, pos, cap, ",\"service\":\"");
        pos = fmg_put_word(out, pos, cap, &r);
        uint32_t rg = fmg_below(&r, 6);
        pos = fmg_put_str(out, pos, cap, rg == 0 ? "\",\"region\":\"eu-west-1\"" : rg == 1 ? "\",\"region\":\"eu-central-1\"" : rg == 2 ? "\",\"region\":\"us-east-1\"" :
                                          rg == 3 ? "\",\"region\":\"us-west-2\"" : rg == 4 ? "\",\"region\":\"ap-south-1\"" : "\",\"region\":\"sa-east-1\"");
        pos = fmg_put_str(out, pos, cap, ",\"latency_ms\":");
        pos = fmg_put_dec(out, pos, cap, fmg_below(&r, 900), 1); if (pos < cap) out[pos++] = '.';
        pos = fmg_put_dec(out, pos, cap, fmg_below(&r, 100), 2);
        uint32_t sc = fmg_below(&r, 100);
        pos = fmg_put_str(out, pos, cap, sc < 80 ? ",\"status\":200" : sc < 88 ? ",\"status\":204" : sc < 94 ? ",\"status\":404" : sc < 98 ? ",\"status\":500" : ",\"status\":503");
        pos = fmg_put_str(out, pos, cap, ",\"bytes\":");
        pos = fmg_put_dec(out, pos, cap, fmg_below(&r, 1u << (8 + fmg_below(&r, 14))), 1);
        pos = fmg_put_str(out, pos, cap, fmg_below(&r, 10) ? ",\"ok\":true,\"tags\":[\"" : ",\"ok\":false,\"tags\":[\"");
        pos = fmg_put_word(out, pos, cap, &r);
        pos = fmg_put_str(out, pos, cap, "\",\"");
        pos = fmg_put_word(out, pos, cap, &r);
        pos = fmg_put_str(out, pos, cap, "\"],\"msg\":\"");
        uint32_t nw = 3 + fmg_below(&r, 6);
        for (uint32_t k = 0; k < nw; k++) {
            if (k && pos < cap) out[pos++] = ' ';
            pos = fmg_put_word(out, pos, cap, &r);
        }
        pos = fmg_put_str(out, pos, cap, "\"}\n");
    }
}

FMG_HD void fmg_mix_page(uint64_t seed, uint64_t page, uint8_t *out)
{
    const uint32_t cap = FMG_PAGE;
    const uint64_t block = page >> 10;                       /* 1024 pages = one 4 MiB block */
    fmg_rng t; t.s = seed * 0x9E3779B97F4A7C15ull + block;
    const uint32_t type = fmg_below(&t, 6);
    if (type == 0) { fmg_logtext_page(seed, page, out); return; }
    if (type == 1) { fmg_json_page(seed, page, out); return; }
    fmg_rng r; r.s = seed ^ (page * 0xD1342543DE82EF95ull + 0x9FB21C651E98DF25ull);
    uint32_t pos = 0;
    if (type == 2) {                                         /* prose: words, punctuation, lines of ~80 */
        uint32_t col = 0;
        while (pos < cap) {
            uint32_t before = pos;
            pos = fmg_put_word(out, pos, cap, &r);
            col += pos - before + 1;
            uint32_t p = fmg_below(&r, 12);
            if (p == 0 && pos < cap) out[pos++] = ',';
            if (p == 1 && pos < cap) out[pos++] = '.';
            if (pos < cap) out[pos++] = (uint8_t)(col > 72 ? '\n' : ' ');
            if (col > 72) col = 0;
        }
    } else if (type == 3) {                                  /* 16-bit PCM-like random walk */
        int32_t x = (int32_t)fmg_below(&r, 65536) - 32768;
        for (; pos + 1 < cap; pos += 2) {
            x += (int32_t)fmg_below(&r, 257) - 128;
            if (x > 32767) x = 32767;
            if (x < -32768) x = -32768;
            out[pos] = (uint8_t)x; out[pos + 1] = (uint8_t)((uint32_t)x >> 8);
        }
    } else if (type == 4) {                                  /* already compressed: random bytes */
        for (; pos + 7 < cap; pos += 8) { uint64_t v = fmg_next(&r); for (int i = 0; i < 8; i++) out[pos + i] = (uint8_t)(v >> (8 * i)); }
    } else {                                                 /* sparse binary: zero runs with short records */
        for (uint32_t i = 0; i < cap; i++) out[i] = 0;
        while (pos < cap) {
            pos += fmg_below(&r, 600);
            uint64_t v = fmg_next(&r);
            for (int i = 0; i < 8 && pos < cap; i++) out[pos++] = (uint8_t)(v >> (8 * i));
        }
    }
}

/* kind: 0 log-text, 1 JSON, 2 silesia-like mix.  Returns 0, or -1 for an unknown kind. */
FMG_HD int fmg_page(int kind, uint64_t seed, uint64_t page, uint8_t *out)
{
    if (kind == 0) fmg_logtext_page(seed, page, out);
    else if (kind == 1) fmg_json_page(seed, page, out);
    else if (kind == 2) fmg_mix_page(seed, page, out);
    else return -1;
    return 0;
}

#endif
