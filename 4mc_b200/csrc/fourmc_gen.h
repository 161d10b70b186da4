/*
 * fourmc_gen.h -- deterministic synthetic inputs for the BASELINE.json configs.
 *
 * One source for host (gcc/g++) and device (nvcc): every 4 KiB page of the input is a pure
 * function of (seed, global page index), integer arithmetic only, so any 4 MiB block can be
 * regenerated on the CPU and on the GPU bit-identically.  A page is log lines concatenated and
 * hard-truncated at 4096 bytes.
 *
 *   log-text line:  "<epoch> host-%03d svc[%d] <LEVEL> req=%08x path=/api/v1/<w>/<w> latency=%dms msg=\"<3-12 w>\"\n"
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

#endif
