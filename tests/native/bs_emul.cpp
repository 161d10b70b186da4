// bs_emul.cpp -- test helper: the block-stream framing (4mc_b200/csrc/blockstream.h, the host code the C-ABI uses for
// the raw Lz4Codec / ZstdCodec streams) bound to a CPU codec handed in as function pointers (the oracle's LZ4, or the
// reference's ZSTD_compress with the oracle's zstd decoder), so that the writer / reader rules are checked without a
// GPU and GPU-made streams can be cross-decoded.
#include <climits>

#include "../../4mc_b200/csrc/blockstream.h"

namespace {
struct Fns { int kind; void *c, *d; };          // kind 0: LZ4 (oracle signatures), 1: zstd (ZSTD_compress, fmo_zstd_decompress)

long long comp(void *u, int level, const uint8_t *s, uint32_t n, uint8_t *d, size_t cap)
{
    const Fns *f = (const Fns *)u;
    if (f->kind == 0) {
        typedef int (*fn)(const char *, char *, int, int);
        return ((fn)f->c)((const char *)s, (char *)d, (int)n, (int)std::min<size_t>(cap, INT_MAX));
    }
    typedef size_t (*fn)(void *, size_t, const void *, size_t, int);
    static const int lv[5] = {1, 1, 3, 6, 12};              // ZstdCodec / Medium / High / Ultra
    const size_t r = ((fn)f->c)(d, cap, s, n, lv[level < 1 ? 1 : level > 4 ? 4 : level]);
    return r > cap ? -1 : (long long)r;                    // ZSTD_isError values are huge
}
long long decomp(void *u, const uint8_t *s, uint32_t c, uint8_t *d, uint32_t cap)
{
    const Fns *f = (const Fns *)u;
    if (f->kind == 0) {
        typedef int (*fn)(const char *, char *, int, int);
        return ((fn)f->d)((const char *)s, (char *)d, (int)c, (int)cap);
    }
    typedef long long (*fn)(uint8_t *, long long, const uint8_t *, long long);
    return ((fn)f->d)(d, cap, s, c);
}
fbs::Codec codec(Fns *f)
{
    const uint32_t n = fbs::BUFFER;
    const uint32_t bound = f->kind == 0 ? n + n / 255 + 16 : n + (n >> 8);     // lz4.h:212, zstd.h:204 (n >= 128 KiB)
    return fbs::Codec{f, bound, comp, decomp};
}
}  // namespace

extern "C" {

unsigned bs_emul_max_input(int kind) { Fns f{kind, nullptr, nullptr}; return fbs::max_input(codec(&f)); }

size_t bs_emul_bound(int kind, size_t n, size_t write_size) { Fns f{kind, nullptr, nullptr}; return fbs::bound(codec(&f), n, write_size); }

long long bs_emul_compress(int kind, void *cfn, int level, const uint8_t *in, size_t n, size_t write_size, uint8_t *out, size_t cap)
{
    Fns f{kind, cfn, nullptr};
    return fbs::compress(codec(&f), level, in, n, write_size, out, cap);
}

long long bs_emul_decompress(int kind, void *dfn, const uint8_t *in, size_t n, uint8_t *out, size_t cap)
{
    Fns f{kind, nullptr, dfn};
    return fbs::decompress(codec(&f), in, n, out, cap);
}

// number of blocks; raw lengths to raws[] (up to cap entries); *trailing_zero as in plan_blocks
long long bs_emul_plan(int kind, size_t n, size_t write_size, unsigned *raws, size_t cap, int *trailing_zero)
{
    Fns f{kind, nullptr, nullptr};
    std::vector<fbs::Block> b;
    bool tz = false;
    fbs::plan_blocks(codec(&f), n, write_size, b, &tz);
    for (size_t i = 0; i < b.size() && i < cap; ++i) raws[i] = b[i].raw;
    *trailing_zero = tz ? 1 : 0;
    return (long long)b.size();
}

// chunks predicted from the framing alone: count, or -1 when the stream is not the reference writer's
long long bs_emul_predict(int kind, const uint8_t *in, size_t n, unsigned *usizes, size_t cap, size_t *total)
{
    Fns f{kind, nullptr, nullptr};
    std::vector<fbs::Chunk> c;
    if (!fbs::predict_chunks(codec(&f), in, n, c, total)) return -1;
    for (size_t i = 0; i < c.size() && i < cap; ++i) usizes[i] = c[i].usize;
    return (long long)c.size();
}

}
