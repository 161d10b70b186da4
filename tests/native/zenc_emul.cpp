// zenc_emul.cpp -- TEST-ONLY host build of the product's zstd block encoder
// (4mc_b200/csrc/zstd_encode.h).  The encoder source is written as CTA phases; here a phase is a
// loop over thread ids, so everything except the GPU's barriers and atomics is exercised on a
// machine without a GPU and checked against the reference's ZSTD_decompress.
// The LZ sequences come from a plain greedy matcher (the GPU takes them from the region kernel of
// lz4_encode.cuh); it is not part of the product.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../4mc_b200/csrc/zstd_encode.h"

namespace {

struct EmuExec {
    template <class F> void phase(F f) { for (int t = 0; t < fmz::ZE_THREADS; t++) f(t); }
    void excl_scan(uint32_t *arr, uint32_t *, uint32_t *total)
    {
        uint32_t run = 0;
        for (int t = 0; t < fmz::ZE_THREADS; t++) { const uint32_t v = arr[t]; arr[t] = run; run += v; }
        *total = run;
    }
    void add32(uint32_t *p, uint32_t v) { *p += v; }
    void max32(uint32_t *p, uint32_t v) { if (v > *p) *p = v; }
};

inline uint32_t rd4(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }

// Greedy matchers over one region (test-only stand-ins for the region kernel):
//   mode 0  most recent occurrence of a 4-byte hash
//   mode 1  FIRST occurrence (what the GPU match finder indexes)
//   mode 2  first occurrence, but the previous match's offset is tried first (repeat offset)
void find_sequences(const uint8_t *d, int n, int mm, int mode, std::vector<uint16_t> &ll, std::vector<uint16_t> &ml,
                    std::vector<uint16_t> &off, std::vector<uint8_t> &lits)
{
    std::vector<int> table(1 << 14, -1);
    if (mode >= 1)
        for (int q = n - 4; q >= 0; q--) table[(rd4(d + q) * 2654435761u) >> 18] = q;
    int p = 0, anchor = 0, rep = 0;
    while (p + 4 <= n) {
        const uint32_t v = rd4(d + p);
        const uint32_t h = (v * 2654435761u) >> 18;
        int c = table[h];
        if (mode == 0) table[h] = p;
        int best = 0, bo = 0;
        if (c >= 0 && c < p && rd4(d + c) == v) {
            int len = 4;
            while (p + len < n && d[p + len] == d[c + len]) len++;
            best = len; bo = p - c;
        }
        if (mode == 2 && rep > 0 && p - rep >= 0 && p > anchor && rd4(d + p - rep) == v) {
            int len = 4;
            while (p + len < n && d[p + len] == d[p - rep + len]) len++;
            if (len + 3 >= best) { best = len; bo = rep; }
        }
        if (best >= (bo == rep && mode == 2 ? 4 : mm) && best <= 65535 && p - anchor <= 65535) {
            ll.push_back((uint16_t)(p - anchor)); ml.push_back((uint16_t)best); off.push_back((uint16_t)bo);
            lits.insert(lits.end(), d + anchor, d + p);
            p += best; anchor = p; rep = bo;
            continue;
        }
        p++;
    }
    lits.insert(lits.end(), d + anchor, d + n);
}

}  // namespace

// One zstd frame for src[0..n), n <= 4 MiB, assembled the way the GPU kernels assemble it.
// Returns the frame size, or -1 when it does not fit.
// min_match: low 4 bits = minimum match length, bits 4.. = matcher mode (see find_sequences)
extern "C" long long zenc_emul_compress(uint8_t *dst, long long cap, const uint8_t *src, long long n, int min_match)
{
    static fmz::Tables T;
    static bool init = false;
    if (!init) { fmz::make_tables(T); init = true; }
    std::vector<uint8_t> out(fmz::ZE_FRAME_HDR);
    fmz::ze_write_frame_header(out.data(), (uint32_t)n);
    const long long nreg = std::max<long long>(1, (n + fmz::ZE_REGION - 1) / fmz::ZE_REGION);
    fmz::ZShared *sh = (fmz::ZShared *)malloc(sizeof(fmz::ZShared));
    std::vector<uint32_t> slot(fmz::ZE_OUT_SLOT / 4);
    for (long long r = 0; r < nreg; r++) {
        const long long o = r * fmz::ZE_REGION;
        const int rlen = (int)std::min<long long>(fmz::ZE_REGION, n - o);
        std::vector<uint16_t> ll, ml, off;
        std::vector<uint8_t> lits;
        find_sequences(src + o, rlen, min_match & 15, min_match >> 4, ll, ml, off, lits);
        lits.resize(lits.size() + 8);
        fmz::ZRegionIn in{ll.data(), ml.data(), off.data(), lits.data(), (uint32_t)ll.size(), (uint32_t)(lits.size() - 8), (uint32_t)rlen, (uint32_t)fmz::ZE_REGION};
        fmz::ZRegionOut ro{0, 0};
        EmuExec ex;
        memset(sh, 0xA5, sizeof(*sh));                               // nothing may rely on zeroed shared memory
        std::fill(slot.begin(), slot.end(), 0xDEADBEEFu);            // ... or on a zeroed slot
        fmz::zenc_region(ex, *sh, in, slot.data(), &ro, T);
        uint8_t bh[3];
        const bool last = r == nreg - 1;
        if (ro.raw || rlen == 0) {
            fmz::ze_write_block_header(bh, last, 0, (uint32_t)rlen);
            out.insert(out.end(), bh, bh + 3);
            out.insert(out.end(), src + o, src + o + rlen);
        } else {
            fmz::ze_write_block_header(bh, last, 2, ro.bytes);
            out.insert(out.end(), bh, bh + 3);
            out.insert(out.end(), (uint8_t *)slot.data(), (uint8_t *)slot.data() + ro.bytes);
        }
    }
    free(sh);
    if ((long long)out.size() > cap) return -1;
    memcpy(dst, out.data(), out.size());
    return (long long)out.size();
}

extern "C" int zenc_emul_shared_bytes() { return (int)sizeof(fmz::ZShared); }

// ---- unit hooks ---------------------------------------------------------------------------------

// normalise `count[0..max_sym]`, write the table description, read it back with the decoder's
// parser.  Returns the description size, or a negative number on any disagreement.
extern "C" int zenc_emul_ncount_roundtrip(const uint32_t *count, int max_sym, int log)
{
    short norm[64], back[64];
    uint32_t total = 0;
    for (int s = 0; s <= max_sym; s++) total += count[s];
    fmz::fse_normalize(norm, log, count, total, max_sym);
    int sum = 0;
    for (int s = 0; s <= max_sym; s++) { if (count[s] && norm[s] < 1) return -1; if (!count[s] && norm[s]) return -2; sum += norm[s]; }
    if (sum != (1 << log)) return -3;
    uint8_t buf[256];
    const int n = fmz::fse_write_ncount(buf, norm, max_sym, log);
    int ms = 63, lg = 0;
    const int used = fmz::read_ncount(back, &ms, &lg, 9, buf, n);
    if (used != n) return -4;
    if (ms != max_sym || lg != log) return -5;
    for (int s = 0; s <= max_sym; s++) if (back[s] != norm[s]) return -6;
    return n;
}

// Huffman lengths for count[0..255]: returns the table log, or a negative number when the code is
// not a complete prefix code within the length limit.
extern "C" int zenc_emul_huffman_check(const uint32_t *count, uint8_t *nbits_out)
{
    static fmz::HufBuild hb;
    uint16_t sorted[256];
    std::vector<std::pair<uint32_t, int>> v;
    for (int s = 0; s < 256; s++) if (count[s]) v.push_back({count[s], s});
    if (v.size() < 2) return -1;
    std::sort(v.begin(), v.end());
    for (size_t i = 0; i < v.size(); i++) sorted[i] = (uint16_t)v[i].second;
    memset(nbits_out, 0, 256);
    const int log = fmz::huf_build_lengths(hb, count, sorted, (int)v.size(), nbits_out);
    uint64_t kraft = 0;
    for (int s = 0; s < 256; s++) {
        if ((count[s] != 0) != (nbits_out[s] != 0)) return -2;
        if (nbits_out[s] > fmz::ZE_HUF_MAXBITS) return -3;
        if (nbits_out[s]) kraft += 1ull << (fmz::ZE_HUF_MAXBITS - nbits_out[s]);
    }
    if (kraft != (1ull << fmz::ZE_HUF_MAXBITS)) return -4;
    return log;
}
