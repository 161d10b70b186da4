// enc_emul.cpp -- TEST-ONLY sequential emulation of the encode kernels' algorithm
// (4mc_b200/csrc/lz4_encode.cuh: region parse / stitch / emit, block size, block write).
// It mirrors the kernels step by step with "for each thread" loops so that the algorithm (slice
// stitching, region joins, end-of-block rules) can be checked against the oracle decoder on a
// machine without a GPU.  It is not part of the product and is never shipped.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {
constexpr int BLOCK = 4 << 20, REGION = 65536, THREADS = 512, SLICE = 132, HB = 14, PAD = 64;
struct Meta { uint32_t body_bytes, tail_lits, lead, nseq; };
struct Seq { int st, len, off; };
inline uint32_t rd4(const uint8_t *d, int p) { uint32_t v; memcpy(&v, d + p, 4); return v; }
inline uint32_t hsh(uint32_t v) { return (v * 2654435761u) >> (32 - HB); }
inline int ext(int v) { return v < 15 ? 0 : 1 + (v - 15) / 255; }
inline uint8_t *emit_len(uint8_t *o, int v) { while (v >= 255) { *o++ = 255; v -= 255; } *o++ = (uint8_t)v; return o; }

// depth == 0: the Fast parse (first-occurrence table).  depth > 0: the chain parse of the higher
// levels -- every position linked to the previous position with the same hash (exact links over the whole
// block, at most 65535 back: lz4_chain_kernel), `depth` candidates per search with the chain swap, a search at
// every position, a cost-optimal choice per 64-byte slice and two walks (lz4_region_kernel<., true>).
// The window is 64 KiB of look-back (earlier bytes of the block: searched, not parsed) + a 32 KiB region of new
// bytes; positions are relative to the window, new bytes are [lb, rlen).
constexpr int CHAIN_REGION = 32768, CHAIN_LOOKBACK = 65536, CHAIN_WINDOW = CHAIN_LOOKBACK + CHAIN_REGION, CHAIN_SLICE = 64,
              CHAIN_THREADS = 1024, LENCAP = 255, CREDIT = 6, SHORTER = 8, SWAPSCAN = 16;
// chain links of one block: distance to the previous position with the same hash, 0 = none within 65535
std::vector<uint16_t> chain_links(const uint8_t *blk, uint32_t blk_len)
{
    std::vector<uint16_t> link(blk_len, 0);
    std::vector<int> head(1 << HB, -1);
    for (int p = 0; p + 4 <= (int)blk_len; p++) {
        const uint32_t h = hsh(rd4(blk, p));
        const int q = head[h];
        link[p] = (q >= 0 && p - q <= 65535) ? (uint16_t)(p - q) : 0;
        head[h] = p;
    }
    return link;
}
// lz4_chain_kernel restated lane by lane: the block is cut into chunks, every chunk is linked on its own after 64 KiB of
// warm-up, in steps of 32 positions; a step reads the heads, stores its positions (equal hashes race: here the LOWEST lane
// lands, the opposite of what the kernel finally wants, so the repair path is always exercised), reads back, and repairs
// the step lane to lane when some lane lost.  Must equal chain_links() for every chunk size.
std::vector<uint16_t> chain_links_chunked(const uint8_t *blk, uint32_t blk_len, uint32_t chunk)
{
    std::vector<uint16_t> link(blk_len, 0xAAAA);
    const int last = (int)blk_len - 4;
    for (uint32_t c0u = 0; c0u < blk_len; c0u += chunk) {
        const int c0 = (int)c0u, c1 = (int)std::min<uint32_t>(c0u + chunk, blk_len), w0 = std::max(c0 - 65536, 0);
        std::vector<uint32_t> head(1 << HB, 0xffffffffu);
        for (int base = w0; base < c1; base += 32) {
            uint32_t h[32], q[32]; bool live[32];
            for (int l = 0; l < 32; l++) {
                const int p = base + l;
                live[l] = p <= last;
                h[l] = live[l] ? hsh(rd4(blk, p)) : 0x10000u + (uint32_t)l;
                q[l] = live[l] ? head[h[l]] : 0xffffffffu;
            }
            for (int l = 31; l >= 0; l--) if (live[l]) head[h[l]] = (uint32_t)(base + l);      // the lowest lane lands
            bool any_lost = false;
            for (int l = 0; l < 32; l++) if (live[l] && head[h[l]] != (uint32_t)(base + l)) any_lost = true;
            if (any_lost) {
                for (int l = 0; l < 32; l++) {
                    int below = -1, above = -1;
                    for (int m = 0; m < 32; m++) if (h[m] == h[l]) { if (m < l) below = m; if (m > l) above = m; }
                    if (below >= 0) q[l] = (uint32_t)(base + below);
                    if (live[l] && above < 0) head[h[l]] = (uint32_t)(base + l);
                }
            }
            for (int l = 0; l < 32; l++) {
                const int p = base + l;
                const uint32_t dist = (uint32_t)p - q[l];
                if (p >= c0 && p < c1) link[p] = (q[l] != 0xffffffffu && dist <= 65535u) ? (uint16_t)dist : (uint16_t)0;
            }
        }
    }
    return link;
}
void region(const uint8_t *blk, uint32_t blk_len, uint32_t r_new, int min_match, std::vector<uint8_t> &body, Meta &mt,
            int depth = 0, const uint16_t *link = nullptr)
{
    const int lb = depth > 0 ? (int)std::min<uint32_t>(CHAIN_LOOKBACK, r_new) : 0;
    const uint32_t r_off = r_new - (uint32_t)lb;
    const int rlen = lb + (int)std::min<uint32_t>(depth > 0 ? CHAIN_REGION : REGION, blk_len - r_new);
    std::vector<uint8_t> data(CHAIN_WINDOW + PAD, 0);
    memcpy(data.data(), blk + r_off, rlen);
    const int mf_limit = std::min(rlen - 1, (int)blk_len - 12 - (int)r_off);
    const int match_limit = std::min(rlen, (int)blk_len - 5 - (int)r_off);
    std::vector<uint16_t> table(1 << HB, 0xffff);
    // index: descending sweep in steps of THREADS; within a step the order is unspecified on
    // the GPU -- emulate "highest thread wins" (any order is legal)
    if (depth == 0) for (int base = ((rlen - 1) / THREADS) * THREADS; base >= 0; base -= THREADS)
        for (int t = THREADS - 1; t >= 0; t--) { int p = base + t; if (p <= rlen - 4 && (p & 1) == 0) table[hsh(rd4(data.data(), p))] = (uint16_t)p; }   // even positions only
    const int NTH = depth > 0 ? CHAIN_THREADS : THREADS;
    std::vector<std::vector<Seq>> inner(NTH);
    std::vector<Seq> last(NTH, Seq{0, 0, 0});
    const uint8_t *d = data.data();
    std::vector<uint16_t> moff;
    std::vector<uint8_t> mlen;
    if (depth > 0) {
        // search: the longest match of every new position
        const uint16_t *gp = link + r_off;
        const int nnew = rlen - lb;
        moff.assign(nnew, 0); mlen.assign(nnew, 0);
        for (int k = 0; k < nnew; k++) {
            const int q = lb + k;
            int best = 0, bo = 0;
            if (q <= mf_limit) {
                const uint32_t v = rd4(d, q);
                const int maxlen = std::min(match_limit - q, LENCAP);
                int c = q, kpos = 0;
                for (int a = 0; a < depth; a++) {
                    const int dl = gp[c + kpos];
                    if (!dl) break;
                    c -= dl;
                    if (c < 0 || q - c > 65535) break;
                    if (rd4(d, c) != v) continue;
                    if (best >= 4 && d[q + best] != d[c + best]) continue;
                    int len = 4;
                    while (len < maxlen && d[q + len] == d[c + len]) len++;
                    len = std::min(len, maxlen);
                    if (len > best) {
                        best = len; bo = q - c;
                        if (len >= maxlen) break;
                        if (c + len <= q) {
                            int far = 1, kb = 0;
                            const int ns = std::min(len - 3, SWAPSCAN);
                            for (int j = 0; j < ns; j++) if (gp[c + j] > far) { far = gp[c + j]; kb = j; }
                            if (far > 1) kpos = kb;
                        }
                    }
                }
            }
            moff[k] = (uint16_t)bo; mlen[k] = (uint8_t)(best >= min_match ? best : 0);
        }
        // parse: cost-optimal choices per slice, back to front; walk 1
        std::vector<int> end1(NTH, 0);
        for (int t = 0; t < NTH; t++) {
            const int ss = lb + t * CHAIN_SLICE;
            if (ss >= rlen) continue;
            const int se = std::min(ss + CHAIN_SLICE, rlen);
            int cost[CHAIN_SLICE];
            auto at = [&](int i) { return i >= se ? -(i - se) * CREDIT : cost[i - ss]; };
            for (int i = se - 1; i >= ss; i--) {
                int bc = 16 + at(i + 1), bn = 0;
                const int L = mlen[i - lb];
                if (L) {
                    const int cf = 16 * (3 + ext(L - 4)) + at(i + L);
                    if (cf <= bc) { bc = cf; bn = L; }
                    const int l0 = std::min(L - 1, se - i);
                    for (int l = l0; l >= min_match && l > l0 - SHORTER; l--) {
                        const int cc = 16 * (3 + ext(l - 4)) + at(i + l);
                        if (cc < bc) { bc = cc; bn = l; }
                    }
                }
                cost[i - ss] = bc; mlen[i - lb] = (uint8_t)bn;
            }
            for (int p = ss; p < se;) { const int l = mlen[p - lb]; if (l) { p += l; end1[t] = p; } else p++; }
        }
        // walk 2 from what earlier slices leave
        int cov1 = 0;
        for (int t = 0; t < NTH; t++) {
            const int ss = lb + t * CHAIN_SLICE;
            if (ss < rlen) {
                const int se = std::min(ss + CHAIN_SLICE, rlen);
                for (int p = std::max(ss, cov1); p < se;) {
                    int len = mlen[p - lb];
                    if (!len) { p++; continue; }
                    const int off = moff[p - lb];
                    if (len == LENCAP) { const int maxlen = match_limit - p; while (len < maxlen && d[p + len] == d[p + len - off]) len++; }
                    if (last[t].len) inner[t].push_back(last[t]);
                    last[t] = Seq{p, len, off};
                    p += len;
                }
            }
            cov1 = std::max(cov1, end1[t]);
        }
    }
    for (int t = 0; t < NTH; t++) {
        const int ss = depth > 0 ? lb + t * CHAIN_SLICE : t * SLICE;
        if (ss >= rlen) continue;
        const int se = std::min(ss + (depth > 0 ? CHAIN_SLICE : SLICE), rlen);
        int p = ss, anchor = ss;
        if (depth > 0) continue;                                 // parsed above
        while (p < se && p <= mf_limit) {
            const uint32_t v = rd4(d, p);
            const int c = table[hsh(v)];
            if (c < p && rd4(d, c) == v) {
                int len = 4; const int maxlen = match_limit - p;
                while (len < maxlen && d[p + len] == d[c + len]) len++;
                len = std::min(len, maxlen);
                int st = p, m = c;
                while (st > anchor && m > 0 && d[st - 1] == d[m - 1]) { st--; m--; len++; }
                if (len >= min_match) {
                    if (last[t].len) inner[t].push_back(last[t]);
                    last[t] = Seq{st, len, st - m};
                    p = st + len; anchor = p; continue;
                }
            }
            p++;
        }
    }
    // stitch 1
    std::vector<int> surv_end(NTH, 0);
    int cov = 0;
    for (int t = 0; t < NTH; t++) {
        const int my_end = last[t].len ? last[t].st + last[t].len : 0;
        std::vector<Seq> keep;
        for (auto s : inner[t]) {
            const int end = s.st + s.len;
            if (s.st < cov) { s.len = end - cov; s.st = cov; }
            if (s.len < 4 || s.st > mf_limit) continue;
            keep.push_back(s); surv_end[t] = end;
        }
        inner[t] = keep;
        if (last[t].len) {
            const int end = last[t].st + last[t].len;
            if (last[t].st < cov) { last[t].len = end - cov; last[t].st = cov; }
            if (last[t].len < 4 || last[t].st > mf_limit) last[t].len = 0; else surv_end[t] = end;
        }
        cov = std::max(cov, my_end);
    }
    // stitch 2 + emit
    body.clear();
    int anchor = lb; uint32_t nseq = 0; mt.lead = 0;
    for (int t = 0; t < NTH; t++) {
        std::vector<Seq> all = inner[t];
        if (last[t].len) all.push_back(last[t]);
        int a = anchor;
        for (auto &s : all) {
            const int lit = s.st - a, ml = s.len - 4;
            if (nseq == 0) mt.lead = lit;
            size_t o0 = body.size();
            body.resize(o0 + 1 + ext(lit) + lit + 2 + ext(ml));
            uint8_t *o = body.data() + o0;
            *o++ = (uint8_t)((std::min(lit, 15) << 4) | std::min(ml, 15));
            if (lit >= 15) o = emit_len(o, lit - 15);
            memcpy(o, d + a, lit); o += lit;
            *o++ = (uint8_t)(s.off & 0xff); *o++ = (uint8_t)(s.off >> 8);
            if (ml >= 15) o = emit_len(o, ml - 15);
            a = s.st + s.len; nseq++;
        }
        anchor = std::max(anchor, surv_end[t]);
    }
    mt.body_bytes = (uint32_t)body.size(); mt.tail_lits = (uint32_t)(rlen - anchor); mt.nseq = nseq;
}
}  // namespace

// Compresses one block (n <= 4 MiB) the way the kernels do.  Returns the LZ4 payload size written
// to dst (capacity must be >= n + n/255 + 64), never "stored".
static int emul_block(const uint8_t *src, int n, uint8_t *dst, int min_match, int depth);
// 0 when the chunked, step-wise links (lz4_chain_kernel) equal the sequential ones; else 1 + the first differing position
extern "C" long long enc_emul_chain_links_check(const uint8_t *src, int n, int chunk)
{
    const std::vector<uint16_t> a = chain_links(src, (uint32_t)n), b = chain_links_chunked(src, (uint32_t)n, (uint32_t)chunk);
    for (int i = 0; i < n; i++) if (a[i] != b[i]) return 1 + i;
    return 0;
}
extern "C" int enc_emul_block(const uint8_t *src, int n, uint8_t *dst, int min_match) { return emul_block(src, n, dst, min_match, 0); }
// the chain parse of levels 2..4 (capi.cu level_chain_depth / level_lazy)
extern "C" int enc_emul_block_chain(const uint8_t *src, int n, uint8_t *dst, int min_match, int depth)
{
    return emul_block(src, n, dst, min_match, depth);
}
static int emul_block(const uint8_t *src, int n, uint8_t *dst, int min_match, int depth)
{
    std::vector<uint16_t> link;
    if (depth > 0) link = chain_links(src, (uint32_t)n);
    const int REG = depth > 0 ? CHAIN_REGION : REGION, RPB = BLOCK / REG;
    std::vector<Meta> meta(RPB, Meta{0, 0, 0, 0});
    std::vector<std::vector<uint8_t>> bodies(RPB);
    for (int r = 0; r < RPB; r++) if ((uint32_t)r * REG < (uint32_t)n) region(src, (uint32_t)n, (uint32_t)r * REG, min_match, bodies[r], meta[r], depth, link.data());
    uint8_t *o = dst; uint32_t carry = 0;
    for (int r = 0; r < RPB; r++) {
        const Meta &x = meta[r];
        if (x.nseq == 0) { carry += x.tail_lits; continue; }
        const int old_hdr = 1 + ext((int)x.lead), lit = (int)(x.lead + carry);
        *o++ = (uint8_t)((std::min(lit, 15) << 4) | (bodies[r][0] & 15));
        if (lit >= 15) o = emit_len(o, lit - 15);
        memcpy(o, src + (size_t)r * REG - carry, carry); o += carry;
        memcpy(o, bodies[r].data() + old_hdr, x.body_bytes - old_hdr); o += x.body_bytes - old_hdr;
        carry = x.tail_lits;
    }
    *o++ = (uint8_t)(std::min((int)carry, 15) << 4);
    if (carry >= 15) o = emit_len(o, (int)carry - 15);
    memcpy(o, src + n - carry, carry); o += carry;
    return (int)(o - dst);
}
