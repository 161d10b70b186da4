// hc_study.cpp -- CPU study for the chain parse of levels 2..4 (development aid, not shipped, not a test):
// what ratio does a GPU-shaped parse reach with (a) an exact hash chain over a sliding 64 KiB history,
// (b) a search at EVERY position and (c) a cost-optimal choice per slice (dynamic programming over the
// slice's positions), against the lazy parse the kernel had in round 1?
//   g++ -O2 -o /tmp/hc_study tools/hc_study.cpp && /tmp/hc_study file [block]
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static inline uint32_t rd4(const uint8_t *d, int p) { uint32_t v; memcpy(&v, d + p, 4); return v; }
static inline int ext(int v) { return v < 15 ? 0 : 1 + (v - 15) / 255; }

struct M { int len, off; };

int main(int argc, char **argv)
{
    FILE *f = fopen(argv[1], "rb");
    const int blk = argc > 2 ? atoi(argv[2]) : 0;
    std::vector<uint8_t> buf(4 << 20);
    fseek(f, (long)blk * (4 << 20), SEEK_SET);
    const int n = (int)fread(buf.data(), 1, buf.size(), f);
    const uint8_t *d = buf.data();
    const int mf_limit = n - 12, match_limit = n - 5;
    const int CREDIT = getenv("CREDIT") ? atoi(getenv("CREDIT")) : 6;
    const int SWAPCAP = getenv("SWAPCAP") ? atoi(getenv("SWAPCAP")) : 1 << 30;
    std::vector<int> DEPTHS; { const char *e = getenv("DEPTHS"); if (!e) e = "16,64"; for (const char *q = e; *q;) { DEPTHS.push_back(atoi(q)); while (*q && *q != ',') q++; if (*q) q++; } }
    const int TRUNC = getenv("TRUNC") ? atoi(getenv("TRUNC")) : 1 << 30;
    const int NICE = getenv("NICE") ? atoi(getenv("NICE")) : 1 << 30;
    const int HLEN = getenv("HLEN") ? atoi(getenv("HLEN")) : 4;
    const int SWAPMIN = getenv("SWAPMIN") ? atoi(getenv("SWAPMIN")) : 4;
    const int SWAP = getenv("SWAP") ? atoi(getenv("SWAP")) : 1;
    for (int HB : {14}) {
        std::vector<uint16_t> prev(n, 0);
        {
            std::vector<int> head(1 << HB, -1);
            for (int p = 0; p + 4 <= n; p++) {
                uint32_t h;
                if (HLEN == 4) h = (rd4(d, p) * 2654435761u) >> (32 - HB);
                else { uint64_t v8 = 0; memcpy(&v8, d + p, p + 8 <= n ? 8 : n - p); h = (uint32_t)(((v8 << (64 - 8 * HLEN)) * 0x9E3779B185EBCA87ull) >> (64 - HB)); }
                const int q = head[h];
                prev[p] = (q >= 0 && p - q <= 65535) ? (uint16_t)(p - q) : 0;
                head[h] = p;
            }
        }
        for (int depth : DEPTHS) {
            // longest match at every position
            std::vector<M> best(n, M{0, 0});
            double hops = 0, swaps = 0, swapev = 0;
            for (int p = 0; p <= mf_limit; p++) {
                const uint32_t v = rd4(d, p);
                const int maxlen = match_limit - p;
                int c = p, bl = 0, bo = 0, kpos = 0;
                for (int k = 0; k < depth; k++) {
                    const int dl = prev[c + kpos];
                    if (!dl) break;
                    c -= dl;
                    if (c < 0 || p - c > 65535) break;
                    hops++;
                    if (rd4(d, c) != v) continue;
                    if (bl >= 4 && d[p + bl] != d[c + bl]) continue;
                    int len = 4;
                    while (len < maxlen && d[p + len] == d[c + len]) len++;
                    if (len > bl) {
                        bl = len; bo = p - c;
                        if (len >= NICE) break;
                        if (SWAP && len >= SWAPMIN && c + len <= p) { swapev++;
                            // chain swap: a longer match must also continue every 4-byte window of this one, so
                            // follow the chain of the window whose previous occurrence lies furthest back
                            int far = 1, kb = 0;
                            for (int j = 0; j <= len - 4 && j < SWAPCAP; j++) { swaps++; if (prev[c + j] > far) { far = prev[c + j]; kb = j; } }
                            if (far > 1) kpos = kb;
                        }
                    }
                }
                best[p] = M{bl, bo};
            }
            for (int S : {32, 64, 128}) {
                for (int mode : {0, 1, 2}) {     // 0 lazy (round-1 kernel), 1 DP per slice + one restart round, 2 DP + fixpoint
                    if (S == (1 << 22) && mode == 2) continue;
                    // per slice: sequences from a fresh start, then stitched like the kernel does
                    long cost = 0; int anchor = 0, cov = 0, nseq = 0;
                    std::vector<int> nxt(n + 1, 0);   // DP: nxt[i] = 0 literal, else match length to take at i
                    if (mode) {
                        std::vector<int> dp(S == (1 << 22) ? n + 2 : S + 2);
                        for (int ss = 0; ss < n; ss += S) {
                            const int se = std::min(ss + S, n);
                            // dp over [ss, se); positions >= se cost 0
                            std::vector<int> &c = dp;
                            auto at = [&](int i) { return i >= se ? -(i - se) * CREDIT : c[i - ss]; };
                            for (int i = se - 1; i >= ss; i--) {
                                int bc = 16 + at(i + 1), bn = 0;
                                const int L = i <= mf_limit ? best[i].len : 0;
                                if (L >= 4) {
                                    // full length, and every shorter length that still ends inside the slice
                                    { const int cc = 16 * (3 + ext(L - 4)) + at(i + L); if (cc <= bc) { bc = cc; bn = L; } }
                                    for (int l = std::min(L - 1, se - i), l0 = l; l >= 4 && l > l0 - TRUNC; l--) {
                                        const int cc = 16 * (3 + ext(l - 4)) + at(i + l);
                                        if (cc < bc) { bc = cc; bn = l; }
                                    }
                                }
                                c[i - ss] = bc; nxt[i] = bn;
                            }
                        }
                    }
                    int start_next = 0;
                    for (int ss = 0; ss < n; ss += S) {
                        const int se = std::min(ss + S, n);
                        int p = mode ? std::max(ss, start_next) : ss;
                        if (mode == 0) {
                            int a = ss;
                            while (p < se && p <= mf_limit) {
                                M m = best[p];
                                if (m.len < 4) { p++; continue; }
                                if (p + 1 <= mf_limit && best[p + 1].len > m.len) { p++; continue; }
                                int st = p, len = m.len, src = p - m.off;
                                while (st > a && src > 0 && d[st - 1] == d[src - 1]) { st--; src--; len++; }
                                // stitch: trim against cov
                                int e = st + len;
                                if (st < cov) { len = e - cov; st = cov; }
                                if (len >= 4 && st <= mf_limit) { cost += 3 + ext(len - 4) + (st - anchor) + ext(st - anchor); anchor = e; nseq++; cov = std::max(cov, e); }
                                p = e; a = p;
                            }
                        } else {
                            while (p < se) {
                                if (!nxt[p]) { p++; continue; }
                                const int len = nxt[p], e = p + len;
                                cost += 3 + ext(len - 4) + (p - anchor) + ext(p - anchor); anchor = e; nseq++;
                                p = e;
                            }
                            start_next = p;
                        }
                    }
                    if (mode == 2) {
                        // the kernel's way: every slice walks from its own start, a max-scan gives what earlier slices
                        // cover, every slice walks again from there, and what still overlaps is trimmed
                        cost = 0; anchor = 0; nseq = 0;
                        const int NS = (n + S - 1) / S;
                        std::vector<int> end1(NS, 0);
                        auto walk = [&](int p, int se, std::vector<std::pair<int,int>> *out) {
                            int e = 0;
                            while (p < se) { if (!nxt[p]) { p++; continue; } if (out) out->push_back({p, nxt[p]}); p += nxt[p]; e = p; }
                            return e;
                        };
                        for (int t = 0; t < NS; t++) end1[t] = walk(t * S, std::min(t * S + S, n), nullptr);
                        int cv = 0, cov2 = 0;
                        for (int t = 0; t < NS; t++) {
                            const int ss = t * S, se = std::min(ss + S, n);
                            std::vector<std::pair<int,int>> sq;
                            walk(std::max(ss, cv), se, &sq);
                            cv = std::max(cv, end1[t]);
                            for (auto &q : sq) {
                                int st = q.first, len = q.second; const int e = st + len;
                                if (st < cov2) { len = e - cov2; st = cov2; }
                                if (len < 4 || st > mf_limit) continue;
                                cost += 3 + ext(len - 4) + (st - anchor) + ext(st - anchor); anchor = e; nseq++; cov2 = std::max(cov2, e);
                            }
                        }
                    }
                    cost += 1 + (n - anchor) + ext(n - anchor);
                    printf("HB %d depth %3d slice %7d mode %d: ratio %.4f  nseq %d  hops/pos %.1f swapreads/pos %.1f swap events/pos %.2f\n", HB, depth, S, mode, (double)n / cost, nseq, hops / n, swaps / n, swapev / n);
                }
            }
        }
    }
    return 0;
}
