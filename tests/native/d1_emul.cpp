// d1_emul.cpp -- TEST-ONLY sequential emulation of lz4_parse_kernel's bulk + tail phases
// (4mc_b200/csrc/lz4_decode.cuh), lane by lane, so that the warp-parallel parse can be checked
// against the plain serial walk (lz4_parse.h) on a machine without a GPU.  Not part of the product.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../4mc_b200/csrc/lz4_parse.h"

namespace {
constexpr int HALF = 2048, RING = 4096, SUB = 64, SPECIAL = 0x8000, NONE = 0xffff;

struct Stream {                 // the aligned stream: q = ip + d; bytes outside the payload are filler
    const uint8_t *src; int csize, d; uint8_t filler;
    unsigned at_q(int q) const { int ip = q - d; return (ip >= 0 && ip < csize) ? src[ip] : filler; }
    unsigned operator()(int ip) const { return (ip >= 0 && ip < csize) ? src[ip] : filler; }
    unsigned near(int ip) const { return (*this)(ip); }
};
struct SeqDec { int lit, ml, off, next; bool clean; };
SeqDec decode_slow(const Stream &rd, int ip, int limit)
{
    SeqDec r{0, 0, 0, ip, false};
    if (ip >= limit) return r;
    const unsigned tok = rd(ip++);
    int lit = (int)(tok >> 4);
    if (lit == 15) { unsigned b; do { if (ip >= limit) return r; b = rd(ip++); lit += (int)b; } while (b == 255); }
    if (lit > limit - ip) return r;
    ip += lit;
    if (ip + 2 > limit) return r;
    r.off = (int)rd(ip) | ((int)rd(ip + 1) << 8); ip += 2;
    int ml = (int)(tok & 15);
    if (ml == 15) { unsigned b; do { if (ip >= limit) return r; b = rd(ip++); ml += (int)b; } while (b == 255); }
    r.lit = lit; r.ml = ml + 4; r.next = ip; r.clean = true;
    return r;
}
struct Sink {
    std::vector<uint32_t> *map; std::vector<int64_t> *cop; int d; uint32_t cur_chunk;
    void token(int ip, int op)
    {
        const uint32_t q = (uint32_t)(ip + d);
        (*map)[q >> 5] |= 1u << (q & 31);
        if ((q >> 7) != cur_chunk) { (*cop)[q >> 7] = op; cur_chunk = q >> 7; }
    }
};
}  // namespace

// Returns the parse result; fills map (1 bit per aligned position) and cop (per 128-byte chunk, -1 = unset).
extern "C" int d1_emul(const uint8_t *src, int csize, int cap, int d, uint32_t *map_out, int64_t *cop_out, int n_words, int n_chunks)
{
    std::vector<uint32_t> map(n_words, 0);
    std::vector<int64_t> cop(n_chunks, -1);
    Stream rd{src, csize, d, 0xF7};
    fm::ParseState st;
    fm::lz4_parse_init(st, src == nullptr, csize, cap, (src && csize > 0) ? src[0] : 0u);
    Sink sink{&map, &cop, d, 0xffffffffu};
    const int oend = cap;
    if (!st.status) {
        const int clean_ip = csize - 32, clean_op = oend - 64;
        int e_q = d, e_op = 0;
        bool bulk = clean_ip > 0 && clean_op > 0;
        std::vector<int> A(HALF), Bv(HALF);
        while (bulk) {
            const int k = e_q >> 11, base_q = k * HALF;
            for (int lane = 0; lane < 32; lane++) {                // phase B (per sub-chunk, local exits)
                for (int jj = SUB - 1; jj >= 0; jj--) {
                    const unsigned tok = rd.at_q(base_q + lane * SUB + jj);
                    int lit = (int)(tok >> 4), ml = (int)(tok & 15), n = jj + 3;
                    bool special = false;
                    if (lit == 15) { const unsigned x = rd.at_q(base_q + lane * SUB + jj + 1); special = x == 255; lit += (int)x; n++; }
                    n += lit;
                    if (ml == 15 && !special) { const unsigned x = rd.at_q(base_q + lane * SUB + n); special = x == 255; ml += (int)x; n++; }
                    int ex, os = lit + ml + 4;
                    if (special) { ex = jj | 0x8000; os = 0; }
                    else if (n < SUB) { ex = A[lane * SUB + n]; os += Bv[lane * SUB + n]; }
                    else ex = n;
                    A[lane * SUB + jj] = ex; Bv[lane * SUB + jj] = os;
                }
            }
            int entry[32], eop[32];
            for (int i = 0; i < 32; i++) entry[i] = NONE;
            int e = e_q - base_q, op = e_op;                       // phase C
            int last_sc = -1;
            while (e < HALF) {
                const int sc = e >> 6, jj = e & (SUB - 1);
                if (sc != last_sc) { entry[sc] = e; eop[sc] = op; last_sc = sc; }
                const int x = A[sc * SUB + jj]; op += Bv[sc * SUB + jj];
                if (x & 0x8000) {
                    const SeqDec sd = decode_slow(rd, base_q + sc * SUB + (x & (SUB - 1)) - d, clean_ip);
                    if (!sd.clean) break;
                    op += sd.lit + sd.ml; e = sd.next + d - base_q;
                } else e = sc * SUB + x;
            }
            int pu = 0x7fffffff, pv = 0x7fffffff, pu_op = 0, pv_next = 0;   // phase D
            int first_op[32];
            for (int lane = 0; lane < 32; lane++) {
                first_op[lane] = -1;
                int p = entry[lane];
                if (p == NONE) continue;
                int o = eop[lane]; const int s1 = (lane + 1) * SUB;
                while (p < s1) {
                    const int ip = base_q + p - d; int lit, mlen, off, next;
                    const unsigned tok = rd.at_q(base_q + p);
                    bool slow = false;
                    {
                        int qo = base_q + p + 1;
                        lit = (int)(tok >> 4); mlen = (int)(tok & 15) + 4;
                        if (lit == 15) { const unsigned x = rd.at_q(qo); slow = x == 255; lit += (int)x; qo++; }
                        qo += lit;
                        off = (int)rd.at_q(qo) | ((int)rd.at_q(qo + 1) << 8);
                        qo += 2;
                        if (mlen == 19 && !slow) { const unsigned x = rd.at_q(qo); slow = x == 255; mlen += (int)x; qo++; }
                        next = qo - d;
                    }
                    if (slow) {
                        const SeqDec sd = decode_slow(rd, ip, clean_ip);
                        if (!sd.clean) { if (p < pu) { pu = p; pu_op = o; } break; }
                        lit = sd.lit; mlen = sd.ml; off = sd.off; next = sd.next;
                    }
                    if (next > clean_ip || o + lit + mlen >= clean_op) { if (p < pu) { pu = p; pu_op = o; } break; }
                    if (off > o + lit) { if (p < pv) { pv = p; pv_next = next; } break; }
                    if (first_op[lane] < 0) first_op[lane] = o;
                    const uint32_t q = (uint32_t)(base_q + p);
                    map[q >> 5] |= 1u << (q & 31);
                    o += lit + mlen; p = next + d - base_q;
                }
            }
            if (pv < pu) { st.status = 1; st.result = -pv_next - 1; break; }
            for (int c = 0; c < 16; c++) {
                const int f = first_op[2 * c] >= 0 ? first_op[2 * c] : first_op[2 * c + 1];
                if (f >= 0) cop[(base_q >> 7) + c] = f;
            }
            if (pu != 0x7fffffff) {
                st.ip = base_q + pu - d; st.op = pu_op;
                const int su = pu >> 6;
                const bool chunk_has = (su & 1) ? (first_op[su & ~1] >= 0 || first_op[su] >= 0) : (first_op[su] >= 0);
                sink.cur_chunk = chunk_has ? (uint32_t)(base_q + pu) >> 7 : 0xffffffffu;
                break;
            }
            e_q = base_q + e; e_op = op;
            if (e_q - d >= csize) { st.status = 1; st.result = -(e_q - d) - 1; break; }
        }
    }
    if (!st.status) {
        int stop = st.ip + 1;                       // resume in small steps to exercise resumability
        while (!st.status) { fm::lz4_parse_run(st, rd, sink, csize, cap, stop); stop += 777; }
    }
    memcpy(map_out, map.data(), n_words * 4);
    memcpy(cop_out, cop.data(), n_chunks * 8);
    return st.result;
}

// Plain serial walk with the same outputs, for comparison.
extern "C" int d1_serial(const uint8_t *src, int csize, int cap, int d, uint32_t *map_out, int64_t *cop_out, int n_words, int n_chunks)
{
    std::vector<uint32_t> map(n_words, 0);
    std::vector<int64_t> cop(n_chunks, -1);
    Sink sink{&map, &cop, d, 0xffffffffu};
    const int r = fm::lz4_parse_block(src, csize, cap, sink);
    memcpy(map_out, map.data(), n_words * 4);
    memcpy(cop_out, cop.data(), n_chunks * 8);
    return r;
}
