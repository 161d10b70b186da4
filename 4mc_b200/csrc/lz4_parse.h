// lz4_parse.h -- sequential validation/parse of one LZ4 block (host + device).
//
// Decoding a 4 MiB LZ4 block on the GPU is split in two (DESIGN.md "LZ4 decode"):
//   D1  this file: ONE lane walks the token chain of a block, reproducing every accept/reject
//       decision of the reference decoder and emitting where each sequence starts;
//   D2  lz4_decode.cuh: warps copy literals and matches of up to 32 sequences at a time.
// All of LZ4_decompress_safe's error conditions depend only on positions and lengths, never on
// the decoded bytes (native/lz4/lz4.c:1936-2339), so D1 alone decides the call's return value,
// including the exact negative value -(ip)-1 (:2337).
//
// The reference runs a "fast" loop while >= 64 bytes of output space remain (:1996-2115) and a
// "safe" loop afterwards (:2120-2328); their checks differ slightly on malformed input (the
// shortcuts at :2014, :2132 skip the end-of-input test of :2186), so both are kept, as two
// states over integer positions.  The same structure is restated for the CPU in
// oracle/fourmc_oracle.c (test infrastructure); this is the product's own copy.
//
// The walk is a dependent chain (token -> lengths -> next token position), so its speed is the
// latency of one iteration.  Two things keep that short on the GPU: the parser is resumable at
// sequence boundaries (lz4_parse_run stops at `ip_stop`), which lets the caller stream the
// compressed bytes through shared memory ahead of it; and in the fast loop the "offset points
// before the start of the output" test (:2065) is carried into the next iteration (`pend`), so
// the two offset bytes are never waited for -- the error value is unchanged because that test's
// failure position is exactly the next token's position.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define FM_HD __host__ __device__ __forceinline__
#define FM_HDM __host__ __device__ __forceinline__
#else
#define FM_HD static inline
#define FM_HDM inline
#endif

namespace fm {

struct ParseState {
    int ip, op;
    int status;        // 0 running, 1 finished (result valid)
    int result;
    bool fast;
    bool pend;         // a deferred offset test is outstanding
    int pend_off, pend_op;
};

// Reader interface: unsigned operator()(int ip) const  -- byte ip of the compressed block;
//                   unsigned near(int ip) const        -- same, for ip < (current token) + 32:
//                   a staged reader may skip its window test there (see ip_stop below).
// Sink interface:   void token(int token_pos, int op_before_literals)

// variable-length field (lz4.c:1903-1928).  Returns false on error.
template <class Reader>
FM_HD bool lz4_rvl(const Reader &rd, int &ip, int ilimit, bool initial_check, int &length)
{
    if (initial_check && ip >= ilimit) return false;
    unsigned s;
    do {
        s = rd(ip);
        ip++;
        length += (int)s;
        if (ip > ilimit) return false;
    } while (s == 255);
    return true;
}

FM_HD void lz4_parse_init(ParseState &s, bool src_null, int src_size, int dst_capacity, unsigned first_byte)
{
    s.ip = 0; s.op = 0; s.status = 0; s.result = 0; s.pend = false; s.pend_off = 0; s.pend_op = 0;
    s.fast = dst_capacity >= 64;                                                       // :1990
    if (src_null || dst_capacity < 0) { s.status = 1; s.result = -1; return; }         // :1951
    if (dst_capacity == 0) { s.status = 1; s.result = (src_size == 1 && first_byte == 0) ? 0 : -1; return; }   // :1977-1981
    if (src_size == 0) { s.status = 1; s.result = -1; return; }                        // :1982
}

// Runs sequences while the next token lies below ip_stop (tested at sequence boundaries only).
// Contract for staged readers: every position below ip_stop + 32 (capped at iend) is readable
// through rd.near().
template <class Reader, class Sink>
FM_HD void lz4_parse_run(ParseState &s, const Reader &rd, Sink &sink, const int iend, const int oend, const int ip_stop)
{
    int ip = s.ip, op = s.op;
    bool fast = s.fast, pend = s.pend;
    int pend_off = s.pend_off, pend_op = s.pend_op;
    int length, offset, cpy;
    unsigned token;

    while (ip < ip_stop) {
        if (fast) {
            // ---- fast loop, :1996-2115
            token = rd.near(ip);
            if (pend && pend_off > pend_op) goto error;                                // deferred :2065
            pend = false;
            sink.token(ip, op);
            ip++;
            length = (int)(token >> 4);
            if (length == 15) {
                if (!lz4_rvl(rd, ip, iend - 15, true, length)) goto error;
                cpy = op + length;
                if (cpy > oend - 32 || ip + length > iend - 32) { fast = false; goto safe_literal_copy; }
            } else {
                cpy = op + length;
                if (ip > iend - 17) { fast = false; goto safe_literal_copy; }
            }
            ip += length; op = cpy;

            // a long literal run may have carried ip far ahead: only then take the checked read
            offset = (token >= 0xF0) ? ((int)rd(ip) | ((int)rd(ip + 1) << 8))
                                     : ((int)rd.near(ip) | ((int)rd.near(ip + 1) << 8));
            ip += 2;
            length = (int)(token & 15);
            if (length == 15) {
                if (!lz4_rvl(rd, ip, iend - 5 + 1, false, length)) goto error;
                length += 4;
                if (offset > op) goto error;                                           // :2041
                if (op + length >= oend - 64) { fast = false; goto safe_match_copy; }
                op += length;
                continue;
            }
            length += 4;
            if (op + length >= oend - 64) { fast = false; goto safe_match_copy; }
            pend = true; pend_off = offset; pend_op = op;                              // :2065, tested next round
            op += length;
            continue;
        }

        // ---- safe loop, :2120-2328
        sink.token(ip, op);
        token = rd(ip); ip++;
        length = (int)(token >> 4);

        if (length != 15 && ip < iend - 16 && op <= oend - 32) {                       // :2132
            op += length; ip += length;
            length = (int)(token & 15);
            offset = (int)rd(ip) | ((int)rd(ip + 1) << 8); ip += 2;
            if (length != 15 && offset >= 8 && offset <= op) { op += length + 4; continue; }
            goto copy_match;
        }

        if (length == 15) {
            if (!lz4_rvl(rd, ip, iend - 15, true, length)) goto error;
        }
        cpy = op + length;
safe_literal_copy:
        if (cpy > oend - 12 || ip + length > iend - (2 + 1 + 5)) {                     // :2186
            if (ip + length != iend || cpy > oend) goto error;                         // :2208
            ip += length; op += length;
            s.status = 1; s.result = op;
            goto save;
        }
        ip += length; op = cpy;

        offset = (int)rd(ip) | ((int)rd(ip + 1) << 8); ip += 2;
        length = (int)(token & 15);
copy_match:
        if (length == 15) {
            if (!lz4_rvl(rd, ip, iend - 5 + 1, false, length)) goto error;
        }
        length += 4;
safe_match_copy:
        if (offset > op) goto error;                                                   // :2250
        cpy = op + length;
        if (cpy > oend - 5) goto error;                                                // :2317
        op = cpy;
    }
save:
    s.ip = ip; s.op = op; s.fast = fast; s.pend = pend; s.pend_off = pend_off; s.pend_op = pend_op;
    return;

error:
    s.status = 1; s.result = -ip - 1;                                                  // :2337
    goto save;
}

// One-shot convenience over a plain buffer (host tests, tiny inputs).
struct PtrReader {
    const uint8_t *p;
    FM_HDM unsigned operator()(int i) const { return p[i]; }
    FM_HDM unsigned near(int i) const { return p[i]; }
};

template <class Sink>
FM_HD int lz4_parse_block(const uint8_t *src, int src_size, int dst_capacity, Sink &sink, int step = 0x7fffffff)
{
    ParseState s;
    lz4_parse_init(s, src == nullptr, src_size, dst_capacity, (src && src_size > 0) ? src[0] : 0u);
    PtrReader rd{src};
    int stop = step;
    while (!s.status) {
        lz4_parse_run(s, rd, sink, src_size, dst_capacity, stop);
        stop = (stop > 0x7fffffff - step) ? 0x7fffffff : stop + step;
    }
    return s.result;
}

}  // namespace fm
