// lz4_parse.h -- sequential validation/parse of one LZ4 block (host + device).
//
// Decoding a 4 MiB LZ4 block on the GPU is split in two (DESIGN.md "LZ4 decode"):
//   D1  this file: ONE thread walks the token chain of a block, reproducing every accept/reject
//       decision of the reference decoder and emitting where each sequence starts;
//   D2  lz4_decode.cuh: warps copy literals and matches of 32 sequences at a time.
// All of LZ4_decompress_safe's error conditions depend only on positions and lengths, never on
// the decoded bytes (native/lz4/lz4.c:1936-2339), so D1 alone decides the call's return value,
// including the exact negative value -(ip)-1 (:2337).
//
// The reference runs a "fast" loop while >= 64 bytes of output space remain (:1996-2115) and a
// "safe" loop afterwards (:2120-2328); their checks differ slightly on malformed input (the
// shortcuts at :2014, :2132 skip the end-of-input test of :2186), so both are kept, as two
// states over integer positions.  The same structure is restated for the CPU in
// oracle/fourmc_oracle.c (test infrastructure); this is the product's own copy.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define FM_HD __host__ __device__ __forceinline__
#else
#define FM_HD static inline
#endif

namespace fm {

// variable-length field (lz4.c:1903-1928).  Returns false on error.
FM_HD bool lz4_rvl(const uint8_t *src, int &ip, int ilimit, bool initial_check, int &length)
{
    if (initial_check && ip >= ilimit) return false;
    unsigned s;
    do {
        s = src[ip];
        ip++;
        length += (int)s;
        if (ip > ilimit) return false;
    } while (s == 255);
    return true;
}

// Sink interface: void token(int token_pos, int op_before_literals)
// Called once per sequence, in stream order, before the sequence is validated any further than
// its token; a failing block's sink output is discarded by the caller.
template <class Sink>
FM_HD int lz4_parse_block(const uint8_t *src, int src_size, int dst_capacity, Sink &sink)
{
    int ip = 0, op = 0;
    const int iend = src_size, oend = dst_capacity;
    int length, offset, cpy;
    unsigned token;

    if (src == nullptr || dst_capacity < 0) return -1;                                 // :1951
    if (dst_capacity == 0) return (src_size == 1 && src[0] == 0) ? 0 : -1;             // :1977-1981
    if (src_size == 0) return -1;                                                      // :1982

    bool fast = (oend - op) >= 64;                                                     // :1990

    for (;;) {
        if (fast) {
            // ---- fast loop, :1996-2115
            sink.token(ip, op);
            token = src[ip++];
            length = (int)(token >> 4);
            if (length == 15) {
                if (!lz4_rvl(src, ip, iend - 15, true, length)) goto error;
                cpy = op + length;
                if (cpy > oend - 32 || ip + length > iend - 32) { fast = false; goto safe_literal_copy; }
            } else {
                cpy = op + length;
                if (ip > iend - 17) { fast = false; goto safe_literal_copy; }
            }
            ip += length; op = cpy;

            offset = (int)src[ip] | ((int)src[ip + 1] << 8); ip += 2;
            length = (int)(token & 15);
            if (length == 15) {
                if (!lz4_rvl(src, ip, iend - 5 + 1, false, length)) goto error;
                length += 4;
                if (offset > op) goto error;                                           // :2041
                if (op + length >= oend - 64) { fast = false; goto safe_match_copy; }
            } else {
                length += 4;
                if (op + length >= oend - 64) { fast = false; goto safe_match_copy; }
            }
            if (offset > op) goto error;                                               // :2065
            op += length;
            continue;
        }

        // ---- safe loop, :2120-2328
        sink.token(ip, op);
        token = src[ip++];
        length = (int)(token >> 4);

        if (length != 15 && ip < iend - 16 && op <= oend - 32) {                       // :2132
            op += length; ip += length;
            length = (int)(token & 15);
            offset = (int)src[ip] | ((int)src[ip + 1] << 8); ip += 2;
            if (length != 15 && offset >= 8 && offset <= op) { op += length + 4; continue; }
            goto copy_match;
        }

        if (length == 15) {
            if (!lz4_rvl(src, ip, iend - 15, true, length)) goto error;
        }
        cpy = op + length;
safe_literal_copy:
        if (cpy > oend - 12 || ip + length > iend - (2 + 1 + 5)) {                     // :2186
            if (ip + length != iend || cpy > oend) goto error;                         // :2208
            ip += length; op += length;
            break;
        }
        ip += length; op = cpy;

        offset = (int)src[ip] | ((int)src[ip + 1] << 8); ip += 2;
        length = (int)(token & 15);
copy_match:
        if (length == 15) {
            if (!lz4_rvl(src, ip, iend - 5 + 1, false, length)) goto error;
        }
        length += 4;
safe_match_copy:
        if (offset > op) goto error;                                                   // :2250
        cpy = op + length;
        if (cpy > oend - 5) goto error;                                                // :2317
        op = cpy;
    }
    return op;

error:
    return -ip - 1;                                                                    // :2337
}

}  // namespace fm
