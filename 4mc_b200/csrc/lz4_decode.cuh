// lz4_decode.cuh -- LZ4 block decode kernels (reference: LZ4_decompress_safe,
// native/lz4/lz4.c:2345, called per block at native/4mc.c:661 and native/jniDecompressor.c:88).
//
//  D1  lz4_parse_kernel   one WARP per block finds where every sequence starts, decides the return
//                         value, and leaves two small side tables in HBM (details at the kernel):
//                           tokmap   1 bit per compressed byte, set where a sequence's token sits
//                           chunk_op u32 per 128 compressed bytes: output position of the first
//                                    sequence whose token lies in that 128-byte chunk
//                         (0.16 bytes of scratch per compressed byte).
//  D2  lz4_copy_kernel    one CTA per block.  A warp takes 1 KiB of the compressed stream at a time
//                         (128 B when several warps share a block), lists its tokens and takes them 32
//                         at a time, one sequence per lane; output positions by prefix sum.  The batch's
//                         output span (~420 B on text) is assembled in shared memory as aligned 32-bit
//                         WORDS -- literal words from aligned words of the stream, match words from the
//                         aligned 16-byte chunks around the source in HBM / L2, both through a funnel
//                         shift; a word shared by two neighbouring sequences is stored once, by the left
//                         lane, with the right lane's half OR-ed in (shuffle) -- and flushed with 128-bit
//                         stores.  The rare rest is patched in with byte stores: sources inside the span
//                         (resolved in rounds, lowest destination first), offset 0, long runs.  A source
//                         may belong to a work item another warp is still copying, so warps publish the
//                         lowest output position they still owe (`owed[w]`) and a match waits until the
//                         minimum over all warps has passed the bytes it needs.  Spans over 2 KiB (long
//                         runs) are copied in place instead, by the whole warp.
//  D0  lz4_stored_kernel  csize == usize blocks are raw copies (native/4mc.c:635-642).
#pragma once

#include "fm_common.cuh"
#include "lz4_parse.h"

namespace fm {

constexpr int LZ4_CHUNK = 128;            // compressed bytes per D2 work item
constexpr int LZ4_CHUNK_WORDS = 4;        // tokmap words per chunk
constexpr int LZ4_MAX_TOK = 43;           // ceil(128 / 3): a sequence is at least 3 bytes
constexpr int LZ4_LONG = 48;              // copies this long are done by the whole warp

struct BlockDesc {                         // one per block, built on the device
    const uint8_t *src;                    // payload
    uint8_t *dst;
    uint32_t csize, usize;                 // usize = capacity offered to the decoder
    uint32_t chunk_base;                   // first chunk index in tokmap/chunk_op
    uint32_t stored;                       // 0: compressed, 1: raw copy, 2: nothing to decode (checksum-only item)
};

// ---- D1 ------------------------------------------------------------------------------------

// Token positions are recorded in ALIGNED coordinates q = ip + d, d = (payload address & 15), so
// that the ring, the 2 KiB halves and the 128-byte D2 chunks all line up with 16-byte loads.
struct TokSink {
    uint32_t *tokmap;      // block's first word
    uint32_t *chunk_op;    // block's first entry
    int d;
    uint32_t cur_word;     // index of the word being accumulated
    uint32_t bits;
    uint32_t cur_chunk;
    __device__ __forceinline__ void token(int ip, int op)
    {
        const uint32_t q = (uint32_t)(ip + d);
        const uint32_t w = q >> 5;
        if (w != cur_word) {
            if (bits) atomicOr(&tokmap[cur_word], bits);
            cur_word = w; bits = 0;
        }
        bits |= 1u << (q & 31);
        const uint32_t c = q >> 7;
        if (c != cur_chunk) { chunk_op[c] = (uint32_t)op; cur_chunk = c; }
    }
    __device__ __forceinline__ void flush() { if (bits) atomicOr(&tokmap[cur_word], bits); bits = 0; }
};

constexpr int D1_HALF = 2048;             // bytes per ring half
constexpr int D1_RING = 2 * D1_HALF;
constexpr int D1_WARPS = 4;               // blocks per CTA
constexpr int D1_SUB = 64;                // positions per lane per half (32 lanes x 64 = one half)
constexpr int D1_XS_STRIDE = D1_SUB + 1;                   // exit table: one word per position (bytes produced << 16 |
                                                           // exit), one word of padding per sub-chunk: both halves of
                                                           // an entry come with ONE scattered load
constexpr int D1_WARP_SMEM = D1_RING + 32 * D1_XS_STRIDE * 4 + 32 * 4 + 32 * 2;
constexpr int D1_SMEM = D1_WARPS * ((D1_WARP_SMEM + 15) & ~15);
constexpr int D1_SPECIAL = 0x8000;         // exit-table flag: the chain stops at a token with a long continued length
constexpr int D1_NONE = 0xffff;

// bytes of the compressed block through this warp's shared-memory ring; positions outside the
// staged window (far look-ahead after a long literal run) fall back to global memory
struct RingReader {
    const uint8_t *ring;
    const uint8_t *src;
    int d;                 // misalignment of src: ring holds the 16-byte-aligned stream
    int lo, span;          // staged window [lo, lo + span) in block coordinates
    __device__ __forceinline__ unsigned near(int ip) const { return ring[(ip + d) & (D1_RING - 1)]; }
    __device__ __forceinline__ unsigned operator()(int ip) const
    {
        if ((unsigned)(ip - lo) < (unsigned)span) return ring[(ip + d) & (D1_RING - 1)];
        return src[ip];
    }
};

// One sequence, any length, for the bulk phase: everything read lies below `limit` (block
// coordinates) or the sequence is reported as not clean.  next = position of the next token.
struct SeqDec { int lit, ml, off, next; bool clean; };
__device__ __forceinline__ SeqDec d1_decode_slow(const RingReader &rd, int ip, int limit)
{
    SeqDec r; r.clean = false; r.lit = r.ml = r.off = 0; r.next = ip;
    if (ip >= limit) return r;
    const unsigned tok = rd(ip++);
    int lit = (int)(tok >> 4);
    if (lit == 15) {
        unsigned b;
        do { if (ip >= limit) return r; b = rd(ip++); lit += (int)b; } while (b == 255);
    }
    if (lit > limit - ip) return r;
    ip += lit;
    if (ip + 2 > limit) return r;
    r.off = (int)rd(ip) | ((int)rd(ip + 1) << 8); ip += 2;
    int ml = (int)(tok & 15);
    if (ml == 15) {
        unsigned b;
        do { if (ip >= limit) return r; b = rd(ip++); ml += (int)b; } while (b == 255);
    }
    r.lit = lit; r.ml = ml + 4; r.next = ip; r.clean = true;
    return r;
}

// Phase B of the bulk parse (see lz4_parse_kernel): the exit function of this lane's 64-position sub-chunk of
// the half at aligned position base_q, folded back to front.  Each position is read AS IF a token started there:
// a sequence without continued lengths is 3 + lit bytes long and produces lit + ml + 4 bytes; a length continued by
// ONE byte is folded in (literal runs of 15+ and matches of 19+ are common); longer continuations stop the fold
// (D1_SPECIAL: the chain walk decodes that sequence byte by byte).
__device__ __forceinline__ void d1_fold(const uint8_t *ring, uint32_t *XS, const int base_q, const int lane)
{
    const uint4 *mine = (const uint4 *)(ring + ((base_q + lane * D1_SUB) & (D1_RING - 1)));
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 4; i++) { const uint4 v = mine[i]; w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
    uint32_t *xs = XS + lane * D1_XS_STRIDE;
    const int sub_q = base_q + lane * D1_SUB;
#pragma unroll
    for (int jj = D1_SUB - 1; jj >= 0; jj--) {
        const unsigned tok = (w[jj >> 2] >> ((jj & 3) * 8)) & 0xffu;
        int lit = (int)(tok >> 4), ml = (int)(tok & 15), n = jj + 3;
        bool special = false;
        if (lit == 15) {
            const unsigned x = (jj + 1 < D1_SUB) ? ((w[(jj + 1) >> 2 & 15] >> (((jj + 1) & 3) * 8)) & 0xffu)
                                                : (unsigned)ring[(sub_q + D1_SUB) & (D1_RING - 1)];
            special = x == 255;
            lit += (int)x; n++;
        }
        n += lit;
        if (ml == 15 && !special) {
            const unsigned x = ring[(sub_q + n) & (D1_RING - 1)];
            special = x == 255;
            ml += (int)x; n++;
        }
        // (bytes produced << 16) | exit: at most 21 sequences of at most 544 bytes each fit the upper half
        uint32_t e = ((uint32_t)(lit + ml + 4) << 16) | (uint32_t)n;
        if (special) e = (uint32_t)(jj | D1_SPECIAL);
        else if (n < D1_SUB) e = xs[n] + ((uint32_t)(lit + ml + 4) << 16);
        xs[jj] = e;
    }
}

// Phase D: this lane's sub-chunk from its real entry (s_entry / s_op, set by the chain walk): token bits, the
// offset test (lz4.c:2041/:2065) and the first sequence that is not clean (next token beyond clean_ip, or output
// at or beyond clean_op).
struct D1Mark { uint32_t w0, w1; int first_op, viol, viol_next, uncl, uncl_op; };
__device__ __forceinline__ D1Mark d1_mark(const uint8_t *ring, const RingReader &rd, const uint16_t *s_entry, const uint32_t *s_op,
                                          const int base_q, const int d, const int clean_ip, const int clean_op, const int lane)
{
    D1Mark m;
    m.w0 = 0; m.w1 = 0; m.first_op = -1; m.viol = 0x7fffffff; m.viol_next = 0; m.uncl = 0x7fffffff; m.uncl_op = 0;
    int p = (int)s_entry[lane];
    if (p == D1_NONE) return m;
    int o = (int)s_op[lane];
    const int s1 = (lane + 1) * D1_SUB;
    while (p < s1) {
        const int ip = base_q + p - d;
        int lit, mlen, off, next;
        const unsigned tok = ring[(base_q + p) & (D1_RING - 1)];
        bool slow = false;
        {
            int qo = base_q + p + 1;
            lit = (int)(tok >> 4); mlen = (int)(tok & 15) + 4;
            if (lit == 15) {                  // continued literal length: one more byte, usually the last
                const unsigned x = ring[qo & (D1_RING - 1)];
                slow = x == 255;
                lit += (int)x; qo++;
            }
            qo += lit;
            // up to 269 literals ahead: still inside the two staged halves
            off = (int)ring[qo & (D1_RING - 1)] | ((int)ring[(qo + 1) & (D1_RING - 1)] << 8);
            qo += 2;
            if (mlen == 19 && !slow) {        // continued match length
                const unsigned x = ring[qo & (D1_RING - 1)];
                slow = x == 255;
                mlen += (int)x; qo++;
            }
            next = qo - d;
        }
        if (slow) {
            const SeqDec sd = d1_decode_slow(rd, ip, clean_ip);
            if (!sd.clean) { m.uncl = p; m.uncl_op = o; break; }
            lit = sd.lit; mlen = sd.ml; off = sd.off; next = sd.next;
        }
        if (next > clean_ip || o + lit + mlen >= clean_op) { m.uncl = p; m.uncl_op = o; break; }
        if (off > o + lit) { m.viol = p; m.viol_next = next; break; }         // lz4.c:2041/:2065
        if (m.first_op < 0) m.first_op = o;
        const int bit = p - lane * D1_SUB;
        if (bit < 32) m.w0 |= 1u << bit; else m.w1 |= 1u << (bit - 32);
        o += lit + mlen;
        p = next + d - base_q;
    }
    return m;
}

// D1.  One WARP per block.  All lanes stream the payload through a double-buffered ring.
//
// BULK phase (all of the block except its last few sequences), one 2 KiB half at a time:
//   B  each lane reads every byte position of its own 64-byte sub-chunk AS IF a token started
//      there and folds the sub-chunk back to front into an "exit function": for every entry
//      position, where the chain leaves the sub-chunk and how many bytes it produced on the way
//      (straight-line code over 16 registers; the only memory traffic is the table itself);
//   C  lane 0 walks the real chain with ONE table lookup per sub-chunk instead of one dependent
//      step per sequence (32 hops per half instead of ~400 sequences);
//   D  every lane re-walks its sub-chunk from its real entry: sets the token bits, checks the
//      one test that can fail this far from the ends of the buffers (offset beyond the start of
//      the output, lz4.c:2065) and watches for the first sequence that comes within 32 bytes of
//      the end of the input or 64 bytes of the end of the output.
// From that first "not clean" sequence on, lane 0 runs the exact two-loop state machine of
// lz4_parse.h (TAIL phase); every end-of-buffer rule of the reference lives there.  A sequence is
// clean iff next_token <= iend - 32 and op_after < oend - 64; both grow along the chain, so the
// clean sequences are a prefix of the stream and none of the reference's fast-loop exits
// (:2010, :2014, :2045, :2050) nor any length-field limit (:1911-1920) can trigger inside it.
__global__ void __launch_bounds__(D1_WARPS * 32)
lz4_parse_kernel(const BlockDesc *blocks, uint32_t n_blocks, uint32_t *tokmap, uint32_t *chunk_op, int32_t *result)
{
    extern __shared__ __align__(16) uint8_t d1_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * D1_WARPS + warp;
    if (b >= n_blocks) return;
    const BlockDesc bd = blocks[b];
    if (bd.stored) { if (lane == 0) result[b] = (int32_t)bd.usize; return; }

    uint8_t *ring = d1_smem + (size_t)warp * ((D1_WARP_SMEM + 15) & ~15);
    uint32_t *XS = (uint32_t *)(ring + D1_RING);                   // [32][D1_XS_STRIDE] from each entry: bytes produced << 16 | exit
    uint32_t *s_op = XS + 32 * D1_XS_STRIDE;                       // per sub-chunk: op at its entry
    uint16_t *s_entry = (uint16_t *)(s_op + 32);                   // per sub-chunk: entry position

    const uintptr_t a = (uintptr_t)bd.src;
    const uint4 *base = (const uint4 *)(a & ~(uintptr_t)15);
    const int d = (int)(a & 15);
    const int csize = (int)bd.csize, oend = (int)bd.usize;
    const int nchunks = (d + csize + 15) >> 4;            // 16-byte chunks holding payload bytes
    constexpr int HC = D1_HALF / 16;                      // chunks per half
    uint32_t *my_map = tokmap + (size_t)bd.chunk_base * LZ4_CHUNK_WORDS;
    uint32_t *my_cop = chunk_op + bd.chunk_base;

    uint4 r0, r1, r2, r3;
    auto load_half = [&](int h) {
        const int c = h * HC + lane;
        const uint4 z = make_uint4(0, 0, 0, 0);
        r0 = (c < nchunks) ? ldg_nc_v4(base + c) : z;
        r1 = (c + 32 < nchunks) ? ldg_nc_v4(base + c + 32) : z;
        r2 = (c + 64 < nchunks) ? ldg_nc_v4(base + c + 64) : z;
        r3 = (c + 96 < nchunks) ? ldg_nc_v4(base + c + 96) : z;
    };
    auto store_half = [&](int h) {
        uint4 *q = (uint4 *)(ring + (h & 1) * D1_HALF);
        q[lane] = r0; q[lane + 32] = r1; q[lane + 64] = r2; q[lane + 96] = r3;
    };

    ParseState st;
    TokSink sink;
    sink.tokmap = my_map; sink.chunk_op = my_cop; sink.d = d;
    sink.cur_word = 0; sink.bits = 0; sink.cur_chunk = 0xffffffffu;
    lz4_parse_init(st, bd.src == nullptr, csize, oend,
                   (lane == 0 && bd.src != nullptr && csize > 0) ? (unsigned)bd.src[0] : 0u);

    int k = 0;                                            // ring holds halves k and k+1
    if (!st.status) {
        // ================= BULK =================
        const int clean_ip = csize - 32;                  // next token must not pass this
        const int clean_op = oend - 64;                   // op after the sequence must stay below
        int e_q = d, e_op = 0;                            // chain entry (aligned coordinate) and its op
        bool bulk = clean_ip > 0 && clean_op > 0;
        bool staged = false;                              // ring already holds halves k, k+1
        while (bulk) {
            k = e_q >> 11;
            if (!staged) { load_half(k); store_half(k); load_half(k + 1); store_half(k + 1); __syncwarp(); }
            const bool more = (k + 2) * HC < nchunks;
            if (more) load_half(k + 2);                   // in flight during the phases below
            const int base_q = k * D1_HALF;
            RingReader rd;
            rd.ring = ring; rd.src = bd.src; rd.d = d;
            rd.lo = base_q - d; rd.span = min(base_q + D1_RING - d, csize) - rd.lo;

            // ---- B: exit function of my 64-position sub-chunk, back to front
            d1_fold(ring, XS, base_q, lane);
            s_entry[lane] = (uint16_t)D1_NONE;
            __syncwarp();
            // ---- C: lane 0 hops sub-chunk to sub-chunk along the real chain
            int e = e_q - base_q, op = e_op;
            if (lane == 0) {
                int last_sc = -1;
                while (e < D1_HALF) {
                    const int sc = e >> 6, jj = e & (D1_SUB - 1);
                    if (sc != last_sc) { s_entry[sc] = (uint16_t)e; s_op[sc] = (uint32_t)op; last_sc = sc; }
                    const uint32_t xs = XS[sc * D1_XS_STRIDE + jj];
                    const int x = (int)(xs & 0xffffu);
                    op += (int)(xs >> 16);
                    if (x & D1_SPECIAL) {
                        const SeqDec sd = d1_decode_slow(rd, base_q + sc * D1_SUB + (x & (D1_SUB - 1)) - d, clean_ip);   // flag | position 0..63
                        if (!sd.clean) break;             // phase D finds it too and ends the bulk phase
                        op += sd.lit + sd.ml;
                        e = sd.next + d - base_q;
                    } else e = sc * D1_SUB + x;
                }
            }
            e = __shfl_sync(FM_FULL, e, 0); op = __shfl_sync(FM_FULL, op, 0);
            __syncwarp();
            // ---- D: my sub-chunk from its real entry: bits, offset test, first unclean sequence
            const D1Mark mk = d1_mark(ring, rd, s_entry, s_op, base_q, d, clean_ip, clean_op, lane);
            const uint32_t w0 = mk.w0, w1 = mk.w1;
            const int first_op = mk.first_op, viol = mk.viol, viol_next = mk.viol_next, uncl = mk.uncl, uncl_op = mk.uncl_op;
            const int pu = warp_min(uncl), pv = warp_min(viol);
            if (pv < pu) {                                // corrupt: offset before the start of the output
                const int src_lane = __ffs(__ballot_sync(FM_FULL, viol == pv)) - 1;
                const int vn = __shfl_sync(FM_FULL, viol_next, src_lane);
                st.status = 1; st.result = -vn - 1;       // lz4.c:2337 with ip at the next token
                bulk = false;
                break;
            }
            // token bits and per-chunk output positions of this half
            {
                uint32_t *mw = my_map + (base_q >> 5) + 2 * lane;
                if (w0) mw[0] = w0;
                if (w1) mw[1] = w1;
                const int f_even = __shfl_sync(FM_FULL, first_op, lane & ~1), f_odd = __shfl_sync(FM_FULL, first_op, lane | 1);
                const int f = f_even >= 0 ? f_even : f_odd;
                if (!(lane & 1) && f >= 0) my_cop[(base_q >> 7) + (lane >> 1)] = (uint32_t)f;
            }
            if (pu != 0x7fffffff) {                       // hand over to the exact state machine
                const int src_lane = __ffs(__ballot_sync(FM_FULL, uncl == pu)) - 1;
                st.ip = base_q + pu - d;
                st.op = __shfl_sync(FM_FULL, uncl_op, src_lane);
                // the word and chunk holding the hand-over token may already carry earlier tokens
                const int c_first = __shfl_sync(FM_FULL, first_op, (pu >> 6) & ~1);
                const int c_mine = __shfl_sync(FM_FULL, first_op, pu >> 6);
                const bool chunk_has = ((pu >> 6) & 1) ? (c_first >= 0 || c_mine >= 0) : (c_mine >= 0);
                sink.cur_word = (uint32_t)(base_q + pu) >> 5; sink.bits = 0;
                sink.cur_chunk = chunk_has ? (uint32_t)(base_q + pu) >> 7 : 0xffffffffu;
                bulk = false;
                __syncwarp();
                break;
            }
            // next half
            e_q = base_q + e; e_op = op;
            __syncwarp();
            if (e < D1_RING && more) { store_half(k + 2); staged = true; }
            else staged = false;                          // jumped far ahead (or nothing left): restage
            __syncwarp();
            if (e_q - d >= csize) {                       // cannot happen after a clean sequence; be safe
                st.status = 1; st.result = -(e_q - d) - 1;
                bulk = false;
            }
        }
    }

    if (!st.status) {
        // ================= TAIL =================
        k = (st.ip + d) >> 11;
        __syncwarp();
        load_half(k); store_half(k);
        load_half(k + 1); store_half(k + 1);
        __syncwarp();
        for (;; k++) {
            const bool more = (k + 2) * HC < nchunks;
            if (more) load_half(k + 2);                   // in flight during the walk below
            if (lane == 0 && !st.status) {
                RingReader rd;
                rd.ring = ring; rd.src = bd.src; rd.d = d;
                rd.lo = k * D1_HALF - d;
                rd.span = min((k + 2) * D1_HALF - d, csize) - rd.lo;
                // tokens below the middle of the window keep their next 32 bytes staged
                const int stop = more ? (k + 1) * D1_HALF - d : 0x7fffffff;
                lz4_parse_run(st, rd, sink, csize, oend, stop);
            }
            if (__shfl_sync(FM_FULL, st.status, 0)) break;
            __syncwarp();
            if (more) store_half(k + 2);
            __syncwarp();
        }
    }
    if (lane == 0) { sink.flush(); result[b] = st.result; }
}

// D1 for FEW blocks: one CTA of 16 warps per block (lz4_parse_kernel gives a block one warp: 9 ms however few blocks
// there are -- the per-block call of the JNI path, a split of two or three blocks, the slices of the host pipeline).
// The three phases of the bulk parse become a pipeline over the block's 2 KiB halves:
//   warps 1..15  WORKERS: worker i takes halves i, i + 15, ...: stages the half (and the next one, for look-ahead) in
//                its own ring, folds it (phase B), hands it to the chain warp, and when the chain has passed through
//                it marks its tokens (phase D);
//   warp 0       CHAIN: walks the real chain through the halves in order (phase C), one table lookup per 64-byte
//                sub-chunk -- the only serial part: about 1 ms for a 4 MiB block.
// The first sequence that is not clean (and any offset violation) is reported through shared memory; after the
// pipeline has drained, warp 0 runs the exact state machine from there, exactly like lz4_parse_kernel's tail.
// The chain warp, not the workers, bounds a block's time (2.7 ms); 7 workers keep up with it as well as 15 do and leave
// room for a second CTA on the SM: D1W_WORKERS = 15 for up to one block per SM, 7 for up to two.
constexpr int D1W_SLOT = (D1_WARP_SMEM + 15) & ~15;
constexpr int D1W_MAX_WORKERS = 15;
template <int D1W_WORKERS> constexpr int d1w_smem_bytes() { return (D1W_WORKERS + 1) * D1W_SLOT + 256; }

struct D1WCtl {
    volatile int b_ready[D1W_MAX_WORKERS]; // half whose exit tables stand in the worker's slot
    volatile int c_done[D1W_MAX_WORKERS];  // half the chain has passed through
    int stop_h;                            // halves beyond this one need no work (first unclean sequence / violation / end)
    unsigned long long uncl, viol;         // (aligned position << 32) | op, resp. | next token
};

template <int D1W_WORKERS>
__global__ void __launch_bounds__((D1W_WORKERS + 1) * 32, D1W_WORKERS > 7 ? 1 : 2)
lz4_parse_wide_kernel(const BlockDesc *blocks, uint32_t n_blocks, uint32_t *tokmap, uint32_t *chunk_op, int32_t *result)
{
    extern __shared__ __align__(16) uint8_t d1w_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x;
    if (b >= n_blocks) return;
    const BlockDesc bd = blocks[b];
    if (bd.stored) { if (threadIdx.x == 0) result[b] = (int32_t)bd.usize; return; }

    D1WCtl *ctl = (D1WCtl *)(d1w_smem + (size_t)(D1W_WORKERS + 1) * D1W_SLOT);
    uint8_t *ring = d1w_smem + (size_t)warp * D1W_SLOT;             // this warp's slot
    uint32_t *XS = (uint32_t *)(ring + D1_RING);
    uint32_t *s_op = XS + 32 * D1_XS_STRIDE;
    uint16_t *s_entry = (uint16_t *)(s_op + 32);

    const uintptr_t a = (uintptr_t)bd.src;
    const uint4 *base = (const uint4 *)(a & ~(uintptr_t)15);
    const int d = (int)(a & 15);
    const int csize = (int)bd.csize, oend = (int)bd.usize;
    const int nchunks = (d + csize + 15) >> 4;
    constexpr int HC = D1_HALF / 16;
    const int n_halves = (d + csize + D1_HALF - 1) / D1_HALF;
    uint32_t *my_map = tokmap + (size_t)bd.chunk_base * LZ4_CHUNK_WORDS;
    uint32_t *my_cop = chunk_op + bd.chunk_base;

    uint4 r0, r1, r2, r3;
    auto load_half = [&](int h) {
        const int c = h * HC + lane;
        const uint4 z = make_uint4(0, 0, 0, 0);
        r0 = (c < nchunks) ? ldg_nc_v4(base + c) : z;
        r1 = (c + 32 < nchunks) ? ldg_nc_v4(base + c + 32) : z;
        r2 = (c + 64 < nchunks) ? ldg_nc_v4(base + c + 64) : z;
        r3 = (c + 96 < nchunks) ? ldg_nc_v4(base + c + 96) : z;
    };
    auto store_half = [&](int h) {
        uint4 *q = (uint4 *)(ring + (h & 1) * D1_HALF);
        q[lane] = r0; q[lane + 32] = r1; q[lane + 64] = r2; q[lane + 96] = r3;
    };

    ParseState st;
    TokSink sink;
    sink.tokmap = my_map; sink.chunk_op = my_cop; sink.d = d;
    sink.cur_word = 0; sink.bits = 0; sink.cur_chunk = 0xffffffffu;
    lz4_parse_init(st, bd.src == nullptr, csize, oend, (bd.src != nullptr && csize > 0) ? (unsigned)bd.src[0] : 0u);
    const int clean_ip = csize - 32, clean_op = oend - 64;
    const bool bulk = !st.status && clean_ip > 0 && clean_op > 0;

    if (threadIdx.x < D1W_WORKERS) { ctl->b_ready[threadIdx.x] = -1; ctl->c_done[threadIdx.x] = -1; }
    if (threadIdx.x == 0) { ctl->stop_h = n_halves - 1; ctl->uncl = ~0ull; ctl->viol = ~0ull; }
    __syncthreads();
    volatile int *stop_h = &ctl->stop_h;

    if (bulk && warp > 0) {
        // ================= WORKER =================
        const int me = warp - 1;
        for (int h = me; h < n_halves; h += D1W_WORKERS) {
            if (h > *stop_h) break;
            const int base_q = h * D1_HALF;
            load_half(h); store_half(h);
            load_half(h + 1); store_half(h + 1);
            __syncwarp();
            d1_fold(ring, XS, base_q, lane);
            s_entry[lane] = (uint16_t)D1_NONE;
            __syncwarp();
            __threadfence_block();
            if (lane == 0) ctl->b_ready[me] = h;
            bool go = true;
            while (ctl->c_done[me] != h) {
                if (h > *stop_h) { go = false; break; }
                __nanosleep(64);
            }
            if (!go) break;
            __threadfence_block();
            RingReader rd;
            rd.ring = ring; rd.src = bd.src; rd.d = d;
            rd.lo = base_q - d; rd.span = min(base_q + D1_RING - d, csize) - rd.lo;
            const D1Mark mk = d1_mark(ring, rd, s_entry, s_op, base_q, d, clean_ip, clean_op, lane);
            {
                uint32_t *mw = my_map + (base_q >> 5) + 2 * lane;
                if (mk.w0) mw[0] = mk.w0;
                if (mk.w1) mw[1] = mk.w1;
                const int f_even = __shfl_sync(FM_FULL, mk.first_op, lane & ~1), f_odd = __shfl_sync(FM_FULL, mk.first_op, lane | 1);
                const int f = f_even >= 0 ? f_even : f_odd;
                if (!(lane & 1) && f >= 0) my_cop[(base_q >> 7) + (lane >> 1)] = (uint32_t)f;
            }
            const int pu = warp_min(mk.uncl), pv = warp_min(mk.viol);
            if (pv != 0x7fffffff) {
                const int src_lane = __ffs(__ballot_sync(FM_FULL, mk.viol == pv)) - 1;
                const int vn = __shfl_sync(FM_FULL, mk.viol_next, src_lane);
                if (lane == 0) { atomicMin(&ctl->viol, ((unsigned long long)(base_q + pv) << 32) | (uint32_t)vn); atomicMin(&ctl->stop_h, h); }
            }
            if (pu != 0x7fffffff) {
                const int src_lane = __ffs(__ballot_sync(FM_FULL, mk.uncl == pu)) - 1;
                const int uo = __shfl_sync(FM_FULL, mk.uncl_op, src_lane);
                if (lane == 0) { atomicMin(&ctl->uncl, ((unsigned long long)(base_q + pu) << 32) | (uint32_t)uo); atomicMin(&ctl->stop_h, h); }
            }
            __syncwarp();
        }
    } else if (bulk) {
        // ================= CHAIN =================
        int e_q = d, op = 0;
        for (int h = 0; h < n_halves; h++) {
            if (h > *stop_h) break;
            const int me = h % D1W_WORKERS;
            uint8_t *wring = d1w_smem + (size_t)(me + 1) * D1W_SLOT;
            uint32_t *wXS = (uint32_t *)(wring + D1_RING);
            uint32_t *w_op = wXS + 32 * D1_XS_STRIDE;
            uint16_t *w_entry = (uint16_t *)(w_op + 32);
            bool go = true;
            while (ctl->b_ready[me] != h) {
                if (h > *stop_h) { go = false; break; }
                __nanosleep(32);
            }
            if (!go) break;
            __threadfence_block();
            const int base_q = h * D1_HALF;
            int e = e_q - base_q;
            if (lane == 0 && e < D1_HALF) {
                RingReader rd;
                rd.ring = wring; rd.src = bd.src; rd.d = d;
                rd.lo = base_q - d; rd.span = min(base_q + D1_RING - d, csize) - rd.lo;
                int last_sc = -1;
                while (e < D1_HALF) {
                    const int sc = e >> 6, jj = e & (D1_SUB - 1);
                    if (sc != last_sc) { w_entry[sc] = (uint16_t)e; w_op[sc] = (uint32_t)op; last_sc = sc; }
                    const uint32_t xs = wXS[sc * D1_XS_STRIDE + jj];
                    const int x = (int)(xs & 0xffffu);
                    op += (int)(xs >> 16);
                    if (x & D1_SPECIAL) {
                        const SeqDec sd = d1_decode_slow(rd, base_q + sc * D1_SUB + (x & (D1_SUB - 1)) - d, clean_ip);
                        if (!sd.clean) { e = 0x40000000; break; }      // phase D finds it and reports; nothing clean follows
                        op += sd.lit + sd.ml;
                        e = sd.next + d - base_q;
                    } else e = sc * D1_SUB + x;
                }
            }
            e = __shfl_sync(FM_FULL, e, 0); op = __shfl_sync(FM_FULL, op, 0);
            __threadfence_block();
            if (lane == 0) ctl->c_done[me] = h;
            if (e >= 0x40000000) { if (lane == 0) atomicMin(&ctl->stop_h, h); break; }
            e_q = base_q + e;
        }
    }
    __syncthreads();
    if (warp != 0) return;

    // ================= hand-over and TAIL (warp 0) =================
    if (bulk) {
        const unsigned long long u = ctl->uncl, v = ctl->viol;
        if ((uint32_t)(v >> 32) < (uint32_t)(u >> 32)) {          // corrupt: offset before the start of the output
            st.status = 1; st.result = -(int)(uint32_t)v - 1;     // lz4.c:2337 with ip at the next token
        } else if (u != ~0ull) {
            const uint32_t q = (uint32_t)(u >> 32);
            st.ip = (int)q - d; st.op = (int)(uint32_t)u;
            // the chunk holding the hand-over token may already carry earlier tokens (then its position is set)
            const uint32_t c = q >> 7;
            bool chunk_has = false;
            for (uint32_t w = c * 4; w <= (q >> 5); w++) {
                uint32_t bits = my_map[w];
                if (w == (q >> 5)) bits &= (1u << (q & 31)) - 1u;
                chunk_has |= bits != 0;
            }
            sink.cur_word = q >> 5; sink.bits = 0;
            sink.cur_chunk = chunk_has ? c : 0xffffffffu;
        }
        // neither reported (cannot happen: the last sequence of a block is never clean): the exact machine from the start
    }
    if (!st.status) {
        int k = (st.ip + d) >> 11;
        __syncwarp();
        load_half(k); store_half(k);
        load_half(k + 1); store_half(k + 1);
        __syncwarp();
        for (;; k++) {
            const bool more = (k + 2) * HC < nchunks;
            if (more) load_half(k + 2);
            if (lane == 0 && !st.status) {
                RingReader rd;
                rd.ring = ring; rd.src = bd.src; rd.d = d;
                rd.lo = k * D1_HALF - d;
                rd.span = min((k + 2) * D1_HALF - d, csize) - rd.lo;
                const int stop = more ? (k + 1) * D1_HALF - d : 0x7fffffff;
                lz4_parse_run(st, rd, sink, csize, oend, stop);
            }
            if (__shfl_sync(FM_FULL, st.status, 0)) break;
            __syncwarp();
            if (more) store_half(k + 2);
            __syncwarp();
        }
    }
    if (lane == 0) { sink.flush(); result[b] = st.result; }
}

// ---- D2 ------------------------------------------------------------------------------------

__device__ __forceinline__ void warp_copy_bytes(uint8_t *dst, const uint8_t *src, int n)
{
    for (int i = lane_id(); i < n; i += 32) dst[i] = src[i];
}

// match copy by the whole warp; all source bytes below `d` are already complete.
__device__ __forceinline__ void warp_copy_match(uint8_t *out, int d, int off, int ml)
{
    const int lane = lane_id();
    if (off == 0) {                                   // lz4.c:2300-2303: offset 0 yields zeros
        for (int i = lane; i < ml; i += 32) out[d + i] = 0;
    } else if (off >= ml) {
        for (int i = lane; i < ml; i += 32) out[d + i] = out[d - off + i];
    } else if (off < 32) {                            // periodic fill from the `off` bytes before d
        for (int i = lane; i < ml; i += 32) out[d + i] = out[d - off + (i % off)];
    } else {                                          // 32 <= off < ml: 32 bytes per step
        for (int i = 0; i < ml; i += 32) {
            if (i + lane < ml) out[d + i + lane] = out[d - off + i + lane];
            __syncwarp();
        }
    }
}

// sequential byte copy inside shared memory (dst > src, may self-overlap when dst - src < n)
__device__ __forceinline__ void smem_copy_seq(uint8_t *dst, const uint8_t *src, int n)
{
    if (dst - src >= 8) {                             // 8 loads in flight, then 8 stores
        int i = 0;
        for (; i + 8 <= n; i += 8) {
            uint8_t t[8];
#pragma unroll
            for (int j = 0; j < 8; j++) t[j] = src[i + j];
#pragma unroll
            for (int j = 0; j < 8; j++) dst[i + j] = t[j];
        }
        for (; i < n; i++) dst[i] = src[i];
    } else {
        for (int i = 0; i < n; i++) dst[i] = src[i];
    }
}

constexpr int LZ4_SPAN = 2048;            // output bytes a warp assembles in shared memory at once
constexpr int LZ4_SCR = 80;               // scratch bytes per lane (80 keeps 128-bit stores conflict-free)
constexpr int LZ4_MAC_CHUNKS = 8;         // 128-byte chunks per D2 work item when a warp has a block to itself: 1 KiB,
                                          // 32 tokmap words, one per lane
constexpr int LZ4_MAC_TOK = 352;          // > ceil(1024 / 3): a sequence is at least 3 bytes
constexpr int LZ4_WLIT = 64;              // literal runs up to this long travel as words with their sequence
constexpr int LZ4_WMATCH = 32;            // and matches up to this long (three aligned 16-byte loads cover them)

// Shared memory of one copy warp / of the W warps that share a block.
struct CopyWarpSmem {
    __align__(16) uint8_t span[LZ4_SPAN + 32];
    __align__(16) uint8_t scr[32 * LZ4_SCR + 16]; // per lane: 16 bytes of slack + the aligned 16-byte chunks of a match source
    uint16_t tokpos[LZ4_MAC_TOK];                 // token positions of the work item, in stream order
};
template <int W>
struct CopyBlockSmem {
    int owed[W];
    int ticket;
};

// The copy of one block by W warps (`warp` = 0 .. W-1 within the block).
//
// Work item = C 128-byte chunks of the compressed stream (4 C tokmap words, one per lane): its tokens are listed in
// shared memory and taken 32 at a time, one sequence per lane.  C = 8 when the warp has the block to itself: full
// batches whatever the sequences' sizes (a 128-byte chunk of the log text holds 20), a span of ~420 B.  C = 1 when W
// warps share a block: an item's matches mostly point a few KiB back, i.e. into the items of the sibling warps, so
// the window in flight (W items) must stay small (r02n: 1 KiB items with 8 warps, 8.4 -> 19 ms for 64 blocks).
// A batch's output span is assembled in shared memory and flushed with 128-bit stores.
//
// Assembly.  99 % of the sequences are a literal run of at most LZ4_WLIT bytes and a match of at most LZ4_WMATCH bytes
// whose source lies wholly before the span.  Such a lane builds its sequence's bytes as aligned 32-bit WORDS of the
// span: literal words from aligned words of the stream, match words from the aligned 16-byte chunks of the source
// (HBM / L2 -> registers -> the lane's scratch slot), both through a funnel shift, bytes outside the sequence zeroed.
// Only a lane's first and last word can be shared with a neighbour (a sequence is at least 4 bytes): the LEFT lane
// stores the shared word, OR-ing in the right lane's half that it gets by shuffle.  No byte stores, no read-modify-write.
// Everything else is patched in afterwards with byte stores, which disturb nobody: long literal runs (whole warp per
// run), and "slow" matches -- a source inside the span (0.7 % on text: resolved in rounds, lowest destination first),
// offset 0 (zeros, lz4.c:2300-2303), long matches, and (W > 1) a source the sibling warps have not delivered yet.
// A lane with a slow match takes no part in the word stage at all.
template <int W, int C>
__device__ __forceinline__ void lz4_copy_block(const BlockDesc &bd, const uint32_t *tokmap, const uint32_t *chunk_op,
                                               const int warp, const int lane, CopyBlockSmem<W> *bs, CopyWarpSmem *ws)
{
    const uint8_t *__restrict__ src = bd.src;
    uint8_t *out = bd.dst;
    const int csize = (int)bd.csize;
    const int dq = (int)((uintptr_t)bd.src & 15);     // D1 records tokens at aligned positions ip + dq
    const int nchunks = (dq + csize + LZ4_CHUNK - 1) / LZ4_CHUNK;
    const int nmac = (nchunks + C - 1) / C;
    const int nwords = nchunks * LZ4_CHUNK_WORDS;
    const uint32_t *maps = tokmap + (size_t)bd.chunk_base * LZ4_CHUNK_WORDS;
    const uint32_t *cops = chunk_op + bd.chunk_base;
    // literal words are fetched as the aligned words around their bytes: never outside the payload's own words
    const uint32_t *src_w_lo = (const uint32_t *)((uintptr_t)src & ~(uintptr_t)3);
    const int src_w_last = (int)((((uintptr_t)(src + (csize > 0 ? csize - 1 : 0)) & ~(uintptr_t)3) - (uintptr_t)src_w_lo) >> 2);
    volatile int *owed = bs->owed;
    uint8_t *span = ws->span;
    uint32_t *span32 = (uint32_t *)ws->span;
    uint16_t *s_tokpos_w = ws->tokpos;
    uint8_t *s_scr_w = ws->scr;

    // lowest output position any warp of this block still owes
    auto high_water = [&]() -> int {
        if (W == 1) return 0x7fffffff;
        int v = lane < W ? owed[lane] : 0x7fffffff;
#pragma unroll
        for (int s = W / 2; s > 0; s >>= 1) v = min(v, __shfl_xor_sync(FM_FULL, v, s));
        return __shfl_sync(FM_FULL, v, 0);
    };
    auto publish = [&](int pos) {
        if (W == 1) return;
        __threadfence_block();                        // our finished bytes before the new mark
        if (lane == 0) owed[warp] = pos;
    };

    // Work items are taken one AHEAD: while item k is copied, the token bits and the output position of the item
    // this warp takes next are already on their way, and the lines of compressed bytes it will read are being
    // pulled into L1.
    int kseq = 0;
    auto take = [&]() -> int {
        int t = 0;
        if (W == 1) t = kseq++;
        else {
            if (lane == 0) t = atomicAdd(&bs->ticket, 1);
            t = __shfl_sync(FM_FULL, t, 0);
        }
        return t;
    };
    uint32_t w_n = 0, co_n = 0;
    auto fetch = [&](int kk) {
        if (kk >= nmac) return;
        const int wi = kk * (C * LZ4_CHUNK_WORDS) + lane;
        w_n = (lane < C * LZ4_CHUNK_WORDS && wi < nwords) ? maps[wi] : 0u;
        const int ci = kk * C + (lane & (C - 1));
        co_n = ci < nchunks ? cops[ci] : 0u;
        const int at = max(kk * (C * LZ4_CHUNK) - dq + lane * 128, 0);
        if (lane < C + 1 && at < csize) asm volatile("prefetch.global.L1 [%0];" ::"l"(src + at));
    };
    int kn = take();
    fetch(kn);
    for (;;) {
        const int k = kn;
        if (k >= nmac) break;
        uint32_t word = w_n;
        const uint32_t co = co_n;
        kn = take();
        fetch(kn);
        const int cnt = __popc(word);
        const int incl_c = warp_incl_scan_add(cnt);
        const int ntok = min(__shfl_sync(FM_FULL, incl_c, 31), LZ4_MAC_TOK);
        if (ntok == 0) continue;
        // output position of the item's first sequence: D1 left it with the first 128-byte chunk that holds a token
        int op0 = (int)__shfl_sync(FM_FULL, co, (__ffs(__ballot_sync(FM_FULL, cnt > 0)) - 1) >> 2);
        for (int base = incl_c - cnt; word; word &= word - 1, base++)
            if (base < LZ4_MAC_TOK) s_tokpos_w[base] = (uint16_t)(lane * 32 + __ffs(word) - 1);
        __syncwarp();
        const int ip_base = k * (C * LZ4_CHUNK) - dq;

        for (int batch = 0; batch < ntok; batch += 32) {
            const bool active = batch + lane < ntok;
            int lit = 0, ml = 0, off = 0, lit_src = 0;
            if (active) {
                int ip = ip_base + (int)s_tokpos_w[batch + lane];
                const unsigned tok = src[ip++];
                lit = (int)(tok >> 4);
                if (lit == 15) { unsigned s; do { s = src[ip++]; lit += (int)s; } while (s == 255); }
                lit_src = ip;
                ip += lit;
                if (ip != csize) {                    // not the closing literal run
                    off = (int)src[ip] | ((int)src[ip + 1] << 8); ip += 2;
                    ml = (int)(tok & 15);
                    if (ml == 15) { unsigned s; do { s = src[ip++]; ml += (int)s; } while (s == 255); }
                    ml += 4;
                }
            }
            const int outlen = lit + ml;
            const int incl = warp_incl_scan_add(outlen);
            const int my_op = op0 + incl - outlen;
            const int total = __shfl_sync(FM_FULL, incl, 31);
            const int d = my_op + lit;                // where my match starts
            const int mstart = d - off;

            if (total <= LZ4_SPAN) {
                // ---------- path A: assemble [op0, op0 + total) in shared memory, flush once ----------
                const int shift = (int)((uintptr_t)(out + op0) & 15);   // equal 16-byte phases in smem and HBM
                uint8_t *sp = span + shift - op0;     // sp[x] holds output byte x
                publish(op0);                         // nothing below op0 is owed by this warp

                // which bytes of my sequence travel as words: [rs, re)
                // (W > 1) a source the sibling warps still owe makes a slow match: it waits there, alone (a short wait
                // here, for all lanes, in the hope of saving that lane the slow way: r02p, 0 / 8 / 32 polls: 9.8 / 10.1 / 10.6 ms)
                const int hwm0 = high_water();
                if (W > 1) __threadfence_block();
                const bool fastm = ml > 0 && off != 0 && mstart + ml <= min(op0, hwm0) && ml <= LZ4_WMATCH;
                const bool slowm = ml > 0 && !fastm;
                // literal words are the aligned words of the stream around the run: all of them inside the payload's own words
                const uint32_t la = (uint32_t)(shift + (my_op - op0));
                const int dd = (int)(la & 3u);
                const int nwl = (dd + lit + 3) >> 2;
                const uintptr_t sb = (uintptr_t)(src + lit_src) - (uintptr_t)dd;      // the stream byte under byte 0 of word 0
                const int wo = (int)(((intptr_t)(sb & ~(uintptr_t)3) - (intptr_t)(uintptr_t)src_w_lo) >> 2);
                const bool lit_words = lit <= LZ4_WLIT && wo >= 0 && wo + nwl <= src_w_last;
                int rs = my_op, re = my_op + outlen;
                if (slowm) re = rs;
                else if (!lit_words) rs = d;
                if (re - rs < 4) { rs = my_op; re = my_op; }          // a word must not be shared by three lanes
                const bool in = re > rs;
                const bool lit_in = in && rs == my_op && lit > 0;
                const bool m_in = in && ml > 0;
                const bool lit_later = lit > 0 && !lit_in;

                // is my first word the left neighbour's last word?  (the same test, seen from the other side, below)
                const int p_re = __shfl_up_sync(FM_FULL, in ? re : -1, 1);
                const bool head_skip = in && lane > 0 && p_re == rs && ((shift + rs - op0) & 3) != 0;
                // match sources: the aligned 16-byte chunks around them
                const int so = (int)((uintptr_t)(out + mstart) & 15);
                const int nch = (so + ml + 15) >> 4;                  // 1..3 when m_in
                // literal words: aligned words of the stream through a funnel shift, bytes outside the run zeroed
                uint32_t l0 = 0, llast = 0;
                const uint32_t lwi = la >> 2;
                if (lit_in) {
                    const int nb = dd + lit;
                    const int sh = (int)(sb & 3) * 8;
                    const uint32_t *sw = src_w_lo + wo;
                    const uint32_t tmask = 0xffffffffu >> (8 * (4 * nwl - nb));
                    // word j is built, word j - 1 stored
                    uint32_t lo = sw[1];
                    uint32_t cur = __funnelshift_r(sw[0], lo, sh) & (0xffffffffu << (8 * dd));
                    if (nwl == 1) cur &= tmask;
                    l0 = cur;
                    for (int j = 1; j < nwl; j++) {
                        const uint32_t hi = sw[j + 1];
                        uint32_t v = __funnelshift_r(lo, hi, sh);
                        lo = hi;
                        if (j == nwl - 1) v &= tmask;
                        if (!(j == 1 && head_skip)) span32[lwi + j - 1] = cur;
                        cur = v;
                    }
                    llast = cur;
                }
                // first match word
                uint32_t mcur = 0, mlo = 0, mwi = 0, m_tmask = 0xffffffffu;
                const uint32_t *msw = (const uint32_t *)s_scr_w;
                int nwm = 0, msh = 0;
                if (m_in) {
                    const uint4 *gb = (const uint4 *)(out + mstart - so);
                    const uint4 z = make_uint4(0, 0, 0, 0);
                    const uint4 q0 = ldg_v4(gb), q1 = nch > 1 ? ldg_v4(gb + 1) : z, q2 = nch > 2 ? ldg_v4(gb + 2) : z;
                    uint4 *scr = (uint4 *)(s_scr_w + lane * LZ4_SCR + 16);
                    scr[0] = q0; if (nch > 1) scr[1] = q1; if (nch > 2) scr[2] = q2;
                    const uint32_t ma = (uint32_t)(shift + (d - op0));
                    const int ddm = (int)(ma & 3u);
                    const int nbm = ddm + ml;
                    nwm = (nbm + 3) >> 2;
                    const uint32_t sbo = (uint32_t)(lane * LZ4_SCR + 16 + so - ddm);  // scratch byte under byte 0 of word 0
                    msh = (int)(sbo & 3u) * 8;
                    msw = (const uint32_t *)(s_scr_w + (sbo & ~3u));
                    const uint32_t b0 = msw[0];
                    mlo = msw[1];
                    mcur = __funnelshift_r(b0, mlo, msh) & (0xffffffffu << (8 * ddm));
                    m_tmask = 0xffffffffu >> (8 * (4 * nwm - nbm));
                    if (nwm == 1) mcur &= m_tmask;
                    mwi = ma >> 2;
                }
                // the word where my literals end and my match begins holds both
                const bool share = lit_in && m_in && mwi == lwi + (uint32_t)nwl - 1u;
                if (share) mcur |= llast;
                else if (lit_in && m_in && !(nwl == 1 && head_skip)) span32[lwi + nwl - 1] = llast;
                const bool head_is_m = m_in && (!lit_in || (share && nwl == 1));
                const uint32_t hv = head_is_m ? mcur : l0;            // my first word, complete
                // match words: word j is built, word j - 1 stored
                for (int j = 1; j < nwm; j++) {
                    const uint32_t hi = msw[j + 1];
                    uint32_t v = __funnelshift_r(mlo, hi, msh);
                    mlo = hi;
                    if (j == nwm - 1) v &= m_tmask;
                    if (!(j == 1 && head_is_m && head_skip)) span32[mwi + j - 1] = mcur;
                    mcur = v;
                }
                // my last word, with the right neighbour's first word folded in when they are the same word
                {
                    const uint32_t n_hv = __shfl_down_sync(FM_FULL, hv, 1);
                    const int n_rs = __shfl_down_sync(FM_FULL, in ? rs : -1, 1);
                    const bool adj = lane < 31 && n_rs == re && ((shift + re - op0) & 3) != 0;
                    if (in) {
                        const uint32_t tv = m_in ? mcur : llast;
                        const uint32_t twi = m_in ? mwi + (uint32_t)nwm - 1u : lwi + (uint32_t)nwl - 1u;
                        span32[twi] = tv | (adj ? n_hv : 0u);
                    }
                }
                __syncwarp();
                // ---- patches, all with byte stores ----
                if (lit_later && lit <= LZ4_WLIT)                  // beside a slow match, or a closing run under 4 bytes
                    copy_batched(sp + my_op, src + lit_src, lit);
                for (unsigned lm = __ballot_sync(FM_FULL, lit > LZ4_WLIT); lm; lm &= lm - 1) {
                    const int l = __ffs(lm) - 1;
                    const int n = __shfl_sync(FM_FULL, lit, l), sp0 = __shfl_sync(FM_FULL, my_op, l);
                    const int ls = __shfl_sync(FM_FULL, lit_src, l);
                    for (int i = lane; i < n; i += 32) sp[sp0 + i] = src[ls + i];
                }
                if (__any_sync(FM_FULL, slowm)) {
                    // Bytes from before the span come from HBM once the block-wide mark has passed them; bytes inside
                    // the span come from shared memory once every earlier slow match that could overlap them is done
                    // (lowest destination first; everything that is not a slow match is complete by now).
                    bool pending = slowm;
                    if (pending && off == 0) {            // lz4.c:2300-2303: offset 0 yields zeros
                        for (int i = 0; i < ml; i++) sp[d + i] = 0;
                        pending = false;
                    }
                    const int ext = (pending && mstart < op0) ? min(ml, op0 - mstart) : 0;
                    const int need_ext = ext ? mstart + ext : 0;
                    const int need_in = (ml > ext) ? min(mstart + ml, d) : 0;
                    __syncwarp();
                    int idle = 0;
                    for (;;) {
                        const int low = warp_min(pending ? d : 0x7fffffff);
                        if (low == 0x7fffffff) break;
                        const int hwm = high_water();
                        if (W > 1) __threadfence_block();
                        const bool go = pending && need_ext <= hwm && need_in <= low;
                        if (go) {
                            if (ext > 0 && ext < LZ4_LONG) {
                                const uint8_t *gs = out + mstart;
                                const int so2 = (int)((uintptr_t)gs & 15);
                                const uint4 *gb = (const uint4 *)(gs - so2);
                                const int nch = (so2 + ext + 15) >> 4;
                                uint4 *scr = (uint4 *)(s_scr_w + lane * LZ4_SCR);
                                const uint4 z = make_uint4(0, 0, 0, 0);
                                const uint4 q0 = ldg_v4(gb), q1 = nch > 1 ? ldg_v4(gb + 1) : z;
                                const uint4 q2 = nch > 2 ? ldg_v4(gb + 2) : z, q3 = nch > 3 ? ldg_v4(gb + 3) : z;
                                scr[0] = q0; if (nch > 1) scr[1] = q1; if (nch > 2) scr[2] = q2; if (nch > 3) scr[3] = q3;
                                const uint8_t *sc = (const uint8_t *)scr + so2;
                                for (int i = 0; i < ext; i++) sp[d + i] = sc[i];
                            }
                            if (ml > ext && ext < LZ4_LONG) smem_copy_seq(sp + d + ext, sp + mstart + ext, ml - ext);
                        }
                        // long reads from HBM: the whole warp per match, then its in-span remainder
                        for (unsigned lm = __ballot_sync(FM_FULL, go && ext >= LZ4_LONG); lm; lm &= lm - 1) {
                            const int l = __ffs(lm) - 1;
                            const int n = __shfl_sync(FM_FULL, ext, l), dd = __shfl_sync(FM_FULL, d, l);
                            const int ms = __shfl_sync(FM_FULL, mstart, l), mm = __shfl_sync(FM_FULL, ml, l);
                            for (int i = lane; i < n; i += 32) sp[dd + i] = out[ms + i];
                            __syncwarp();
                            if (lane == l && mm > n) smem_copy_seq(sp + dd + n, sp + ms + n, mm - n);
                        }
                        const bool any = __any_sync(FM_FULL, go);
                        if (go) pending = false;
                        __syncwarp();
                        if (!any) { if (++idle > 2) __nanosleep(idle > 16 ? 256 : 32); } else idle = 0;
                    }
                }
                __syncwarp();
                // flush
                {
                    const int head = min(total, (16 - shift) & 15);
                    if (lane < head) out[op0 + lane] = span[shift + lane];
                    const int body = (total - head) >> 4;
                    const uint4 *sv = (const uint4 *)(span + shift + head);
                    uint4 *dv = (uint4 *)(out + op0 + head);
                    for (int i = lane; i < body; i += 32) dv[i] = sv[i];
                    const int done = head + (body << 4);
                    if (done + lane < total) out[op0 + done + lane] = span[shift + done + lane];
                }
                op0 += total;
                __syncwarp();
                publish(op0);
                continue;
            }

            // ---------- path B: long sequences, copied in place ----------
            op0 += total;
            if (active && lit < LZ4_LONG) copy_batched(out + my_op, src + lit_src, lit);
            for (unsigned long_m = __ballot_sync(FM_FULL, active && lit >= LZ4_LONG); long_m; long_m &= long_m - 1) {
                const int l = __ffs(long_m) - 1;
                warp_copy_bytes(out + __shfl_sync(FM_FULL, my_op, l), src + __shfl_sync(FM_FULL, lit_src, l),
                                __shfl_sync(FM_FULL, lit, l));
            }
            bool pending = active && ml > 0;
            const int need = (off == 0) ? 0 : min(mstart + ml, d);   // everything below this must exist
            int idle = 0;
            for (;;) {
                const int wmin = warp_min(pending ? d : 0x7fffffff);
                publish(wmin == 0x7fffffff ? op0 : wmin);
                if (wmin == 0x7fffffff) break;
                __syncwarp();
                // alone in the block, everything below our own lowest pending match is complete
                const int hwm = (W == 1) ? wmin : high_water();
                __threadfence_block();
                const bool go = pending && need <= hwm;
                if (go && ml < LZ4_LONG) {
                    if (off == 0) { for (int i = 0; i < ml; i++) out[d + i] = 0; }
                    else { for (int i = 0; i < ml; i++) out[d + i] = out[mstart + i]; }
                }
                for (unsigned long_m = __ballot_sync(FM_FULL, go && ml >= LZ4_LONG); long_m; long_m &= long_m - 1) {
                    const int l = __ffs(long_m) - 1;
                    warp_copy_match(out, __shfl_sync(FM_FULL, d, l), __shfl_sync(FM_FULL, off, l),
                                    __shfl_sync(FM_FULL, ml, l));
                }
                const bool any = __any_sync(FM_FULL, go);
                if (go) pending = false;
                __syncwarp();
                if (!any) { if (++idle > 2) __nanosleep(idle > 16 ? 256 : 32); } else idle = 0;
            }
        }
        __syncwarp();                                 // the token list is rewritten by the next work item
    }
    // a finished warp must not hold the mark down
    publish(0x7fffffff);
}

// ---- D2 when W warps SHARE a block (few blocks: the per-block calls, the slices of the host pipeline, a rank's
// share of a stream on many GPUs).  This is the round-1 design, kept for exactly this case: work items of ONE
// 128-byte chunk (~20 sequences), every match waiting on its own for the block-wide mark, byte-granular assembly.
// A block's matches mostly point a few KiB back -- into the items the sibling warps are still copying -- so what
// counts here is how soon a single item is delivered, not how few instructions it costs: the word-stage kernel
// above with its full batches is 1.7x faster when a warp has a block to itself and 1.1-1.4x SLOWER than this one
// when warps share a block (r02v: 1024 blocks, 4 warps each: 28.2 ms against 19.6 ms; 256 blocks, 8 warps: 10.9
// against 9.5).

// Shared memory of one copy warp / of the W warps that share a block.
struct SharedCopyWarpSmem {
    __align__(16) uint8_t span[LZ4_SPAN + 32];
    __align__(16) uint8_t scr[32 * LZ4_SCR];      // per lane: 4 aligned 16-byte chunks of a match source
    uint8_t tokpos[64];
};

// The copy of one block by W warps (`warp` = 0 .. W-1 within the block).
template <int W>
__device__ __forceinline__ void lz4_copy_block_shared(const BlockDesc &bd, const uint32_t *tokmap, const uint32_t *chunk_op,
                                               const int warp, const int lane, CopyBlockSmem<W> *bs, SharedCopyWarpSmem *ws)
{
    const uint8_t *__restrict__ src = bd.src;
    uint8_t *out = bd.dst;
    const int csize = (int)bd.csize;
    const int dq = (int)((uintptr_t)bd.src & 15);     // D1 records tokens at aligned positions ip + dq
    const int nchunks = (dq + csize + LZ4_CHUNK - 1) / LZ4_CHUNK;
    const uint4 *maps = (const uint4 *)(tokmap + (size_t)bd.chunk_base * LZ4_CHUNK_WORDS);
    const uint32_t *cops = chunk_op + bd.chunk_base;
    volatile int *owed = bs->owed;
    uint8_t *span = ws->span;
    uint8_t *s_tokpos_w = ws->tokpos;
    uint8_t *s_scr_w = ws->scr;

    // lowest output position any warp of this block still owes
    auto high_water = [&]() -> int {
        if (W == 1) return 0x7fffffff;
        int v = lane < W ? owed[lane] : 0x7fffffff;
#pragma unroll
        for (int s = W / 2; s > 0; s >>= 1) v = min(v, __shfl_xor_sync(FM_FULL, v, s));
        return __shfl_sync(FM_FULL, v, 0);
    };
    auto publish = [&](int pos) {
        if (W == 1) return;
        __threadfence_block();                        // our finished bytes before the new mark
        if (lane == 0) owed[warp] = pos;
    };

    // Chunks are taken one AHEAD: while chunk k is copied, the token bits and the output position of the chunk
    // this warp takes next are already on their way, and the two lines of compressed bytes it will read are
    // being pulled into L1 -- a warp's chunks are a dependent chain (bits -> tokens -> offsets -> match bytes),
    // so every load taken off that chain shortens the block's time.
    int kseq = 0;
    auto take = [&]() -> int {
        int t = 0;
        if (W == 1) t = kseq++;
        else {
            if (lane == 0) t = atomicAdd(&bs->ticket, 1);
            t = __shfl_sync(FM_FULL, t, 0);
        }
        return t;
    };
    uint4 m_n = make_uint4(0u, 0u, 0u, 0u);
    int op_n = 0;
    auto fetch = [&](int kk) {
        if (kk >= nchunks) return;
        m_n = maps[kk]; op_n = (int)cops[kk];
        const int at = kk * LZ4_CHUNK - dq + lane * 128;
        if (lane < 2 && at >= 0 && at < csize) asm volatile("prefetch.global.L1 [%0];" ::"l"(src + at));
    };
    int kn = take();
    fetch(kn);
    for (;;) {
        const int k = kn;
        if (k >= nchunks) break;
        const uint4 m = m_n;
        int op0 = op_n;
        kn = take();
        fetch(kn);
        const int c0 = __popc(m.x), c1 = c0 + __popc(m.y), c2 = c1 + __popc(m.z), ntok = c2 + __popc(m.w);
        if (ntok == 0) continue;

        // lane r takes the r-th token of the chunk: scatter positions by rank
        {
            const uint32_t lt = (1u << lane) - 1u;
            if ((m.x >> lane) & 1u) s_tokpos_w[__popc(m.x & lt)] = (uint8_t)lane;
            if ((m.y >> lane) & 1u) s_tokpos_w[c0 + __popc(m.y & lt)] = (uint8_t)(32 + lane);
            if ((m.z >> lane) & 1u) s_tokpos_w[c1 + __popc(m.z & lt)] = (uint8_t)(64 + lane);
            if ((m.w >> lane) & 1u) s_tokpos_w[c2 + __popc(m.w & lt)] = (uint8_t)(96 + lane);
        }
        __syncwarp();

        for (int batch = 0; batch < ntok; batch += 32) {
            const bool active = batch + lane < ntok;
            int lit = 0, ml = 0, off = 0, lit_src = 0;
            if (active) {
                int ip = k * LZ4_CHUNK + (int)s_tokpos_w[batch + lane] - dq;
                const unsigned tok = src[ip++];
                lit = (int)(tok >> 4);
                if (lit == 15) { unsigned s; do { s = src[ip++]; lit += (int)s; } while (s == 255); }
                lit_src = ip;
                ip += lit;
                if (ip != csize) {                    // not the closing literal run
                    off = (int)src[ip] | ((int)src[ip + 1] << 8); ip += 2;
                    ml = (int)(tok & 15);
                    if (ml == 15) { unsigned s; do { s = src[ip++]; ml += (int)s; } while (s == 255); }
                    ml += 4;
                }
            }
            const int outlen = lit + ml;
            const int incl = warp_incl_scan_add(outlen);
            const int my_op = op0 + incl - outlen;
            const int total = __shfl_sync(FM_FULL, incl, 31);
            const int d = my_op + lit;                // where my match starts
            const int mstart = d - off;

            if (total <= LZ4_SPAN) {
                // ---------- path A: assemble [op0, op0 + total) in shared memory, flush once ----------
                const int shift = (int)((uintptr_t)(out + op0) & 15);   // equal 16-byte phases in smem and HBM
                uint8_t *sp = span + shift - op0;     // sp[x] holds output byte x
                publish(op0);                         // nothing below op0 is owed by this warp
                // literals
                if (active && lit < LZ4_LONG) copy_batched(sp + my_op, src + lit_src, lit);
                for (unsigned lm = __ballot_sync(FM_FULL, active && lit >= LZ4_LONG); lm; lm &= lm - 1) {
                    const int l = __ffs(lm) - 1;
                    const int n = __shfl_sync(FM_FULL, lit, l), sp0 = __shfl_sync(FM_FULL, my_op, l);
                    const int ls = __shfl_sync(FM_FULL, lit_src, l);
                    for (int i = lane; i < n; i += 32) sp[sp0 + i] = src[ls + i];
                }
                // matches.  Bytes from before the span come from HBM once the block-wide mark has
                // passed them; bytes inside the span come from shared memory once every earlier
                // match of this warp that could overlap them is done (lowest destination first).
                bool pending = ml > 0;
                if (pending && off == 0) {            // lz4.c:2300-2303: offset 0 yields zeros
                    for (int i = 0; i < ml; i++) sp[d + i] = 0;
                    pending = false;
                }
                const int ext = (pending && mstart < op0) ? min(ml, op0 - mstart) : 0;
                const int need_ext = ext ? mstart + ext : 0;
                const int need_in = (ml > ext) ? min(mstart + ml, d) : 0;
                __syncwarp();
                int idle = 0;
                for (;;) {
                    const int low = warp_min(pending ? d : 0x7fffffff);
                    if (low == 0x7fffffff) break;
                    const int hwm = high_water();
                    if (W > 1) __threadfence_block();
                    const bool go = pending && need_ext <= hwm && need_in <= low;
                    if (go) {
                        if (ext > 0 && ext < LZ4_LONG) {
                            // the source is somewhere in HBM/L2: fetch it with (at most four) aligned
                            // 128-bit loads instead of one load per byte -- every load of a scattered
                            // address costs the L1 data pipe a wavefront per lane
                            const uint8_t *gs = out + mstart;
                            const int so = (int)((uintptr_t)gs & 15);
                            const uint4 *gb = (const uint4 *)(gs - so);
                            const int nch = (so + ext + 15) >> 4;
                            uint4 *scr = (uint4 *)(s_scr_w + lane * LZ4_SCR);
                            const uint4 z = make_uint4(0, 0, 0, 0);
                            const uint4 q0 = ldg_v4(gb), q1 = nch > 1 ? ldg_v4(gb + 1) : z;
                            const uint4 q2 = nch > 2 ? ldg_v4(gb + 2) : z, q3 = nch > 3 ? ldg_v4(gb + 3) : z;
                            scr[0] = q0; if (nch > 1) scr[1] = q1; if (nch > 2) scr[2] = q2; if (nch > 3) scr[3] = q3;
                            copy_batched(sp + d, (const uint8_t *)scr + so, ext);
                        }
                        if (ml > ext && ext < LZ4_LONG) smem_copy_seq(sp + d + ext, sp + mstart + ext, ml - ext);
                    }
                    // long reads from HBM: the whole warp per match, then its in-span remainder
                    for (unsigned lm = __ballot_sync(FM_FULL, go && ext >= LZ4_LONG); lm; lm &= lm - 1) {
                        const int l = __ffs(lm) - 1;
                        const int n = __shfl_sync(FM_FULL, ext, l), dd = __shfl_sync(FM_FULL, d, l);
                        const int ms = __shfl_sync(FM_FULL, mstart, l), mm = __shfl_sync(FM_FULL, ml, l);
                        for (int i = lane; i < n; i += 32) sp[dd + i] = out[ms + i];
                        __syncwarp();
                        if (lane == l && mm > n) smem_copy_seq(sp + dd + n, sp + ms + n, mm - n);
                    }
                    const bool any = __any_sync(FM_FULL, go);
                    if (go) pending = false;
                    __syncwarp();
                    if (!any) { if (++idle > 2) __nanosleep(idle > 16 ? 256 : 32); } else idle = 0;
                }
                __syncwarp();
                // flush
                {
                    const int head = min(total, (16 - shift) & 15);
                    if (lane < head) out[op0 + lane] = span[shift + lane];
                    const int body = (total - head) >> 4;
                    const uint4 *sv = (const uint4 *)(span + shift + head);
                    uint4 *dv = (uint4 *)(out + op0 + head);
                    for (int i = lane; i < body; i += 32) dv[i] = sv[i];
                    const int done = head + (body << 4);
                    if (done + lane < total) out[op0 + done + lane] = span[shift + done + lane];
                }
                op0 += total;
                __syncwarp();
                publish(op0);
                continue;
            }

            // ---------- path B: long sequences, copied in place ----------
            op0 += total;
            if (active && lit < LZ4_LONG) copy_batched(out + my_op, src + lit_src, lit);
            for (unsigned long_m = __ballot_sync(FM_FULL, active && lit >= LZ4_LONG); long_m; long_m &= long_m - 1) {
                const int l = __ffs(long_m) - 1;
                warp_copy_bytes(out + __shfl_sync(FM_FULL, my_op, l), src + __shfl_sync(FM_FULL, lit_src, l),
                                __shfl_sync(FM_FULL, lit, l));
            }
            bool pending = active && ml > 0;
            const int need = (off == 0) ? 0 : min(mstart + ml, d);   // everything below this must exist
            int idle = 0;
            for (;;) {
                const int wmin = warp_min(pending ? d : 0x7fffffff);
                publish(wmin == 0x7fffffff ? op0 : wmin);
                if (wmin == 0x7fffffff) break;
                __syncwarp();
                // alone in the block, everything below our own lowest pending match is complete
                const int hwm = (W == 1) ? wmin : high_water();
                __threadfence_block();
                const bool go = pending && need <= hwm;
                if (go && ml < LZ4_LONG) {
                    if (off == 0) { for (int i = 0; i < ml; i++) out[d + i] = 0; }
                    else { for (int i = 0; i < ml; i++) out[d + i] = out[mstart + i]; }
                }
                for (unsigned long_m = __ballot_sync(FM_FULL, go && ml >= LZ4_LONG); long_m; long_m &= long_m - 1) {
                    const int l = __ffs(long_m) - 1;
                    warp_copy_match(out, __shfl_sync(FM_FULL, d, l), __shfl_sync(FM_FULL, off, l),
                                    __shfl_sync(FM_FULL, ml, l));
                }
                const bool any = __any_sync(FM_FULL, go);
                if (go) pending = false;
                __syncwarp();
                if (!any) { if (++idle > 2) __nanosleep(idle > 16 ? 256 : 32); } else idle = 0;
            }
        }
        __syncwarp();
    }
    // a finished warp must not hold the mark down
    publish(0x7fffffff);
}


// D2 as a kernel of its own (D1 ran before it): one CTA of W warps per block.
// W = warps per block (1, 2, 4 or 8): the host picks it from the batch size -- many blocks in flight need few warps
// each (and then hardly ever wait on one another), few blocks need many.  More than 8 only wait for one another: a
// chunk's matches mostly point into the last few KiB, i.e. into the chunks the sibling warps still work on
// (r02i, one block: 8 warps 8.5 ms, 32 warps 12.5 ms).
// A warp alone with its block: 28 CTAs per SM at 72 registers (r02q: forced down to 64 registers for 32 CTAs per SM,
// 42 -> 47 ms per 16 GiB).
// BYTES = true: the byte-granular copy (items of one chunk, W > 1 only).
template <int W, bool BYTES>
__global__ void __launch_bounds__(W * 32, W == 1 ? 28 : 1)
lz4_copy_kernel(const BlockDesc *blocks, const uint32_t *tokmap, const uint32_t *chunk_op,
                const int32_t *result)
{
    __shared__ CopyBlockSmem<W> s_block;
    const BlockDesc bd = blocks[blockIdx.x];
    if (bd.stored || result[blockIdx.x] < 0) return;
    if (threadIdx.x < W) s_block.owed[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_block.ticket = 0;
    if (W > 1) __syncthreads();
    if constexpr (!BYTES) {
        __shared__ CopyWarpSmem s_warp[W];
        lz4_copy_block<W, (W == 1 ? LZ4_MAC_CHUNKS : 1)>(bd, tokmap, chunk_op, threadIdx.x >> 5, threadIdx.x & 31, &s_block, &s_warp[threadIdx.x >> 5]);
    } else {
        __shared__ SharedCopyWarpSmem s_warp[W];
        lz4_copy_block_shared<W>(bd, tokmap, chunk_op, threadIdx.x >> 5, threadIdx.x & 31, &s_block, &s_warp[threadIdx.x >> 5]);
    }
}

// ---- D0 ------------------------------------------------------------------------------------

__global__ void lz4_stored_kernel(const BlockDesc *blocks, uint32_t n_blocks)
{
    // grid.y = block index, grid.x tiles the payload
    const BlockDesc bd = blocks[blockIdx.y];
    if (bd.stored != 1) return;                           // 2 = checksum-only item (a footer): nothing to copy
    const uint32_t n = bd.csize;
    const uint8_t *__restrict__ s = bd.src;
    uint8_t *d = bd.dst;
    const uint32_t head = min(n, (uint32_t)((16 - ((uintptr_t)d & 15)) & 15));
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (((uintptr_t)(s + head) & 3) == 0) {
        // destination 16-byte aligned after `head`, source word aligned: vector body
        for (uint32_t i = tid; i < head; i += nth) d[i] = s[i];
        const uint32_t body = (n - head) >> 4;
        const uint32_t *sw = (const uint32_t *)(s + head);
        uint4 *dv = (uint4 *)(d + head);
        for (uint32_t i = tid; i < body; i += nth)
            dv[i] = make_uint4(sw[4 * i], sw[4 * i + 1], sw[4 * i + 2], sw[4 * i + 3]);
        for (uint32_t i = head + (body << 4) + tid; i < n; i += nth) d[i] = s[i];
    } else {
        const uint32_t sh = ((uintptr_t)(s + head) & 3) * 8;
        const uint32_t *sw = (const uint32_t *)((uintptr_t)(s + head) & ~(uintptr_t)3);
        for (uint32_t i = tid; i < head; i += nth) d[i] = s[i];
        // keep one word of slack at the end: the funnel reads sw[4i+4]
        const uint32_t body = (n - head) >= 20 ? ((n - head - 4) >> 4) : 0;
        uint4 *dv = (uint4 *)(d + head);
        for (uint32_t i = tid; i < body; i += nth) {
            const uint32_t a = sw[4 * i], b = sw[4 * i + 1], c = sw[4 * i + 2], e = sw[4 * i + 3], f = sw[4 * i + 4];
            dv[i] = make_uint4(__funnelshift_r(a, b, sh), __funnelshift_r(b, c, sh),
                               __funnelshift_r(c, e, sh), __funnelshift_r(e, f, sh));
        }
        for (uint32_t i = head + (body << 4) + tid; i < n; i += nth) d[i] = s[i];
    }
}

}  // namespace fm
