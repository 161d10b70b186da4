// lz4_decode.cuh -- LZ4 block decode kernels (reference: LZ4_decompress_safe,
// native/lz4/lz4.c:2345, called per block at native/4mc.c:661 and native/jniDecompressor.c:88).
//
//  D1  lz4_parse_kernel   one WARP per block: 32 lanes stream the payload through a shared-memory
//                         ring, lane 0 walks the token chain (lz4_parse.h), decides the return
//                         value, and leaves two small side tables in HBM:
//                           tokmap   1 bit per compressed byte, set where a sequence's token sits
//                           chunk_op u32 per 128 compressed bytes: output position of the first
//                                    sequence whose token lies in that 128-byte chunk
//                         (0.16 bytes of scratch per compressed byte).
//  D2  lz4_copy_kernel    one CTA per block.  Warps take 128-byte chunks of the compressed
//                         stream in order (a shared ticket), turn the chunk's <= 43 token bits
//                         into one sequence per lane and prefix-sum the output lengths.  The
//                         chunk's output span (~250 B on text) is assembled in shared memory --
//                         literals from the stream, match bytes from HBM when the source precedes
//                         the span, from the span itself otherwise (resolved in rounds, lowest
//                         destination first) -- and flushed with 128-bit stores.  A source in HBM
//                         may belong to a chunk another warp is still working on, so warps publish
//                         the lowest output position they still owe (`owed[w]`) and wait until the
//                         minimum over all warps has passed the bytes they need.  Spans over 2 KiB
//                         (long runs) are copied in place instead, by the whole warp.
//  D0  lz4_stored_kernel  csize == usize blocks are raw copies (native/4mc.c:635-642).
#pragma once

#include "fm_common.cuh"
#include "lz4_parse.h"

namespace fm {

constexpr int LZ4_CHUNK = 128;            // compressed bytes per D2 work item
constexpr int LZ4_CHUNK_WORDS = 4;        // tokmap words per chunk
constexpr int LZ4_MAX_TOK = 43;           // ceil(128 / 3): a sequence is at least 3 bytes
constexpr int LZ4_COPY_WARPS = 8;
constexpr int LZ4_LONG = 48;              // copies this long are done by the whole warp

struct BlockDesc {                         // one per block, built on the device
    const uint8_t *src;                    // payload
    uint8_t *dst;
    uint32_t csize, usize;                 // usize = capacity offered to the decoder
    uint32_t chunk_base;                   // first chunk index in tokmap/chunk_op
    uint32_t stored;                       // 1: raw copy
};

// ---- D1 ------------------------------------------------------------------------------------

struct TokSink {
    uint32_t *tokmap;      // block's first word
    uint32_t *chunk_op;    // block's first entry
    uint32_t cur_word;     // index of the word being accumulated
    uint32_t bits;
    uint32_t cur_chunk;
    __device__ __forceinline__ void token(int pos, int op)
    {
        const uint32_t w = (uint32_t)pos >> 5;
        if (w != cur_word) {
            if (bits) tokmap[cur_word] = bits;
            cur_word = w; bits = 0;
        }
        bits |= 1u << (pos & 31);
        const uint32_t c = (uint32_t)pos >> 7;
        if (c != cur_chunk) { chunk_op[c] = (uint32_t)op; cur_chunk = c; }
    }
    __device__ __forceinline__ void flush() { if (bits) tokmap[cur_word] = bits; }
};

constexpr int D1_HALF = 2048;             // bytes per ring half
constexpr int D1_RING = 2 * D1_HALF;
constexpr int D1_WARPS = 4;               // blocks per CTA

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

// One WARP per block: all lanes stream the payload into a double-buffered ring (coalesced 128-bit
// loads, the next half in flight while the current one is parsed), lane 0 walks the chain.
__global__ void __launch_bounds__(D1_WARPS * 32)
lz4_parse_kernel(const BlockDesc *blocks, uint32_t n_blocks, uint32_t *tokmap, uint32_t *chunk_op, int32_t *result)
{
    __shared__ __align__(16) uint8_t s_ring[D1_WARPS][D1_RING];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * D1_WARPS + warp;
    if (b >= n_blocks) return;
    const BlockDesc bd = blocks[b];
    if (bd.stored) { if (lane == 0) result[b] = (int32_t)bd.usize; return; }

    uint8_t *ring = s_ring[warp];
    const uintptr_t a = (uintptr_t)bd.src;
    const uint4 *base = (const uint4 *)(a & ~(uintptr_t)15);
    const int d = (int)(a & 15);
    const int csize = (int)bd.csize;
    const int nchunks = (d + csize + 15) >> 4;            // 16-byte chunks holding payload bytes
    constexpr int HC = D1_HALF / 16;                      // chunks per half

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
    sink.tokmap = tokmap + (size_t)bd.chunk_base * LZ4_CHUNK_WORDS;
    sink.chunk_op = chunk_op + bd.chunk_base;
    sink.cur_word = 0; sink.bits = 0; sink.cur_chunk = 0xffffffffu;
    lz4_parse_init(st, bd.src == nullptr, csize, (int)bd.usize,
                   (lane == 0 && bd.src != nullptr && csize > 0) ? (unsigned)bd.src[0] : 0u);

    if (!st.status) {
        load_half(0); store_half(0);
        load_half(1); store_half(1);
        __syncwarp();
        for (int k = 0;; k++) {
            const bool more = (k + 2) * HC < nchunks;
            if (more) load_half(k + 2);                   // in flight during the walk below
            if (lane == 0 && !st.status) {
                RingReader rd;
                rd.ring = ring; rd.src = bd.src; rd.d = d;
                rd.lo = k * D1_HALF - d;
                rd.span = min((k + 2) * D1_HALF - d, csize) - rd.lo;
                // tokens below the middle of the window keep their next 32 bytes staged
                const int stop = more ? (k + 1) * D1_HALF - d : 0x7fffffff;
                lz4_parse_run(st, rd, sink, csize, (int)bd.usize, stop);
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

__global__ void __launch_bounds__(LZ4_COPY_WARPS * 32)
lz4_copy_kernel(const BlockDesc *blocks, const uint32_t *tokmap, const uint32_t *chunk_op,
                const int32_t *result)
{
    __shared__ int s_owed[LZ4_COPY_WARPS];
    __shared__ int s_ticket;
    __shared__ uint8_t s_tokpos[LZ4_COPY_WARPS][64];
    __shared__ __align__(16) uint8_t s_span[LZ4_COPY_WARPS][LZ4_SPAN + 32];

    const BlockDesc bd = blocks[blockIdx.x];
    if (bd.stored || result[blockIdx.x] < 0) return;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint8_t *__restrict__ src = bd.src;
    uint8_t *out = bd.dst;
    const int csize = (int)bd.csize;
    const int nchunks = (csize + LZ4_CHUNK - 1) / LZ4_CHUNK;
    const uint4 *maps = (const uint4 *)(tokmap + (size_t)bd.chunk_base * LZ4_CHUNK_WORDS);
    const uint32_t *cops = chunk_op + bd.chunk_base;
    volatile int *owed = s_owed;
    uint8_t *span = s_span[warp];

    if (threadIdx.x < LZ4_COPY_WARPS) s_owed[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_ticket = 0;
    __syncthreads();

    for (;;) {
        int k = 0;
        if (lane == 0) k = atomicAdd(&s_ticket, 1);
        k = __shfl_sync(FM_FULL, k, 0);
        if (k >= nchunks) break;

        const uint4 m = maps[k];
        const int c0 = __popc(m.x), c1 = c0 + __popc(m.y), c2 = c1 + __popc(m.z), ntok = c2 + __popc(m.w);
        if (ntok == 0) continue;
        int op0 = (int)cops[k];

        // lane r takes the r-th token of the chunk: scatter positions by rank
        {
            const uint32_t lt = (1u << lane) - 1u;
            if ((m.x >> lane) & 1u) s_tokpos[warp][__popc(m.x & lt)] = (uint8_t)lane;
            if ((m.y >> lane) & 1u) s_tokpos[warp][c0 + __popc(m.y & lt)] = (uint8_t)(32 + lane);
            if ((m.z >> lane) & 1u) s_tokpos[warp][c1 + __popc(m.z & lt)] = (uint8_t)(64 + lane);
            if ((m.w >> lane) & 1u) s_tokpos[warp][c2 + __popc(m.w & lt)] = (uint8_t)(96 + lane);
        }
        __syncwarp();

        for (int batch = 0; batch < ntok; batch += 32) {
            const bool active = batch + lane < ntok;
            int lit = 0, ml = 0, off = 0, lit_src = 0;
            if (active) {
                int ip = k * LZ4_CHUNK + (int)s_tokpos[warp][batch + lane];
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

            if (total <= LZ4_SPAN) {
                // ---------- path A: assemble [op0, op0 + total) in shared memory, flush once ----------
                const int shift = (int)((uintptr_t)(out + op0) & 15);   // equal 16-byte phases in smem and HBM
                uint8_t *sp = span + shift - op0;     // sp[x] holds output byte x
                __threadfence_block();
                if (lane == 0) owed[warp] = op0;      // nothing below op0 is owed by this warp
                // literals
                if (active && lit < LZ4_LONG)
                    for (int i = 0; i < lit; i++) sp[my_op + i] = src[lit_src + i];
                for (unsigned lm = __ballot_sync(FM_FULL, active && lit >= LZ4_LONG); lm; lm &= lm - 1) {
                    const int l = __ffs(lm) - 1;
                    const int n = __shfl_sync(FM_FULL, lit, l), sp0 = __shfl_sync(FM_FULL, my_op, l);
                    const int ls = __shfl_sync(FM_FULL, lit_src, l);
                    for (int i = lane; i < n; i += 32) sp[sp0 + i] = src[ls + i];
                }
                // bytes a match takes from before the span come from HBM, once they exist
                const int mstart = d - off;
                const int ext = (ml > 0 && off > 0 && mstart < op0) ? min(ml, op0 - mstart) : 0;
                int need = ext ? mstart + ext : 0;
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) need = max(need, __shfl_xor_sync(FM_FULL, need, s));
                if (need > 0) {
                    for (;;) {
                        const int hwm = warp_min(lane < LZ4_COPY_WARPS ? owed[lane] : 0x7fffffff);
                        if (hwm >= need) break;
                    }
                    __threadfence_block();
                }
                if (ext > 0) {
                    if (ext < LZ4_LONG) { for (int i = 0; i < ext; i++) sp[d + i] = out[mstart + i]; }
                }
                for (unsigned lm = __ballot_sync(FM_FULL, ext >= LZ4_LONG); lm; lm &= lm - 1) {
                    const int l = __ffs(lm) - 1;
                    const int n = __shfl_sync(FM_FULL, ext, l), dd = __shfl_sync(FM_FULL, d, l);
                    const int ms = __shfl_sync(FM_FULL, mstart, l);
                    for (int i = lane; i < n; i += 32) sp[dd + i] = out[ms + i];
                }
                __syncwarp();
                // the rest of every match reads bytes of this span: resolve in rounds, lowest first
                bool pending = ml > ext;
                if (ml > 0 && off == 0) {             // lz4.c:2300-2303: offset 0 yields zeros
                    for (int i = 0; i < ml; i++) sp[d + i] = 0;
                    pending = false;
                }
                const int need_in = min(mstart + ml, d);          // everything below this must exist
                for (;;) {
                    const int low = warp_min(pending ? d : 0x7fffffff);
                    if (low == 0x7fffffff) break;
                    if (pending && need_in <= low) {
                        smem_copy_seq(sp + d + ext, sp + mstart + ext, ml - ext);
                        pending = false;
                    }
                    __syncwarp();
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
                __threadfence_block();
                __syncwarp();
                if (lane == 0) owed[warp] = op0;
                continue;
            }

            // ---------- path B: long sequences, copied in place ----------
            op0 += total;
            if (active && lit < LZ4_LONG)
                for (int i = 0; i < lit; i++) out[my_op + i] = src[lit_src + i];
            for (unsigned long_m = __ballot_sync(FM_FULL, active && lit >= LZ4_LONG); long_m; long_m &= long_m - 1) {
                const int l = __ffs(long_m) - 1;
                warp_copy_bytes(out + __shfl_sync(FM_FULL, my_op, l), src + __shfl_sync(FM_FULL, lit_src, l),
                                __shfl_sync(FM_FULL, lit, l));
            }
            bool pending = active && ml > 0;
            const int need = (off == 0) ? 0 : min(d - off + ml, d);   // everything below this must exist
            for (;;) {
                const int wmin = warp_min(pending ? d : 0x7fffffff);
                __threadfence_block();                // our finished bytes before the new mark
                if (lane == 0) owed[warp] = (wmin == 0x7fffffff) ? op0 : wmin;
                if (wmin == 0x7fffffff) break;
                __syncwarp();
                const int hwm = warp_min(lane < LZ4_COPY_WARPS ? owed[lane] : 0x7fffffff);
                __threadfence_block();
                const bool go = pending && need <= hwm;
                if (go && ml < LZ4_LONG) {
                    if (off == 0) { for (int i = 0; i < ml; i++) out[d + i] = 0; }
                    else { for (int i = 0; i < ml; i++) out[d + i] = out[d - off + i]; }
                }
                for (unsigned long_m = __ballot_sync(FM_FULL, go && ml >= LZ4_LONG); long_m; long_m &= long_m - 1) {
                    const int l = __ffs(long_m) - 1;
                    warp_copy_match(out, __shfl_sync(FM_FULL, d, l), __shfl_sync(FM_FULL, off, l),
                                    __shfl_sync(FM_FULL, ml, l));
                }
                if (go) pending = false;
            }
        }
        __syncwarp();
    }
    // a finished warp must not hold the mark down
    __threadfence_block();
    if (lane == 0) owed[warp] = 0x7fffffff;
}

// ---- D0 ------------------------------------------------------------------------------------

__global__ void lz4_stored_kernel(const BlockDesc *blocks, uint32_t n_blocks)
{
    // grid.y = block index, grid.x tiles the payload
    const BlockDesc bd = blocks[blockIdx.y];
    if (!bd.stored) return;
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
