// lz4_decode.cuh -- LZ4 block decode kernels (reference: LZ4_decompress_safe,
// native/lz4/lz4.c:2345, called per block at native/4mc.c:661 and native/jniDecompressor.c:88).
//
//  D1  lz4_parse_kernel   one THREAD per block walks the token chain (lz4_parse.h), decides the
//                         return value, and leaves two small side tables in HBM:
//                           tokmap   1 bit per compressed byte, set where a sequence's token sits
//                           chunk_op u32 per 128 compressed bytes: output position of the first
//                                    sequence whose token lies in that 128-byte chunk
//                         (0.16 bytes of scratch per compressed byte).
//  D2  lz4_copy_kernel    one CTA per block.  Warps take 128-byte chunks of the compressed
//                         stream in order (a shared ticket), turn the chunk's <= 43 token bits
//                         into one sequence per lane, prefix-sum the output lengths, and copy.
//                         Literals have no dependencies.  A match may read bytes produced by an
//                         earlier sequence, so warps publish the lowest output position they still
//                         owe (`owed[w]`, shared memory) and a match is copied once everything
//                         below the end of its source is below the minimum over all warps
//                         (multi-round resolution, generalised from one warp to the CTA).
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

__global__ void lz4_parse_kernel(const BlockDesc *blocks, uint32_t n_blocks,
                                 uint32_t *tokmap, uint32_t *chunk_op, int32_t *result)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const BlockDesc bd = blocks[b];
    if (bd.stored) { result[b] = (int32_t)bd.usize; return; }
    TokSink sink;
    sink.tokmap = tokmap + (size_t)bd.chunk_base * LZ4_CHUNK_WORDS;
    sink.chunk_op = chunk_op + bd.chunk_base;
    sink.cur_word = 0; sink.bits = 0; sink.cur_chunk = 0xffffffffu;
    const int r = lz4_parse_block(bd.src, (int)bd.csize, (int)bd.usize, sink);
    sink.flush();
    result[b] = r;
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

__global__ void __launch_bounds__(LZ4_COPY_WARPS * 32)
lz4_copy_kernel(const BlockDesc *blocks, const uint32_t *tokmap, const uint32_t *chunk_op,
                const int32_t *result)
{
    __shared__ int s_owed[LZ4_COPY_WARPS];
    __shared__ int s_ticket;
    __shared__ uint8_t s_tokpos[LZ4_COPY_WARPS][64];

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
            op0 += __shfl_sync(FM_FULL, incl, 31);

            // literals: independent of everything else
            if (active && lit < LZ4_LONG)
                for (int i = 0; i < lit; i++) out[my_op + i] = src[lit_src + i];
            for (unsigned long_m = __ballot_sync(FM_FULL, active && lit >= LZ4_LONG); long_m; long_m &= long_m - 1) {
                const int l = __ffs(long_m) - 1;
                warp_copy_bytes(out + __shfl_sync(FM_FULL, my_op, l), src + __shfl_sync(FM_FULL, lit_src, l),
                                __shfl_sync(FM_FULL, lit, l));
            }

            // matches: multi-round resolution against the CTA-wide high-water mark
            const int d = my_op + lit;
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
