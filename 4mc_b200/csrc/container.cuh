// container.cuh -- 4mc container bookkeeping on the device.
//
// Writer side (native/4mc.c:263-274 header, :285-293 block offsets, :335-362 EOS + footer):
//   scan_lens_kernel    exclusive scan of the block record lengths (12 + csize) -> absolute offsets
//   write_index_kernel  file header, EOS mark, footer = size, version, delta[], size, magic, XXH32
// Reader side (native/4mc.c:577-585 header, :603-668 block headers, :670-688 footer;
// FourMcInputStream.java:163-239 footer index):
//   read_index_kernel   validates header + footer, prefix-sums the deltas, reads every block header
//                       and cross-checks it against the index, builds the decoder's BlockDesc table
//   xxh_verify_kernel   XXH32 of every payload against its header checksum (:637/:645)
//   finalize_kernel     first failing block in stream order decides the result, like the serial loop
#pragma once

#include "fm_common.cuh"
#include "lz4_decode.cuh"
#include "xxh32.cuh"

namespace fm {

constexpr int SCAN_THREADS = 1024;

// CTA-wide inclusive scan of one value per thread (SCAN_THREADS threads); tmp holds 32 entries.
__device__ __forceinline__ unsigned long long cta_incl_scan_u64(unsigned long long v, unsigned long long *tmp,
                                                                unsigned long long *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(FM_FULL, v, d);
        if (lane >= d) v += t;
    }
    if (lane == 31) tmp[warp] = v;
    __syncthreads();
    unsigned long long w = tmp[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(FM_FULL, w, d);
        if (lane >= d) w += t;
    }
    const unsigned long long base = warp ? __shfl_sync(FM_FULL, w, warp - 1) : 0ull;
    *total = __shfl_sync(FM_FULL, w, 31);
    __syncthreads();
    return v + base;
}

// block_off[b] = base + sum_{i<b} lens[i];  *span = sum of all lens.   One CTA.
__global__ void __launch_bounds__(SCAN_THREADS)
scan_lens_kernel(const uint32_t *lens, uint32_t n_blocks, uint64_t base, uint64_t *block_off, uint64_t *span)
{
    __shared__ unsigned long long tmp[32];
    unsigned long long carry = 0;
    for (uint32_t i0 = 0; i0 < n_blocks; i0 += SCAN_THREADS) {
        const uint32_t i = i0 + threadIdx.x;
        const unsigned long long v = i < n_blocks ? lens[i] : 0;
        unsigned long long total;
        const unsigned long long incl = cta_incl_scan_u64(v, tmp, &total);
        if (i < n_blocks) block_off[i] = base + carry + incl - v;
        carry += total;
    }
    if (threadIdx.x == 0) *span = carry;
}

// header (12 bytes, may be NULL) and tail = EOS (12 zero bytes) + footer (20 + 4n bytes).
// first_delta is the absolute offset of block 0 (12 for a stream that starts at file offset 0).
// When tail_off is not NULL the tail goes to tail + *tail_off (the span length the scan produced),
// and *total (optional) receives header + span + EOS + footer.
__global__ void __launch_bounds__(SCAN_THREADS)
write_index_kernel(const uint32_t *lens, uint32_t n_blocks, uint32_t first_delta, uint32_t magic,
                   uint8_t *header, uint8_t *tail, const uint64_t *tail_off, uint64_t *total)
{
    __shared__ __align__(16) uint32_t s_stage[XXH_WARP_SMEM_WORDS];
    const uint32_t fsize = 20 + 4 * n_blocks;
    if (tail_off) tail += *tail_off;
    if (total && threadIdx.x == 0) *total = 12ull + (tail_off ? *tail_off : 0ull) + 12ull + fsize;
    uint8_t *foot = tail + 12;
    if (threadIdx.x == 0) {
        if (header) {
            st_be32(header, magic); st_be32(header + 4, FOURMC_VERSION);
            st_be32(header + 8, xxh32_thread(header, 8, 0));
        }
        for (int i = 0; i < 12; i++) tail[i] = 0;
        st_be32(foot, fsize); st_be32(foot + 4, 1);
        st_be32(foot + 8 + 4 * n_blocks, fsize);
        st_be32(foot + 12 + 4 * n_blocks, magic);
    }
    for (uint32_t i = threadIdx.x; i < n_blocks; i += SCAN_THREADS)
        st_be32(foot + 8 + 4 * i, i == 0 ? first_delta : lens[i - 1]);
    __threadfence_block();
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t h = xxh32_warp<false>(foot, fsize - 4, 0, s_stage);
        if (threadIdx.x == 0) st_be32(foot + fsize - 4, h);
    }
}

// ---- reader ----------------------------------------------------------------------------------

struct IndexInfo {              // device-side summary of one parsed stream
    int32_t status;             // FOURMC_OK or FOURMC_E_*
    uint32_t n_blocks;
    uint64_t total_usize;
    uint32_t total_chunks;
    uint32_t pad;
};

// One CTA.  n_blocks_host is what the host derived from the footer size field; the kernel
// re-derives it and fails if they disagree.
__global__ void __launch_bounds__(SCAN_THREADS)
read_index_kernel(const uint8_t *in, uint64_t n, uint32_t n_blocks_host, uint8_t *out, uint64_t out_cap,
                  BlockDesc *desc, uint32_t *xxh_expect, uint8_t *status, IndexInfo *info, uint32_t magic,
                  uint32_t first = 0, uint32_t count = 0xffffffffu)
{
    // [first, first + count): the block range the caller decodes (a rank's share of the stream).
    // The whole index is validated either way; output offsets and chunk indices count from block `first`.
    __shared__ __align__(16) uint32_t s_stage[XXH_WARP_SMEM_WORDS];
    __shared__ unsigned long long tmp[32];
    __shared__ int s_err;
    __shared__ uint32_t s_foot_hash;
    __shared__ unsigned long long s_base_out, s_base_chunk;

    if (threadIdx.x == 0) { s_err = FOURMC_OK; s_base_out = 0; s_base_chunk = 0; }
    __syncthreads();

    // header :577-585, footer :670-688 / FourMcInputStream.java:187-228
    uint32_t fsize = 0;
    const uint8_t *foot = nullptr;
    if (n < 12 + 12 + 20) {
        if (threadIdx.x == 0) s_err = FOURMC_E_INPUT;
    } else {
        fsize = ld_be32(in + n - 12);
        if (fsize < 20 || (uint64_t)fsize > n - 24 || ((fsize - 20) & 3) || (fsize - 20) / 4 != n_blocks_host) {
            if (threadIdx.x == 0) s_err = FOURMC_E_CONTENT;
        } else {
            foot = in + n - fsize;
        }
    }
    __syncthreads();
    if (s_err != FOURMC_OK) {
        if (threadIdx.x == 0) { info->status = s_err; info->n_blocks = 0; info->total_usize = 0; info->total_chunks = 0; }
        return;
    }
    if (threadIdx.x < 32) {
        const uint32_t h = xxh32_warp<true>(foot, fsize - 4, 0, s_stage);
        if (threadIdx.x == 0) s_foot_hash = h;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int e = FOURMC_OK;
        if (ld_be32(in) != magic) e = FOURMC_E_CONTENT;
        else if (ld_be32(in + 4) != FOURMC_VERSION) e = FOURMC_E_CONTENT;
        else if (ld_be32(in + 8) != xxh32_thread(in, 8, 0)) e = FOURMC_E_CONTENT;
        else if (ld_be32(foot) != fsize || ld_be32(foot + 4) != 1) e = FOURMC_E_CONTENT;
        else if (ld_be32(in + n - 8) != magic) e = FOURMC_E_CONTENT;
        else if (ld_be32(in + n - 4) != s_foot_hash) e = FOURMC_E_CONTENT;
        s_err = e;
    }
    __syncthreads();
    if (s_err != FOURMC_OK) {
        if (threadIdx.x == 0) { info->status = s_err; info->n_blocks = 0; info->total_usize = 0; info->total_chunks = 0; }
        return;
    }

    const uint32_t nb = n_blocks_host;
    const uint64_t eos_pos = n - fsize - 12;
    unsigned long long c_off = 0, c_out = 0, c_chunks = 0;
    for (uint32_t i0 = 0; i0 < nb || i0 == 0; i0 += SCAN_THREADS) {
        const uint32_t i = i0 + threadIdx.x;
        const bool live = i < nb;
        unsigned long long total;
        // absolute offset of block i = prefix sum of deltas (FourMcInputStream.java:230-236)
        const unsigned long long delta = live ? ld_be32(foot + 8 + 4 * i) : 0;
        const unsigned long long off = c_off + cta_incl_scan_u64(delta, tmp, &total);
        c_off += total;
        uint32_t u = 0, c = 0, ck = 0;
        bool bad = false;
        // a range reader looks at its own blocks only (the others may not even be there: the ranks of a multi-GPU
        // reader each hold their part of the stream); the index itself is validated as a whole either way
        const bool looked_at = live && (count == 0xffffffffu || (i >= first && i - first < count));
        if (live && !looked_at) { u = FOURMC_BLOCKSIZE; c = 0; }
        if (looked_at) {
            if (off + 12 > eos_pos) bad = true;
            else {
                u = ld_be32(in + off); c = ld_be32(in + off + 4); ck = ld_be32(in + off + 8);
                // the serial reader would reach the next header at off + 12 + c (:631); the index must agree
                const unsigned long long next = off + 12 + c;
                const unsigned long long expect = (i + 1 < nb) ? off + ld_be32(foot + 8 + 4 * (i + 1)) : eos_pos;
                if (next != expect) bad = true;
                if (i == 0 && off != 12) bad = true;
                if (u == 0 && c == 0 && ck == 0) bad = true;          // an EOS mark inside the index range
            }
        }
        const bool toolarge = live && !bad && (c > FOURMC_BLOCKSIZE || (c != u && u > FOURMC_BLOCKSIZE));  // :618, :651
        const bool usable = live && !bad && !toolarge;
        const unsigned long long oincl = cta_incl_scan_u64(usable ? u : 0, tmp, &total);
        const unsigned long long dst_off = c_out + oincl - (usable ? u : 0);
        c_out += total;
        const uint32_t nch = usable && c != u ? (c + 15 + LZ4_CHUNK - 1) / LZ4_CHUNK : 0;   // +15: aligned coordinates
        const unsigned long long cincl = cta_incl_scan_u64(nch, tmp, &total);
        const unsigned long long chunk_base = c_chunks + cincl - nch;
        c_chunks += total;
        if (live && i == first) { s_base_out = dst_off; s_base_chunk = chunk_base; }
        __syncthreads();
        if (live) {
            const bool mine = i >= first && i - first < count;
            const unsigned long long my_out = dst_off - s_base_out;
            BlockDesc d;
            d.src = in + off + 12; d.dst = mine ? out + my_out : nullptr;
            d.csize = c; d.usize = u; d.chunk_base = mine ? (uint32_t)(chunk_base - s_base_chunk) : 0u; d.stored = (c == u) ? 1u : 0u;
            uint8_t st = FOURMC_BLOCK_OK;
            if (bad) { st = FOURMC_BLOCK_CORRUPT; atomicMin(&s_err, FOURMC_E_CONTENT); d.csize = 0; d.usize = 0; d.stored = 1; }
            else if (toolarge) { st = FOURMC_BLOCK_TOOLARGE; d.csize = 0; d.usize = 0; d.stored = 1; }
            else if (mine && my_out + u > out_cap) { atomicMin(&s_err, FOURMC_E_OUTPUT); d.csize = 0; d.usize = 0; d.stored = 1; }
            desc[i] = d; xxh_expect[i] = ck; status[i] = st;
        }
        if (nb == 0) break;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int e = s_err;
        // EOS mark :616
        for (int k = 0; k < 12 && e == FOURMC_OK; k++) if (in[eos_pos + k] != 0) e = FOURMC_E_CONTENT;
        if (nb == 0 && eos_pos != 12 && e == FOURMC_OK) e = FOURMC_E_CONTENT;
        info->status = e; info->n_blocks = nb; info->total_usize = c_out; info->total_chunks = (uint32_t)c_chunks;
    }
}

// one warp per block
constexpr int VERIFY_WARPS = 4;
__global__ void __launch_bounds__(VERIFY_WARPS * 32)
xxh_verify_kernel(const BlockDesc *desc, const uint32_t *xxh_expect, uint32_t n_blocks, uint8_t *status)
{
    __shared__ __align__(16) uint32_t s_stage[VERIFY_WARPS][XXH_WARP_SMEM_WORDS];
    const uint32_t b = blockIdx.x * VERIFY_WARPS + (threadIdx.x >> 5);
    if (b >= n_blocks) return;
    if (status[b] != FOURMC_BLOCK_OK) return;
    const BlockDesc d = desc[b];
    const uint32_t h = xxh32_warp<true>(d.src, d.csize, 0, s_stage[threadIdx.x >> 5]);
    if ((threadIdx.x & 31) == 0 && h != xxh_expect[b]) status[b] = FOURMC_BLOCK_CHECKSUM;
}

__global__ void __launch_bounds__(VERIFY_WARPS * 32)
xxh_batch_kernel(const uint8_t *base, const uint64_t *off, const uint32_t *len, uint32_t n_items,
                 uint32_t seed, uint32_t *out)
{
    __shared__ __align__(16) uint32_t s_stage[VERIFY_WARPS][XXH_WARP_SMEM_WORDS];
    const uint32_t b = blockIdx.x * VERIFY_WARPS + (threadIdx.x >> 5);
    if (b >= n_items) return;
    const uint32_t h = xxh32_warp<true>(base + off[b], len[b], seed, s_stage[threadIdx.x >> 5]);
    if ((threadIdx.x & 31) == 0) out[b] = h;
}

// Folds D1's verdict into the per-block status and reduces to the stream result:
// result[0] = decoded size or FOURMC_E_*, result[1] = first failing block or -1.  One CTA.
//
// A compressed block may decode to FEWER bytes than its header announces: the reference writes what the
// decoder returned (native/4mc.c:661-666, :810-815 `filesize += decodedBytes`).  *short_blocks counts such
// blocks; compact_kernel then closes the gaps the way the serial reader's output has none.
__global__ void __launch_bounds__(SCAN_THREADS)
finalize_kernel(const BlockDesc *desc, const int32_t *parse_result, uint32_t n_blocks, uint8_t *status,
                int32_t *out_size, const IndexInfo *info, long long *result, uint32_t *short_blocks)
{
    __shared__ unsigned int s_first, s_short;
    __shared__ unsigned long long s_total;
    if (threadIdx.x == 0) { s_first = 0xffffffffu; s_total = 0; s_short = 0; }
    __syncthreads();
    unsigned long long mine = 0;
    unsigned int shorts = 0;
    for (uint32_t i = threadIdx.x; i < n_blocks; i += SCAN_THREADS) {
        uint8_t st = status[i];
        const int32_t r = parse_result[i];
        if (st == FOURMC_BLOCK_OK && r < 0) { st = FOURMC_BLOCK_CORRUPT; status[i] = st; }
        if (out_size) out_size[i] = r;
        if (st != FOURMC_BLOCK_OK) atomicMin(&s_first, i);
        else { mine += (unsigned long long)r; if ((uint32_t)r != desc[i].usize) shorts++; }
    }
    atomicAdd(&s_total, mine);
    if (shorts) atomicAdd(&s_short, shorts);
    __syncthreads();
    if (threadIdx.x == 0 && short_blocks) *short_blocks = s_short;
    if (threadIdx.x == 0 && result) {
        const int e = info ? info->status : FOURMC_OK;
        if (e != FOURMC_OK) { result[0] = e; result[1] = -1; }
        else if (s_first != 0xffffffffu) { result[0] = FOURMC_E_CONTENT; result[1] = (long long)s_first; }
        else { result[0] = (long long)s_total; result[1] = -1; }
    }
}

// Blocks that decoded short leave gaps between consecutive blocks of a stream (every block was placed at the
// sum of the ANNOUNCED sizes before it).  One CTA moves the blocks down in stream order; nothing to do -- the
// common case -- when *short_blocks is zero.  Only for descriptors whose destinations are consecutive.
__global__ void __launch_bounds__(SCAN_THREADS)
compact_kernel(const BlockDesc *desc, const int32_t *parse_result, const uint8_t *status, uint32_t n_blocks,
               const uint32_t *short_blocks)
{
    if (*short_blocks == 0) return;
    size_t shift = 0;
    for (uint32_t b = 0; b < n_blocks; b++) {
        if (status[b] != FOURMC_BLOCK_OK) return;                // the stream fails at this block: nothing after it counts
        const uint32_t got = (uint32_t)parse_result[b];
        if (shift) {
            const uint8_t *s = desc[b].dst;
            uint8_t *d = desc[b].dst - shift;
            for (uint32_t i0 = 0; i0 < got; i0 += SCAN_THREADS) {    // moving down: read a slab, then write it
                const uint32_t i = i0 + threadIdx.x;
                uint8_t v = 0;
                if (i < got) v = s[i];
                __syncthreads();
                if (i < got) d[i] = v;
                __syncthreads();
            }
        }
        shift += desc[b].usize - got;
    }
}

}  // namespace fm
