// zstd_encode.cuh -- 4mz block compression kernels ("4mz Fast").
//
// Reference behaviour being replaced: the ZSTD writer loop native/4mc.c:446-500 (ZSTD_compress per
// 4 MiB block :467, stored fallback :469-485, XXH32 of the payload, 12-byte header) and
// native/jniZstdCompressor.c:59-199.  Compressed bytes need not match the reference; every frame
// decodes with ZSTD_decompress (tests).  Pipeline per batch of 4 MiB blocks:
//
//  Z1 lz4_region_kernel<true>   (lz4_encode.cuh) the shared-memory match finder, emitting per 64 KiB
//       region the sequence arrays (literal length, match length, offset) and the gathered literals
//  Z2 zstd_entropy_kernel       one CTA of 128 threads per region = one zstd block: Huffman literals
//       (4 streams, one warp each), FSE sequences (three state chains on three lanes, bit packing
//       by all threads from prefix-summed bit positions).  The code is zenc_region() of
//       zstd_encode.h, shared with the CPU emulation test.
//  Z3 zstd_block_size_kernel    per 4 MiB block: frame size = 9 + sum of (3 + block body or raw
//       region), compressed vs stored (native/4mc.c:469-485)
//     zstd_block_write_kernel   per 4 MiB block: frame header, block headers, bodies copied to their
//       final place in the .4mz stream, XXH32 of the payload, 12-byte 4mz block header
#pragma once

#include "container.cuh"
#include "fm_common.cuh"
#include "lz4_encode.cuh"
#include "xxh32.cuh"
#include "zstd_encode.h"

namespace fm {

struct CtaExec {
    template <class F> __device__ __forceinline__ void phase(F f) { f((int)threadIdx.x); __syncthreads(); }
    __device__ __forceinline__ void excl_scan(uint32_t *arr, uint32_t *tmp, uint32_t *total)
    {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const uint32_t v = arr[tid];
        const uint32_t incl = (uint32_t)warp_incl_scan_add((int)v);
        if (lane == 31) tmp[warp] = incl;
        __syncthreads();
        uint32_t base = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < fmz::ZE_WARPS; w++) { const uint32_t t = tmp[w]; if (w < warp) base += t; tot += t; }
        arr[tid] = base + incl - v;
        if (tid == 0) *total = tot;
        __syncthreads();
    }
    __device__ __forceinline__ void add32(uint32_t *p, uint32_t v) { atomicAdd(p, v); }
    __device__ __forceinline__ void max32(uint32_t *p, uint32_t v) { atomicMax(p, v); }
};

struct ZEncParams {
    const RegionMeta *meta;        // from Z1: nseq, body_bytes = literal count
    const uint8_t *scratch_in;     // n_regions * ZE_IN_SLOT
    uint8_t *scratch_out;          // n_regions * ZE_OUT_SLOT
    fmz::ZRegionOut *rout;         // n_regions
    const fmz::Tables *tables;
    uint64_t n;                    // input bytes of the batch
    uint32_t n_regions;
    uint32_t region_bytes, regions_per_block;      // 65536 x 64 (Fast parse) or 32768 x 128 (chain parse)
    uint32_t block_bytes;                          // 0 = FOURMC_BLOCKSIZE; smaller for the raw codec streams' chunks
};

__global__ void __launch_bounds__(fmz::ZE_THREADS) zstd_entropy_kernel(ZEncParams P)
{
    __shared__ fmz::ZShared sh;
    const uint32_t rg = blockIdx.x;
    const uint32_t blk = rg / P.regions_per_block, rib = rg % P.regions_per_block;
    const uint32_t block_bytes = P.block_bytes ? P.block_bytes : (uint32_t)FOURMC_BLOCKSIZE;
    const uint64_t blk_off = (uint64_t)blk * block_bytes;
    const uint32_t blk_len = (uint32_t)min((uint64_t)block_bytes, P.n - blk_off);
    const uint32_t r_off = rib * P.region_bytes;
    if (r_off >= blk_len) return;                                  // region beyond a short last block
    const RegionMeta m = P.meta[rg];
    fmz::ZRegionIn in;
    const uint8_t *slot_in = P.scratch_in + (size_t)rg * fmz::ze_in_slot(P.region_bytes);
    const uint32_t stride = fmz::ze_seq_stride(m.nseq);
    in.ll = (const uint16_t *)slot_in; in.ml = in.ll + stride; in.off = in.ml + stride;
    in.lits = (const uint8_t *)(in.off + stride);
    in.nseq = m.nseq; in.nlit = m.body_bytes;
    in.rlen = min(P.region_bytes, blk_len - r_off);
    in.cap = P.region_bytes;
    CtaExec ex;
    fmz::zenc_region(ex, sh, in, (uint32_t *)(P.scratch_out + (size_t)rg * fmz::ze_out_slot(P.region_bytes)), &P.rout[rg], *P.tables);
}

// raw_limit < 0: container mode, a block is stored when its frame reaches its raw size
// (native/4mc.c:469-485: ZSTD_compress is offered u-1 bytes).  raw_limit >= 0: bare frame for the
// per-block API; "stored" then means "does not fit in raw_limit bytes".
__global__ void zstd_block_size_kernel(const fmz::ZRegionOut *rout, uint32_t n_blocks, uint64_t n,
                                       BlockPlan *plan, uint32_t *block_lens, int64_t raw_limit,
                                       uint32_t region_bytes, uint32_t regions_per_block,
                                       uint32_t block_bytes = FOURMC_BLOCKSIZE)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint64_t blk_off = (uint64_t)b * block_bytes;
    const uint32_t u = (uint32_t)min((uint64_t)block_bytes, n - blk_off);
    const fmz::ZRegionOut *r = rout + (size_t)b * regions_per_block;
    uint32_t c = fmz::ZE_FRAME_HDR;
    const uint32_t nreg = (u + region_bytes - 1) / region_bytes;
    for (uint32_t k = 0; k < nreg; k++) {
        const uint32_t rlen = min(region_bytes, u - k * region_bytes);
        c += 3u + (r[k].raw ? rlen : r[k].bytes);
    }
    if (nreg == 0) c += 3u;                                         // empty input: one empty raw block
    BlockPlan p;
    p.usize = u;
    if (raw_limit >= 0) {
        p.stored = ((int64_t)c > raw_limit) ? 1u : 0u;
        p.payload = p.stored ? 0u : c;
    } else {
        p.stored = (c >= u) ? 1u : 0u;
        p.payload = p.stored ? u : c;
    }
    p.final_lits = 0;
    plan[b] = p;
    if (block_lens) block_lens[b] = 12u + p.payload;
}

__global__ void __launch_bounds__(ENC_WRITE_THREADS)
zstd_block_write_kernel(const uint8_t *in, const uint8_t *scratch_out, const fmz::ZRegionOut *rout,
                        const BlockPlan *plan, const uint64_t *block_off, uint8_t *out_base, int raw_mode,
                        uint32_t region_bytes, uint32_t regions_per_block, uint32_t block_bytes = FOURMC_BLOCKSIZE)
{
    __shared__ __align__(16) uint32_t s_stage[XXH_WARP_SMEM_WORDS];
    __shared__ uint32_t s_dst[ENC_MAX_REGIONS_PER_BLOCK + 1];

    const uint32_t b = blockIdx.x;
    const BlockPlan p = plan[b];
    const uint64_t blk_off = (uint64_t)b * block_bytes;
    const uint8_t *src = in + blk_off;
    uint8_t *rec = out_base + block_off[b];
    uint8_t *pay = rec + 12;
    const fmz::ZRegionOut *r = rout + (size_t)b * regions_per_block;
    const uint32_t nreg = (p.usize + region_bytes - 1) / region_bytes;

    if (p.stored) {
        if (raw_mode) return;
        cta_copy(pay, src, p.usize);
    } else {
        if (threadIdx.x == 0) {
            uint32_t c = fmz::ZE_FRAME_HDR;
            for (uint32_t k = 0; k < nreg; k++) {
                s_dst[k] = c;
                const uint32_t rlen = min(region_bytes, p.usize - k * region_bytes);
                c += 3u + (r[k].raw ? rlen : r[k].bytes);
            }
            fmz::ze_write_frame_header(pay, p.usize);
            if (nreg == 0) fmz::ze_write_block_header(pay + fmz::ZE_FRAME_HDR, true, 0, 0);
        }
        __syncthreads();
        for (uint32_t k = 0; k < nreg; k++) {
            const uint32_t rlen = min(region_bytes, p.usize - k * region_bytes);
            const fmz::ZRegionOut x = r[k];
            uint8_t *o = pay + s_dst[k];
            const bool last = k + 1 == nreg;
            if (threadIdx.x == 0) fmz::ze_write_block_header(o, last, x.raw ? 0 : 2, x.raw ? rlen : x.bytes);
            if (x.raw) cta_copy(o + 3, src + (size_t)k * region_bytes, rlen);
            else cta_copy(o + 3, scratch_out + ((size_t)b * regions_per_block + k) * fmz::ze_out_slot(region_bytes), x.bytes);
        }
    }
    __threadfence_block();
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t h = xxh32_warp<false>(pay, p.payload, 0, s_stage);
        if (threadIdx.x == 0) { st_be32(rec, p.usize); st_be32(rec + 4, p.payload); st_be32(rec + 8, h); }
    }
}

// block_off[b] = *carry + sum_{i<b} lens[i]; then *carry += sum of all lens and *span (optional)
// receives that sum.  Lets batches of blocks be appended to one stream without a host round trip.
__global__ void __launch_bounds__(SCAN_THREADS)
scan_lens_carry_kernel(const uint32_t *lens, uint32_t n_blocks, uint64_t *carry_io, uint64_t *block_off, uint64_t *span_acc)
{
    __shared__ unsigned long long tmp[32];
    const unsigned long long base = *carry_io;
    unsigned long long carry = 0;
    for (uint32_t i0 = 0; i0 < n_blocks; i0 += SCAN_THREADS) {
        const uint32_t i = i0 + threadIdx.x;
        const unsigned long long v = i < n_blocks ? lens[i] : 0;
        unsigned long long total;
        const unsigned long long incl = cta_incl_scan_u64(v, tmp, &total);
        if (i < n_blocks) block_off[i] = base + carry + incl - v;
        carry += total;
    }
    __syncthreads();
    if (threadIdx.x == 0) { *carry_io = base + carry; if (span_acc) *span_acc += carry; }
}

}  // namespace fm
