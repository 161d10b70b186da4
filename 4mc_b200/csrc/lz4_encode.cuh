// lz4_encode.cuh -- LZ4 block compression kernels ("4mc Fast").
//
// Reference behaviour being replaced: LZ4_compress_default / LZ4_compress_generic_validated
// (native/lz4/lz4.c:1435, :910-1302) as called per 4 MiB block at native/4mc.c:301 and
// native/jniCompressor.c:91.  The reference is a serial greedy parse with a 4096-entry hash of
// "last seen" positions.  Compressed bytes need not match it, only be a valid LZ4 block, so the
// algorithm here is shaped for the GPU instead (DESIGN.md "LZ4 encode"):
//
//  E1 lz4_region_kernel   persistent CTAs; each takes 64 KiB REGIONS of a block:
//       stage   region -> shared memory with one bulk async copy (cp.async.bulk + mbarrier)
//       index   16384-entry table of the FIRST (even) position of each 4-byte hash in the region,
//               order-independent (descending sweep; inside a step the lowest position wins), so all
//               threads build it at once and every run produces the same bytes; any earlier
//               occurrence is a usable LZ4 match candidate
//       parse   every thread owns a 132-byte SLICE (132 = 33 words: slices start in distinct
//               shared-memory banks).  Pass 1 tests EVERY position of the slice against the table
//               (uniform work, one bit per position); pass 2 walks greedily over the set bits only,
//               extending matches forwards and backwards -- they may run past the slice end
//       stitch  two CTA-wide max-scans: matches of later slices are trimmed to what earlier
//               slices left uncovered, and every slice learns where its first literal run starts
//       emit    prefix-sum of encoded sizes, then every thread writes its own sequences
//     Output per region: the encoded sequences ("body", as if the region stood alone) in a
//     scratch slot, plus a RegionMeta.  Literals after the region's last match are not stored:
//     they are input bytes.
//  E2 lz4_block_size_kernel   per block: walks its <= 64 RegionMeta, merges each region's
//     leading literals with the literals carried over from the previous region (re-encoding one
//     token per region join), and decides compressed vs stored (native/4mc.c:301-329).
//  E3 lz4_block_write_kernel  per block: copies the pieces to their final place in the .4mc
//     stream, hashes the payload (XXH32, native/4mc.c:311/323) and writes the 12-byte header.
#pragma once

#include "fm_common.cuh"
#include "xxh32.cuh"

namespace fm {

constexpr int ENC_REGION = 65536;
constexpr int ENC_REGIONS_PER_BLOCK = FOURMC_BLOCKSIZE / ENC_REGION;   // 64
constexpr int ENC_SLOT = ENC_REGION + 512;       // scratch bytes per region body (worst case R + R/255 + 16)
constexpr int ENC_THREADS = 512;
constexpr int ENC_WARPS = ENC_THREADS / 32;
constexpr int ENC_SLICE = 132;
constexpr int ENC_HASH_BITS = 14;
constexpr int ENC_MAXREC = ENC_SLICE / 4 + 1;    // inner records of one slice
constexpr int ENC_PAD = 64;
constexpr int ENC_STAGE = 48 * 1024;             // output staging: the (dead) hash table + 16 KiB
constexpr size_t ENC_SMEM = ENC_REGION + ENC_PAD + ENC_STAGE;
// Chain parse (levels 2..4).  Geometry: the shared-memory window holds 64 KiB of LOOK-BACK (the block's previous
// bytes, searchable but not parsed: LZ4's whole offset range, so every position sees the same sliding history the
// reference's HC search sees, native/lz4/lz4hc.c:262-263) + a 32 KiB region of new bytes; 128 regions per block.
// The chain links live in HBM (lz4_chain_kernel); shared memory holds the window and, per new position, the longest
// match found (u16 offset + u8 length), which becomes the output staging area once the parse is done.
constexpr int ENC_CHAIN_REGION = 32768;
constexpr int ENC_CHAIN_LOOKBACK = 65536;
constexpr int ENC_CHAIN_WINDOW = ENC_CHAIN_LOOKBACK + ENC_CHAIN_REGION;
constexpr int ENC_CHAIN_THREADS = 1024;          // the search is latency bound per thread (chain links come from L2)
constexpr int ENC_CHAIN_SLICE = 64;              // positions per cost-optimal parse (one thread each, 512 per region)
constexpr int ENC_CHAIN_MAXREC = ENC_CHAIN_SLICE / 4 + 1;
constexpr int ENC_CHAIN_LENCAP = 255;            // match lengths are searched and stored up to here, longer ones are extended when taken
constexpr int ENC_CHAIN_CREDIT = 6;              // parse cost model: 16 per output byte; a byte covered beyond the slice end is worth 6
constexpr int ENC_CHAIN_SHORTER = 8;             // shorter lengths of a match tried by the parse
constexpr int ENC_CHAIN_SWAPSCAN = 16;           // 4-byte windows of a new longest match inspected for a better chain
constexpr int ENC_STAGE_CHAIN = 3 * ENC_CHAIN_REGION;
constexpr int ENC_CHAIN_SLOT = ENC_CHAIN_REGION + 512;
constexpr int ENC_MAX_REGIONS_PER_BLOCK = FOURMC_BLOCKSIZE / ENC_CHAIN_REGION;   // 128
static_assert(ENC_CHAIN_THREADS * ENC_CHAIN_SLICE >= ENC_CHAIN_REGION, "chain slices must cover the region");
constexpr size_t ENC_SMEM_CHAIN = ENC_CHAIN_WINDOW + ENC_PAD + ENC_STAGE_CHAIN;
// chain links: one warp per CHUNK of a block, 64 KiB of warm-up before it (links reach at most 65535 back)
constexpr int ENC_LINK_PIECE = 512;              // positions per staged piece of lz4_chain_kernel
constexpr int ENC_LINK_STAGE_BYTES = ENC_LINK_PIECE + 32;
constexpr int ENC_LINK_SMEM = (4 << ENC_HASH_BITS) + 2 * ENC_LINK_STAGE_BYTES;
static_assert((sizeof(uint16_t) << ENC_HASH_BITS) <= ENC_STAGE, "the hash table lives inside the staging area");

static_assert(ENC_THREADS * ENC_SLICE >= ENC_REGION, "slices must cover the region");

struct RegionMeta {
    uint32_t body_bytes;     // bytes in the scratch slot
    uint32_t tail_lits;      // input bytes after the region's last match (whole region if nseq == 0)
    uint32_t lead;           // literal count of the region's first sequence
    uint32_t nseq;
};

struct EncParams {
    const uint8_t *in;       // contiguous input
    uint64_t n;              // input bytes
    uint32_t n_regions;      // total over all blocks (64 per block, trailing ones may be empty)
    uint8_t *scratch;        // n_regions * ENC_SLOT
    RegionMeta *meta;        // n_regions
    uint32_t *work_counter;  // zeroed before launch
    int min_match;           // >= 4
    uint32_t slot_bytes;     // scratch bytes per region (ENC_SLOT; fmz::ZE_IN_SLOT for sequence output)
    int depth;               // chain parse: candidates tried per search (levels 2..4); unused by the Fast parse
    uint32_t region_bytes;   // new bytes per region: ENC_REGION (Fast) or ENC_CHAIN_REGION (chain)
    uint32_t regions_per_block;
    uint32_t block_bytes;    // bytes per block: 0 = FOURMC_BLOCKSIZE (the containers); the raw codec streams cut smaller chunks
    int reproducible;        // ties between racing table stores are settled by position: the same bytes on every run
    const uint16_t *chain;   // chain parse: per input byte of this launch, distance to the previous position with the same hash (0: none)
};

__device__ __forceinline__ uint32_t enc_hash(uint32_t v) { return (v * 2654435761u) >> (32 - ENC_HASH_BITS); }

// bytes of LZ4 length continuation for a field value v (0 when v < 15)
__device__ __forceinline__ int enc_ext_bytes(int v) { return v < 15 ? 0 : 1 + (v - 15) / 255; }

__device__ __forceinline__ uint32_t smem_read4(const uint32_t *data32, int p)
{
    const int w = p >> 2;
    return __funnelshift_r(data32[w], data32[w + 1], (p & 3) * 8);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// CTA-wide exclusive scan over ENC_THREADS values (max or add); `tmp` holds ENC_WARPS ints.
// Identity is 0 for both (all values are >= 0).  *total (optional) = reduction over the CTA.
template <bool IS_MAX, int NWARPS = ENC_WARPS>
__device__ __forceinline__ int cta_excl_scan(int v, int *tmp, int *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int incl = IS_MAX ? warp_incl_scan_max(v) : warp_incl_scan_add(v);
    if (lane == 31) tmp[warp] = incl;
    __syncthreads();
    const int wv = lane < NWARPS ? tmp[lane] : 0;
    const int wincl = IS_MAX ? warp_incl_scan_max(wv) : warp_incl_scan_add(wv);
    int wexcl = __shfl_up_sync(FM_FULL, wincl, 1);
    if (lane == 0) wexcl = 0;
    const int base = __shfl_sync(FM_FULL, wexcl, warp);
    if (total) *total = __shfl_sync(FM_FULL, wincl, 31);
    int excl = __shfl_up_sync(FM_FULL, incl, 1);
    if (lane == 0) excl = 0;
    __syncthreads();                                           // tmp reusable
    return IS_MAX ? max(base, excl) : base + excl;
}

__device__ __forceinline__ uint8_t *emit_len(uint8_t *o, int v)
{
    while (v >= 255) { *o++ = 255; v -= 255; }
    *o++ = (uint8_t)v;
    return o;
}

// inner record: start (relative to the slice, 8 bits) | length (8 bits) | offset (16 bits)
__device__ __forceinline__ uint32_t enc_pack(int st_rel, int len, int off)
{
    return ((uint32_t)st_rel << 24) | ((uint32_t)len << 16) | (uint32_t)off;
}

// Chain links for levels 2..4: chain[p] = distance from position p of a block to the previous position of the
// same block with the same 4-byte hash, 0 when there is none within 65535 (what the reference keeps in its
// chainTable, native/lz4/lz4hc.c:120-141; here for every position of the block at once, 2 bytes per input byte).
// A link never reaches further than 65535 back, so a block is cut into chunks that are linked independently:
// one warp per chunk walks 64 KiB of warm-up and then its chunk in steps of 32 positions, a table of the last
// position per hash in shared memory (64 KiB).  Equal hashes inside a step are detected by reading the table back
// after the step's stores; only then are they linked lane to lane (__match_any_sync) and the highest lane made the
// head, so the links are exact and the same on every run.
__global__ void __launch_bounds__(32) lz4_chain_kernel(const uint8_t *in, uint64_t n, uint32_t block_bytes, uint32_t chunk,
                                                       uint32_t chunks_per_block, uint32_t n_items, uint16_t *chain)
{
    extern __shared__ __align__(16) uint32_t link_head[];           // 1 << ENC_HASH_BITS, then two staging buffers
    uint8_t *stage = (uint8_t *)(link_head + (1 << ENC_HASH_BITS)); // 2 x ENC_LINK_STAGE_BYTES
    const int lane = threadIdx.x;
    const uint8_t *in_end = in + n;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t blk = item / chunks_per_block, ci = item % chunks_per_block;
        const uint64_t blk_off = (uint64_t)blk * block_bytes;
        if (blk_off >= n) continue;
        const int blk_len = (int)min((uint64_t)block_bytes, n - blk_off);
        const int c0 = (int)(ci * chunk);
        if (c0 >= blk_len) continue;
        const int c1 = min(c0 + (int)chunk, blk_len);
        const int w0 = max(c0 - 65536, 0);
        const int last = blk_len - 4;                               // last position with four bytes
        for (int i = lane; i < (1 << ENC_HASH_BITS); i += 32) link_head[i] = 0xffffffffu;
        const uint8_t *b = in + blk_off;
        uint16_t *out = chain + blk_off;
        // The bytes travel in pieces of 512 positions (16 steps): 34 aligned 16-byte loads per piece, requested one
        // piece ahead and parked in shared memory, so that no step waits for HBM.
        const uint8_t *ab = (const uint8_t *)((uintptr_t)(b + w0) & ~(uintptr_t)15);
        const int m16 = (int)((b + w0) - ab);
        auto fetch = [&](const uint8_t *a) -> uint4 {               // 16 bytes at a (aligned); bytes outside the input read as 0
            if (a >= in && a + 16 <= in_end) return ldg_nc_v4((const uint4 *)a);
            uint32_t w[4] = {0u, 0u, 0u, 0u};
            for (int j = 0; j < 16; j++) if (a + j >= in && a + j < in_end) w[j >> 2] |= (uint32_t)a[j] << (8 * (j & 3));
            return make_uint4(w[0], w[1], w[2], w[3]);
        };
        const int n_pieces = (c1 - w0 + ENC_LINK_PIECE - 1) / ENC_LINK_PIECE;
        uint4 r0 = fetch(ab + 16 * lane), r1 = make_uint4(0u, 0u, 0u, 0u);
        if (lane < 2) r1 = fetch(ab + 16 * (32 + lane));
        for (int t = 0; t < n_pieces; t++) {
            uint8_t *sb = stage + (t & 1) * ENC_LINK_STAGE_BYTES;
            ((uint4 *)sb)[lane] = r0;
            if (lane < 2) ((uint4 *)sb)[32 + lane] = r1;
            __syncwarp();
            if (t + 1 < n_pieces) {
                const uint8_t *na = ab + (size_t)(t + 1) * ENC_LINK_PIECE;
                r0 = fetch(na + 16 * lane);
                if (lane < 2) r1 = fetch(na + 16 * (32 + lane));
            }
            const uint32_t *sw = (const uint32_t *)sb;
            const int pbase = w0 + t * ENC_LINK_PIECE;
#pragma unroll 4
            for (int sidx = 0; sidx < ENC_LINK_PIECE / 32; sidx++) {
                const int base = pbase + 32 * sidx;
                if (base >= c1) break;
                const int p = base + lane;
                const int bo = m16 + 32 * sidx + lane;              // byte offset within the piece
                const bool live = p <= last;
                const uint32_t v = __funnelshift_r(sw[bo >> 2], sw[(bo >> 2) + 1], (bo & 3) * 8);
                const uint32_t h = live ? enc_hash(v) : 0x10000u + (uint32_t)lane;
                uint32_t q = live ? link_head[h] : 0xffffffffu;
                __syncwarp();
                if (live) link_head[h] = (uint32_t)p;               // equal hashes inside the step race: one of them lands
                __syncwarp();
                const bool lost = live && link_head[h] != (uint32_t)p;
                if (__any_sync(FM_FULL, lost)) {
                    // some hash occurs twice among the 32 positions (rare on text, the rule on short periods): link
                    // lane to lane and make the highest lane the head
                    const uint32_t same = __match_any_sync(FM_FULL, h);
                    const uint32_t below = same & ((1u << lane) - 1u);
                    if (below) q = (uint32_t)(base + 31 - __clz(below));
                    __syncwarp();
                    if (live && (same >> lane) == 1u) link_head[h] = (uint32_t)p;
                    __syncwarp();
                }
                const uint32_t dist = (uint32_t)p - q;              // q == 0xffffffff: p + 1, never a valid link below
                if (p >= c0 && p < c1) out[p] = (q != 0xffffffffu && dist <= 65535u) ? (uint16_t)dist : (uint16_t)0;
            }
        }
        __syncwarp();
    }
}

// ZSEQ = false: LZ4 sequences as bytes (4mc).  ZSEQ = true: the same parse, emitted as
// (literal length, match length, offset) arrays plus the gathered literals for the zstd entropy
// stage (zstd_encode.cuh): slot = u16 ll[n] | u16 ml[n] | u16 off[n] | literals, n rounded up to 8;
// RegionMeta.body_bytes then holds the literal count (tail literals included).
//
// CHAIN = false: the Fast parse (level 1), first-occurrence table.  CHAIN = true: levels 2..4
// (LZ4 MC and HC are hash-chain searches, native/lz4/lz4mc.c:518-579,
// native/lz4/lz4hc.c:239-447).  The chain -- every position linked to the previous position with the same
// 4-byte hash, at most 65535 back -- is built for the whole block beforehand (lz4_chain_kernel, HBM, read
// through L1 / L2 here).  Per region:
//   search  EVERY new position looks for its longest match: up to `depth` links, and, like the reference's
//           search (lz4hc.c:321-343), a new longest match switches the walk to the chain of that one of its
//           4-byte windows whose previous occurrence lies furthest back (a longer match must repeat every window)
//   parse   one thread per 64-byte slice chooses, back to front, literal or match (at its full length or up to 8
//           bytes shorter) at every position so that the encoded size of the slice is minimal (16 per byte; bytes
//           a match covers beyond the slice are credited 6 each) -- a cost-optimal parse instead of the
//           reference's lazy arbitration between overlapping matches (lz4hc.c:553-788)
//   walk    every slice follows the choices from its own start; a max-scan tells every slice what earlier slices
//           cover; it walks again from there.  What still overlaps is trimmed by the stitching shared with the Fast parse.
// One CTA of 1024 threads per SM (192 KiB of shared memory).
template <bool ZSEQ, bool CHAIN>
__global__ void __launch_bounds__(CHAIN ? ENC_CHAIN_THREADS : ENC_THREADS, CHAIN ? 1 : 2) lz4_region_kernel(EncParams P)
{
    constexpr int NT = CHAIN ? ENC_CHAIN_THREADS : ENC_THREADS, NW = NT / 32;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *data = smem;                                          // window + ENC_PAD
    uint32_t *data32 = (uint32_t *)smem;
    uint16_t *table = (uint16_t *)(smem + (CHAIN ? ENC_CHAIN_WINDOW : ENC_REGION) + ENC_PAD);
    __shared__ int s_scan[NW];
    __shared__ uint32_t s_work;
    __shared__ int s_flush;
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x, lane = tid & 31;
    uint32_t phase = 0;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    for (;;) {
        if (tid == 0) s_work = atomicAdd(P.work_counter, 1u);
        __syncthreads();
        const uint32_t rg = s_work;
        if (rg >= P.n_regions) break;

        // ---- locate the region
        const uint32_t blk = rg / P.regions_per_block, rib = rg % P.regions_per_block;
        const uint32_t block_bytes = P.block_bytes ? P.block_bytes : (uint32_t)FOURMC_BLOCKSIZE;
        const uint64_t blk_off = (uint64_t)blk * block_bytes;
        const uint32_t blk_len = (uint32_t)min((uint64_t)block_bytes, P.n - blk_off);
        const uint32_t r_new = rib * P.region_bytes;               // first new byte, within the block
        if (r_new >= blk_len) {                                    // region beyond a short last block
            if (tid == 0) P.meta[rg] = RegionMeta{0, 0, 0, 0};
            __syncthreads();
            continue;
        }
        // the window starts `lb` bytes earlier (chain parse: look-back into the block's previous bytes);
        // positions below are relative to the window, new bytes are [lb, rlen)
        const int lb = CHAIN ? (int)min((uint32_t)ENC_CHAIN_LOOKBACK, r_new) : 0;
        const uint32_t r_off = r_new - (uint32_t)lb;
        const int rlen = lb + (int)min(P.region_bytes, blk_len - r_new);
        const uint8_t *gsrc = P.in + blk_off + r_off;
        // LZ4 end-of-block rules, relative to the region (native/lz4/lz4.c:243-247): the last
        // match starts at least 12 bytes before the block end and ends at least 5 before it.
        const int mf_limit = min(rlen - 1, (int)blk_len - 12 - (int)r_off);
        const int match_limit = min(rlen, (int)blk_len - 5 - (int)r_off);

        // ---- stage: bulk async copy of the 16-byte multiple, plain loads for the rest
        const int bulk = (((uintptr_t)gsrc & 15) == 0) ? (rlen & ~15) : 0;
        if (tid == 0 && bulk) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(bulk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(data)), "l"(gsrc), "r"(bulk), "r"(smem_u32(&s_bar)) : "memory");
        }
        for (int i = bulk + tid; i < rlen; i += NT) data[i] = gsrc[i];
        if (tid < ENC_PAD) data[rlen + tid] = 0;                   // reads past the end see zeros
        uint16_t *moff = table;                                     // CHAIN only: per new position, offset and length
        uint8_t *mlen = (uint8_t *)(table + ENC_CHAIN_REGION);      //   of the longest match found / of the parse's choice
        if constexpr (!CHAIN) { for (int i = tid; i < (1 << ENC_HASH_BITS) / 2; i += NT) ((uint32_t *)table)[i] = 0xffffffffu; }
        if (bulk) {
            uint32_t done = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\t"
                             "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                             "selp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&s_bar)), "r"(phase) : "memory");
            }
            phase ^= 1;
        }
        __syncthreads();

        // ---- index: first occurrence of every 4-byte hash.  Descending sweep, 4 consecutive
        // positions per thread (two word loads, three funnel shifts); within a thread and between
        // steps the lower position is stored last, within a step the order is left to the race.
        if constexpr (CHAIN) {
            // ---- search: the longest match of every new position, positions dealt out round robin
            const uint16_t *gp = P.chain + blk_off + r_off;         // gp[c]: link of window position c
            const int nnew = rlen - lb;
            for (int k = tid; k < nnew; k += NT) {
                const int q = lb + k;
                int best = 0, bo = 0;
                if (q <= mf_limit) {
                    const uint32_t v = smem_read4(data32, q);
                    const int maxlen = min(match_limit - q, ENC_CHAIN_LENCAP);
                    int c = q, kpos = 0;
                    int dl = (int)__ldg(gp + q);                    // the link to follow next; always requested one candidate ahead
                    for (int a = 0; a < P.depth; a++) {
                        if (!dl) break;
                        c -= dl;
                        if (c < 0 || q - c > 65535) break;
                        dl = (int)__ldg(gp + c + kpos);             // travels while this candidate is evaluated
                        if (smem_read4(data32, c) != v) continue;
                        if (best >= 4 && data[q + best] != data[c + best]) continue;      // cannot beat the best so far
                        int len = 4;
                        while (len < maxlen) {
                            const uint32_t x = smem_read4(data32, q + len) ^ smem_read4(data32, c + len);
                            if (x) { len += (__ffs(x) - 1) >> 3; break; }
                            len += 4;
                        }
                        len = min(len, maxlen);
                        if (len > best) {
                            best = len; bo = q - c;
                            if (len >= maxlen) break;
                            if (c + len <= q) {
                                // chain swap: the links of the match's first 4-byte windows, fetched as aligned pairs in one batch
                                const int ns = min(len - 3, ENC_CHAIN_SWAPSCAN);
                                const int c2 = c & ~1;               // window j sits in pair (c - c2 + j) >> 1
                                const uint32_t *gw = (const uint32_t *)(gp + c2);
                                uint32_t pr[ENC_CHAIN_SWAPSCAN / 2 + 1];
#pragma unroll
                                for (int u = 0; u < ENC_CHAIN_SWAPSCAN / 2 + 1; u++) pr[u] = 2 * u < ns + 1 ? __ldg(gw + u) : 0u;
                                int far = 1, kb = 0;
#pragma unroll
                                for (int u = 0; u < ENC_CHAIN_SWAPSCAN / 2 + 1; u++) {
                                    const int j0 = 2 * u - (c - c2), j1 = j0 + 1;
                                    const int d0 = (int)(pr[u] & 0xffffu), d1 = (int)(pr[u] >> 16);
                                    if (j0 >= 0 && j0 < ns && d0 > far) { far = d0; kb = j0; }
                                    if (j1 < ns && d1 > far) { far = d1; kb = j1; }
                                }
                                if (far > 1) { kpos = kb; dl = far; }
                            }
                        }
                    }
                }
                moff[k] = (uint16_t)bo;
                mlen[k] = (uint8_t)(best >= P.min_match ? best : 0);
            }
            __syncthreads();
        } else {
            const int last = rlen - 4;                             // last position with 4 bytes
            constexpr int STEP = NT * 4;
            for (int base = ((rlen - 1) / STEP) * STEP; base >= 0; base -= STEP) {
                const int p = base + 4 * tid;
                uint32_t h0 = 0, h2 = 0;
                const bool live0 = p <= last, live2 = p + 2 <= last;
                if (live0) {
                    // every second position is enough: a repeat first seen at an odd position is
                    // found one byte later and the backward extension recovers that byte
                    const uint32_t w0 = data32[p >> 2], w1 = data32[(p >> 2) + 1];
                    h0 = enc_hash(w0); h2 = enc_hash(__funnelshift_r(w0, w1, 16));
                    if (live2) table[h2] = (uint16_t)(p + 2);
                    table[h0] = (uint16_t)p;
                }
                // Equal hashes inside one step race; whoever finds a HIGHER position in its bucket stores again
                // until the lowest one holds it: the table -- and with it every compressed byte -- is then the
                // same on every run (one extra round on text, where a step rarely holds a hash twice).
                // (P.reproducible; costs a third of this kernel's time on text.  Without it any of the racing positions
                // stays -- all of them are legal candidates, only the bytes differ from run to run.)
                if (!P.reproducible) { __syncthreads(); continue; }
                for (;;) {
                    const bool lost0 = live0 && table[h0] > (uint16_t)p, lost2 = live2 && table[h2] > (uint16_t)(p + 2);
                    if (!__syncthreads_or(lost0 || lost2)) break;
                    if (lost2) table[h2] = (uint16_t)(p + 2);
                    if (lost0) table[h0] = (uint16_t)p;
                    __syncthreads();
                }
            }
        }

        // ---- parse: one slice per thread.  Inner sequences (all but the last) end inside the
        // slice and fit a packed word; the last one may be long and lives in registers.
        uint32_t rec[CHAIN ? ENC_CHAIN_MAXREC : ENC_MAXREC];
        int nrec = 0;
        int l_st = 0, l_len = 0, l_off = 0;                        // last sequence (l_len == 0: none)
        const int ss = CHAIN ? lb + tid * ENC_CHAIN_SLICE : tid * ENC_SLICE;
        if constexpr (CHAIN) {
            const bool own = ss < rlen;                             // 512 of the 1024 threads own a slice
            const int se = min(ss + ENC_CHAIN_SLICE, rlen);
            int end1 = 0;
            if (own) {
                // cost-optimal choices, back to front; cost[i]: encoded size (x16) of [i, se) minus the credit for bytes beyond se
                short cost[ENC_CHAIN_SLICE] = {};
                auto at = [&](int i) -> int { return i >= se ? -(i - se) * ENC_CHAIN_CREDIT : (int)cost[i - ss]; };
                for (int i = se - 1; i >= ss; i--) {
                    int bc = 16 + at(i + 1), bn = 0;
                    const int L = (int)mlen[i - lb];
                    if (L) {
                        const int cf = 16 * (3 + (L >= 19)) + at(i + L);            // lengths <= 255: one length byte from 19 on
                        if (cf <= bc) { bc = cf; bn = L; }
                        const int l0 = min(L - 1, se - i);
                        for (int l = l0; l >= P.min_match && l > l0 - ENC_CHAIN_SHORTER; l--) {
                            const int cc = 16 * (3 + (l >= 19)) + at(i + l);
                            if (cc < bc) { bc = cc; bn = l; }
                        }
                    }
                    cost[i - ss] = (short)bc;
                    mlen[i - lb] = (uint8_t)bn;
                }
                // walk 1: where do my choices end when I start at my own first byte (capped lengths not yet extended)
                for (int p = ss; p < se;) {
                    const int l = (int)mlen[p - lb];
                    if (l) { p += l; end1 = p; } else p++;
                }
            }
            const int cov1 = cta_excl_scan<true, NW>(end1, s_scan, nullptr);
            if (own) {
                // walk 2: from the first byte earlier slices leave to me
                for (int p = max(ss, cov1); p < se;) {
                    int len = (int)mlen[p - lb];
                    if (!len) { p++; continue; }
                    const int off = (int)moff[p - lb];
                    if (len == ENC_CHAIN_LENCAP) {
                        const int maxlen = match_limit - p;
                        while (len < maxlen) {
                            const uint32_t x = smem_read4(data32, p + len) ^ smem_read4(data32, p + len - off);
                            if (x) { len += (__ffs(x) - 1) >> 3; break; }
                            len += 4;
                        }
                        len = min(len, maxlen);
                    }
                    if (l_len) rec[nrec++] = enc_pack(l_st - ss, l_len, l_off);
                    l_st = p; l_len = len; l_off = off;
                    p += len;
                }
            }
        }
        if (!CHAIN && ss < rlen) {
            const int se = min(ss + ENC_SLICE, rlen);
            // pass 1 -- every position of the slice, no skipping: does the table hold an earlier
            // position with the same four bytes?  One bit per position (uniform work for all lanes).
            unsigned long long cand[3] = {0ull, 0ull, 0ull};
            {
                // One slice word per step = four positions with compile-time byte shifts (17 instructions per position
                // instead of 37 for the position-by-position loop, r02 SASS); positions at or after `stop` are
                // masked off at the end (their lookups read table entries and region bytes that exist).
                const int stop_rel = min(se, mf_limit + 1) - ss;
                const uint32_t *sw = data32 + (ss >> 2);                           // ss is word aligned
                uint32_t w_lo = sw[0];
                uint32_t half[5] = {0u, 0u, 0u, 0u, 0u};                           // 8 words = 32 positions each
                auto probe = [&](const uint32_t v, const int p) -> bool {
                    const int c = (int)table[enc_hash(v)];
                    // filter on the ONE word that holds the candidate's first byte (4 - (c & 3) of the four
                    // bytes): half the scattered loads of a full compare; pass 2 checks the rest
                    const uint32_t cw = data32[c >> 2];
                    const int cs = (c & 3) * 8;
                    return c < p && ((((cw >> cs) ^ v) << cs) == 0);
                };
#pragma unroll
                for (int h = 0; h < 5; h++) {
                    const int nw = h < 4 ? 8 : 1;                                  // 33 words per slice
                    uint32_t m = 0;
                    for (int j = 0; j < nw; j++) {
                        const int wi = h * 8 + j;
                        const uint32_t w_hi = sw[wi + 1];
                        const int p = ss + 4 * wi;
                        uint32_t nib = 0;
                        if (probe(w_lo, p)) nib |= 1u;
                        if (probe(__funnelshift_r(w_lo, w_hi, 8), p + 1)) nib |= 2u;
                        if (probe(__funnelshift_r(w_lo, w_hi, 16), p + 2)) nib |= 4u;
                        if (probe(__funnelshift_r(w_lo, w_hi, 24), p + 3)) nib |= 8u;
                        m |= nib << (4 * j);
                        w_lo = w_hi;
                    }
                    half[h] = m;
                }
                cand[0] = (unsigned long long)half[0] | ((unsigned long long)half[1] << 32);
                cand[1] = (unsigned long long)half[2] | ((unsigned long long)half[3] << 32);
                cand[2] = (unsigned long long)half[4];
#pragma unroll
                for (int g = 0; g < 3; g++) {
                    const int k = stop_rel - g * 64;                               // positions of this group before `stop`
                    if (k <= 0) cand[g] = 0ull; else if (k < 64) cand[g] &= (1ull << k) - 1ull;
                }
            }
            // pass 2 -- the greedy walk only visits positions whose bit is set
            int p = ss, anchor = ss;
            for (;;) {
                // next candidate at or after p
                int rel = p - ss, nxt = -1;
#pragma unroll
                for (int g = 0; g < 3; g++) {
                    unsigned long long m = cand[g];
                    const int sh = rel - g * 64;
                    if (sh >= 64) m = 0; else if (sh > 0) m &= ~0ull << sh;
                    if (nxt < 0 && m) nxt = g * 64 + __ffsll((long long)m) - 1;
                }
                if (nxt < 0) break;
                p = ss + nxt;
                const int c = (int)table[enc_hash(smem_read4(data32, p))];
                int len = (c & 3) ? 0 : 4;                    // pass 1 compared all four bytes only for aligned candidates
                const int maxlen = match_limit - p;
                while (len < maxlen) {
                    const uint32_t x = smem_read4(data32, p + len) ^ smem_read4(data32, c + len);
                    if (x) { len += (__ffs(x) - 1) >> 3; break; }
                    len += 4;
                }
                len = min(len, maxlen);
                if (len < 4) { p++; continue; }                // the filter let a partial match through
                int st = p, m = c;
                while (st > anchor && m > 0 && data[st - 1] == data[m - 1]) { st--; m--; len++; }
                if (len >= P.min_match) {
                    if (l_len) rec[nrec++] = enc_pack(l_st - ss, l_len, l_off);
                    l_st = st; l_len = len; l_off = st - m;
                    p = st + len; anchor = p;
                } else p++;
            }
        }

        // ---- stitch 1: trim against everything earlier slices cover
        const int cov = cta_excl_scan<true, NW>(l_len ? l_st + l_len : 0, s_scan, nullptr);
        int surv_end = 0;                                           // end of my last surviving sequence
        {
            int w = 0;
            for (int k = 0; k < nrec; k++) {
                const uint32_t r = rec[k];
                int st = ss + (int)(r >> 24), len = (int)((r >> 16) & 255);
                const int end = st + len;
                if (st < cov) { len = end - cov; st = cov; }
                if (len < 4 || st > mf_limit) continue;             // dropped: its bytes become literals
                rec[w++] = enc_pack(st - ss, len, (int)(r & 0xffffu));
                surv_end = end;
            }
            nrec = w;
            if (l_len) {
                const int end = l_st + l_len;
                if (l_st < cov) { l_len = end - cov; l_st = cov; }
                if (l_len < 4 || l_st > mf_limit) l_len = 0; else surv_end = end;
            }
        }
        // ---- stitch 2: where does my first literal run start; encoded size of my sequences
        int total_anchor;
        int anchor0 = cta_excl_scan<true, NW>(surv_end, s_scan, &total_anchor);
        anchor0 = max(anchor0, lb);                                 // the first literal run starts at the first NEW byte
        total_anchor = max(total_anchor, lb);
        int bytes = 0;
        {
            int a = anchor0;
            for (int k = 0; k < nrec; k++) {
                const uint32_t r = rec[k];
                const int st = ss + (int)(r >> 24), len = (int)((r >> 16) & 255);
                const int lit = st - a;
                bytes += 1 + enc_ext_bytes(lit) + lit + 2 + enc_ext_bytes(len - 4);
                a = st + len;
            }
            if (l_len) {
                const int lit = l_st - a;
                bytes += 1 + enc_ext_bytes(lit) + lit + 2 + enc_ext_bytes(l_len - 4);
            }
        }
        const int myseq = nrec + (l_len ? 1 : 0);
        int total_bytes, total_seq;
        const int out_off = cta_excl_scan<false, NW>(bytes, s_scan, &total_bytes);
        const int seq_before = cta_excl_scan<false, NW>(myseq, s_scan, &total_seq);

        uint8_t *slot = P.scratch + (size_t)rg * P.slot_bytes;
        if constexpr (ZSEQ) {
            // ---- emit sequence arrays + gathered literals straight to HBM
            int mbytes = l_len;
            for (int k = 0; k < nrec; k++) mbytes += (int)((rec[k] >> 16) & 255);
            int total_matched;
            const int matched_before = cta_excl_scan<false, NW>(mbytes, s_scan, &total_matched);
            const uint32_t stride = ((uint32_t)total_seq + 7u) & ~7u;
            uint16_t *zll = (uint16_t *)slot, *zml = zll + stride, *zoff = zml + stride;
            uint8_t *zlit = (uint8_t *)(zoff + stride);
            int long_n = 0, long_src = 0;
            uint8_t *long_dst = nullptr;
            {
                int a = anchor0, lp = anchor0 - lb - matched_before, idx = seq_before;
                for (int k = 0; k <= nrec; k++) {
                    int st, len, off;
                    if (k < nrec) {
                        const uint32_t r = rec[k];
                        st = ss + (int)(r >> 24); len = (int)((r >> 16) & 255); off = (int)(r & 0xffffu);
                    } else {
                        if (!l_len) break;
                        st = l_st; len = l_len; off = l_off;
                    }
                    const int lit = st - a;
                    zll[idx] = (uint16_t)lit; zml[idx] = (uint16_t)len; zoff[idx] = (uint16_t)off;
                    if (lit > ENC_SLICE) { long_n = lit; long_src = a; long_dst = zlit + lp; }
                    else copy_batched(zlit + lp, data + a, lit);
                    lp += lit; a = st + len; idx++;
                }
            }
            for (unsigned mm = __ballot_sync(FM_FULL, long_n > 0); mm; mm &= mm - 1) {
                const int l = __ffs(mm) - 1;
                const int n = __shfl_sync(FM_FULL, long_n, l);
                const int sp = __shfl_sync(FM_FULL, long_src, l);
                uint8_t *dp = (uint8_t *)__shfl_sync(FM_FULL, (unsigned long long)long_dst, l);
                for (int i = lane; i < n; i += 32) dp[i] = data[sp + i];
            }
            {   // literals after the region's last match
                uint8_t *dp = zlit + (total_anchor - lb - total_matched);
                for (int i = total_anchor + tid; i < rlen; i += NT) dp[i - total_anchor] = data[i];
            }
            total_bytes = rlen - lb - total_matched;
        } else {
        // ---- emit into the staging area (the hash table is dead now), flushed below with 128-bit
        // stores; sequences beyond its capacity (poorly compressible regions) go straight to HBM.
        // Only a slice's first literal run can be long (it may reach back over match-free slices);
        // those are copied by the whole warp afterwards.
        uint8_t *stage = (uint8_t *)table;
        __syncthreads();                                            // every thread is done with the table
        constexpr int STAGE_CAP = CHAIN ? ENC_STAGE_CHAIN : ENC_STAGE;
        const bool staged = out_off + bytes <= STAGE_CAP;
        if (tid == 0) s_flush = min(total_bytes, STAGE_CAP);
        __syncthreads();
        if (!staged && bytes > 0) atomicMin(&s_flush, out_off);     // the staged prefix ends at the first direct writer
        int long_n = 0, long_src = 0;
        uint8_t *long_dst = nullptr;
        {
            uint8_t *o = (staged ? stage : slot) + out_off;
            int a = anchor0;
            for (int k = 0; k <= nrec; k++) {
                int st, len, off;
                if (k < nrec) {
                    const uint32_t r = rec[k];
                    st = ss + (int)(r >> 24); len = (int)((r >> 16) & 255); off = (int)(r & 0xffffu);
                } else {
                    if (!l_len) break;
                    st = l_st; len = l_len; off = l_off;
                }
                const int lit = st - a, ml = len - 4;
                if (k == 0 && seq_before == 0) P.meta[rg].lead = (uint32_t)lit;
                *o++ = (uint8_t)((min(lit, 15) << 4) | min(ml, 15));
                if (lit >= 15) o = emit_len(o, lit - 15);
                if (lit > ENC_SLICE) { long_n = lit; long_src = a; long_dst = o; }
                else copy_batched(o, data + a, lit);
                o += lit;
                *o++ = (uint8_t)(off & 0xff); *o++ = (uint8_t)(off >> 8);
                if (ml >= 15) o = emit_len(o, ml - 15);
                a = st + len;
            }
        }
        for (unsigned mm = __ballot_sync(FM_FULL, long_n > 0); mm; mm &= mm - 1) {
            const int l = __ffs(mm) - 1;
            const int n = __shfl_sync(FM_FULL, long_n, l);
            const int sp = __shfl_sync(FM_FULL, long_src, l);
            uint8_t *dp = (uint8_t *)__shfl_sync(FM_FULL, (unsigned long long)long_dst, l);
            for (int i = lane; i < n; i += 32) dp[i] = data[sp + i];
        }
        __syncthreads();
        {
            const int nflush = s_flush;
            const uint4 *sv = (const uint4 *)stage;
            uint4 *dv = (uint4 *)slot;
            for (int i = tid; i < (nflush >> 4); i += NT) dv[i] = sv[i];
            for (int i = (nflush & ~15) + tid; i < nflush; i += NT) slot[i] = stage[i];
        }
        }
        if (tid == 0) {
            RegionMeta *mt = &P.meta[rg];
            mt->body_bytes = (uint32_t)total_bytes;
            mt->tail_lits = (uint32_t)(rlen - total_anchor);
            mt->nseq = (uint32_t)total_seq;
            if (total_seq == 0) mt->lead = 0;
        }
        __syncthreads();        // smem is reused by the next region
    }
}

// ---- E2 / E3 ---------------------------------------------------------------------------------

struct BlockPlan {              // produced by E2, consumed by the index scan and E3
    uint32_t usize;
    uint32_t payload;           // csize, or usize when stored
    uint32_t stored;
    uint32_t final_lits;        // literals of the closing sequence
};

// raw_limit < 0: container mode, a block is stored when its compressed size reaches its raw size.
// raw_limit >= 0: bare LZ4 block for the per-block API; "stored" then means "does not fit in
// raw_limit bytes" (LZ4_compress_default returns 0, native/lz4/lz4.c:1290-1300).
__global__ void lz4_block_size_kernel(const RegionMeta *meta, uint32_t n_blocks, uint64_t n,
                                      BlockPlan *plan, uint32_t *block_lens, int64_t raw_limit, uint32_t regions_per_block,
                                      uint32_t block_bytes = FOURMC_BLOCKSIZE)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint64_t blk_off = (uint64_t)b * block_bytes;
    const uint32_t u = (uint32_t)min((uint64_t)block_bytes, n - blk_off);
    const RegionMeta *m = meta + (size_t)b * regions_per_block;
    uint32_t carry = 0, c = 0;
    for (uint32_t r = 0; r < regions_per_block; r++) {
        const RegionMeta x = m[r];
        if (x.nseq == 0) { carry += x.tail_lits; continue; }
        const int old_hdr = 1 + enc_ext_bytes((int)x.lead);
        const int new_hdr = 1 + enc_ext_bytes((int)(x.lead + carry));
        c += (uint32_t)new_hdr + carry + (x.body_bytes - (uint32_t)old_hdr);
        carry = x.tail_lits;
    }
    c += 1u + (uint32_t)enc_ext_bytes((int)carry) + carry;
    BlockPlan p;
    p.usize = u;
    if (raw_limit >= 0) {
        p.stored = ((int64_t)c > raw_limit) ? 1u : 0u;
        p.payload = p.stored ? 0u : c;
    } else {
        p.stored = (c >= u) ? 1u : 0u;      // the reference offers u-1 bytes (native/4mc.c:301)
        p.payload = p.stored ? u : c;
    }
    p.final_lits = carry;
    plan[b] = p;
    if (block_lens) block_lens[b] = 12u + p.payload;
}

// CTA-cooperative copy, arbitrary alignment: byte head, 16-byte stores with funnel-shifted loads.
__device__ __forceinline__ void cta_copy(uint8_t *d, const uint8_t *s, uint32_t n)
{
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    if (n < 64) { for (uint32_t i = tid; i < n; i += nth) d[i] = s[i]; return; }
    const uint32_t head = (uint32_t)((16 - ((uintptr_t)d & 15)) & 15);
    for (uint32_t i = tid; i < head; i += nth) d[i] = s[i];
    const uint8_t *s2 = s + head;
    const uint32_t sh = (uint32_t)((uintptr_t)s2 & 3) * 8;
    const uint32_t *sw = (const uint32_t *)((uintptr_t)s2 & ~(uintptr_t)3);
    const uint32_t body = (n - head - 4) >> 4;              // keep the funnel's extra word in range
    uint4 *dv = (uint4 *)(d + head);
    for (uint32_t i = tid; i < body; i += nth) {
        const uint32_t a = sw[4 * i], b = sw[4 * i + 1], c = sw[4 * i + 2], e = sw[4 * i + 3], f = sw[4 * i + 4];
        dv[i] = make_uint4(__funnelshift_r(a, b, sh), __funnelshift_r(b, c, sh),
                           __funnelshift_r(c, e, sh), __funnelshift_r(e, f, sh));
    }
    for (uint32_t i = head + (body << 4) + tid; i < n; i += nth) d[i] = s[i];
}

constexpr int ENC_WRITE_THREADS = 256;

// out_base + block_off[b] is where block b's 12-byte header goes.
__global__ void __launch_bounds__(ENC_WRITE_THREADS)
lz4_block_write_kernel(const uint8_t *in, const uint8_t *scratch, const RegionMeta *meta,
                       const BlockPlan *plan, const uint64_t *block_off, uint8_t *out_base, int raw_mode,
                       uint32_t regions_per_block, uint32_t region_bytes, uint32_t slot_bytes,
                       uint32_t block_bytes = FOURMC_BLOCKSIZE)
{
    __shared__ __align__(16) uint32_t s_stage[XXH_WARP_SMEM_WORDS];
    __shared__ uint32_t s_dst[ENC_MAX_REGIONS_PER_BLOCK + 1];   // payload offset of each region's piece
    __shared__ uint32_t s_carry[ENC_MAX_REGIONS_PER_BLOCK + 1]; // literals carried INTO the region

    const uint32_t b = blockIdx.x;
    const BlockPlan p = plan[b];
    const uint64_t blk_off = (uint64_t)b * block_bytes;
    const uint8_t *src = in + blk_off;
    uint8_t *rec = out_base + block_off[b];
    uint8_t *pay = rec + 12;
    const RegionMeta *m = meta + (size_t)b * regions_per_block;
    const int RPB = (int)regions_per_block;

    if (p.stored) {
        if (raw_mode) return;               // nothing to write: the caller reports 0
        cta_copy(pay, src, p.usize);
    } else {
        if (threadIdx.x == 0) {
            uint32_t carry = 0, c = 0;
            for (int r = 0; r < RPB; r++) {
                const RegionMeta x = m[r];
                s_dst[r] = c; s_carry[r] = carry;
                if (x.nseq == 0) { carry += x.tail_lits; continue; }
                const int old_hdr = 1 + enc_ext_bytes((int)x.lead);
                const int new_hdr = 1 + enc_ext_bytes((int)(x.lead + carry));
                c += (uint32_t)new_hdr + carry + (x.body_bytes - (uint32_t)old_hdr);
                carry = x.tail_lits;
            }
            s_dst[RPB] = c; s_carry[RPB] = carry;
        }
        __syncthreads();
        for (int r = 0; r < RPB; r++) {
            const RegionMeta x = m[r];
            if (x.nseq == 0) continue;
            const uint32_t carry = s_carry[r];
            const uint8_t *slot = scratch + ((size_t)b * RPB + r) * slot_bytes;
            uint8_t *o = pay + s_dst[r];
            const int old_hdr = 1 + enc_ext_bytes((int)x.lead);
            const int lit = (int)(x.lead + carry);
            const int new_hdr = 1 + enc_ext_bytes(lit);
            if (threadIdx.x == 0) {
                uint8_t *q = o;
                *q++ = (uint8_t)((min(lit, 15) << 4) | (slot[0] & 15));
                if (lit >= 15) emit_len(q, lit - 15);
            }
            // carried literals are input bytes just before the region
            cta_copy(o + new_hdr, src + (size_t)r * region_bytes - carry, carry);
            cta_copy(o + new_hdr + carry, slot + old_hdr, x.body_bytes - (uint32_t)old_hdr);
        }
        {   // closing sequence: literals only
            const uint32_t carry = s_carry[RPB];
            uint8_t *o = pay + s_dst[RPB];
            const int hdr = 1 + enc_ext_bytes((int)carry);
            if (threadIdx.x == 0) {
                uint8_t *q = o;
                *q++ = (uint8_t)(min((int)carry, 15) << 4);
                if (carry >= 15) emit_len(q, (int)carry - 15);
            }
            cta_copy(o + hdr, src + p.usize - carry, carry);
        }
    }
    __threadfence_block();
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t h = xxh32_warp<false>(pay, p.payload, 0, s_stage);
        if (threadIdx.x == 0) { st_be32(rec, p.usize); st_be32(rec + 4, p.payload); st_be32(rec + 8, h); }
    }
}

}  // namespace fm
