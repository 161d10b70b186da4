// capi.cu -- the C-ABI of lib4mcgpu.so (include/fourmc.h): contexts, workspaces, kernel launches.
// Host code here is plumbing only; every byte of codec / checksum / index arithmetic runs in the
// kernels of xxh32.cuh, lz4_decode.cuh, lz4_encode.cuh and container.cuh.  There is no CPU codec
// in this library: without a CUDA device every entry point fails with FOURMC_E_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <thread>
#include <vector>

#include "blockstream.h"
#include "../../include/fourmc.h"
#include "container.cuh"
#include "fourmc_gen.h"
#include "lz4_decode.cuh"
#include "lz4_encode.cuh"
#include "xxh32.cuh"
#include "zstd_decode.cuh"
#include "zstd_encode.cuh"

using namespace fm;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct EncWs {
    DevBuf scratch, meta, plan, lens, off, misc;   // misc: [0] work counter (u32), [8] span (u64), [16] total (u64), [24] carry (u64)
    DevBuf zout, zrout;                            // 4mz: entropy-stage output slots and their sizes
    DevBuf chain;                                  // levels 2..4: chain links of the blocks in flight (2 bytes per input byte)
};

struct DecWs {
    DevBuf desc, xxh, status, tokmap, chunkop, result, info, tables, outsize, final_, zwork, zlane;
    // the checksum pass runs beside the parse on its own stream (both only read the payloads)
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};

}  // namespace

constexpr int FM_PIPE_MAX = 8;

struct fourmc_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;       // default stream of the context
    cudaStream_t aux[FM_PIPE_MAX] = {};     // host-buffer pipelines: one stream + buffer set per slice in flight
    cudaEvent_t ev[FM_PIPE_MAX] = {};
    std::string err;
    uint64_t launches = 0;
    // compressed bytes identical from run to run (ties between racing match-finder stores settled by position):
    // -1 = per entry point (host / file / per-block calls: yes; device-resident calls: no, they favour speed), 0 / 1 = always
    int reproducible = -1;
    int repro_call = 1;                  // what the entry point in progress resolved it to
    // the block index of the file the split reader saw last (a reader works through many splits of one file)
    const void *idx_file = nullptr;
    size_t idx_size = 0;
    uint8_t idx_tail[12] = {};
    std::vector<int64_t> idx_offs;
    EncWs enc[FM_PIPE_MAX];
    DecWs dec[FM_PIPE_MAX];
    DevBuf stage_in[FM_PIPE_MAX], stage_out[FM_PIPE_MAX];    // device staging for the host-pointer entry points
    void *pinned = nullptr;              // small pinned scratch for scalars
    size_t pinned_cap = 0;
    void *pin_up[2] = {nullptr, nullptr};   // pinned bounce buffers for uploads from pageable host memory (upload_pieces), per staging slot
    size_t pin_up_cap[2] = {0, 0};
    bool region_attr_set = false, chain_attr_set = false, d1_attr_set = false, d1w_attr_set = false, zd_attr_set = false, gen_attr_set = false;   // per context = per device
    DevBuf ztables;                      // fmz::Tables (constant decode tables), uploaded once
    // optional per-kernel timing (fourmc_timing_enable): CUDA event pairs around every launch
    bool timing = false;
    std::vector<cudaEvent_t> tev;        // pool: [2i] start, [2i+1] stop
    std::vector<const char *> tname;     // kernel name of pair i
    size_t tused = 0;
};

namespace {

int fail(fourmc_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess)
{
    if (c) {
        c->err = what;
        if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
    }
    return code;
}

#define CK(call)                                                                      \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) return fail(ctx, FOURMC_E_CUDA, #call, e__);          \
    } while (0)

// development aid (FOURMC_PROFILE=1): serialises every launch and prints its wall time
bool profile_mode()
{
    static int v = -1;
    if (v < 0) v = getenv("FOURMC_PROFILE") ? 1 : 0;
    return v == 1;
}
void profile_mark(const char *what)
{
    static double last = 0;
    cudaDeviceSynchronize();
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    const double now = ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
    fprintf(stderr, "[fourmc profile] %-28s %9.3f ms\n", what, last ? now - last : 0.0);
    last = now;
}

#define CKL(what)                                                                     \
    do {                                                                              \
        ctx->launches++;                                                              \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) return fail(ctx, FOURMC_E_CUDA, what, e__);           \
        if (profile_mode()) profile_mark(what);                                       \
    } while (0)

void ktime_mark(fourmc_ctx *ctx, const char *name, cudaStream_t st, int which)
{
    if (!ctx->timing) return;
    if (which == 0) {
        if (ctx->tused * 2 + 2 > ctx->tev.size()) {
            cudaEvent_t a = nullptr, b = nullptr;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) {
                // no events: stop timing for this context rather than leave the closing mark without a slot
                if (a) cudaEventDestroy(a);
                (void)cudaGetLastError();
                ctx->timing = 0;
                return;
            }
            ctx->tev.push_back(a); ctx->tev.push_back(b); ctx->tname.push_back(name);
        }
        ctx->tname[ctx->tused] = name;
        cudaEventRecord(ctx->tev[ctx->tused * 2], st);
    } else {
        cudaEventRecord(ctx->tev[ctx->tused * 2 + 1], st);
        ctx->tused++;
    }
}

// KL(name, stream, kernel<<<...>>>(...)): launch with optional event timing and error check
#define KL(name, st, ...)                                                             \
    do {                                                                              \
        ktime_mark(ctx, name, st, 0);                                                 \
        __VA_ARGS__;                                                                  \
        ktime_mark(ctx, name, st, 1);                                                 \
        CKL(name);                                                                    \
    } while (0)

int ensure(fourmc_ctx *ctx, DevBuf &b, size_t need)
{
    if (b.cap >= need && b.p) return FOURMC_OK;
    if (b.p) { CK(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    // 6 % of headroom: several workspaces are sized from compressed sizes, which differ a little from call to
    // call (and from run to run: the encoder's index is filled by racing stores); without it a batch that is a
    // few bytes larger than every earlier one costs a cudaFree + cudaMalloc of gigabytes in the middle of a call
    size_t cap = (need + need / 16 + 255) & ~(size_t)255;
    if (cap < 256) cap = 256;
    CK(cudaMalloc(&b.p, cap));
    b.cap = cap;
    return FOURMC_OK;
}

void release(DevBuf &b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
}

inline cudaStream_t pick(fourmc_ctx *ctx, void *stream) { return stream ? (cudaStream_t)stream : ctx->stream; }

inline uint32_t blocks_of(size_t n) { return (uint32_t)((n + FOURMC_BLOCKSIZE - 1) / FOURMC_BLOCKSIZE); }

int slice_blocks()
{
    static int v = 0;
    if (!v) {
        const char *e = getenv("FOURMC_SLICE_BLOCKS");
        v = e ? atoi(e) : 128;
        if (v < 1) v = 1;
        if (v > 4096) v = 4096;
    }
    return v;
}

// The 4mc host writer is bound by the upload (PCIe): smaller slices shorten the pipeline's fill and
// drain (B200: 50.3 GB/s at 32 blocks, 47.8 at 128); the reader's kernels want the larger slices.
int cslice_blocks()
{
    static int v = 0;
    if (!v) {
        const char *e = getenv("FOURMC_CSLICE_BLOCKS");
        v = e ? atoi(e) : 32;
        if (v < 1) v = 1;
        if (v > 4096) v = 4096;
    }
    return v;
}

int zslice_blocks()
{
    static int v = 0;
    if (!v) {
        const char *e = getenv("FOURMC_ZSLICE_BLOCKS");
        v = e ? atoi(e) : 384;
        if (v < 1) v = 1;
        if (v > 4096) v = 4096;
    }
    return v;
}

int pipe_depth()
{
    static int v = 0;
    if (!v) {
        const char *e = getenv("FOURMC_PIPE");
        v = e ? atoi(e) : 4;
        if (v < 2) v = 2;
        if (v > FM_PIPE_MAX) v = FM_PIPE_MAX;
    }
    return v;
}

// threads per CTA of the block write kernels (one CTA per block: copies, then one warp hashes the payload)
int write_threads()
{
    static int v = 0;
    if (!v) {
        const char *e = getenv("FOURMC_WRITE_THREADS");
        v = e ? atoi(e) : ENC_WRITE_THREADS;
        if (v != 64 && v != 128 && v != 256) v = ENC_WRITE_THREADS;
    }
    return v;
}

int level_min_match(int level)
{
    // level 1 (Fast): minimum match 5 gives fewer, longer sequences at the same ratio on text
    // (DESIGN.md); the chain parse of levels 2..4 takes every match of 4 or more.
    const char *e = getenv("FOURMC_MIN_MATCH");
    if (e) { int v = atoi(e); if (v >= 4 && v <= 16) return v; }
    return level >= 2 ? 4 : 5;
}

// Levels as in native/4mc.c:243-253 (1 fast, 2 medium = LZ4 MC, 3 high = HC 4, 4 ultra = HC 8; the
// 4mz twins :415-425): candidates tried per chain search; 0 selects the Fast parse.  32 / 64 candidates with the
// cost-optimal parse reach the ratios of the reference's HC 4 / HC 8 (16 / 256 attempts, lz4hc.c:75-90 of its level
// table) on the bench text: 2.67 / 2.69 against 2.66 / 2.69 (tools/hc_study.cpp).
int level_chain_depth(int level)
{
    const char *e = getenv("FOURMC_CHAIN_DEPTH");
    if (e) { int v = atoi(e); if (v >= 0 && v <= 4096) return v; }
    return level <= 1 ? 0 : level == 2 ? 4 : level == 3 ? 32 : 128;
}

enum { CODEC_LZ4 = 0, CODEC_ZSTD = 1 };

int ensure_ztables(fourmc_ctx *ctx)
{
    if (ctx->ztables.p) return FOURMC_OK;
    int r;
    if ((r = ensure(ctx, ctx->ztables, sizeof(fmz::Tables)))) return r;
    fmz::Tables T;
    fmz::make_tables(T);                 // format constants only (base values, extra bits, default distributions)
    CK(cudaMemcpy(ctx->ztables.p, &T, sizeof(T), cudaMemcpyHostToDevice));
    return FOURMC_OK;
}

// ---- encode --------------------------------------------------------------------------------

// Levels 2..4: blocks of one launch of the region kernel share a chain-link buffer (2 bytes per input byte), so large
// inputs are parsed in groups of blocks.
uint32_t chain_group_blocks(uint32_t nb)
{
    const char *e = getenv("FOURMC_CHAIN_GROUP");  // read per call: the tests shrink it to exercise the loop over groups
    int v = e ? atoi(e) : 256;
    if (v < 1) v = 1;
    if (v > 4096) v = 4096;
    return std::min<uint32_t>(nb, (uint32_t)v);
}

// chain links of `gb` blocks starting at d_in (gn bytes), then nothing else: the caller launches the region kernel
int launch_chain_links(fourmc_ctx *ctx, cudaStream_t st, EncWs &ws, const uint8_t *d_in, size_t gn, uint32_t gb, uint32_t block_bytes)
{
    if (!ctx->chain_attr_set) {
        CK(cudaFuncSetAttribute(lz4_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_LINK_SMEM));
        ctx->chain_attr_set = true;
    }
    // chunk size: enough warps to fill the GPU three times over when the group is small, little warm-up overhead when it is large
    uint32_t chunk = 1u << 20;
    const uint64_t want = 9ull * (uint64_t)ctx->sm_count;
    while (chunk > (1u << 16) && (uint64_t)gb * ((block_bytes + chunk - 1) / chunk) < want) chunk >>= 1;
    const uint32_t cpb = (block_bytes + chunk - 1) / chunk;
    const uint32_t items = gb * cpb;
    const uint32_t grid = std::min<uint32_t>(items, 3u * (uint32_t)ctx->sm_count);
    KL("lz4_chain_kernel", st, lz4_chain_kernel<<<grid, 32, ENC_LINK_SMEM, st>>>(d_in, gn, block_bytes, chunk, cpb, items,
                                                                                       (uint16_t *)ws.chain.p));
    return FOURMC_OK;
}

// d_in[0..n) -> block records back to back at d_span (block b at d_span + off[b], off[0] = base).
// raw_limit >= 0 selects the bare-block mode of the per-block API (single block, no header use).
int enc_span(fourmc_ctx *ctx, cudaStream_t st, EncWs &ws, int level, const uint8_t *d_in, size_t n,
             uint8_t *d_out_base, uint64_t base, uint32_t *d_block_lens_out, int64_t raw_limit,
             uint32_t block_bytes = FOURMC_BLOCKSIZE)
{
    const uint32_t nb = (uint32_t)((n + block_bytes - 1) / block_bytes);
    if (nb == 0) {
        int r;
        if ((r = ensure(ctx, ws.misc, 64))) return r;
        CK(cudaMemsetAsync(ws.misc.p, 0, 64, st));
        return FOURMC_OK;
    }
    // geometry: 64 regions of 64 KiB (Fast parse) or 128 regions of 32 KiB behind 32 KiB of look-back (chain parse)
    const int depth = level_chain_depth(level);
    const uint32_t region_bytes = depth > 0 ? ENC_CHAIN_REGION : ENC_REGION;
    const uint32_t rpb = FOURMC_BLOCKSIZE / region_bytes;
    const uint32_t slot_bytes = depth > 0 ? ENC_CHAIN_SLOT : ENC_SLOT;
    const uint32_t nreg = nb * rpb;
    int r;
    if ((r = ensure(ctx, ws.scratch, (size_t)nreg * slot_bytes))) return r;
    if ((r = ensure(ctx, ws.meta, (size_t)nreg * sizeof(RegionMeta)))) return r;
    if ((r = ensure(ctx, ws.plan, (size_t)nb * sizeof(BlockPlan)))) return r;
    if ((r = ensure(ctx, ws.lens, (size_t)nb * 4))) return r;
    if ((r = ensure(ctx, ws.off, (size_t)nb * 8))) return r;
    if ((r = ensure(ctx, ws.misc, 64))) return r;
    if (!ctx->region_attr_set) {
        CK(cudaFuncSetAttribute(lz4_region_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_SMEM));
        CK(cudaFuncSetAttribute(lz4_region_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_SMEM));
        CK(cudaFuncSetAttribute(lz4_region_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_SMEM_CHAIN));
        CK(cudaFuncSetAttribute(lz4_region_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_SMEM_CHAIN));
        ctx->region_attr_set = true;
    }
    CK(cudaMemsetAsync(ws.misc.p, 0, 64, st));
    EncParams P = {};
    P.in = d_in; P.n = n; P.n_regions = nreg;
    P.scratch = (uint8_t *)ws.scratch.p; P.meta = (RegionMeta *)ws.meta.p;
    P.work_counter = (uint32_t *)ws.misc.p;
    P.min_match = level_min_match(level);
    P.slot_bytes = slot_bytes;
    P.depth = depth;
    P.region_bytes = region_bytes; P.regions_per_block = rpb;
    P.block_bytes = block_bytes;
    P.reproducible = ctx->repro_call;
    if (P.depth > 0) {
        const uint32_t G = chain_group_blocks(nb);
        if ((r = ensure(ctx, ws.chain, (size_t)G * block_bytes * 2))) return r;
        for (uint32_t g0 = 0; g0 < nb; g0 += G) {
            const uint32_t gb = std::min<uint32_t>(G, nb - g0);
            const size_t goff = (size_t)g0 * block_bytes;
            const size_t gn = std::min<size_t>(n - goff, (size_t)gb * block_bytes);
            if (g0) CK(cudaMemsetAsync(ws.misc.p, 0, 4, st));
            if ((r = launch_chain_links(ctx, st, ws, d_in + goff, gn, gb, block_bytes))) return r;
            P.in = d_in + goff; P.n = gn; P.n_regions = gb * rpb;
            P.scratch = (uint8_t *)ws.scratch.p + (size_t)g0 * rpb * slot_bytes;
            P.meta = (RegionMeta *)ws.meta.p + (size_t)g0 * rpb;
            P.chain = (const uint16_t *)ws.chain.p;
            const uint32_t grid = std::min<uint32_t>(P.n_regions, (uint32_t)ctx->sm_count);
            KL("lz4_region_chain_kernel", st, lz4_region_kernel<false, true><<<grid, ENC_CHAIN_THREADS, ENC_SMEM_CHAIN, st>>>(P));
        }
    } else {
        const uint32_t grid = std::min<uint32_t>(nreg, 2u * (uint32_t)ctx->sm_count);
        KL("lz4_region_kernel", st, lz4_region_kernel<false, false><<<grid, ENC_THREADS, ENC_SMEM, st>>>(P));
    }
    uint32_t *lens = d_block_lens_out ? d_block_lens_out : (uint32_t *)ws.lens.p;
    KL("lz4_block_size_kernel", st, lz4_block_size_kernel<<<(nb + 127) / 128, 128, 0, st>>>((const RegionMeta *)ws.meta.p, nb, n,
                                                           (BlockPlan *)ws.plan.p, lens, raw_limit, rpb, block_bytes));
    KL("scan_lens_kernel", st, scan_lens_kernel<<<1, SCAN_THREADS, 0, st>>>(lens, nb, base, (uint64_t *)ws.off.p,
                                                 (uint64_t *)((uint8_t *)ws.misc.p + 8)));
    KL("lz4_block_write_kernel", st, lz4_block_write_kernel<<<nb, write_threads(), 0, st>>>(d_in, (const uint8_t *)ws.scratch.p,
                                                             (const RegionMeta *)ws.meta.p, (const BlockPlan *)ws.plan.p,
                                                             (const uint64_t *)ws.off.p, d_out_base, raw_limit >= 0 ? 1 : 0,
                                                             rpb, region_bytes, slot_bytes, block_bytes));
    return FOURMC_OK;
}

// ---- 4mz encode ------------------------------------------------------------------------------

__global__ void init_carry_kernel(uint8_t *misc, uint64_t base)
{
    *(uint32_t *)misc = 0; *(uint64_t *)(misc + 8) = 0; *(uint64_t *)(misc + 16) = 0; *(uint64_t *)(misc + 24) = base;
}

int zgroup_blocks()
{
    const char *e = getenv("FOURMC_ZGROUP");       // read per call: the tests shrink it to exercise the carry
    int v = e ? atoi(e) : 512;
    if (v < 1) v = 1;
    if (v > 4096) v = 4096;
    return v;
}

// Same contract as enc_span, for zstd frames.  Blocks are processed in groups so that the
// per-region scratch (sequence arrays in, block bodies out: 160 KiB per 64 KiB region) stays
// bounded; each group appends to the stream through a device-side carry (no host round trip).
int enc_span_zstd(fourmc_ctx *ctx, cudaStream_t st, EncWs &ws, int level, const uint8_t *d_in, size_t n,
                  uint8_t *d_out_base, uint64_t base, uint32_t *d_block_lens_out, int64_t raw_limit,
                  uint32_t block_bytes = FOURMC_BLOCKSIZE)
{
    const uint32_t nb = (uint32_t)((n + block_bytes - 1) / block_bytes);
    int r;
    if ((r = ensure(ctx, ws.misc, 64))) return r;
    if (nb == 0) { CK(cudaMemsetAsync(ws.misc.p, 0, 64, st)); return FOURMC_OK; }
    if ((r = ensure_ztables(ctx))) return r;
    const int depth = level_chain_depth(level);
    const uint32_t region_bytes = depth > 0 ? ENC_CHAIN_REGION : ENC_REGION;      // = bytes per zstd block
    const uint32_t rpb = FOURMC_BLOCKSIZE / region_bytes;
    const uint32_t G = std::min<uint32_t>(nb, (uint32_t)zgroup_blocks());
    const uint32_t greg = G * rpb;
    if ((r = ensure(ctx, ws.scratch, (size_t)greg * fmz::ze_in_slot(region_bytes)))) return r;
    if ((r = ensure(ctx, ws.zout, (size_t)greg * fmz::ze_out_slot(region_bytes)))) return r;
    if ((r = ensure(ctx, ws.zrout, (size_t)greg * sizeof(fmz::ZRegionOut)))) return r;
    if ((r = ensure(ctx, ws.meta, (size_t)greg * sizeof(RegionMeta)))) return r;
    if ((r = ensure(ctx, ws.plan, (size_t)nb * sizeof(BlockPlan)))) return r;
    if ((r = ensure(ctx, ws.lens, (size_t)nb * 4))) return r;
    if ((r = ensure(ctx, ws.off, (size_t)nb * 8))) return r;
    if (depth > 0 && (r = ensure(ctx, ws.chain, (size_t)G * block_bytes * 2))) return r;
    if (!ctx->region_attr_set) {
        CK(cudaFuncSetAttribute(lz4_region_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_SMEM));
        CK(cudaFuncSetAttribute(lz4_region_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_SMEM));
        CK(cudaFuncSetAttribute(lz4_region_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_SMEM_CHAIN));
        CK(cudaFuncSetAttribute(lz4_region_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_SMEM_CHAIN));
        ctx->region_attr_set = true;
    }
    uint8_t *misc = (uint8_t *)ws.misc.p;
    KL("init_carry_kernel", st, init_carry_kernel<<<1, 1, 0, st>>>(misc, base));
    uint32_t *lens = d_block_lens_out ? d_block_lens_out : (uint32_t *)ws.lens.p;
    for (uint32_t g0 = 0; g0 < nb; g0 += G) {
        const uint32_t gb = std::min<uint32_t>(G, nb - g0);
        const size_t goff = (size_t)g0 * block_bytes;
        const size_t gn = std::min<size_t>(n - goff, (size_t)gb * block_bytes);
        const uint32_t nreg = gb * rpb;
        if (g0) CK(cudaMemsetAsync(misc, 0, 4, st));
        EncParams P = {};
        P.in = d_in + goff; P.n = gn; P.n_regions = nreg;
        P.scratch = (uint8_t *)ws.scratch.p; P.meta = (RegionMeta *)ws.meta.p;
        P.work_counter = (uint32_t *)misc;
        P.min_match = level_min_match(level);
        P.slot_bytes = fmz::ze_in_slot(region_bytes);
        P.depth = depth;
        P.region_bytes = region_bytes; P.regions_per_block = rpb;
        P.block_bytes = block_bytes;
        P.reproducible = ctx->repro_call;
        if (P.depth > 0) {
            if ((r = launch_chain_links(ctx, st, ws, d_in + goff, gn, gb, block_bytes))) return r;
            P.chain = (const uint16_t *)ws.chain.p;
            const uint32_t grid = std::min<uint32_t>(nreg, (uint32_t)ctx->sm_count);
            KL("lz4_region_chain_kernel", st, lz4_region_kernel<true, true><<<grid, ENC_CHAIN_THREADS, ENC_SMEM_CHAIN, st>>>(P));
        } else {
            const uint32_t grid = std::min<uint32_t>(nreg, 2u * (uint32_t)ctx->sm_count);
            KL("lz4_region_kernel", st, lz4_region_kernel<true, false><<<grid, ENC_THREADS, ENC_SMEM, st>>>(P));
        }
        ZEncParams Z = {};
        Z.meta = (const RegionMeta *)ws.meta.p; Z.scratch_in = (const uint8_t *)ws.scratch.p;
        Z.scratch_out = (uint8_t *)ws.zout.p; Z.rout = (fmz::ZRegionOut *)ws.zrout.p;
        Z.tables = (const fmz::Tables *)ctx->ztables.p; Z.n = gn; Z.n_regions = nreg;
        Z.region_bytes = region_bytes; Z.regions_per_block = rpb; Z.block_bytes = block_bytes;
        KL("zstd_entropy_kernel", st, zstd_entropy_kernel<<<nreg, fmz::ZE_THREADS, 0, st>>>(Z));
        KL("zstd_block_size_kernel", st, zstd_block_size_kernel<<<(gb + 127) / 128, 128, 0, st>>>(
            (const fmz::ZRegionOut *)ws.zrout.p, gb, gn, (BlockPlan *)ws.plan.p + g0, lens + g0, raw_limit, region_bytes, rpb, block_bytes));
        KL("scan_lens_carry_kernel", st, scan_lens_carry_kernel<<<1, SCAN_THREADS, 0, st>>>(
            lens + g0, gb, (uint64_t *)(misc + 24), (uint64_t *)ws.off.p + g0, (uint64_t *)(misc + 8)));
        KL("zstd_block_write_kernel", st, zstd_block_write_kernel<<<gb, write_threads(), 0, st>>>(
            d_in + goff, (const uint8_t *)ws.zout.p, (const fmz::ZRegionOut *)ws.zrout.p, (const BlockPlan *)ws.plan.p + g0,
            (const uint64_t *)ws.off.p + g0, d_out_base, raw_limit >= 0 ? 1 : 0, region_bytes, rpb, block_bytes));
    }
    return FOURMC_OK;
}

int enc_span_codec(fourmc_ctx *ctx, cudaStream_t st, EncWs &ws, int codec, int level, const uint8_t *d_in, size_t n,
                   uint8_t *d_out_base, uint64_t base, uint32_t *d_block_lens_out, int64_t raw_limit)
{
    return codec == CODEC_ZSTD ? enc_span_zstd(ctx, st, ws, level, d_in, n, d_out_base, base, d_block_lens_out, raw_limit)
                               : enc_span(ctx, st, ws, level, d_in, n, d_out_base, base, d_block_lens_out, raw_limit);
}

// ---- decode --------------------------------------------------------------------------------

int ensure_side(fourmc_ctx *ctx, DecWs &ws)
{
    if (ws.side) return FOURMC_OK;
    CK(cudaStreamCreateWithFlags(&ws.side, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ws.fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ws.join, cudaEventDisableTiming));
    return FOURMC_OK;
}

// Runs verify + D1 + D0 + D2 + finalize over n_blocks descriptors already in ws.desc / ws.xxh /
// ws.status.  max_chunks bounds the chunk indices used by the descriptors.

// compact: the descriptors' destinations are consecutive (a stream): blocks that decode short are moved down.
// first: the nb blocks are entries [first, first + nb) of the workspace's descriptor / checksum / status tables
// (a block range of a longer stream).
int dec_blocks(fourmc_ctx *ctx, cudaStream_t st, DecWs &ws, uint32_t nb, size_t max_chunks, int check_xxh,
               int32_t *d_out_size, const IndexInfo *d_info, long long *d_result, int codec = CODEC_LZ4, int compact = 0,
               uint32_t first = 0, bool pipelined = false)
{
    int r;
    if (codec == CODEC_ZSTD) max_chunks = 0;
    if ((r = ensure(ctx, ws.tokmap, (max_chunks + 1) * LZ4_CHUNK_WORDS * 4))) return r;
    if ((r = ensure(ctx, ws.chunkop, (max_chunks + 1) * 4))) return r;
    if ((r = ensure(ctx, ws.result, (size_t)std::max<uint32_t>(nb, 1) * 4))) return r;
    const BlockDesc *const desc_all = (const BlockDesc *)ws.desc.p + first;
    const uint32_t *const xxh_all = (const uint32_t *)ws.xxh.p + first;
    uint8_t *const status_all = (uint8_t *)ws.status.p + first;
    bool verify_forked = false;
    if (nb) {
        CK(cudaMemsetAsync(ws.tokmap.p, 0, (max_chunks + 1) * LZ4_CHUNK_WORDS * 4, st));
        const BlockDesc *desc = desc_all;
        uint8_t *status = status_all;
        // The checksum pass runs on a side stream beside the decode.  Its launch comes AFTER the parse kernel's
        // (verify_order 1) or after the copy kernel's (2): submitted first (0, round 1), its CTAs -- 16 KiB of shared
        // memory each, seven per SM -- took the room of half the parse kernel's CTAs for as long as they ran.
        static int verify_order = -1;
        if (verify_order < 0) { const char *e = getenv("FOURMC_VERIFY_ORDER"); verify_order = e ? atoi(e) : 1; }
        if (check_xxh) {
            int rs;
            if ((rs = ensure_side(ctx, ws))) return rs;
            CK(cudaEventRecord(ws.fork, st));
            CK(cudaStreamWaitEvent(ws.side, ws.fork, 0));
        }
        auto launch_verify = [&]() -> int {
            if (!check_xxh || verify_forked) return FOURMC_OK;
            KL("xxh_verify_kernel", ws.side, xxh_verify_kernel<<<(nb + VERIFY_WARPS - 1) / VERIFY_WARPS, VERIFY_WARPS * 32, 0, ws.side>>>(
                desc, xxh_all, nb, status));
            CK(cudaEventRecord(ws.join, ws.side));
            verify_forked = true;
            return FOURMC_OK;
        };
        if (verify_order == 0 || codec == CODEC_ZSTD) { if ((r = launch_verify())) return r; }
        if (!ctx->d1_attr_set) {
            CK(cudaFuncSetAttribute(lz4_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D1_SMEM));
            ctx->d1_attr_set = true;
        }
        if (codec == CODEC_ZSTD) {
            if ((r = ensure_ztables(ctx))) return r;
            if ((r = ensure(ctx, ws.zwork, (size_t)nb * sizeof(fmz::Work)))) return r;
            if (!ctx->zd_attr_set) {
                CK(cudaFuncSetAttribute(zstd_frames_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZD_SMEM));
                ctx->zd_attr_set = true;
            }
            // FOURMC_ZD_MODE = lane (default) | warp | serial: which fast path runs before the exact serial decoder
            static int zd_mode = -1;
            if (zd_mode < 0) {
                const char *e = getenv("FOURMC_ZD_MODE");
                zd_mode = !e ? 2 : !strcmp(e, "serial") ? 0 : !strcmp(e, "warp") ? 1 : 2;
            }
            const int zd_serial = zd_mode == 0;
            if (zd_mode == 1)
                KL("zstd_frames_warp_kernel", st, zstd_frames_warp_kernel<<<(nb + ZD_WARPS - 1) / ZD_WARPS, ZD_WARPS * 32, ZD_SMEM, st>>>(
                    desc, nb, (fmz::Work *)ws.zwork.p, (const fmz::Tables *)ctx->ztables.p, (int32_t *)ws.result.p));
            if (zd_mode == 2) {
                // resident warps per SM: the fewest that give the smallest number of rounds over the
                // frames (a frame's latency grows with the warps sharing an SM); FOURMC_ZL_PER_SM overrides
                static int forced = -1;
                if (forced < 0) { const char *e = getenv("FOURMC_ZL_PER_SM"); forced = e ? atoi(e) : 0; if (forced > 32) forced = 32; }
                uint32_t per_sm = (uint32_t)forced;
                if (!per_sm) {
                    const uint32_t sm = (uint32_t)ctx->sm_count, most = 28;
                    const uint32_t rounds = (nb + sm * most - 1) / (sm * most);
                    per_sm = (nb + sm * rounds - 1) / (sm * rounds);
                    if (per_sm < 1) per_sm = 1;
                    if (per_sm > most) per_sm = most;
                }
                const uint32_t grid = std::min<uint32_t>(nb, (uint32_t)ctx->sm_count * per_sm);
                if ((r = ensure(ctx, ws.zlane, (size_t)grid * sizeof(ZlScratch) + 64))) return r;
                uint32_t *counter = (uint32_t *)((uint8_t *)ws.zlane.p + (size_t)grid * sizeof(ZlScratch));
                CK(cudaMemsetAsync(counter, 0, 4, st));
                KL("zstd_frames_lane_kernel", st, zstd_frames_lane_kernel<<<grid, 32, 0, st>>>(
                    desc, nb, (ZlScratch *)ws.zlane.p, (const fmz::Tables *)ctx->ztables.p, (int32_t *)ws.result.p, counter));
            }
            KL("zstd_frames_kernel", st, zstd_frames_kernel<<<(nb + 31) / 32, 32, 0, st>>>(desc, nb, (fmz::Work *)ws.zwork.p,
                                                           (const fmz::Tables *)ctx->ztables.p, (int32_t *)ws.result.p, zd_serial ? 0 : 1));
        } else {
            // few blocks: a CTA of 16 warps per block (its chain walk is the only serial part, ~1 ms per block and
            // one block per SM at a time); many blocks: a warp per block.  FOURMC_D1_WIDE=0|1 forces one of them.
            static int wide_forced = -1;
            if (wide_forced < 0) { const char *e = getenv("FOURMC_D1_WIDE"); wide_forced = e ? (atoi(e) ? 1 : 0) : 2; }
            // r02i: 2.7 ms per wave of 148 blocks, the warp-per-block kernel 9.4 ms up to ~600 blocks.  Not for the slices of
            // the host pipeline: several of them are in flight, and a 16-warp CTA with 200 KiB of shared memory keeps the
            // other slices' kernels off its SM (r02j: host-buffer decode 37 GB/s with it, 42-45 without)
            const uint32_t sm = (uint32_t)ctx->sm_count;
            const bool wide = wide_forced == 2 ? (!pipelined && nb <= sm * 4) : wide_forced == 1;
            if (wide) {
                if (!ctx->d1w_attr_set) {
                    CK(cudaFuncSetAttribute(lz4_parse_wide_kernel<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, d1w_smem_bytes<15>()));
                    CK(cudaFuncSetAttribute(lz4_parse_wide_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, d1w_smem_bytes<7>()));
                    ctx->d1w_attr_set = true;
                }
                if (nb <= sm)
                    KL("lz4_parse_wide_kernel", st, lz4_parse_wide_kernel<15><<<nb, 16 * 32, d1w_smem_bytes<15>(), st>>>(desc, nb, (uint32_t *)ws.tokmap.p,
                                                                               (uint32_t *)ws.chunkop.p, (int32_t *)ws.result.p));
                else
                    KL("lz4_parse_wide_kernel", st, lz4_parse_wide_kernel<7><<<nb, 8 * 32, d1w_smem_bytes<7>(), st>>>(desc, nb, (uint32_t *)ws.tokmap.p,
                                                                              (uint32_t *)ws.chunkop.p, (int32_t *)ws.result.p));
            } else
                KL("lz4_parse_kernel", st, lz4_parse_kernel<<<(nb + D1_WARPS - 1) / D1_WARPS, D1_WARPS * 32, D1_SMEM, st>>>(desc, nb, (uint32_t *)ws.tokmap.p,
                                                               (uint32_t *)ws.chunkop.p, (int32_t *)ws.result.p));
        }
        if (verify_order == 1) { if ((r = launch_verify())) return r; }
        for (uint32_t b0 = 0; b0 < nb; b0 += 32768) {
            const uint32_t cnt = std::min<uint32_t>(32768, nb - b0);
            KL("lz4_stored_kernel", st, lz4_stored_kernel<<<dim3(32, cnt), 256, 0, st>>>(desc + b0, cnt));
        }
        if (codec == CODEC_LZ4) {
            // warps per block.  Many blocks: one warp each, the word-stage kernel with 1 KiB items (full batches; a
            // lone warp needs ~25 ms for its block however few blocks there are).  Fewer: four warps share a block,
            // same kernel with 128-byte items.  Few: eight warps, the byte-granular kernel, which delivers a single
            // item soonest.  D2 ms at 2048 / 1024 / 512 / 256 / 64 / 1 blocks (r02v, r02w): one warp 29.8 / 26.1 / 25.2 /
            // 24.8 / - / 21; four 44.5 / 28.2 / 14.5 / 13.2 / 14.4 / -; eight, byte-granular - / - / 17.9 / 9.0 / 8.1 / 7.7
            static int forced = -1, forced_bytes = -1;
            if (forced < 0) { const char *e = getenv("FOURMC_D2_WARPS"); forced = e ? atoi(e) : 0; }
            if (forced_bytes < 0) { const char *e = getenv("FOURMC_D2_BYTES"); forced_bytes = e ? (atoi(e) ? 1 : 0) : 2; }
            const uint32_t sm = (uint32_t)ctx->sm_count;
            int w = forced;
            if (w != 1 && w != 2 && w != 4 && w != 8) w = nb >= sm * 6 ? 1 : nb * 10 >= sm * 26 ? 4 : 8;
            const bool bytes = w > 1 && (forced_bytes == 2 ? w == 8 : forced_bytes == 1);
            const uint32_t *tm = (const uint32_t *)ws.tokmap.p, *co = (const uint32_t *)ws.chunkop.p;
            const int32_t *rs = (const int32_t *)ws.result.p;
#define D2_LAUNCH(WW, BB) KL("lz4_copy_kernel", st, (lz4_copy_kernel<WW, BB><<<nb, WW * 32, 0, st>>>(desc, tm, co, rs)))
            if (w == 1) D2_LAUNCH(1, false);
            else if (w == 2) { if (bytes) D2_LAUNCH(2, true); else D2_LAUNCH(2, false); }
            else if (w == 4) { if (bytes) D2_LAUNCH(4, true); else D2_LAUNCH(4, false); }
            else { if (bytes) D2_LAUNCH(8, true); else D2_LAUNCH(8, false); }
#undef D2_LAUNCH
        }
        if ((r = launch_verify())) return r;          // verify_order 2, and whatever has not launched it yet
    }
    if (verify_forked) CK(cudaStreamWaitEvent(st, ws.join, 0));
    if ((r = ensure(ctx, ws.final_, 16))) return r;
    KL("finalize_kernel", st, finalize_kernel<<<1, SCAN_THREADS, 0, st>>>(desc_all, (const int32_t *)ws.result.p, nb,
                                                status_all, d_out_size, d_info, d_result, (uint32_t *)ws.final_.p));
    if (compact && nb)
        KL("compact_kernel", st, compact_kernel<<<1, SCAN_THREADS, 0, st>>>(desc_all, (const int32_t *)ws.result.p,
                                                  status_all, nb, (const uint32_t *)ws.final_.p));
    return FOURMC_OK;
}

// Builds BlockDesc[] from user tables (all device arrays).  One CTA.
__global__ void __launch_bounds__(SCAN_THREADS)
build_desc_kernel(uint32_t nb, const uint8_t *src, const uint64_t *src_off, const uint32_t *csize,
                  const uint32_t *usize, uint8_t *dst, const uint64_t *dst_off, BlockDesc *desc, uint8_t *status)
{
    __shared__ unsigned long long tmp[32];
    unsigned long long carry = 0;
    for (uint32_t i0 = 0; i0 < nb; i0 += SCAN_THREADS) {
        const uint32_t i = i0 + threadIdx.x;
        const bool live = i < nb;
        const uint32_t c = live ? csize[i] : 0, u = live ? usize[i] : 0;
        const bool hash_only = u == 0xffffffffu;          // a footer travelling with the blocks: checksum, no decode
        const bool toolarge = !hash_only && (c > FOURMC_BLOCKSIZE || (c != u && u > FOURMC_BLOCKSIZE));
        const uint32_t nch = (live && !toolarge && !hash_only && c != u) ? (c + 15 + LZ4_CHUNK - 1) / LZ4_CHUNK : 0;
        unsigned long long total;
        const unsigned long long incl = cta_incl_scan_u64(nch, tmp, &total);
        if (live) {
            BlockDesc d;
            d.src = src + src_off[i]; d.dst = dst + dst_off[i];
            d.csize = c; d.usize = u; d.chunk_base = (uint32_t)(carry + incl - nch); d.stored = (c == u) ? 1u : 0u;
            uint8_t s = FOURMC_BLOCK_OK;
            if (toolarge) { s = FOURMC_BLOCK_TOOLARGE; d.csize = 0; d.usize = 0; d.stored = 1; }
            if (hash_only) { d.usize = 0; d.stored = 2; }
            desc[i] = d; status[i] = s;
        }
        carry += total;
    }
}

__global__ void gen_kernel(int kind, uint64_t seed, uint64_t first_page, uint64_t n_pages, uint8_t *out)
{
    // one warp per CTA: each lane builds a page in shared memory (padded rows -> distinct banks),
    // then the warp streams the 32 pages out with 16-byte stores.
    extern __shared__ __align__(16) uint8_t pages[];
    constexpr int ROW = FMG_PAGE + 16;
    const uint64_t p0 = (uint64_t)blockIdx.x * 32;
    const uint64_t mine = p0 + threadIdx.x;
    if (mine < n_pages) fmg_page(kind, seed, first_page + mine, pages + (size_t)threadIdx.x * ROW);
    __syncwarp();
    for (int r = 0; r < 32 && p0 + r < n_pages; r++) {
        const uint4 *s = (const uint4 *)(pages + (size_t)r * ROW);
        uint4 *d = (uint4 *)(out + (p0 + r) * FMG_PAGE);
        for (int i = threadIdx.x; i < (int)(FMG_PAGE / 16); i += 32) d[i] = s[i];
    }
}

int pinned_scratch(fourmc_ctx *ctx, size_t need)
{
    if (ctx->pinned_cap >= need) return FOURMC_OK;
    if (ctx->pinned) { cudaFreeHost(ctx->pinned); ctx->pinned = nullptr; ctx->pinned_cap = 0; }
    need += need / 8;                                 // headroom, as in ensure()
    CK(cudaMallocHost(&ctx->pinned, need));
    ctx->pinned_cap = need;
    return FOURMC_OK;
}

inline uint32_t be32(const uint8_t *p)
{
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

}  // namespace

// =================================================================================================
// C-ABI
// =================================================================================================

extern "C" {

int fourmc_ctx_create(fourmc_ctx **out, int device)
{
    if (!out) return FOURMC_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return FOURMC_E_CUDA;
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) return FOURMC_E_CUDA; }
    if (device >= ndev) return FOURMC_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return FOURMC_E_CUDA;
    fourmc_ctx *ctx = new fourmc_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return FOURMC_E_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    if (const char *e = getenv("FOURMC_REPRODUCIBLE")) { ctx->reproducible = atoi(e) ? 1 : 0; ctx->repro_call = ctx->reproducible; }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return FOURMC_E_CUDA; }
    for (int i = 0; i < FM_PIPE_MAX; i++) {
        if (cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming) != cudaSuccess) {
            fourmc_ctx_destroy(ctx);
            return FOURMC_E_CUDA;
        }
    }
    *out = ctx;
    return FOURMC_OK;
}

void fourmc_ctx_destroy(fourmc_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < FM_PIPE_MAX; i++) {
        EncWs &e = ctx->enc[i];
        release(e.scratch); release(e.meta); release(e.plan); release(e.lens); release(e.off); release(e.misc); release(e.zout); release(e.zrout); release(e.chain);
        DecWs &d = ctx->dec[i];
        release(d.desc); release(d.xxh); release(d.status); release(d.tokmap); release(d.chunkop);
        release(d.result); release(d.info); release(d.tables); release(d.outsize); release(d.final_); release(d.zwork); release(d.zlane);
        if (d.side) cudaStreamDestroy(d.side);
        if (d.fork) cudaEventDestroy(d.fork);
        if (d.join) cudaEventDestroy(d.join);
        release(ctx->stage_in[i]); release(ctx->stage_out[i]);
        if (ctx->aux[i]) cudaStreamDestroy(ctx->aux[i]);
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    }
    release(ctx->ztables);
    for (auto e : ctx->tev) cudaEventDestroy(e);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (void *q : ctx->pin_up) if (q) cudaFreeHost(q);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *fourmc_last_error(const fourmc_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }

uint64_t fourmc_kernel_launches(const fourmc_ctx *ctx) { return ctx ? ctx->launches : 0; }

int fourmc_sync(fourmc_ctx *ctx, void *stream)
{
    if (!ctx) return FOURMC_E_ARG;
    CK(cudaStreamSynchronize(pick(ctx, stream)));
    return FOURMC_OK;
}

int fourmc_ctx_set_reproducible(fourmc_ctx *ctx, int mode)
{
    if (!ctx || mode < -1 || mode > 1) return FOURMC_E_ARG;
    ctx->reproducible = mode;
    ctx->repro_call = mode < 0 ? 1 : mode;
    return FOURMC_OK;
}

int fourmc_timing_enable(fourmc_ctx *ctx, int on)
{
    if (!ctx) return FOURMC_E_ARG;
    ctx->timing = on != 0;
    ctx->tused = 0;
    return FOURMC_OK;
}

long long fourmc_timing_collect(fourmc_ctx *ctx, char *buf, size_t cap)
{
    if (!ctx || !buf || cap == 0) return FOURMC_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    struct Acc { const char *name; int count; double ms; };
    std::vector<Acc> acc;
    for (size_t i = 0; i < ctx->tused; i++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->tev[2 * i], ctx->tev[2 * i + 1]) != cudaSuccess) continue;
        bool found = false;
        for (auto &a : acc) if (!strcmp(a.name, ctx->tname[i])) { a.count++; a.ms += ms; found = true; break; }
        if (!found) acc.push_back(Acc{ctx->tname[i], 1, ms});
    }
    ctx->tused = 0;
    size_t pos = 0;
    buf[0] = 0;
    for (auto &a : acc) {
        const int w = snprintf(buf + pos, cap - pos, "%s %d %.6f\n", a.name, a.count, a.ms);
        if (w < 0 || (size_t)w >= cap - pos) break;
        pos += (size_t)w;
    }
    return (long long)pos;
}

int fourmc_lz4_compress_bound(int n)
{
    return ((unsigned)n > 0x7E000000u) ? 0 : n + n / 255 + 16;      // native/lz4/lz4.h:211-212
}

size_t fourmc_4mc_bound(size_t n)
{
    const size_t nb = (n + FOURMC_BLOCKSIZE - 1) / FOURMC_BLOCKSIZE;
    return 12 + n + 12 * nb + 12 + 20 + 4 * nb;
}

// ---- device-resident ---------------------------------------------------------------------------

static int compress_span_impl(fourmc_ctx *ctx, void *stream, int codec, int level, const void *d_in, size_t n,
                              void *d_span, size_t span_capacity, uint64_t *d_span_size, uint32_t *d_block_lens)
{
    if (!ctx || (!d_in && n) || !d_span) return FOURMC_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint32_t nb = blocks_of(n);
    if (span_capacity < n + 12ull * nb) return fail(ctx, FOURMC_E_OUTPUT, "span capacity below the all-stored bound");
    cudaStream_t st = pick(ctx, stream);
    EncWs &ws = ctx->enc[0];
    ctx->repro_call = ctx->reproducible < 0 ? 0 : ctx->reproducible;
    int r = enc_span_codec(ctx, st, ws, codec, level, (const uint8_t *)d_in, n, (uint8_t *)d_span, 0, d_block_lens, -1);
    ctx->repro_call = ctx->reproducible < 0 ? 1 : ctx->reproducible;
    if (r) return r;
    if (d_span_size)
        CK(cudaMemcpyAsync(d_span_size, (uint8_t *)ws.misc.p + 8, 8, cudaMemcpyDeviceToDevice, st));
    return FOURMC_OK;
}

int fourmc_4mc_compress_span_device(fourmc_ctx *ctx, void *stream, int level, const void *d_in, size_t n,
                                    void *d_span, size_t span_capacity, uint64_t *d_span_size, uint32_t *d_block_lens)
{
    return compress_span_impl(ctx, stream, CODEC_LZ4, level, d_in, n, d_span, span_capacity, d_span_size, d_block_lens);
}

int fourmc_4mz_compress_span_device(fourmc_ctx *ctx, void *stream, int level, const void *d_in, size_t n,
                                    void *d_span, size_t span_capacity, uint64_t *d_span_size, uint32_t *d_block_lens)
{
    return compress_span_impl(ctx, stream, CODEC_ZSTD, level, d_in, n, d_span, span_capacity, d_span_size, d_block_lens);
}

static int build_index_impl(fourmc_ctx *ctx, void *stream, uint32_t magic, const uint32_t *d_block_lens, uint32_t n_blocks,
                            void *d_header, void *d_tail)
{
    if (!ctx || !d_tail || (n_blocks && !d_block_lens)) return FOURMC_E_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    KL("write_index_kernel", st, write_index_kernel<<<1, SCAN_THREADS, 0, st>>>(d_block_lens, n_blocks, 12, magic, (uint8_t *)d_header,
                                                   (uint8_t *)d_tail, nullptr, nullptr));
    return FOURMC_OK;
}

int fourmc_4mc_build_index_device(fourmc_ctx *ctx, void *stream, const uint32_t *d_block_lens, uint32_t n_blocks,
                                  void *d_header, void *d_tail)
{
    return build_index_impl(ctx, stream, FOURMC_MAGIC_4MC, d_block_lens, n_blocks, d_header, d_tail);
}

int fourmc_4mz_build_index_device(fourmc_ctx *ctx, void *stream, const uint32_t *d_block_lens, uint32_t n_blocks,
                                  void *d_header, void *d_tail)
{
    return build_index_impl(ctx, stream, FOURMC_MAGIC_4MZ, d_block_lens, n_blocks, d_header, d_tail);
}

static int compress_device_impl(fourmc_ctx *ctx, void *stream, int codec, int level, const void *d_in, size_t n, void *d_out,
                                size_t out_capacity, uint64_t *d_out_size, uint32_t *d_block_lens)
{
    if (!ctx || (!d_in && n) || !d_out) return FOURMC_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (out_capacity < fourmc_4mc_bound(n)) return fail(ctx, FOURMC_E_OUTPUT, "output capacity below fourmc_4mc_bound(n)");
    cudaStream_t st = pick(ctx, stream);
    EncWs &ws = ctx->enc[0];
    const uint32_t nb = blocks_of(n);
    int r;
    if ((r = ensure(ctx, ws.lens, (size_t)std::max<uint32_t>(nb, 1) * 4))) return r;
    uint32_t *lens = d_block_lens ? d_block_lens : (uint32_t *)ws.lens.p;
    // block b's header lands at d_out + 12 + sum of earlier record lengths
    ctx->repro_call = ctx->reproducible < 0 ? 0 : ctx->reproducible;
    r = enc_span_codec(ctx, st, ws, codec, level, (const uint8_t *)d_in, n, (uint8_t *)d_out, 12, lens, -1);
    ctx->repro_call = ctx->reproducible < 0 ? 1 : ctx->reproducible;
    if (r) return r;
    KL("write_index_kernel", st, write_index_kernel<<<1, SCAN_THREADS, 0, st>>>(lens, nb, 12, codec == CODEC_ZSTD ? FOURMC_MAGIC_4MZ : FOURMC_MAGIC_4MC,
                                                   (uint8_t *)d_out, (uint8_t *)d_out + 12,
                                                   (const uint64_t *)((uint8_t *)ws.misc.p + 8),
                                                   (uint64_t *)((uint8_t *)ws.misc.p + 16)));
    if (d_out_size)
        CK(cudaMemcpyAsync(d_out_size, (uint8_t *)ws.misc.p + 16, 8, cudaMemcpyDeviceToDevice, st));
    return FOURMC_OK;
}

int fourmc_4mc_compress_device(fourmc_ctx *ctx, void *stream, int level, const void *d_in, size_t n, void *d_out,
                               size_t out_capacity, uint64_t *d_out_size, uint32_t *d_block_lens)
{
    return compress_device_impl(ctx, stream, CODEC_LZ4, level, d_in, n, d_out, out_capacity, d_out_size, d_block_lens);
}

int fourmc_4mz_compress_device(fourmc_ctx *ctx, void *stream, int level, const void *d_in, size_t n, void *d_out,
                               size_t out_capacity, uint64_t *d_out_size, uint32_t *d_block_lens)
{
    return compress_device_impl(ctx, stream, CODEC_ZSTD, level, d_in, n, d_out, out_capacity, d_out_size, d_block_lens);
}

// first / count: the block range to decode (count = 0xffffffff: to the end); the output of block `first` lands at d_out.
static int decompress_device_impl(fourmc_ctx *ctx, void *stream, int codec, const void *d_in, size_t n, void *d_out,
                                  size_t out_capacity, long long *d_result, uint32_t first = 0, uint32_t count = 0xffffffffu)
{
    if (!ctx || !d_in || !d_result) return FOURMC_E_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    DecWs &ws = ctx->dec[0];
    int r;
    if ((r = pinned_scratch(ctx, 4096))) return r;
    long long *h = (long long *)ctx->pinned;
    if (n < 44) {
        h[0] = FOURMC_E_INPUT; h[1] = -1;
        CK(cudaMemcpyAsync(d_result, h, 16, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        return FOURMC_OK;
    }
    // the only host round trip: the footer size field tells how many blocks to launch for
    uint8_t *ft = (uint8_t *)ctx->pinned + 64;
    CK(cudaMemcpyAsync(ft, (const uint8_t *)d_in + n - 12, 12, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t fsize = be32(ft);
    if (fsize < 20 || (uint64_t)fsize > n - 24 || ((fsize - 20) & 3)) {
        h[0] = FOURMC_E_CONTENT; h[1] = -1;
        CK(cudaMemcpyAsync(d_result, h, 16, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        return FOURMC_OK;
    }
    const uint32_t nb = (fsize - 20) / 4;
    if ((r = ensure(ctx, ws.desc, (size_t)std::max<uint32_t>(nb, 1) * sizeof(BlockDesc)))) return r;
    if ((r = ensure(ctx, ws.xxh, (size_t)std::max<uint32_t>(nb, 1) * 4))) return r;
    if ((r = ensure(ctx, ws.status, (size_t)std::max<uint32_t>(nb, 1)))) return r;
    if ((r = ensure(ctx, ws.info, sizeof(IndexInfo)))) return r;
    // a rejected index leaves the descriptors untouched: zero them first so that the kernels
    // queued behind (no host round trip to skip them) see empty blocks
    CK(cudaMemsetAsync(ws.desc.p, 0, (size_t)std::max<uint32_t>(nb, 1) * sizeof(BlockDesc), st));
    CK(cudaMemsetAsync(ws.status.p, 0, (size_t)std::max<uint32_t>(nb, 1), st));
    CK(cudaMemsetAsync(ws.xxh.p, 0, (size_t)std::max<uint32_t>(nb, 1) * 4, st));
    KL("read_index_kernel", st, read_index_kernel<<<1, SCAN_THREADS, 0, st>>>((const uint8_t *)d_in, n, nb, (uint8_t *)d_out, out_capacity,
                                                  (BlockDesc *)ws.desc.p, (uint32_t *)ws.xxh.p, (uint8_t *)ws.status.p,
                                                  (IndexInfo *)ws.info.p, codec == CODEC_ZSTD ? FOURMC_MAGIC_4MZ : FOURMC_MAGIC_4MC,
                                                  first, count));
    first = std::min(first, nb);
    const uint32_t cnt = std::min(count, nb - first);
    // chunk indices are rebased to the range's first block: the range's payload bounds the side tables
    const size_t max_chunks = (cnt == nb ? n : std::min<size_t>(n, (size_t)cnt * (FOURMC_BLOCKSIZE + 64))) / LZ4_CHUNK + 2 * (size_t)cnt + 2;
    return dec_blocks(ctx, st, ws, cnt, max_chunks, 1, nullptr, (const IndexInfo *)ws.info.p, d_result, codec, 1, first);
}

int fourmc_4mc_decompress_device(fourmc_ctx *ctx, void *stream, const void *d_in, size_t n, void *d_out,
                                 size_t out_capacity, long long *d_result)
{
    return decompress_device_impl(ctx, stream, CODEC_LZ4, d_in, n, d_out, out_capacity, d_result);
}

int fourmc_4mz_decompress_device(fourmc_ctx *ctx, void *stream, const void *d_in, size_t n, void *d_out,
                                 size_t out_capacity, long long *d_result)
{
    return decompress_device_impl(ctx, stream, CODEC_ZSTD, d_in, n, d_out, out_capacity, d_result);
}

int fourmc_4mc_decompress_range_device(fourmc_ctx *ctx, void *stream, const void *d_in, size_t n, uint32_t first_block,
                                       uint32_t n_blocks, void *d_out, size_t out_capacity, long long *d_result)
{
    return decompress_device_impl(ctx, stream, CODEC_LZ4, d_in, n, d_out, out_capacity, d_result, first_block, n_blocks);
}

int fourmc_4mz_decompress_range_device(fourmc_ctx *ctx, void *stream, const void *d_in, size_t n, uint32_t first_block,
                                       uint32_t n_blocks, void *d_out, size_t out_capacity, long long *d_result)
{
    return decompress_device_impl(ctx, stream, CODEC_ZSTD, d_in, n, d_out, out_capacity, d_result, first_block, n_blocks);
}

// batch decode over caller tables with an explicit workspace (the host pipeline keeps two in flight)
static int dec_batch(fourmc_ctx *ctx, cudaStream_t st, DecWs &ws, uint32_t nb, const void *d_src,
                     const uint64_t *d_src_off, const uint32_t *d_csize, const uint32_t *d_usize, const uint32_t *d_xxh,
                     int check_xxh, void *d_dst, const uint64_t *d_dst_off, int32_t *d_out_size, uint8_t *d_status,
                     int codec = CODEC_LZ4, int compact = 0, bool pipelined = false)
{
    int r;
    if ((r = ensure(ctx, ws.desc, (size_t)std::max<uint32_t>(nb, 1) * sizeof(BlockDesc)))) return r;
    if ((r = ensure(ctx, ws.xxh, (size_t)std::max<uint32_t>(nb, 1) * 4))) return r;
    if ((r = ensure(ctx, ws.status, (size_t)std::max<uint32_t>(nb, 1)))) return r;
    if (nb == 0) return FOURMC_OK;
    KL("build_desc_kernel", st, build_desc_kernel<<<1, SCAN_THREADS, 0, st>>>(nb, (const uint8_t *)d_src, d_src_off, d_csize, d_usize,
                                                  (uint8_t *)d_dst, d_dst_off, (BlockDesc *)ws.desc.p, (uint8_t *)ws.status.p));
    if (check_xxh) CK(cudaMemcpyAsync(ws.xxh.p, d_xxh, (size_t)nb * 4, cudaMemcpyDeviceToDevice, st));
    // every compressed block has csize <= 4 MiB: bound the chunk count by that
    const size_t max_chunks = (size_t)nb * (FOURMC_BLOCKSIZE / LZ4_CHUNK + 2);
    if ((r = dec_blocks(ctx, st, ws, nb, max_chunks, check_xxh, d_out_size, nullptr, nullptr, codec, compact, 0, pipelined))) return r;
    if (d_status) CK(cudaMemcpyAsync(d_status, ws.status.p, nb, cudaMemcpyDeviceToDevice, st));
    return FOURMC_OK;
}

int fourmc_lz4_decompress_batch_device(fourmc_ctx *ctx, void *stream, uint32_t n_blocks, const void *d_src,
                                       const uint64_t *d_src_off, const uint32_t *d_csize, const uint32_t *d_usize,
                                       const uint32_t *d_xxh, int check_xxh, void *d_dst, const uint64_t *d_dst_off,
                                       int32_t *d_out_size, uint8_t *d_status)
{
    if (!ctx || !d_src || !d_src_off || !d_csize || !d_usize || !d_dst_off) return FOURMC_E_ARG;
    if (check_xxh && !d_xxh) return FOURMC_E_ARG;
    CK(cudaSetDevice(ctx->device));
    return dec_batch(ctx, pick(ctx, stream), ctx->dec[0], n_blocks, d_src, d_src_off, d_csize, d_usize, d_xxh, check_xxh,
                     d_dst, d_dst_off, d_out_size, d_status);
}

int fourmc_xxh32_batch_device(fourmc_ctx *ctx, void *stream, uint32_t n_items, const void *d_base,
                              const uint64_t *d_off, const uint32_t *d_len, uint32_t seed, uint32_t *d_out)
{
    if (!ctx || !d_off || !d_len || !d_out) return FOURMC_E_ARG;
    if (n_items == 0) return FOURMC_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    KL("xxh_batch_kernel", st, xxh_batch_kernel<<<(n_items + VERIFY_WARPS - 1) / VERIFY_WARPS, VERIFY_WARPS * 32, 0, st>>>(
        (const uint8_t *)d_base, d_off, d_len, n_items, seed, d_out));
    return FOURMC_OK;
}

int fourmc_gen_device(fourmc_ctx *ctx, void *stream, int kind, uint64_t seed, uint64_t first_page,
                      uint64_t n_pages, void *d_out)
{
    if (!ctx || !d_out) return FOURMC_E_ARG;
    if (kind < 0 || kind > 2) return fail(ctx, FOURMC_E_ARG, "kind: 0 log-text, 1 JSON, 2 silesia-like mix");
    if (n_pages == 0) return FOURMC_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = pick(ctx, stream);
    const size_t smem = 32 * (FMG_PAGE + 16);
    if (!ctx->gen_attr_set) { CK(cudaFuncSetAttribute(gen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); ctx->gen_attr_set = true; }
    const uint64_t max_grid = 1u << 30;
    for (uint64_t p0 = 0; p0 < n_pages; p0 += max_grid * 32) {
        const uint64_t cnt = std::min<uint64_t>(n_pages - p0, max_grid * 32);
        KL("gen_kernel", st, gen_kernel<<<(unsigned)((cnt + 31) / 32), 32, smem, st>>>(kind, seed, first_page + p0, cnt,
                                                                  (uint8_t *)d_out + p0 * FMG_PAGE));
    }
    return FOURMC_OK;
}

int fourmc_gen_host(int kind, uint64_t seed, uint64_t first_page, uint64_t n_pages, void *out)
{
    if (kind < 0 || kind > 2 || !out) return FOURMC_E_ARG;
    for (uint64_t p = 0; p < n_pages; p++) fmg_page(kind, seed, first_page + p, (uint8_t *)out + p * FMG_PAGE);
    return FOURMC_OK;
}

// ---- per-block, host pointers ------------------------------------------------------------------

uint32_t fourmc_xxh32(fourmc_ctx *ctx, const void *data, size_t len, uint32_t seed, int *status)
{
    auto done = [&](int s, uint32_t v) { if (status) *status = s; return v; };
    if (!ctx || (!data && len) || len > 0xffffffffull) return done(FOURMC_E_ARG, 0);
    if (cudaSetDevice(ctx->device) != cudaSuccess) return done(fail(ctx, FOURMC_E_CUDA, "cudaSetDevice"), 0);
    cudaStream_t st = ctx->stream;
    if (ensure(ctx, ctx->stage_in[0], len + 64) || pinned_scratch(ctx, 4096)) return done(FOURMC_E_CUDA, 0);
    // table: off (u64) | len (u32) | out (u32) in one small device buffer
    if (ensure(ctx, ctx->dec[0].tables, 64)) return done(FOURMC_E_CUDA, 0);
    uint8_t *tb = (uint8_t *)ctx->dec[0].tables.p;
    uint64_t *h = (uint64_t *)ctx->pinned;
    h[0] = 0; ((uint32_t *)h)[2] = (uint32_t)len; ((uint32_t *)h)[3] = 0;
    cudaError_t e;
    if (len && (e = cudaMemcpyAsync(ctx->stage_in[0].p, data, len, cudaMemcpyHostToDevice, st)) != cudaSuccess)
        return done(fail(ctx, FOURMC_E_CUDA, "H2D", e), 0);
    if ((e = cudaMemcpyAsync(tb, h, 16, cudaMemcpyHostToDevice, st)) != cudaSuccess)
        return done(fail(ctx, FOURMC_E_CUDA, "H2D", e), 0);
    int r = fourmc_xxh32_batch_device(ctx, st, 1, ctx->stage_in[0].p, (const uint64_t *)tb, (const uint32_t *)(tb + 8),
                                      seed, (uint32_t *)(tb + 12));
    if (r) return done(r, 0);
    if ((e = cudaMemcpyAsync((uint8_t *)ctx->pinned + 32, tb + 12, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(st)) != cudaSuccess)
        return done(fail(ctx, FOURMC_E_CUDA, "D2H", e), 0);
    return done(FOURMC_OK, *(uint32_t *)((uint8_t *)ctx->pinned + 32));
}

int fourmc_lz4_compress(fourmc_ctx *ctx, int level, const void *src, int src_size, void *dst, int dst_capacity)
{
    if (!ctx || !src || !dst || src_size < 0 || dst_capacity < 0) return FOURMC_E_ARG;
    if (src_size > FOURMC_BLOCKSIZE) return fail(ctx, FOURMC_E_ARG, "per-block calls take at most 4 MiB (native/4mc.c:116)");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int r;
    const size_t bound = (size_t)fourmc_lz4_compress_bound(src_size) + 64;
    if ((r = ensure(ctx, ctx->stage_in[0], (size_t)src_size + 64))) return r;
    if ((r = ensure(ctx, ctx->stage_out[0], bound + 64))) return r;
    if ((r = pinned_scratch(ctx, 4096))) return r;
    if (src_size == 0) {
        // LZ4_compress_default(src, dst, 0, cap): one token, no literals (native/lz4/lz4.c:1266-1293)
        if (dst_capacity < 1) return 0;
        ((uint8_t *)dst)[0] = 0;
        return 1;
    }
    CK(cudaMemcpyAsync(ctx->stage_in[0].p, src, (size_t)src_size, cudaMemcpyHostToDevice, st));
    EncWs &ws = ctx->enc[0];
    // the block record goes to stage_out + 4 so that the payload (record + 12) is 16-byte aligned
    uint8_t *rec = (uint8_t *)ctx->stage_out[0].p + 4;
    if ((r = enc_span(ctx, st, ws, level, (const uint8_t *)ctx->stage_in[0].p, (size_t)src_size, rec, 0,
                      nullptr, (int64_t)std::min<size_t>((size_t)dst_capacity, bound))))
        return r;
    uint32_t *h = (uint32_t *)ctx->pinned;
    CK(cudaMemcpyAsync(h, ws.plan.p, sizeof(BlockPlan), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const BlockPlan *p = (const BlockPlan *)h;
    if (p->stored) return 0;                                      // does not fit in dst_capacity
    const uint32_t c = p->payload;
    CK(cudaMemcpyAsync(dst, rec + 12, c, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return (int)c;
}

size_t fourmc_zstd_compress_bound(size_t n) { return fmz::ze_compress_bound(n); }     // native/zstd/zstd.h:204

// ZSTD_compress(dst, cap, src, n, level) on one block: native/4mc.c:467, native/jniZstdCompressor.c:93.
// Returns the frame size, or ZSTD's dstSize_tooSmall (-70) when the frame does not fit.
long long fourmc_zstd_compress(fourmc_ctx *ctx, int level, const void *src, size_t src_size, void *dst, size_t dst_capacity)
{
    if (!ctx || (!src && src_size) || !dst) return FOURMC_E_ARG;
    if (src_size > FOURMC_BLOCKSIZE) return fail(ctx, FOURMC_E_ARG, "per-block calls take at most 4 MiB (native/4mc.c:116)");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int r;
    if (src_size == 0) {                                          // frame header + one empty raw block
        if (dst_capacity < (size_t)fmz::ZE_FRAME_HDR + 3) return fmz::ERR_DSTSIZE;
        fmz::ze_write_frame_header((uint8_t *)dst, 0);
        fmz::ze_write_block_header((uint8_t *)dst + fmz::ZE_FRAME_HDR, true, 0, 0);
        return fmz::ZE_FRAME_HDR + 3;
    }
    const size_t bound = src_size + 3 * ENC_MAX_REGIONS_PER_BLOCK + fmz::ZE_FRAME_HDR + 64;
    if ((r = ensure(ctx, ctx->stage_in[0], src_size + 64))) return r;
    if ((r = ensure(ctx, ctx->stage_out[0], bound + 64))) return r;
    if ((r = pinned_scratch(ctx, 4096))) return r;
    CK(cudaMemcpyAsync(ctx->stage_in[0].p, src, src_size, cudaMemcpyHostToDevice, st));
    EncWs &ws = ctx->enc[0];
    uint8_t *rec = (uint8_t *)ctx->stage_out[0].p + 4;            // payload (record + 12) 16-byte aligned
    if ((r = enc_span_zstd(ctx, st, ws, level, (const uint8_t *)ctx->stage_in[0].p, src_size, rec, 0, nullptr,
                           (int64_t)std::min<size_t>(dst_capacity, bound))))
        return r;
    uint32_t *h = (uint32_t *)ctx->pinned;
    CK(cudaMemcpyAsync(h, ws.plan.p, sizeof(BlockPlan), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const BlockPlan *p = (const BlockPlan *)h;
    if (p->stored) return fmz::ERR_DSTSIZE;
    const uint32_t c = p->payload;
    CK(cudaMemcpyAsync(dst, rec + 12, c, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return (long long)c;
}

int fourmc_lz4_decompress_safe(fourmc_ctx *ctx, const void *src, int compressed_size, void *dst, int dst_capacity)
{
    if (!ctx) return FOURMC_E_ARG;
    ctx->err.clear();                                  // -10 / -11 are also legal LZ4 error values
    if (!src || dst_capacity < 0) return -1;                                        // lz4.c:1951
    if (dst_capacity == 0) return (compressed_size == 1 && ((const uint8_t *)src)[0] == 0) ? 0 : -1;   // :1977-1981
    if (compressed_size <= 0) return -1;                                            // :1982
    if (!dst) return FOURMC_E_ARG;
    if (compressed_size > fourmc_lz4_compress_bound(FOURMC_BLOCKSIZE) || dst_capacity > (1 << 30))
        return fail(ctx, FOURMC_E_ARG, "per-block calls take at most one 4 MiB block");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DecWs &ws = ctx->dec[0];
    int r;
    if ((r = ensure(ctx, ctx->stage_in[0], (size_t)compressed_size + 64))) return r;
    if ((r = ensure(ctx, ctx->stage_out[0], (size_t)dst_capacity + 64))) return r;
    if ((r = ensure(ctx, ws.desc, sizeof(BlockDesc)))) return r;
    if ((r = ensure(ctx, ws.status, 16))) return r;
    if ((r = pinned_scratch(ctx, 4096))) return r;
    CK(cudaMemcpyAsync(ctx->stage_in[0].p, src, (size_t)compressed_size, cudaMemcpyHostToDevice, st));
    BlockDesc *hd = (BlockDesc *)ctx->pinned;
    hd->src = (const uint8_t *)ctx->stage_in[0].p; hd->dst = (uint8_t *)ctx->stage_out[0].p;
    hd->csize = (uint32_t)compressed_size; hd->usize = (uint32_t)dst_capacity; hd->chunk_base = 0; hd->stored = 0;
    CK(cudaMemcpyAsync(ws.desc.p, hd, sizeof(BlockDesc), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ws.status.p, 0, 16, st));
    const size_t max_chunks = (size_t)compressed_size / LZ4_CHUNK + 3;
    if ((r = ensure(ctx, ws.outsize, 16))) return r;
    if ((r = dec_blocks(ctx, st, ws, 1, max_chunks, 0, (int32_t *)ws.outsize.p, nullptr, nullptr))) return r;
    int32_t *hr = (int32_t *)((uint8_t *)ctx->pinned + 256);
    CK(cudaMemcpyAsync(hr, ws.outsize.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int32_t res = *hr;
    if (res > 0) {
        CK(cudaMemcpyAsync(dst, ctx->stage_out[0].p, (size_t)res, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return res;
}

// ---- whole stream, host buffers ----------------------------------------------------------------

static long long compress_host_impl(fourmc_ctx *ctx, int codec, int level, const void *in, size_t n, void *out, size_t out_capacity)
{
    if (!ctx || (!in && n) || !out) return FOURMC_E_ARG;
    if (out_capacity < fourmc_4mc_bound(n)) return fail(ctx, FOURMC_E_OUTPUT, "output capacity below fourmc_4mc_bound(n)");
    CK(cudaSetDevice(ctx->device));
    const uint32_t nb = blocks_of(n);
    const size_t sl_blocks = codec == CODEC_LZ4 && level <= 1 && !getenv("FOURMC_SLICE_BLOCKS") ? (size_t)cslice_blocks() : (size_t)slice_blocks();
    const size_t sl_bytes = sl_blocks * FOURMC_BLOCKSIZE;
    const size_t nslices = (n + sl_bytes - 1) / sl_bytes;
    int r;
    if ((r = pinned_scratch(ctx, 4096))) return r;
    // all block lengths of the file, on the device, for the footer
    DevBuf &all_lens = ctx->dec[1].tables;
    if ((r = ensure(ctx, all_lens, (size_t)std::max<uint32_t>(nb, 1) * 4 + 64 + 32 + 4 * (size_t)nb))) return r;
    const int np = pipe_depth();
    uint64_t *h_span = (uint64_t *)ctx->pinned;          // span size of the slice in flight, per pipeline slot
    size_t pos = 12;                                     // output position of the next span
    const size_t cap_in = std::min(n, sl_bytes) + 64, cap_out = std::min(n, sl_bytes) + 12 * sl_blocks + 64;
    for (int i = 0; i < np && (size_t)i < std::max<size_t>(nslices, 1); i++) {
        if ((r = ensure(ctx, ctx->stage_in[i], cap_in))) return r;
        if ((r = ensure(ctx, ctx->stage_out[i], cap_out))) return r;
    }
    // software pipeline: slice s is uploaded and compressed on its slot's stream while earlier
    // slices download; spans are appended in order (each needs the sizes of all before it)
    for (size_t s = 0; s < nslices + (size_t)np - 1; s++) {
        if (s < nslices) {
            const int b = (int)(s % np);
            cudaStream_t st = ctx->aux[b];
            const size_t off = s * sl_bytes, len = std::min(sl_bytes, n - off);
            CK(cudaMemcpyAsync(ctx->stage_in[b].p, (const uint8_t *)in + off, len, cudaMemcpyHostToDevice, st));
            if ((r = enc_span_codec(ctx, st, ctx->enc[b], codec, level, (const uint8_t *)ctx->stage_in[b].p, len,
                                    (uint8_t *)ctx->stage_out[b].p, 0, (uint32_t *)all_lens.p + s * sl_blocks, -1)))
                return r;
            CK(cudaMemcpyAsync(&h_span[b], (uint8_t *)ctx->enc[b].misc.p + 8, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(ctx->ev[b], st));
        }
        if (s + 1 >= (size_t)np && s + 1 - np < nslices) {
            const int b = (int)((s + 1 - np) % np);
            cudaStream_t st = ctx->aux[b];
            CK(cudaEventSynchronize(ctx->ev[b]));
            const size_t span = (size_t)h_span[b];
            CK(cudaMemcpyAsync((uint8_t *)out + pos, ctx->stage_out[b].p, span, cudaMemcpyDeviceToHost, st));
            pos += span;
            // the slot is reused by slice s+1: its upload is queued on the same stream, in order
        }
    }
    for (int i = 0; i < np; i++) CK(cudaStreamSynchronize(ctx->aux[i]));
    // header + EOS + footer, assembled on the device from the gathered lengths
    cudaStream_t st = ctx->stream;
    uint8_t *d_hdr = (uint8_t *)all_lens.p + (((size_t)std::max<uint32_t>(nb, 1) * 4 + 15) & ~(size_t)15);
    uint8_t *d_tail = d_hdr + 16;
    if ((r = build_index_impl(ctx, st, codec == CODEC_ZSTD ? FOURMC_MAGIC_4MZ : FOURMC_MAGIC_4MC, (const uint32_t *)all_lens.p, nb, d_hdr, d_tail))) return r;
    const size_t tail_bytes = 12 + 20 + 4 * (size_t)nb;
    CK(cudaMemcpyAsync(out, d_hdr, 12, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync((uint8_t *)out + pos, d_tail, tail_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return (long long)(pos + tail_bytes);
}

long long fourmc_4mc_compress_host(fourmc_ctx *ctx, int level, const void *in, size_t n, void *out, size_t out_capacity)
{
    return compress_host_impl(ctx, CODEC_LZ4, level, in, n, out, out_capacity);
}

long long fourmc_4mz_compress_host(fourmc_ctx *ctx, int level, const void *in, size_t n, void *out, size_t out_capacity)
{
    return compress_host_impl(ctx, CODEC_ZSTD, level, in, n, out, out_capacity);
}

namespace {
struct HostBlock { uint64_t src_off, dst_off; uint32_t csize, usize, xxh; };
}

// Walks one or more concatenated streams exactly like decodeFourMC (native/4mc.c:560-707) minus the
// payload work.  Returns FOURMC_OK or the container error the serial reader would hit first,
// together with the blocks seen before it (they are decoded first, so that an earlier block error
// takes precedence).  *walk_err_at = number of blocks preceding the container error.
static int walk_streams(const uint8_t *in, size_t n, std::vector<HostBlock> &blocks, uint64_t *total_out,
                        uint32_t magic = FOURMC_MAGIC_4MC, uint32_t hdr_ck = 0xA4B73443u)
{
    size_t pos = 0;
    uint64_t opos = 0;
    while (pos < n) {
        const uint64_t opos_before = opos;
        if (n - pos < 4) return FOURMC_E_CONTENT;                                  // :868
        if (be32(in + pos) != magic) return FOURMC_E_CONTENT;                      // :873
        if (n - pos < 12) return FOURMC_E_CONTENT;                                 // :577
        if (be32(in + pos + 4) != FOURMC_VERSION) return FOURMC_E_CONTENT;         // :583
        // header checksum (:584): XXH32 of magic + version 1 is a constant (0xA4B73443 4mc, 0x289A1C9A 4mz)
        if (be32(in + pos + 8) != hdr_ck) return FOURMC_E_CONTENT;
        pos += 12;
        for (;;) {
            if (n - pos < 12) { *total_out = opos; return FOURMC_E_INPUT; }       // :610
            const uint32_t u = be32(in + pos), c = be32(in + pos + 4), ck = be32(in + pos + 8);
            pos += 12;
            if (u == 0 && c == 0 && ck == 0) break;                                // :616
            if (c > FOURMC_BLOCKSIZE) { *total_out = opos; return FOURMC_E_CONTENT; }   // :618
            if (n - pos < c) { *total_out = opos; return FOURMC_E_INPUT; }         // :632
            if (u != c && u > FOURMC_BLOCKSIZE) { *total_out = opos; return FOURMC_E_CONTENT; }   // :651
            blocks.push_back(HostBlock{(uint64_t)pos, opos, c, u, ck});
            opos += u;
            pos += c;
        }
        *total_out = opos;
        // footer :670-688 -- its checksum is verified on the device with the payloads
        if (n - pos < 4) return FOURMC_E_GENERIC;                                  // :672
        const uint32_t fsize = be32(in + pos);
        if (fsize < 4 || n - pos < fsize) return FOURMC_E_INPUT;                   // :680
        if (fsize < 8) return FOURMC_E_CONTENT;
        // represented as a pseudo block: csize = fsize - 4, usize = 0xffffffff marks "hash only"
        blocks.push_back(HostBlock{(uint64_t)pos, opos, fsize - 4, 0xffffffffu, be32(in + pos + fsize - 4)});
        if (be32(in + pos + 4) != 1) return FOURMC_E_CONTENT;                      // :687 (after the checksum, checked below)
        pos += fsize;
        // :909-913 `do { ... } while (decodedSize)`: a stream that announces no bytes ends the loop, whatever follows it.
        // (A stream whose blocks all DECODE to nothing while announcing sizes would also end the reference's loop; the
        // walk only sees the announced sizes -- such streams are written by no 4mc writer.)
        if (opos == opos_before) break;
    }
    return FOURMC_OK;
}

static long long decoded_size_impl(const void *in, size_t n, int codec)
{
    if (!in && n) return FOURMC_E_ARG;
    std::vector<HostBlock> blocks;
    uint64_t total = 0;
    const int e = walk_streams((const uint8_t *)in, n, blocks, &total, codec == CODEC_ZSTD ? FOURMC_MAGIC_4MZ : FOURMC_MAGIC_4MC,
                               codec == CODEC_ZSTD ? 0x289A1C9Au : 0xA4B73443u);
    return e ? e : (long long)total;
}

long long fourmc_4mc_decoded_size_host(const void *in, size_t n) { return decoded_size_impl(in, n, CODEC_LZ4); }
long long fourmc_4mz_decoded_size_host(const void *in, size_t n) { return decoded_size_impl(in, n, CODEC_ZSTD); }

static long long decompress_host_impl(fourmc_ctx *ctx, int codec, const void *in, size_t n, void *out, size_t out_capacity)
{
    if (!ctx || (!in && n) || (!out && out_capacity)) return FOURMC_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint8_t *src = (const uint8_t *)in;
    std::vector<HostBlock> blocks;
    uint64_t total = 0;
    const int walk_err = walk_streams(src, n, blocks, &total, codec == CODEC_ZSTD ? FOURMC_MAGIC_4MZ : FOURMC_MAGIC_4MC,
                                      codec == CODEC_ZSTD ? 0x289A1C9Au : 0xA4B73443u);
    if (total > out_capacity) return fail(ctx, FOURMC_E_OUTPUT, "destination too small");
    // slices of consecutive blocks; footers travel as hash-only items.  A zstd frame takes tens of
    // milliseconds however few are in flight (its entropy streams are serial), so 4mz slices are larger.
    const size_t sl_blocks = codec == CODEC_ZSTD ? (size_t)zslice_blocks() : (size_t)slice_blocks();
    int r;
    struct Slice { size_t b0, b1; uint64_t s0, s1, d0, d1; };
    std::vector<Slice> slices;
    // The download can only start when the first slice has been uploaded and decoded, so the first slices are small
    // (16, 32, 64 ... blocks, FOURMC_SLICE_RAMP=0 switches the ramp off): the first bytes come back after ~11 ms
    // instead of ~28 ms; the LZ4 slices of the ramp use the 16-warp parse (nothing else is on the GPU yet).
    static int ramp = -1;
    if (ramp < 0) { const char *e = getenv("FOURMC_SLICE_RAMP"); ramp = e ? atoi(e) : 16; if (ramp < 0 || ramp > 4096) ramp = 0; }
    size_t n_ramp = 0;
    for (size_t b0 = 0, step = ramp ? (size_t)ramp : sl_blocks; b0 < blocks.size();) {
        const size_t cnt = std::min(step, sl_blocks);
        if (cnt < sl_blocks) n_ramp++;
        const size_t b1 = std::min(blocks.size(), b0 + cnt);
        Slice s{b0, b1, blocks[b0].src_off, blocks[b1 - 1].src_off + blocks[b1 - 1].csize, blocks[b0].dst_off, 0};
        s.d1 = blocks[b1 - 1].dst_off + (blocks[b1 - 1].usize == 0xffffffffu ? 0 : blocks[b1 - 1].usize);
        slices.push_back(s);
        b0 = b1; step *= 2;
    }
    size_t max_in = 0, max_out = 0, max_cnt = 0;
    for (auto &s : slices) {
        max_in = std::max<size_t>(max_in, s.s1 - s.s0); max_out = std::max<size_t>(max_out, s.d1 - s.d0);
        max_cnt = std::max<size_t>(max_cnt, s.b1 - s.b0);
    }
    // everything the device reads from / writes to the host besides the payload lives in pinned
    // memory, so that no copy blocks the host and the two slice pipelines really overlap
    const size_t tb_host = (max_cnt * 28 + 63) & ~(size_t)63;
    const size_t pin_need = 4096 + FM_PIPE_MAX * tb_host + blocks.size() * 9 + 64;
    if ((r = pinned_scratch(ctx, pin_need))) return r;
    uint8_t *pin = (uint8_t *)ctx->pinned + 4096;
    const int np = codec == CODEC_ZSTD ? std::min(pipe_depth(), 3) : pipe_depth();
    uint8_t *h_tables[FM_PIPE_MAX];
    for (int i = 0; i < FM_PIPE_MAX; i++) h_tables[i] = pin + (size_t)i * tb_host;
    int32_t *sizes = (int32_t *)(pin + FM_PIPE_MAX * tb_host);         // decoded size of every item
    uint8_t *status = (uint8_t *)(sizes + blocks.size());
    for (size_t k = 0; k < slices.size(); k++) {
        const int b = (int)(k % np);
        const Slice &s = slices[k];
        cudaStream_t st = ctx->aux[b];
        const uint32_t cnt = (uint32_t)(s.b1 - s.b0);
        if ((r = ensure(ctx, ctx->stage_in[b], max_in + 64))) return r;
        if ((r = ensure(ctx, ctx->stage_out[b], max_out + 64))) return r;
        DecWs &ws = ctx->dec[b];
        // tables: src_off u64[cnt] | dst_off u64[cnt] | csize u32[cnt] | usize u32[cnt] | xxh u32[cnt] | hash_out u32[cnt] | out_size i32[cnt] | status u8[cnt]
        const size_t tb_bytes = (size_t)cnt * (8 + 8 + 4 + 4 + 4 + 4 + 4 + 1) + 64;
        if ((r = ensure(ctx, ws.tables, tb_bytes))) return r;
        CK(cudaStreamSynchronize(st));                       // previous use of this buffer pair is complete
        uint64_t *t_src = (uint64_t *)h_tables[b];
        uint64_t *t_dst = t_src + cnt;
        uint32_t *t_c = (uint32_t *)(t_dst + cnt), *t_u = t_c + cnt, *t_x = t_u + cnt;
        for (uint32_t i = 0; i < cnt; i++) {
            const HostBlock &hb = blocks[s.b0 + i];
            t_src[i] = hb.src_off - s.s0; t_dst[i] = hb.dst_off - s.d0;
            t_c[i] = hb.csize; t_x[i] = hb.xxh;
            t_u[i] = hb.usize;               // footers carry 0xffffffff: hashed, never decoded
        }
        uint8_t *d_tb = (uint8_t *)ws.tables.p;
        CK(cudaMemcpyAsync(ctx->stage_in[b].p, src + s.s0, s.s1 - s.s0, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_tb, h_tables[b], (size_t)cnt * 28, cudaMemcpyHostToDevice, st));
        const uint64_t *d_src_off = (const uint64_t *)d_tb, *d_dst_off = d_src_off + cnt;
        const uint32_t *d_c = (const uint32_t *)(d_dst_off + cnt), *d_u = d_c + cnt, *d_x = d_u + cnt;
        uint32_t *d_hash = (uint32_t *)(d_x + cnt);
        int32_t *d_osz = (int32_t *)(d_hash + cnt);
        uint8_t *d_st = (uint8_t *)(d_osz + cnt);
        // every item's payload checksum (blocks and footers) is verified inside the decode batch (:637/:645)
        if ((r = dec_batch(ctx, st, ws, cnt, ctx->stage_in[b].p, d_src_off, d_c, d_u, d_x, 1,
                           ctx->stage_out[b].p, d_dst_off, d_osz, d_st, codec, 0, slices.size() > 1 && k >= n_ramp)))
            return r;
        if (s.d1 > s.d0)
            CK(cudaMemcpyAsync((uint8_t *)out + s.d0, ctx->stage_out[b].p, s.d1 - s.d0, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(status + s.b0, d_st, cnt, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(sizes + s.b0, d_osz, (size_t)cnt * 4, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < np; i++) CK(cudaStreamSynchronize(ctx->aux[i]));
    // verdict in stream order: the first item whose checksum (:637/:645) or decode (:662) failed
    bool short_block = false;
    for (size_t i = 0; i < blocks.size(); i++) {
        if (status[i] != FOURMC_BLOCK_OK) return FOURMC_E_CONTENT;
        if (blocks[i].usize != 0xffffffffu && (uint32_t)sizes[i] != blocks[i].usize) short_block = true;
    }
    if (short_block) {
        // a block that decodes to fewer bytes than announced: the reference writes what the decoder returned
        // (native/4mc.c:661-666, :810-815), so the following blocks move down
        uint64_t wpos = 0;
        for (size_t i = 0; i < blocks.size(); i++) {
            if (blocks[i].usize == 0xffffffffu) continue;
            if (wpos != blocks[i].dst_off) memmove((uint8_t *)out + wpos, (uint8_t *)out + blocks[i].dst_off, (size_t)sizes[i]);
            wpos += (uint64_t)sizes[i];
        }
        memset((uint8_t *)out + wpos, 0, (size_t)(total - wpos));      // nothing of the staging buffers is left behind
        total = wpos;
    }
    if (walk_err) return walk_err;
    return (long long)total;
}

long long fourmc_4mc_decompress_host(fourmc_ctx *ctx, const void *in, size_t n, void *out, size_t out_capacity)
{
    return decompress_host_impl(ctx, CODEC_LZ4, in, n, out, out_capacity);
}

long long fourmc_4mz_decompress_host(fourmc_ctx *ctx, const void *in, size_t n, void *out, size_t out_capacity)
{
    return decompress_host_impl(ctx, CODEC_ZSTD, in, n, out, out_capacity);
}

// ---- raw codec streams: Hadoop's block-stream framing around the per-block natives ----------
// Lz4Codec.java:95-104,128-138 / ZstdCodec.java:103-112,136-146 -> BlockCompressorStream / BlockDecompressorStream.
// The framing (blockstream.h) is host code; chunks are compressed through the per-block calls (one native call per
// chunk, as the Java stream drives them) and decompressed as ONE batch when the stream is the reference writer's.

namespace {

long long bs_lz4_compress(void *u, int level, const uint8_t *src, uint32_t n, uint8_t *dst, size_t cap)
{
    // LZ4_compress(in, out, n) with a bound-sized destination: native/jniCompressor.c:91
    const int c = (int)std::min<size_t>(cap, (size_t)fourmc_lz4_compress_bound((int)n));
    return fourmc_lz4_compress((fourmc_ctx *)u, level, src, (int)n, dst, c);
}
long long bs_lz4_decompress(void *u, const uint8_t *src, uint32_t c, uint8_t *dst, uint32_t cap)
{
    return fourmc_lz4_decompress_safe((fourmc_ctx *)u, src, (int)c, dst, (int)cap);       // native/jniDecompressor.c:88
}
long long bs_zstd_compress(void *u, int level, const uint8_t *src, uint32_t n, uint8_t *dst, size_t cap)
{
    const long long r = fourmc_zstd_compress((fourmc_ctx *)u, level, src, n, dst, cap);     // native/jniZstdCompressor.c:93
    return r == 0 ? FOURMC_E_GENERIC : r;
}
long long bs_zstd_decompress(void *u, const uint8_t *src, uint32_t c, uint8_t *dst, uint32_t cap)
{
    return fourmc_zstd_decompress((fourmc_ctx *)u, src, c, dst, cap);                       // native/jniZstdDecompressor.c:90
}

fbs::Codec bs_codec(fourmc_ctx *ctx, int codec)
{
    if (codec == CODEC_ZSTD)
        return fbs::Codec{ctx, (uint32_t)fourmc_zstd_compress_bound(fbs::BUFFER), bs_zstd_compress, bs_zstd_decompress};
    return fbs::Codec{ctx, (uint32_t)fourmc_lz4_compress_bound((int)fbs::BUFFER), bs_lz4_compress, bs_lz4_decompress};
}

// All chunks at once.  Returns the decoded size, or 1 when the prediction did not hold (the caller then runs the
// serial reader, whose verdict is the reference's), or a negative FOURMC_E_*.
long long bs_decode_batch(fourmc_ctx *ctx, int codec, const uint8_t *in, size_t n, const std::vector<fbs::Chunk> &chunks,
                          size_t total, uint8_t *out)
{
    const uint32_t nb = (uint32_t)chunks.size();
    if (nb == 0) return 0;
    cudaStream_t st = ctx->stream;
    DecWs &ws = ctx->dec[0];
    int r;
    if ((r = ensure(ctx, ctx->stage_in[0], n + 64))) return r;
    if ((r = ensure(ctx, ctx->stage_out[0], total + 64))) return r;
    if ((r = ensure(ctx, ws.desc, (size_t)nb * sizeof(BlockDesc)))) return r;
    if ((r = ensure(ctx, ws.status, nb))) return r;
    if ((r = ensure(ctx, ws.outsize, (size_t)nb * 4))) return r;
    std::vector<BlockDesc> hd(nb);
    size_t chunk_base = 0;
    for (uint32_t i = 0; i < nb; ++i) {
        hd[i].src = (const uint8_t *)ctx->stage_in[0].p + chunks[i].src_off;
        hd[i].dst = (uint8_t *)ctx->stage_out[0].p + chunks[i].dst_off;
        hd[i].csize = chunks[i].clen; hd[i].usize = chunks[i].usize;
        hd[i].chunk_base = (uint32_t)chunk_base; hd[i].stored = 0;            // a chunk is never raw in this format
        chunk_base += (chunks[i].clen + 15 + LZ4_CHUNK - 1) / LZ4_CHUNK;
    }
    if (chunk_base > 0xfffffff0ull) return 1;
    CK(cudaMemcpyAsync(ctx->stage_in[0].p, in, n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ws.desc.p, hd.data(), (size_t)nb * sizeof(BlockDesc), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));                                            // hd is pageable and about to go away
    CK(cudaMemsetAsync(ws.status.p, 0, nb, st));
    if ((r = dec_blocks(ctx, st, ws, nb, chunk_base, 0, (int32_t *)ws.outsize.p, nullptr, nullptr, codec))) return r;
    std::vector<int32_t> sizes(nb);
    CK(cudaMemcpyAsync(sizes.data(), ws.outsize.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < nb; ++i)
        if (sizes[i] != (int32_t)chunks[i].usize) return 1;
    CK(cudaMemcpyAsync(out, ctx->stage_out[0].p, total, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return (long long)total;
}

}  // namespace

size_t fourmc_blockstream_bound(int zstd, size_t n, size_t write_size)
{
    const fbs::Codec c = bs_codec(nullptr, zstd ? CODEC_ZSTD : CODEC_LZ4);
    return fbs::bound(c, n, write_size);
}

// Writer, all chunks at once: possible when every chunk but the last has the same size (a constant write size,
// or one large write), since the encode kernels then see the input as blocks of `chunk` bytes instead of 4 MiB.
// Returns the stream size, 1 when the plan is not uniform (the caller goes chunk by chunk), or a negative FOURMC_E_*.
static long long bs_encode_batch(fourmc_ctx *ctx, int codec, const fbs::Codec &c, int level, const uint8_t *in, size_t n,
                                 const std::vector<fbs::Block> &blocks, bool trailing_zero, uint8_t *out, size_t cap)
{
    const uint32_t MAX = fbs::max_input(c);
    struct Piece { uint32_t len; bool first; uint32_t raw; };
    std::vector<Piece> pieces;
    for (const fbs::Block &b : blocks)
        for (uint32_t done = 0; done < b.raw;) {
            const uint32_t len = std::min(MAX, b.raw - done);
            pieces.push_back(Piece{len, done == 0, b.raw});
            done += len;
        }
    if (pieces.size() < 2) return 1;
    const uint32_t chunk = pieces[0].len;
    for (size_t i = 0; i + 1 < pieces.size(); ++i) if (pieces[i].len != chunk) return 1;
    if (pieces.back().len > chunk || (chunk & 15)) return 1;       // 16-byte multiples keep the regions' bulk copies aligned
    const uint32_t nb = (uint32_t)pieces.size();
    const size_t pay_bound = codec == CODEC_ZSTD ? fourmc_zstd_compress_bound(chunk) : (size_t)fourmc_lz4_compress_bound((int)chunk);
    const size_t rec_bound = 12 + pay_bound;
    cudaStream_t st = ctx->stream;
    EncWs &ws = ctx->enc[0];
    int r;
    if ((r = ensure(ctx, ctx->stage_in[0], n + 64))) return r;
    if ((r = ensure(ctx, ctx->stage_out[0], (size_t)nb * rec_bound + 64))) return r;
    CK(cudaMemcpyAsync(ctx->stage_in[0].p, in, n, cudaMemcpyHostToDevice, st));
    // records (12-byte header + payload) back to back from stage_out + 4: payloads 16-byte aligned for the first one only,
    // the write kernel copes with any alignment
    uint8_t *d_rec = (uint8_t *)ctx->stage_out[0].p + 4;
    if ((r = (codec == CODEC_ZSTD ? enc_span_zstd : enc_span)(ctx, st, ws, level, (const uint8_t *)ctx->stage_in[0].p, n, d_rec, 0,
                                                              nullptr, (int64_t)pay_bound, chunk)))
        return r;
    std::vector<BlockPlan> plan(nb);
    std::vector<uint64_t> off(nb);
    CK(cudaMemcpyAsync(plan.data(), ws.plan.p, (size_t)nb * sizeof(BlockPlan), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(off.data(), ws.off.p, (size_t)nb * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    size_t op = 0;
    for (uint32_t i = 0; i < nb; ++i) {
        if (plan[i].stored || plan[i].usize != pieces[i].len) return fail(ctx, FOURMC_E_GENERIC, "block-stream chunk did not fit its bound");
        const uint32_t csz = plan[i].payload;
        if (cap - op < (pieces[i].first ? 8u : 4u) || cap - op - (pieces[i].first ? 8u : 4u) < csz) return FOURMC_E_OUTPUT;
        if (pieces[i].first) { fbs::put32(out + op, pieces[i].raw); op += 4; }
        fbs::put32(out + op, csz); op += 4;
        CK(cudaMemcpyAsync(out + op, d_rec + off[i] + 12, csz, cudaMemcpyDeviceToHost, st));
        op += csz;
    }
    if (trailing_zero) {
        if (cap - op < 4) return FOURMC_E_OUTPUT;
        fbs::put32(out + op, 0); op += 4;
    }
    CK(cudaStreamSynchronize(st));
    return (long long)op;
}

long long fourmc_blockstream_compress_host(fourmc_ctx *ctx, int zstd, int level, const void *in, size_t n, size_t write_size,
                                           void *out, size_t out_capacity)
{
    if (!ctx || (!in && n) || !out) return FOURMC_E_ARG;
    const fbs::Codec c = bs_codec(ctx, zstd ? CODEC_ZSTD : CODEC_LZ4);
    if (!getenv("FOURMC_BS_SERIAL")) {
        CK(cudaSetDevice(ctx->device));
        std::vector<fbs::Block> blocks;
        bool trailing_zero = false;
        fbs::plan_blocks(c, n, write_size, blocks, &trailing_zero);
        const long long r = bs_encode_batch(ctx, zstd ? CODEC_ZSTD : CODEC_LZ4, c, level, (const uint8_t *)in, n, blocks, trailing_zero,
                                            (uint8_t *)out, out_capacity);
        if (r != 1) return r;
    }
    return fbs::compress(c, level, (const uint8_t *)in, n, write_size, (uint8_t *)out, out_capacity);
}

long long fourmc_blockstream_decompress_host(fourmc_ctx *ctx, int zstd, const void *in, size_t n, void *out, size_t out_capacity)
{
    if (!ctx || (!in && n) || (!out && out_capacity)) return FOURMC_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const int codec = zstd ? CODEC_ZSTD : CODEC_LZ4;
    const fbs::Codec c = bs_codec(ctx, codec);
    std::vector<fbs::Chunk> chunks;
    size_t total = 0;
    const bool serial_only = getenv("FOURMC_BS_SERIAL") != nullptr;          // tests: force the chunk-by-chunk reader
    if (!serial_only && fbs::predict_chunks(c, (const uint8_t *)in, n, chunks, &total) && total <= out_capacity) {
        const long long r = bs_decode_batch(ctx, codec, (const uint8_t *)in, n, chunks, total, (uint8_t *)out);
        if (r != 1) return r;
    }
    return fbs::decompress(c, (const uint8_t *)in, n, (uint8_t *)out, out_capacity);
}

// ---- block index and splits: the step either side of the path (BASELINE.json configs[4]) ------

// FourMcBlockIndex.findNextPosition (FourMcBlockIndex.java:92-104): the first block offset >= pos.
int64_t fourmc_index_find_next_position(const int64_t *offsets, int n, int64_t pos)
{
    if (!offsets || n <= 0) return FOURMC_NOT_FOUND;
    const int64_t *it = std::lower_bound(offsets, offsets + n, pos);
    return it == offsets + n ? FOURMC_NOT_FOUND : *it;
}

// FourMcBlockIndex.findBelongingBlockIndex (:111-124): the block whose offset is the last one <= pos.
int64_t fourmc_index_find_belonging_block(const int64_t *offsets, int n, int64_t pos)
{
    if (!offsets || n <= 0) return FOURMC_NOT_FOUND;
    const int64_t *it = std::upper_bound(offsets, offsets + n, pos);
    return it == offsets ? FOURMC_NOT_FOUND : (int64_t)(it - offsets) - 1;
}

// FourMcBlockIndex.alignSliceStartToIndex (:142-153)
int64_t fourmc_index_align_slice_start(const int64_t *offsets, int n, int64_t start, int64_t end)
{
    if (start == 0) return 0;
    const int64_t s = fourmc_index_find_next_position(offsets, n, start);
    return (s == FOURMC_NOT_FOUND || s >= end) ? FOURMC_NOT_FOUND : s;
}

// FourMcBlockIndex.alignSliceEndToIndex (:163-173)
int64_t fourmc_index_align_slice_end(const int64_t *offsets, int n, int64_t end, int64_t file_size)
{
    const int64_t e = fourmc_index_find_next_position(offsets, n, end);
    return e == FOURMC_NOT_FOUND ? file_size : e;
}

// FourMcInputFormat.getSplits for one file (FourMcInputFormat.java:126-173): Hadoop's default byte-range
// splits (FileInputFormat: pieces of split_size, the last one up to 1.1 x split_size) nudged to block starts;
// splits that contain no block start disappear.  Returns the number of splits (may exceed cap: call again).
int fourmc_plan_splits(const int64_t *offsets, int n, int64_t file_size, int64_t split_size,
                       int64_t *starts, int64_t *lengths, int cap)
{
    if (file_size < 0 || split_size <= 0 || (n > 0 && !offsets)) return FOURMC_E_ARG;
    int count = 0;
    auto emit = [&](int64_t s, int64_t len) {
        int64_t a = s, b = s + len;
        if (n > 0) {                                              // an empty index leaves the default split (:153-156)
            a = fourmc_index_align_slice_start(offsets, n, s, s + len);
            b = fourmc_index_align_slice_end(offsets, n, s + len, file_size);
            if (a == FOURMC_NOT_FOUND || b == FOURMC_NOT_FOUND) return;
        }
        if (count < cap && starts && lengths) { starts[count] = a; lengths[count] = b - a; }
        count++;
    };
    int64_t remaining = file_size, pos = 0;
    while ((double)remaining / (double)split_size > 1.1) { emit(pos, split_size); pos += split_size; remaining -= split_size; }
    if (remaining != 0) emit(pos, remaining);
    return count;
}

// FourMcInputStream.readIndex (FourMcInputStream.java:163-239): block offsets from the footer of a whole
// .4mc / .4mz file in host memory; the footer checksum is verified on the device.  Returns the number of
// blocks (offsets[] receives min(count, cap) of them), 0 for a file too short to hold an index, or
// FOURMC_E_CONTENT.
long long fourmc_read_index_host(fourmc_ctx *ctx, const void *file, size_t file_size, int64_t *offsets, size_t cap)
{
    if (!ctx || (!file && file_size)) return FOURMC_E_ARG;
    const uint8_t *f = (const uint8_t *)file;
    if (file_size < 12 + 20) return 0;
    const uint32_t fsize = be32(f + file_size - 12), magic = be32(f + file_size - 8), ck = be32(f + file_size - 4);
    if (magic != FOURMC_MAGIC_4MC && magic != FOURMC_MAGIC_4MZ) return FOURMC_E_CONTENT;
    if (fsize >= file_size - 12 || fsize < 20 || ((fsize - 20) & 3)) return FOURMC_E_CONTENT;
    const uint8_t *foot = f + file_size - fsize;
    if (be32(foot) != fsize || be32(foot + 4) != FOURMC_VERSION) return FOURMC_E_CONTENT;
    int st = FOURMC_OK;
    const uint32_t h = fourmc_xxh32(ctx, foot, fsize - 4, 0, &st);
    if (st != FOURMC_OK) return st;
    if (h != ck) return FOURMC_E_CONTENT;
    const size_t nb = (fsize - 20) / 4;
    int64_t cur = 0;
    for (size_t i = 0; i < nb; i++) {
        cur += be32(foot + 8 + 4 * i);
        if (i < cap && offsets) offsets[i] = cur;
    }
    return (long long)nb;
}

namespace {
// First line terminator of p[0, n) the way Hadoop's LineReader sees it (the reader behind FourMcLineRecordReader):
// a line ends at LF, at CR, or at CR LF.  *first = 2 * (index of the terminator's first byte) + (1 if it is CR LF);
// the smallest key is the first terminator.
__global__ void find_eol_kernel(const uint8_t *p, unsigned long long n, unsigned long long n_avail, unsigned long long *first)
{   // n: positions searched; n_avail >= n: bytes that may be looked at (the LF of a CR LF straddling the range's end)
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint8_t c = p[i];
        if (c == '\n' || c == '\r') {
            const unsigned long long crlf = (c == '\r' && i + 1 < n_avail && p[i + 1] == '\n') ? 1ull : 0ull;
            atomicMin(first, 2 * i + crlf);
            break;
        }
    }
}
}  // namespace

// FourMcLineRecordReader over one split (FourMcLineRecordReader.java:116-163): the bytes of all the records
// (lines, with their terminators) the reader would return for the split [start, start + length) of a whole
// .4mc / .4mz file in host memory -- when start != 0 the first (partial) line is skipped, and the line that is
// being read when the position passes the split end is finished from the following block(s).  The split's
// blocks are decoded on the device in one batch; the two line boundaries are found there too.
// Host -> device for a list of pieces.  Pinned sources go straight to the copy engine.  Pageable ones (a file read
// into ordinary memory, a Java byte array) would pass through the driver's own bounce buffer at the speed of ONE
// copying thread (r02k: 7 GB/s, the whole of a split read's time); here FOURMC_COPY_THREADS (default 4) threads move
// 1 MiB chunks into a pinned buffer of the context and queue each chunk's transfer as soon as it is there.
struct UpPiece { size_t dev_off; const uint8_t *src; size_t len; };
static int upload_pieces(fourmc_ctx *ctx, cudaStream_t st, uint8_t *dev, const std::vector<UpPiece> &pieces, const int slot)
{
    if (pieces.empty()) return FOURMC_OK;
    size_t total = 0;
    for (const UpPiece &q : pieces) total += q.len;
    cudaPointerAttributes at;
    bool is_pinned = false;
    if (cudaPointerGetAttributes(&at, pieces[0].src) == cudaSuccess) is_pinned = at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
    else (void)cudaGetLastError();
    static int n_threads = -1;
    if (n_threads < 0) {
        const char *e = getenv("FOURMC_COPY_THREADS");
        n_threads = e ? atoi(e) : 4;
        const int hw = (int)std::thread::hardware_concurrency();
        if (hw > 0 && n_threads > hw) n_threads = hw;
        if (n_threads > 16) n_threads = 16;
    }
    if (is_pinned || n_threads <= 0 || total < ((size_t)4 << 20)) {
        for (const UpPiece &q : pieces) if (q.len) CK(cudaMemcpyAsync(dev + q.dev_off, q.src, q.len, cudaMemcpyHostToDevice, st));
        return FOURMC_OK;
    }
    if (ctx->pin_up_cap[slot] < total) {
        if (ctx->pin_up[slot]) { cudaFreeHost(ctx->pin_up[slot]); ctx->pin_up[slot] = nullptr; ctx->pin_up_cap[slot] = 0; }
        const size_t cap = total + total / 8 + ((size_t)1 << 20);
        CK(cudaMallocHost(&ctx->pin_up[slot], cap));
        ctx->pin_up_cap[slot] = cap;
    }
    struct Chunk { size_t dev_off, pin_off; const uint8_t *src; size_t len; };
    std::vector<Chunk> chunks;
    size_t pin_off = 0;
    constexpr size_t CH = (size_t)1 << 20;
    for (const UpPiece &q : pieces)
        for (size_t o = 0; o < q.len; o += CH) {
            const size_t l = std::min(CH, q.len - o);
            chunks.push_back(Chunk{q.dev_off + o, pin_off, q.src + o, l});
            pin_off += l;
        }
    std::atomic<size_t> next{0};
    std::atomic<int> err{(int)cudaSuccess};
    uint8_t *pin = (uint8_t *)ctx->pin_up[slot];
    auto work = [&]() {
        cudaSetDevice(ctx->device);
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= chunks.size()) break;
            const Chunk &c = chunks[i];
            memcpy(pin + c.pin_off, c.src, c.len);
            const cudaError_t e = cudaMemcpyAsync(dev + c.dev_off, pin + c.pin_off, c.len, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) { err.store((int)e); break; }
        }
    };
    std::vector<std::thread> th;
    const int nt = (int)std::min<size_t>((size_t)n_threads, chunks.size());
    for (int t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (std::thread &t : th) t.join();
    if (err.load() != (int)cudaSuccess) return fail(ctx, FOURMC_E_CUDA, "upload", (cudaError_t)err.load());
    return FOURMC_OK;
}


static long long read_split_impl(fourmc_ctx *ctx, const void *file, size_t file_size, int64_t start, int64_t length,
                                 void *out, size_t out_capacity, const int slot)
{
    if (!ctx || !file || start < 0 || length < 0 || (!out && out_capacity)) return FOURMC_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint8_t *f = (const uint8_t *)file;
    if (file_size < 12) return FOURMC_E_CONTENT;
    const uint32_t magic = be32(f);
    if (magic != FOURMC_MAGIC_4MC && magic != FOURMC_MAGIC_4MZ) return FOURMC_E_CONTENT;
    const int codec = magic == FOURMC_MAGIC_4MZ ? CODEC_ZSTD : CODEC_LZ4;
    // the index is read (and its checksum verified on the device) once per file, not once per split
    if (ctx->idx_file != file || ctx->idx_size != file_size || memcmp(ctx->idx_tail, f + file_size - 12, 12) != 0) {
        ctx->idx_file = nullptr;
        const long long nb_ll = fourmc_read_index_host(ctx, file, file_size, nullptr, 0);
        if (nb_ll < 0) return nb_ll;
        ctx->idx_offs.assign((size_t)nb_ll, 0);
        if (nb_ll) fourmc_read_index_host(ctx, file, file_size, ctx->idx_offs.data(), ctx->idx_offs.size());
        ctx->idx_file = file; ctx->idx_size = file_size; memcpy(ctx->idx_tail, f + file_size - 12, 12);
    }
    const std::vector<int64_t> &offs = ctx->idx_offs;
    if (offs.empty()) return 0;
    const int n = (int)offs.size();
    const int64_t end = start + length;
    // first block of the split: the reader seeks to `start`, which the planner put on a block start
    int b0 = 0;
    if (start != 0) {
        const int64_t s = fourmc_index_find_next_position(offs.data(), n, start);
        if (s == FOURMC_NOT_FOUND || s >= end) return 0;
        b0 = (int)(std::lower_bound(offs.begin(), offs.end(), s) - offs.begin());
    }
    const int b1 = (int)(std::lower_bound(offs.begin(), offs.end(), end) - offs.begin());      // blocks [b0, b1) start before `end`
    if (b0 >= b1) return 0;
    // block headers (native/4mc.c:603-668 field order), the split's blocks plus the ones after it while a line is open
    struct HB { uint64_t src; uint32_t c, u, x; };
    auto header = [&](int b, HB &h) -> bool {
        const uint64_t o = (uint64_t)offs[b];
        if (o + 12 > file_size) return false;
        h.u = be32(f + o); h.c = be32(f + o + 4); h.x = be32(f + o + 8); h.src = o + 12;
        return h.c <= FOURMC_BLOCKSIZE && (h.u == h.c || h.u <= FOURMC_BLOCKSIZE) && h.src + h.c <= file_size && h.u != 0;
    };
    cudaStream_t st = ctx->stream;
    DecWs &ws = ctx->dec[slot];
    int r;
    if ((r = pinned_scratch(ctx, 4096))) return r;
    int t1 = std::min(n, b1 + 1);                                  // decode through block t1 - 1
    for (;;) {
        const uint32_t cnt = (uint32_t)(t1 - b0);
        std::vector<HB> hb(cnt);
        uint64_t total_u = 0;
        for (uint32_t i = 0; i < cnt; i++) { if (!header(b0 + (int)i, hb[i])) return FOURMC_E_CONTENT; total_u += hb[i].u; }
        const uint64_t src0 = hb[0].src, src1 = hb[cnt - 1].src + hb[cnt - 1].c;
        if ((r = ensure(ctx, ctx->stage_in[slot], (size_t)(src1 - src0) + 64))) return r;
        if ((r = ensure(ctx, ctx->stage_out[slot], (size_t)total_u + 64))) return r;
        const size_t tb = (size_t)cnt * 33 + 64 + 16;
        if ((r = ensure(ctx, ws.tables, tb))) return r;
        std::vector<uint8_t> ht((size_t)cnt * 28);
        uint64_t *t_src = (uint64_t *)ht.data(), *t_dst = t_src + cnt;
        uint32_t *t_c = (uint32_t *)(t_dst + cnt), *t_u = t_c + cnt, *t_x = t_u + cnt;
        uint64_t dpos = 0, u_split = 0;
        for (uint32_t i = 0; i < cnt; i++) {
            t_src[i] = hb[i].src - src0; t_dst[i] = dpos; t_c[i] = hb[i].c; t_u[i] = hb[i].u; t_x[i] = hb[i].x;
            dpos += hb[i].u;
            if ((int)i < b1 - b0) u_split = dpos;                  // decoded bytes of the split's own blocks
        }
        uint8_t *d_tb = (uint8_t *)ws.tables.p;
        const std::vector<UpPiece> up{UpPiece{0, f + src0, (size_t)(src1 - src0)}};
        r = upload_pieces(ctx, st, (uint8_t *)ctx->stage_in[slot].p, up, slot);
        if (r) return r;
        CK(cudaMemcpyAsync(d_tb, ht.data(), ht.size(), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));                             // ht is pageable host memory
        const uint64_t *d_so = (const uint64_t *)d_tb, *d_do = d_so + cnt;
        const uint32_t *d_c = (const uint32_t *)(d_do + cnt), *d_u = d_c + cnt, *d_x = d_u + cnt;
        int32_t *d_osz = (int32_t *)(d_x + cnt);
        uint8_t *d_st = (uint8_t *)(d_osz + cnt);
        unsigned long long *d_first = (unsigned long long *)(((uintptr_t)(d_st + cnt) + 15) & ~(uintptr_t)15);
        if ((r = dec_batch(ctx, st, ws, cnt, ctx->stage_in[slot].p, d_so, d_c, d_u, d_x, 1, ctx->stage_out[slot].p, d_do, d_osz, d_st, codec, 1)))
            return r;
        // what the blocks really decoded to (a block may decode short, native/4mc.c:661-666; the batch closed the gaps)
        std::vector<uint8_t> status(cnt);
        std::vector<int32_t> osz(cnt);
        CK(cudaMemcpyAsync(status.data(), d_st, cnt, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(osz.data(), d_osz, (size_t)cnt * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (uint32_t i = 0; i < cnt; i++) if (status[i] != FOURMC_BLOCK_OK) return FOURMC_E_CONTENT;
        total_u = 0; u_split = 0;
        for (uint32_t i = 0; i < cnt; i++) { total_u += (uint64_t)osz[i]; if ((int)i < b1 - b0) u_split = total_u; }
        // first line terminator of the split (skipped line) and the first one at or after the split's end
        CK(cudaMemsetAsync(d_first, 0xff, 16, st));
        const uint8_t *d_out = (const uint8_t *)ctx->stage_out[slot].p;
        if (start != 0 && u_split)
            KL("find_eol_kernel", st, find_eol_kernel<<<1024, 256, 0, st>>>(d_out, u_split, total_u, d_first));
        if (total_u > u_split)
            KL("find_eol_kernel", st, find_eol_kernel<<<256, 256, 0, st>>>(d_out + u_split, total_u - u_split, total_u - u_split, d_first + 1));
        uint8_t *hs = (uint8_t *)ctx->pinned;
        CK(cudaMemcpyAsync(hs, d_first, 16, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const unsigned long long first_nl = ((unsigned long long *)hs)[0], tail_nl = ((unsigned long long *)hs)[1];
        uint64_t from = 0, to;
        if (start != 0) {
            // the skipped line must end inside the split's own blocks, else the reader's position is past `end`
            if (first_nl == ~0ull) return 0;
            from = (first_nl >> 1) + 1 + (first_nl & 1);
        }
        if (tail_nl != ~0ull) to = u_split + (tail_nl >> 1) + 1 + (tail_nl & 1);
        else if (t1 < n) { t1 = std::min(n, t1 + 4); continue; }    // the open line runs through every block decoded so far
        else to = total_u;                                         // end of the file: the last line has no terminator
        if (to < from) to = from;
        if (to - from > out_capacity) return fail(ctx, FOURMC_E_OUTPUT, "destination too small");
        if (to > from) {
            CK(cudaMemcpyAsync(out, d_out + from, (size_t)(to - from), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
        return (long long)(to - from);
    }
}

long long fourmc_read_split_lines_host(fourmc_ctx *ctx, const void *file, size_t file_size, int64_t start, int64_t length,
                                       void *out, size_t out_capacity)
{
    return read_split_impl(ctx, file, file_size, start, length, out, out_capacity, 0);
}

namespace {
struct EolJob { unsigned long long off, n, n_avail; };        // a search of d_out + off .. for the first line end
__global__ void find_eol_batch_kernel(const uint8_t *base, const EolJob *jobs, unsigned long long *first)
{
    const EolJob j = jobs[blockIdx.y];
    const uint8_t *p = base + j.off;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < j.n; i += stride) {
        const uint8_t c = p[i];
        if (c == '\n' || c == '\r') {
            const unsigned long long crlf = (c == '\r' && i + 1 < j.n_avail && p[i + 1] == '\n') ? 1ull : 0ull;
            atomicMin(&first[blockIdx.y], 2 * i + crlf);
            break;
        }
    }
}
}  // namespace

// Many splits of one file in one call (what a node running many map tasks over one file asks for; BASELINE.json
// configs[4]).  Same records per split as fourmc_read_split_lines_host; the blocks of ALL the splits are decoded as one
// batch, which is what the GPU is good at -- a split alone is two or three blocks and those take 13 ms however idle the
// chip is.  The records of split i land at out + out_offsets[i] (back to back, split order); out_offsets has
// n_splits + 1 entries, the last one is the total, which is also the return value.
long long fourmc_read_splits_lines_host(fourmc_ctx *ctx, const void *file, size_t file_size, int n_splits, const int64_t *starts,
                                        const int64_t *lengths, void *out, size_t out_capacity, int64_t *out_offsets)
{
    if (!ctx || !file || n_splits < 0 || (n_splits && (!starts || !lengths)) || !out_offsets || (!out && out_capacity)) return FOURMC_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint8_t *f = (const uint8_t *)file;
    out_offsets[0] = 0;
    if (n_splits == 0) return 0;
    if (file_size < 12) return FOURMC_E_CONTENT;
    const uint32_t magic = be32(f);
    if (magic != FOURMC_MAGIC_4MC && magic != FOURMC_MAGIC_4MZ) return FOURMC_E_CONTENT;
    const int codec = magic == FOURMC_MAGIC_4MZ ? CODEC_ZSTD : CODEC_LZ4;
    if (ctx->idx_file != file || ctx->idx_size != file_size || memcmp(ctx->idx_tail, f + file_size - 12, 12) != 0) {
        ctx->idx_file = nullptr;
        const long long nb_ll = fourmc_read_index_host(ctx, file, file_size, nullptr, 0);
        if (nb_ll < 0) return nb_ll;
        ctx->idx_offs.assign((size_t)nb_ll, 0);
        if (nb_ll) fourmc_read_index_host(ctx, file, file_size, ctx->idx_offs.data(), ctx->idx_offs.size());
        ctx->idx_file = file; ctx->idx_size = file_size; memcpy(ctx->idx_tail, f + file_size - 12, 12);
    }
    const std::vector<int64_t> &offs = ctx->idx_offs;
    const int n = (int)offs.size();
    struct HB { uint64_t src; uint32_t c, u, x; };
    auto header = [&](int b, HB &h) -> bool {
        const uint64_t o = (uint64_t)offs[b];
        if (o + 12 > file_size) return false;
        h.u = be32(f + o); h.c = be32(f + o + 4); h.x = be32(f + o + 8); h.src = o + 12;
        return h.c <= FOURMC_BLOCKSIZE && (h.u == h.c || h.u <= FOURMC_BLOCKSIZE) && h.src + h.c <= file_size && h.u != 0;
    };
    // per split: its blocks [b0, b1) plus one more that finishes the open line
    struct Sp { int b0, b1, t1; uint32_t first_item; uint64_t in_off, out_off, u_split, u_total; bool empty, fallback; };
    std::vector<Sp> sp((size_t)n_splits);
    std::vector<HB> hb;
    uint64_t in_total = 0, out_total = 0;
    for (int i = 0; i < n_splits; i++) {
        Sp &q = sp[i];
        q.empty = true; q.fallback = false;
        if (n == 0 || starts[i] < 0 || lengths[i] < 0) { if (starts[i] < 0 || lengths[i] < 0) return FOURMC_E_ARG; continue; }
        const int64_t start = starts[i], end = start + lengths[i];
        int b0 = 0;
        if (start != 0) {
            const int64_t s0 = fourmc_index_find_next_position(offs.data(), n, start);
            if (s0 == FOURMC_NOT_FOUND || s0 >= end) continue;
            b0 = (int)(std::lower_bound(offs.begin(), offs.end(), s0) - offs.begin());
        }
        const int b1 = (int)(std::lower_bound(offs.begin(), offs.end(), end) - offs.begin());
        if (b0 >= b1) continue;
        q.empty = false; q.b0 = b0; q.b1 = b1; q.t1 = std::min(n, b1 + 1);
        q.first_item = (uint32_t)hb.size(); q.in_off = in_total; q.out_off = out_total; q.u_split = 0; q.u_total = 0;
        for (int b = b0; b < q.t1; b++) {
            HB h;
            if (!header(b, h)) return FOURMC_E_CONTENT;
            hb.push_back(h);
            q.u_total += h.u;
            if (b < b1) q.u_split = q.u_total;
        }
        const HB &h0 = hb[q.first_item], &h1 = hb.back();
        in_total += ((h1.src + h1.c - h0.src) + 15) & ~(uint64_t)15;
        out_total += (q.u_total + 15) & ~(uint64_t)15;
    }
    const uint32_t cnt = (uint32_t)hb.size();
    std::vector<std::vector<uint8_t>> fb((size_t)n_splits);        // records of the splits that went through the single-split call
    std::vector<uint64_t> from((size_t)n_splits, 0), to((size_t)n_splits, 0);
    cudaStream_t st = ctx->stream;
    DecWs &ws = ctx->dec[0];
    int r;
    if (cnt) {
        if ((r = ensure(ctx, ctx->stage_in[0], (size_t)in_total + 64))) return r;
        if ((r = ensure(ctx, ctx->stage_out[0], (size_t)out_total + 64))) return r;
        const size_t tb = (size_t)cnt * 33 + 64 + (size_t)n_splits * 2 * (sizeof(EolJob) + 8) + 64;
        if ((r = ensure(ctx, ws.tables, tb))) return r;
        std::vector<uint8_t> ht((size_t)cnt * 28);
        uint64_t *t_src = (uint64_t *)ht.data(), *t_dst = t_src + cnt;
        uint32_t *t_c = (uint32_t *)(t_dst + cnt), *t_u = t_c + cnt, *t_x = t_u + cnt;
        std::vector<EolJob> jobs((size_t)n_splits * 2, EolJob{0, 0, 0});
        std::vector<UpPiece> up;
        for (int i = 0; i < n_splits; i++) {
            const Sp &q = sp[i];
            if (q.empty) continue;
            const HB &h0 = hb[q.first_item];
            const uint32_t k = (uint32_t)(q.t1 - q.b0);
            const uint64_t bytes = hb[q.first_item + k - 1].src + hb[q.first_item + k - 1].c - h0.src;
            up.push_back(UpPiece{(size_t)q.in_off, f + h0.src, (size_t)bytes});
            uint64_t dpos = q.out_off;
            for (uint32_t j = 0; j < k; j++) {
                const HB &h = hb[q.first_item + j];
                t_src[q.first_item + j] = q.in_off + (h.src - h0.src); t_dst[q.first_item + j] = dpos;
                t_c[q.first_item + j] = h.c; t_u[q.first_item + j] = h.u; t_x[q.first_item + j] = h.x;
                dpos += h.u;
            }
            if (starts[i] != 0) jobs[2 * i] = EolJob{q.out_off, q.u_split, q.u_total};
            if (q.u_total > q.u_split) jobs[2 * i + 1] = EolJob{q.out_off + q.u_split, q.u_total - q.u_split, q.u_total - q.u_split};
        }
        uint8_t *d_tb = (uint8_t *)ws.tables.p;
        CK(cudaMemcpyAsync(d_tb, ht.data(), ht.size(), cudaMemcpyHostToDevice, st));
        if ((r = upload_pieces(ctx, st, (uint8_t *)ctx->stage_in[0].p, up, 0))) return r;
        const uint64_t *d_so = (const uint64_t *)d_tb, *d_do = d_so + cnt;
        const uint32_t *d_c = (const uint32_t *)(d_do + cnt), *d_u = d_c + cnt, *d_x = d_u + cnt;
        int32_t *d_osz = (int32_t *)(d_x + cnt);
        uint8_t *d_st = (uint8_t *)(d_osz + cnt);
        EolJob *d_jobs = (EolJob *)(((uintptr_t)(d_st + cnt) + 15) & ~(uintptr_t)15);
        unsigned long long *d_first = (unsigned long long *)(d_jobs + (size_t)n_splits * 2);
        CK(cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(EolJob), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));                                 // ht and jobs are pageable
        if ((r = dec_batch(ctx, st, ws, cnt, ctx->stage_in[0].p, d_so, d_c, d_u, d_x, 1, ctx->stage_out[0].p, d_do, d_osz, d_st, codec)))
            return r;
        CK(cudaMemsetAsync(d_first, 0xff, (size_t)n_splits * 16, st));
        KL("find_eol_batch_kernel", st, find_eol_batch_kernel<<<dim3(32, (unsigned)n_splits * 2), 256, 0, st>>>(
            (const uint8_t *)ctx->stage_out[0].p, d_jobs, d_first));
        std::vector<uint8_t> status(cnt);
        std::vector<int32_t> osz(cnt);
        std::vector<unsigned long long> first((size_t)n_splits * 2);
        CK(cudaMemcpyAsync(status.data(), d_st, cnt, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(osz.data(), d_osz, (size_t)cnt * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(first.data(), d_first, first.size() * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (uint32_t i = 0; i < cnt; i++) if (status[i] != FOURMC_BLOCK_OK) return FOURMC_E_CONTENT;
        for (int i = 0; i < n_splits; i++) {
            Sp &q = sp[i];
            if (q.empty) continue;
            // a block that decoded short, or a line still open after the extra block: the single-split call knows how
            for (int j = 0; j < q.t1 - q.b0; j++) if ((uint32_t)osz[q.first_item + j] != hb[q.first_item + j].u) q.fallback = true;
            const unsigned long long first_nl = first[2 * i], tail_nl = first[2 * i + 1];
            if (starts[i] != 0) {
                if (first_nl == ~0ull) { q.empty = true; continue; }
                from[i] = (first_nl >> 1) + 1 + (first_nl & 1);
            }
            if (tail_nl != ~0ull) to[i] = q.u_split + (tail_nl >> 1) + 1 + (tail_nl & 1);
            else if (q.t1 < n) q.fallback = true;
            else to[i] = q.u_total;
            if (to[i] < from[i]) to[i] = from[i];
        }
    }
    // splits that need the general path
    for (int i = 0; i < n_splits; i++) {
        if (sp[i].empty || !sp[i].fallback) continue;
        size_t cap = (size_t)(sp[i].u_total + 4 * FOURMC_BLOCKSIZE);
        for (;;) {
            fb[i].resize(cap);
            const long long got = read_split_impl(ctx, file, file_size, starts[i], lengths[i], fb[i].data(), cap, 1);   // slot 1: slot 0 holds the batch
            if (got == FOURMC_E_OUTPUT && cap < ((size_t)1 << 34)) { cap *= 4; continue; }
            if (got < 0) return got;
            fb[i].resize((size_t)got);
            break;
        }
    }
    // pack in split order
    uint64_t pos = 0;
    for (int i = 0; i < n_splits; i++) {
        out_offsets[i] = (int64_t)pos;
        if (sp[i].empty) continue;
        const uint64_t len = sp[i].fallback ? fb[i].size() : to[i] - from[i];
        if (pos + len > out_capacity) return fail(ctx, FOURMC_E_OUTPUT, "destination too small");
        if (sp[i].fallback) memcpy((uint8_t *)out + pos, fb[i].data(), (size_t)len);
        else if (len) CK(cudaMemcpyAsync((uint8_t *)out + pos, (const uint8_t *)ctx->stage_out[0].p + sp[i].out_off + from[i], (size_t)len,
                                         cudaMemcpyDeviceToHost, st));
        pos += len;
    }
    out_offsets[n_splits] = (int64_t)pos;
    CK(cudaStreamSynchronize(st));
    return (long long)pos;
}

// ZSTD_decompress on one block (native/4mc.c:810, native/jniZstdDecompressor.c): decoded size, or a
// negative value when ZSTD_isError() would be true for the reference.
long long fourmc_zstd_decompress(fourmc_ctx *ctx, const void *src, size_t compressed_size, void *dst, size_t dst_capacity)
{
    if (!ctx || !src || (!dst && dst_capacity)) return FOURMC_E_ARG;
    if (compressed_size > (size_t)(8 << 20) || dst_capacity > (size_t)(1 << 30))
        return fail(ctx, FOURMC_E_ARG, "per-block calls take at most one 4 MiB block");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DecWs &ws = ctx->dec[0];
    int r;
    if ((r = ensure(ctx, ctx->stage_in[0], compressed_size + 64))) return r;
    if ((r = ensure(ctx, ctx->stage_out[0], dst_capacity + 64))) return r;
    if ((r = ensure(ctx, ws.desc, sizeof(BlockDesc)))) return r;
    if ((r = ensure(ctx, ws.status, 16))) return r;
    if ((r = ensure(ctx, ws.outsize, 16))) return r;
    if ((r = pinned_scratch(ctx, 4096))) return r;
    if (compressed_size) CK(cudaMemcpyAsync(ctx->stage_in[0].p, src, compressed_size, cudaMemcpyHostToDevice, st));
    BlockDesc *hd = (BlockDesc *)ctx->pinned;
    hd->src = (const uint8_t *)ctx->stage_in[0].p; hd->dst = (uint8_t *)ctx->stage_out[0].p;
    hd->csize = (uint32_t)compressed_size; hd->usize = (uint32_t)dst_capacity; hd->chunk_base = 0; hd->stored = 0;
    CK(cudaMemcpyAsync(ws.desc.p, hd, sizeof(BlockDesc), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ws.status.p, 0, 16, st));
    if ((r = dec_blocks(ctx, st, ws, 1, 0, 0, (int32_t *)ws.outsize.p, nullptr, nullptr, CODEC_ZSTD))) return r;
    int32_t *hr = (int32_t *)((uint8_t *)ctx->pinned + 256);
    CK(cudaMemcpyAsync(hr, ws.outsize.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int32_t res = *hr;
    if (res > 0) {
        CK(cudaMemcpyAsync(dst, ctx->stage_out[0].p, (size_t)res, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return res;
}

}  // extern "C"

#include "fileio.h"
