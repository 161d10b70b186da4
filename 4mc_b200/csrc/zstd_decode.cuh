// zstd_decode.cuh -- 4mz block decode kernels.
//
// Reference behaviour: ZSTD_decompress(out, usize, in, csize) per 4mz block, native/4mc.c:810.
//
//  zstd_frames_warp_kernel   one WARP per zstd frame (= per 4mz block), the fast path for
//      well-formed single frames.  The frame's blocks are decoded in order with the entropy tables
//      of the current block in shared memory (treeless literals and "repeat" sequence tables just
//      keep them): lane 0 parses headers and builds the tables (zstd_decode.h, the same source as
//      the serial decoder), lanes 0..3 decode the four Huffman streams, lane 0 decodes the FSE
//      sequence stream 32 sequences at a time and the whole warp executes them (literal and match
//      copies, overlap handled per offset).  Throughput comes from the number of frames in flight.
//      Anything unusual -- a damaged stream, several frames or a skippable frame in one payload, a
//      Huffman table log of 12 -- is not judged here: the frame is marked ZD_RETRY.
//  zstd_frames_kernel        one THREAD per frame, the exact serial decoder (zstd_decode.h
//      fmz::decompress, pinned to the reference's accept / reject behaviour on damaged input).  Runs
//      only for frames marked ZD_RETRY, so return values and error verdicts are always the serial ones.
#pragma once

#include "fm_common.cuh"
#include "lz4_decode.cuh"
#include "zstd_decode.h"

namespace fm {

constexpr int ZD_WARPS = 4;                         // frames per CTA in the warp kernel
constexpr int32_t ZD_RETRY = INT32_MIN;             // result marker: "decide serially"
constexpr int ZD_BATCH = 32;                        // sequences decoded by lane 0 per round
constexpr int ZD_LONG = 48;                         // copies above this many bytes are done by the whole warp

// Per-frame global scratch of the warp kernel = the serial decoder's fmz::Work (its table-building
// scratch and its 128 KiB literal buffer; the decode tables themselves live in shared memory here).
using ZdScratch = fmz::Work;

struct ZdWarp {                                     // per warp, shared memory; the members zstd_decode.h's table builders use
    static constexpr int HUF_MAX_LOG = 11;
    fmz::HufEntry huf[2048];
    fmz::SeqEntry ll[512], ml[512], of[256];
    int huf_log, ll_log, ml_log, of_log, huf_ok, ll_ok, ml_ok, of_ok, huf_x2;
    uint32_t rep[3];
    short *norm;
    uint16_t *symnext;
    uint8_t *weights;
    uint32_t *rank;
    fmz::WtEntry *wt;
    int32_t s_ll[ZD_BATCH], s_ml[ZD_BATCH];
    uint32_t s_off[ZD_BATCH];
};
constexpr size_t ZD_SMEM = ZD_WARPS * sizeof(ZdWarp);

// One Huffman stream, single-symbol semantics (huf_decompress.c:555-640): `count` symbols, the
// stream must end exactly at its first bit.  A 64-bit container is refilled every four symbols;
// the last few symbols (container at the start of the stream) go through the generic reader.
__device__ __forceinline__ int zd_huf_stream(const fmz::HufEntry *tbl, int log, uint8_t *dst, int count,
                                             const uint8_t *src, long long n)
{
    if (n < 1 || src[n - 1] == 0) return fmz::ERR_CORRUPT;
    long long pos = 8 * (n - 1) + fmz::highbit(src[n - 1]);       // unread bits
    int i = 0;
    while (count - i >= 4 && pos >= 64 + 7) {
        // container = 64 bits whose top bit is bit (pos - 1) of the stream
        const long long lo = pos - 64;
        const uintptr_t a = (uintptr_t)(src + (lo >> 3));
        const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(a & 3) * 8;
        const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
        // 96 bits starting at byte (lo >> 3); drop (lo & 7) low bits
        const uint32_t b0 = __funnelshift_r(w0, w1, sh), b1 = __funnelshift_r(w1, w2, sh), b2 = __funnelshift_r(w2, w3, sh);
        const uint32_t s2 = (uint32_t)(lo & 7);
        uint64_t c = (uint64_t)__funnelshift_r(b0, b1, s2) | ((uint64_t)__funnelshift_r(b1, b2, s2) << 32);
        uint32_t used = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const fmz::HufEntry e = tbl[(uint32_t)((c << used) >> (64 - log))];
            dst[i + k] = e.sym;
            used += e.nbits;
        }
        i += 4;
        pos -= used;
    }
    fmz::BackBits b;
    b.p = src; b.pos = pos;
    for (; i < count; i++) {
        fmz::BackBits t = b;
        const fmz::HufEntry e = tbl[t.read(log)];
        dst[i] = e.sym;
        b.pos -= e.nbits;
    }
    return b.pos == 0 ? 0 : fmz::ERR_CORRUPT;
}

__device__ __forceinline__ void zd_warp_copy(uint8_t *d, const uint8_t *s, long long n, int lane)
{
    for (long long i = lane; i < n; i += 32) d[i] = s[i];
}

// One compressed block (zstd_decode.h decode_block, same checks in the same places); returns the
// new output position or < 0.  All lanes carry the same scalars.
__device__ long long zd_block(ZdWarp &w, ZdScratch &sc, const fmz::Tables &T, uint8_t *dst, long long op, long long cap,
                              const uint8_t *src, long long n, int lane)
{
    using namespace fmz;
    if (n < 1) return ERR_CORRUPT;
    const int ltype = src[0] & 3, sf = (src[0] >> 2) & 3;
    long long hdr, regen, comp = 0;
    const uint8_t *lit = nullptr;
    bool own_lit = true;
    if (ltype < 2) {
        if (sf == 0 || sf == 2) { hdr = 1; regen = src[0] >> 3; }
        else if (sf == 1) { if (n < 2) return ERR_CORRUPT; hdr = 2; regen = (src[0] >> 4) + ((long long)src[1] << 4); }
        else { if (n < 3) return ERR_CORRUPT; hdr = 3; regen = (src[0] >> 4) + ((long long)src[1] << 4) + ((long long)src[2] << 12); }
        if (regen > BLOCK_MAX) return ERR_CORRUPT;
        if (ltype == 0) {
            if (hdr + regen > n) return ERR_CORRUPT;
            lit = src + hdr;
            own_lit = hdr + regen + 32 > n;
            hdr += regen;
        } else {
            if (hdr + 1 > n) return ERR_CORRUPT;
            const uint8_t v = src[hdr];
            for (long long i = lane; i < regen; i += 32) sc.lit[i] = v;
            lit = sc.lit;
            hdr += 1;
        }
    } else {
        if (n < 5 && sf == 3) return ERR_CORRUPT;
        if (n < 3) return ERR_CORRUPT;
        int streams = 4;
        if (sf <= 1) {
            const uint32_t v = src[0] | (src[1] << 8) | (src[2] << 16);
            hdr = 3; regen = (v >> 4) & 0x3FF; comp = (v >> 14) & 0x3FF;
            streams = sf == 0 ? 1 : 4;
        } else if (sf == 2) {
            if (n < 4) return ERR_CORRUPT;
            const uint32_t v = src[0] | (src[1] << 8) | (src[2] << 16) | ((uint32_t)src[3] << 24);
            hdr = 4; regen = (v >> 4) & 0x3FFF; comp = v >> 18;
        } else {
            if (n < 5) return ERR_CORRUPT;
            const uint64_t v = (uint64_t)src[0] | ((uint64_t)src[1] << 8) | ((uint64_t)src[2] << 16) | ((uint64_t)src[3] << 24) | ((uint64_t)src[4] << 32);
            hdr = 5; regen = (long long)((v >> 4) & 0x3FFFF); comp = (long long)(v >> 22);
        }
        if (regen > BLOCK_MAX) return ERR_CORRUPT;
        if (hdr + comp > n) return ERR_CORRUPT;
        const uint8_t *cs = src + hdr;
        long long cn = comp;
        if (ltype == 2) {
            int used = 0;
            if (lane == 0) used = read_huffman(w, cs, cn);
            used = __shfl_sync(FM_FULL, used, 0);
            if (used < 0) return used;
            cs += used; cn -= used;
        } else if (!w.huf_ok) return ERR_CORRUPT;
        __syncwarp();
        const int log = w.huf_log;
        int e = 0;
        if (streams == 1) {
            if (lane == 0) e = zd_huf_stream(w.huf, log, sc.lit, (int)regen, cs, cn);
        } else {
            if (cn < 10) return ERR_CORRUPT;
            const long long s1 = cs[0] | (cs[1] << 8), s2 = cs[2] | (cs[3] << 8), s3 = cs[4] | (cs[5] << 8);
            const long long s4 = cn - 6 - s1 - s2 - s3;
            if (s4 < 1 || s1 < 1 || s2 < 1 || s3 < 1) return ERR_CORRUPT;
            const int seg = (int)((regen + 3) / 4);
            if (3LL * seg > regen) return ERR_CORRUPT;
            const uint8_t *q = cs + 6;
            if (lane < 4) {
                const long long so = lane == 0 ? 0 : lane == 1 ? s1 : lane == 2 ? s1 + s2 : s1 + s2 + s3;
                const long long sn = lane == 0 ? s1 : lane == 1 ? s2 : lane == 2 ? s3 : s4;
                const int cnt = lane < 3 ? seg : (int)(regen - 3 * seg);
                e = zd_huf_stream(w.huf, log, sc.lit + (size_t)lane * seg, cnt, q + so, sn);
            }
        }
        if (__any_sync(FM_FULL, e < 0)) return ERR_CORRUPT;
        lit = sc.lit;
        hdr += comp;
    }
    __syncwarp();                                              // regenerated literals visible to every lane
    long long out_cap = cap;
    if (own_lit && cap - op > BLOCK_MAX + 32 + regen + 32) out_cap = op + BLOCK_MAX + 32;
    // ---- sequences section header
    const uint8_t *sp = src + hdr;
    long long sn = n - hdr;
    if (sn < 1) return ERR_SRCSIZE;
    long long nseq = sp[0];
    long long shdr = 1;
    if (nseq == 0 && sn != 1) return ERR_SRCSIZE;
    if (nseq >= 128) {
        if (nseq == 255) { if (sn < 3) return ERR_SRCSIZE; nseq = sp[1] + (sp[2] << 8) + 0x7F00; shdr = 3; }
        else { if (sn < 2) return ERR_SRCSIZE; nseq = ((nseq - 128) << 8) + sp[1]; shdr = 2; }
    }
    long long lit_pos = 0;
    if (nseq > 0) {
        if (shdr + 1 > sn) return ERR_SRCSIZE;
        const int modes = sp[shdr];
        shdr += 1;
        long long tb = 0;                                      // table descriptions: bytes used, or < 0
        if (lane == 0) {
            long long h = shdr;
            int used;
            if ((used = build_mode((modes >> 6) & 3, w.ll, &w.ll_log, &w.ll_ok, w, T, 0, sp + h, sn - h)) < 0) tb = used;
            else {
                h += used;
                if ((used = build_mode((modes >> 4) & 3, w.of, &w.of_log, &w.of_ok, w, T, 1, sp + h, sn - h)) < 0) tb = used;
                else {
                    h += used;
                    if ((used = build_mode((modes >> 2) & 3, w.ml, &w.ml_log, &w.ml_ok, w, T, 2, sp + h, sn - h)) < 0) tb = used;
                    else tb = h + used - shdr;
                }
            }
        }
        tb = __shfl_sync(FM_FULL, tb, 0);
        if (tb < 0) return tb;
        shdr += tb;
        __syncwarp();
        // ---- lane 0 decodes, the warp executes
        SeqBits b;
        uint32_t sl = 0, so = 0, sm = 0, r0 = 0, r1 = 0, r2 = 0;
        long long d_op = op, d_lit = 0;                        // lane 0's running positions (for the checks)
        int err = 0;
        if (lane == 0) {
            if (!b.init(sp + shdr, sn - shdr)) err = ERR_CORRUPT;
            else {
                sl = b.read(w.ll_log); b.reload();
                so = b.read(w.of_log); b.reload();
                sm = b.read(w.ml_log); b.reload();
                r0 = w.rep[0]; r1 = w.rep[1]; r2 = w.rep[2];
            }
        }
        if (__shfl_sync(FM_FULL, err, 0)) return ERR_CORRUPT;
        for (long long k0 = 0; k0 < nseq; k0 += ZD_BATCH) {
            const int cnt = (int)(nseq - k0 < ZD_BATCH ? nseq - k0 : ZD_BATCH);
            if (lane == 0) {
                for (int j = 0; j < cnt; j++) {
                    const SeqEntry el = w.ll[sl], eo = w.of[so], em = w.ml[sm];
                    uint32_t offset;
                    if (eo.extra > 1) {
                        offset = eo.base + b.read_fast(eo.extra);
                        r2 = r1; r1 = r0; r0 = offset;
                    } else {
                        const uint32_t ll0 = el.base == 0;
                        if (eo.extra == 0) {
                            offset = ll0 ? r1 : r0;
                            r1 = ll0 ? r0 : r1;
                            r0 = offset;
                        } else {
                            const uint32_t code = eo.base + ll0 + b.read_fast(1);
                            uint32_t t = code == 3 ? r0 - 1 : (code == 1 ? r1 : code == 2 ? r2 : r0);
                            t += !t;
                            if (code != 1) r2 = r1;
                            r1 = r0;
                            r0 = offset = t;
                        }
                    }
                    long long mlen = em.base, llen = el.base;
                    if (em.extra) mlen += b.read_fast(em.extra);
                    if (el.extra + em.extra + eo.extra >= 31) b.reload();
                    if (el.extra) llen += b.read_fast(el.extra);
                    sl = el.next + b.read(el.nbits);
                    sm = em.next + b.read(em.nbits);
                    so = eo.next + b.read(eo.nbits);
                    if (llen + mlen > out_cap - d_op) { err = ERR_DSTSIZE; break; }
                    if (llen > regen - d_lit) { err = ERR_CORRUPT; break; }
                    d_op += llen; d_lit += llen;
                    if ((long long)offset > d_op) { err = ERR_CORRUPT; break; }
                    d_op += mlen;
                    w.s_ll[j] = (int32_t)llen; w.s_ml[j] = (int32_t)mlen; w.s_off[j] = offset;
                    if (k0 + j + 1 < nseq) b.reload();
                }
            }
            if (__shfl_sync(FM_FULL, err, 0)) return ERR_CORRUPT;
            __syncwarp();
            // ---- execute the batch: lane j owns sequence j
            {
                const int my_ll = lane < cnt ? w.s_ll[lane] : 0, my_ml = lane < cnt ? w.s_ml[lane] : 0;
                const long long my_off = lane < cnt ? (long long)w.s_off[lane] : 1;
                const int incl = warp_incl_scan_add(my_ll + my_ml), incl_ll = warp_incl_scan_add(my_ll);
                uint8_t *my_dst = dst + op + (incl - my_ll - my_ml);                 // my literals go here, my match right after
                const uint8_t *my_lit = lit + lit_pos + (incl_ll - my_ll);
                const int batch_bytes = __shfl_sync(FM_FULL, incl, 31), batch_lits = __shfl_sync(FM_FULL, incl_ll, 31);
                // A: literals -- sources are never in dst, so all of them go at once
                if (my_ll <= ZD_LONG) copy_batched(my_dst, my_lit, my_ll);
                for (unsigned mm = __ballot_sync(FM_FULL, my_ll > ZD_LONG); mm; mm &= mm - 1) {
                    const int l = __ffs(mm) - 1;
                    const int nn = __shfl_sync(FM_FULL, my_ll, l);
                    uint8_t *dp = (uint8_t *)__shfl_sync(FM_FULL, (unsigned long long)my_dst, l);
                    const uint8_t *spp = (const uint8_t *)__shfl_sync(FM_FULL, (unsigned long long)my_lit, l);
                    for (int i = lane; i < nn; i += 32) dp[i] = spp[i];
                }
                // B: matches whose source ends before this batch's output: independent of the batch
                uint8_t *m_dst = my_dst + my_ll;
                const uint8_t *m_src = m_dst - my_off;
                const bool indep = m_src + my_ml <= dst + op;
                if (indep && my_ml <= ZD_LONG) copy_batched(m_dst, m_src, my_ml);
                for (unsigned mm = __ballot_sync(FM_FULL, indep && my_ml > ZD_LONG); mm; mm &= mm - 1) {
                    const int l = __ffs(mm) - 1;
                    const int nn = __shfl_sync(FM_FULL, my_ml, l);
                    uint8_t *dp = (uint8_t *)__shfl_sync(FM_FULL, (unsigned long long)m_dst, l);
                    const uint8_t *spp = (const uint8_t *)__shfl_sync(FM_FULL, (unsigned long long)m_src, l);
                    for (int i = lane; i < nn; i += 32) dp[i] = spp[i];
                }
                __syncwarp();
                // C: matches that read this batch's output, in order, by the whole warp
                for (unsigned mm = __ballot_sync(FM_FULL, !indep && my_ml > 0); mm; mm &= mm - 1) {
                    const int l = __ffs(mm) - 1;
                    const long long mlen = __shfl_sync(FM_FULL, my_ml, l);
                    const long long off = __shfl_sync(FM_FULL, my_off, l);
                    uint8_t *o = (uint8_t *)__shfl_sync(FM_FULL, (unsigned long long)m_dst, l);
                    const uint8_t *m = o - off;
                    if (off >= mlen) { for (long long i = lane; i < mlen; i += 32) o[i] = m[i]; }
                    else if (off >= 32) { for (long long i0 = 0; i0 < mlen; i0 += 32) { const long long i = i0 + lane; if (i < mlen) o[i] = m[i]; __syncwarp(); } }
                    else { for (long long i = lane; i < mlen; i += 32) o[i] = m[i % off]; }      // periodic: sources lie before o
                    __syncwarp();
                }
                op += batch_bytes; lit_pos += batch_lits;
            }
            __syncwarp();                                      // the batch arrays are reused
        }
        if (lane == 0) {
            if (b.reload() < 2) err = ERR_CORRUPT;
            w.rep[0] = r0; w.rep[1] = r1; w.rep[2] = r2;
        }
        if (__shfl_sync(FM_FULL, err, 0)) return ERR_CORRUPT;
    }
    {
        const long long rest = regen - lit_pos;
        if (rest > out_cap - op) return ERR_DSTSIZE;
        zd_warp_copy(dst + op, lit + lit_pos, rest, lane);
        op += rest;
    }
    __syncwarp();
    return op;
}

// One zstd frame filling the whole payload.  Returns the decoded size or < 0 ("decide serially").
__device__ long long zd_frame(ZdWarp &w, ZdScratch &sc, const fmz::Tables &T, uint8_t *dst, long long cap,
                              const uint8_t *src, long long n, int lane)
{
    using namespace fmz;
    if (n < 6) return -1;
    const uint32_t magic = src[0] | (src[1] << 8) | (src[2] << 16) | ((uint32_t)src[3] << 24);
    if (magic != 0xFD2FB528u) return -1;
    const int fhd = src[4];
    const int did = fhd & 3, cksum = (fhd >> 2) & 1, single = (fhd >> 5) & 1, fcsid = fhd >> 6;
    if ((fhd & 8) || did || cksum) return -1;
    long long h = 5;
    uint64_t window = 0;
    if (!single) {
        if (h >= n) return -1;
        const int wl = (src[h] >> 3) + 10;
        if (wl > 31) return -1;
        window = (1ull << wl) + ((1ull << wl) >> 3) * (src[h] & 7);
        h++;
    }
    const int fsz = fcsid == 0 ? (single ? 1 : 0) : fcsid == 1 ? 2 : fcsid == 2 ? 4 : 8;
    if (h + fsz > n) return -1;
    uint64_t fcs = 0;
    for (int i = 0; i < fsz; i++) fcs |= (uint64_t)src[h + i] << (8 * i);
    if (fsz == 2) fcs += 256;
    h += fsz;
    if (single) window = fcs;
    const long long block_max = (long long)(window < (uint64_t)BLOCK_MAX ? window : (uint64_t)BLOCK_MAX);
    if (lane == 0) {
        w.huf_ok = w.ll_ok = w.ml_ok = w.of_ok = w.huf_x2 = 0;
        w.rep[0] = 1; w.rep[1] = 4; w.rep[2] = 8;
    }
    __syncwarp();
    long long ip = h, op = 0;
    for (;;) {
        if (n - ip < 3) return -1;
        const uint32_t bh = src[ip] | (src[ip + 1] << 8) | (src[ip + 2] << 16);
        ip += 3;
        const int last = bh & 1, type = (bh >> 1) & 3;
        const long long bsz = bh >> 3;
        if (type == 3) return -1;
        if (type == 1) {
            if (n - ip < 1 || bsz > block_max || bsz > cap - op) return -1;
            const uint8_t v = src[ip];
            for (long long i = lane; i < bsz; i += 32) dst[op + i] = v;
            op += bsz; ip += 1;
        } else {
            if (bsz > n - ip || bsz > block_max) return -1;
            if (type == 2 && bsz >= BLOCK_MAX) return -1;
            if (type == 0) {
                if (bsz > cap - op) return -1;
                zd_warp_copy(dst + op, src + ip, bsz, lane);
                op += bsz;
            } else {
                const long long r = zd_block(w, sc, T, dst, op, cap, src + ip, bsz, lane);
                if (r < 0) return r;
                op = r;
            }
            ip += bsz;
        }
        __syncwarp();
        if (last) break;
    }
    if (ip != n) return -1;                                    // more frames / trailing bytes: the serial decoder judges
    if (fsz && (uint64_t)op != fcs) return -1;
    return op;
}

__global__ void __launch_bounds__(ZD_WARPS * 32)
zstd_frames_warp_kernel(const BlockDesc *blocks, uint32_t n_blocks, ZdScratch *scratch, const fmz::Tables *tables, int32_t *result)
{
    extern __shared__ __align__(16) uint8_t zd_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * ZD_WARPS + warp;
    if (b >= n_blocks) return;
    const BlockDesc bd = blocks[b];
    if (bd.stored) { if (lane == 0) result[b] = (int32_t)bd.usize; return; }
    ZdWarp &w = ((ZdWarp *)zd_smem)[warp];
    ZdScratch &sc = scratch[b];
    if (lane == 0) { w.norm = sc.norm; w.symnext = sc.symnext; w.weights = sc.weights; w.rank = sc.rank; w.wt = sc.wt; }
    __syncwarp();
    const long long r = zd_frame(w, sc, *tables, bd.dst, (long long)bd.usize, bd.src, (long long)bd.csize, lane);
    if (lane == 0) result[b] = r < 0 ? ZD_RETRY : (int32_t)r;
}

// =================================================================================================
// Lane-parallel variant: the FSE sequence streams of up to 32 blocks of a frame are decoded at
// once, one LANE per block (tables and decoded sequences in global scratch), then the warp walks the
// blocks in order: Huffman literals (4 lanes, table in shared memory), repeat-offset resolution
// and batched execution.  Cuts the warp-instructions per sequence of the FSE stage by the number
// of blocks decoded side by side.  Persistent: one warp per CTA pulls frames from a counter and
// owns one scratch area.
// =================================================================================================

constexpr int ZL_SEQ_CAP = 16384;                   // sequences of one block held in scratch; more: ZD_RETRY

struct ZlLane {                                     // per lane, global: the tables a block defines
    fmz::SeqEntry ll[512], ml[512], of[256];
};
struct ZlLaneW { short *norm; uint16_t *symnext; }; // what fmz::build_mode needs of a work area

struct ZlScratch {                                  // per resident warp, global
    ZlLane lane[32];
    fmz::SeqEntry c_ll[512], c_ml[512], c_of[256];  // tables in force at the end of the previous chunk of blocks
    unsigned long long seq[32][ZL_SEQ_CAP];         // literal length | match length << 18 | (offset + 3 or repeat code) << 36
    fmz::Work work;                                 // Huffman build scratch and the literal buffer
};

struct ZlSmem {                                     // per warp (= per CTA), shared; the members fmz::read_huffman uses
    static constexpr int HUF_MAX_LOG = 11;
    fmz::HufEntry huf[2048];
    int huf_log, huf_ok;
    short *norm;
    uint16_t *symnext;
    uint8_t *weights;
    uint32_t *rank;
    fmz::WtEntry *wt;
    uint32_t sub_off[32], sub_size[32];
    uint8_t sub_type[32];
    uint32_t s_off[32];
};

// bytes of the literals section of a compressed block (header + content), or < 0
__device__ __forceinline__ long long zd_lit_section_bytes(const uint8_t *src, long long n)
{
    using namespace fmz;
    if (n < 1) return -1;
    const int ltype = src[0] & 3, sf = (src[0] >> 2) & 3;
    long long hdr, regen, comp;
    if (ltype < 2) {
        if (sf == 0 || sf == 2) { hdr = 1; regen = src[0] >> 3; }
        else if (sf == 1) { if (n < 2) return -1; hdr = 2; regen = (src[0] >> 4) + ((long long)src[1] << 4); }
        else { if (n < 3) return -1; hdr = 3; regen = (src[0] >> 4) + ((long long)src[1] << 4) + ((long long)src[2] << 12); }
        if (regen > BLOCK_MAX) return -1;
        const long long tot = hdr + (ltype == 0 ? regen : 1);
        return tot > n ? -1 : tot;
    }
    if (n < 3) return -1;
    if (sf <= 1) { const uint32_t v = src[0] | (src[1] << 8) | (src[2] << 16); hdr = 3; comp = (v >> 14) & 0x3FF; }
    else if (sf == 2) { if (n < 4) return -1; const uint32_t v = src[0] | (src[1] << 8) | (src[2] << 16) | ((uint32_t)src[3] << 24); hdr = 4; comp = v >> 18; }
    else { if (n < 5) return -1; const uint64_t v = (uint64_t)src[0] | ((uint64_t)src[1] << 8) | ((uint64_t)src[2] << 16) | ((uint64_t)src[3] << 24) | ((uint64_t)src[4] << 32); hdr = 5; comp = (long long)(v >> 22); }
    const long long tot = hdr + comp;
    return tot > n ? -1 : tot;
}

__device__ __forceinline__ fmz::SeqEntry zl_entry(const fmz::SeqEntry *t, uint32_t i)
{
    const uint2 v = *(const uint2 *)(t + i);
    fmz::SeqEntry e;
    e.base = v.x; e.next = (uint16_t)(v.y & 0xffffu); e.nbits = (uint8_t)((v.y >> 16) & 0xffu); e.extra = (uint8_t)(v.y >> 24);
    return e;
}

// One lane decodes one block's sequence stream (zstd_decode.h decode_block, same reads and reloads).
__device__ int zl_decode(const fmz::SeqEntry *tll, const fmz::SeqEntry *tof, const fmz::SeqEntry *tml, int lll, int lof, int lml,
                         const uint8_t *bs, long long bn, long long nseq, unsigned long long *out)
{
    using namespace fmz;
    SeqBits b;
    if (!b.init(bs, bn)) return -1;
    uint32_t sl = b.read(lll); b.reload();
    uint32_t so = b.read(lof); b.reload();
    uint32_t sm = b.read(lml); b.reload();
    for (long long k = 0; k < nseq; k++) {
        const SeqEntry el = zl_entry(tll, sl), eo = zl_entry(tof, so), em = zl_entry(tml, sm);
        unsigned long long offv;
        if (eo.extra > 1) {
            const unsigned long long o = (unsigned long long)eo.base + b.read_fast(eo.extra) + 3ull;
            offv = o < (1ull << 28) ? o : (1ull << 28) - 1;          // beyond any frame here: rejected at execution
        } else {
            const uint32_t ll0 = el.base == 0;
            offv = eo.extra == 0 ? ll0 : eo.base + ll0 + b.read_fast(1);
        }
        unsigned long long mlen = em.base, llen = el.base;
        if (em.extra) mlen += b.read_fast(em.extra);
        if (el.extra + em.extra + eo.extra >= 31) b.reload();
        if (el.extra) llen += b.read_fast(el.extra);
        sl = el.next + b.read(el.nbits);
        sm = em.next + b.read(em.nbits);
        so = eo.next + b.read(eo.nbits);
        out[k] = llen | (mlen << 18) | (offv << 36);
        if (k + 1 < nseq) b.reload();
    }
    return b.reload() < 2 ? -1 : 0;
}

// Literals + execution of one compressed block whose sequences are already decoded.  r0..r2 is the
// repeat-offset history (same value on every lane).
__device__ long long zl_block_exec(ZlSmem &w, ZlScratch &sc, uint8_t *dst, long long op, long long cap, const uint8_t *src, long long n,
                                   long long nseq, const unsigned long long *seq, uint32_t &r0, uint32_t &r1, uint32_t &r2, int lane)
{
    using namespace fmz;
    uint8_t *litbuf = sc.work.lit;
    const int ltype = src[0] & 3, sf = (src[0] >> 2) & 3;          // sizes were validated by zd_lit_section_bytes
    long long hdr, regen, comp = 0;
    const uint8_t *lit = nullptr;
    bool own_lit = true;
    if (ltype < 2) {
        if (sf == 0 || sf == 2) { hdr = 1; regen = src[0] >> 3; }
        else if (sf == 1) { hdr = 2; regen = (src[0] >> 4) + ((long long)src[1] << 4); }
        else { hdr = 3; regen = (src[0] >> 4) + ((long long)src[1] << 4) + ((long long)src[2] << 12); }
        if (ltype == 0) { lit = src + hdr; own_lit = hdr + regen + 32 > n; hdr += regen; }
        else {
            const uint8_t v = src[hdr];
            for (long long i = lane; i < regen; i += 32) litbuf[i] = v;
            lit = litbuf; hdr += 1;
        }
    } else {
        if (n < 5 && sf == 3) return ERR_CORRUPT;
        int streams = 4;
        if (sf <= 1) {
            const uint32_t v = src[0] | (src[1] << 8) | (src[2] << 16);
            hdr = 3; regen = (v >> 4) & 0x3FF; comp = (v >> 14) & 0x3FF;
            streams = sf == 0 ? 1 : 4;
        } else if (sf == 2) {
            const uint32_t v = src[0] | (src[1] << 8) | (src[2] << 16) | ((uint32_t)src[3] << 24);
            hdr = 4; regen = (v >> 4) & 0x3FFF; comp = v >> 18;
        } else {
            const uint64_t v = (uint64_t)src[0] | ((uint64_t)src[1] << 8) | ((uint64_t)src[2] << 16) | ((uint64_t)src[3] << 24) | ((uint64_t)src[4] << 32);
            hdr = 5; regen = (long long)((v >> 4) & 0x3FFFF); comp = (long long)(v >> 22);
        }
        if (regen > BLOCK_MAX) return ERR_CORRUPT;
        const uint8_t *cs = src + hdr;
        long long cn = comp;
        if (ltype == 2) {
            int used = 0;
            if (lane == 0) used = read_huffman(w, cs, cn);
            used = __shfl_sync(FM_FULL, used, 0);
            if (used < 0) return used;
            cs += used; cn -= used;
        } else if (!w.huf_ok) return ERR_CORRUPT;
        __syncwarp();
        const int log = w.huf_log;
        int e = 0;
        if (streams == 1) {
            if (lane == 0) e = zd_huf_stream(w.huf, log, litbuf, (int)regen, cs, cn);
        } else {
            if (cn < 10) return ERR_CORRUPT;
            const long long s1 = cs[0] | (cs[1] << 8), s2 = cs[2] | (cs[3] << 8), s3 = cs[4] | (cs[5] << 8);
            const long long s4 = cn - 6 - s1 - s2 - s3;
            if (s4 < 1 || s1 < 1 || s2 < 1 || s3 < 1) return ERR_CORRUPT;
            const int seg = (int)((regen + 3) / 4);
            if (3LL * seg > regen) return ERR_CORRUPT;
            const uint8_t *q = cs + 6;
            if (lane < 4) {
                const long long so = lane == 0 ? 0 : lane == 1 ? s1 : lane == 2 ? s1 + s2 : s1 + s2 + s3;
                const long long sn = lane == 0 ? s1 : lane == 1 ? s2 : lane == 2 ? s3 : s4;
                const int cnt = lane < 3 ? seg : (int)(regen - 3 * seg);
                e = zd_huf_stream(w.huf, log, litbuf + (size_t)lane * seg, cnt, q + so, sn);
            }
        }
        if (__any_sync(FM_FULL, e < 0)) return ERR_CORRUPT;
        lit = litbuf;
    }
    __syncwarp();
    long long out_cap = cap;
    if (own_lit && cap - op > BLOCK_MAX + 32 + regen + 32) out_cap = op + BLOCK_MAX + 32;
    long long lit_pos = 0;
    unsigned long long rec_next = lane < nseq ? seq[lane] : 0ull;
    for (long long k0 = 0; k0 < nseq; k0 += 32) {
        const int cnt = (int)(nseq - k0 < 32 ? nseq - k0 : 32);
        const unsigned long long rec = rec_next;
        rec_next = k0 + 32 + lane < nseq ? seq[k0 + 32 + lane] : 0ull;      // the next batch's records travel while this one executes
        const int my_ll = (int)(rec & 0x3FFFFu), my_ml = (int)((rec >> 18) & 0x3FFFFu);
        const uint32_t offv = lane < cnt ? (uint32_t)(rec >> 36) : 4u;
        uint32_t my_off;
        if (__ballot_sync(FM_FULL, offv < 4u)) {                   // repeat codes present: resolve in order
            w.s_off[lane] = offv;
            __syncwarp();
            if (lane == 0) {
                for (int j = 0; j < cnt; j++) {
                    const uint32_t v = w.s_off[j];
                    uint32_t o;
                    if (v >= 4u) { o = v - 3u; r2 = r1; r1 = r0; r0 = o; }
                    else if (v == 0u) o = r0;
                    else if (v == 1u) { o = r1; r1 = r0; r0 = o; }
                    else if (v == 2u) { o = r2; r2 = r1; r1 = r0; r0 = o; }
                    else { o = r0 - 1u; o += !o; r2 = r1; r1 = r0; r0 = o; }
                    w.s_off[j] = o;
                }
            }
            __syncwarp();
            my_off = lane < cnt ? w.s_off[lane] : 1u;
            r0 = __shfl_sync(FM_FULL, r0, 0); r1 = __shfl_sync(FM_FULL, r1, 0); r2 = __shfl_sync(FM_FULL, r2, 0);
            __syncwarp();
        } else {
            my_off = offv - 3u;
            const uint32_t a = __shfl_sync(FM_FULL, my_off, cnt - 1);
            const uint32_t b2 = __shfl_sync(FM_FULL, my_off, cnt >= 2 ? cnt - 2 : 0);
            const uint32_t c2 = __shfl_sync(FM_FULL, my_off, cnt >= 3 ? cnt - 3 : 0);
            const uint32_t n2 = cnt >= 3 ? c2 : cnt == 2 ? r0 : r1;
            const uint32_t n1 = cnt >= 2 ? b2 : r0;
            r0 = a; r1 = n1; r2 = n2;
        }
        // positions and the serial decoder's per-sequence checks
        const int incl = warp_incl_scan_add(my_ll + my_ml), incl_ll = warp_incl_scan_add(my_ll);
        const long long my_op = op + (incl - my_ll - my_ml);
        int bad = 0;
        if (lane < cnt) {
            if ((long long)my_ll + my_ml > out_cap - my_op) bad = 1;
            if ((long long)incl_ll > regen - lit_pos) bad = 1;
            if ((long long)my_off > my_op + my_ll) bad = 1;
        }
        if (__any_sync(FM_FULL, bad)) return ERR_CORRUPT;
        uint8_t *my_dst = dst + my_op;
        const uint8_t *my_lit = lit + lit_pos + (incl_ll - my_ll);
        const int batch_bytes = __shfl_sync(FM_FULL, incl, 31), batch_lits = __shfl_sync(FM_FULL, incl_ll, 31);
        if (my_ll <= ZD_LONG) copy_batched(my_dst, my_lit, my_ll);
        for (unsigned mm = __ballot_sync(FM_FULL, my_ll > ZD_LONG); mm; mm &= mm - 1) {
            const int l = __ffs(mm) - 1;
            const int nn = __shfl_sync(FM_FULL, my_ll, l);
            uint8_t *dp = (uint8_t *)__shfl_sync(FM_FULL, (unsigned long long)my_dst, l);
            const uint8_t *spp = (const uint8_t *)__shfl_sync(FM_FULL, (unsigned long long)my_lit, l);
            for (int i = lane; i < nn; i += 32) dp[i] = spp[i];
        }
        uint8_t *m_dst = my_dst + my_ll;
        const uint8_t *m_src = m_dst - my_off;
        const bool indep = m_src + my_ml <= dst + op;
        if (indep && my_ml <= ZD_LONG) copy_batched(m_dst, m_src, my_ml);
        for (unsigned mm = __ballot_sync(FM_FULL, indep && my_ml > ZD_LONG); mm; mm &= mm - 1) {
            const int l = __ffs(mm) - 1;
            const int nn = __shfl_sync(FM_FULL, my_ml, l);
            uint8_t *dp = (uint8_t *)__shfl_sync(FM_FULL, (unsigned long long)m_dst, l);
            const uint8_t *spp = (const uint8_t *)__shfl_sync(FM_FULL, (unsigned long long)m_src, l);
            for (int i = lane; i < nn; i += 32) dp[i] = spp[i];
        }
        __syncwarp();
        for (unsigned mm = __ballot_sync(FM_FULL, !indep && my_ml > 0); mm; mm &= mm - 1) {
            const int l = __ffs(mm) - 1;
            const long long mlen = __shfl_sync(FM_FULL, my_ml, l);
            const long long off = __shfl_sync(FM_FULL, (long long)my_off, l);
            uint8_t *o = (uint8_t *)__shfl_sync(FM_FULL, (unsigned long long)m_dst, l);
            const uint8_t *m = o - off;
            if (off >= mlen) { for (long long i = lane; i < mlen; i += 32) o[i] = m[i]; }
            else if (off >= 32) { for (long long i0 = 0; i0 < mlen; i0 += 32) { const long long i = i0 + lane; if (i < mlen) o[i] = m[i]; __syncwarp(); } }
            else { for (long long i = lane; i < mlen; i += 32) o[i] = m[i % off]; }
            __syncwarp();
        }
        op += batch_bytes; lit_pos += batch_lits;
    }
    {
        const long long rest = regen - lit_pos;
        if (rest > out_cap - op) return ERR_DSTSIZE;
        zd_warp_copy(dst + op, lit + lit_pos, rest, lane);
        op += rest;
    }
    __syncwarp();
    return op;
}

__device__ long long zl_frame(ZlSmem &w, ZlScratch &sc, const fmz::Tables &T, uint8_t *dst, long long cap,
                              const uint8_t *src, long long n, int lane)
{
    using namespace fmz;
    if (n < 6) return -1;
    const uint32_t magic = src[0] | (src[1] << 8) | (src[2] << 16) | ((uint32_t)src[3] << 24);
    if (magic != 0xFD2FB528u) return -1;
    const int fhd = src[4];
    const int did = fhd & 3, cksum = (fhd >> 2) & 1, single = (fhd >> 5) & 1, fcsid = fhd >> 6;
    if ((fhd & 8) || did || cksum) return -1;
    long long h = 5;
    uint64_t window = 0;
    if (!single) {
        if (h >= n) return -1;
        const int wl = (src[h] >> 3) + 10;
        if (wl > 31) return -1;
        window = (1ull << wl) + ((1ull << wl) >> 3) * (src[h] & 7);
        h++;
    }
    const int fsz = fcsid == 0 ? (single ? 1 : 0) : fcsid == 1 ? 2 : fcsid == 2 ? 4 : 8;
    if (h + fsz > n) return -1;
    uint64_t fcs = 0;
    for (int i = 0; i < fsz; i++) fcs |= (uint64_t)src[h + i] << (8 * i);
    if (fsz == 2) fcs += 256;
    h += fsz;
    if (single) window = fcs;
    const long long block_max = (long long)(window < (uint64_t)BLOCK_MAX ? window : (uint64_t)BLOCK_MAX);
    if (lane == 0) w.huf_ok = 0;
    uint32_t r0 = 1, r1 = 4, r2 = 8;
    int c_ok[3] = {0, 0, 0}, c_log[3] = {0, 0, 0};                // carried tables: LL, OF, ML
    long long ip = h, op = 0;
    for (bool done = false; !done;) {
        // ---- lane 0 walks the next (up to 32) block headers
        int count = 0, werr = 0, lastf = 0;
        long long ipn = ip;
        if (lane == 0) {
            while (count < 32) {
                if (n - ipn < 3) { werr = 1; break; }
                const uint32_t bh = src[ipn] | (src[ipn + 1] << 8) | (src[ipn + 2] << 16);
                ipn += 3;
                const int last = bh & 1, type = (bh >> 1) & 3;
                const long long bsz = bh >> 3;
                if (type == 3 || bsz > block_max) { werr = 1; break; }
                w.sub_off[count] = (uint32_t)ipn; w.sub_size[count] = (uint32_t)bsz; w.sub_type[count] = (uint8_t)type;
                if (type == 1) { if (n - ipn < 1) { werr = 1; break; } ipn += 1; }
                else { if (bsz > n - ipn || (type == 2 && bsz >= BLOCK_MAX)) { werr = 1; break; } ipn += bsz; }
                count++;
                if (last) { lastf = 1; break; }
            }
        }
        count = __shfl_sync(FM_FULL, count, 0); werr = __shfl_sync(FM_FULL, werr, 0); lastf = __shfl_sync(FM_FULL, lastf, 0);
        ipn = __shfl_sync(FM_FULL, ipn, 0);
        if (werr) return -1;
        __syncwarp();
        // ---- every lane locates its block's sequences section
        int st = 0, modes = 0;
        long long my_nseq = 0, sn = 0, shdr = 0;
        const uint8_t *sp = nullptr;
        const bool mine = lane < count && w.sub_type[lane] == 2;
        if (mine) {
            const uint8_t *bsrc = src + w.sub_off[lane];
            const long long bn = w.sub_size[lane];
            const long long ls = zd_lit_section_bytes(bsrc, bn);
            if (ls < 0) st = -1;
            else {
                sp = bsrc + ls; sn = bn - ls;
                if (sn < 1) st = -1;
                else {
                    my_nseq = sp[0]; shdr = 1;
                    if (my_nseq == 0 && sn != 1) st = -1;
                    if (my_nseq >= 128) {
                        if (my_nseq == 255) { if (sn < 3) st = -1; else { my_nseq = sp[1] + (sp[2] << 8) + 0x7F00; shdr = 3; } }
                        else { if (sn < 2) st = -1; else { my_nseq = ((my_nseq - 128) << 8) + sp[1]; shdr = 2; } }
                    }
                    if (st == 0 && my_nseq > 0) { if (shdr + 1 > sn) st = -1; else { modes = sp[shdr]; shdr += 1; } }
                    if (my_nseq > ZL_SEQ_CAP) st = -1;
                }
            }
        }
        if (__any_sync(FM_FULL, st < 0)) return -1;
        const bool has = mine && my_nseq > 0;
        // ---- who defines which table; repeat mode takes the nearest earlier definition
        int owner[3], mode[3], my_log[3] = {0, 0, 0};
        unsigned defmask[3];
#pragma unroll
        for (int x = 0; x < 3; x++) {
            mode[x] = (modes >> (6 - 2 * x)) & 3;
            defmask[x] = __ballot_sync(FM_FULL, has && mode[x] != 3);
            const unsigned prior = defmask[x] & ((1u << lane) - 1u);
            owner[x] = prior ? 31 - __clz(prior) : -1;
            if (has && mode[x] == 3 && owner[x] < 0 && !c_ok[x]) st = -1;
        }
        if (__any_sync(FM_FULL, st < 0)) return -1;
        ZlLane &ml = sc.lane[lane];
        if (has) {
            // thread-local scratch: local memory interleaves the lanes, so lanes building in lockstep coalesce
            short l_norm[64];
            uint16_t l_symnext[64];
            ZlLaneW lw{l_norm, l_symnext};
            long long hh = shdr;
            int ok = 0, used;
            if (mode[0] != 3) { used = build_mode(mode[0], ml.ll, &my_log[0], &ok, lw, T, 0, sp + hh, sn - hh); if (used < 0) st = -1; else hh += used; }
            if (st == 0 && mode[1] != 3) { used = build_mode(mode[1], ml.of, &my_log[1], &ok, lw, T, 1, sp + hh, sn - hh); if (used < 0) st = -1; else hh += used; }
            if (st == 0 && mode[2] != 3) { used = build_mode(mode[2], ml.ml, &my_log[2], &ok, lw, T, 2, sp + hh, sn - hh); if (used < 0) st = -1; else hh += used; }
            shdr = hh;
        }
        if (__any_sync(FM_FULL, st < 0)) return -1;
        __syncwarp();                                              // tables written by other lanes are visible
        const SeqEntry *tp[3];
        int lg[3];
#pragma unroll
        for (int x = 0; x < 3; x++) {
            const int o_log = __shfl_sync(FM_FULL, my_log[x], owner[x] < 0 ? 0 : owner[x]);
            const ZlLane &ol = sc.lane[owner[x] < 0 ? 0 : owner[x]];
            if (mode[x] != 3) { tp[x] = x == 0 ? ml.ll : x == 1 ? ml.of : ml.ml; lg[x] = my_log[x]; }
            else if (owner[x] >= 0) { tp[x] = x == 0 ? ol.ll : x == 1 ? ol.of : ol.ml; lg[x] = o_log; }
            else { tp[x] = x == 0 ? sc.c_ll : x == 1 ? sc.c_of : sc.c_ml; lg[x] = c_log[x]; }
        }
        if (has) st = zl_decode(tp[0], tp[1], tp[2], lg[0], lg[1], lg[2], sp + shdr, sn - shdr, my_nseq, sc.seq[lane]);
        if (__any_sync(FM_FULL, st < 0)) return -1;
        __syncwarp();
        // ---- the blocks in order
        for (int j = 0; j < count; j++) {
            const int type = w.sub_type[j];
            const uint8_t *bsrc = src + w.sub_off[j];
            const long long bsz = w.sub_size[j];
            const long long nseq_j = __shfl_sync(FM_FULL, my_nseq, j);
            if (type == 1) {
                if (bsz > cap - op) return -1;
                const uint8_t v = bsrc[0];
                for (long long i = lane; i < bsz; i += 32) dst[op + i] = v;
                op += bsz;
            } else if (type == 0) {
                if (bsz > cap - op) return -1;
                zd_warp_copy(dst + op, bsrc, bsz, lane);
                op += bsz;
            } else {
                const long long r = zl_block_exec(w, sc, dst, op, cap, bsrc, bsz, nseq_j, sc.seq[j], r0, r1, r2, lane);
                if (r < 0) return r;
                op = r;
            }
            __syncwarp();
        }
        // ---- carry the tables in force into the next chunk of blocks
        if (!lastf) {
#pragma unroll
            for (int x = 0; x < 3; x++) {
                if (!defmask[x]) continue;
                const int l = 31 - __clz(defmask[x]);
                const ZlLane &ol = sc.lane[l];
                const SeqEntry *from = x == 0 ? ol.ll : x == 1 ? ol.of : ol.ml;
                SeqEntry *to = x == 0 ? sc.c_ll : x == 1 ? sc.c_of : sc.c_ml;
                const int cnt = x == 1 ? 256 : 512;
                for (int i = lane; i < cnt; i += 32) to[i] = from[i];
                c_log[x] = __shfl_sync(FM_FULL, my_log[x], l);
                c_ok[x] = 1;
            }
            __syncwarp();
        }
        ip = ipn;
        done = lastf != 0;
    }
    if (ip != n) return -1;
    if (fsz && (uint64_t)op != fcs) return -1;
    return op;
}

__global__ void __launch_bounds__(32)
zstd_frames_lane_kernel(const BlockDesc *blocks, uint32_t n_blocks, ZlScratch *scratch, const fmz::Tables *tables,
                        int32_t *result, uint32_t *counter)
{
    __shared__ ZlSmem w;
    const int lane = threadIdx.x;
    ZlScratch &sc = scratch[blockIdx.x];
    if (lane == 0) { w.norm = sc.work.norm; w.symnext = sc.work.symnext; w.weights = sc.work.weights; w.rank = sc.work.rank; w.wt = sc.work.wt; }
    __syncwarp();
    for (;;) {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(counter, 1u);
        b = __shfl_sync(FM_FULL, b, 0);
        if (b >= n_blocks) break;
        const BlockDesc bd = blocks[b];
        if (bd.stored) { if (lane == 0) result[b] = (int32_t)bd.usize; continue; }
        const long long r = zl_frame(w, sc, *tables, bd.dst, (long long)bd.usize, bd.src, (long long)bd.csize, lane);
        if (lane == 0) result[b] = r < 0 ? ZD_RETRY : (int32_t)r;
        __syncwarp();
    }
}

// The exact serial decoder, for the frames the warp kernel did not settle.
__global__ void __launch_bounds__(32)
zstd_frames_kernel(const BlockDesc *blocks, uint32_t n_blocks, fmz::Work *work, const fmz::Tables *tables, int32_t *result, int only_retry)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const BlockDesc bd = blocks[b];
    if (only_retry && result[b] != ZD_RETRY) return;
    if (bd.stored) { result[b] = (int32_t)bd.usize; return; }
    const long long r = fmz::decompress(bd.dst, (long long)bd.usize, bd.src, (long long)bd.csize, work[b], *tables);
    result[b] = (int32_t)r;
}

}  // namespace fm
