// xxh32.cuh -- XXH32 on the device (reference: native/lz4/xxhash.c:263-416).
//
// XXH32 is four serial lane recurrences over 16-byte stripes; the recurrence
//     v = rotl13(v + w * P2) * P1                        (xxhash.c:269-275)
// is not associative, so one buffer cannot be split.  Parallelism comes from hashing many
// payloads at once: one WARP per payload.  All 32 lanes stream the payload from HBM with
// coalesced 128-bit loads into a double-buffered shared-memory stage (2 KiB per stage, the next
// stage's loads are in flight while the current one is consumed); lanes 0..3 each own one of the
// four accumulators and walk the stage.  The dependent chain (IMAD, SHF, IMUL) is ~13 cycles
// per 16 bytes per payload, so a 2 MiB payload costs ~1 ms regardless of how many are in flight.
#pragma once

#include "fm_common.cuh"

namespace fm {

constexpr uint32_t XP1 = 0x9E3779B1u;
constexpr uint32_t XP2 = 0x85EBCA77u;
constexpr uint32_t XP3 = 0xC2B2AE3Du;
constexpr uint32_t XP4 = 0x27D4EB2Fu;
constexpr uint32_t XP5 = 0x165667B1u;

constexpr int XXH_STAGE_STRIPES = 128;                        // 2 KiB of payload per stage
constexpr int XXH_STAGE_WORDS = (XXH_STAGE_STRIPES + 1) * 4;  // +1 chunk: stripes may straddle
constexpr int XXH_WARP_SMEM_WORDS = 2 * XXH_STAGE_WORDS;      // double buffer

__device__ __forceinline__ uint32_t xxh_round(uint32_t acc, uint32_t w)
{
    return rotl32(acc + w * XP2, 13) * XP1;
}

__device__ __forceinline__ uint32_t xxh_avalanche(uint32_t h)
{
    h ^= h >> 15; h *= XP2;
    h ^= h >> 13; h *= XP3;
    h ^= h >> 16;
    return h;
}

// Hash of a short range by one thread (tails, 8-byte headers): plain byte loads.
__device__ inline uint32_t xxh32_thread(const uint8_t *p, uint32_t len, uint32_t seed)
{
    const uint8_t *end = p + len;
    uint32_t h;
    auto rd = [](const uint8_t *q) {
        return (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
    };
    if (len >= 16) {
        uint32_t v1 = seed + XP1 + XP2, v2 = seed + XP2, v3 = seed, v4 = seed - XP1;
        const uint8_t *limit = end - 15;
        do {
            v1 = xxh_round(v1, rd(p)); v2 = xxh_round(v2, rd(p + 4));
            v3 = xxh_round(v3, rd(p + 8)); v4 = xxh_round(v4, rd(p + 12));
            p += 16;
        } while (p < limit);
        h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    } else {
        h = seed + XP5;
    }
    h += len;
    while (p + 4 <= end) { h = rotl32(h + rd(p) * XP3, 17) * XP4; p += 4; }
    while (p < end) { h = rotl32(h + (uint32_t)(*p) * XP5, 11) * XP1; p++; }
    return xxh_avalanche(h);
}

// Warp-cooperative XXH32 of [p, p+len).  All 32 lanes call with identical arguments;
// `stage` is this warp's XXH_WARP_SMEM_WORDS words of shared memory.  Result in every lane.
// NC = true reads through the non-coherent path (payload written by an EARLIER kernel);
// NC = false uses ordinary loads (payload written by this CTA before a __syncthreads()).
template <bool NC>
__device__ inline uint32_t xxh32_warp(const uint8_t *p, uint32_t len, uint32_t seed, uint32_t *stage)
{
    const int lane = lane_id();
    const uintptr_t a = (uintptr_t)p;
    const uint4 *base = (const uint4 *)(a & ~(uintptr_t)15);
    const uint32_t d = (uint32_t)(a & 15);
    const uint32_t nstripes = len >> 4;
    const uint32_t nchunks = (d + len + 15) >> 4;   // aligned 16-byte chunks holding any payload byte
    const uint32_t shift = (d & 3) * 8;

    uint32_t acc = 0;
    if (lane == 0) acc = seed + XP1 + XP2;
    else if (lane == 1) acc = seed + XP2;
    else if (lane == 2) acc = seed;
    else if (lane == 3) acc = seed - XP1;

    const uint32_t nstages = (nstripes + XXH_STAGE_STRIPES - 1) / XXH_STAGE_STRIPES;
    uint4 r0, r1, r2, r3, rx;
    auto ld = [](const uint4 *q) { return NC ? ldg_nc_v4(q) : ldg_v4(q); };
    auto load_stage = [&](uint32_t t) {
        const uint32_t c0 = t * XXH_STAGE_STRIPES + lane;
        const uint4 z = make_uint4(0, 0, 0, 0);
        r0 = (c0 < nchunks) ? ld(base + c0) : z;
        r1 = (c0 + 32 < nchunks) ? ld(base + c0 + 32) : z;
        r2 = (c0 + 64 < nchunks) ? ld(base + c0 + 64) : z;
        r3 = (c0 + 96 < nchunks) ? ld(base + c0 + 96) : z;
        rx = (lane == 0 && c0 + 128 < nchunks) ? ld(base + c0 + 128) : z;
    };
    auto store_stage = [&](uint32_t *buf) {
        uint4 *b = (uint4 *)buf;
        b[lane] = r0; b[lane + 32] = r1; b[lane + 64] = r2; b[lane + 96] = r3;
        if (lane == 0) b[128] = rx;
    };

    if (nstages > 0) {
        load_stage(0);
        store_stage(stage);
        __syncwarp();
    }
    for (uint32_t t = 0; t < nstages; t++) {
        uint32_t *cur = stage + (t & 1) * XXH_STAGE_WORDS;
        if (t + 1 < nstages) load_stage(t + 1);              // in flight during the walk below
        const uint32_t ns = min((uint32_t)XXH_STAGE_STRIPES, nstripes - t * XXH_STAGE_STRIPES);
        if (lane < 4) {
            const uint32_t *w = cur + ((d >> 2) + lane);
            uint32_t s = 0;
            for (; s + 4 <= ns; s += 4) {
                uint32_t a0 = __funnelshift_r(w[0], w[1], shift);
                uint32_t a1 = __funnelshift_r(w[4], w[5], shift);
                uint32_t a2 = __funnelshift_r(w[8], w[9], shift);
                uint32_t a3 = __funnelshift_r(w[12], w[13], shift);
                acc = xxh_round(acc, a0); acc = xxh_round(acc, a1);
                acc = xxh_round(acc, a2); acc = xxh_round(acc, a3);
                w += 16;
            }
            for (; s < ns; s++) { acc = xxh_round(acc, __funnelshift_r(w[0], w[1], shift)); w += 4; }
        }
        __syncwarp();
        if (t + 1 < nstages) {
            store_stage(stage + ((t + 1) & 1) * XXH_STAGE_WORDS);
            __syncwarp();
        }
    }

    uint32_t v1 = __shfl_sync(FM_FULL, acc, 0), v2 = __shfl_sync(FM_FULL, acc, 1);
    uint32_t v3 = __shfl_sync(FM_FULL, acc, 2), v4 = __shfl_sync(FM_FULL, acc, 3);
    uint32_t h = (len >= 16) ? rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18) : seed + XP5;
    h += len;
    if (lane == 0) {
        const uint8_t *q = p + ((size_t)nstripes << 4);
        const uint8_t *end = p + len;
        while (q + 4 <= end) {
            uint32_t w = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
            h = rotl32(h + w * XP3, 17) * XP4;
            q += 4;
        }
        while (q < end) { h = rotl32(h + (uint32_t)(*q) * XP5, 11) * XP1; q++; }
        h = xxh_avalanche(h);
    }
    return __shfl_sync(FM_FULL, h, 0);
}

}  // namespace fm
