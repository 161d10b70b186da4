// fm_common.cuh -- shared device helpers for lib4mcgpu (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fourmc.h"

#define FM_WARP 32
#define FM_FULL 0xffffffffu

namespace fm {

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// big-endian u32 store/load at arbitrary alignment (container fields, native/4mc.c:122-123)
__device__ __forceinline__ void st_be32(uint8_t *p, uint32_t v)
{
    p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}
__device__ __forceinline__ uint32_t ld_be32(const uint8_t *p)
{
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

__device__ __forceinline__ uint4 ldg_nc_v4(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ uint4 ldg_v4(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}

__device__ __forceinline__ int warp_incl_scan_add(int v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FM_FULL, v, d);
        if (lane_id() >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ int warp_incl_scan_max(int v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FM_FULL, v, d);
        if (lane_id() >= d) v = max(v, t);
    }
    return v;
}

__device__ __forceinline__ int warp_min(int v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(FM_FULL, v, d));
    return v;
}

__device__ __forceinline__ int warp_max(int v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(FM_FULL, v, d));
    return v;
}

// Byte copy of a short, non-overlapping range with every load issued before the first store, so a
// lane pays one memory latency per 8 bytes instead of one per byte (warps issue in order: a store
// that waits for its load's data blocks the loads behind it).
template <class D, class S>
__device__ __forceinline__ void copy_batched(D *dst, const S *src, int n)
{
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        uint8_t t[8];
#pragma unroll
        for (int j = 0; j < 8; j++) t[j] = src[i + j];
#pragma unroll
        for (int j = 0; j < 8; j++) dst[i + j] = t[j];
    }
    if (i < n) {
        uint8_t t[8];
#pragma unroll
        for (int j = 0; j < 7; j++) if (i + j < n) t[j] = src[i + j];
#pragma unroll
        for (int j = 0; j < 7; j++) if (i + j < n) dst[i + j] = t[j];
    }
}

}  // namespace fm
