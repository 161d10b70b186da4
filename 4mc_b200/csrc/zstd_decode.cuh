// zstd_decode.cuh -- 4mz block decode kernel: one THREAD per zstd frame (= per 4mz block), the
// frame decoder of zstd_decode.h run as is.  All blocks of a batch decode concurrently; inside a
// frame everything is serial (Huffman literals, FSE sequences, execution), so throughput comes
// from the number of blocks in flight.  First correct version of SURVEY.md row a9; splitting the
// entropy stages across a warp is the next step.
#pragma once

#include "fm_common.cuh"
#include "lz4_decode.cuh"
#include "zstd_decode.h"

namespace fm {

__global__ void __launch_bounds__(32)
zstd_frames_kernel(const BlockDesc *blocks, uint32_t n_blocks, fmz::Work *work, const fmz::Tables *tables, int32_t *result)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const BlockDesc bd = blocks[b];
    if (bd.stored) { result[b] = (int32_t)bd.usize; return; }
    const long long r = fmz::decompress(bd.dst, (long long)bd.usize, bd.src, (long long)bd.csize, work[b], *tables);
    result[b] = (int32_t)r;
}

}  // namespace fm
