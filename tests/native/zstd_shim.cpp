// zstd_shim.cpp -- TEST-ONLY host build of the product's zstd frame decoder
// (4mc_b200/csrc/zstd_decode.h) so that it is checked against the reference's ZSTD_decompress and
// the committed golden .4mz files on a machine without a GPU.
#include <cstdint>
#include <cstdlib>
#include "../../4mc_b200/csrc/zstd_decode.h"

extern "C" long long zstd_shim_decompress(uint8_t *dst, long long cap, const uint8_t *src, long long n)
{
    static fmz::Tables T;
    static bool init = false;
    if (!init) { fmz::make_tables(T); init = true; }
    fmz::Work *w = (fmz::Work *)malloc(sizeof(fmz::Work));
    const long long r = fmz::decompress(dst, cap, src, n, *w, T);
    free(w);
    return r;
}
