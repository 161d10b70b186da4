/*
 * gen_host.c -- libfourmcgen.so: the synthetic inputs of SURVEY.md 8d on the host, for callers that must not map
 * the CUDA library (bench.py --impl reference: the reference's CPU arm runs none of this repo's product code).
 * Same source as the device generator (4mc_b200/csrc/fourmc_gen.h): a page is a pure function of
 * (kind, seed, page index).
 */
#include "../csrc/fourmc_gen.h"

__attribute__((visibility("default")))
int fourmcgen_pages(int kind, uint64_t seed, uint64_t first_page, uint64_t n_pages, void *out)
{
    if (kind < 0 || kind > 2 || !out) return -1;
    for (uint64_t p = 0; p < n_pages; p++) fmg_page(kind, seed, first_page + p, (uint8_t *)out + p * FMG_PAGE);
    return 0;
}
