// parse_shim.cpp -- TEST-ONLY host build of the product's D1 parser (4mc_b200/csrc/lz4_parse.h),
// so that its accept/reject logic is checked against the oracle on a machine without a GPU.
#include <cstdint>
#include <vector>
#include "../../4mc_b200/csrc/lz4_parse.h"

namespace {
struct Rec { std::vector<int> pos, op; void token(int p, int o) { pos.push_back(p); op.push_back(o); } };
}

extern "C" int parse_shim(const uint8_t *src, int n, int cap, int *n_tokens, long long *pos_sum, long long *op_sum, int step)
{
    Rec r;
    const int ret = fm::lz4_parse_block(src, n, cap, r, step > 0 ? step : 0x7fffffff);
    long long ps = 0, os = 0;
    for (size_t i = 0; i < r.pos.size(); i++) { ps += r.pos[i]; os += r.op[i]; }
    if (n_tokens) *n_tokens = (int)r.pos.size();
    if (pos_sum) *pos_sum = ps;
    if (op_sum) *op_sum = os;
    return ret;
}
