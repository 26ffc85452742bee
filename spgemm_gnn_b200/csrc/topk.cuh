// Device helpers shared by the MaxK top-k kernels (topk.cu, topk_tile.cu).
#pragma once

#include "common.cuh"

namespace mk {

// Order-preserving uint32 key of a float: larger value <=> larger key, every NaN above +inf (the
// torch.topk convention), -0.0 ties with +0.0.  Three instructions: `v + 0.0f` turns -0.0 into +0.0
// and every NaN into the canonical 0x7FFFFFFF (an fp32 add never returns another NaN pattern), then
// positive values get their sign bit set and negative ones are complemented.
__device__ __forceinline__ uint32_t order_key(float v) {
    const int32_t b = __float_as_int(__fadd_rn(v, 0.0f));
    return static_cast<uint32_t>(b) ^ (static_cast<uint32_t>(b >> 31) | 0x80000000u);
}

// c += (key >= cand), cand != 0, given ncand = -cand: the carry out of key + (2^32 - cand).  Two integer
// instructions per element, and ptxas folds two carries into one IADD3.X (the compiler's own
// `c += key >= cand` is a compare, an add and a predicated move per element).
__device__ __forceinline__ void count_ge(int& c, uint32_t key, uint32_t ncand) {
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %0, 0;\n\t}" : "+r"(c) : "r"(key), "r"(ncand));
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

}  // namespace mk
