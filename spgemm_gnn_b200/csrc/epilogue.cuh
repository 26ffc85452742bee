// f-3  The epilogue of the aggregation, shared by the stand-alone kernel (layernorm.cu) and the forward
// SpGEMM that applies it to the row it has just finished (banked.cu, spgemm_fwd.cu fold):
//     z = a + b (+ bias),  y = LayerNorm(z) * gamma + beta          (utils/maxk_layers.py:174-182)
// One warp per row; lane holds columns j*128 + lane*4 .. +3 of the row as float4, NV4 of them
// (D <= 128*NV4, D % 4 == 0).  Every user runs this exact sequence of operations, so that the fused
// forward reproduces the stand-alone kernel bit for bit.
#pragma once

#include "common.cuh"

namespace mk {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// What the forward SpGEMM needs to finish a row as `y = LayerNorm(h_self + row + bias)`.
// gamma == nullptr: no epilogue.  z / mean / rstd may be null (inference: nothing kept for a backward).
struct FwdEpilogue {
    const float* h_self;  // [n_rows, d] or null
    const float* bias;    // [d] or null
    const float* gamma;   // [d]
    const float* beta;    // [d]
    float* z;             // [n_rows, d] pre-normalisation sum, or null
    float* mean;          // [n_rows] or null
    float* rstd;          // [n_rows] or null
    float eps;
};

// `first(c)` / `second(c)`: float4 of the two addends at column c (second may be absent: HAS_B false).
template <int NV4, bool HAS_B, typename FA, typename FB>
__device__ __forceinline__ void add_layernorm_row(FA first, FB second, const float* __restrict__ bias,
                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                  float* __restrict__ z_row, float* __restrict__ y_row,
                                                  float* __restrict__ mean_at, float* __restrict__ rstd_at, int d,
                                                  float eps, int lane) {
    float v[NV4 * 4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV4; ++j) {
        const int c = j * 128 + lane * 4;
        if (c < d) {
            float4 x = first(c);
            if (HAS_B) {
                const float4 r = second(c);
                x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w;
            }
            if (bias) {
                const float4 r = *reinterpret_cast<const float4*>(bias + c);
                x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w;
            }
            v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
            s += (x.x + x.y) + (x.z + x.w);
            if (z_row) *reinterpret_cast<float4*>(z_row + c) = x;
        } else {
            v[4 * j] = v[4 * j + 1] = v[4 * j + 2] = v[4 * j + 3] = 0.f;
        }
    }
    const float mu = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV4; ++j) {
        const int c = j * 128 + lane * 4;
        if (c < d) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float t = v[4 * j + i] - mu;
                q += t * t;
            }
        }
    }
    const float rs = rsqrtf(warp_sum(q) / d + eps);
    if (lane == 0) {
        if (mean_at) *mean_at = mu;
        if (rstd_at) *rstd_at = rs;
    }
#pragma unroll
    for (int j = 0; j < NV4; ++j) {
        const int c = j * 128 + lane * 4;
        if (c < d) {
            const float4 g = *reinterpret_cast<const float4*>(gamma + c);
            const float4 bt = *reinterpret_cast<const float4*>(beta + c);
            float4 o;
            o.x = (v[4 * j] - mu) * rs * g.x + bt.x;
            o.y = (v[4 * j + 1] - mu) * rs * g.y + bt.y;
            o.z = (v[4 * j + 2] - mu) * rs * g.z + bt.z;
            o.w = (v[4 * j + 3] - mu) * rs * g.w + bt.w;
            st_stream_f4(y_row + c, o);
        }
    }
}

// The forward's use: `agg(c)` is the aggregated row (shared memory or folded partials), D <= 512.
// `hself(c)`: this lane's float4 of h_self at column c (prefetched registers, or a load).
template <typename FAgg, typename FSelf>
__device__ __forceinline__ void fwd_epilogue_row(const FwdEpilogue& ep, FAgg agg, FSelf hself,
                                                 float* __restrict__ y, int64_t row, int d, int lane) {
    float* z_row = ep.z ? ep.z + row * d : nullptr;
    float* y_row = y + row * d;
    float* m = ep.mean ? ep.mean + row : nullptr;
    float* r = ep.rstd ? ep.rstd + row : nullptr;
    if (ep.h_self != nullptr) {
        add_layernorm_row<4, true>(hself, agg, ep.bias, ep.gamma, ep.beta, z_row, y_row, m, r, d, ep.eps, lane);
    } else {
        add_layernorm_row<4, false>(agg, agg, ep.bias, ep.gamma, ep.beta, z_row, y_row, m, r, d, ep.eps, lane);
    }
}

}  // namespace mk
