// f-3  Epilogue of the aggregation: z = a + b (+ bias), y = LayerNorm(z) * gamma + beta.
//
// In the reference this is three passes after the SpGEMM: `h_self + aggregated`
// (utils/maxk_layers.py:174), `self.norm(output)` (:181-182, nn.LayerNorm) and their autograd twins.
// Around the sparse kernels these dense N x D passes are the next-largest traffic of an epoch
// (SURVEY.md section 8 f-3): torch's LayerNorm kernels take 0.40 ms forward and 0.95 ms backward
// per call at N = 232,965, D = 256 (13 % of a MaxK-SAGE epoch) where the bytes moved would allow
// ~0.15 ms.  One warp per row, the row in registers as float4 (D % 4 == 0, D <= 1024):
//   forward : reads a, b (and bias), writes z (kept for the backward), y, mean, rstd;
//   backward: reads gy, z, writes gz (= d loss / d z, which is also d/da and d/db), and per-CTA
//             partial sums of d gamma / d beta that a second tiny kernel folds in fixed order.
#include "common.cuh"
#include "epilogue.cuh"

namespace mk {

constexpr int kLNMaxV4 = 8;  // float4 per lane -> D <= 1024

template <int NV4>
__global__ void __launch_bounds__(256)
add_layernorm_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                         const float* __restrict__ bias, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float* __restrict__ z,
                         float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd,
                         int64_t n, int d, float eps) {
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const int lane = lane_id();
    const float* __restrict__ ar = a + row * d;
    float* z_row = z ? z + row * d : nullptr;
    auto fa = [&](int c) { return ld_stream_f4(ar + c); };
    if (b) {
        const float* __restrict__ br = b + row * d;
        add_layernorm_row<NV4, true>(fa, [&](int c) { return ld_stream_f4(br + c); }, bias, gamma, beta, z_row,
                                     y + row * d, mean + row, rstd + row, d, eps, lane);
    } else {
        add_layernorm_row<NV4, false>(fa, fa, bias, gamma, beta, z_row, y + row * d, mean + row, rstd + row, d,
                                      eps, lane);
    }
}

// Persistent grid: every warp walks rows with a grid stride and keeps its d gamma / d beta
// contributions in registers; the 8 warps of a CTA are folded through shared memory and the CTA
// writes one partial row pair.
template <int NV4>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ z,
                     const float* __restrict__ gamma, const float* __restrict__ mean,
                     const float* __restrict__ rstd, float* __restrict__ gz,
                     float* __restrict__ part_dgamma, float* __restrict__ part_dbeta,
                     float* __restrict__ part_dbias, int64_t n, int d) {
    extern __shared__ float sm[];  // 3 * 8 warps * d
    const int lane = lane_id();
    const int w = threadIdx.x >> 5;
    const int64_t warps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
    float dg[NV4 * 4], db[NV4 * 4], dz[NV4 * 4], gm[NV4 * 4];
#pragma unroll
    for (int j = 0; j < NV4; ++j) {
        const int c = j * 128 + lane * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            dg[4 * j + i] = 0.f;
            db[4 * j + i] = 0.f;
            dz[4 * j + i] = 0.f;
            gm[4 * j + i] = c < d ? gamma[c + i] : 0.f;
        }
    }
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + w; row < n; row += warps) {
        const float mu = mean[row], rs = rstd[row];
        float g[NV4 * 4], xh[NV4 * 4];
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV4; ++j) {
            const int c = j * 128 + lane * 4;
            if (c < d) {
                const float4 u = ld_stream_f4(gy + row * d + c);
                const float4 x = ld_stream_f4(z + row * d + c);
                const float uu[4] = {u.x, u.y, u.z, u.w};
                const float xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int e = 4 * j + i;
                    xh[e] = (xx[i] - mu) * rs;
                    g[e] = uu[i] * gm[e];
                    dg[e] += uu[i] * xh[e];
                    db[e] += uu[i];
                    c1 += g[e];
                    c2 += g[e] * xh[e];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) { g[4 * j + i] = 0.f; xh[4 * j + i] = 0.f; }
            }
        }
        c1 = warp_sum(c1) / d;
        c2 = warp_sum(c2) / d;
#pragma unroll
        for (int j = 0; j < NV4; ++j) {
            const int c = j * 128 + lane * 4;
            if (c < d) {
                float4 o;
                o.x = rs * (g[4 * j] - c1 - xh[4 * j] * c2);
                o.y = rs * (g[4 * j + 1] - c1 - xh[4 * j + 1] * c2);
                o.z = rs * (g[4 * j + 2] - c1 - xh[4 * j + 2] * c2);
                o.w = rs * (g[4 * j + 3] - c1 - xh[4 * j + 3] * c2);
                dz[4 * j] += o.x; dz[4 * j + 1] += o.y; dz[4 * j + 2] += o.z; dz[4 * j + 3] += o.w;
                st_stream_f4(gz + row * d + c, o);
            }
        }
    }
    float* sg = sm + w * d;
    float* sb = sm + (8 + w) * d;
    float* sz = sm + (16 + w) * d;
#pragma unroll
    for (int j = 0; j < NV4; ++j) {
        const int c = j * 128 + lane * 4;
        if (c < d) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                sg[c + i] = dg[4 * j + i];
                sb[c + i] = db[4 * j + i];
                sz[c + i] = dz[4 * j + i];
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        float tg = 0.f, tb = 0.f, tz = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            tg += sm[q * d + c];
            tb += sm[(8 + q) * d + c];
            tz += sm[(16 + q) * d + c];
        }
        part_dgamma[static_cast<int64_t>(blockIdx.x) * d + c] = tg;
        part_dbeta[static_cast<int64_t>(blockIdx.x) * d + c] = tb;
        part_dbias[static_cast<int64_t>(blockIdx.x) * d + c] = tz;
    }
}

// 32 columns per CTA, 8 threads per column each summing every 8th partial row (fixed order), then
// a shared-memory fold of the 8.
__global__ void __launch_bounds__(256)
layernorm_fold_kernel(const float* __restrict__ part_dgamma, const float* __restrict__ part_dbeta,
                      const float* __restrict__ part_dbias, int parts, int d,
                      float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias) {
    __shared__ float sm[3][8][32];
    const int cl = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    float tg = 0.f, tb = 0.f, tz = 0.f;
    if (c < d) {
        for (int p = rg; p < parts; p += 8) {
            tg += part_dgamma[static_cast<int64_t>(p) * d + c];
            tb += part_dbeta[static_cast<int64_t>(p) * d + c];
            tz += part_dbias[static_cast<int64_t>(p) * d + c];
        }
    }
    sm[0][rg][cl] = tg; sm[1][rg][cl] = tb; sm[2][rg][cl] = tz;
    __syncthreads();
    if (rg == 0 && c < d) {
        float a = 0.f, b = 0.f, z = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) { a += sm[0][q][cl]; b += sm[1][q][cl]; z += sm[2][q][cl]; }
        dgamma[c] = a;
        dbeta[c] = b;
        if (dbias) dbias[c] = z;
    }
}

template <int NV4>
static int launch_ln_fwd(const float* a, const float* b, const float* bias, const float* gamma,
                         const float* beta, float* z, float* y, float* mean, float* rstd, int64_t n,
                         int d, float eps, cudaStream_t st) {
    const int64_t blocks = (n * 32 + 255) / 256;
    if (blocks > 0x7fffffffLL) return MK_EUNSUPPORTED;
    add_layernorm_fwd_kernel<NV4><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        a, b, bias, gamma, beta, z, y, mean, rstd, n, d, eps);
    MK_LAUNCH_CHECK("add_layernorm_fwd_kernel");
    return MK_OK;
}

template <int NV4>
static int launch_ln_bwd(const float* gy, const float* z, const float* gamma, const float* mean,
                         const float* rstd, float* gz, float* part_dgamma, float* part_dbeta,
                         float* part_dbias, int parts, int64_t n, int d, cudaStream_t st) {
    const size_t smem = static_cast<size_t>(24) * d * 4;
    auto kern = layernorm_bwd_kernel<NV4>;
    if (smem > 48 * 1024)
        MK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    kern<<<parts, 256, smem, st>>>(gy, z, gamma, mean, rstd, gz, part_dgamma, part_dbeta, part_dbias,
                                   n, d);
    MK_LAUNCH_CHECK("layernorm_bwd_kernel");
    return MK_OK;
}

}  // namespace mk

extern "C" int mk_layernorm_parts(void) { return 148 * 4; }

extern "C" int mk_add_layernorm_fwd(const float* a, const float* b, const float* bias,
                                    const float* gamma, const float* beta, float* z, float* y,
                                    float* mean, float* rstd, int64_t n, int d, float eps,
                                    void* stream) {
    if (n < 0 || d < 4 || d % 4 != 0 || d > 128 * mk::kLNMaxV4) return n < 0 ? MK_EINVAL : MK_EUNSUPPORTED;
    if (n == 0) return MK_OK;
    if (!a || !gamma || !beta || !y || !mean || !rstd) return MK_EINVAL;
    const uintptr_t al = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                         reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(gamma) |
                         reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(z) |
                         reinterpret_cast<uintptr_t>(y);
    if (al % 16) return MK_EINVAL;
    cudaStream_t st = mk::as_stream(stream);
    const int nv4 = (d + 127) / 128;
    switch (nv4) {
        case 1: return mk::launch_ln_fwd<1>(a, b, bias, gamma, beta, z, y, mean, rstd, n, d, eps, st);
        case 2: return mk::launch_ln_fwd<2>(a, b, bias, gamma, beta, z, y, mean, rstd, n, d, eps, st);
        case 3: return mk::launch_ln_fwd<3>(a, b, bias, gamma, beta, z, y, mean, rstd, n, d, eps, st);
        case 4: return mk::launch_ln_fwd<4>(a, b, bias, gamma, beta, z, y, mean, rstd, n, d, eps, st);
        default: return mk::launch_ln_fwd<8>(a, b, bias, gamma, beta, z, y, mean, rstd, n, d, eps, st);
    }
}

extern "C" int mk_layernorm_bwd(const float* gy, const float* z, const float* gamma,
                                const float* mean, const float* rstd, float* gz, float* dgamma,
                                float* dbeta, float* dbias, float* workspace, int64_t n, int d,
                                void* stream) {
    if (n < 0 || d < 4 || d % 4 != 0 || d > 128 * mk::kLNMaxV4) return n < 0 ? MK_EINVAL : MK_EUNSUPPORTED;
    if (!dgamma || !dbeta) return MK_EINVAL;
    cudaStream_t st = mk::as_stream(stream);
    if (n == 0) {
        MK_CUDA_TRY(cudaMemsetAsync(dgamma, 0, sizeof(float) * d, st));
        MK_CUDA_TRY(cudaMemsetAsync(dbeta, 0, sizeof(float) * d, st));
        if (dbias) MK_CUDA_TRY(cudaMemsetAsync(dbias, 0, sizeof(float) * d, st));
        return MK_OK;
    }
    if (!gy || !z || !gamma || !mean || !rstd || !gz || !workspace) return MK_EINVAL;
    const uintptr_t al = reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(z) |
                         reinterpret_cast<uintptr_t>(gz);
    if (al % 16) return MK_EINVAL;
    const int parts = mk_layernorm_parts();
    float* pg = workspace;                                    // [parts, d]
    float* pb = workspace + static_cast<size_t>(parts) * d;   // [parts, d]
    float* pz = workspace + static_cast<size_t>(2) * parts * d;
    const int nv4 = (d + 127) / 128;
    int rc;
    switch (nv4) {
        case 1: rc = mk::launch_ln_bwd<1>(gy, z, gamma, mean, rstd, gz, pg, pb, pz, parts, n, d, st); break;
        case 2: rc = mk::launch_ln_bwd<2>(gy, z, gamma, mean, rstd, gz, pg, pb, pz, parts, n, d, st); break;
        case 3: rc = mk::launch_ln_bwd<3>(gy, z, gamma, mean, rstd, gz, pg, pb, pz, parts, n, d, st); break;
        case 4: rc = mk::launch_ln_bwd<4>(gy, z, gamma, mean, rstd, gz, pg, pb, pz, parts, n, d, st); break;
        default: rc = mk::launch_ln_bwd<8>(gy, z, gamma, mean, rstd, gz, pg, pb, pz, parts, n, d, st); break;
    }
    if (rc != MK_OK) return rc;
    mk::layernorm_fold_kernel<<<(d + 31) / 32, 256, 0, st>>>(pg, pb, pz, parts, d, dgamma, dbeta, dbias);
    MK_LAUNCH_CHECK("layernorm_fold_kernel");
    return MK_OK;
}
