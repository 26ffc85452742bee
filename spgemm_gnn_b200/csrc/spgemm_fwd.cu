// a-3  Forward row-wise-product SpGEMM:  out = A(CSR) x Xs(CBSR)  -> dense [n_rows, D].
//
// One warp per work record (<= max_nz stored entries of one CSR row), one record per CTA so
// that the hardware CTA scheduler does the dynamic load balancing over the skewed degree
// distribution.  The warp owns a dense fp32 accumulator of D floats in shared memory
// (one per lane group when k < 32).  Per stored entry (r <- j) a group of LPN = min(k,32)
// lanes gathers j's CBSR row with one coalesced load of the values (k*4 B) and one of the
// column ids (k*w B) and adds val*data into acc[col] -- plain LDS/FFMA/STS, race-free because
// the columns of a CBSR row are distinct.  Gathers of U neighbours are issued before their
// accumulation so that U*2 loads per lane are in flight.
//
// Differences from the reference kernel (spmm_kernel_opt2_sparse_v3, so@0x24b60):
//   * records come from mk_partition on the GPU, not from a .warp4 file;
//   * a row with a single record is written with full-line vector stores, a row with several
//     records goes through a partial buffer folded in fixed order: no float atomics
//     (the reference does D RED.E.ADD.F32 per <=64-entry record), bit-reproducible output;
//   * for k < 32 all 32 lanes work (32/k neighbours per step) instead of the first 12
//     sub-warps of a 384-thread block;
//   * entries whose value is exactly 0.0 are skipped, which makes the zero-padded rows of
//     utils/maxk_layers.py:245-257 harmless (they race in the reference).
//
// Algorithmic bytes per launch:
//   E*(4 + 4 + k*(4+w)) + N*D*4 + (N+1)*4 + P*16.
#include <stdlib.h>

#include "common.cuh"
#include "epilogue.cuh"

namespace mk {

template <int K>
struct FwdShape {
    static constexpr int LPN = K < 32 ? K : 32;  // lanes per neighbour
    static constexpr int EPL = K / LPN;          // entries per lane
    static constexpr int G = 32 / LPN;           // neighbours per warp step
    static_assert(K % LPN == 0, "K must be a power of two <= 32 or a multiple of 32");
};

template <int K, typename IdxT, int U, bool VEC>
__global__ void __launch_bounds__(32)
spgemm_fwd_kernel(const mk_part* __restrict__ parts, const int* __restrict__ idx,
                  const float* __restrict__ val, const float* __restrict__ sp_data,
                  const IdxT* __restrict__ sp_index, float* __restrict__ out,
                  float* __restrict__ partial, int d) {
    using S = FwdShape<K>;
    extern __shared__ __align__(16) float acc[];  // S::G accumulators of dpad floats
    const int dpad = (d + 3) & ~3;
    const int lane = lane_id();
    const int g = lane / S::LPN;
    const int t = lane % S::LPN;
    const mk_part rec = parts[blockIdx.x];

    for (int c = lane * 4; c < S::G * dpad; c += 128)
        *reinterpret_cast<float4*>(acc + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    float* __restrict__ my = acc + g * dpad;
    const unsigned gmask = (S::LPN >= 32 ? kFull : ((1u << (S::LPN & 31)) - 1u)) << (g * S::LPN);
    const int end = rec.loc + rec.len;
    for (int base = rec.loc; base < end; base += 32) {
        const int n_here = min(32, end - base);
        int my_nz = 0;
        float my_v = 0.f;
        if (lane < n_here) {
            my_nz = ld_stream_i1(idx + base + lane);
            my_v = ld_stream_f1(val + base + lane);
        }
        for (int i = 0; i < n_here; i += S::G * U) {
            float dv[U][S::EPL];
            int cv[U][S::EPL];
            float vv[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = i + u * S::G + g;
                const int nz = __shfl_sync(kFull, my_nz, e & 31);
                vv[u] = __shfl_sync(kFull, my_v, e & 31);
                ok[u] = e < n_here;
                if (ok[u]) {
                    const float* __restrict__ dp = sp_data + static_cast<int64_t>(nz) * K;
                    const IdxT* __restrict__ ip = sp_index + static_cast<int64_t>(nz) * K;
#pragma unroll
                    for (int q = 0; q < S::EPL; ++q) {
                        dv[u][q] = __ldg(dp + t + q * S::LPN);
                        cv[u][q] = static_cast<int>(__ldg(ip + t + q * S::LPN));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (ok[u]) {  // uniform over the lanes of a group
#pragma unroll
                    for (int q = 0; q < S::EPL; ++q)
                        if (dv[u][q] != 0.0f) my[cv[u][q]] += vv[u] * dv[u][q];
                    accum_fence_group(gmask);
                }
                accum_fence_warp();
            }
        }
    }
    __syncwarp();

    float* __restrict__ o = rec.slot < 0 ? out + static_cast<int64_t>(rec.row) * d
                                         : partial + static_cast<int64_t>(rec.slot) * d;
    if (VEC) {
        for (int c = lane * 4; c < d; c += 128) {
            float4 s = *reinterpret_cast<const float4*>(acc + c);
#pragma unroll
            for (int q = 1; q < S::G; ++q) {
                const float4 a = *reinterpret_cast<const float4*>(acc + q * dpad + c);
                s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
            }
            st_stream_f4(o + c, s);
        }
    } else {
        for (int c = lane; c < d; c += 32) {
            float s = acc[c];
#pragma unroll
            for (int q = 1; q < S::G; ++q) s += acc[q * dpad + c];
            o[c] = s;
        }
    }
}


// ---- vectorised variant ----------------------------------------------------------------------
// Each lane handles EPL consecutive entries of one neighbour (one 16-byte load of values and
// one 4/8-byte load of column ids when EPL == 4), so a warp step covers G = 32 / (K/EPL)
// neighbours.  Lane groups working on different neighbours could hit the same column, so each
// group accumulates into its own copy; the G copies are summed when the row is written.
// Against the scalar variant this removes 3/4 of the SHFL broadcasts and 3/8 of the global-load
// wavefronts; the shared-memory RMW traffic (the L1TEX-pipe bound, see profiles/) is unchanged.
template <int EPL, typename IdxT>
__device__ __forceinline__ void load_entries(const float* __restrict__ dp,
                                             const IdxT* __restrict__ ip, float (&dv)[EPL],
                                             int (&cv)[EPL]);
template <>
__device__ __forceinline__ void load_entries<4, uint8_t>(const float* __restrict__ dp,
                                                         const uint8_t* __restrict__ ip,
                                                         float (&dv)[4], int (&cv)[4]) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(dp));
    const uchar4 q = __ldg(reinterpret_cast<const uchar4*>(ip));
    dv[0] = f.x; dv[1] = f.y; dv[2] = f.z; dv[3] = f.w;
    cv[0] = q.x; cv[1] = q.y; cv[2] = q.z; cv[3] = q.w;
}
template <>
__device__ __forceinline__ void load_entries<4, uint16_t>(const float* __restrict__ dp,
                                                          const uint16_t* __restrict__ ip,
                                                          float (&dv)[4], int (&cv)[4]) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(dp));
    const ushort4 q = __ldg(reinterpret_cast<const ushort4*>(ip));
    dv[0] = f.x; dv[1] = f.y; dv[2] = f.z; dv[3] = f.w;
    cv[0] = q.x; cv[1] = q.y; cv[2] = q.z; cv[3] = q.w;
}
template <>
__device__ __forceinline__ void load_entries<2, uint8_t>(const float* __restrict__ dp,
                                                         const uint8_t* __restrict__ ip,
                                                         float (&dv)[2], int (&cv)[2]) {
    const float2 f = __ldg(reinterpret_cast<const float2*>(dp));
    const uchar2 q = __ldg(reinterpret_cast<const uchar2*>(ip));
    dv[0] = f.x; dv[1] = f.y;
    cv[0] = q.x; cv[1] = q.y;
}
template <>
__device__ __forceinline__ void load_entries<2, uint16_t>(const float* __restrict__ dp,
                                                          const uint16_t* __restrict__ ip,
                                                          float (&dv)[2], int (&cv)[2]) {
    const float2 f = __ldg(reinterpret_cast<const float2*>(dp));
    const ushort2 q = __ldg(reinterpret_cast<const ushort2*>(ip));
    dv[0] = f.x; dv[1] = f.y;
    cv[0] = q.x; cv[1] = q.y;
}
template <>
__device__ __forceinline__ void load_entries<1, uint8_t>(const float* __restrict__ dp,
                                                         const uint8_t* __restrict__ ip,
                                                         float (&dv)[1], int (&cv)[1]) {
    dv[0] = __ldg(dp);
    cv[0] = __ldg(ip);
}
template <>
__device__ __forceinline__ void load_entries<1, uint16_t>(const float* __restrict__ dp,
                                                          const uint16_t* __restrict__ ip,
                                                          float (&dv)[1], int (&cv)[1]) {
    dv[0] = __ldg(dp);
    cv[0] = __ldg(ip);
}

// Measurement variant (off): group-private BANKS instead of group-private copies laid end to end.  On
// paper a step then costs the largest bank load inside a group (~3.0) instead of 32 entries into 32
// banks (3.6); measured on a B200 it is SLOWER on every short-record shape
// (profiles/r2/fwd_vec_group_banks_call28.log: products 5.69 -> 6.55 ms, Flickr 0.079 -> 0.111, Yelp
// 0.90 -> 1.14).  Not profiled further; the balls-into-bins estimate alone does not decide this kernel.
#ifndef MK_FWD_VEC_GROUP_BANKS
#define MK_FWD_VEC_GROUP_BANKS 0
#endif

template <int K, int EPL, typename IdxT, int U>
__global__ void __launch_bounds__(32)
spgemm_fwd_vec_kernel(const mk_part* __restrict__ parts, const int* __restrict__ idx,
                      const float* __restrict__ val, const float* __restrict__ sp_data,
                      const IdxT* __restrict__ sp_index, float* __restrict__ out,
                      float* __restrict__ partial, int d) {
    constexpr int LPN = K / EPL;  // lanes per neighbour
    constexpr int G = 32 / LPN;   // neighbours per warp step == accumulator copies
    static_assert(K % EPL == 0 && 32 % LPN == 0 && LPN <= 32, "unsupported K / EPL");
    extern __shared__ __align__(16) float acc[];
    const int dpad = (d + 3) & ~3;
    const int lane = lane_id();
    const int g = lane / LPN;
    const int t = lane % LPN;
    const mk_part rec = parts[blockIdx.x];
#if MK_FWD_VEC_GROUP_BANKS
    // column c of group g lives in bank LPN*g + c % LPN of row c / LPN (see the note at the macro)
    constexpr int LOG_BG = LPN >= 32 ? 5 : LPN >= 16 ? 4 : LPN >= 8 ? 3 : LPN >= 4 ? 2 : LPN >= 2 ? 1 : 0;
    static_assert((1 << LOG_BG) == LPN, "lanes per neighbour must be a power of two");
    const int rows = (dpad + LPN - 1) >> LOG_BG;
    const int cells = rows << 5;
    auto cell = [](int c) { return ((c >> LOG_BG) << 5) + (c & (LPN - 1)); };
    float* __restrict__ my = acc + LPN * g;
#else
    const int cells = G * dpad;
    auto cell = [](int c) { return c; };
    float* __restrict__ my = acc + g * dpad;
#endif

    for (int c = lane * 4; c < cells; c += 128)
        *reinterpret_cast<float4*>(acc + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    const unsigned gmask = (LPN >= 32 ? kFull : ((1u << (LPN & 31)) - 1u)) << (g * LPN);
    const int end = rec.loc + rec.len;
    for (int base = rec.loc; base < end; base += 32) {
        const int n_here = min(32, end - base);
        int my_nz = 0;
        float my_v = 0.f;
        if (lane < n_here) {
            my_nz = ld_stream_i1(idx + base + lane);
            my_v = ld_stream_f1(val + base + lane);
        }
        for (int i = 0; i < n_here; i += G * U) {
            float dv[U][EPL];
            int cv[U][EPL];
            float vv[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = i + u * G + g;
                const int nz = __shfl_sync(kFull, my_nz, e & 31);
                vv[u] = __shfl_sync(kFull, my_v, e & 31);
                ok[u] = e < n_here;
                if (ok[u])
                    load_entries<EPL, IdxT>(sp_data + static_cast<int64_t>(nz) * K + EPL * t,
                                            sp_index + static_cast<int64_t>(nz) * K + EPL * t,
                                            dv[u], cv[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (ok[u]) {  // uniform over the lanes of a group
#pragma unroll
                    for (int q = 0; q < EPL; ++q)
                        if (dv[u][q] != 0.0f) my[cell(cv[u][q])] += vv[u] * dv[u][q];
                    accum_fence_group(gmask);
                }
                accum_fence_warp();
            }
        }
    }
    __syncwarp();

    float* __restrict__ o = rec.slot < 0 ? out + static_cast<int64_t>(rec.row) * d
                                         : partial + static_cast<int64_t>(rec.slot) * d;
    for (int c = lane * 4; c < d; c += 128) {
#if MK_FWD_VEC_GROUP_BANKS
        const float* __restrict__ base = acc + cell(c);   // LPN >= 4: the four columns share a row segment
        float4 s = *reinterpret_cast<const float4*>(base);
#pragma unroll
        for (int q = 1; q < G; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(base + LPN * q);
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
#else
        float4 s = *reinterpret_cast<const float4*>(acc + c);
#pragma unroll
        for (int q = 1; q < G; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(acc + q * dpad + c);
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
#endif
        st_stream_f4(o + c, s);
    }
}

// Any k: one neighbour per warp step, lanes stride over its entries.
template <typename IdxT>
__global__ void __launch_bounds__(32)
spgemm_fwd_generic_kernel(const mk_part* __restrict__ parts, const int* __restrict__ idx,
                          const float* __restrict__ val, const float* __restrict__ sp_data,
                          const IdxT* __restrict__ sp_index, float* __restrict__ out,
                          float* __restrict__ partial, int k, int d) {
    extern __shared__ __align__(16) float acc[];
    const int lane = lane_id();
    const mk_part rec = parts[blockIdx.x];
    for (int c = lane; c < d; c += 32) acc[c] = 0.f;
    __syncwarp();
    const int end = rec.loc + rec.len;
    for (int e = rec.loc; e < end; ++e) {
        const int64_t nz = idx[e];
        const float v = val[e];
        for (int q = lane; q < k; q += 32) {
            const float dv = __ldg(sp_data + nz * k + q);
            const int c = static_cast<int>(__ldg(sp_index + nz * k + q));
            if (dv != 0.0f) acc[c] += v * dv;
        }
        __syncwarp();
    }
    float* __restrict__ o = rec.slot < 0 ? out + static_cast<int64_t>(rec.row) * d
                                         : partial + static_cast<int64_t>(rec.slot) * d;
    for (int c = lane; c < d; c += 32) o[c] = acc[c];
}

// Folds the partial rows of every multi-record row in slot order.  One THREAD per record finds the
// first records of such rows (3 % of the records of a Reddit-shaped graph), the block's 8 warps then
// fold them, float4 wide -- the round-1 form spent a whole warp on every record (0.049 -> 0.0xx ms).
// The sums of a row are formed in slot order whichever warp takes it, so the result does not depend
// on the order in which the block lists its rows.
template <bool LN>
__global__ void __launch_bounds__(256)
spgemm_fold_kernel(const mk_part* __restrict__ parts, int64_t num_parts, const float* __restrict__ partial,
                   float* __restrict__ out, int d, int accumulate, const FwdEpilogue ep) {
    __shared__ int list[256];
    __shared__ int n_list;
    const int64_t base = static_cast<int64_t>(blockIdx.x) * 256;
    const int64_t p = base + threadIdx.x;
    if (threadIdx.x == 0) n_list = 0;
    __syncthreads();
    if (p < num_parts) {
        const mk_part rec = parts[p];
        if (rec.slot >= 0 && (p == 0 || parts[p - 1].row != rec.row)) list[atomicAdd(&n_list, 1)] = threadIdx.x;
    }
    __syncthreads();
    const int n = n_list;
    const int lane = lane_id();
    const bool vec = (d % 4 == 0) && (reinterpret_cast<uintptr_t>(partial) % 16 == 0) &&
                     (reinterpret_cast<uintptr_t>(out) % 16 == 0);
    for (int i = threadIdx.x >> 5; i < n; i += 8) {
        const int64_t q0 = base + list[i];
        const mk_part rec = parts[q0];
        int cnt = 1;
        while (q0 + cnt < num_parts && parts[q0 + cnt].row == rec.row) ++cnt;
        const float* __restrict__ src = partial + static_cast<int64_t>(rec.slot) * d;
        if constexpr (LN) {
            // followed by the f-3 epilogue (same sums as the plain fold, then epilogue.cuh): the rows the
            // forward kernel could not finish by itself
            auto agg = [&](int c) {
                float4 s4 = *reinterpret_cast<const float4*>(src + c);
                for (int q = 1; q < cnt; ++q) {
                    const float4 t = *reinterpret_cast<const float4*>(src + static_cast<int64_t>(q) * d + c);
                    s4.x += t.x; s4.y += t.y; s4.z += t.z; s4.w += t.w;
                }
                return s4;
            };
            const float* __restrict__ hrow = ep.h_self ? ep.h_self + static_cast<int64_t>(rec.row) * d : nullptr;
            fwd_epilogue_row(ep, agg, [&](int c) { return ld_stream_f4(hrow + c); }, out, rec.row, d, lane);
        } else {
            float* __restrict__ o = out + static_cast<int64_t>(rec.row) * d;
            if (vec) {
                for (int c = lane * 4; c < d; c += 128) {
                    float4 s4 = *reinterpret_cast<const float4*>(src + c);
                    for (int q = 1; q < cnt; ++q) {
                        const float4 t = *reinterpret_cast<const float4*>(src + static_cast<int64_t>(q) * d + c);
                        s4.x += t.x; s4.y += t.y; s4.z += t.z; s4.w += t.w;
                    }
                    if (accumulate) {
                        const float4 a = *reinterpret_cast<const float4*>(o + c);
                        s4.x = a.x + s4.x; s4.y = a.y + s4.y; s4.z = a.z + s4.z; s4.w = a.w + s4.w;
                    }
                    *reinterpret_cast<float4*>(o + c) = s4;
                }
            } else {
                for (int c = lane; c < d; c += 32) {
                    float s1 = src[c];
                    for (int q = 1; q < cnt; ++q) s1 += src[static_cast<int64_t>(q) * d + c];
                    o[c] = accumulate ? o[c] + s1 : s1;
                }
            }
        }
    }
}

int launch_fold(const mk_part* parts, int64_t num_parts, const float* partial, float* out, int d,
                cudaStream_t st, int accumulate, const FwdEpilogue* ep) {
    const int64_t blocks = (num_parts + 255) / 256;
    if (blocks > 0x7fffffffLL) return MK_EUNSUPPORTED;
    if (ep != nullptr) {
        spgemm_fold_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, st>>>(parts, num_parts, partial, out, d, 0, *ep);
        MK_LAUNCH_CHECK("spgemm_fold_ln_kernel");
        return MK_OK;
    }
    spgemm_fold_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, st>>>(parts, num_parts, partial, out, d,
                                                                             accumulate, FwdEpilogue{});
    MK_LAUNCH_CHECK("spgemm_fold_kernel");
    return MK_OK;
}

template <int K, typename IdxT>
static int launch_fwd_k(const mk_part* parts, int64_t num_parts, const int* idx, const float* val,
                        const float* sp_data, const void* sp_index, float* out, float* partial,
                        int d, cudaStream_t st) {
    using S = FwdShape<K>;
    constexpr int U = K >= 64 ? 4 : 8;
    const int dpad = (d + 3) & ~3;
    const size_t smem = static_cast<size_t>(S::G) * dpad * 4;
    if (smem > 200 * 1024) return MK_EUNSUPPORTED;
    const bool vec = (d % 4 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
                     (partial == nullptr || reinterpret_cast<uintptr_t>(partial) % 16 == 0);
    const IdxT* si = static_cast<const IdxT*>(sp_index);
    if (vec) {
        auto kern = spgemm_fwd_kernel<K, IdxT, U, true>;
        if (smem > 48 * 1024)
            MK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem)));
        kern<<<static_cast<unsigned>(num_parts), 32, smem, st>>>(parts, idx, val, sp_data, si, out,
                                                                 partial, d);
    } else {
        auto kern = spgemm_fwd_kernel<K, IdxT, U, false>;
        if (smem > 48 * 1024)
            MK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem)));
        kern<<<static_cast<unsigned>(num_parts), 32, smem, st>>>(parts, idx, val, sp_data, si, out,
                                                                 partial, d);
    }
    MK_LAUNCH_CHECK("spgemm_fwd_kernel");
    return MK_OK;
}


template <int K, int EPL, typename IdxT>
static int launch_fwd_vec(const mk_part* parts, int64_t num_parts, const int* idx, const float* val,
                          const float* sp_data, const void* sp_index, float* out, float* partial,
                          int d, cudaStream_t st) {
    constexpr int G = 32 / (K / EPL);
#ifdef MK_FWD_VEC_U
    constexpr int U = MK_FWD_VEC_U;
#else
    constexpr int U = (EPL >= 4) ? 4 : (32 / G >= 8 ? 8 : 32 / G);  // <= 32 / G steps per slice
#endif
#if MK_FWD_VEC_GROUP_BANKS
    constexpr int LPN = K / EPL;
    const size_t smem = static_cast<size_t>((((d + 3) & ~3) + LPN - 1) / LPN) * 32 * 4;
#else
    const size_t smem = static_cast<size_t>(G) * d * 4;
#endif
    if (smem > 200 * 1024) return MK_EUNSUPPORTED;
    auto kern = spgemm_fwd_vec_kernel<K, EPL, IdxT, U>;
    if (smem > 48 * 1024)
        MK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    kern<<<static_cast<unsigned>(num_parts), 32, smem, st>>>(
        parts, idx, val, sp_data, static_cast<const IdxT*>(sp_index), out, partial, d);
    MK_LAUNCH_CHECK("spgemm_fwd_vec_kernel");
    return MK_OK;
}

static bool use_scalar_variant() {
    static const bool v = [] {
        const char* e = getenv("MAXK_FWD_SCALAR");
        return e && e[0] == '1';
    }();
    return v;
}

template <typename IdxT>
static int launch_fwd(const mk_part* parts, int64_t num_parts, const int* idx, const float* val,
                      const float* sp_data, const void* sp_index, float* out, float* partial,
                      int k, int d, cudaStream_t st) {
    const bool vec_ok = !use_scalar_variant() && d % 4 == 0 &&
                        reinterpret_cast<uintptr_t>(out) % 16 == 0 &&
                        reinterpret_cast<uintptr_t>(sp_data) % 16 == 0 &&
                        reinterpret_cast<uintptr_t>(sp_index) % (4 * sizeof(IdxT)) == 0 &&
                        (partial == nullptr || reinterpret_cast<uintptr_t>(partial) % 16 == 0);
    if (vec_ok) {
        switch (k) {
            case 4: return launch_fwd_vec<4, 1, IdxT>(parts, num_parts, idx, val, sp_data, sp_index, out, partial, d, st);
            case 8: return launch_fwd_vec<8, 1, IdxT>(parts, num_parts, idx, val, sp_data, sp_index, out, partial, d, st);
            case 16: return launch_fwd_vec<16, 2, IdxT>(parts, num_parts, idx, val, sp_data, sp_index, out, partial, d, st);
            case 32: return launch_fwd_vec<32, 4, IdxT>(parts, num_parts, idx, val, sp_data, sp_index, out, partial, d, st);
            case 64: return launch_fwd_vec<64, 4, IdxT>(parts, num_parts, idx, val, sp_data, sp_index, out, partial, d, st);
            case 128: return launch_fwd_vec<128, 4, IdxT>(parts, num_parts, idx, val, sp_data, sp_index, out, partial, d, st);
            default: break;
        }
    }
    switch (k) {
        case 4: return launch_fwd_k<4, IdxT>(parts, num_parts, idx, val, sp_data, sp_index, out, partial, d, st);
        case 8: return launch_fwd_k<8, IdxT>(parts, num_parts, idx, val, sp_data, sp_index, out, partial, d, st);
        case 16: return launch_fwd_k<16, IdxT>(parts, num_parts, idx, val, sp_data, sp_index, out, partial, d, st);
        case 32: return launch_fwd_k<32, IdxT>(parts, num_parts, idx, val, sp_data, sp_index, out, partial, d, st);
        case 64: return launch_fwd_k<64, IdxT>(parts, num_parts, idx, val, sp_data, sp_index, out, partial, d, st);
        case 96: return launch_fwd_k<96, IdxT>(parts, num_parts, idx, val, sp_data, sp_index, out, partial, d, st);
        case 128: return launch_fwd_k<128, IdxT>(parts, num_parts, idx, val, sp_data, sp_index, out, partial, d, st);
        default: break;
    }
    const size_t smem = static_cast<size_t>(d) * 4;
    if (smem > 200 * 1024) return MK_EUNSUPPORTED;
    if (smem > 48 * 1024)
        MK_CUDA_TRY(cudaFuncSetAttribute(spgemm_fwd_generic_kernel<IdxT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    spgemm_fwd_generic_kernel<IdxT><<<static_cast<unsigned>(num_parts), 32, smem, st>>>(
        parts, idx, val, sp_data, static_cast<const IdxT*>(sp_index), out, partial, k, d);
    MK_LAUNCH_CHECK("spgemm_fwd_generic_kernel");
    return MK_OK;
}

}  // namespace mk

extern "C" int mk_spgemm_fwd(const mk_part* parts, int64_t num_parts, int64_t num_slots,
                             const int32_t* idx, const float* val, const float* sp_data,
                             const void* sp_index, int index_bytes, float* out, float* partial,
                             int64_t n_rows, int k, int d, void* stream) {
    if (n_rows < 0 || num_parts < 0 || num_slots < 0 || d < 1 || k < 1 || k > d) return MK_EINVAL;
    if (index_bytes != 1 && index_bytes != 2) return MK_EINVAL;
    if ((index_bytes == 1 && d > 256) || d > 65536) return MK_EINVAL;
    if (n_rows == 0 || num_parts == 0) return MK_OK;
    if (!parts || !out || !sp_data || !sp_index) return MK_EINVAL;  // idx / val may be NULL: a graph without stored entries
    if (num_slots > 0 && !partial) return MK_EINVAL;
    if (num_parts > 0x7fffffffLL) return MK_EUNSUPPORTED;
    cudaStream_t st = mk::as_stream(stream);
    const int rc = index_bytes == 1
                       ? mk::launch_fwd<uint8_t>(parts, num_parts, idx, val, sp_data, sp_index, out,
                                                 partial, k, d, st)
                       : mk::launch_fwd<uint16_t>(parts, num_parts, idx, val, sp_data, sp_index,
                                                  out, partial, k, d, st);
    if (rc != MK_OK) return rc;
    if (num_slots > 0) return mk::launch_fold(parts, num_parts, partial, out, d, st, 0, nullptr);
    return MK_OK;
}
