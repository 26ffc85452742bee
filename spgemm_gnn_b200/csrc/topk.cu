// a-1  MaxK nonlinearity: exact per-row top-k emitted as CBSR.
//
// One warp per row.  The row sits in registers (D <= 1024) as NV4 float4 per lane, turned
// into order-preserving uint32 keys; the k-th largest key is found by a 32-step bitwise
// search in which every step is one compare per element plus one REDUX (warp integer
// reduction), with an early exit as soon as a candidate splits the row into exactly k / D-k.
// Ties on the threshold are broken towards the lower column with a warp prefix sum.  The
// kept entries are written in ascending column order.
//
// Replaces: maxk_kernel (so@0x21110) -- one thread per row, 8 bisection steps, approximate --
// and torch.topk + zeros_like + scatter_ + mul (utils/models.py:14-20).
// HBM traffic: reads N*D*4 once, writes N*k*(4+w).
#include <stdlib.h>

#include "common.cuh"
#include "topk.cuh"

namespace mk {

// Column of element (j, i) held by `lane`: j*128 + lane*4 + i.
template <int NV4, typename IdxT, bool VEC>
__global__ void __launch_bounds__(256)
topk_cbsr_reg_kernel(const float* __restrict__ x, int64_t n, int d, int k,
                     float* __restrict__ sp_data, IdxT* __restrict__ sp_index) {
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const int lane = lane_id();
    const float* __restrict__ xr = x + row * d;

    float v[NV4 * 4];
    uint32_t key[NV4 * 4];
#pragma unroll
    for (int j = 0; j < NV4; ++j) {
        const int c0 = j * 128 + lane * 4;
        if (VEC && c0 + 3 < d) {
            const float4 f = ld_stream_f4(xr + c0);
            v[4 * j + 0] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
#pragma unroll
            for (int i = 0; i < 4; ++i) key[4 * j + i] = order_key(v[4 * j + i]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool in = c0 + i < d;
                v[4 * j + i] = in ? ld_stream_f1(xr + (in ? c0 + i : 0)) : 0.0f;
                key[4 * j + i] = in ? order_key(v[4 * j + i]) : 0u;  // below every real key
            }
        }
    }

    // largest T with #{key >= T} >= k
    uint32_t thr = 0;
    bool exact = false;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = thr | (1u << bit);
        const uint32_t ncand = 0u - cand;
        int c = 0;
#pragma unroll
        for (int e = 0; e < NV4 * 4; ++e) count_ge(c, key[e], ncand);
        c = __reduce_add_sync(kFull, c);
        if (c >= k) {
            thr = cand;
            if (c == k) { exact = true; break; }
        }
    }

    uint32_t selmask = 0;  // bit e: element e is kept
    if (exact) {
#pragma unroll
        for (int e = 0; e < NV4 * 4; ++e) selmask |= (key[e] >= thr ? 1u : 0u) << e;
    } else {
        // more than k keys are >= thr: all keys > thr are kept, the lowest-column ties fill up
        int gt = 0;
#pragma unroll
        for (int e = 0; e < NV4 * 4; ++e) gt += (key[e] > thr) ? 1 : 0;
        gt = __reduce_add_sync(kFull, gt);
        const int need = k - gt;
        int before = 0;
#pragma unroll
        for (int j = 0; j < NV4; ++j) {
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) cnt += (key[4 * j + i] == thr) ? 1 : 0;
            const int incl = warp_incl_scan(cnt, lane);
            int rank = before + incl - cnt;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = 4 * j + i;
                if (key[e] > thr) {
                    selmask |= 1u << e;
                } else if (key[e] == thr) {
                    if (rank < need) selmask |= 1u << e;
                    ++rank;
                }
            }
            before += __shfl_sync(kFull, incl, 31);
        }
    }

    // ascending-column slot of every kept element
    float* __restrict__ od = sp_data + row * k;
    IdxT* __restrict__ oi = sp_index + row * k;
    int before = 0;
#pragma unroll
    for (int j = 0; j < NV4; ++j) {
        const int cnt = __popc((selmask >> (4 * j)) & 0xFu);
        const int incl = warp_incl_scan(cnt, lane);
        int pos = before + incl - cnt;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = 4 * j + i;
            if ((selmask >> e) & 1u) {
                od[pos] = v[e];
                oi[pos] = static_cast<IdxT>(j * 128 + lane * 4 + i);
                ++pos;
            }
        }
        before += __shfl_sync(kFull, incl, 31);
    }
}

// Second mapping of the same algorithm (the default): lane-CONTIGUOUS columns -- lane l holds columns
// [l*EPL, (l+1)*EPL), EPL = NV4*4 -- so ascending column order is (lane, element) order and ONE warp
// prefix sum ranks every kept entry (the strided mapping above needs one per 128 columns); the kept
// entries are compacted through shared memory and leave as coalesced row stores instead of 2*EPL
// predicated scalar stores with 64-bit address arithmetic each.  Rows come in with 256-bit loads
// (LDG.E.256) when they are 32-byte aligned.  Same search, same tie rule, same output.
// FULL: the row is exactly NV4*128 columns wide (the usual 128 / 256 / 512 ...): no bounds logic at all.
template <int NV4, typename IdxT, int VECW, bool FULL>
__global__ void __launch_bounds__(256)
topk_cbsr_lane_kernel(const float* __restrict__ x, int64_t n, int d_arg, int k,
                      float* __restrict__ sp_data, IdxT* __restrict__ sp_index) {
    extern __shared__ __align__(16) uint2 stage[];  // [warps][k] {value bits, column}
    constexpr int EPL = NV4 * 4;
    const int d = FULL ? NV4 * 128 : d_arg;
    const int w = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * nw + w;
    if (row >= n) return;
    const int lane = lane_id();
    const float* __restrict__ xr = x + row * d;
    const int cb = lane * EPL;

    float v[EPL];
    uint32_t key[EPL];
#pragma unroll
    for (int j = 0; j < NV4; ++j) {
        const int c0 = cb + 4 * j;
        if (VECW == 8 && (j & 1) == 1 && c0 + 3 < d) continue;  // second half of the 256-bit load below
        if (VECW == 8 && (j & 1) == 0 && j + 1 < NV4 && c0 + 7 < d) {
            asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(v[4 * j + 0]), "=f"(v[4 * j + 1]), "=f"(v[4 * j + 2]), "=f"(v[4 * j + 3]),
                           "=f"(v[4 * j + 4]), "=f"(v[4 * j + 5]), "=f"(v[4 * j + 6]), "=f"(v[4 * j + 7])
                         : "l"(xr + c0));
#pragma unroll
            for (int i = 0; i < 8; ++i) key[4 * j + i] = order_key(v[4 * j + i]);
        } else if (VECW >= 4 && c0 + 3 < d) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(xr + c0));  // neighbours share sectors: via L1
            v[4 * j + 0] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
#pragma unroll
            for (int i = 0; i < 4; ++i) key[4 * j + i] = order_key(v[4 * j + i]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool in = c0 + i < d;
                v[4 * j + i] = in ? __ldg(xr + (in ? c0 + i : 0)) : 0.0f;
                key[4 * j + i] = in ? order_key(v[4 * j + i]) : 0u;  // below every real key
            }
        }
    }

    // A lower bound of the answer that costs a handful of instructions: every lane holds at least j keys
    // >= its own j-th largest, so at least 32*j >= k keys of the row are >= the smallest of those over the
    // lanes (j = 1 for k <= 32, 2 for k <= 64).  Candidates not above the bound are accepted without counting.
    uint32_t lb = 0;
    if (k <= 32) {
        uint32_t m1 = key[0];
#pragma unroll
        for (int e = 1; e < EPL; ++e) m1 = max(m1, key[e]);
        lb = __reduce_min_sync(kFull, m1);
    } else if (k <= 64 && EPL >= 2) {
        uint32_t m1 = max(key[0], key[1]), m2 = min(key[0], key[1]);
#pragma unroll
        for (int e = 2; e < EPL; ++e) {
            m2 = max(m2, min(m1, key[e]));
            m1 = max(m1, key[e]);
        }
        lb = __reduce_min_sync(kFull, m2);
    }

    // largest T with #{key >= T} >= k
    uint32_t thr = 0;
    bool exact = false;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = thr | (1u << bit);
        if (cand <= lb) { thr = cand; continue; }
        const uint32_t ncand = 0u - cand;
        int c = 0;
#pragma unroll
        for (int e = 0; e < EPL; ++e) count_ge(c, key[e], ncand);
        c = __reduce_add_sync(kFull, c);
        if (c >= k) {
            thr = cand;
            if (c == k) { exact = true; break; }
        }
    }

    uint32_t selmask = 0;  // bit e: element e is kept
    if (exact) {
#pragma unroll
        for (int e = 0; e < EPL; ++e) selmask |= (key[e] >= thr ? 1u : 0u) << e;
    } else {
        // more than k keys are >= thr: all keys > thr are kept, the lowest-column ties fill up
        int gt = 0, eq = 0;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            gt += (key[e] > thr) ? 1 : 0;
            eq += (key[e] == thr) ? 1 : 0;
        }
        const int need = k - __reduce_add_sync(kFull, gt);
        int rank = warp_incl_scan(eq, lane) - eq;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            if (key[e] > thr) {
                selmask |= 1u << e;
            } else if (key[e] == thr) {
                if (rank < need) selmask |= 1u << e;
                ++rank;
            }
        }
    }

    // ascending-column slot of every kept element; the row is assembled in shared memory as
    // {value, column} pairs: one predicated 64-bit store per element, no branches
    uint2* __restrict__ sp = stage + w * k;
    const int cnt = __popc(selmask);
    // exclusive prefix sum of a count that fits NB bits: one ballot + popc per bit
    constexpr int NB = EPL <= 4 ? 3 : EPL <= 8 ? 4 : EPL <= 16 ? 5 : 6;
    unsigned lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
    int pos = 0;
#pragma unroll
    for (int b = 0; b < NB; ++b) pos += __popc(__ballot_sync(kFull, (cnt & (1 << b)) != 0) & lt) << b;
    uint32_t pa = static_cast<uint32_t>(__cvta_generic_to_shared(sp + pos));
#pragma unroll
    for (int e = 0; e < EPL; ++e)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.shared.v2.b32 [%0], {%1, %2};\n\t"
                     "@p add.u32 %0, %0, 8;\n\t}"
                     : "+r"(pa)
                     : "r"(__float_as_uint(v[e])), "r"(cb + e), "r"(selmask & (1u << e))
                     : "memory");
    __syncwarp();
    float* __restrict__ od = sp_data + row * k;
    IdxT* __restrict__ oi = sp_index + row * k;
    if (k <= 32) {  // the usual case without the loop's trip-count arithmetic
        if (lane < k) {
            const uint2 pr = sp[lane];
            od[lane] = __uint_as_float(pr.x);
            oi[lane] = static_cast<IdxT>(pr.y);
        }
        return;
    }
    for (int t = lane; t < k; t += 32) {
        const uint2 pr = sp[t];
        od[t] = __uint_as_float(pr.x);
        oi[t] = static_cast<IdxT>(pr.y);
    }
}

// Any D: same algorithm with strided loops.  STAGED: the keys live in shared memory (D <= 49152);
// otherwise (up to the 65536 columns a uint16 id can name) every pass re-derives them from the row,
// which the first pass left in L2.
template <typename IdxT, bool STAGED>
__global__ void __launch_bounds__(32)
topk_cbsr_smem_kernel(const float* __restrict__ x, int64_t n, int d, int k,
                      float* __restrict__ sp_data, IdxT* __restrict__ sp_index) {
    extern __shared__ uint32_t skey_buf[];
    const int lane = lane_id();
    for (int64_t row = blockIdx.x; row < n; row += gridDim.x) {
        const float* __restrict__ xr = x + row * d;
        struct Keys {
            const uint32_t* s;
            const float* g;
            __device__ __forceinline__ uint32_t operator[](int c) const {
                return STAGED ? s[c] : order_key(__ldg(g + c));
            }
        } skey{skey_buf, xr};
        if (STAGED) {
            for (int c = lane; c < d; c += 32) skey_buf[c] = order_key(ld_stream_f1(xr + c));
            __syncwarp();
        }
        uint32_t thr = 0;
        bool exact = false;
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t cand = thr | (1u << bit);
            int c = 0;
            for (int q = lane; q < d; q += 32) c += (skey[q] >= cand) ? 1 : 0;
            c = __reduce_add_sync(kFull, c);
            if (c >= k) {
                thr = cand;
                if (c == k) { exact = true; break; }
            }
        }
        int need = 0;
        if (!exact) {
            int gt = 0;
            for (int q = lane; q < d; q += 32) gt += (skey[q] > thr) ? 1 : 0;
            need = k - __reduce_add_sync(kFull, gt);
        }
        float* __restrict__ od = sp_data + row * k;
        IdxT* __restrict__ oi = sp_index + row * k;
        int pos = 0, ties = 0;
        const unsigned lt = (1u << lane) - 1u;
        for (int base = 0; base < d; base += 32) {
            const int c = base + lane;
            const uint32_t kc = c < d ? skey[c] : 0u;
            bool keep;
            if (exact) {
                keep = c < d && kc >= thr;
            } else {
                const bool eq = c < d && kc == thr;
                const unsigned be = __ballot_sync(kFull, eq);
                const int rank = ties + __popc(be & lt);
                keep = c < d && (kc > thr || (eq && rank < need));
                ties += __popc(be);
            }
            const unsigned bk = __ballot_sync(kFull, keep);
            if (keep) {
                const int p = pos + __popc(bk & lt);
                od[p] = xr[c];
                oi[p] = static_cast<IdxT>(c);
            }
            pos += __popc(bk);
        }
        __syncwarp();
    }
}

template <int NV4, typename IdxT>
static int launch_reg(const float* x, int64_t n, int d, int k, float* sp_data, void* sp_index,
                      cudaStream_t st) {
    const int64_t blocks = (n * 32 + 255) / 256;
    if (blocks > 0x7fffffffLL) return MK_EUNSUPPORTED;
    const bool vec = (d % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0);
    if (vec)
        topk_cbsr_reg_kernel<NV4, IdxT, true><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
            x, n, d, k, sp_data, static_cast<IdxT*>(sp_index));
    else
        topk_cbsr_reg_kernel<NV4, IdxT, false><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
            x, n, d, k, sp_data, static_cast<IdxT*>(sp_index));
    MK_LAUNCH_CHECK("topk_cbsr_reg_kernel");
    return MK_OK;
}

template <int NV4, typename IdxT>
static int launch_lane(const float* x, int64_t n, int d, int k, float* sp_data, void* sp_index,
                       cudaStream_t st) {
    // 8 rows per CTA while the 8 staging rows (8 bytes per kept entry) stay within 48 KB
    const int warps = static_cast<size_t>(k) * 8 * 8 <= 48 * 1024 ? 8 : 4;
    const size_t smem = static_cast<size_t>(warps) * k * sizeof(uint2);
    const int64_t blocks = (n + warps - 1) / warps;
    if (blocks > 0x7fffffffLL) return MK_EUNSUPPORTED;
    const bool a16 = (d % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0);
    const bool a32 = (d % 8 == 0) && (reinterpret_cast<uintptr_t>(x) % 32 == 0) && NV4 % 2 == 0;
    const bool full = d == NV4 * 128;
    IdxT* oi = static_cast<IdxT*>(sp_index);
    const unsigned g = static_cast<unsigned>(blocks), b = static_cast<unsigned>(warps * 32);
    if (a32 && full)
        topk_cbsr_lane_kernel<NV4, IdxT, 8, true><<<g, b, smem, st>>>(x, n, d, k, sp_data, oi);
    else if (a32)
        topk_cbsr_lane_kernel<NV4, IdxT, 8, false><<<g, b, smem, st>>>(x, n, d, k, sp_data, oi);
    else if (a16 && full)
        topk_cbsr_lane_kernel<NV4, IdxT, 4, true><<<g, b, smem, st>>>(x, n, d, k, sp_data, oi);
    else if (a16)
        topk_cbsr_lane_kernel<NV4, IdxT, 4, false><<<g, b, smem, st>>>(x, n, d, k, sp_data, oi);
    else
        topk_cbsr_lane_kernel<NV4, IdxT, 0, false><<<g, b, smem, st>>>(x, n, d, k, sp_data, oi);
    MK_LAUNCH_CHECK("topk_cbsr_lane_kernel");
    return MK_OK;
}

template <typename IdxT>
static int launch_topk(const float* x, int64_t n, int d, int k, float* sp_data, void* sp_index,
                       cudaStream_t st) {
    // MAXK_TOPK_STRIDED=1: the round-1 mapping (columns strided over the lanes), kept for A/B runs
    static const bool strided = [] { const char* e = getenv("MAXK_TOPK_STRIDED"); return e && e[0] == '1'; }();
    if (!strided && d > 128) {  // up to 128 columns (4 per lane) the strided mapping measured faster (0.093 vs 0.101 ms)
        if (d <= 256) return launch_lane<2, IdxT>(x, n, d, k, sp_data, sp_index, st);
        if (d <= 384) return launch_lane<3, IdxT>(x, n, d, k, sp_data, sp_index, st);
        if (d <= 512) return launch_lane<4, IdxT>(x, n, d, k, sp_data, sp_index, st);
        if (d <= 768) return launch_lane<6, IdxT>(x, n, d, k, sp_data, sp_index, st);
        if (d <= 1024) return launch_lane<8, IdxT>(x, n, d, k, sp_data, sp_index, st);
    }
    if (d <= 128) return launch_reg<1, IdxT>(x, n, d, k, sp_data, sp_index, st);
    if (d <= 256) return launch_reg<2, IdxT>(x, n, d, k, sp_data, sp_index, st);
    if (d <= 384) return launch_reg<3, IdxT>(x, n, d, k, sp_data, sp_index, st);
    if (d <= 512) return launch_reg<4, IdxT>(x, n, d, k, sp_data, sp_index, st);
    if (d <= 768) return launch_reg<6, IdxT>(x, n, d, k, sp_data, sp_index, st);
    if (d <= 1024) return launch_reg<8, IdxT>(x, n, d, k, sp_data, sp_index, st);
    const unsigned blocks = static_cast<unsigned>(n < 148 * 32 ? n : 148 * 32);
    if (d > 49152) {
        topk_cbsr_smem_kernel<IdxT, false><<<blocks, 32, 0, st>>>(x, n, d, k, sp_data,
                                                                  static_cast<IdxT*>(sp_index));
        MK_LAUNCH_CHECK("topk_cbsr_smem_kernel");
        return MK_OK;
    }
    const size_t smem = static_cast<size_t>(d) * 4;
    if (smem > 48 * 1024)
        MK_CUDA_TRY(cudaFuncSetAttribute(topk_cbsr_smem_kernel<IdxT, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    topk_cbsr_smem_kernel<IdxT, true><<<blocks, 32, smem, st>>>(x, n, d, k, sp_data,
                                                                static_cast<IdxT*>(sp_index));
    MK_LAUNCH_CHECK("topk_cbsr_smem_kernel");
    return MK_OK;
}

int launch_topk_tile(const float* x, int64_t n, int d, int k, float* sp_data, void* sp_index, int index_bytes,
                     cudaStream_t st);  // topk_tile.cu

}  // namespace mk

extern "C" int mk_topk_cbsr(const float* x, int64_t n, int d, int k, float* sp_data,
                            void* sp_index, int index_bytes, void* stream) {
    if (n < 0 || d < 1 || k < 1 || k > d) return MK_EINVAL;
    if (index_bytes != 1 && index_bytes != 2) return MK_EINVAL;
    if ((index_bytes == 1 && d > 256) || d > 65536) return MK_EINVAL;
    if (n == 0) return MK_OK;
    if (!x || !sp_data || !sp_index) return MK_EINVAL;
    cudaStream_t st = mk::as_stream(stream);
    // MAXK_TOPK_TILE=1: the second-generation kernel (topk_tile.cu: interpolation search, shared-memory
    // compaction, row prefetch).  Measured SLOWER than the kernel below on a B200 (profiles/r2/topk_tile.log:
    // 0.197 vs 0.164 ms on the Reddit shape): its ~5.5 counting steps carry ~45 warp-uniform
    // instructions of bracket arithmetic each, against 13.7 steps of 20 instructions here.
    static const bool tile = [] { const char* e = getenv("MAXK_TOPK_TILE"); return e && e[0] == '1'; }();
    if (tile && d <= 1024) {
        const int rc = mk::launch_topk_tile(x, n, d, k, sp_data, sp_index, index_bytes, st);
        if (rc != MK_EUNSUPPORTED) return rc;
    }
    return index_bytes == 1 ? mk::launch_topk<uint8_t>(x, n, d, k, sp_data, sp_index, st)
                            : mk::launch_topk<uint16_t>(x, n, d, k, sp_data, sp_index, st);
}
