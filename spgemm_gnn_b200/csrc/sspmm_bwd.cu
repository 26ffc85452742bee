// a-4  Backward sampled SpMM (SSpMM):  dXs = sample(A^T x dY) at the forward's top-k positions,
//      emitted as a CBSR gradient [n_src, k].
//
// Outer-product (push) form, like the reference (spmm_kernel_opt2_sparse_backward_v3,
// so@0x257a0): one warp per work record of CSR row r; dY[r, :] is staged once in shared
// memory; for every stored entry (r <- j) the warp reads j's k column ids, picks dY[r, col]
// out of shared memory, scales by val and adds into dXs[j, :].
// What is different:
//   * each lane handles FOUR consecutive entries of a neighbour: one 4-byte (uint8) or
//     8-byte (uint16) load of column ids and ONE vector reduction `red.global.add.v4.f32`
//     (SASS REDG.E.ADD.F32x4) instead of four scalar RED -- k/4 lanes per neighbour, 128/k
//     neighbours per warp step, all lanes busy for every k in {4..128};
//   * U steps of column ids are loaded before any is consumed (the staged dY row is
//     read-only, so the compiler is free to overlap them);
//   * the record list is the one mk_partition builds on the GPU.
// The sums land in L2 (dXs is n_src*k*4 B: 30 MB for the Reddit shape, L2-resident); their
// order is not fixed, exactly as with the reference's RED.E.ADD.F32.
//
// Algorithmic bytes per launch:
//   E*(4 + 4 + k*w + k*4) + N*D*4 + N*k*4 + (N+1)*4 + P*16.
#include "common.cuh"

// Records per CTA of the plain backward (measurement knob, VERDICT r1 "occupancy ceiling": one 32-thread
// CTA per record caps residency at 32 CTAs = 32 warps per SM).  W warps per CTA, one record each, lift
// the cap to 64 warps.  Measured on a B200 (profiles/r2/bwd_warps_per_cta_call17.log): SLOWER -- Reddit shape
// k = 32: 2.559 ms at 1, 2.862 at 2, 2.867 at 4 warps per CTA; k = 64: 4.744 / 5.276 / 5.291; products shape
// (plain kernel) 6.46 -> 7.27 ms.  The kernel is bound by the SM->L2 request path (93 %), not by latency;
// more warps in flight only add contention for it.  The default stays 1.
#ifndef MK_BWD_WARPS
#define MK_BWD_WARPS 1
#endif

namespace mk {

template <typename IdxT>
__device__ __forceinline__ void load_cols4(const IdxT* p, int (&c)[4]);
template <>
__device__ __forceinline__ void load_cols4<uint8_t>(const uint8_t* p, int (&c)[4]) {
    const uchar4 q = __ldg(reinterpret_cast<const uchar4*>(p));
    c[0] = q.x; c[1] = q.y; c[2] = q.z; c[3] = q.w;
}
template <>
__device__ __forceinline__ void load_cols4<uint16_t>(const uint16_t* p, int (&c)[4]) {
    const ushort4 q = __ldg(reinterpret_cast<const ushort4*>(p));
    c[0] = q.x; c[1] = q.y; c[2] = q.z; c[3] = q.w;
}

template <int K, typename IdxT, int U>
__device__ __forceinline__ void
sspmm_bwd_body(const mk_part* __restrict__ parts, [[maybe_unused]] int64_t num_parts, const int* __restrict__ idx,
               const float* __restrict__ val, const float* __restrict__ dy,
               const IdxT* __restrict__ sp_index, float* __restrict__ dxs, int d, int vec_dy) {
    constexpr int LPN = K / 4;     // lanes per neighbour
    constexpr int G = 32 / LPN;    // neighbours per warp step
    static_assert(K % 4 == 0 && (32 % LPN) == 0, "K must be 4, 8, 16, 32, 64 or 128");
    extern __shared__ __align__(16) float dys_all[];
    const int lane = lane_id();
    const int g = lane / LPN;
    const int t = lane % LPN;
#if MK_BWD_WARPS > 1
    const int64_t rid = static_cast<int64_t>(blockIdx.x) * MK_BWD_WARPS + (threadIdx.x >> 5);
    if (rid >= num_parts) return;
    float* __restrict__ dys = dys_all + (threadIdx.x >> 5) * ((d + 3) & ~3);
    const mk_part rec = parts[rid];
#else
    float* __restrict__ dys = dys_all;
    const mk_part rec = parts[blockIdx.x];
#endif
    if (rec.len == 0) return;

    const float* __restrict__ dyr = dy + static_cast<int64_t>(rec.row) * d;
    if (vec_dy) {
        for (int c = lane * 4; c < d; c += 128)
            *reinterpret_cast<float4*>(dys + c) = ld_stream_f4(dyr + c);
    } else {
        for (int c = lane; c < d; c += 32) dys[c] = ld_stream_f1(dyr + c);
    }
    __syncwarp();

    const int end = rec.loc + rec.len;
#ifdef MK_EDGE_CPASYNC
    // Measured variant (BASELINE.json part 2: "cp.async/TMA to stage ... edge slices"): the next
    // 32-entry slice of idx / val is copied to shared memory with cp.async (LDGSTS) while the current
    // one is consumed, and read back with group-uniform LDS instead of SHFL broadcasts.  Slower on
    // every shape tried (profiles/r2/cpasync_edges.log): the kernel is bound by the L1TEX data pipe
    // and the SM->L2 request path, LDGSTS adds shared-memory write wavefronts and an LDS per broadcast
    // to exactly that pipe, and the latency it hides is already hidden by 32 resident warps per SM.
    __shared__ int s_nz[2][32];
    __shared__ float s_v[2][32];
    auto stage = [&](int buf, int base) {
        if (base + lane < end) {
            const unsigned dn = static_cast<unsigned>(__cvta_generic_to_shared(&s_nz[buf][lane]));
            const unsigned dv = static_cast<unsigned>(__cvta_generic_to_shared(&s_v[buf][lane]));
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dn), "l"(idx + base + lane) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dv), "l"(val + base + lane) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0, rec.loc);
    int cur = 0;
#endif
    for (int base = rec.loc; base < end; base += 32) {
        const int n_here = min(32, end - base);
#ifdef MK_EDGE_CPASYNC
        stage(cur ^ 1, base + 32);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
#else
        int my_nz = 0;
        float my_v = 0.f;
        if (lane < n_here) {
            my_nz = ld_stream_i1(idx + base + lane);
            my_v = ld_stream_f1(val + base + lane);
        }
#endif
        for (int i = 0; i < n_here; i += G * U) {
            int cv[U][4];
            int nzv[U];
            float vv[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = i + u * G + g;
#ifdef MK_EDGE_CPASYNC
                nzv[u] = s_nz[cur][e & 31];
                vv[u] = s_v[cur][e & 31];
#else
                nzv[u] = __shfl_sync(kFull, my_nz, e & 31);
                vv[u] = __shfl_sync(kFull, my_v, e & 31);
#endif
                ok[u] = e < n_here;
                if (ok[u]) load_cols4<IdxT>(sp_index + static_cast<int64_t>(nzv[u]) * K + 4 * t, cv[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (ok[u]) {
                    const float v = vv[u];
                    red_add_f4(dxs + static_cast<int64_t>(nzv[u]) * K + 4 * t, v * dys[cv[u][0]],
                               v * dys[cv[u][1]], v * dys[cv[u][2]], v * dys[cv[u][3]]);
                }
            }
        }
#ifdef MK_EDGE_CPASYNC
        __syncwarp();
        cur ^= 1;
#endif
    }
}

// Two entry points over the same body.  k <= 32: 32 resident CTAs per SM (64 registers), measured
// 2.60 ms against 2.63 ms at 28 CTAs (72 registers) on the Reddit shape; at k = 64 any register cap
// spills and loses (4.90-4.96 vs 4.78 ms), so the wide variants keep the compiler's own choice.
template <int K, typename IdxT, int U>
__global__ void __launch_bounds__(32 * MK_BWD_WARPS, 32 / MK_BWD_WARPS)
sspmm_bwd_kernel_occ32(const mk_part* __restrict__ parts, int64_t num_parts, const int* __restrict__ idx,
                       const float* __restrict__ val, const float* __restrict__ dy,
                       const IdxT* __restrict__ sp_index, float* __restrict__ dxs, int d, int vec_dy) {
    sspmm_bwd_body<K, IdxT, U>(parts, num_parts, idx, val, dy, sp_index, dxs, d, vec_dy);
}
template <int K, typename IdxT, int U>
__global__ void __launch_bounds__(32 * MK_BWD_WARPS)
sspmm_bwd_kernel(const mk_part* __restrict__ parts, int64_t num_parts, const int* __restrict__ idx,
                 const float* __restrict__ val, const float* __restrict__ dy,
                 const IdxT* __restrict__ sp_index, float* __restrict__ dxs, int d, int vec_dy) {
    sspmm_bwd_body<K, IdxT, U>(parts, num_parts, idx, val, dy, sp_index, dxs, d, vec_dy);
}

// ---- column-blocked, row-tiled form for CBSR gradients that do not fit L2 -------------------------
// On a products-shaped graph (2.45 M nodes, mean degree 51) dXs is 313 MB: every vector reduction of
// the kernel above misses L2 and becomes a DRAM read-modify-write (ncu: 36.2 GB of DRAM traffic for
// 24 GB algorithmic, profiles/r1_products_plain_k32.summary.txt).  Here the sources are cut into
// `n_blocks` column blocks whose slice of dXs (+ of the column ids) fits L2, and the grid walks
// block 0 of every row first, then block 1, ...: all CTAs in flight reduce into the same ~64-80 MB.
// One record = one column block of TR consecutive rows (blk_ptr from mk_block_ptr gives every row's
// segment), so a CTA stages TR rows of dY once and works through ~TR * deg / n_blocks stored entries
// -- round 1's per-row blocked records (one row, ~12 entries, dY re-staged each time) were too
// short to pay.  Price: dY is read n_blocks times (2.5 GB each on that shape) against ~25 GB of
// read-modify-write traffic saved.
template <int K, typename IdxT, int TR>
__global__ void __launch_bounds__(32)
sspmm_bwd_tiled_kernel(const int* __restrict__ blk_ptr, const int* __restrict__ idx,
                       const float* __restrict__ val, const float* __restrict__ dy,
                       const IdxT* __restrict__ sp_index, float* __restrict__ dxs, int d, int64_t n_rows,
                       int n_tiles) {
    constexpr int LPN = K / 4;   // lanes per neighbour
    constexpr int G = 32 / LPN;  // neighbours per warp step
    constexpr int U = (K / 4 >= 8) ? 4 : (K / 4);
    extern __shared__ __align__(16) float dys[];  // TR rows of d floats
    const int lane = lane_id();
    const int g = lane / LPN;
    const int t = lane % LPN;
    const int b = blockIdx.x / n_tiles;            // column block: the slow index of the grid
    const int64_t row0 = static_cast<int64_t>(blockIdx.x % n_tiles) * TR;
    const int nr = static_cast<int>(min(static_cast<int64_t>(TR), n_rows - row0));
    const int* __restrict__ seg_lo = blk_ptr + static_cast<int64_t>(b) * n_rows + row0;
    const int* __restrict__ seg_hi = seg_lo + n_rows;

    // segment bounds of the tile's rows (lane i: row i), and whether there is anything to do
    int lo = 0, hi = 0;
    if (lane < nr) {
        lo = __ldg(seg_lo + lane);
        hi = __ldg(seg_hi + lane);
    }
    if (__ballot_sync(kFull, hi > lo) == 0u) return;

    const float* __restrict__ dyr = dy + row0 * d;
    for (int c = lane * 4; c < nr * d; c += 128)
        *reinterpret_cast<float4*>(dys + c) = ld_stream_f4(dyr + c);
    __syncwarp();

    for (int i = 0; i < nr; ++i) {
        const int r_lo = __shfl_sync(kFull, lo, i), r_hi = __shfl_sync(kFull, hi, i);
        const float* __restrict__ dyi = dys + i * d;
        for (int base = r_lo; base < r_hi; base += 32) {
            const int n_here = min(32, r_hi - base);
            int my_nz = 0;
            float my_v = 0.f;
            if (lane < n_here) {
                my_nz = ld_stream_i1(idx + base + lane);
                my_v = ld_stream_f1(val + base + lane);
            }
            for (int j = 0; j < n_here; j += G * U) {
                int cv[U][4];
                int nzv[U];
                float vv[U];
                bool ok[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int e = j + u * G + g;
                    nzv[u] = __shfl_sync(kFull, my_nz, e & 31);
                    vv[u] = __shfl_sync(kFull, my_v, e & 31);
                    ok[u] = e < n_here;
                    if (ok[u]) load_cols4<IdxT>(sp_index + static_cast<int64_t>(nzv[u]) * K + 4 * t, cv[u]);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (ok[u]) {
                        const float v = vv[u];
                        red_add_f4(dxs + static_cast<int64_t>(nzv[u]) * K + 4 * t, v * dyi[cv[u][0]],
                                   v * dyi[cv[u][1]], v * dyi[cv[u][2]], v * dyi[cv[u][3]]);
                    }
                }
            }
        }
    }
}

template <int K, typename IdxT>
static int launch_bwd_tiled(const int* blk_ptr, int n_blocks, const int* idx, const float* val,
                            const float* dy, const void* sp_index, float* dxs, int d, int64_t n_rows,
                            cudaStream_t st) {
    constexpr int TR = 8;
    const int64_t n_tiles = (n_rows + TR - 1) / TR;
    const int64_t grid = n_tiles * n_blocks;
    if (grid > 0x7fffffffLL || n_tiles > 0x7fffffffLL) return MK_EUNSUPPORTED;
    const size_t smem = static_cast<size_t>(TR) * d * 4;
    if (smem > 200 * 1024) return MK_EUNSUPPORTED;
    auto kern = sspmm_bwd_tiled_kernel<K, IdxT, TR>;
    if (smem > 48 * 1024)
        MK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<static_cast<unsigned>(grid), 32, smem, st>>>(blk_ptr, idx, val, dy, static_cast<const IdxT*>(sp_index),
                                                        dxs, d, n_rows, static_cast<int>(n_tiles));
    MK_LAUNCH_CHECK("sspmm_bwd_tiled_kernel");
    return MK_OK;
}

// ---- experimental: part of the reductions through the TMA unit ---------------------------------
// The kernel above is bound by the SM -> L2 request path (l1tex__m_l1tex2xbar_req_cycles_active 91 %,
// profiles/r1_final_*): every stored entry costs k/4 lane-level REDG.128.  Here NT of the G
// neighbours of a warp step take another road: their k contributions are staged in shared memory
// (one STS.128 per lane) and leave as ONE bulk reduction per neighbour,
// `cp.reduce.async.bulk.global.shared::cta.add.f32` (SASS UBLKRED), issued by the first lane of
// the group -- k*4 bytes per request through the TMA unit instead of k/4 requests through L1TEX.
// NBUF staging buffers per group; the issuing lane waits for the bulk group of NBUF steps ago to
// have READ its buffer before the group overwrites it.  NT = G sends everything through TMA,
// smaller NT splits the traffic between the two paths.
// MEASURED SLOWER on a B200 (tools/bwd_tma_probe.py, profiles/peer_r1/bwd_tma_probe.log; Reddit shape,
// k = 32): 2.61 ms shipped kernel, 4.02 / 4.26 / 4.43 ms with NT = 1 / 2 / 4 -- about 11 cycles per
// 128-byte bulk reduction and SM, against 6.6 cycles per stored entry for the eight REDG.128.  Kept
// behind MAXK_BWD_TMA as the record of that experiment; not a product path.
__device__ __forceinline__ void bulk_red_add_f32(float* gdst, const float* ssrc, int bytes) {
    const uint32_t sa = static_cast<uint32_t>(__cvta_generic_to_shared(ssrc));
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst),
                 "r"(sa), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <int K, typename IdxT, int U, int NT>
__global__ void __launch_bounds__(32)
sspmm_bwd_tma_kernel(const mk_part* __restrict__ parts, const int* __restrict__ idx,
                     const float* __restrict__ val, const float* __restrict__ dy,
                     const IdxT* __restrict__ sp_index, float* __restrict__ dxs, int d, int vec_dy,
                     int stage_off) {
    constexpr int LPN = K / 4;   // lanes per neighbour
    constexpr int G = 32 / LPN;  // neighbours per warp step
    constexpr int NBUF = 4;
    static_assert(NT >= 1 && NT <= G, "NT counts neighbours of a step");
    extern __shared__ __align__(16) float dys[];   // dynamic shared memory starts 128-byte aligned
    float* __restrict__ stage = dys + stage_off;  // [NBUF][NT][K]; stage_off is a multiple of 32 floats
    const int lane = lane_id();
    const int g = lane / LPN;
    const int t = lane % LPN;
    const bool via_tma = g < NT;
    const bool issuer = via_tma && t == 0;
    const mk_part rec = parts[blockIdx.x];
    if (rec.len == 0) return;

    const float* __restrict__ dyr = dy + static_cast<int64_t>(rec.row) * d;
    if (vec_dy) {
        for (int c = lane * 4; c < d; c += 128)
            *reinterpret_cast<float4*>(dys + c) = ld_stream_f4(dyr + c);
    } else {
        for (int c = lane; c < d; c += 32) dys[c] = ld_stream_f1(dyr + c);
    }
    __syncwarp();

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
            int cv[U][4];
            int nzv[U];
            float vv[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = i + u * G + g;
                nzv[u] = __shfl_sync(kFull, my_nz, e & 31);
                vv[u] = __shfl_sync(kFull, my_v, e & 31);
                ok[u] = e < n_here;
                if (ok[u]) load_cols4<IdxT>(sp_index + static_cast<int64_t>(nzv[u]) * K + 4 * t, cv[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float v = vv[u];
                float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok[u]) c = make_float4(v * dys[cv[u][0]], v * dys[cv[u][1]], v * dys[cv[u][2]], v * dys[cv[u][3]]);
                float* __restrict__ buf = stage + ((u % NBUF) * NT + g) * K;
                // the bulk reduction that read this buffer NBUF steps ago must be done with it
                if (issuer) bulk_wait_read<NBUF - 1>();
                __syncwarp();
                if (via_tma) {
                    if (ok[u]) {
                        *reinterpret_cast<float4*>(buf + 4 * t) = c;
                        fence_async_smem();
                    }
                } else if (ok[u]) {
                    red_add_f4(dxs + static_cast<int64_t>(nzv[u]) * K + 4 * t, c.x, c.y, c.z, c.w);
                }
                __syncwarp();
                if (issuer) {
                    if (ok[u]) bulk_red_add_f32(dxs + static_cast<int64_t>(nzv[u]) * K, buf, K * 4);
                    bulk_commit();
                }
            }
        }
    }
    if (issuer) bulk_wait_all();
}

template <typename IdxT>
__global__ void __launch_bounds__(32)
sspmm_bwd_generic_kernel(const mk_part* __restrict__ parts, const int* __restrict__ idx,
                         const float* __restrict__ val, const float* __restrict__ dy,
                         const IdxT* __restrict__ sp_index, float* __restrict__ dxs, int k, int d) {
    extern __shared__ __align__(16) float dys[];
    const int lane = lane_id();
    const mk_part rec = parts[blockIdx.x];
    if (rec.len == 0) return;
    const float* __restrict__ dyr = dy + static_cast<int64_t>(rec.row) * d;
    for (int c = lane; c < d; c += 32) dys[c] = dyr[c];
    __syncwarp();
    const int end = rec.loc + rec.len;
    for (int e = rec.loc; e < end; ++e) {
        const int64_t nz = idx[e];
        const float v = val[e];
        for (int q = lane; q < k; q += 32) {
            const int c = static_cast<int>(__ldg(sp_index + nz * k + q));
            red_add_f1(dxs + nz * k + q, v * dys[c]);
        }
    }
}

template <int K, typename IdxT>
static int launch_bwd_k(const mk_part* parts, int64_t num_parts, const int* idx, const float* val,
                        const float* dy, const void* sp_index, float* dxs, int d,
                        cudaStream_t st) {
    // steps of column ids in flight; a 32-entry slice holds 32 / G steps (G = 128 / K neighbours per
    // step), more would only add predicated-off work
#ifdef MK_BWD_U
    constexpr int U = MK_BWD_U;
#else
    constexpr int U = (K / 4 >= 8) ? 8 : (K / 4);
#endif
    const int dpad = (d + 3) & ~3;
#ifdef MK_BWD_EXTRA_SMEM   // measurement knob: FEWER resident CTAs (the request path is the bound, not latency)
    const size_t smem = static_cast<size_t>(dpad) * 4 * MK_BWD_WARPS + MK_BWD_EXTRA_SMEM;
#else
    const size_t smem = static_cast<size_t>(dpad) * 4 * MK_BWD_WARPS;
#endif
    if (smem > 200 * 1024) return MK_EUNSUPPORTED;
    auto kern = K <= 32 ? sspmm_bwd_kernel_occ32<K, IdxT, U> : sspmm_bwd_kernel<K, IdxT, U>;
    if (smem > 48 * 1024)
        MK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    const int vec_dy = (d % 4 == 0) && (reinterpret_cast<uintptr_t>(dy) % 16 == 0);
    kern<<<static_cast<unsigned>((num_parts + MK_BWD_WARPS - 1) / MK_BWD_WARPS), 32 * MK_BWD_WARPS, smem, st>>>(
        parts, num_parts, idx, val, dy, static_cast<const IdxT*>(sp_index), dxs, d, vec_dy);
    MK_LAUNCH_CHECK("sspmm_bwd_kernel");
    return MK_OK;
}

template <typename IdxT>
static int launch_bwd(const mk_part* parts, int64_t num_parts, const int* idx, const float* val,
                      const float* dy, const void* sp_index, float* dxs, int k, int d,
                      cudaStream_t st) {
    const bool aligned = (reinterpret_cast<uintptr_t>(dxs) % 16 == 0) &&
                         (reinterpret_cast<uintptr_t>(sp_index) % (4 * sizeof(IdxT)) == 0);
    if (aligned) {
        switch (k) {
            case 4: return launch_bwd_k<4, IdxT>(parts, num_parts, idx, val, dy, sp_index, dxs, d, st);
            case 8: return launch_bwd_k<8, IdxT>(parts, num_parts, idx, val, dy, sp_index, dxs, d, st);
            case 16: return launch_bwd_k<16, IdxT>(parts, num_parts, idx, val, dy, sp_index, dxs, d, st);
            case 32: return launch_bwd_k<32, IdxT>(parts, num_parts, idx, val, dy, sp_index, dxs, d, st);
            case 64: return launch_bwd_k<64, IdxT>(parts, num_parts, idx, val, dy, sp_index, dxs, d, st);
            case 128: return launch_bwd_k<128, IdxT>(parts, num_parts, idx, val, dy, sp_index, dxs, d, st);
            default: break;
        }
    }
    const size_t smem = static_cast<size_t>(d) * 4;
    if (smem > 200 * 1024) return MK_EUNSUPPORTED;
    if (smem > 48 * 1024)
        MK_CUDA_TRY(cudaFuncSetAttribute(sspmm_bwd_generic_kernel<IdxT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    sspmm_bwd_generic_kernel<IdxT><<<static_cast<unsigned>(num_parts), 32, smem, st>>>(
        parts, idx, val, dy, static_cast<const IdxT*>(sp_index), dxs, k, d);
    MK_LAUNCH_CHECK("sspmm_bwd_generic_kernel");
    return MK_OK;
}

}  // namespace mk

extern "C" int mk_sspmm_bwd(const mk_part* parts, int64_t num_parts, const int32_t* idx,
                            const float* val, const float* dy, const void* sp_index,
                            int index_bytes, float* dxs, int64_t n_rows, int64_t n_src, int k,
                            int d, void* stream) {
    if (n_rows < 0 || n_src < 0 || num_parts < 0 || d < 1 || k < 1 || k > d) return MK_EINVAL;
    if (index_bytes != 1 && index_bytes != 2) return MK_EINVAL;
    if ((index_bytes == 1 && d > 256) || d > 65536) return MK_EINVAL;
    if (n_src == 0) return MK_OK;
    if (!dxs) return MK_EINVAL;
    cudaStream_t st = mk::as_stream(stream);
    MK_CUDA_TRY(cudaMemsetAsync(dxs, 0, static_cast<size_t>(n_src) * k * sizeof(float), st));
    if (n_rows == 0 || num_parts == 0) return MK_OK;
    if (!parts || !dy || !sp_index) return MK_EINVAL;  // idx / val may be NULL: a graph without stored entries
    if (num_parts > 0x7fffffffLL) return MK_EUNSUPPORTED;
    return index_bytes == 1
               ? mk::launch_bwd<uint8_t>(parts, num_parts, idx, val, dy, sp_index, dxs, k, d, st)
               : mk::launch_bwd<uint16_t>(parts, num_parts, idx, val, dy, sp_index, dxs, k, d, st);
}

namespace mk {

template <int K, typename IdxT, int NT>
static int launch_bwd_tma_k(const mk_part* parts, int64_t num_parts, const int* idx, const float* val,
                            const float* dy, const void* sp_index, float* dxs, int d, cudaStream_t st) {
    constexpr int U = (K / 4 >= 8) ? 8 : (K / 4);
    constexpr int G = 128 / K;
    static_assert(U % 4 == 0 || U < 4, "staging buffers are indexed by u % 4");
    const int stage_off = (d + 31) & ~31;  // floats; keeps the staging area 128-byte aligned
    const size_t smem = (static_cast<size_t>(stage_off) + 4 * NT * K) * 4;
    if (smem > 200 * 1024) return MK_EUNSUPPORTED;
    auto kern = sspmm_bwd_tma_kernel<K, IdxT, U, (NT <= G ? NT : G)>;
    if (smem > 48 * 1024)
        MK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    const int vec_dy = (d % 4 == 0) && (reinterpret_cast<uintptr_t>(dy) % 16 == 0);
    kern<<<static_cast<unsigned>(num_parts), 32, smem, st>>>(
        parts, idx, val, dy, static_cast<const IdxT*>(sp_index), dxs, d, vec_dy, stage_off);
    MK_LAUNCH_CHECK("sspmm_bwd_tma_kernel");
    return MK_OK;
}

template <int K, typename IdxT>
static int launch_bwd_tma(const mk_part* parts, int64_t num_parts, const int* idx, const float* val,
                          const float* dy, const void* sp_index, float* dxs, int d, int nt,
                          cudaStream_t st) {
    switch (nt) {
        case 1: return launch_bwd_tma_k<K, IdxT, 1>(parts, num_parts, idx, val, dy, sp_index, dxs, d, st);
        case 2: return launch_bwd_tma_k<K, IdxT, 2>(parts, num_parts, idx, val, dy, sp_index, dxs, d, st);
        case 4: return launch_bwd_tma_k<K, IdxT, 4>(parts, num_parts, idx, val, dy, sp_index, dxs, d, st);
        default: return MK_EUNSUPPORTED;
    }
}

}  // namespace mk

extern "C" int mk_sspmm_bwd_tma(const mk_part* parts, int64_t num_parts, const int32_t* idx,
                                const float* val, const float* dy, const void* sp_index,
                                int index_bytes, float* dxs, int64_t n_rows, int64_t n_src, int k,
                                int d, int tma_neighbours, void* stream) {
    if (n_rows < 0 || n_src < 0 || num_parts < 0 || d < 1 || k < 1 || k > d) return MK_EINVAL;
    if (index_bytes != 1 || d > 256) return MK_EUNSUPPORTED;
    if (k != 32) return MK_EUNSUPPORTED;  // the only width measured on a B200 so far
    if (tma_neighbours != 1 && tma_neighbours != 2 && tma_neighbours != 4) return MK_EUNSUPPORTED;
    if (n_src == 0) return MK_OK;
    if (!dxs || (reinterpret_cast<uintptr_t>(dxs) & 15) || (reinterpret_cast<uintptr_t>(sp_index) & 3))
        return MK_EINVAL;
    cudaStream_t st = mk::as_stream(stream);
    MK_CUDA_TRY(cudaMemsetAsync(dxs, 0, static_cast<size_t>(n_src) * k * sizeof(float), st));
    if (n_rows == 0 || num_parts == 0) return MK_OK;
    if (!parts || !dy || !sp_index) return MK_EINVAL;
    if (num_parts > 0x7fffffffLL) return MK_EUNSUPPORTED;
    return mk::launch_bwd_tma<32, uint8_t>(parts, num_parts, idx, val, dy, sp_index, dxs, d, tma_neighbours, st);
}

extern "C" int mk_sspmm_bwd_tiled(const int32_t* blk_ptr, int n_blocks, const int32_t* idx, const float* val,
                                  const float* dy, const void* sp_index, int index_bytes, float* dxs,
                                  int64_t n_rows, int64_t n_src, int k, int d, void* stream) {
    if (n_rows < 0 || n_src < 0 || n_blocks < 1 || d < 1 || k < 1 || k > d) return MK_EINVAL;
    if (index_bytes != 1 && index_bytes != 2) return MK_EINVAL;
    if ((index_bytes == 1 && d > 256) || d > 65536) return MK_EINVAL;
    if (k != 8 && k != 16 && k != 32 && k != 64) return MK_EUNSUPPORTED;
    if (d % 4 != 0) return MK_EUNSUPPORTED;
    if (n_src == 0) return MK_OK;
    if (!dxs || (reinterpret_cast<uintptr_t>(dxs) & 15)) return MK_EINVAL;
    cudaStream_t st = mk::as_stream(stream);
    MK_CUDA_TRY(cudaMemsetAsync(dxs, 0, static_cast<size_t>(n_src) * k * sizeof(float), st));
    if (n_rows == 0) return MK_OK;
    if (!blk_ptr || !dy || !sp_index || !idx || !val) return MK_EINVAL;
    if ((reinterpret_cast<uintptr_t>(dy) & 15) || (reinterpret_cast<uintptr_t>(sp_index) & (4 * index_bytes - 1)))
        return MK_EINVAL;
#define MK_TILED(KK)                                                                                              \
    return index_bytes == 1                                                                                       \
               ? mk::launch_bwd_tiled<KK, uint8_t>(blk_ptr, n_blocks, idx, val, dy, sp_index, dxs, d, n_rows, st)  \
               : mk::launch_bwd_tiled<KK, uint16_t>(blk_ptr, n_blocks, idx, val, dy, sp_index, dxs, d, n_rows, st)
    switch (k) {
        case 8: MK_TILED(8);
        case 16: MK_TILED(16);
        case 32: MK_TILED(32);
        default: MK_TILED(64);
    }
#undef MK_TILED
}
