// Forward SpGEMM and backward SSpMM on a BANKED CBSR table (see bank.cu for the format).
//
// Same work decomposition as spgemm_fwd.cu / sspmm_bwd.cu -- one warp per work record, one record
// per CTA, 8 lanes per neighbour, 4 neighbours per warp step, CAP = k/8 entries per lane fetched
// with one vector load of values and one of cell offsets -- but every shared-memory access goes to
// the cell `slot + 8*g` chosen by mk_cbsr_bank: group g only ever touches banks [8g, 8g+8), and
// inside a group the 8 lanes of a step hit 8 different banks except for the few duplicates the
// assignment could not avoid.  The L1TEX data pipe (the bound of the unbanked kernels) then sees
// ~1.4 wavefronts per LDS / STS instead of 3.6.
//
// Shared memory per warp: 32 banks x ROWS rows of fp32, ROWS = ceil(D/8) + 8*ceil(D/64)
// (copy A rows, then copy B rows; 8 KB at D = 256).  Forward: the cells are accumulators, folded
// over the 4 groups and the 2 copies when the row is written.  Backward: the cells are replicas
// of dY[r, :], built once per record.
#include "common.cuh"
#include "epilogue.cuh"
#include "peer.cuh"

// Which widths prefetch the epilogue's h_self row at CTA start (-DMK_EPI_PREFETCH_MASK=<or of k> to vary it).
#ifndef MK_EPI_PREFETCH_MASK
#define MK_EPI_PREFETCH_MASK (8 | 32)
#endif
#define MK_EPI_PREFETCH(K) (((MK_EPI_PREFETCH_MASK) & (K)) != 0)
// Which widths run their PLAIN forward on the instantiation without the epilogue's code (see launch_fwd_banked).
#ifndef MK_FWD_PLAIN_BARE_MASK
#define MK_FWD_PLAIN_BARE_MASK 16
#endif

namespace mk {

// Loads of a lane's CAP consecutive entries.  CG = false: read-only path (ld.global.nc), the
// table is complete before the kernel starts.  CG = true: L2-coherent loads (ld.global.cg) for the
// forward that runs while peers are still storing other rows of the table over NVLink (the L1 hit
// rate of these gathers is 0.15 % -- profiles/r1_final_* -- so nothing is lost).
// -DMK_TAB_NOALLOC: the complete-table gathers with L1::no_allocate instead of the default read-only path
// (measurement knob).  Their L1 hit rate is 0.15 %, and still the no-allocate form is much SLOWER: forward at k = 8 /
// 16 / 32 / 64 1.459 / 2.005 / 2.816 / 5.758 ms as shipped, 1.537 / 2.161 / 3.454 / 9.411 ms without allocation
// (profiles/r2/tab_noalloc_call57.log).
template <bool CG>
__device__ __forceinline__ float4 ld_tab_f4(const float* p) {
#ifdef MK_TAB_NOALLOC
    if (!CG) return ld_stream_f4(p);
#endif
    if (!CG) return __ldg(reinterpret_cast<const float4*>(p));
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
template <bool CG>
__device__ __forceinline__ float2 ld_tab_f2(const float* p) {
    if (!CG) return __ldg(reinterpret_cast<const float2*>(p));
    float2 r;
    asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
template <bool CG>
__device__ __forceinline__ float ld_tab_f1(const float* p) {
    if (!CG) return __ldg(p);
    float r;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
template <bool CG>
__device__ __forceinline__ uint4 ld_tab_u4(const void* p) {
#ifdef MK_TAB_NOALLOC
    if (!CG) {
        uint4 r;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
        return r;
    }
#endif
    if (!CG) return __ldg(reinterpret_cast<const uint4*>(p));
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
template <bool CG>
__device__ __forceinline__ uint2 ld_tab_u2(const void* p) {
#ifdef MK_TAB_NOALLOC
    if (!CG) {
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
        return r;
    }
#endif
    if (!CG) return __ldg(reinterpret_cast<const uint2*>(p));
    uint2 r;
    asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
template <bool CG>
__device__ __forceinline__ uint32_t ld_tab_u1(const void* p) {
    if (!CG) return __ldg(reinterpret_cast<const uint32_t*>(p));
    uint32_t r;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
template <bool CG>
__device__ __forceinline__ uint32_t ld_tab_h1(const uint16_t* p) {
    if (!CG) return __ldg(p);
    uint16_t r;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return r;
}

template <int CAP, bool CG = false>
struct BankedLoad;
template <bool CG>
struct BankedLoad<1, CG> {
    static __device__ __forceinline__ void data(const float* p, float (&v)[1]) { v[0] = ld_tab_f1<CG>(p); }
    static __device__ __forceinline__ void slots(const uint16_t* p, int (&s)[1]) { s[0] = ld_tab_h1<CG>(p); }
};
template <bool CG>
struct BankedLoad<2, CG> {
    static __device__ __forceinline__ void data(const float* p, float (&v)[2]) {
        const float2 f = ld_tab_f2<CG>(p);
        v[0] = f.x; v[1] = f.y;
    }
    static __device__ __forceinline__ void slots(const uint16_t* p, int (&s)[2]) {
        const uint32_t q = ld_tab_u1<CG>(p);
        s[0] = q & 0xffff; s[1] = q >> 16;
    }
};
template <bool CG>
struct BankedLoad<4, CG> {
    static __device__ __forceinline__ void data(const float* p, float (&v)[4]) {
        const float4 f = ld_tab_f4<CG>(p);
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    }
    static __device__ __forceinline__ void slots(const uint16_t* p, int (&s)[4]) {
        const uint2 q = ld_tab_u2<CG>(p);
        s[0] = q.x & 0xffff; s[1] = q.x >> 16; s[2] = q.y & 0xffff; s[3] = q.y >> 16;
    }
};
template <bool CG>
struct BankedLoad<8, CG> {
    static __device__ __forceinline__ void data(const float* p, float (&v)[8]) {
        const float4 f = ld_tab_f4<CG>(p);
        const float4 h = ld_tab_f4<CG>(p + 4);
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        v[4] = h.x; v[5] = h.y; v[6] = h.z; v[7] = h.w;
    }
    static __device__ __forceinline__ void slots(const uint16_t* p, int (&s)[8]) {
        const uint4 q = ld_tab_u4<CG>(p);
        s[0] = q.x & 0xffff; s[1] = q.x >> 16; s[2] = q.y & 0xffff; s[3] = q.y >> 16;
        s[4] = q.z & 0xffff; s[5] = q.z >> 16; s[6] = q.w & 0xffff; s[7] = q.w >> 16;
    }
};

// Packed form (k = 8, 16): one entry = {value bits, cell | column << 16} in 8 bytes, a lane's CAP
// entries in ONE load.  At these widths the two-array form pays a full L1 wavefront for a 32-byte
// row of values and another for 8-16 bytes of offsets (the sectors are fetched whole anyway), and the
// forward is bound by exactly those wavefronts; packing halves them without moving more bytes.
template <int CAP, bool CG>
struct PackedLoad;
template <bool CG>
struct PackedLoad<1, CG> {
    static __device__ __forceinline__ void both(const uint2* p, float (&v)[1], int (&s)[1]) {
        const uint2 q = ld_tab_u2<CG>(p);
        v[0] = __uint_as_float(q.x); s[0] = q.y & 0xffff;
    }
};
template <bool CG>
struct PackedLoad<2, CG> {
    static __device__ __forceinline__ void both(const uint2* p, float (&v)[2], int (&s)[2]) {
        const uint4 q = ld_tab_u4<CG>(p);
        v[0] = __uint_as_float(q.x); s[0] = q.y & 0xffff;
        v[1] = __uint_as_float(q.z); s[1] = q.w & 0xffff;
    }
};
template <int CAP, bool CG>
struct PackedLoadAny {  // CAP > 2 is never packed; keeps the template instantiable
    static __device__ __forceinline__ void both(const uint2*, float (&)[CAP], int (&)[CAP]) {}
};
template <bool CG>
struct PackedLoadAny<1, CG> : PackedLoad<1, CG> {};
template <bool CG>
struct PackedLoadAny<2, CG> : PackedLoad<2, CG> {};

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// Forward over a table that may still be ARRIVING (multi-GPU, peer.cuh): rank q's rows of the table
// are complete once done[q] >= the window's epoch.  `hdr` is the OWN window's header.
struct FwdWait {
    const uint32_t* hdr;  // null: the table is complete (single GPU, NCCL form)
    int world, rank;
    int rows_per_rank;    // table rows [q*rows_per_rank, (q+1)*rows_per_rank) come from rank q
    uint64_t timeout_ns;
    // The all-gather itself, fused into this kernel: the first `push.pushers` CTAs TO START copy this
    // rank's rows to the peers (push_rows, peer.cuh).  They wait for nothing, so every rank's rows
    // leave while its other CTAs already work on the blocks that have arrived.  pushers == 0:
    // somebody else moves the rows (mk_peer_push_sm before the kernel, copy engines, NCCL).
    PushDesc push;
};

// One PHASE of a forward that is cut by source block (mk_fwd_phase): the launch only walks the stored
// entries whose column lies in the blocks [a0, a1) and then [b0, b1) (blk[b * stride + row] = position
// in idx of the row's first entry of block b, mk_block_ptr), and either writes the output rows or adds
// to what the earlier phases wrote.  A row-partitioned forward runs the phases in the order the
// peers' rows arrive -- own block, the next few senders, the rest -- so that whole launches, not
// just the CTAs that happen to be resident, overlap the transfer.
struct FwdPhase {
    const int* blk;  // null: no phases (the record's whole range, `split` order)
    int64_t stride;
    int a0, a1, b0, b1;
    int accumulate;  // add to the rows already in `out`
    int last;        // the kernel's completion must mean "the whole table has arrived"
};

// `split` (nullable): per CSR row, the position in idx of the first stored entry whose column is
// >= rank*rows_per_rank.  A record then walks [split, end) first and [begin, split) second, i.e.
// the source blocks in the order rank, rank+1, ..., world-1, 0, ..., rank-1 -- the order in which
// the pushes (peer.cu, push_rows) make them arrive.  The summation order of a row is fixed either way.
// EPI: the f-3 epilogue variant (its own instantiation: the 16 registers of the prefetched h_self row would
// otherwise be carried through the accumulation loop of every launch -- that alone moved the k = 16 forward
// from 2.14 to 2.39 ms, profiles/r2/k16_bisect_call35.log).
// -DMK_FWD_MINBLOCKS=n: a resident-CTA hint for ptxas (measurement knob).  Measured, forward ms at k = 8 / 16 / 32 /
// 64 (profiles/r2/fwd_minblocks_call37.log): no hint 1.521 / 2.138 / 2.870 / 5.979 (shipped); 25: 1.508 / 2.706 / 2.872 /
// 6.053; 32: 1.509 / 2.537 / 2.884 / 6.098; 1 or 20: 1.554 / 2.401 / 2.967 / 6.192.
#ifdef MK_FWD_MINBLOCKS
#define MK_FWD_BOUNDS __launch_bounds__(32, MK_FWD_MINBLOCKS)
#else
#define MK_FWD_BOUNDS __launch_bounds__(32)
#endif
template <int K, int U, bool WAIT, bool PACKED = false, bool EPI = false, bool PF = false>
__global__ void MK_FWD_BOUNDS
spgemm_fwd_banked_kernel(const mk_part* __restrict__ parts, const int* __restrict__ idx,
                         const float* __restrict__ val, const float* __restrict__ bk_data,
                         const uint16_t* __restrict__ bk_slot, float* __restrict__ out,
                         float* __restrict__ partial, int d, int rows, const int* __restrict__ split,
                         const FwdWait fw, const FwdPhase ph, const FwdEpilogue ep) {
    constexpr int CAP = K / 8;
    extern __shared__ __align__(16) float acc[];  // 32 * rows
    const int lane = lane_id();
    const int g = lane >> 3;
    const int t = lane & 7;
    // Role of this CTA.  With pushers, roles are handed out in the order the CTAs START (a ticket),
    // not by block index: the first `pushers` CTAs that run move the rows, whatever order the
    // hardware dispatches blocks in -- a waiting consumer can never keep a pusher off the machine.
    int role = static_cast<int>(blockIdx.x);
    uint32_t epoch = 0;
    if (WAIT) {
        epoch = fw.hdr[kHdrEpoch];
        if (fw.push.pushers > 0) {
            if (lane == 0) role = static_cast<int>(atomicAdd(const_cast<uint32_t*>(fw.hdr) + kHdrTicket, 1u));
            role = __shfl_sync(kFull, role, 0);
            if (role < fw.push.pushers) {
                push_rows(fw.push, role, epoch);
                return;
            }
            role -= fw.push.pushers;
        }
    }
    const mk_part rec = parts[role];

    // f-3 epilogue: this row's h_self, asked for now and consumed when the row is finished (the load's
    // latency would otherwise sit at the end of every CTA)
    // (EPI: epilogue code compiled in, still switched by ep.gamma at run time; PF: with the prefetch.  Measured per
    // width, profiles/r2/epi_prefetch_call36.log: the prefetch pays at k = 8 and 32 -- 1.593 -> 1.550, 2.933 -> 2.899 ms
    // -- and costs at k = 16, whose 8-neighbour step already fills the register file, and at k = 64 -- 2.216 -> 2.537,
    // 5.923 -> 5.972 ms; there the row is loaded when it is needed.)
    constexpr bool PREFETCH_SELF = EPI && PF;
    [[maybe_unused]] float4 hs[PREFETCH_SELF ? 4 : 1];
    if constexpr (PREFETCH_SELF) {
        if (ep.gamma != nullptr && ep.h_self != nullptr && rec.slot < 0) {
            const float* __restrict__ hrow = ep.h_self + static_cast<int64_t>(rec.row) * d;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j * 128 + lane * 4 < d) hs[j] = ld_stream_f4(hrow + j * 128 + lane * 4);
        }
    }

    // which source blocks have arrived (bit q); refreshed only while something is missing
    unsigned have = 0xffffffffu;
    if (WAIT) {
        const bool there = lane >= fw.world ||
                           static_cast<int32_t>(ld_acquire_sys(fw.hdr + kHdrDone + lane) - epoch) >= 0;
        have = __ballot_sync(kFull, there);
    }

    for (int c = lane * 4; c < 32 * rows; c += 128)
        *reinterpret_cast<float4*>(acc + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    float* __restrict__ my = acc + 8 * g;
    [[maybe_unused]] const unsigned gmask = 0xffu << (8 * g);  // the lanes that share this group's banks
    const int lo = rec.loc, hi = rec.loc + rec.len;
    int r0b, r0e, r1b, r1e;  // the two ranges of idx this launch walks, in this order
    if (ph.blk != nullptr) {
        const int* __restrict__ bp = ph.blk + rec.row;
        r0b = min(max(__ldg(bp + ph.a0 * ph.stride), lo), hi);
        r0e = min(max(__ldg(bp + ph.a1 * ph.stride), lo), hi);
        r1b = r1e = lo;
        if (ph.b1 > ph.b0) {
            r1b = min(max(__ldg(bp + ph.b0 * ph.stride), lo), hi);
            r1e = min(max(__ldg(bp + ph.b1 * ph.stride), lo), hi);
        }
    } else {
        int sp = lo;
        if (split != nullptr) sp = min(max(__ldg(split + rec.row), lo), hi);
        r0b = sp; r0e = hi; r1b = lo; r1e = sp;
    }
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int b0 = pass == 0 ? r0b : r1b;
        const int end = pass == 0 ? r0e : r1e;
        for (int base = b0; base < end; base += 32) {
            const int n_here = min(32, end - base);
            int my_nz = 0;
            float my_v = 0.f;
            if (lane < n_here) {
                my_nz = ld_stream_i1(idx + base + lane);
                my_v = ld_stream_f1(val + base + lane);
            }
            if (WAIT && have != 0xffffffffu) {  // uniform; false for every CTA once all rows are in
                int blk = __float2int_rd(__fdividef(static_cast<float>(my_nz), static_cast<float>(fw.rows_per_rank)));
                blk += (my_nz >= (blk + 1) * fw.rows_per_rank) ? 1 : 0;
                blk -= (my_nz < blk * fw.rows_per_rank) ? 1 : 0;
                const unsigned need = __reduce_or_sync(kFull, lane < n_here ? (1u << blk) : 0u);
                const unsigned missing = need & ~have;
                if (missing) {
                    if ((missing >> lane) & 1u)
                        wait_flag(fw.hdr + kHdrDone + lane, epoch,
                                  const_cast<uint32_t*>(fw.hdr) + kHdrError, fw.timeout_ns);
                    __syncwarp();
                    have |= missing;
                }
            }
            for (int i = 0; i < n_here; i += 4 * U) {
                float dv[U][CAP];
                int sl[U][CAP];
                float vv[U];
                bool ok[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int e = i + u * 4 + g;
                    const int nz = __shfl_sync(kFull, my_nz, e & 31);
                    vv[u] = __shfl_sync(kFull, my_v, e & 31);
                    ok[u] = e < n_here;
                    if (ok[u]) {
                        const int64_t off = static_cast<int64_t>(nz) * K + CAP * t;
                        if (PACKED) {  // bk_data is the packed table
                            PackedLoadAny<CAP, WAIT>::both(reinterpret_cast<const uint2*>(bk_data) + off, dv[u], sl[u]);
                        } else {
                            BankedLoad<CAP, WAIT>::data(bk_data + off, dv[u]);
                            BankedLoad<CAP, WAIT>::slots(bk_slot + off, sl[u]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (ok[u]) {  // uniform over the 8 lanes of a group
#pragma unroll
                        for (int q = 0; q < CAP; ++q)
                            if (dv[u][q] != 0.0f) my[sl[u][q]] += vv[u] * dv[u][q];
                        accum_fence_group(gmask);
                    }
                    accum_fence_warp();
                }
            }
        }
    }
    __syncwarp();
    // the kernel's completion must mean "the whole table has arrived" (the backward reads the
    // gathered column ids after it): one CTA waits for every sender
    if (WAIT && ph.last && role == 0 && lane < fw.world)
        wait_flag(fw.hdr + kHdrDone + lane, epoch, const_cast<uint32_t*>(fw.hdr) + kHdrError, fw.timeout_ns);

    // ---- fold the 4 groups and the 2 copies.  Lane (s, b) sums, for row 4i+s, the cells of bank
    //      b of every group, visiting the groups in the order (j+s)&3 so that the four rows read
    //      by one instruction sit in four different bank octets.  Copy-A rows produce 8
    //      consecutive columns each; their sums are parked in acc[0..D) (rows already consumed);
    //      copy-B rows are then added on top.
    const int ra = (d + 7) >> 3;
    const int s = lane >> 3, b = lane & 7;
    for (int row0 = 0; row0 < ra; row0 += 4) {
        const int row = row0 + s;
        float sum = 0.f;
        if (row < ra) {
#pragma unroll
            for (int j = 0; j < 4; ++j) sum += acc[32 * row + 8 * ((j + s) & 3) + b];
        }
        __syncwarp();
        if (row < ra) acc[8 * row + b] = sum;  // column 8*row + b
    }
    __syncwarp();
    for (int row0 = ra; row0 < rows; row0 += 4) {
        const int row = row0 + s;
        if (row < rows) {
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) sum += acc[32 * row + 8 * ((j + s) & 3) + b];
            const int rb = row - ra;
            const int c = ((rb >> 3) << 6) + (b << 3) + (rb & 7);
            if (c < d) acc[c] += sum;
        }
    }
    __syncwarp();
    float* __restrict__ o = rec.slot < 0 ? out + static_cast<int64_t>(rec.row) * d
                                         : partial + static_cast<int64_t>(rec.slot) * d;
    if constexpr (EPI) {
        if (ep.gamma != nullptr && rec.slot < 0) {
            // f-3: the finished row goes straight into y = LayerNorm(h_self + row + bias) * gamma + beta
            // (and z, mean, rstd for the backward) -- no dense round trip of the aggregated row
            if constexpr (PREFETCH_SELF) {
                fwd_epilogue_row(ep, [&](int c) { return *reinterpret_cast<const float4*>(acc + c); },
                                 [&](int c) { return hs[c >> 7]; }, out, rec.row, d, lane);
            } else {
                const float* __restrict__ hrow = ep.h_self + static_cast<int64_t>(rec.row) * d;
                fwd_epilogue_row(ep, [&](int c) { return *reinterpret_cast<const float4*>(acc + c); },
                                 [&](int c) { return ld_stream_f4(hrow + c); }, out, rec.row, d, lane);
            }
            return;
        }
    }
    if (ph.accumulate && rec.slot < 0) {  // later phase: on top of what the earlier ones wrote
        for (int c = lane * 4; c < d; c += 128) {
            float4 a = *reinterpret_cast<const float4*>(acc + c);
            const float4 p = *reinterpret_cast<const float4*>(o + c);
            a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
            st_stream_f4(o + c, a);
        }
        return;
    }
    for (int c = lane * 4; c < d; c += 128)
        st_stream_f4(o + c, *reinterpret_cast<const float4*>(acc + c));
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
template <int CAP>
__device__ __forceinline__ void red_add_cap(float* p, const float (&g)[CAP]);
template <>
__device__ __forceinline__ void red_add_cap<1>(float* p, const float (&g)[1]) { red_add_f1(p, g[0]); }
template <>
__device__ __forceinline__ void red_add_cap<2>(float* p, const float (&g)[2]) { red_add_f2(p, g[0], g[1]); }
template <>
__device__ __forceinline__ void red_add_cap<4>(float* p, const float (&g)[4]) {
    red_add_f4(p, g[0], g[1], g[2], g[3]);
}
template <>
__device__ __forceinline__ void red_add_cap<8>(float* p, const float (&g)[8]) {
    red_add_f4(p, g[0], g[1], g[2], g[3]);
    red_add_f4(p + 4, g[4], g[5], g[6], g[7]);
}

template <int K, int U>
__global__ void __launch_bounds__(32)
sspmm_bwd_banked_kernel(const mk_part* __restrict__ parts, const int* __restrict__ idx,
                        const float* __restrict__ val, const float* __restrict__ dy,
                        const uint16_t* __restrict__ bk_slot, float* __restrict__ dxs, int d,
                        int rows) {
    constexpr int CAP = K / 8;
    extern __shared__ __align__(16) float sm[];  // 32*rows replicated cells, then d floats of dY
    float* __restrict__ cells = sm;
    float* __restrict__ tmp = sm + 32 * rows;
    const int lane = lane_id();
    const int g = lane >> 3;
    const int t = lane & 7;
    const mk_part rec = parts[blockIdx.x];
    if (rec.len == 0) return;

    const float* __restrict__ dyr = dy + static_cast<int64_t>(rec.row) * d;
    for (int c = lane * 4; c < d; c += 128)
        *reinterpret_cast<float4*>(tmp + c) = ld_stream_f4(dyr + c);
    __syncwarp();
    const int ra = (d + 7) >> 3;
    for (int row = 0; row < ra; ++row)  // copy A: column 8*row + t in bank t of every group
        cells[32 * row + lane] = tmp[8 * row + t];
    for (int row = ra; row < rows; ++row) {  // copy B: column 64h + 8*t + l in bank t
        const int rb = row - ra;
        const int c = ((rb >> 3) << 6) + (t << 3) + (rb & 7);
        cells[32 * row + lane] = c < d ? tmp[c] : 0.f;
    }
    __syncwarp();

    const float* __restrict__ my = cells + 8 * g;
    const int end = rec.loc + rec.len;
    for (int base = rec.loc; base < end; base += 32) {
        const int n_here = min(32, end - base);
        int my_nz = 0;
        float my_v = 0.f;
        if (lane < n_here) {
            my_nz = ld_stream_i1(idx + base + lane);
            my_v = ld_stream_f1(val + base + lane);
        }
        for (int i = 0; i < n_here; i += 4 * U) {
            int sl[U][CAP];
            int nzv[U];
            float vv[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = i + u * 4 + g;
                nzv[u] = __shfl_sync(kFull, my_nz, e & 31);
                vv[u] = __shfl_sync(kFull, my_v, e & 31);
                ok[u] = e < n_here;
                if (ok[u])
                    BankedLoad<CAP>::slots(bk_slot + static_cast<int64_t>(nzv[u]) * K + CAP * t, sl[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (ok[u]) {
                    float gq[CAP];
#pragma unroll
                    for (int q = 0; q < CAP; ++q) gq[q] = vv[u] * my[sl[u][q]];
                    red_add_cap<CAP>(dxs + static_cast<int64_t>(nzv[u]) * K + CAP * t, gq);
                }
            }
        }
    }
}

template <int K, bool PACKED>
static int launch_fwd_banked(const mk_part* parts, int64_t num_parts, const int* idx,
                             const float* val, const float* bk_data, const uint16_t* bk_slot,
                             float* out, float* partial, int d, int rows, const int* split,
                             const FwdWait& fw, const FwdPhase& ph, const FwdEpilogue& ep, cudaStream_t st) {
#ifdef MK_FWD_U
    constexpr int U = MK_FWD_U;
#else
    // neighbour steps in flight, measured per width (profiles/r2/fwd_unroll_call39.log; forward ms at U = 2 / 4 / 8:
    // k = 8: 1.894 / 1.581 / 1.493, k = 16: 2.339 / 2.042 / 2.138, k = 32: 2.977 / 2.857 / 3.867, k = 64: 5.806 / 6.222 / 11.7)
    constexpr int U = K >= 64 ? 2 : (K >= 16 ? 4 : 8);
#endif
#ifdef MK_FWD_EXTRA_SMEM   // measurement knob: what fewer resident CTAs (a third copy of the cells) would cost
    const size_t smem = static_cast<size_t>(32) * rows * 4 + MK_FWD_EXTRA_SMEM;
#else
    const size_t smem = static_cast<size_t>(32) * rows * 4;
#endif
    if (fw.hdr != nullptr) {
        // (the multi-GPU form keeps the instantiation its 2- to 8-GPU measurements were taken with)
        auto kern = spgemm_fwd_banked_kernel<K, U, true, PACKED, true, true>;
        if (smem > 48 * 1024)
            MK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem)));
        const int64_t grid = num_parts + fw.push.pushers;
        if (grid > 0x7fffffffLL) return MK_EUNSUPPORTED;
        kern<<<static_cast<unsigned>(grid), 32, smem, st>>>(parts, idx, val, bk_data, bk_slot, out,
                                                            partial, d, rows, split, fw, ph, ep);
    } else {
        // Which instantiation runs.  With the epilogue: prefetch where it pays (above).  Without: the kernel
        // that carries the epilogue's code and registers is, as ptxas schedules it, the faster PLAIN forward at
        // k = 8 / 32 / 64 (1.495 / 2.856 / 5.82 ms against 1.521 / 2.870 / 5.98) and the slower one at k = 16 (2.39
        // against 2.14 ms: profiles/r2/k16_bisect_call35.log, epi_prefetch_call36.log) -- so each width gets its own.
        auto go = [&](auto kern) -> int {
            if (smem > 48 * 1024)
                MK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(smem)));
            kern<<<static_cast<unsigned>(num_parts), 32, smem, st>>>(parts, idx, val, bk_data, bk_slot, out,
                                                                     partial, d, rows, split, fw, ph, ep);
            return MK_OK;
        };
        int rc;
        if (ep.gamma != nullptr) {
            rc = go(spgemm_fwd_banked_kernel<K, U, false, PACKED, true, MK_EPI_PREFETCH(K)>);
        } else {
            if constexpr ((MK_FWD_PLAIN_BARE_MASK & K) != 0)
                rc = go(spgemm_fwd_banked_kernel<K, U, false, PACKED, false, false>);
            else
                rc = go(spgemm_fwd_banked_kernel<K, U, false, PACKED, true, true>);
        }
        if (rc != MK_OK) return rc;
    }
    MK_LAUNCH_CHECK("spgemm_fwd_banked_kernel");
    return MK_OK;
}

template <int K>
static int launch_bwd_banked(const mk_part* parts, int64_t num_parts, const int* idx,
                             const float* val, const float* dy, const uint16_t* bk_slot, float* dxs,
                             int d, int rows, cudaStream_t st) {
    constexpr int U = K >= 64 ? 4 : 8;
    const size_t smem = (static_cast<size_t>(32) * rows + d) * 4;
    auto kern = sspmm_bwd_banked_kernel<K, U>;
    if (smem > 48 * 1024)
        MK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    kern<<<static_cast<unsigned>(num_parts), 32, smem, st>>>(parts, idx, val, dy, bk_slot, dxs, d,
                                                             rows);
    MK_LAUNCH_CHECK("sspmm_bwd_banked_kernel");
    return MK_OK;
}

int launch_fold(const mk_part* parts, int64_t num_parts, const float* partial, float* out, int d,
                cudaStream_t st, int accumulate = 0, const FwdEpilogue* ep = nullptr);  // spgemm_fwd.cu

}  // namespace mk

extern "C" int mk_banked_supported(int k, int d);
extern "C" int mk_banked_rows(int d);
extern "C" int mk_peer_wait_all(void* window, int world, int timeout_ms, void* stream);

static int fwd_banked_any(bool packed, const mk_part* parts, int64_t num_parts, int64_t num_slots,
                         const mk_part* exec_parts, const int32_t* idx, const float* val,
                         const float* bk_data, const uint16_t* bk_slot, float* out, float* partial,
                         int64_t n_rows, int k, int d, const int32_t* split, const mk_fwd_exchange* x,
                         const mk_fwd_phase* phase, const mk_fwd_epilogue* epi, void* stream) {
    if (n_rows < 0 || num_parts < 0 || num_slots < 0 || d < 1 || k < 1 || k > d) return MK_EINVAL;
    mk::FwdEpilogue ep{};
    if (epi != nullptr && epi->gamma != nullptr) {
        if (phase != nullptr || x != nullptr || !epi->beta || d % 4 != 0 || d > 512) return MK_EUNSUPPORTED;
        const uintptr_t al = reinterpret_cast<uintptr_t>(epi->h_self) | reinterpret_cast<uintptr_t>(epi->bias) |
                             reinterpret_cast<uintptr_t>(epi->gamma) | reinterpret_cast<uintptr_t>(epi->beta) |
                             reinterpret_cast<uintptr_t>(epi->z);
        if (al & 15) return MK_EINVAL;
        ep.h_self = epi->h_self; ep.bias = epi->bias; ep.gamma = epi->gamma; ep.beta = epi->beta;
        ep.z = epi->z; ep.mean = epi->mean; ep.rstd = epi->rstd; ep.eps = epi->eps;
    }
    mk::FwdPhase ph{};
    ph.last = 1;
    if (phase != nullptr && phase->blk_ptr != nullptr) {
        if (phase->n_blocks < 1 || phase->a0 < 0 || phase->a1 < phase->a0 || phase->a1 > phase->n_blocks ||
            phase->b0 < 0 || phase->b1 < phase->b0 || phase->b1 > phase->n_blocks || phase->row_stride < n_rows)
            return MK_EINVAL;
        ph.blk = phase->blk_ptr;
        ph.stride = phase->row_stride;
        ph.a0 = phase->a0; ph.a1 = phase->a1; ph.b0 = phase->b0; ph.b1 = phase->b1;
        ph.accumulate = phase->accumulate ? 1 : 0;
        ph.last = phase->last ? 1 : 0;
    }
    if (!mk_banked_supported(k, d) || (packed && k > 16)) return MK_EUNSUPPORTED;
    const bool waiting = x != nullptr && x->window != nullptr;
    if (waiting) {
        if (x->world < 1 || x->world > mk::kMaxPeers || x->rank < 0 || x->rank >= x->world ||
            x->rows_per_rank < 1 || x->rows_per_rank > 0x7fffffffLL)
            return MK_EINVAL;
        if (x->h_windows != nullptr) {
            if (x->n_seg < 1 || x->n_seg > 3 || !x->h_offsets || !x->h_bytes || x->pushers < 1 || x->pushers > 65535)
                return MK_EINVAL;
            for (int g = 0; g < x->n_seg; ++g)
                if (x->h_bytes[g] < 0 || (x->h_bytes[g] & 15) || x->h_offsets[g] < MK_PEER_HEADER_BYTES ||
                    (x->h_offsets[g] & 15))
                    return MK_EINVAL;
            for (int q = 0; q < x->world; ++q)
                if (!x->h_windows[q] || (reinterpret_cast<uintptr_t>(x->h_windows[q]) & 15)) return MK_EINVAL;
            if (x->h_windows[x->rank] != x->window) return MK_EINVAL;
        }
    }
    const bool pushing = waiting && x->h_windows != nullptr && x->world > 1;
    if ((n_rows == 0 || num_parts == 0) && !pushing) {
        // nothing to compute, but the collective's contract stands: return once the table is complete
        if (waiting && ph.last) return mk_peer_wait_all(const_cast<void*>(x->window), x->world, x->timeout_ms, stream);
        return MK_OK;
    }
    if (!parts || !out || !bk_data || (!packed && !bk_slot)) return MK_EINVAL;
    if (num_slots > 0 && !partial) return MK_EINVAL;
    if (num_parts > 0x7fffffffLL) return MK_EUNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(out) % 16 || reinterpret_cast<uintptr_t>(bk_data) % 16 ||
        reinterpret_cast<uintptr_t>(bk_slot) % 16 || (partial && reinterpret_cast<uintptr_t>(partial) % 16))
        return MK_EINVAL;
    cudaStream_t st = mk::as_stream(stream);
    mk::FwdWait fw{};
    if (waiting) {
        fw.hdr = static_cast<const uint32_t*>(x->window);
        fw.world = x->world;
        fw.rank = x->rank;
        fw.rows_per_rank = static_cast<int>(x->rows_per_rank);
        fw.timeout_ns = static_cast<uint64_t>(x->timeout_ms > 0 ? x->timeout_ms : 30000) * 1000000ull;
        if (pushing) {
            for (int q = 0; q < x->world; ++q) fw.push.ps.win[q] = static_cast<unsigned char*>(x->h_windows[q]);
            fw.push.world = x->world;
            fw.push.rank = x->rank;
            fw.push.nseg = x->n_seg;
            for (int g = 0; g < x->n_seg; ++g) {
                fw.push.off[g] = x->h_offsets[g];
                fw.push.bytes[g] = x->h_bytes[g];
            }
            fw.push.pushers = x->pushers;
        }
    }
    const mk_part* ex = exec_parts ? exec_parts : parts;
    const int rows = mk_banked_rows(d);
    int rc;
    if (packed) {
        rc = k == 8 ? mk::launch_fwd_banked<8, true>(ex, num_parts, idx, val, bk_data, nullptr, out, partial, d, rows, split, fw, ph, ep, st)
                    : mk::launch_fwd_banked<16, true>(ex, num_parts, idx, val, bk_data, nullptr, out, partial, d, rows, split, fw, ph, ep, st);
    } else {
        switch (k) {
            case 8: rc = mk::launch_fwd_banked<8, false>(ex, num_parts, idx, val, bk_data, bk_slot, out, partial, d, rows, split, fw, ph, ep, st); break;
            case 16: rc = mk::launch_fwd_banked<16, false>(ex, num_parts, idx, val, bk_data, bk_slot, out, partial, d, rows, split, fw, ph, ep, st); break;
            case 32: rc = mk::launch_fwd_banked<32, false>(ex, num_parts, idx, val, bk_data, bk_slot, out, partial, d, rows, split, fw, ph, ep, st); break;
            default: rc = mk::launch_fwd_banked<64, false>(ex, num_parts, idx, val, bk_data, bk_slot, out, partial, d, rows, split, fw, ph, ep, st); break;
        }
    }
    if (rc != MK_OK) return rc;
    if (num_slots > 0) return mk::launch_fold(parts, num_parts, partial, out, d, st, ph.accumulate, ep.gamma ? &ep : nullptr);
    return MK_OK;
}

extern "C" int mk_spgemm_fwd_banked_ex(const mk_part* parts, int64_t num_parts, int64_t num_slots,
                                       const mk_part* exec_parts, const int32_t* idx, const float* val,
                                       const float* bk_data, const uint16_t* bk_slot, float* out,
                                       float* partial, int64_t n_rows, int k, int d,
                                       const int32_t* split, const mk_fwd_exchange* xchg, void* stream) {
    return fwd_banked_any(false, parts, num_parts, num_slots, exec_parts, idx, val, bk_data, bk_slot, out,
                          partial, n_rows, k, d, split, xchg, nullptr, nullptr, stream);
}

extern "C" int mk_spgemm_fwd_banked_phase(const mk_part* parts, int64_t num_parts, int64_t num_slots,
                                          const mk_part* exec_parts, const int32_t* idx, const float* val,
                                          const float* bk_data, const uint16_t* bk_slot, float* out,
                                          float* partial, int64_t n_rows, int k, int d,
                                          const mk_fwd_exchange* xchg, const mk_fwd_phase* phase,
                                          void* stream) {
    return fwd_banked_any(false, parts, num_parts, num_slots, exec_parts, idx, val, bk_data, bk_slot, out,
                          partial, n_rows, k, d, nullptr, xchg, phase, nullptr, stream);
}

extern "C" int mk_spgemm_fwd_packed_ex(const mk_part* parts, int64_t num_parts, int64_t num_slots,
                                       const mk_part* exec_parts, const int32_t* idx, const float* val,
                                       const void* bk_pack, float* out, float* partial, int64_t n_rows,
                                       int k, int d, const int32_t* split, const mk_fwd_exchange* xchg,
                                       void* stream) {
    return fwd_banked_any(true, parts, num_parts, num_slots, exec_parts, idx, val,
                          static_cast<const float*>(bk_pack), nullptr, out, partial, n_rows, k, d, split,
                          xchg, nullptr, nullptr, stream);
}

extern "C" int mk_spgemm_fwd_banked_ln(const mk_part* parts, int64_t num_parts, int64_t num_slots,
                                       const mk_part* exec_parts, const int32_t* idx, const float* val,
                                       const void* table, const uint16_t* bk_slot, float* y, float* partial,
                                       int64_t n_rows, int k, int d, const mk_fwd_epilogue* epilogue,
                                       void* stream) {
    if (epilogue == nullptr || epilogue->gamma == nullptr) return MK_EINVAL;
    return fwd_banked_any(bk_slot == nullptr, parts, num_parts, num_slots, exec_parts, idx, val,
                          static_cast<const float*>(table), bk_slot, y, partial, n_rows, k, d, nullptr, nullptr,
                          nullptr, epilogue, stream);
}

extern "C" int mk_spgemm_fwd_banked(const mk_part* parts, int64_t num_parts, int64_t num_slots,
                                    const int32_t* idx, const float* val, const float* bk_data,
                                    const uint16_t* bk_slot, float* out, float* partial,
                                    int64_t n_rows, int k, int d, void* stream) {
    return mk_spgemm_fwd_banked_ex(parts, num_parts, num_slots, nullptr, idx, val, bk_data, bk_slot, out,
                                   partial, n_rows, k, d, nullptr, nullptr, stream);
}

extern "C" int mk_sspmm_bwd_banked(const mk_part* parts, int64_t num_parts, const int32_t* idx,
                                   const float* val, const float* dy, const uint16_t* bk_slot,
                                   float* dxs, int64_t n_rows, int64_t n_src, int k, int d,
                                   void* stream) {
    if (n_rows < 0 || n_src < 0 || num_parts < 0 || d < 1 || k < 1 || k > d) return MK_EINVAL;
    if (!mk_banked_supported(k, d)) return MK_EUNSUPPORTED;
    if (n_src == 0) return MK_OK;
    if (!dxs) return MK_EINVAL;
    cudaStream_t st = mk::as_stream(stream);
    MK_CUDA_TRY(cudaMemsetAsync(dxs, 0, static_cast<size_t>(n_src) * k * sizeof(float), st));
    if (n_rows == 0 || num_parts == 0) return MK_OK;
    if (!parts || !dy || !bk_slot) return MK_EINVAL;
    if (num_parts > 0x7fffffffLL) return MK_EUNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(dxs) % 16 || reinterpret_cast<uintptr_t>(dy) % 16 ||
        reinterpret_cast<uintptr_t>(bk_slot) % 16)
        return MK_EINVAL;
    const int rows = mk_banked_rows(d);
    switch (k) {
        case 8: return mk::launch_bwd_banked<8>(parts, num_parts, idx, val, dy, bk_slot, dxs, d, rows, st);
        case 16: return mk::launch_bwd_banked<16>(parts, num_parts, idx, val, dy, bk_slot, dxs, d, rows, st);
        case 32: return mk::launch_bwd_banked<32>(parts, num_parts, idx, val, dy, bk_slot, dxs, d, rows, st);
        default: return mk::launch_bwd_banked<64>(parts, num_parts, idx, val, dy, bk_slot, dxs, d, rows, st);
    }
}
