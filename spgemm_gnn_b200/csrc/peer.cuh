// Device side of the peer-memory exchange: windows, flags, epochs.
//
// A WINDOW is one cudaMalloc'ed buffer per rank, same size everywhere, mapped into every peer
// process with CUDA IPC (mk_peer_export / mk_peer_open), so that a kernel on rank p can store to
// and load from rank q's copy over NVLink.  It starts with a 1 KB header of 32-bit words:
//
//     word 0        epoch     collectives completed through this window (written by the owner)
//     word 1        ticket    CTAs of the running collective that have finished their part
//     word 2        error     collective number a kernel gave up in (bounded wait), 0 = none
//     word 8 + s    step[s]   CTAs that have finished step s of a progressive push (owner-local)
//     word 32 + q   ready[q]  written BY rank q: "for collective #e you may touch my copy"
//     word 64 + q   done[q]   written BY rank q: "for collective #e my part in your copy is over"
//
// and the payload from byte 1024 on.  Every collective through a window runs the same protocol
// (all ranks launch the same collectives in the same order on their streams):
//
//   1. e = epoch + 1.  Block 0 stores ready = e into every peer's header: everything this rank
//      enqueued before the collective has completed (stream order), so its copy may be written
//      (all-gather) / read (reduce-scatter) by the peers.
//   2. A block waits for ready[q] >= e before it touches rank q's copy.
//   3. COMPLETE form (mk_peer_reduce_scatter, mk_peer_reduce_scatter_mc): after its part a block fences
//      (system scope) and takes a ticket; the block that takes the last one stores done = e into
//      every peer's header, waits for done[q] >= e from every peer, and publishes epoch = e.  When
//      the kernel has finished, this rank's copy is complete / no longer read by anybody.
//      PROGRESSIVE form (push_rows: mk_peer_push_sm, pusher CTAs of the forward): the sender visits the destinations one after the other
//      (rank, rank-1, rank-2, ...); the block that finishes step s last stores done = e into THAT
//      destination's header, nobody waits for the peers, and the last block publishes epoch = e.
//      The consumer (the forward SpGEMM, banked.cu) checks done[q] before it reads rank q's rows and
//      walks every CSR row in arrival order, so the transfer overlaps the kernel that needs it.
//
// Every block of a collective kernel may spin on a flag that block 0 of a PEER's kernel writes, so
// the grids are sized to be co-resident (occupancy query, grid-stride loops): a waiting block can
// never keep block 0 of its own grid from being scheduled.
//
// The epoch lives in device memory, not in a kernel argument, so a captured CUDA graph replays
// correctly.  Waits are bounded: after `timeout_ns` without progress the waiter writes the
// collective's number into the error word of its own header and stops waiting (results are then
// garbage, the host finds the error word through mk_peer_epoch and raises -- peer.check_errors()).
#pragma once

#include "common.cuh"

namespace mk {

constexpr int kMaxPeers = MK_PEER_MAX_RANKS;
constexpr int kHdrEpoch = 0, kHdrTicket = 1, kHdrError = 2, kHdrStep = 8, kHdrReady = 32, kHdrDone = 64;
constexpr int kHdrBytes = MK_PEER_HEADER_BYTES;
static_assert(kHdrDone + kMaxPeers <= kHdrBytes / 4, "header too small");
static_assert(kHdrStep + kMaxPeers <= kHdrReady, "step counters overlap the flags");

struct PeerSet {
    unsigned char* win[kMaxPeers];  // base of rank q's window in THIS process's address space
};

// Which rank / block of the collective a CTA is.  Production kernels: (rank argument, blockIdx.x,
// gridDim.x).  The single-GPU emulation used by the tests runs ALL ranks in one launch
// (blockIdx.y = rank), because separate launches that wait on one another are not guaranteed to
// run at the same time on one device.
struct PeerCtx {
    int rank, bid, nblk;
};
__device__ __forceinline__ PeerCtx peer_ctx(int rank_arg, bool virt) {
    PeerCtx c;
    c.rank = virt ? static_cast<int>(blockIdx.y) : rank_arg;
    c.bid = static_cast<int>(blockIdx.x);
    c.nblk = static_cast<int>(gridDim.x);
    return c;
}

__device__ __forceinline__ uint32_t* peer_hdr(const PeerSet& ps, int q) {
    return reinterpret_cast<uint32_t*>(ps.win[q]);
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Spin until *p >= e (wrap-safe).  After timeout_ns: error word = e, give up (returns false).
// A kernel that sees the error word set does not wait any more: the job is lost, it must end.
__device__ __forceinline__ bool wait_flag(const uint32_t* p, uint32_t e, uint32_t* err,
                                          uint64_t timeout_ns) {
    if (static_cast<int32_t>(ld_acquire_sys(p) - e) >= 0) return true;
    if (*reinterpret_cast<volatile uint32_t*>(err) != 0u) return false;
    const uint64_t t0 = global_ns();
    unsigned nap = 200;  // thousands of CTAs may wait for the same word: back off to one poll per ~4 us
    while (static_cast<int32_t>(ld_acquire_sys(p) - e) < 0) {
        __nanosleep(nap);
        if (nap < 4000) nap <<= 1;
        if (global_ns() - t0 > timeout_ns) {
            *reinterpret_cast<volatile uint32_t*>(err) = e ? e : 1u;
            __threadfence_system();
            return false;
        }
    }
    return true;
}

// Step 1.  Call from every thread of every block; returns the collective's number e.
__device__ __forceinline__ uint32_t peer_begin(const PeerSet& ps, int world, const PeerCtx& c) {
    __shared__ uint32_t s_epoch;
    uint32_t* mine = peer_hdr(ps, c.rank);
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(mine + kHdrEpoch) + 1u;
    __syncthreads();
    const uint32_t e = s_epoch;
    if (c.bid == 0 && threadIdx.x < world && static_cast<int>(threadIdx.x) != c.rank)
        st_release_sys(peer_hdr(ps, threadIdx.x) + kHdrReady + c.rank, e);
    return e;
}

// Step 2 for one peer (block-wide; contains a barrier).
__device__ __forceinline__ void peer_wait_ready(const PeerSet& ps, const PeerCtx& c, int q, uint32_t e,
                                                uint64_t timeout_ns) {
    if (threadIdx.x == 0 && q != c.rank) {
        uint32_t* mine = peer_hdr(ps, c.rank);
        wait_flag(mine + kHdrReady + q, e, mine + kHdrError, timeout_ns);
    }
    __syncthreads();
}

// Step 2 for every peer at once.
__device__ __forceinline__ void peer_wait_all_ready(const PeerSet& ps, int world, const PeerCtx& c,
                                                    uint32_t e, uint64_t timeout_ns) {
    if (threadIdx.x < world && static_cast<int>(threadIdx.x) != c.rank) {
        uint32_t* mine = peer_hdr(ps, c.rank);
        wait_flag(mine + kHdrReady + threadIdx.x, e, mine + kHdrError, timeout_ns);
    }
    __syncthreads();
}

// Step 3, COMPLETE form.  Call from every thread of every block after the block's loads / stores.
__device__ __forceinline__ void peer_end(const PeerSet& ps, int world, const PeerCtx& c, uint32_t e,
                                         uint64_t timeout_ns) {
    __shared__ int s_last;
    uint32_t* mine = peer_hdr(ps, c.rank);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t t = atomicAdd(mine + kHdrTicket, 1u);
        s_last = (t == static_cast<uint32_t>(c.nblk) - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();  // the other blocks' stores are ordered before the flags below
    if (threadIdx.x < world && static_cast<int>(threadIdx.x) != c.rank) {
        st_release_sys(peer_hdr(ps, threadIdx.x) + kHdrDone + c.rank, e);
        wait_flag(mine + kHdrDone + threadIdx.x, e, mine + kHdrError, timeout_ns);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mine[kHdrTicket] = 0u;
        st_release_sys(mine + kHdrDone + c.rank, e);
        mine[kHdrEpoch] = e;
    }
}

// Step 3, PROGRESSIVE form, one step: the block has finished its stores into rank q's copy.  The
// block that is last to say so tells rank q (done[rank] = e in q's header).  Block-wide.
__device__ __forceinline__ void peer_step_done(const PeerSet& ps, const PeerCtx& c, int step, int q,
                                               uint32_t e) {
    uint32_t* mine = peer_hdr(ps, c.rank);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t t = atomicAdd(mine + kHdrStep + step, 1u);
        if (t == static_cast<uint32_t>(c.nblk) - 1u) {
            mine[kHdrStep + step] = 0u;
            __threadfence_system();  // the other blocks' stores (fenced before their tickets) first
            st_release_sys(peer_hdr(ps, q) + kHdrDone + c.rank, e);
        }
    }
}

// Step 3, PROGRESSIVE form, end of the kernel: the last block publishes the epoch; no waiting.
__device__ __forceinline__ void peer_end_nowait(const PeerSet& ps, const PeerCtx& c, uint32_t e) {
    uint32_t* mine = peer_hdr(ps, c.rank);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t t = atomicAdd(mine + kHdrTicket, 1u);
        if (t == static_cast<uint32_t>(c.nblk) - 1u) {
            mine[kHdrTicket] = 0u;
            __threadfence();
            mine[kHdrEpoch] = e;
        }
    }
}

// 16-byte loads / stores that may cross NVLink.  Stores bypass L1; loads are cached at L2 only
// (every address is read once per collective, after the owner's flag has been seen).
__device__ __forceinline__ void st_peer_16(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_cg_16(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}

// The rows one rank contributes to a table that lives in every rank's window: rank r's part of
// segment g is the `bytes[g]` bytes (a multiple of 16) at off[g] + r*bytes[g] of every window.
struct PushDesc {
    PeerSet ps;
    int world, rank;
    int64_t off[3], bytes[3];
    int nseg, pushers;
};

// The all-gather by NVLink stores.  `pushers` CTAs of 32 threads (id in [0, pushers)) copy this rank's
// rows to the peers rank-1, rank-2, ... (every rank then receives from one sender at a time, and
// from its successor first); the pusher that finishes a peer last raises done[rank] = epoch there.
// A pusher CTA (32 threads).  `id` in [0, pushers).
__device__ __forceinline__ void push_rows(const PushDesc& fw, int id, uint32_t epoch) {
    const int lane = lane_id();
    uint32_t* hdr = peer_hdr(fw.ps, fw.rank);
    const int64_t stride = static_cast<int64_t>(fw.pushers) * 32;
    for (int s = 1; s < fw.world; ++s) {
        const int q = (fw.rank - s + fw.world) % fw.world;
        for (int g = 0; g < fw.nseg; ++g) {
            const int64_t base = fw.off[g] + fw.rank * fw.bytes[g];
            const int64_t n16 = fw.bytes[g] >> 4;
            const uint4* __restrict__ src = reinterpret_cast<const uint4*>(fw.ps.win[fw.rank] + base);
            uint4* __restrict__ dst = reinterpret_cast<uint4*>(fw.ps.win[q] + base);
            int64_t i = static_cast<int64_t>(id) * 32 + lane;
            for (; i + 3 * stride < n16; i += 4 * stride) {  // four loads in flight per thread
                const uint4 a = ld_cg_16(src + i), b = ld_cg_16(src + i + stride);
                const uint4 c = ld_cg_16(src + i + 2 * stride), d = ld_cg_16(src + i + 3 * stride);
                st_peer_16(dst + i, a);
                st_peer_16(dst + i + stride, b);
                st_peer_16(dst + i + 2 * stride, c);
                st_peer_16(dst + i + 3 * stride, d);
            }
            for (; i < n16; i += stride) st_peer_16(dst + i, ld_cg_16(src + i));
        }
        __threadfence_system();
        __syncwarp();
        if (lane == 0) {
            const uint32_t t = atomicAdd(hdr + kHdrStep + s, 1u);
            if (t == static_cast<uint32_t>(fw.pushers) - 1u) {
                hdr[kHdrStep + s] = 0u;
                __threadfence_system();  // the other pushers' stores (fenced before their tickets) first
                st_release_sys(peer_hdr(fw.ps, q) + kHdrDone + fw.rank, epoch);
            }
        }
    }
}

// Largest grid of `threads`-wide blocks of kernel `fn` that is resident at once on the device.
template <typename F>
static int coresident_blocks(F fn, int threads, size_t smem) {
    int dev = 0, sms = 148, per = 1;
    if (cudaGetDevice(&dev) == cudaSuccess)
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, fn, threads, smem) != cudaSuccess || per < 1)
        per = 1;
    return sms * per;
}

}  // namespace mk
