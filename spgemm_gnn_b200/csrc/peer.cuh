// Device side of the peer-memory exchange: windows, flags, epochs.
//
// A WINDOW is one cudaMalloc'ed buffer per rank, same size everywhere, mapped into every peer
// process with CUDA IPC (mk_peer_export / mk_peer_open), so that a kernel on rank p can store to
// and load from rank q's copy over NVLink.  It starts with a 1 KB header of 32-bit words:
//
//     word 0        epoch     collectives completed through this window (written by the owner)
//     word 1        ticket    CTAs of the running collective that have finished their part
//     word 2        error     set before a kernel gives up on a peer (bounded wait)
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
//   3. After its part a block fences (system scope) and takes a ticket; the block that takes the
//      last one stores done = e into every peer's header, waits for done[q] >= e from every
//      peer, and publishes epoch = e.  When the kernel has finished, this rank's copy is
//      complete (all-gather) / no longer read by anybody (reduce-scatter).
//
// The epoch lives in device memory, not in a kernel argument, so a captured CUDA graph replays
// correctly.  Waits are bounded: after `timeout_ns` without progress the kernel sets the error
// word and traps -- a dead peer ends the job with a CUDA error instead of hanging the GPU.
#pragma once

#include "common.cuh"

namespace mk {

constexpr int kMaxPeers = MK_PEER_MAX_RANKS;
constexpr int kHdrEpoch = 0, kHdrTicket = 1, kHdrError = 2, kHdrReady = 32, kHdrDone = 64;
constexpr int kHdrBytes = MK_PEER_HEADER_BYTES;
static_assert(kHdrDone + kMaxPeers <= kHdrBytes / 4, "header too small");

struct PeerSet {
    unsigned char* win[kMaxPeers];  // base of rank q's window in THIS process's address space
};

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

// Spin until *p >= e (wrap-safe); give up loudly after timeout_ns.
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t e, uint32_t* err,
                                          uint64_t timeout_ns) {
    if (static_cast<int32_t>(ld_acquire_sys(p) - e) >= 0) return;
    const uint64_t t0 = global_ns();
    while (static_cast<int32_t>(ld_acquire_sys(p) - e) < 0) {
        __nanosleep(100);
        if (global_ns() - t0 > timeout_ns) {
            *err = e;
            __threadfence_system();
            __trap();
        }
    }
}

// Step 1.  Call from every thread of every block; returns the collective's number e.
__device__ __forceinline__ uint32_t peer_begin(const PeerSet& ps, int world, int rank) {
    __shared__ uint32_t s_epoch;
    uint32_t* mine = peer_hdr(ps, rank);
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(mine + kHdrEpoch) + 1u;
    __syncthreads();
    const uint32_t e = s_epoch;
    if (blockIdx.x == 0 && threadIdx.x < world && static_cast<int>(threadIdx.x) != rank)
        st_release_sys(peer_hdr(ps, threadIdx.x) + kHdrReady + rank, e);
    return e;
}

// Step 2 for one peer (block-wide; contains a barrier).
__device__ __forceinline__ void peer_wait_ready(const PeerSet& ps, int rank, int q, uint32_t e,
                                                uint64_t timeout_ns) {
    if (threadIdx.x == 0 && q != rank) {
        uint32_t* mine = peer_hdr(ps, rank);
        wait_flag(mine + kHdrReady + q, e, mine + kHdrError, timeout_ns);
    }
    __syncthreads();
}

// Step 2 for every peer at once.
__device__ __forceinline__ void peer_wait_all_ready(const PeerSet& ps, int world, int rank,
                                                    uint32_t e, uint64_t timeout_ns) {
    if (threadIdx.x < world && static_cast<int>(threadIdx.x) != rank) {
        uint32_t* mine = peer_hdr(ps, rank);
        wait_flag(mine + kHdrReady + threadIdx.x, e, mine + kHdrError, timeout_ns);
    }
    __syncthreads();
}

// Step 3.  Call from every thread of every block after the block's loads / stores.
__device__ __forceinline__ void peer_end(const PeerSet& ps, int world, int rank, uint32_t e,
                                         uint64_t timeout_ns) {
    __shared__ int s_last;
    uint32_t* mine = peer_hdr(ps, rank);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t t = atomicAdd(mine + kHdrTicket, 1u);
        s_last = (t == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();  // the other blocks' stores are ordered before the flags below
    if (threadIdx.x < world && static_cast<int>(threadIdx.x) != rank) {
        st_release_sys(peer_hdr(ps, threadIdx.x) + kHdrDone + rank, e);
        wait_flag(mine + kHdrDone + threadIdx.x, e, mine + kHdrError, timeout_ns);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mine[kHdrTicket] = 0u;
        mine[kHdrEpoch] = e;
    }
}

// 16-byte loads / stores that may cross NVLink.  Stores bypass L1; loads are cached at L2 only
// (every address is read once per collective, after the owner's flag has been seen).
__device__ __forceinline__ void st_peer_16(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}

}  // namespace mk
