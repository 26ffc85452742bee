// e  Peer-memory exchange of the row-partitioned hot path.
//
// The reference is single-GPU (README_INTEGRATED.md:382 lists multi-GPU as future work).  With
// the adjacency split by rows over the GPUs of one NVSwitch box, a layer has two exchanges:
//
//   forward   every rank needs the whole CBSR table:  all-gather of N*k*(4+w) bytes;
//   backward  every rank holds a full-height partial CBSR gradient:  reduce-scatter of N*k*4.
//
// torch.distributed / NCCL do both (dist.py, the fallback).  This file is the same pair over
// peer-mapped memory (protocol: peer.cuh):
//
//   mk_peer_publish + mk_peer_push   the all-gather, OVERLAPPED with the kernel that consumes it:
//       the rank's rows are produced straight into its own window (mk_cbsr_bank / mk_topk_cbsr write
//       there), mk_peer_publish opens collective e, and mk_peer_push enqueues -- on a side stream --
//       one copy-engine transfer per peer and table (rank-1, rank-2, ... so that every rank receives
//       from one sender at a time), each peer's transfers followed by a 4-byte copy of e into
//       done[rank] of that peer's header.  No SM is involved and nothing waits: the forward SpGEMM
//       (banked.cu, WAIT form) starts at once on the local rows and checks done[q] before it touches
//       rank q's rows, walking every CSR row in arrival order.  Overwriting a table buffer needs the
//       peers to be done with its previous contents: mk_peer_release (after the consumer) and
//       mk_peer_begin_push (before the producer) are that handshake, two one-block kernels on the
//       main stream; with the two table buffers peer.py alternates, the wait is for a collective two
//       steps back and never blocks in practice.
//   mk_peer_wait_all                 for consumers that cannot wait per block (un-banked kernels).
//   mk_peer_reduce_scatter           each rank LOADS its block of rows from every rank's partial
//       buffer and folds them in rank order 0..P-1 (fixed order: bit-reproducible, which NCCL's
//       ring/tree order is not obliged to be).  Grid sized to be co-resident.
//
// Why the SpGEMM / SSpMM themselves do not reach into peer memory: a CBSR row is re-read ~deg/P
// times by a rank, so gathering rows over NVLink inside the kernel would move E/P*k*5 bytes per
// rank against N*k*5 once for the all-gather (6-60x more on the BASELINE shapes); the same holds
// for pushing reductions to the owner.  The exchange sits next to the kernels, not inside them.
#include <string.h>

#include "peer.cuh"

namespace mk {

constexpr int kHdrSig = 3;      // the value mk_peer_push copies into the peers' done[] words
constexpr int kHdrLastUse = 4;  // [2]: collective that last filled table buffer 0 / 1 of this window

// Before a rank overwrites its rows of table buffer `buf` (and lets the copy engines overwrite the
// peers'): every peer must be done READING the collective that filled that buffer last.  Peers say
// so with mk_peer_release (ready[q] = number of the last collective whose table q has finished
// with).  One block, on the rank's main stream, where no consumer of this window is running any
// more -- it never competes with spinning CTAs for an SM.  With two buffers the wait is for a
// collective two steps back and is over before it starts.
__global__ void peer_wait_ready_kernel(uint32_t* hdr, int world, int rank, int buf, uint64_t timeout_ns) {
    const uint32_t need = hdr[kHdrLastUse + buf];
    if (need != 0u && threadIdx.x < world && static_cast<int>(threadIdx.x) != rank)
        wait_flag(hdr + kHdrReady + threadIdx.x, need, hdr + kHdrError, timeout_ns);
}

// Opens collective e = epoch + 1 on the calling rank's window: the rank's own rows are complete
// (stream order), so done[rank] = e; `sig` = e is what the copy engines hand to the peers.
__global__ void peer_publish_kernel(uint32_t* hdr, int rank, int buf) {
    const uint32_t e = hdr[kHdrEpoch] + 1u;
    hdr[kHdrSig] = e;
    hdr[kHdrLastUse + buf] = e;
    hdr[kHdrTicket] = 0u;  // role tickets of the forward kernel that pushes and consumes this collective
    st_release_sys(hdr + kHdrDone + rank, e);
    hdr[kHdrEpoch] = e;
}

// "I have finished with the table of my current collective": ready[rank] = epoch in every peer's header.
__global__ void peer_release_kernel(const PeerSet ps, int world, int rank) {
    const uint32_t e = peer_hdr(ps, rank)[kHdrEpoch];
    if (threadIdx.x < world && static_cast<int>(threadIdx.x) != rank)
        st_release_sys(peer_hdr(ps, threadIdx.x) + kHdrReady + rank, e);
}

// One block: returns when every sender's rows of the current collective have arrived.
__global__ void peer_wait_all_kernel(uint32_t* hdr, int world, uint64_t timeout_ns) {
    const uint32_t e = hdr[kHdrEpoch];
    if (threadIdx.x < world) wait_flag(hdr + kHdrDone + threadIdx.x, e, hdr + kHdrError, timeout_ns);
}

// The all-gather as a kernel of its own (consumers that cannot wait per block; tests).
__global__ void __launch_bounds__(32) peer_push_sm_kernel(const PushDesc pd) {
    const uint32_t e = peer_hdr(pd.ps, pd.rank)[kHdrEpoch];
    push_rows(pd, static_cast<int>(blockIdx.x), e);
}

struct OutSet {
    float4* out[kMaxPeers];
};

// out[i] = sum over q = 0..world-1 of window_q[off + rank*n4 + i]   (float4 units)
template <int WORLD>
__global__ void __launch_bounds__(256)
peer_reduce_scatter_kernel(const PeerSet ps, const int world_rt, const int rank_arg, const int virt,
                           const int64_t off, const int64_t n4, const OutSet outs,
                           const uint64_t timeout_ns) {
    const int world = WORLD > 0 ? WORLD : world_rt;
    const PeerCtx c = peer_ctx(rank_arg, virt != 0);
    const int rank = c.rank;
    float4* __restrict__ out = outs.out[virt ? rank : 0];
    const uint32_t e = peer_begin(ps, world, c);
    peer_wait_all_ready(ps, world, c, e, timeout_ns);
    const int64_t tid = static_cast<int64_t>(c.bid) * blockDim.x + threadIdx.x;
    const int64_t nthr = static_cast<int64_t>(c.nblk) * blockDim.x;
    for (int64_t i = tid; i < n4; i += nthr) {
        float4 v[WORLD > 0 ? WORLD : 1];
        if (WORLD > 0) {
#pragma unroll
            for (int q = 0; q < WORLD; ++q)  // all loads in flight, folded in rank order
                v[q] = ld_peer_f4(reinterpret_cast<const float*>(ps.win[q] + off) + 4 * (rank * n4 + i));
            float4 acc = v[0];
#pragma unroll
            for (int q = 1; q < WORLD; ++q) {
                acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w;
            }
            out[i] = acc;
        } else {
            float4 acc = ld_peer_f4(reinterpret_cast<const float*>(ps.win[0] + off) + 4 * (rank * n4 + i));
            for (int q = 1; q < world; ++q) {
                const float4 t = ld_peer_f4(reinterpret_cast<const float*>(ps.win[q] + off) + 4 * (rank * n4 + i));
                acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
            }
            out[i] = acc;
        }
    }
    peer_end(ps, world, c, e, timeout_ns);
}

// ---- NVLink multicast (NVLS) forms ----------------------------------------------------------------
// With the windows allocated as symmetric memory bound to a multicast object (peer.py: torch's
// symmetric-memory allocator does the driver plumbing), one address reaches the same offset of EVERY
// rank's window: a store to it is replicated by the NVSwitch, a `multimem.ld_reduce` is summed by it.
// The all-gather then sends every row once instead of P-1 times, the reduce-scatter receives one
// reduced row instead of P.
__device__ __forceinline__ void mc_store_16(void* p, uint4 v) {  // SASS: STG.E.128.STRONG.SYS on the multicast VA
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}
__device__ __forceinline__ float4 mc_load_reduce_f4(const float* p) {  // SASS: LDGMC.E.ADD.F32x4
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}

// All-gather through the multicast address: the rank's rows of every segment are stored ONCE, the
// switch delivers them to all windows (the own one included: same bytes).  The block that finishes
// last raises done[rank] = epoch in every peer's header with ordinary release stores.
__global__ void __launch_bounds__(256)
peer_push_mc_kernel(const PushDesc pd, unsigned char* __restrict__ mc) {
    uint32_t* hdr = peer_hdr(pd.ps, pd.rank);
    const uint32_t e = *reinterpret_cast<volatile uint32_t*>(hdr + kHdrEpoch);
    const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int g = 0; g < pd.nseg; ++g) {
        const int64_t base = pd.off[g] + pd.rank * pd.bytes[g];
        const int64_t n16 = pd.bytes[g] >> 4;
        const uint4* __restrict__ src = reinterpret_cast<const uint4*>(pd.ps.win[pd.rank] + base);
        uint4* __restrict__ dst = reinterpret_cast<uint4*>(mc + base);
        int64_t i = tid;
        for (; i + 3 * nthr < n16; i += 4 * nthr) {  // four loads in flight per thread
            const uint4 a = ld_cg_16(src + i), b = ld_cg_16(src + i + nthr);
            const uint4 c = ld_cg_16(src + i + 2 * nthr), d = ld_cg_16(src + i + 3 * nthr);
            mc_store_16(dst + i, a);
            mc_store_16(dst + i + nthr, b);
            mc_store_16(dst + i + 2 * nthr, c);
            mc_store_16(dst + i + 3 * nthr, d);
        }
        for (; i < n16; i += nthr) mc_store_16(dst + i, ld_cg_16(src + i));
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        const uint32_t t = atomicAdd(hdr + kHdrStep + 1, 1u);
        s_last = (t == gridDim.x - 1u) ? 1 : 0;
        if (s_last) hdr[kHdrStep + 1] = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();  // the other blocks' stores (fenced before their tickets) first
    if (threadIdx.x < pd.world && static_cast<int>(threadIdx.x) != pd.rank)
        st_release_sys(peer_hdr(pd.ps, threadIdx.x) + kHdrDone + pd.rank, e);
}

// out[i] = sum over all ranks of window_q[off + rank*n4 + i], summed by the switch.
__global__ void __launch_bounds__(256)
peer_reduce_scatter_mc_kernel(const PeerSet ps, const int world, const int rank, const unsigned char* __restrict__ mc,
                              const int64_t off, const int64_t n4, float4* __restrict__ out,
                              const uint64_t timeout_ns) {
    PeerCtx c;
    c.rank = rank;
    c.bid = static_cast<int>(blockIdx.x);
    c.nblk = static_cast<int>(gridDim.x);
    const uint32_t e = peer_begin(ps, world, c);
    peer_wait_all_ready(ps, world, c, e, timeout_ns);
    const float* __restrict__ src = reinterpret_cast<const float*>(mc + off) + 4 * (rank * n4);
    const int64_t tid = static_cast<int64_t>(c.bid) * blockDim.x + threadIdx.x;
    const int64_t nthr = static_cast<int64_t>(c.nblk) * blockDim.x;
    for (int64_t i = tid; i < n4; i += nthr) out[i] = mc_load_reduce_f4(src + 4 * i);
    peer_end(ps, world, c, e, timeout_ns);
}

static int fill_peers(PeerSet& ps, void* const* h_windows, int world, int rank) {
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || !h_windows) return MK_EINVAL;
    memset(&ps, 0, sizeof(ps));
    for (int q = 0; q < world; ++q) {
        if (!h_windows[q] || (reinterpret_cast<uintptr_t>(h_windows[q]) & 15)) return MK_EINVAL;
        ps.win[q] = static_cast<unsigned char*>(h_windows[q]);
    }
    return MK_OK;
}

template <int WORLD>
static int launch_rs(const PeerSet& ps, int world, int rank, int virt, int64_t off, int64_t n4,
                     const OutSet& outs, int grid, uint64_t tmo, cudaStream_t st) {
    auto kern = peer_reduce_scatter_kernel<WORLD>;
    // every block may wait for a flag that block 0 of a peer's grid writes: the grid must be resident
    // at once (all virtual ranks together in the single-launch emulation)
    int cap = coresident_blocks(kern, 256, 0);
    if (virt) cap /= world;
    if (cap < 1) return MK_EUNSUPPORTED;
    int64_t want = grid > 0 ? grid : (n4 + 255) / 256;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    const dim3 g(static_cast<unsigned>(want), virt ? static_cast<unsigned>(world) : 1u);
    kern<<<g, 256, 0, st>>>(ps, world, rank, virt, off, n4, outs, tmo);
    MK_LAUNCH_CHECK("peer_reduce_scatter_kernel");
    return MK_OK;
}

static int reduce_scatter_any(void* const* h_windows, int world, int rank, int virt, int64_t offset,
                              int64_t block_bytes, float* const* h_outs, int grid, int timeout_ms,
                              void* stream) {
    PeerSet ps;
    const int rc = fill_peers(ps, h_windows, world, rank);
    if (rc != MK_OK) return rc;
    if (offset < MK_PEER_HEADER_BYTES || (offset & 15) || block_bytes < 0 || (block_bytes & 15)) return MK_EINVAL;
    OutSet outs;
    memset(&outs, 0, sizeof(outs));
    for (int q = 0; q < (virt ? world : 1); ++q) {
        if (block_bytes > 0 && (!h_outs || !h_outs[q] || (reinterpret_cast<uintptr_t>(h_outs[q]) & 15))) return MK_EINVAL;
        outs.out[q] = reinterpret_cast<float4*>(h_outs ? h_outs[q] : nullptr);
    }
    const int64_t n4 = block_bytes / 16;
    const uint64_t tmo = static_cast<uint64_t>(timeout_ms > 0 ? timeout_ms : 30000) * 1000000ull;
    cudaStream_t st = as_stream(stream);
    switch (world) {
        case 2: return launch_rs<2>(ps, world, rank, virt, offset, n4, outs, grid, tmo, st);
        case 4: return launch_rs<4>(ps, world, rank, virt, offset, n4, outs, grid, tmo, st);
        case 8: return launch_rs<8>(ps, world, rank, virt, offset, n4, outs, grid, tmo, st);
        default: return launch_rs<0>(ps, world, rank, virt, offset, n4, outs, grid, tmo, st);
    }
}

}  // namespace mk

// ---- windows ----------------------------------------------------------------------------------
extern "C" int mk_peer_alloc(int64_t bytes, void** window) {
    if (bytes < MK_PEER_HEADER_BYTES || !window) return MK_EINVAL;
    void* p = nullptr;
    MK_CUDA_TRY(cudaMalloc(&p, static_cast<size_t>(bytes)));
    cudaError_t e = cudaMemset(p, 0, static_cast<size_t>(bytes));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(p);
        mk::set_cuda_error(e, "cudaMemset(window)");
        return MK_ECUDA;
    }
    *window = p;
    return MK_OK;
}

extern "C" int mk_peer_free(void* window) {
    if (!window) return MK_OK;
    MK_CUDA_TRY(cudaFree(window));
    return MK_OK;
}

extern "C" int mk_peer_export(void* window, unsigned char* h_handle) {
    if (!window || !h_handle) return MK_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == MK_PEER_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    MK_CUDA_TRY(cudaIpcGetMemHandle(&h, window));
    memcpy(h_handle, &h, sizeof(h));
    return MK_OK;
}

extern "C" int mk_peer_open(const unsigned char* h_handle, void** window) {
    if (!h_handle || !window) return MK_EINVAL;
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle, sizeof(h));
    void* p = nullptr;
    MK_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *window = p;
    return MK_OK;
}

extern "C" int mk_peer_close(void* window) {
    if (!window) return MK_OK;
    MK_CUDA_TRY(cudaIpcCloseMemHandle(window));
    return MK_OK;
}

extern "C" int mk_peer_epoch(const void* window, uint32_t* h_epoch, uint32_t* h_error, void* stream) {
    if (!window || !h_epoch) return MK_EINVAL;
    uint32_t hdr[4] = {0, 0, 0, 0};
    cudaStream_t st = mk::as_stream(stream);
    MK_CUDA_TRY(cudaMemcpyAsync(hdr, window, sizeof(hdr), cudaMemcpyDeviceToHost, st));
    MK_CUDA_TRY(cudaStreamSynchronize(st));
    *h_epoch = hdr[mk::kHdrEpoch];
    if (h_error) *h_error = hdr[mk::kHdrError];
    return MK_OK;
}

// ---- collectives ------------------------------------------------------------------------------
extern "C" int mk_peer_begin_push(void* window, int world, int rank, int buffer, int timeout_ms,
                                  void* stream) {
    if (!window || world < 1 || world > mk::kMaxPeers || rank < 0 || rank >= world) return MK_EINVAL;
    if (buffer < 0 || buffer > 1) return MK_EINVAL;
    const uint64_t tmo = static_cast<uint64_t>(timeout_ms > 0 ? timeout_ms : 30000) * 1000000ull;
    mk::peer_wait_ready_kernel<<<1, 32, 0, mk::as_stream(stream)>>>(static_cast<uint32_t*>(window), world,
                                                                    rank, buffer, tmo);
    MK_LAUNCH_CHECK("peer_wait_ready_kernel");
    return MK_OK;
}

extern "C" int mk_peer_publish(void* window, int rank, int buffer, void* stream) {
    if (!window || rank < 0 || rank >= mk::kMaxPeers || buffer < 0 || buffer > 1) return MK_EINVAL;
    mk::peer_publish_kernel<<<1, 1, 0, mk::as_stream(stream)>>>(static_cast<uint32_t*>(window), rank, buffer);
    MK_LAUNCH_CHECK("peer_publish_kernel");
    return MK_OK;
}

extern "C" int mk_peer_release(void* const* h_windows, int world, int rank, void* stream) {
    mk::PeerSet ps;
    const int rc = mk::fill_peers(ps, h_windows, world, rank);
    if (rc != MK_OK) return rc;
    mk::peer_release_kernel<<<1, 32, 0, mk::as_stream(stream)>>>(ps, world, rank);
    MK_LAUNCH_CHECK("peer_release_kernel");
    return MK_OK;
}

extern "C" int mk_peer_push_steps(void* const* h_windows, int world, int rank, int n_seg,
                                  const int64_t* h_offsets, const int64_t* h_bytes, int first_step,
                                  int step_stride, void* stream);

extern "C" int mk_peer_push(void* const* h_windows, int world, int rank, int n_seg,
                            const int64_t* h_offsets, const int64_t* h_bytes, void* stream) {
    return mk_peer_push_steps(h_windows, world, rank, n_seg, h_offsets, h_bytes, 1, 1, stream);
}

extern "C" int mk_peer_push_steps(void* const* h_windows, int world, int rank, int n_seg,
                                  const int64_t* h_offsets, const int64_t* h_bytes, int first_step,
                                  int step_stride, void* stream) {
    if (first_step < 1 || step_stride < 1) return MK_EINVAL;
    if (world < 1 || world > mk::kMaxPeers || rank < 0 || rank >= world || !h_windows) return MK_EINVAL;
    if (n_seg < 1 || n_seg > 8 || !h_offsets || !h_bytes) return MK_EINVAL;
    for (int q = 0; q < world; ++q)
        if (!h_windows[q]) return MK_EINVAL;
    for (int g = 0; g < n_seg; ++g)
        if (h_bytes[g] < 0 || h_offsets[g] < MK_PEER_HEADER_BYTES) return MK_EINVAL;
    cudaStream_t st = mk::as_stream(stream);
    unsigned char* mine = static_cast<unsigned char*>(h_windows[rank]);
    for (int s = first_step; s < world; s += step_stride) {
        unsigned char* dst = static_cast<unsigned char*>(h_windows[(rank - s + world) % world]);
        for (int g = 0; g < n_seg; ++g) {
            if (h_bytes[g] == 0) continue;
            const int64_t o = h_offsets[g] + rank * h_bytes[g];
            MK_CUDA_TRY(cudaMemcpyAsync(dst + o, mine + o, static_cast<size_t>(h_bytes[g]),
                                        cudaMemcpyDeviceToDevice, st));
        }
        // same stream, same destination: lands behind the rows it announces
        MK_CUDA_TRY(cudaMemcpyAsync(dst + 4 * (mk::kHdrDone + rank), mine + 4 * mk::kHdrSig, 4,
                                    cudaMemcpyDeviceToDevice, st));
    }
    return MK_OK;
}

extern "C" int mk_peer_push_sm(void* const* h_windows, int world, int rank, int n_seg,
                               const int64_t* h_offsets, const int64_t* h_bytes, int pushers, void* stream) {
    mk::PeerSet ps;
    const int rc = mk::fill_peers(ps, h_windows, world, rank);
    if (rc != MK_OK) return rc;
    if (n_seg < 1 || n_seg > 3 || !h_offsets || !h_bytes || pushers < 1 || pushers > 65535) return MK_EINVAL;
    mk::PushDesc pd{};
    pd.ps = ps;
    pd.world = world;
    pd.rank = rank;
    pd.nseg = n_seg;
    pd.pushers = pushers;
    for (int g = 0; g < n_seg; ++g) {
        if (h_bytes[g] < 0 || (h_bytes[g] & 15) || h_offsets[g] < MK_PEER_HEADER_BYTES || (h_offsets[g] & 15))
            return MK_EINVAL;
        pd.off[g] = h_offsets[g];
        pd.bytes[g] = h_bytes[g];
    }
    if (world == 1) return MK_OK;
    mk::peer_push_sm_kernel<<<static_cast<unsigned>(pushers), 32, 0, mk::as_stream(stream)>>>(pd);
    MK_LAUNCH_CHECK("peer_push_sm_kernel");
    return MK_OK;
}

extern "C" int mk_peer_wait_all(void* window, int world, int timeout_ms, void* stream) {
    if (!window || world < 1 || world > mk::kMaxPeers) return MK_EINVAL;
    const uint64_t tmo = static_cast<uint64_t>(timeout_ms > 0 ? timeout_ms : 30000) * 1000000ull;
    mk::peer_wait_all_kernel<<<1, 32, 0, mk::as_stream(stream)>>>(static_cast<uint32_t*>(window), world, tmo);
    MK_LAUNCH_CHECK("peer_wait_all_kernel");
    return MK_OK;
}

extern "C" int mk_peer_reduce_scatter(void* const* h_windows, int world, int rank, int64_t offset,
                                      int64_t block_bytes, float* out, int grid, int timeout_ms,
                                      void* stream) {
    float* outs[1] = {out};
    return mk::reduce_scatter_any(h_windows, world, rank, 0, offset, block_bytes, outs, grid, timeout_ms, stream);
}

extern "C" int mk_peer_reduce_scatter_virtual(void* const* h_windows, int world, int64_t offset,
                                              int64_t block_bytes, float* const* h_outs, int grid,
                                              int timeout_ms, void* stream) {
    return mk::reduce_scatter_any(h_windows, world, 0, 1, offset, block_bytes, h_outs, grid, timeout_ms, stream);
}

extern "C" int mk_peer_push_mc(void* const* h_windows, void* mc_window, int world, int rank, int n_seg,
                               const int64_t* h_offsets, const int64_t* h_bytes, int grid, void* stream) {
    mk::PeerSet ps;
    const int rc = mk::fill_peers(ps, h_windows, world, rank);
    if (rc != MK_OK) return rc;
    if (!mc_window || (reinterpret_cast<uintptr_t>(mc_window) & 15)) return MK_EINVAL;
    if (n_seg < 1 || n_seg > 3 || !h_offsets || !h_bytes) return MK_EINVAL;
    mk::PushDesc pd{};
    pd.ps = ps;
    pd.world = world;
    pd.rank = rank;
    pd.nseg = n_seg;
    int64_t total16 = 0;
    for (int g = 0; g < n_seg; ++g) {
        if (h_bytes[g] < 0 || (h_bytes[g] & 15) || h_offsets[g] < MK_PEER_HEADER_BYTES || (h_offsets[g] & 15))
            return MK_EINVAL;
        pd.off[g] = h_offsets[g];
        pd.bytes[g] = h_bytes[g];
        total16 += h_bytes[g] >> 4;
    }
    if (world == 1) return MK_OK;
    int want = grid > 0 ? grid : 296;  // 2 CTAs per SM saturate the NVLink stores
    const int64_t need = (total16 + 255) / 256;
    if (need < want) want = static_cast<int>(need > 0 ? need : 1);
    mk::peer_push_mc_kernel<<<static_cast<unsigned>(want), 256, 0, mk::as_stream(stream)>>>(
        pd, static_cast<unsigned char*>(mc_window));
    MK_LAUNCH_CHECK("peer_push_mc_kernel");
    return MK_OK;
}

extern "C" int mk_peer_reduce_scatter_mc(void* const* h_windows, const void* mc_window, int world, int rank,
                                         int64_t offset, int64_t block_bytes, float* out, int grid,
                                         int timeout_ms, void* stream) {
    mk::PeerSet ps;
    const int rc = mk::fill_peers(ps, h_windows, world, rank);
    if (rc != MK_OK) return rc;
    if (!mc_window || (reinterpret_cast<uintptr_t>(mc_window) & 15)) return MK_EINVAL;
    if (offset < MK_PEER_HEADER_BYTES || (offset & 15) || block_bytes < 0 || (block_bytes & 15)) return MK_EINVAL;
    if (block_bytes > 0 && (!out || (reinterpret_cast<uintptr_t>(out) & 15))) return MK_EINVAL;
    const int64_t n4 = block_bytes / 16;
    const uint64_t tmo = static_cast<uint64_t>(timeout_ms > 0 ? timeout_ms : 30000) * 1000000ull;
    // every block may wait for a flag that block 0 of a peer's grid writes: the grid must be resident
    int cap = mk::coresident_blocks(mk::peer_reduce_scatter_mc_kernel, 256, 0);
    if (cap < 1) return MK_EUNSUPPORTED;
    int64_t want = grid > 0 ? grid : (n4 + 255) / 256;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    mk::peer_reduce_scatter_mc_kernel<<<static_cast<unsigned>(want), 256, 0, mk::as_stream(stream)>>>(
        ps, world, rank, static_cast<const unsigned char*>(mc_window), offset, n4, reinterpret_cast<float4*>(out),
        tmo);
    MK_LAUNCH_CHECK("peer_reduce_scatter_mc_kernel");
    return MK_OK;
}
