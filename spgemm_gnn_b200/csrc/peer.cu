// e  Peer-memory exchange of the row-partitioned hot path.
//
// The reference is single-GPU (README_INTEGRATED.md:382 lists multi-GPU as future work).  With
// the adjacency split by rows over the GPUs of one NVSwitch box, a layer has two exchanges:
//
//   forward   every rank needs the whole CBSR table:  all-gather of N*k*(4+w) bytes;
//   backward  every rank holds a full-height partial CBSR gradient:  reduce-scatter of N*k*4.
//
// torch.distributed / NCCL do both (dist.py) and stay the default.  This file is the same pair as
// this library's own kernels over peer-mapped memory (protocol: peer.cuh):
//
//   mk_peer_allgather       each rank STORES its rows into every rank's table (16-byte vector
//                           stores over NVLink, destinations visited in rotated order so that at
//                           any moment every rank receives from one sender);
//   mk_peer_bank_push       (bank.cu) the banking kernel writes its output rows straight into
//                           every rank's table -- compute and all-gather in one kernel;
//   mk_peer_reduce_scatter  each rank LOADS its block of rows from every rank's partial buffer
//                           and folds them in rank order 0..P-1 (fixed order: bit-reproducible,
//                           which NCCL's ring/tree order is not obliged to be).
//
// Why the SpGEMM / SSpMM themselves do not reach into peer memory: a CBSR row is re-read ~deg/P
// times by a rank, so gathering rows over NVLink inside the kernel would move E/P*k*5 bytes per
// rank against N*k*5 once for the all-gather (6-60x more on the BASELINE shapes); the same holds
// for pushing reductions to the owner.  The exchange sits in front of / behind the kernels.
#include <string.h>

#include "peer.cuh"

namespace mk {

constexpr int kMaxSegs = 4;

struct GatherSegs {
    const uint4* src[kMaxSegs];  // this rank's rows
    int64_t n16[kMaxSegs];       // 16-byte units per rank
    int64_t off[kMaxSegs];       // byte offset of rank 0's block inside a window
    int nseg;
};

__global__ void __launch_bounds__(256)
peer_allgather_kernel(const PeerSet ps, const int world, const int rank, const GatherSegs segs,
                      const uint64_t timeout_ns) {
    const uint32_t e = peer_begin(ps, world, rank);
    const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int s = 0; s < world; ++s) {
        const int q = (rank + s) % world;
        peer_wait_ready(ps, rank, q, e, timeout_ns);
        for (int g = 0; g < segs.nseg; ++g) {
            const uint4* __restrict__ src = segs.src[g];
            const int64_t n16 = segs.n16[g];
            uint4* dst = reinterpret_cast<uint4*>(ps.win[q] + segs.off[g]) + rank * n16;
            if (q == rank && dst == src) continue;  // produced in place
            int64_t i = tid;
            for (; i + 3 * nthr < n16; i += 4 * nthr) {  // four loads in flight per thread
                const uint4 a = src[i], b = src[i + nthr], c = src[i + 2 * nthr], d = src[i + 3 * nthr];
                st_peer_16(dst + i, a);
                st_peer_16(dst + i + nthr, b);
                st_peer_16(dst + i + 2 * nthr, c);
                st_peer_16(dst + i + 3 * nthr, d);
            }
            for (; i < n16; i += nthr) st_peer_16(dst + i, src[i]);
        }
    }
    peer_end(ps, world, rank, e, timeout_ns);
}

// out[i] = sum over q = 0..world-1 of window_q[off + rank*n4 + i]   (float4 units)
template <int WORLD>
__global__ void __launch_bounds__(256)
peer_reduce_scatter_kernel(const PeerSet ps, const int world_rt, const int rank, const int64_t off,
                           const int64_t n4, float4* __restrict__ out, const uint64_t timeout_ns) {
    const int world = WORLD > 0 ? WORLD : world_rt;
    const uint32_t e = peer_begin(ps, world, rank);
    peer_wait_all_ready(ps, world, rank, e, timeout_ns);
    const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
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
    peer_end(ps, world, rank, e, timeout_ns);
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

static unsigned pick_grid(int grid, int64_t units, int per_sm) {
    if (grid > 0) return static_cast<unsigned>(grid);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want = (units + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sms) * per_sm;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return static_cast<unsigned>(want);
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
extern "C" int mk_peer_allgather(void* const* h_windows, int world, int rank, int n_seg,
                                 const void* const* h_src, const int64_t* h_bytes,
                                 const int64_t* h_offsets, int grid, int timeout_ms, void* stream) {
    mk::PeerSet ps;
    const int rc = mk::fill_peers(ps, h_windows, world, rank);
    if (rc != MK_OK) return rc;
    if (n_seg < 1 || n_seg > mk::kMaxSegs || !h_src || !h_bytes || !h_offsets) return MK_EINVAL;
    mk::GatherSegs segs;
    memset(&segs, 0, sizeof(segs));
    segs.nseg = n_seg;
    int64_t most = 0;
    for (int g = 0; g < n_seg; ++g) {
        if (h_bytes[g] < 0 || (h_bytes[g] & 15) || h_offsets[g] < MK_PEER_HEADER_BYTES || (h_offsets[g] & 15))
            return MK_EINVAL;
        if (h_bytes[g] > 0 && (!h_src[g] || (reinterpret_cast<uintptr_t>(h_src[g]) & 15))) return MK_EINVAL;
        segs.src[g] = static_cast<const uint4*>(h_src[g]);
        segs.n16[g] = h_bytes[g] / 16;
        segs.off[g] = h_offsets[g];
        if (segs.n16[g] > most) most = segs.n16[g];
    }
    const uint64_t tmo = static_cast<uint64_t>(timeout_ms > 0 ? timeout_ms : 30000) * 1000000ull;
    const unsigned nb = mk::pick_grid(grid, (most + 3) / 4, 4);
    mk::peer_allgather_kernel<<<nb, 256, 0, mk::as_stream(stream)>>>(ps, world, rank, segs, tmo);
    MK_LAUNCH_CHECK("peer_allgather_kernel");
    return MK_OK;
}

extern "C" int mk_peer_reduce_scatter(void* const* h_windows, int world, int rank, int64_t offset,
                                      int64_t block_bytes, float* out, int grid, int timeout_ms,
                                      void* stream) {
    mk::PeerSet ps;
    const int rc = mk::fill_peers(ps, h_windows, world, rank);
    if (rc != MK_OK) return rc;
    if (offset < MK_PEER_HEADER_BYTES || (offset & 15) || block_bytes < 0 || (block_bytes & 15)) return MK_EINVAL;
    if (block_bytes > 0 && (!out || (reinterpret_cast<uintptr_t>(out) & 15))) return MK_EINVAL;
    const int64_t n4 = block_bytes / 16;
    const uint64_t tmo = static_cast<uint64_t>(timeout_ms > 0 ? timeout_ms : 30000) * 1000000ull;
    const unsigned nb = mk::pick_grid(grid, n4, 8);
    cudaStream_t st = mk::as_stream(stream);
    float4* o = reinterpret_cast<float4*>(out);
    switch (world) {
        case 2: mk::peer_reduce_scatter_kernel<2><<<nb, 256, 0, st>>>(ps, world, rank, offset, n4, o, tmo); break;
        case 4: mk::peer_reduce_scatter_kernel<4><<<nb, 256, 0, st>>>(ps, world, rank, offset, n4, o, tmo); break;
        case 8: mk::peer_reduce_scatter_kernel<8><<<nb, 256, 0, st>>>(ps, world, rank, offset, n4, o, tmo); break;
        default: mk::peer_reduce_scatter_kernel<0><<<nb, 256, 0, st>>>(ps, world, rank, offset, n4, o, tmo); break;
    }
    MK_LAUNCH_CHECK("peer_reduce_scatter_kernel");
    return MK_OK;
}
