// a-5  On-GPU work partition: CSR row pointer -> {row, loc, len, slot} records.
//
// Replaces the reference's offline kernels/generate_meta.py and the per-call disk read +
// cudaMallocManaged of `../w12_nz64_warp_4/<graph>.warp4` (SPMM_MAXK::do_test,
// so@0x24c50-0x24d2c).  Three small kernels: per-block sums of the record / slot counts,
// a single-block scan of the block sums, and a fill pass that rescans inside the block and
// writes the records.  Runs once per graph; the host layer caches the result.
#include "common.cuh"

namespace mk {

constexpr int kPThreads = 256;
constexpr int kPItems = 4;  // rows per thread
constexpr int kPTile = kPThreads * kPItems;

__device__ __forceinline__ int chunks_of(int deg, int max_nz, int skip_empty) {
    const int c = (deg + max_nz - 1) / max_nz;
    return c < 1 ? (skip_empty ? 0 : 1) : c;
}

struct Pair {
    long long parts;
    long long slots;
};

__device__ __forceinline__ Pair block_reduce(Pair v, Pair* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.parts += __shfl_down_sync(kFull, v.parts, o);
        v.slots += __shfl_down_sync(kFull, v.slots, o);
    }
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sm[w] = v;
    __syncthreads();
    Pair r{0, 0};
    if (threadIdx.x < kPThreads / 32) r = sm[threadIdx.x];
    if (w == 0) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            r.parts += __shfl_down_sync(kFull, r.parts, o);
            r.slots += __shfl_down_sync(kFull, r.slots, o);
        }
    }
    return r;  // valid in thread 0
}

__global__ void __launch_bounds__(kPThreads)
part_block_sums(const int* __restrict__ rs, const int* __restrict__ re, int64_t n, int max_nz,
                int skip_empty, Pair* __restrict__ bsum) {
    __shared__ Pair sm[kPThreads / 32];
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kPTile + threadIdx.x * kPItems;
    Pair v{0, 0};
#pragma unroll
    for (int i = 0; i < kPItems; ++i) {
        const int64_t r = r0 + i;
        if (r < n) {
            const int c = chunks_of(re[r] - rs[r], max_nz, skip_empty);
            v.parts += c;
            v.slots += c > 1 ? c : 0;
        }
    }
    const Pair t = block_reduce(v, sm);
    if (threadIdx.x == 0) bsum[blockIdx.x] = t;
}

// exclusive scan of the block sums in place; totals -> tot[0..1]
__global__ void __launch_bounds__(1024)
part_scan_blocks(Pair* __restrict__ bsum, int nb, long long* __restrict__ tot) {
    __shared__ Pair wsum[32];
    __shared__ Pair carry;
    if (threadIdx.x == 0) carry = Pair{0, 0};
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        Pair v = i < nb ? bsum[i] : Pair{0, 0};
        Pair inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long a = __shfl_up_sync(kFull, inc.parts, o);
            const long long b = __shfl_up_sync(kFull, inc.slots, o);
            if (lane >= o) { inc.parts += a; inc.slots += b; }
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            Pair s = wsum[lane];
            Pair si = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long a = __shfl_up_sync(kFull, si.parts, o);
                const long long b = __shfl_up_sync(kFull, si.slots, o);
                if (lane >= o) { si.parts += a; si.slots += b; }
            }
            wsum[lane] = Pair{si.parts - s.parts, si.slots - s.slots};  // exclusive over warps
        }
        __syncthreads();
        const Pair c = carry;
        const Pair ex{c.parts + wsum[w].parts + inc.parts - v.parts,
                      c.slots + wsum[w].slots + inc.slots - v.slots};
        if (i < nb) bsum[i] = ex;
        __syncthreads();
        if (threadIdx.x == 1023) carry = Pair{ex.parts + v.parts, ex.slots + v.slots};
        __syncthreads();
    }
    if (threadIdx.x == 0) { tot[0] = carry.parts; tot[1] = carry.slots; }
}

__global__ void __launch_bounds__(kPThreads)
part_fill(const int* __restrict__ rs, const int* __restrict__ re, int64_t n, int max_nz,
          int skip_empty, int64_t row_mod, const Pair* __restrict__ boff, mk_part* __restrict__ parts) {
    __shared__ Pair wsum[kPThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kPTile + threadIdx.x * kPItems;
    int cnt[kPItems];
    Pair v{0, 0};
#pragma unroll
    for (int i = 0; i < kPItems; ++i) {
        const int64_t r = r0 + i;
        cnt[i] = r < n ? chunks_of(re[r] - rs[r], max_nz, skip_empty) : 0;
        v.parts += cnt[i];
        v.slots += cnt[i] > 1 ? cnt[i] : 0;
    }
    Pair inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long a = __shfl_up_sync(kFull, inc.parts, o);
        const long long b = __shfl_up_sync(kFull, inc.slots, o);
        if (lane >= o) { inc.parts += a; inc.slots += b; }
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    Pair wo{0, 0};
    for (int q = 0; q < w; ++q) { wo.parts += wsum[q].parts; wo.slots += wsum[q].slots; }
    const Pair b = boff[blockIdx.x];
    long long p = b.parts + wo.parts + inc.parts - v.parts;
    long long s = b.slots + wo.slots + inc.slots - v.slots;
#pragma unroll
    for (int i = 0; i < kPItems; ++i) {
        const int64_t r = r0 + i;
        if (r >= n) break;
        const int lo = rs[r], hi = re[r];
        for (int c = 0; c < cnt[i]; ++c) {
            const int loc = lo + c * max_nz;
            int len = hi - loc;
            len = len > max_nz ? max_nz : (len < 0 ? 0 : len);
            mk_part rec;
            rec.row = static_cast<int>(row_mod > 0 ? r % row_mod : r);
            rec.loc = loc;
            rec.len = len;
            rec.slot = cnt[i] == 1 ? -1 : static_cast<int>(s++);
            parts[p++] = rec;
        }
    }
}

// blk[b * n + r] = first position in CSR row r whose column id is >= b * width (b = 0 .. nb);
// needs ascending column ids inside a row.
__global__ void __launch_bounds__(256)
block_ptr_kernel(const int* __restrict__ ptr, const int* __restrict__ idx, int64_t n, int nb,
                 int width, int* __restrict__ blk) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n * (nb + 1)) return;
    const int64_t r = t % n;
    const int b = static_cast<int>(t / n);
    int lo = ptr[r], hi = ptr[r + 1];
    if (b == 0) { blk[t] = lo; return; }
    if (b == nb) { blk[t] = hi; return; }
    const long long key = static_cast<long long>(b) * width;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (idx[mid] < key) lo = mid + 1; else hi = mid;
    }
    blk[t] = lo;
}

}  // namespace mk

extern "C" int mk_block_ptr(const int32_t* ptr, const int32_t* idx, int64_t n_rows, int n_blocks,
                            int block_width, int32_t* blk_ptr, void* stream) {
    if (n_rows < 0 || n_blocks < 1 || block_width < 1) return MK_EINVAL;
    if (n_rows == 0) return MK_OK;
    if (!ptr || !blk_ptr) return MK_EINVAL;
    const int64_t total = n_rows * (n_blocks + 1);
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) return MK_EUNSUPPORTED;
    mk::block_ptr_kernel<<<static_cast<unsigned>(blocks), 256, 0, mk::as_stream(stream)>>>(
        ptr, idx, n_rows, n_blocks, block_width, blk_ptr);
    MK_LAUNCH_CHECK("block_ptr_kernel");
    return MK_OK;
}

extern "C" int mk_partition_ranges(const int32_t* row_start, const int32_t* row_end, int64_t n_rows,
                                   int64_t row_mod, int max_nz, int skip_empty, mk_part* parts,
                                   int64_t* h_num_parts, int64_t* h_num_slots, void* stream);

extern "C" int mk_partition(const int32_t* ptr, int64_t n_rows, int max_nz, mk_part* parts,
                            int64_t* h_num_parts, int64_t* h_num_slots, void* stream) {
    if (n_rows > 0 && !ptr) return MK_EINVAL;
    return mk_partition_ranges(ptr, ptr ? ptr + 1 : ptr, n_rows, 0, max_nz, 0, parts, h_num_parts,
                               h_num_slots, stream);
}

extern "C" int mk_partition_ranges(const int32_t* row_start, const int32_t* row_end, int64_t n_rows,
                                   int64_t row_mod, int max_nz, int skip_empty, mk_part* parts,
                                   int64_t* h_num_parts, int64_t* h_num_slots, void* stream) {
    const int32_t* ptr = row_start;
    if (n_rows < 0 || max_nz < 1) return MK_EINVAL;
    if (n_rows == 0) {
        if (h_num_parts) *h_num_parts = 0;
        if (h_num_slots) *h_num_slots = 0;
        return MK_OK;
    }
    if (!ptr || !row_end) return MK_EINVAL;
    if (!parts && !h_num_parts) return MK_EINVAL;
    cudaStream_t st = mk::as_stream(stream);
    const int64_t nb64 = (n_rows + mk::kPTile - 1) / mk::kPTile;
    if (nb64 > 0x7fffffffLL) return MK_EUNSUPPORTED;
    const int nb = static_cast<int>(nb64);
    mk::Pair* bsum = nullptr;
    long long* tot = nullptr;
    MK_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&bsum), sizeof(mk::Pair) * nb + 16, st));
    tot = reinterpret_cast<long long*>(bsum + nb);
    mk::part_block_sums<<<nb, mk::kPThreads, 0, st>>>(row_start, row_end, n_rows, max_nz, skip_empty, bsum);
    mk::part_scan_blocks<<<1, 1024, 0, st>>>(bsum, nb, tot);
    int rc = MK_OK;
    if (parts) {
        mk::part_fill<<<nb, mk::kPThreads, 0, st>>>(row_start, row_end, n_rows, max_nz, skip_empty, row_mod,
                                                    bsum, parts);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { mk::set_cuda_error(e, "mk_partition launch"); rc = MK_ECUDA; }
    if (rc == MK_OK && (h_num_parts || h_num_slots)) {
        long long h[2] = {0, 0};
        e = cudaMemcpyAsync(h, tot, sizeof(h), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { mk::set_cuda_error(e, "mk_partition totals"); rc = MK_ECUDA; }
        if (h_num_parts) *h_num_parts = h[0];
        if (h_num_slots) *h_num_slots = h[1];
        if (rc == MK_OK && h[0] > 0x7fffffffLL) rc = MK_EUNSUPPORTED;
    }
    cudaFreeAsync(bsum, st);
    return rc;
}
