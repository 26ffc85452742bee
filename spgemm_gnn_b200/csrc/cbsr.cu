// a-2  CBSR <-> dense row movement.
//
// mk_cbsr_scatter: dense[i,:] = 0; dense[i, sp_index[i,t]] = g[i,t].  One warp per row; the row
// is assembled in shared memory and leaves with full-line vector stores, so the N*D*4 output
// is written exactly once (the reference's maxk_backward_cuda, maxk_cuda_kernels.o@0x4d0, is
// an N*k host loop of .item() copies onto a torch::zeros tensor).
// mk_cbsr_gather:  out[i,t] = dense[i, sp_index[i,t]].
#include "common.cuh"

namespace mk {

constexpr int kRowsPerBlock = 8;  // warps per block

template <typename IdxT>
__global__ void __launch_bounds__(kRowsPerBlock * 32)
cbsr_scatter_kernel(const float* __restrict__ g, const IdxT* __restrict__ sp_index,
                    float* __restrict__ dense, int64_t n, int k, int d) {
    extern __shared__ float srow[];  // kRowsPerBlock * dpad
    const int dpad = (d + 3) & ~3;
    const int w = threadIdx.x >> 5;
    const int lane = lane_id();
    float* __restrict__ buf = srow + w * dpad;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + w;  // rows per CTA = warps
    if (row >= n) return;
    for (int c = lane * 4; c < dpad; c += 128)
        *reinterpret_cast<float4*>(buf + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    const float* __restrict__ gr = g + row * k;
    const IdxT* __restrict__ ir = sp_index + row * k;
    for (int t = lane; t < k; t += 32) {
        const int c = static_cast<int>(ir[t]);
        if (c < d) buf[c] = gr[t];
    }
    __syncwarp();
    float* __restrict__ o = dense + row * d;
    if ((d & 3) == 0 && (reinterpret_cast<uintptr_t>(dense) & 15) == 0) {
        for (int c = lane * 4; c < d; c += 128)
            st_stream_f4(o + c, *reinterpret_cast<const float4*>(buf + c));
    } else {
        for (int c = lane; c < d; c += 32) o[c] = buf[c];
    }
}

// Rows too wide to stage (D > 51,200): the row is zeroed in place, then the k values land on top --
// `__syncwarp()` orders the two sets of stores of the warp.
template <typename IdxT>
__global__ void __launch_bounds__(256)
cbsr_scatter_direct_kernel(const float* __restrict__ g, const IdxT* __restrict__ sp_index,
                           float* __restrict__ dense, int64_t n, int k, int d) {
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const int lane = lane_id();
    float* __restrict__ o = dense + row * d;
    if ((d & 3) == 0 && (reinterpret_cast<uintptr_t>(dense) & 15) == 0) {
        for (int c = lane * 4; c < d; c += 128)
            *reinterpret_cast<float4*>(o + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (int c = lane; c < d; c += 32) o[c] = 0.f;
    }
    __syncwarp();
    const float* __restrict__ gr = g + row * k;
    const IdxT* __restrict__ ir = sp_index + row * k;
    for (int t = lane; t < k; t += 32) {
        const int c = static_cast<int>(ir[t]);
        if (c < d) o[c] = gr[t];
    }
}

template <typename IdxT>
__global__ void __launch_bounds__(256)
cbsr_gather_kernel(const float* __restrict__ dense, const IdxT* __restrict__ sp_index,
                   float* __restrict__ out, int64_t n, int k, int d) {
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const int lane = lane_id();
    const float* __restrict__ dr = dense + row * d;
    const IdxT* __restrict__ ir = sp_index + row * k;
    float* __restrict__ o = out + row * k;
    for (int t = lane; t < k; t += 32) {
        const int c = static_cast<int>(ir[t]);
        o[t] = c < d ? dr[c] : 0.0f;
    }
}

}  // namespace mk

extern "C" int mk_cbsr_scatter(const float* g, const void* sp_index, int index_bytes, float* dense,
                               int64_t n, int k, int d, void* stream) {
    if (n < 0 || d < 1 || k < 1 || k > d) return MK_EINVAL;
    if (index_bytes != 1 && index_bytes != 2) return MK_EINVAL;
    if (n == 0) return MK_OK;
    if (!g || !sp_index || !dense) return MK_EINVAL;
    const int dpad = (d + 3) & ~3;
    // rows (= warps) per CTA: 8 while their staging rows fit shared memory, fewer for very wide rows,
    // un-staged above that -- every D the top-k accepts (<= 65,536) can be scattered back
    int rpb = mk::kRowsPerBlock;
    while (rpb > 1 && static_cast<size_t>(rpb) * dpad * 4 > 200 * 1024) rpb >>= 1;
    const size_t smem = static_cast<size_t>(rpb) * dpad * 4;
    cudaStream_t st = mk::as_stream(stream);
    if (smem > 200 * 1024) {
        const int64_t wide_blocks = (n * 32 + 255) / 256;
        if (wide_blocks > 0x7fffffffLL) return MK_EUNSUPPORTED;
        if (index_bytes == 1)
            mk::cbsr_scatter_direct_kernel<uint8_t><<<static_cast<unsigned>(wide_blocks), 256, 0, st>>>(
                g, static_cast<const uint8_t*>(sp_index), dense, n, k, d);
        else
            mk::cbsr_scatter_direct_kernel<uint16_t><<<static_cast<unsigned>(wide_blocks), 256, 0, st>>>(
                g, static_cast<const uint16_t*>(sp_index), dense, n, k, d);
        MK_LAUNCH_CHECK("cbsr_scatter_direct_kernel");
        return MK_OK;
    }
    const int64_t blocks = (n + rpb - 1) / rpb;
    if (blocks > 0x7fffffffLL) return MK_EUNSUPPORTED;
    if (index_bytes == 1) {
        if (smem > 48 * 1024)
            MK_CUDA_TRY(cudaFuncSetAttribute(mk::cbsr_scatter_kernel<uint8_t>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem)));
        mk::cbsr_scatter_kernel<uint8_t><<<static_cast<unsigned>(blocks), rpb * 32,
                                           smem, st>>>(
            g, static_cast<const uint8_t*>(sp_index), dense, n, k, d);
    } else {
        if (smem > 48 * 1024)
            MK_CUDA_TRY(cudaFuncSetAttribute(mk::cbsr_scatter_kernel<uint16_t>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem)));
        mk::cbsr_scatter_kernel<uint16_t><<<static_cast<unsigned>(blocks), rpb * 32,
                                            smem, st>>>(
            g, static_cast<const uint16_t*>(sp_index), dense, n, k, d);
    }
    MK_LAUNCH_CHECK("cbsr_scatter_kernel");
    return MK_OK;
}

extern "C" int mk_cbsr_gather(const float* dense, const void* sp_index, int index_bytes,
                              float* out, int64_t n, int k, int d, void* stream) {
    if (n < 0 || d < 1 || k < 1 || k > d) return MK_EINVAL;
    if (index_bytes != 1 && index_bytes != 2) return MK_EINVAL;
    if (n == 0) return MK_OK;
    if (!dense || !sp_index || !out) return MK_EINVAL;
    const int64_t blocks = (n * 32 + 255) / 256;
    if (blocks > 0x7fffffffLL) return MK_EUNSUPPORTED;
    cudaStream_t st = mk::as_stream(stream);
    if (index_bytes == 1)
        mk::cbsr_gather_kernel<uint8_t><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
            dense, static_cast<const uint8_t*>(sp_index), out, n, k, d);
    else
        mk::cbsr_gather_kernel<uint16_t><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
            dense, static_cast<const uint16_t*>(sp_index), out, n, k, d);
    MK_LAUNCH_CHECK("cbsr_gather_kernel");
    return MK_OK;
}
