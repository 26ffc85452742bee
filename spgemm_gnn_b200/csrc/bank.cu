// Banked CBSR: a per-row re-ordering of a CBSR table that removes the shared-memory bank
// conflicts which bound the forward SpGEMM and the backward SSpMM (profiles/r1_v2_*: L1TEX data
// pipe 98 % busy, 3.6 wavefronts per LDS/STS against an ideal 1).
//
// The aggregation kernels give each of the 4 lane groups of a warp (8 lanes, one neighbour each)
// a private set of 8 shared-memory banks, and every column c TWO possible cells there:
//     copy A: row  c/8,                      bank  c%8
//     copy B: row  RA + 8*(c/64) + c%8,      bank  (c/8)%8            (RA = ceil(D/8))
// A warp step `q` touches, for each neighbour, the entries at positions {CAP*t + q : t < 8} of its
// row (lane t, CAP = k/8).  This kernel chooses, per row and once per MaxK output, (1) for every
// entry which copy it uses, balancing the 8 banks (sequential two-choice + two improvement
// passes), and (2) the position of every entry, so that each step takes one entry from every
// non-empty bank and the unavoidable duplicates collect in the last steps.  Measured effect on
// random 32-of-256 rows: 1.4 wavefronts per shared-memory access instead of 3.6.
//
// Output ("banked" CBSR, same N x k shapes): values (and, optionally, column ids) in the new order
// -- still a valid CBSR row: distinct columns, just not ascending -- plus the uint16 cell offset
// 32*row + bank of every entry.  The assignment is sequential in nature, so it runs one thread
// per row (32 rows per warp side by side); the data movement runs one warp per row.
#include "bank.cuh"
#include "common.cuh"

namespace mk {

// Phase 1 (one thread per row): choose copy and position of every entry, leave one descriptor
// byte per entry (position | copy << 7) in shared memory.  Phase 2 (one warp per row): move the
// row with coalesced loads and stores.  bk_index may be null (the forward kernel does not need it).
// In the row-partitioned forward the outputs point into the rank's peer window (peer.py), from
// where the copy engines push the finished rows to the other ranks (mk_peer_push).
template <int K, typename IdxT>
__global__ void __launch_bounds__(128)
cbsr_bank_kernel(const float* __restrict__ sp_data, const IdxT* __restrict__ sp_index,
                 float* __restrict__ bk_data, IdxT* __restrict__ bk_index,
                 uint16_t* __restrict__ bk_slot, uint2* __restrict__ bk_pack, int64_t n, int d) {
    constexpr int DSTRIDE = K + 4;
    __shared__ uint8_t desc[128 * DSTRIDE];
    const int64_t row0 = static_cast<int64_t>(blockIdx.x) * 128;
    const int64_t row = row0 + threadIdx.x;
    const int ra = (d + 7) >> 3;

    if (row < n) {
        const IdxT* __restrict__ ir = sp_index + row * K;
        uint8_t* __restrict__ my_desc = desc + threadIdx.x * DSTRIDE;
        int col[K];
#pragma unroll
        for (int e = 0; e < K; ++e) col[e] = static_cast<int>(ir[e]);
        bank_assign<K>([&](int e) { return col[e]; }, my_desc);
    }
    __syncthreads();

    // ---- phase 2: warp w moves rows [32w, 32w+32) of the block, lane = entry; four rows are
    //      loaded before any is stored so that the row-to-row latency chain is a quarter as long
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int EPT = (K + 31) / 32;  // entries per lane
    constexpr int RB = 4;
    for (int r0 = 0; r0 < 32; r0 += RB) {
        int cc[RB][EPT], dd[RB][EPT];
        float vv[RB][EPT];
#pragma unroll
        for (int q = 0; q < RB; ++q) {
            const int lr = w * 32 + r0 + q;
            const int64_t grow = row0 + lr;
#pragma unroll
            for (int j = 0; j < EPT; ++j) {
                const int e = lane + 32 * j;
                if (grow < n && e < K) {
                    cc[q][j] = static_cast<int>(sp_index[grow * K + e]);
                    vv[q][j] = sp_data[grow * K + e];
                    dd[q][j] = desc[lr * DSTRIDE + e];
                }
            }
        }
#pragma unroll
        for (int q = 0; q < RB; ++q) {
            const int64_t grow = row0 + w * 32 + r0 + q;
#pragma unroll
            for (int j = 0; j < EPT; ++j) {
                const int e = lane + 32 * j;
                if (grow < n && e < K) {
                    const int c = cc[q][j], dsc = dd[q][j];
                    const int p = dsc & 0x7f;
                    const uint16_t cell =
                        static_cast<uint16_t>((dsc & 0x80) ? bank_slot_b(c, ra) : bank_slot_a(c));
                    if (bk_pack != nullptr) {  // packed form: value, cell and column of an entry in 8 bytes
                        bk_pack[grow * K + p] = make_uint2(__float_as_uint(vv[q][j]),
                                                           static_cast<uint32_t>(cell) | (static_cast<uint32_t>(c) << 16));
                    } else {
                        bk_data[grow * K + p] = vv[q][j];
                        bk_slot[grow * K + p] = cell;
                        if (bk_index) bk_index[grow * K + p] = static_cast<IdxT>(c);
                    }
                }
            }
        }
    }
}

template <typename IdxT>
static int launch_bank(const float* sp_data, const void* sp_index, float* bk_data, void* bk_index,
                       uint16_t* bk_slot, uint2* bk_pack, int64_t n, int k, int d, cudaStream_t st) {
    const int64_t blocks = (n + 127) / 128;
    if (blocks > 0x7fffffffLL) return MK_EUNSUPPORTED;
    const IdxT* si = static_cast<const IdxT*>(sp_index);
    IdxT* bi = static_cast<IdxT*>(bk_index);
    const unsigned nb = static_cast<unsigned>(blocks);
    switch (k) {
        case 8: cbsr_bank_kernel<8, IdxT><<<nb, 128, 0, st>>>(sp_data, si, bk_data, bi, bk_slot, bk_pack, n, d); break;
        case 16: cbsr_bank_kernel<16, IdxT><<<nb, 128, 0, st>>>(sp_data, si, bk_data, bi, bk_slot, bk_pack, n, d); break;
        case 32: cbsr_bank_kernel<32, IdxT><<<nb, 128, 0, st>>>(sp_data, si, bk_data, bi, bk_slot, bk_pack, n, d); break;
        case 64: cbsr_bank_kernel<64, IdxT><<<nb, 128, 0, st>>>(sp_data, si, bk_data, bi, bk_slot, bk_pack, n, d); break;
        default: return MK_EUNSUPPORTED;
    }
    MK_LAUNCH_CHECK("cbsr_bank_kernel");
    return MK_OK;
}

}  // namespace mk

extern "C" int mk_banked_supported(int k, int d) {
    return (k == 8 || k == 16 || k == 32 || k == 64) && d >= k && d % 8 == 0 && d <= 512 ? 1 : 0;
}

extern "C" int mk_banked_rows(int d) { return ((d + 7) >> 3) + 8 * ((d + 63) >> 6); }

extern "C" int mk_cbsr_bank(const float* sp_data, const void* sp_index, int index_bytes,
                            float* bk_data, void* bk_index, uint16_t* bk_slot, int64_t n, int k,
                            int d, void* stream) {
    if (n < 0 || d < 1 || k < 1 || k > d) return MK_EINVAL;
    if (index_bytes != 1 && index_bytes != 2) return MK_EINVAL;
    if ((index_bytes == 1 && d > 256)) return MK_EINVAL;
    if (!mk_banked_supported(k, d)) return MK_EUNSUPPORTED;
    if (n == 0) return MK_OK;
    if (!sp_data || !sp_index || !bk_data || !bk_slot) return MK_EINVAL;
    cudaStream_t st = mk::as_stream(stream);
    return index_bytes == 1
               ? mk::launch_bank<uint8_t>(sp_data, sp_index, bk_data, bk_index, bk_slot, nullptr, n, k, d, st)
               : mk::launch_bank<uint16_t>(sp_data, sp_index, bk_data, bk_index, bk_slot, nullptr, n, k, d, st);
}

extern "C" int mk_packed_supported(int k, int d) { return (k == 8 || k == 16) && mk_banked_supported(k, d) ? 1 : 0; }

extern "C" int mk_cbsr_bank_packed(const float* sp_data, const void* sp_index, int index_bytes,
                                   void* bk_pack, int64_t n, int k, int d, void* stream) {
    if (n < 0 || d < 1 || k < 1 || k > d) return MK_EINVAL;
    if (index_bytes != 1 && index_bytes != 2) return MK_EINVAL;
    if ((index_bytes == 1 && d > 256)) return MK_EINVAL;
    if (!mk_packed_supported(k, d)) return MK_EUNSUPPORTED;
    if (n == 0) return MK_OK;
    if (!sp_data || !sp_index || !bk_pack || (reinterpret_cast<uintptr_t>(bk_pack) & 15)) return MK_EINVAL;
    cudaStream_t st = mk::as_stream(stream);
    uint2* bp = static_cast<uint2*>(bk_pack);
    return index_bytes == 1
               ? mk::launch_bank<uint8_t>(sp_data, sp_index, nullptr, nullptr, nullptr, bp, n, k, d, st)
               : mk::launch_bank<uint16_t>(sp_data, sp_index, nullptr, nullptr, nullptr, bp, n, k, d, st);
}
