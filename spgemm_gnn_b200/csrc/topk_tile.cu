// a-1 (+ f-3)  MaxK top-k, second generation: fewer search steps, coalesced output, and -- where the
// forward that follows runs on the banked table -- the banking of bank.cu in the same kernel.
//
// topk.cu's kernel is issue-bound (profiles/r1_topk_reddit_k32.summary.txt: 818 warp instructions per
// 256-wide row, issue slots 92 % busy, DRAM 16 %): 13.7 bitwise search steps of ~28 instructions,
// and an output phase of eight divergent store blocks.  Here:
//   * the k-th largest key is bracketed from the data -- [min over lanes of the lane maximum (at
//     least 32 keys are >= it), global maximum] -- and found by interpolation on the counts (Illinois
//     rule, bisection after 8 steps), with two short cuts for ties: a probe of lo + 1 / hi - 1 after
//     four moves of the same end, and a jump to the next key value when the count stops changing.
//     ~6.5 counting steps per row on Gaussian rows instead of 13.7, <= ~8 on rows full of ties (a
//     bitwise search needs all 32 there); counts use the carry chain of topk.cuh;
//   * kept entries are compacted through shared memory (one packed prefix sum for two column groups,
//     predicated STS) and leave as one coalesced store of values and one of column ids per row;
//   * a warp owns a run of consecutive rows and keeps the next row's loads in flight while it works
//     on the current one.
// Fused banking (BK = k in {8,16,32,64}): the warp keeps the compacted rows of a 32-row tile in
// shared memory, then runs bank.cuh's assignment one THREAD per row (it is sequential in nature) and
// moves the tile to the banked table -- what mk_topk_cbsr followed by mk_cbsr_bank produce, bit for
// bit, without the CBSR round trip through L2/HBM and without the second launch.
//
// Replaces: maxk_kernel (so@0x21110) and torch.topk + zeros_like + scatter_ + mul
// (utils/models.py:14-20); same contract as topk.cu (exact, ties -> lower column, NaN above +inf).
#include "bank.cuh"
#include "common.cuh"
#include "topk.cuh"

namespace mk {

constexpr int kTileWarps = 4;  // warps per CTA; every warp is independent (no __syncthreads)

template <int E>
__device__ __forceinline__ int count_ge_all(const uint32_t (&key)[E], uint32_t cand) {
    const uint32_t ncand = 0u - cand;  // cand != 0
    int c = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) count_ge(c, key[e], ncand);
    return __reduce_add_sync(kFull, c);
}

// thr = the k-th largest key of the row spread over the warp (padding keys are 0, below every real
// key).  exact: #{key >= thr} == k; otherwise #{key >= thr} > k > #{key > thr} (ties on thr).
template <int E>
__device__ __forceinline__ void kth_largest_key(const uint32_t (&key)[E], int k, int d, uint32_t& thr,
                                                bool& exact) {
    uint32_t lm = key[0];
#pragma unroll
    for (int e = 1; e < E; ++e) lm = max(lm, key[e]);
    const uint32_t gmax = __reduce_max_sync(kFull, lm);
    exact = false;
#ifdef MK_TILE_BITWISE
    if (true) {  // measurement variant: topk.cu's bitwise search inside this kernel
#else
    if (gmax == 0xFFFFFFFFu) {  // a NaN in the row: hi = gmax + 1 would wrap; plain bitwise search
#endif
        thr = 0;
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t cand = thr | (1u << bit);
            const int c = count_ge_all(key, cand);
            if (c >= k) {
                thr = cand;
                if (c == k) { exact = true; return; }
            }
        }
        return;
    }
    // invariant: #{key >= lo} = clo >= k (> k inside the loop), #{key >= hi} < k
    uint32_t lo = 1u, hi = gmax + 1u;
    int clo = d;
    if (k <= 32) {  // every lane's maximum is >= the smallest of them: at least 32 keys are
        lo = max(__reduce_min_sync(kFull, lm), 1u);
        clo = count_ge_all(key, lo);
    }
    if (clo == k) { thr = lo; exact = true; return; }
    float flo = static_cast<float>(clo - k) + 0.5f, fhi = 0.5f - static_cast<float>(k);
    int side = 0, run_lo = 0, run_hi = 0, it = 0;
    while (hi - lo > 1u) {
        uint32_t cand;
        if (run_hi >= 4) {         // hi keeps falling: is lo itself the answer (many ties at lo)?
            cand = lo + 1u;
            run_hi = 0;
        } else if (run_lo >= 4) {  // lo keeps rising: ties just below hi?
            cand = hi - 1u;
            run_lo = 0;
        } else if (it < 8) {       // regula falsi on the counts, in key space
            const float f = __fdividef(flo, flo - fhi);
            cand = lo + __float2uint_rz(f * __uint2float_rz(hi - lo));
            cand = min(max(cand, lo + 1u), hi - 1u);
        } else {
            cand = lo + ((hi - lo) >> 1);
        }
        ++it;
        const int c = count_ge_all(key, cand);
        if (c == k) { thr = cand; exact = true; return; }
        if (c > k) {
            if (c == clo && cand > lo + 1u) {
                // no key in [lo, cand): jump to the smallest key >= cand (keys below cand wrap to
                // values above every real difference) and test the key value after it
                uint32_t m = 0xFFFFFFFFu;
#pragma unroll
                for (int e = 0; e < E; ++e) m = min(m, key[e] - cand);
                const uint32_t v = cand + __reduce_min_sync(kFull, m);  // <= gmax < 2^32 - 1
                const int c2 = count_ge_all(key, v + 1u);
                if (c2 < k) { thr = v; return; }  // ties on v
                if (c2 == k) { thr = v + 1u; exact = true; return; }
                lo = v + 1u;
                clo = c2;
                flo = static_cast<float>(c2 - k) + 0.5f;
                side = 1;
                run_lo = run_hi = 0;
                continue;
            }
            lo = cand;
            clo = c;
            flo = static_cast<float>(c - k) + 0.5f;
            if (side == 1) fhi *= 0.5f;
            side = 1;
            ++run_lo;
            run_hi = 0;
        } else {
            hi = cand;
            fhi = static_cast<float>(c - k) + 0.5f;
            if (side == -1) flo *= 0.5f;
            side = -1;
            ++run_hi;
            run_lo = 0;
        }
    }
    thr = lo;  // #{key >= lo} > k > #{key >= lo + 1}: ties on lo
}

template <int NV4, bool VEC>
__device__ __forceinline__ void load_row(const float* __restrict__ xr, int d, int lane, float (&v)[NV4 * 4]) {
#pragma unroll
    for (int j = 0; j < NV4; ++j) {
        const int c0 = j * 128 + lane * 4;
        if (VEC && c0 + 3 < d) {
            const float4 f = ld_stream_f4(xr + c0);
            v[4 * j + 0] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[4 * j + i] = (c0 + i < d) ? ld_stream_f1(xr + c0 + i) : 0.0f;
        }
    }
}

// Column of element (j, i) held by `lane`: j*128 + lane*4 + i.
template <int NV4, typename IdxT, bool VEC, int BK>
__global__ void __launch_bounds__(kTileWarps * 32, NV4 > 2 ? 1 : (BK == 64 ? 4 : 8))
topk_tile_kernel(const float* __restrict__ x, int64_t n, int d, int k, int rows_per_warp, int warp_bytes,
                 float* __restrict__ sp_data, IdxT* __restrict__ sp_index, float* __restrict__ bk_data,
                 uint16_t* __restrict__ bk_slot, uint2* __restrict__ bk_pack) {
    constexpr int E = NV4 * 4;
    constexpr bool BANK = BK > 0;
    constexpr int CS = BANK ? BK + 4 / static_cast<int>(sizeof(IdxT)) : 0;  // stride of a row of column ids
    constexpr int DS = BK + 4;                                            // ... of descriptor bytes
    extern __shared__ __align__(16) unsigned char tile_smem[];
    const int lane = lane_id();
    const int w = threadIdx.x >> 5;
    unsigned char* __restrict__ mine = tile_smem + static_cast<size_t>(w) * warp_bytes;
    // BANK: [32][BK] values, [32][CS] column ids, [32][DS] descriptors; else one row of each
    float* __restrict__ sval = reinterpret_cast<float*>(mine);
    IdxT* __restrict__ scol = reinterpret_cast<IdxT*>(mine + (BANK ? 32 * BK : k) * 4);
    [[maybe_unused]] uint8_t* __restrict__ desc =
        reinterpret_cast<uint8_t*>(mine + 32 * BK * 4 + 32 * CS * static_cast<int>(sizeof(IdxT)));

    const int64_t gw = static_cast<int64_t>(blockIdx.x) * kTileWarps + w;
    const int64_t first = gw * rows_per_warp;
    const int64_t last = min(first + rows_per_warp, n);
    if (first >= last) return;
    [[maybe_unused]] const int ra = (d + 7) >> 3;

    float cur[E];
    load_row<NV4, VEC>(x + first * d, d, lane, cur);
    for (int64_t t0 = first; t0 < last; t0 += 32) {
        const int tr = static_cast<int>(min(static_cast<int64_t>(32), last - t0));
        for (int r = 0; r < tr; ++r) {
            const int64_t row = t0 + r;
            float v[E];
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = cur[e];
            if (row + 1 < last) load_row<NV4, VEC>(x + (row + 1) * d, d, lane, cur);  // in flight during this row

            uint32_t key[E];
#pragma unroll
            for (int j = 0; j < NV4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    key[4 * j + i] = (j * 128 + lane * 4 + i < d) ? order_key(v[4 * j + i]) : 0u;

            uint32_t thr;
            bool exact;
            kth_largest_key<E>(key, k, d, thr, exact);

            uint32_t selmask = 0;  // bit e: element e is kept
            if (exact) {
#pragma unroll
                for (int e = 0; e < E; ++e) selmask |= (key[e] >= thr ? 1u : 0u) << e;
            } else {
                // more than k keys are >= thr: all keys > thr are kept, the lowest-column ties fill up
                int gt = 0;
#pragma unroll
                for (int e = 0; e < E; ++e) gt += (key[e] > thr) ? 1 : 0;
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

            // ---- compaction: ascending-column slot of every kept element, two column groups per
            //      prefix sum (counts packed 16 + 16 bits)
            float* __restrict__ sv = BANK ? sval + r * BK : sval;
            IdxT* __restrict__ sc = BANK ? scol + r * CS : scol;
            int before = 0;
#pragma unroll
            for (int j0 = 0; j0 < NV4; j0 += 2) {
                const uint32_t m0 = (selmask >> (4 * j0)) & 0xFu;
                const uint32_t m1 = (j0 + 1 < NV4) ? (selmask >> (4 * j0 + 4)) & 0xFu : 0u;
                const int c0 = __popc(m0), c1 = __popc(m1);
                const int incl = warp_incl_scan(c0 | (c1 << 16), lane);
                const int tot = __shfl_sync(kFull, incl, 31);
                const int p0 = before + (incl & 0xffff) - c0;
                const int p1 = before + (tot & 0xffff) + (incl >> 16) - c1;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if ((m0 >> i) & 1u) {
                        const int p = p0 + __popc(m0 & ((1u << i) - 1u));
                        sv[p] = v[4 * j0 + i];
                        sc[p] = static_cast<IdxT>(j0 * 128 + lane * 4 + i);
                    }
                }
                if (j0 + 1 < NV4) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if ((m1 >> i) & 1u) {
                            const int p = p1 + __popc(m1 & ((1u << i) - 1u));
                            sv[p] = v[4 * j0 + 4 + i];
                            sc[p] = static_cast<IdxT>((j0 + 1) * 128 + lane * 4 + i);
                        }
                    }
                }
                before += (tot & 0xffff) + (tot >> 16);
            }
            __syncwarp();
            // ---- the sorted CBSR row leaves coalesced
            if (sp_data != nullptr)
                for (int t = lane; t < k; t += 32) sp_data[row * k + t] = sv[t];
            for (int t = lane; t < k; t += 32) sp_index[row * k + t] = sc[t];
            if (!BANK) __syncwarp();  // the row buffer is re-used
        }

        if constexpr (BANK) {
            // ---- bank assignment: thread = row of the tile (sequential two-choice, bank.cuh)
            if (lane < tr) {
                const IdxT* __restrict__ ir = scol + lane * CS;  // padded stride: no bank conflicts
                bank_assign<BK>([&](int e) { return static_cast<int>(ir[e]); }, desc + lane * DS);
            }
            __syncwarp();
            // ---- the banked rows: lane = entry, scattered inside the row's own 128-byte lines
            constexpr int EPT = (BK + 31) / 32;
            for (int r = 0; r < tr; ++r) {
                const int64_t grow = t0 + r;
#pragma unroll
                for (int j = 0; j < EPT; ++j) {
                    const int e = lane + 32 * j;
                    if (e < BK) {
                        const int c = static_cast<int>(scol[r * CS + e]);
                        const float val = sval[r * BK + e];
                        const int dsc = desc[r * DS + e];
                        const int p = dsc & 0x7f;
                        const uint16_t cell =
                            static_cast<uint16_t>((dsc & 0x80) ? bank_slot_b(c, ra) : bank_slot_a(c));
                        if (bk_pack != nullptr) {
                            bk_pack[grow * BK + p] = make_uint2(__float_as_uint(val),
                                                                static_cast<uint32_t>(cell) | (static_cast<uint32_t>(c) << 16));
                        } else {
                            bk_data[grow * BK + p] = val;
                            bk_slot[grow * BK + p] = cell;
                        }
                    }
                }
            }
            __syncwarp();  // the tile buffers are re-used
        }
    }
}

template <int NV4, typename IdxT, bool VEC, int BK>
static int launch_tile_one(const float* x, int64_t n, int d, int k, float* sp_data, IdxT* sp_index,
                           float* bk_data, uint16_t* bk_slot, uint2* bk_pack, cudaStream_t st) {
    auto kern = topk_tile_kernel<NV4, IdxT, VEC, BK>;
    constexpr int CS = BK > 0 ? BK + 4 / static_cast<int>(sizeof(IdxT)) : 0;
    int warp_bytes = BK > 0 ? 32 * BK * 4 + 32 * CS * static_cast<int>(sizeof(IdxT)) + 32 * (BK + 4)
                            : k * (4 + static_cast<int>(sizeof(IdxT)));
    warp_bytes = (warp_bytes + 15) & ~15;
    const size_t smem = static_cast<size_t>(warp_bytes) * kTileWarps;
    if (smem > 200 * 1024) return MK_EUNSUPPORTED;
    if (smem > 48 * 1024)
        MK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    // as many warps as are resident at once, each with an equal run of consecutive rows
    int dev = 0, sms = 148, per = 1;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, kTileWarps * 32, smem) != cudaSuccess || per < 1)
        per = 1;
    const int64_t resident_warps = static_cast<int64_t>(sms) * per * kTileWarps;
    int64_t rpw = (n + resident_warps - 1) / resident_warps;
    if (rpw < 1) rpw = 1;
    if (rpw > (1 << 20)) rpw = 1 << 20;
    const int64_t warps = (n + rpw - 1) / rpw;
    const int64_t blocks = (warps + kTileWarps - 1) / kTileWarps;
    if (blocks > 0x7fffffffLL) return MK_EUNSUPPORTED;
    kern<<<static_cast<unsigned>(blocks), kTileWarps * 32, smem, st>>>(
        x, n, d, k, static_cast<int>(rpw), warp_bytes, sp_data, sp_index, bk_data, bk_slot, bk_pack);
    MK_LAUNCH_CHECK("topk_tile_kernel");
    return MK_OK;
}

template <int NV4, typename IdxT, int BK>
static int launch_tile_vec(const float* x, int64_t n, int d, int k, float* sp_data, void* sp_index,
                           float* bk_data, uint16_t* bk_slot, uint2* bk_pack, cudaStream_t st) {
    const bool vec = (d % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0);
    IdxT* si = static_cast<IdxT*>(sp_index);
    return vec ? launch_tile_one<NV4, IdxT, true, BK>(x, n, d, k, sp_data, si, bk_data, bk_slot, bk_pack, st)
               : launch_tile_one<NV4, IdxT, false, BK>(x, n, d, k, sp_data, si, bk_data, bk_slot, bk_pack, st);
}

template <typename IdxT, int BK>
static int launch_tile_d(const float* x, int64_t n, int d, int k, float* sp_data, void* sp_index,
                         float* bk_data, uint16_t* bk_slot, uint2* bk_pack, cudaStream_t st) {
    if (d <= 128) return launch_tile_vec<1, IdxT, BK>(x, n, d, k, sp_data, sp_index, bk_data, bk_slot, bk_pack, st);
    if (d <= 256) return launch_tile_vec<2, IdxT, BK>(x, n, d, k, sp_data, sp_index, bk_data, bk_slot, bk_pack, st);
    if (d <= 384) return launch_tile_vec<3, IdxT, BK>(x, n, d, k, sp_data, sp_index, bk_data, bk_slot, bk_pack, st);
    if (d <= 512) return launch_tile_vec<4, IdxT, BK>(x, n, d, k, sp_data, sp_index, bk_data, bk_slot, bk_pack, st);
    if (BK > 0) return MK_EUNSUPPORTED;  // banked tables stop at d = 512
    if (d <= 768) return launch_tile_vec<6, IdxT, 0>(x, n, d, k, sp_data, sp_index, nullptr, nullptr, nullptr, st);
    if (d <= 1024) return launch_tile_vec<8, IdxT, 0>(x, n, d, k, sp_data, sp_index, nullptr, nullptr, nullptr, st);
    return MK_EUNSUPPORTED;
}

// Plain top-k -> sorted CBSR with the tiled kernel (d <= 1024); MK_EUNSUPPORTED lets topk.cu fall
// back to its shared-memory kernel for wider rows.
int launch_topk_tile(const float* x, int64_t n, int d, int k, float* sp_data, void* sp_index, int index_bytes,
                     cudaStream_t st) {
    if (d > 1024) return MK_EUNSUPPORTED;
    return index_bytes == 1 ? launch_tile_d<uint8_t, 0>(x, n, d, k, sp_data, sp_index, nullptr, nullptr, nullptr, st)
                            : launch_tile_d<uint16_t, 0>(x, n, d, k, sp_data, sp_index, nullptr, nullptr, nullptr, st);
}

template <typename IdxT>
static int launch_tile_bank(const float* x, int64_t n, int d, int k, float* sp_data, void* sp_index,
                            float* bk_data, uint16_t* bk_slot, uint2* bk_pack, cudaStream_t st) {
    switch (k) {
        case 8: return launch_tile_d<IdxT, 8>(x, n, d, k, sp_data, sp_index, bk_data, bk_slot, bk_pack, st);
        case 16: return launch_tile_d<IdxT, 16>(x, n, d, k, sp_data, sp_index, bk_data, bk_slot, bk_pack, st);
        case 32: return launch_tile_d<IdxT, 32>(x, n, d, k, sp_data, sp_index, bk_data, bk_slot, bk_pack, st);
        case 64: return launch_tile_d<IdxT, 64>(x, n, d, k, sp_data, sp_index, bk_data, bk_slot, bk_pack, st);
        default: return MK_EUNSUPPORTED;
    }
}

}  // namespace mk

extern "C" int mk_banked_supported(int k, int d);
extern "C" int mk_packed_supported(int k, int d);

extern "C" int mk_topk_cbsr_bank(const float* x, int64_t n, int d, int k, float* sp_data, void* sp_index,
                                 int index_bytes, float* bk_data, uint16_t* bk_slot, void* bk_pack,
                                 void* stream) {
    if (n < 0 || d < 1 || k < 1 || k > d) return MK_EINVAL;
    if (index_bytes != 1 && index_bytes != 2) return MK_EINVAL;
    if (index_bytes == 1 && d > 256) return MK_EINVAL;
    if (!mk_banked_supported(k, d)) return MK_EUNSUPPORTED;
    if (bk_pack != nullptr && !mk_packed_supported(k, d)) return MK_EUNSUPPORTED;
    if (n == 0) return MK_OK;
    if (!x || !sp_index) return MK_EINVAL;
    if (bk_pack == nullptr && (!bk_data || !bk_slot)) return MK_EINVAL;
    if (bk_pack != nullptr && (reinterpret_cast<uintptr_t>(bk_pack) & 7)) return MK_EINVAL;
    cudaStream_t st = mk::as_stream(stream);
    uint2* bp = static_cast<uint2*>(bk_pack);
    return index_bytes == 1 ? mk::launch_tile_bank<uint8_t>(x, n, d, k, sp_data, sp_index, bk_data, bk_slot, bp, st)
                            : mk::launch_tile_bank<uint16_t>(x, n, d, k, sp_data, sp_index, bk_data, bk_slot, bp, st);
}
