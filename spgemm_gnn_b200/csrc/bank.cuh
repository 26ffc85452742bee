// Bank assignment of one CBSR row (see bank.cu): shared by the stand-alone banking kernel and the
// fused top-k + banking kernel (topk_bank.cu).
#pragma once

#include "common.cuh"

namespace mk {

template <int K>
struct BankMask {
    using type = uint32_t;
};
template <>
struct BankMask<64> {
    using type = unsigned long long;
};

__device__ __forceinline__ int lowest_bit(uint32_t m) { return __ffs(m) - 1; }
__device__ __forceinline__ int lowest_bit(unsigned long long m) { return __ffsll(m) - 1; }
__device__ __forceinline__ int count_bits(uint32_t m) { return __popc(m); }
__device__ __forceinline__ int count_bits(unsigned long long m) { return __popcll(m); }

__device__ __forceinline__ int bank_slot_a(int c) { return 32 * (c >> 3) + (c & 7); }
__device__ __forceinline__ int bank_slot_b(int c, int ra) {
    return 32 * (ra + ((c >> 6) << 3) + (c & 7)) + ((c >> 3) & 7);
}

// One thread, one row: choose the copy (A / B cell) and the position of each of the K entries whose
// columns are col(0..K-1) (a getter: registers in bank.cu, the tile in shared memory in topk_tile.cu);
// leaves one descriptor byte per entry (position | copy << 7) in `my_desc`.
template <int K, typename ColF>
__device__ __forceinline__ void bank_assign(ColF col, uint8_t* __restrict__ my_desc) {
    using Mask = typename BankMask<K>::type;
    constexpr int CAP = K / 8;  // entries per bank when perfectly balanced == steps per neighbour
    // ---- which copy: sequential two-choice on the bank loads (8 x 8-bit counters), tie -> A
    unsigned long long load = 0;
    Mask choice = 0;  // bit e set: entry e uses copy B
#pragma unroll
    for (int e = 0; e < K; ++e) {
        const int a = col(e) & 7, b = (col(e) >> 3) & 7;
        const int la = static_cast<int>((load >> (8 * a)) & 255), lb = static_cast<int>((load >> (8 * b)) & 255);
        const bool pick_b = lb < la;
        load += 1ull << (8 * (pick_b ? b : a));
        choice |= static_cast<Mask>(pick_b ? 1 : 0) << e;
    }
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int e = 0; e < K; ++e) {
            const int a = col(e) & 7, b = (col(e) >> 3) & 7;
            const bool on_b = (choice >> e) & 1;
            const int cur = on_b ? b : a, alt = on_b ? a : b;
            const int lc = static_cast<int>((load >> (8 * cur)) & 255), la = static_cast<int>((load >> (8 * alt)) & 255);
            if (cur != alt && lc > CAP && la + 1 < lc) {
                load += (1ull << (8 * alt)) - (1ull << (8 * cur));
                choice ^= static_cast<Mask>(1) << e;
            }
        }
    }

    // ---- members of every bank, bit-sliced: the three bits of every entry's bank as masks over the entries,
    //      then bank x = the entries whose three bits spell x (3 K + 16 operations instead of 24 K)
    Mask bit0 = 0, bit1 = 0, bit2 = 0;
#pragma unroll
    for (int e = 0; e < K; ++e) {
        const int cls = ((choice >> e) & 1) ? ((col(e) >> 3) & 7) : (col(e) & 7);
        bit0 |= static_cast<Mask>(cls & 1) << e;
        bit1 |= static_cast<Mask>((cls >> 1) & 1) << e;
        bit2 |= static_cast<Mask>((cls >> 2) & 1) << e;
    }
    const Mask all = static_cast<Mask>(K >= 64 ? ~0ull : ((1ull << (K & 63)) - 1ull));
    Mask member[8];
#pragma unroll
    for (int x = 0; x < 8; ++x)
        member[x] = ((x & 1) ? bit0 : ~bit0) & ((x & 2) ? bit1 : ~bit1) & ((x & 4) ? bit2 : ~bit2) & all;

    // ---- positions: step q takes one entry of every non-empty bank, then tops up from the
    //      fullest banks; lane t of a group reads positions [CAP*t, CAP*t + CAP)
    auto emit = [&](int e, int q, int t) {
        my_desc[e] = static_cast<uint8_t>((t * CAP + q) | (((choice >> e) & 1) ? 0x80 : 0));
    };
#pragma unroll 1
    for (int q = 0; q < CAP; ++q) {
        int taken = 0;
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            if (member[x] != 0) {
                const int e = lowest_bit(member[x]);
                member[x] &= member[x] - 1;
                emit(e, q, taken++);
            }
        }
        while (taken < 8) {
            int best = 0, best_cnt = -1;
#pragma unroll
            for (int x = 0; x < 8; ++x) {
                const int cnt = count_bits(member[x]);
                if (cnt > best_cnt) { best_cnt = cnt; best = x; }
            }
            int e = 0;
#pragma unroll
            for (int x = 0; x < 8; ++x) {
                if (x == best) {
                    e = lowest_bit(member[x]);
                    member[x] &= member[x] - 1;
                }
            }
            emit(e, q, taken++);
        }
    }
}

}  // namespace mk
