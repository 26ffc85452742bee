// Shared device/host helpers of the maxk_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "maxk_b200.h"

namespace mk {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- error plumbing (no exceptions across the C ABI) -----------------------------------
void set_cuda_error(cudaError_t e, const char* where);

#define MK_CUDA_TRY(expr)                                  \
    do {                                                   \
        cudaError_t _e = (expr);                           \
        if (_e != cudaSuccess) {                           \
            ::mk::set_cuda_error(_e, #expr);               \
            return MK_ECUDA;                               \
        }                                                  \
    } while (0)

#define MK_LAUNCH_CHECK(name)                              \
    do {                                                   \
        cudaError_t _e = cudaGetLastError();               \
        if (_e != cudaSuccess) {                           \
            ::mk::set_cuda_error(_e, name);                \
            return MK_ECUDA;                               \
        }                                                  \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- memory access flavours ----------------------------------------------------------------
// Streaming read of data touched once (dense inputs, edge arrays): keep it out of L1.
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ int ld_stream_i1(const int* p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
// Streaming store of an output written once.
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
// Vector float reduction in L2 (sm_90+): SASS REDG.E.ADD.F32x4.
__device__ __forceinline__ void red_add_f4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}
__device__ __forceinline__ void red_add_f2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_f1(float* p, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- ordering of the forward's shared-memory accumulation --------------------------------------
// The forward kernels add into shared-memory cells with plain LDS / FFMA / STS.  The lanes that work
// on ONE neighbour hit distinct cells (columns of a CBSR row are distinct), but the same lanes
// handle another neighbour in the next step and may then hit a cell a DIFFERENT lane updated
// before.  That read-after-write between lanes needs an ordering the programming model only gives
// through __syncwarp, so every neighbour step ends with one.  Measured on a B200 (Reddit shape,
// profiles/r2/sync_modes.log; k = 8 / 16 / 32 / 64 forward ms):
//   MK_SYNC_MODE 0  no fence (round 1)                          1.600 / 2.455 / 2.896 / 6.321
//   MK_SYNC_MODE 1  __syncwarp(group mask) inside the branch    1.995 / 3.212 / 3.482 / 6.053
//   MK_SYNC_MODE 2  __syncwarp() by all 32 lanes, placed after  1.604 / 2.456 / 2.916 / 6.328
//                   the group-uniform `if (ok)` has reconverged  (the default)
// Mode 2 compiles to no WARPSYNC at all: ptxas proves the warp converged at that point (BSYNC),
// where in-order issue already gives the ordering -- the fence is free and the guarantee is in the
// source instead of in an assumption.  Mode 1 needs WARPSYNC + the collective machinery per group.
#ifndef MK_SYNC_MODE
#define MK_SYNC_MODE 2
#endif
__device__ __forceinline__ void accum_fence_group(unsigned mask) {
#if MK_SYNC_MODE == 1
    __syncwarp(mask);
#endif
}
__device__ __forceinline__ void accum_fence_warp() {
#if MK_SYNC_MODE == 2
    __syncwarp();
#endif
}

// ---- CBSR index access -------------------------------------------------------------------
template <typename IdxT>
struct Idx4;  // four consecutive column ids in one load
template <>
struct Idx4<uint8_t> {
    using vec = uchar4;
};
template <>
struct Idx4<uint16_t> {
    using vec = ushort4;
};

}  // namespace mk
