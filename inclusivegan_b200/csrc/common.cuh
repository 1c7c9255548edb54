// Tile geometry and small device helpers shared by every kernel of libb200knn.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>
#include "ptx.cuh"

namespace b200 {
// ------------------------------------------------------------------------------------------------
// tile geometry of the distance kernel
// ------------------------------------------------------------------------------------------------
constexpr int BM = 128;          // query rows per CTA tile (UMMA M, one TMEM lane per row)
constexpr int BN = 256;          // pool rows per tile (UMMA N, one TMEM fp32 column per row)
constexpr int BK = 64;           // K elements per pipeline stage: 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;       // K per tcgen05.mma for 16-bit inputs
constexpr int TMEM_COLS = 512;               // two 128 x 256 fp32 accumulators
constexpr int DIST_THREADS = 192;            // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int MAX_KEYS = 4096;               // shortlist entries per query the rerank kernel can merge (slots * C)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// system-scope acquire load / release store: flags written by a peer GPU over NVLink (exchange kernels)
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Opaque use of a register: everything loaded into the arguments must be issued before the compiler may start
// consuming them (it otherwise re-fuses "load all, then convert all" into load/convert pairs that reuse three
// registers, i.e. three loads in flight instead of twenty-four).
__device__ __forceinline__ void keep(float &v) { asm volatile("" : "+f"(v)); }
__device__ __forceinline__ void keep(double &v) { asm volatile("" : "+d"(v)); }

}  // namespace b200
