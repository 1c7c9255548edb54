// Kernel 1 — centre, convert to BF16, norms and exact rounding errors (HBM-bound).
#pragma once
#include "common.cuh"

namespace b200 {
// ------------------------------------------------------------------------------------------------
// Kernel 1: convert + norms.  One warp per row, 8 elements (one 16-byte BF16 store) per lane per step.
// Algorithmic bytes per row: dim * (sizeof(T) + 2) + 8.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void load8(const T *p, double (&d)[8]);

template <>
__device__ __forceinline__ void load8<double>(const double *p, double (&d)[8]) {
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
    const double2 v0 = __ldcs(p2), v1 = __ldcs(p2 + 1), v2 = __ldcs(p2 + 2), v3 = __ldcs(p2 + 3);
    d[0] = v0.x; d[1] = v0.y; d[2] = v1.x; d[3] = v1.y; d[4] = v2.x; d[5] = v2.y; d[6] = v3.x; d[7] = v3.y;
}
template <>
__device__ __forceinline__ void load8<float>(const float *p, double (&d)[8]) {
    const float4 *p4 = reinterpret_cast<const float4 *>(p);
    const float4 v0 = __ldcs(p4), v1 = __ldcs(p4 + 1);
    d[0] = v0.x; d[1] = v0.y; d[2] = v0.z; d[3] = v0.w; d[4] = v1.x; d[5] = v1.y; d[6] = v1.z; d[7] = v1.w;
}

// 256-bit flavours (LDG.256, sm_100): one full 32-byte sector per lane and request — rows must be 32-byte aligned
template <typename T>
__device__ __forceinline__ void load8_wide(const T *p, double (&d)[8]);
template <>
__device__ __forceinline__ void load8_wide<double>(const double *p, double (&d)[8]) {
    asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(d[0]), "=d"(d[1]), "=d"(d[2]), "=d"(d[3]) : "l"(p));
    asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(d[4]), "=d"(d[5]), "=d"(d[6]), "=d"(d[7]) : "l"(p + 4));
}
template <>
__device__ __forceinline__ void load8_wide<float>(const float *p, double (&d)[8]) {
    float f[8];
    asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "l"(p));
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = f[i];
}

// raw 8-element group of a row (no conversion: a float->double conversion right behind its load would make the in-order
// warp wait for that load before issuing the next one)
template <typename T>
__device__ __forceinline__ void load8_raw(const T *p, T (&r)[8], bool wide);
template <>
__device__ __forceinline__ void load8_raw<double>(const double *p, double (&r)[8], bool wide) {
    if (wide) {
        asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r[0]), "=d"(r[1]), "=d"(r[2]), "=d"(r[3]) : "l"(p));
        asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r[4]), "=d"(r[5]), "=d"(r[6]), "=d"(r[7]) : "l"(p + 4));
    } else {
        const double2 *p2 = reinterpret_cast<const double2 *>(p);
        const double2 v0 = __ldcs(p2), v1 = __ldcs(p2 + 1), v2 = __ldcs(p2 + 2), v3 = __ldcs(p2 + 3);
        r[0] = v0.x; r[1] = v0.y; r[2] = v1.x; r[3] = v1.y; r[4] = v2.x; r[5] = v2.y; r[6] = v3.x; r[7] = v3.y;
    }
}
template <>
__device__ __forceinline__ void load8_raw<float>(const float *p, float (&r)[8], bool wide) {
    if (wide) {
        asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p));
    } else {
        const float4 *p4 = reinterpret_cast<const float4 *>(p);
        const float4 v0 = __ldcs(p4), v1 = __ldcs(p4 + 1);
        r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w; r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
    }
}

// 8 doubles through the read-only cached path (the column means: 8*dim bytes, L1/L2 resident)
__device__ __forceinline__ void load8_cached(const double *p, double (&d)[8]) {
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
    const double2 v0 = __ldg(p2), v1 = __ldg(p2 + 1), v2 = __ldg(p2 + 2), v3 = __ldg(p2 + 3);
    d[0] = v0.x; d[1] = v0.y; d[2] = v1.x; d[3] = v1.y; d[4] = v2.x; d[5] = v2.y; d[6] = v3.x; d[7] = v3.y;
}

// Column sums of the pool (float64 atomics): the pool mean is subtracted from pool AND queries before the BF16
// rounding.  Translation changes no distance, but it removes a common offset from the norms the rounding error is
// proportional to (features with a large mean otherwise certify nothing in the first pass).
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T *__restrict__ src, int64_t n, int64_t ld, int dim, double *__restrict__ sums) {
    const int rows_per_block = 256;
    // row blocks on grid.x (up to 2^31-1 blocks), column blocks on grid.y
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
    const int64_t r1 = min(r0 + rows_per_block, n);
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= dim) return;
    double a0 = 0.0, a1 = 0.0;
    int64_t r = r0;
    for (; r + 1 < r1; r += 2) {
        a0 += static_cast<double>(src[r * ld + c]);
        a1 += static_cast<double>(src[(r + 1) * ld + c]);
    }
    if (r < r1) a0 += static_cast<double>(src[r * ld + c]);
    atomicAdd(sums + c, a0 + a1);
}
__global__ void scale_kernel(double *__restrict__ v, int dim, double f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < dim) v[i] *= f;
}

// Outputs per row: the BF16 row x~ of (x - mu) (zero padded to kp), ||x~||^2 (fp32 sum of the exact squares of the
// rounded values) and err = ||(x - mu) - x~|| rounded up - the EXACT size of the rounding perturbation, which is what
// the exactness certificate needs (a worst-case 2^-9 ||x|| bound is ~2.5x looser).  Grid-wide maxima of both are kept
// as float bit patterns (non-negative floats order like unsigned ints).
// vec != 0 requires: dim % 8 == 0 (so kp == dim), src rows 16-byte aligned; vec == 2: 32-byte aligned (256-bit loads).
template <typename T>
__global__ void __launch_bounds__(256)
convert_norm_kernel(const T *__restrict__ src, const double *__restrict__ mu, int64_t n, int64_t ld, int dim, int kp, int vec,
                    __nv_bfloat16 *__restrict__ dst, float *__restrict__ norm_bf, float *__restrict__ err_out,
                    unsigned int *__restrict__ max_norm_bf_bits, unsigned int *__restrict__ max_err_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_per_grid = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
    float mx_bf = 0.f, mx_er = 0.f;
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps_per_grid) {
        const T *s = src + row * ld;
        __nv_bfloat16 *d = dst + row * kp;
        float acc = 0.f;      // sum of squares of the ROUNDED values (exact products, fp32 accumulation)
        double er = 0.0;      // sum of squares of (x - x~)
        if (vec) {
            const int groups = dim >> 3;
            // NG 8-element groups per lane per step: all loads of a step are issued before the first use (NG x 64 bytes of a
            // float64 row in flight per lane; the means come from L1/L2)
            constexpr int NG = sizeof(T) == 8 ? 3 : 4;      // 192 / 128 bytes of the row in flight per lane
            int g = lane;
            for (; g + 32 * (NG - 1) < groups; g += 32 * NG) {
                T raw[NG][8];
#pragma unroll
                for (int h = 0; h < NG; h++) load8_raw<T>(s + ((g + 32 * h) << 3), raw[h], vec == 2);
                if constexpr (sizeof(T) == 4) {
#pragma unroll
                    for (int h = 0; h < NG; h++)
#pragma unroll
                        for (int i = 0; i < 8; i++) keep(raw[h][i]);      // every load issued before the first conversion
                }
                // one group at a time from here on (the raw values of the others wait in registers)
#pragma unroll
                for (int h = 0; h < NG; h++) {
                    double v[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) v[i] = static_cast<double>(raw[h][i]);
                    if (mu) {
                        double m0[8];
                        load8_cached(mu + ((g + 32 * h) << 3), m0);
#pragma unroll
                        for (int i = 0; i < 8; i++) v[i] -= m0[i];
                    }
                    __nv_bfloat162 b[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        b[i] = __floats2bfloat162_rn(static_cast<float>(v[2 * i]), static_cast<float>(v[2 * i + 1]));
                        const float lo = __low2float(b[i]), hi = __high2float(b[i]);
                        acc = fmaf(lo, lo, acc);
                        acc = fmaf(hi, hi, acc);
                        const double e0 = v[2 * i] - static_cast<double>(lo), e1 = v[2 * i + 1] - static_cast<double>(hi);
                        er = fma(e0, e0, er);
                        er = fma(e1, e1, er);
                    }
                    uint4 out;
                    out.x = *reinterpret_cast<uint32_t *>(&b[0]);
                    out.y = *reinterpret_cast<uint32_t *>(&b[1]);
                    out.z = *reinterpret_cast<uint32_t *>(&b[2]);
                    out.w = *reinterpret_cast<uint32_t *>(&b[3]);
                    *reinterpret_cast<uint4 *>(d + ((g + 32 * h) << 3)) = out;
                }
            }
            for (; g < groups; g += 32) {
                double v[8];
                load8<T>(s + (g << 3), v);
                if (mu) {
                    double m0[8];
                    load8_cached(mu + (g << 3), m0);
#pragma unroll
                    for (int i = 0; i < 8; i++) v[i] -= m0[i];
                }
                __nv_bfloat162 b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    b[i] = __floats2bfloat162_rn(static_cast<float>(v[2 * i]), static_cast<float>(v[2 * i + 1]));
                    const float lo = __low2float(b[i]), hi = __high2float(b[i]);
                    acc = fmaf(lo, lo, acc);
                    acc = fmaf(hi, hi, acc);
                    const double e0 = v[2 * i] - static_cast<double>(lo), e1 = v[2 * i + 1] - static_cast<double>(hi);
                    er = fma(e0, e0, er);
                    er = fma(e1, e1, er);
                }
                uint4 out;
                out.x = *reinterpret_cast<uint32_t *>(&b[0]);
                out.y = *reinterpret_cast<uint32_t *>(&b[1]);
                out.z = *reinterpret_cast<uint32_t *>(&b[2]);
                out.w = *reinterpret_cast<uint32_t *>(&b[3]);
                *reinterpret_cast<uint4 *>(d + (g << 3)) = out;
            }
        } else {
            for (int e = lane; e < kp; e += 32) {
                double v = 0.0;
                if (e < dim) v = static_cast<double>(s[e]) - (mu ? __ldg(mu + e) : 0.0);
                const __nv_bfloat16 b = __float2bfloat16_rn(static_cast<float>(v));
                const float fb = __bfloat162float(b);
                acc = fmaf(fb, fb, acc);
                const double e0 = v - static_cast<double>(fb);
                er = fma(e0, e0, er);
                d[e] = b;
            }
        }
        acc = warp_sum(acc);
        er = warp_sum(er);
        const float erf = __double2float_ru(sqrt(er) * (1.0 + 1e-9));
        if (lane == 0) {
            norm_bf[row] = acc;
            err_out[row] = erf;
        }
        mx_bf = fmaxf(mx_bf, acc);
        mx_er = fmaxf(mx_er, erf);
    }
    if (lane == 0) {
        // (read first: after the first wave almost every warp finds a larger maximum already there and skips the atomic —
        // tens of thousands of same-address atomics would otherwise serialise at the end of a short launch)
        if (mx_bf > 0.f && __float_as_uint(mx_bf) > *reinterpret_cast<volatile unsigned int *>(max_norm_bf_bits)) atomicMax(max_norm_bf_bits, __float_as_uint(mx_bf));
        if (mx_er > 0.f && __float_as_uint(mx_er) > *reinterpret_cast<volatile unsigned int *>(max_err_bits)) atomicMax(max_err_bits, __float_as_uint(mx_er));
    }
}

}  // namespace b200
