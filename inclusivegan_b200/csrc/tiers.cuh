// Precision tiers of the tensor pass: operand conversion for split BF16 (bf16x3) and TF32.
//
//   bf16    x~ = bf16(x)                                  1 MMA per product      rounding ~2^-9 per element
//   bf16x3  x^ = hi + lo, hi = bf16(x), lo = bf16(x - hi)  3 MMAs (hi.hi + hi.lo + lo.hi; lo.lo is bounded, not computed)
//                                                                                 rounding ~2^-17 per element
//   tf32    x^ = tf32(x) in an fp32 container              1 MMA at half the BF16 rate, rounding ~2^-12 per element
//
// Whatever the tier, the convert kernel measures what the certificate needs EXACTLY, per row: ||x^||^2 (float64 sum, one
// rounding to fp32), ||x - x^|| (rounded up) and, for the split tier, ||lo|| (the dropped lo.lo term is at most
// ||q_lo|| ||x_lo||).  HBM-bound like convert_norm_kernel: one warp per row, 8 elements per lane and step.
#pragma once
#include "common.cuh"
#include "convert.cuh"

namespace b200 {

__device__ __forceinline__ float tf32_round(float v) {      // round to nearest (ties away), 10 explicit mantissa bits
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}

// TIER 1: hi/lo BF16 rows (hi may be nullptr when the BF16 tier has written it already).  TIER 2: tf32 rows (fp32).
// src rows may be addressed through a list (second pass of the mixed mode: only the uncertified queries are split).
template <typename T, int TIER>
__global__ void __launch_bounds__(256)
convert_tier_kernel(const T *__restrict__ src, const double *__restrict__ mu, int64_t n, int64_t ld, int dim, int kp,
                    __nv_bfloat16 *__restrict__ dst_hi, __nv_bfloat16 *__restrict__ dst_lo, float *__restrict__ dst_tf,
                    float *__restrict__ norm_out, float *__restrict__ err_out, float *__restrict__ lonorm_out,
                    unsigned int *__restrict__ max_bits /* [3]: norm, err, lo norm */) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_per_grid = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
    float mx_n = 0.f, mx_e = 0.f, mx_l = 0.f;
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps_per_grid) {
        const T *s = src + row * ld;
        double nn = 0.0, er = 0.0, ln = 0.0;
        for (int e = lane; e < kp; e += 32) {
            double v = 0.0;
            if (e < dim) v = static_cast<double>(s[e]) - (mu ? __ldg(mu + e) : 0.0);
            if constexpr (TIER == 1) {
                const __nv_bfloat16 h = __float2bfloat16_rn(static_cast<float>(v));
                const double hv = static_cast<double>(__bfloat162float(h));
                const __nv_bfloat16 l = __float2bfloat16_rn(static_cast<float>(v - hv));
                const double lv = static_cast<double>(__bfloat162float(l));
                const double xv = hv + lv;
                nn = fma(xv, xv, nn);
                er = fma(v - xv, v - xv, er);
                ln = fma(lv, lv, ln);
                if (dst_hi) dst_hi[row * kp + e] = h;
                dst_lo[row * kp + e] = l;
            } else {
                const float t = tf32_round(static_cast<float>(v));
                const double tv = static_cast<double>(t);
                nn = fma(tv, tv, nn);
                er = fma(v - tv, v - tv, er);
                dst_tf[row * kp + e] = t;
            }
        }
        nn = warp_sum(nn);
        er = warp_sum(er);
        ln = warp_sum(ln);
        const float nf = __double2float_ru(nn);
        const float ef = __double2float_ru(sqrt(er) * (1.0 + 1e-9));
        const float lf = __double2float_ru(sqrt(ln) * (1.0 + 1e-9));
        if (lane == 0) {
            norm_out[row] = nf;
            err_out[row] = ef;
            if (TIER == 1) lonorm_out[row] = lf;
        }
        mx_n = fmaxf(mx_n, nf);
        mx_e = fmaxf(mx_e, ef);
        mx_l = fmaxf(mx_l, lf);
    }
    if (lane == 0) {
        if (mx_n > 0.f) atomicMax(max_bits, __float_as_uint(mx_n));
        if (mx_e > 0.f) atomicMax(max_bits + 1, __float_as_uint(mx_e));
        if (TIER == 1 && mx_l > 0.f) atomicMax(max_bits + 2, __float_as_uint(mx_l));
    }
}

}  // namespace b200
