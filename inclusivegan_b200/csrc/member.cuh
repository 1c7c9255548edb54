// Ball membership for the k-NN precision/recall metric.
#pragma once
#include "common.cuh"
#include "dist.cuh"
#include "rerank.cuh"

namespace b200 {
// ------------------------------------------------------------------------------------------------
// Ball membership (k-NN precision/recall metric, reference metrics/precision_recall.py:96-134):
// is query i inside ANY ball B(x_j, r_j)?   Filter on the tensor cores, decide exactly in float64.
//   necessary condition from the BF16 pass:  ||q~ - x~_j|| <= r_j + ||x_j - x~_j|| + ||q_i - q~_i||
//   =>  s~_ij - (r_j + e_j)^2  <=  -||q~_i||^2 + 2 Rmax e_i + e_i^2 + eps      (Rmax = max_j (r_j + e_j))
// so the collect-mode distance kernel runs with column terms  ||x~_j||^2 - (r_j + e_j)^2  and row thresholds.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ball_colterm_kernel(const float *__restrict__ xnorm_bf, const float *__restrict__ x_err, const double *__restrict__ radius2, int n,
                    float *__restrict__ colterm, unsigned int *__restrict__ rmax_bits) {
    float mx = 0.f;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const double r2 = radius2[j];
        const double re = (r2 > 0.0 ? sqrt(r2) : 0.0) * (1.0 + 1e-12) + static_cast<double>(x_err[j]);
        colterm[j] = __double2float_rd(static_cast<double>(xnorm_bf[j]) - re * re);     // rounded DOWN: keeps more candidates
        mx = fmaxf(mx, __double2float_ru(re));
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(rmax_bits, __float_as_uint(mx));
}

__global__ void __launch_bounds__(256)
ball_rowthr_kernel(const float *__restrict__ qnorm_bf, const float *__restrict__ q_err, const unsigned int *__restrict__ max_xnorm_bf_bits,
                   const unsigned int *__restrict__ rmax_bits, int kp, int nq, float *__restrict__ thr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const double qn = static_cast<double>(qnorm_bf[i]);
    const double xn = static_cast<double>(__uint_as_float(*max_xnorm_bf_bits));
    const double K = static_cast<double>(kp);
    const double eps = (K + 8.0) * 2.4e-7 * sqrt(qn * xn) * 1.001 + (K / 16.0 + 8.0) * 1.2e-7 * (xn + qn);
    const double e = static_cast<double>(q_err[i]) * (1.0 + 1e-6);
    const double rmax = static_cast<double>(__uint_as_float(*rmax_bits));
    const double t = -qn + 2.0 * rmax * e + e * e + eps;
    thr[i] = __double2float_ru(t + 1e-6 * fabs(t));
}

struct MemberParams {
    const int *coll_count;       // [nq]
    const int *coll_idx;         // [nq][cap]
    int cap;
    const double *radius2;       // [n]
    int dim;
    int64_t ld_x, ld_q;
    unsigned char *out_member;   // [nq], OR-ed into (several shards / radii sets share the buffer)
    int *overflow_count;
    int *overflow_list;          // queries whose list overflowed without a witness: exact scan
};

template <typename TX, typename TQ>
__global__ void __launch_bounds__(128)
ball_member_kernel(const TX *__restrict__ x, const TQ *__restrict__ qmat, const MemberParams p) {
    __shared__ int found_s;
    const int q = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total = p.coll_count[q];
    const int cnt = min(total, p.cap);
    if (tid == 0) found_s = 0;
    __syncthreads();
    const TQ *qr = qmat + static_cast<int64_t>(q) * p.ld_q;
    // four candidates at a time (one per warp, canonical summation order), stop at the first witness
    for (int c0 = 0; c0 < cnt; c0 += 4) {
        const int c = c0 + warp;
        if (c < cnt) {
            const int j = p.coll_idx[static_cast<int64_t>(q) * p.cap + c];
            const double dd = canon_d2_warp(x + static_cast<int64_t>(j) * p.ld_x, qr, p.dim, lane);
            if (lane == 0 && dd <= p.radius2[j]) found_s = 1;
        }
        __syncthreads();
        if (found_s) break;
    }
    if (tid == 0) {
        if (found_s) p.out_member[q] = 1;
        else if (total > p.cap) {
            const int o = atomicAdd(p.overflow_count, 1);
            p.overflow_list[o] = q;
        }
    }
}

// exact scan fallback: d2[s][j] (from scan_dist_kernel) against radius2[j]
__global__ void __launch_bounds__(256)
scan_member_kernel(const double *__restrict__ d2, int n, const int *__restrict__ qlist, const double *__restrict__ radius2,
                   unsigned char *__restrict__ out_member) {
    __shared__ int any_s;
    const int s = blockIdx.x;
    if (threadIdx.x == 0) any_s = 0;
    __syncthreads();
    const double *row = d2 + static_cast<int64_t>(s) * n;
    int hit = 0;
    for (int j = threadIdx.x; j < n && !hit; j += blockDim.x) hit = row[j] <= radius2[j];
    if (hit) any_s = 1;
    __syncthreads();
    if (threadIdx.x == 0 && any_s) out_member[qlist ? qlist[s] : s] = 1;
}

}  // namespace b200
