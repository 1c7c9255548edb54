// Exact float64 CUDA-core scan: overflowed second-pass lists and k > 32.
#pragma once
#include "common.cuh"

namespace b200 {
// ------------------------------------------------------------------------------------------------
// Exact float64 scan (CUDA cores).  Used for uncertified queries and for k > 16.
//   scan_dist_kernel : d2[s][j] = sum_e (q[list[s]][e] - x[j][e])^2     (32 queries x 64 pool rows per block)
//   scan_select_kernel: kk passes of lexicographic (d2, index) arg-min  (kk <= 32)
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_TQ = 64, SCAN_TX = 64, SCAN_TK = 16;

// 64 queries x 64 pool rows per block, 4 x 4 outputs per thread, operands staged k-major in shared memory so a thread
// fetches its four query values and four pool values with two 16-byte loads each (float64 pipe bound, not LDS bound).
template <typename TX, typename TQ>
__global__ void __launch_bounds__(256)
scan_dist_kernel(const TX *__restrict__ x, int64_t ld_x, int n, const TQ *__restrict__ qmat, int64_t ld_q,
                 const int *__restrict__ qlist, int nsub, int dim, double *__restrict__ d2) {
    __shared__ __align__(16) double qs[SCAN_TK][SCAN_TQ + 4];
    __shared__ __align__(16) double xs[SCAN_TK][SCAN_TX + 4];
    const int tx = threadIdx.x & 15;    // pool rows 4*tx .. 4*tx+3
    const int ty = threadIdx.x >> 4;    // queries   4*ty .. 4*ty+3
    const int x0 = blockIdx.x * SCAN_TX;
    const int s0 = blockIdx.y * SCAN_TQ;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < dim; k0 += SCAN_TK) {
        for (int i = threadIdx.x; i < SCAN_TQ * SCAN_TK; i += 256) {
            const int r = i / SCAN_TK, c = i % SCAN_TK;
            const int s = s0 + r, e = k0 + c;
            double v = 0.0;
            if (s < nsub && e < dim) {
                const int qrow = qlist ? qlist[s] : s;
                v = static_cast<double>(qmat[static_cast<int64_t>(qrow) * ld_q + e]);
            }
            qs[c][r] = v;
        }
        for (int i = threadIdx.x; i < SCAN_TX * SCAN_TK; i += 256) {
            const int r = i / SCAN_TK, c = i % SCAN_TK;
            const int j = x0 + r, e = k0 + c;
            xs[c][r] = (j < n && e < dim) ? static_cast<double>(x[static_cast<int64_t>(j) * ld_x + e]) : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < SCAN_TK; c++) {
            const double2 qa = *reinterpret_cast<const double2 *>(&qs[c][ty * 4]);
            const double2 qb = *reinterpret_cast<const double2 *>(&qs[c][ty * 4 + 2]);
            const double2 xa = *reinterpret_cast<const double2 *>(&xs[c][tx * 4]);
            const double2 xb = *reinterpret_cast<const double2 *>(&xs[c][tx * 4 + 2]);
            const double qv[4] = {qa.x, qa.y, qb.x, qb.y};
            const double xv[4] = {xa.x, xa.y, xb.x, xb.y};
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const double df = qv[a] - xv[b];
                    acc[a][b] = fma(df, df, acc[a][b]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int s = s0 + ty * 4 + a;
        if (s >= nsub) continue;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int j = x0 + tx * 4 + b;
            if (j < n) d2[static_cast<int64_t>(s) * n + j] = acc[a][b];
        }
    }
}

// One block per scanned query; kk sequential block-wide lexicographic arg-min passes over d2[s][0..n).
__global__ void __launch_bounds__(256)
scan_select_kernel(const double *__restrict__ d2, int n, const int *__restrict__ qlist, int kk, int64_t index_base,
                   unsigned flags, int32_t *__restrict__ out_idx, double *__restrict__ out_dist) {
    __shared__ double sd[8];
    __shared__ int si[8];
    __shared__ double last_d_s;
    __shared__ int last_i_s;
    const int s = blockIdx.x;
    const int qrow = qlist ? qlist[s] : s;
    const double *row = d2 + static_cast<int64_t>(s) * n;
    if (threadIdx.x == 0) { last_d_s = -1.0; last_i_s = -1; }
    __syncthreads();
    for (int r = 0; r < kk; r++) {
        const double ld = last_d_s;
        const int li = last_i_s;
        double bd = DBL_MAX;
        int bi = 0x7fffffff;
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const double d = row[j];
            const bool after = (d > ld) || (d == ld && j > li);          // strictly after the last pick
            if (after && (d < bd || (d == bd && j < bi))) { bd = d; bi = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        __syncthreads();   // everyone has read last_*_s
        if ((threadIdx.x & 31) == 0) { sd[threadIdx.x >> 5] = bd; si[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; w++)
                if (sd[w] < bd || (sd[w] == bd && si[w] < bi)) { bd = sd[w]; bi = si[w]; }
            last_d_s = bd;
            last_i_s = bi;
            out_idx[static_cast<int64_t>(qrow) * kk + r] = static_cast<int32_t>(index_base + bi);
            out_dist[static_cast<int64_t>(qrow) * kk + r] = (flags & 1u) ? bd : sqrt(bd);
        }
        __syncthreads();
    }
}

// fill segment offsets / iota values for the segmented sort used when kk > 32
__global__ void iota_kernel(int *__restrict__ v, int64_t total, int n) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        v[i] = static_cast<int>(i % n);
}
__global__ void scatter_sorted_kernel(const double *__restrict__ sd, const int *__restrict__ sv, int n, const int *__restrict__ qlist,
                                      int nsub, int kk, int64_t index_base, unsigned flags,
                                      int32_t *__restrict__ out_idx, double *__restrict__ out_dist) {
    const int64_t total = static_cast<int64_t>(nsub) * kk;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int s = static_cast<int>(i / kk), r = static_cast<int>(i % kk);
        const int qrow = qlist ? qlist[s] : s;
        const double d = sd[static_cast<int64_t>(s) * n + r];
        out_idx[static_cast<int64_t>(qrow) * kk + r] = static_cast<int32_t>(index_base + sv[static_cast<int64_t>(s) * n + r]);
        out_dist[static_cast<int64_t>(qrow) * kk + r] = (flags & 1u) ? d : sqrt(d);
    }
}

}  // namespace b200
