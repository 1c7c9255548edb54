// Random projection (float64 GEMM on CUDA cores).
#pragma once
#include "common.cuh"

namespace b200 {
// ------------------------------------------------------------------------------------------------
// Random projection (training/training_loop.py:205-212,362-365,379-381): out = rows(float64) @ projector,
// float64 like the reference's np.matmul.  CUDA-core DGEMM: 128 x 128 output tile per block, 256 threads x (8 x 8)
// accumulators, K step 16 through shared memory, next step prefetched into registers.  Every output element is ONE
// sequential FMA chain over e = 0 .. in_dim-1: the result does not depend on tiling or on how the rows are chunked.
// ------------------------------------------------------------------------------------------------
constexpr int PJ_T = 128, PJ_K = 16;
template <typename T>
__global__ void __launch_bounds__(256)
project_kernel(const T *__restrict__ rows, int64_t ld_rows, int n, const double *__restrict__ proj, int in_dim, int dim,
               double *__restrict__ out, int64_t ld_out) {
    __shared__ __align__(16) double as[PJ_K][PJ_T];   // [k][row]
    __shared__ __align__(16) double ps[PJ_K][PJ_T];   // [k][col]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;           // outputs: rows ty*8 .. ty*8+7, columns j*32 + tx*2 + {0,1}, j = 0..3
    const int r0 = blockIdx.y * PJ_T, c0 = blockIdx.x * PJ_T;
    const int lr = tid & 127, lk = (tid >> 7) * 8;    // loader, rows tile: 8 consecutive k of row lr
    const int pk = tid >> 4, pc = (tid & 15) * 2;     // loader, projector tile: k = pk, columns pc + 32 i + {0,1}
    double acc[8][8] = {};
    T ra[8];
    double rp[8];
    auto load = [&](int k0) {
        const int gr = r0 + lr;
        const T *src = rows + static_cast<int64_t>(gr) * ld_rows + k0 + lk;
#pragma unroll
        for (int i = 0; i < 8; i++) ra[i] = (gr < n && k0 + lk + i < in_dim) ? src[i] : T(0);
        const int e = k0 + pk;
        const double *ps_src = proj + static_cast<int64_t>(e) * dim + c0 + pc;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int c = c0 + pc + 32 * i;
            rp[2 * i] = (e < in_dim && c < dim) ? ps_src[32 * i] : 0.0;
            rp[2 * i + 1] = (e < in_dim && c + 1 < dim) ? ps_src[32 * i + 1] : 0.0;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < 8; i++) as[lk + i][lr] = static_cast<double>(ra[i]);
#pragma unroll
        for (int i = 0; i < 4; i++) *reinterpret_cast<double2 *>(&ps[pk][pc + 32 * i]) = make_double2(rp[2 * i], rp[2 * i + 1]);
    };
    load(0);
    stash();
    __syncthreads();
    for (int k0 = 0; k0 < in_dim; k0 += PJ_K) {
        const bool more = k0 + PJ_K < in_dim;
        if (more) load(k0 + PJ_K);
#pragma unroll
        for (int k = 0; k < PJ_K; k++) {
            double a[8], b[8];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const double2 va = *reinterpret_cast<const double2 *>(&as[k][ty * 8 + 2 * j]);
                const double2 vb = *reinterpret_cast<const double2 *>(&ps[k][j * 32 + tx * 2]);
                a[2 * j] = va.x; a[2 * j + 1] = va.y;
                b[2 * j] = vb.x; b[2 * j + 1] = vb.y;
            }
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
        if (more) {
            stash();
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int r = r0 + ty * 8 + i;
        if (r >= n) continue;
        double *dst = out + static_cast<int64_t>(r) * ld_out;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = c0 + j * 32 + tx * 2;
            if (c < dim) dst[c] = acc[i][2 * j];
            if (c + 1 < dim) dst[c + 1] = acc[i][2 * j + 1];
        }
    }
}

}  // namespace b200
