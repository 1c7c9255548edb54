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

// ------------------------------------------------------------------------------------------------
// kk > 32 (and num_neighbours = -1: all points, dci.py:278-279): the kk smallest of a scanned row, sorted ascending by
// (distance, index).  One block of 1024 threads per scanned query, three phases, all on the row's float64 bit patterns
// (non-negative doubles order like their bits):
//   A  radix select: 8 passes of 8 bits from the top find the kk-th smallest key T and how many keys equal to T belong
//      to the answer (lowest indices first);
//   B  compaction in index order (block-wide prefix sums) of every key < T and the first `need` keys == T into a
//      scratch list of exactly kk (key, index) pairs — index order makes the following sort's ties come out right;
//   C  stable LSD radix sort of the kk pairs, 8-bit digits, ping-pong through the scratch (digits every key shares are
//      skipped: distances of one query share their top bytes), then the rows are written out.
// Hand-written for this path (round 1 used cub::DeviceSegmentedRadixSort over all n keys of every row).
// ------------------------------------------------------------------------------------------------
constexpr int TOPK_THREADS = 1024;

// histogram of an 8-bit digit with one shared-memory atomic per distinct digit and warp
__device__ __forceinline__ void hist_add_warp(unsigned int *hist, bool valid, unsigned int digit) {
    const unsigned int act = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const unsigned int peers = __match_any_sync(act, digit);
    if ((threadIdx.x & 31) == static_cast<unsigned int>(__ffs(peers) - 1)) atomicAdd(&hist[digit], static_cast<unsigned int>(__popc(peers)));
}

// exclusive prefix sum of one value per thread over the block (1024 threads); total returned to every thread
__device__ __forceinline__ unsigned int block_exclusive_scan(unsigned int v, unsigned int *warp_tot /* [32] shared */, unsigned int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const unsigned int w = warp_tot[lane];
        unsigned int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_tot[lane] = winc - w;                 // exclusive offset of warp `lane`
        if (lane == 31) warp_tot[32] = winc;       // block total
    }
    __syncthreads();
    const unsigned int res = warp_tot[warp] + inc - v;
    total = warp_tot[32];
    __syncthreads();                               // warp_tot may be reused by the next call
    return res;
}

__global__ void __launch_bounds__(TOPK_THREADS)
scan_topk_kernel(const double *__restrict__ d2, int n, const int *__restrict__ qlist, int kk, int64_t index_base, unsigned flags,
                 unsigned long long *scratch_key /* [nsub][2][kk] */, int *scratch_idx /* [nsub][2][kk] */,
                 int32_t *__restrict__ out_idx, double *__restrict__ out_dist) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int warp_tot[33];
    __shared__ unsigned int warp_cnt[32][256];       // phase C: per-warp digit counts, then per-warp output offsets
    __shared__ unsigned long long prefix_s;
    __shared__ unsigned int need_s, skip_s;
    const int s = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qrow = qlist ? qlist[s] : s;
    const unsigned long long *row = reinterpret_cast<const unsigned long long *>(d2) + static_cast<int64_t>(s) * n;
    unsigned long long *key_a = scratch_key + static_cast<int64_t>(s) * 2 * kk, *key_b = key_a + kk;
    int *idx_a = scratch_idx + static_cast<int64_t>(s) * 2 * kk, *idx_b = idx_a + kk;

    // ---- A: the kk-th smallest key T (all keys when kk == n: T = max, every key taken)
    unsigned long long T = ~0ull;
    unsigned int need = 0;                           // keys == T that belong to the answer
    if (kk < n) {
        if (tid == 0) { prefix_s = 0ull; need_s = static_cast<unsigned int>(kk); }
        for (int pass = 7; pass >= 0; pass--) {
            for (int i = tid; i < 256; i += TOPK_THREADS) hist[i] = 0u;
            __syncthreads();
            const unsigned long long prefix = prefix_s;
            const int hi_shift = 8 * (pass + 1);
            for (int j0 = 0; j0 < n; j0 += TOPK_THREADS) {
                const int j = j0 + tid;
                bool valid = j < n;
                unsigned long long key = 0ull;
                if (valid) {
                    key = row[j];
                    if (pass < 7) valid = (key >> hi_shift) == (prefix >> hi_shift);
                }
                hist_add_warp(hist, valid, static_cast<unsigned int>(key >> (8 * pass)) & 255u);
            }
            __syncthreads();
            if (tid == 0) {
                unsigned int rem = need_s, b = 0;
                while (b < 255u && hist[b] < rem) { rem -= hist[b]; b++; }
                prefix_s = prefix | (static_cast<unsigned long long>(b) << (8 * pass));
                need_s = rem;
            }
            __syncthreads();
        }
        T = prefix_s;
        need = need_s;
    }
    // ---- B: compaction in index order
    {
        unsigned int out_base = 0, eq_seen = 0;
        for (int j0 = 0; j0 < n; j0 += TOPK_THREADS) {
            const int j = j0 + tid;
            const unsigned long long key = (j < n) ? row[j] : ~0ull;
            const bool less = (j < n) && (kk == n || key < T);
            const bool eq = (j < n) && kk < n && key == T;
            unsigned int total = 0;
            const unsigned int ex = block_exclusive_scan((less ? 1u : 0u) | (eq ? 0x10000u : 0u), warp_tot, total);
            const unsigned int less_before = ex & 0xffffu, eq_before = eq_seen + (ex >> 16);
            const unsigned int eq_taken_before = min(eq_before, need);
            const bool take = less || (eq && eq_before < need);
            if (take) {
                const unsigned int pos = out_base + less_before + (eq_taken_before - min(eq_seen, need));
                key_a[pos] = key;
                idx_a[pos] = j;
            }
            const unsigned int eq_total = eq_seen + (total >> 16);
            out_base += (total & 0xffffu) + (min(eq_total, need) - min(eq_seen, need));
            eq_seen = eq_total;
        }
    }
    __syncthreads();
    // ---- C: stable LSD radix sort of the kk pairs by key (ties keep index order)
    unsigned long long *src_k = key_a, *dst_k = key_b;
    int *src_i = idx_a, *dst_i = idx_b;
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 8 * pass;
        for (int i = tid; i < 256; i += TOPK_THREADS) hist[i] = 0u;
        __syncthreads();
        for (int i0 = 0; i0 < kk; i0 += TOPK_THREADS) {
            const int i = i0 + tid;
            const bool valid = i < kk;
            const unsigned long long key = valid ? src_k[i] : 0ull;
            hist_add_warp(hist, valid, static_cast<unsigned int>(key >> shift) & 255u);
        }
        __syncthreads();
        if (tid < 32) {                              // exclusive scan of the 256 bins (8 per lane); a bin holding every key: skip the pass
            unsigned int v[8], sum = 0, full = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) { v[t] = hist[tid * 8 + t]; sum += v[t]; full |= (v[t] == static_cast<unsigned int>(kk)) ? 1u : 0u; }
            unsigned int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            unsigned int run = inc - sum;
#pragma unroll
            for (int t = 0; t < 8; t++) { hist[tid * 8 + t] = run; run += v[t]; }
            const unsigned int any_full = __ballot_sync(0xffffffffu, full != 0u);
            if (tid == 0) skip_s = any_full ? 1u : 0u;
        }
        __syncthreads();
        if (skip_s) continue;                        // uniform
        for (int i0 = 0; i0 < kk; i0 += TOPK_THREADS) {
            for (int t = tid; t < 32 * 256; t += TOPK_THREADS) (&warp_cnt[0][0])[t] = 0u;
            __syncthreads();
            const int i = i0 + tid;
            const bool valid = i < kk;
            unsigned long long key = 0ull;
            int id = 0;
            if (valid) { key = src_k[i]; id = src_i[i]; }
            const unsigned int digit = static_cast<unsigned int>(key >> shift) & 255u;
            const unsigned int act = __ballot_sync(0xffffffffu, valid);
            unsigned int rank_in_warp = 0;
            if (valid) {
                const unsigned int peers = __match_any_sync(act, digit);
                rank_in_warp = __popc(peers & ((1u << lane) - 1u));
                if (rank_in_warp == 0) warp_cnt[warp][digit] = static_cast<unsigned int>(__popc(peers));
            }
            __syncthreads();
            if (tid < 256) {                         // per digit: running offset over the warps (index order), then over the chunks
                unsigned int run = hist[tid];
#pragma unroll 8
                for (int w = 0; w < 32; w++) {
                    const unsigned int c = warp_cnt[w][tid];
                    warp_cnt[w][tid] = run;
                    run += c;
                }
                hist[tid] = run;
            }
            __syncthreads();
            if (valid) {
                const unsigned int pos = warp_cnt[warp][digit] + rank_in_warp;
                dst_k[pos] = key;
                dst_i[pos] = id;
            }
            __syncthreads();
        }
        { unsigned long long *t = src_k; src_k = dst_k; dst_k = t; }
        { int *t = src_i; src_i = dst_i; dst_i = t; }
    }
    __syncthreads();
    for (int r = tid; r < kk; r += TOPK_THREADS) {
        const double d = __longlong_as_double(static_cast<long long>(src_k[r]));
        out_idx[static_cast<int64_t>(qrow) * kk + r] = static_cast<int32_t>(index_base + src_i[r]);
        out_dist[static_cast<int64_t>(qrow) * kk + r] = (flags & 1u) ? d : sqrt(d);
    }
}

}  // namespace b200
