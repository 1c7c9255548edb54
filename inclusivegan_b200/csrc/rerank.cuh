// Kernel 3 — shortlist merge, exact float64 re-rank in one canonical summation order, certificate; second-pass list re-rank.
#pragma once
#include "common.cuh"

namespace b200 {
// ------------------------------------------------------------------------------------------------
// Kernel 3: shortlist merge + exact re-rank + certificate.  One block (128 threads) per query.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_order_bits(float f) {   // monotone float -> uint
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_bits(uint32_t b) {
    return __uint_as_float((b & 0x80000000u) ? (b & 0x7fffffffu) : ~b);
}

// Canonical exact squared distance: ONE summation order for every code path that emits a distance (first-pass re-rank,
// second-pass list re-rank, ball membership), so a result does not depend on how the pool is sharded, on the batch
// size, or on whether the query needed the second pass.  128 virtual lanes: lane v sums the elements e = v + 128 i,
// even i into one accumulator and odd i into another; xor-shuffle tree inside each of the 4 warps; the 4 warp sums are
// added in order.  Executed by threads 0..127 of the block; the value is returned to every thread.
template <int RB = 16, typename TX, typename TQ>    // RB (even) elements per lane and step: 2 RB loads in flight before any arithmetic
__device__ __forceinline__ void canon_d2_lanes(const TX *__restrict__ xr, const TQ *__restrict__ qr, int dim, int v, double &a0, double &a1) {
    for (int base = 0; base < dim; base += RB * 128) {
        TQ qraw[RB];
        TX xraw[RB];
#pragma unroll
        for (int i = 0; i < RB; i++) {
            const int e = base + v + i * 128;
            qraw[i] = (e < dim) ? qr[e] : TQ(0);     // past the end: 0 - 0 adds nothing
            xraw[i] = (e < dim) ? xr[e] : TX(0);
        }
#pragma unroll
        for (int i = 0; i < RB; i++) { keep(qraw[i]); keep(xraw[i]); }
#pragma unroll
        for (int i = 0; i < RB; i += 2) {
            const double d0 = static_cast<double>(qraw[i]) - static_cast<double>(xraw[i]);
            const double d1 = static_cast<double>(qraw[i + 1]) - static_cast<double>(xraw[i + 1]);
            a0 = fma(d0, d0, a0);
            a1 = fma(d1, d1, a1);
        }
    }
}
template <typename TX, typename TQ>
__device__ __forceinline__ double canon_d2(const TX *__restrict__ xr, const TQ *__restrict__ qr, int dim, int tid, double *partial4) {
    if (tid < 128) {
        double a0 = 0.0, a1 = 0.0;
        canon_d2_lanes(xr, qr, dim, tid, a0, a1);
        const double w = warp_sum(a0 + a1);
        if ((tid & 31) == 0) partial4[tid >> 5] = w;
    }
    __syncthreads();
    const double tot = ((partial4[0] + partial4[1]) + partial4[2]) + partial4[3];
    __syncthreads();
    return tot;
}

// bit-identical to canon_d2, no block barrier — lets the warps of a block work on different candidates.
// U: 256-element groups per step.  Every step issues all its loads (16 U per lane) before any arithmetic, and a step cannot
// start before the previous one has consumed its values: one memory latency per step.  float32 rows take two groups per
// step (the same bytes in flight as one group of float64); the order of additions into each accumulator is unchanged.
template <typename TX, typename TQ>
__device__ __forceinline__ double canon_d2_warp(const TX *__restrict__ xr, const TQ *__restrict__ qr, int dim, int lane) {
    constexpr int U = (sizeof(TX) + sizeof(TQ) <= 8) ? 2 : 1;
    double a[4][2] = {};
    for (int base = 0; base < dim; base += 256 * U) {
        TQ qv[8 * U];
        TX xv[8 * U];
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const int e0 = base + 256 * u + lane + 32 * g, e1 = e0 + 128;
                qv[8 * u + 2 * g] = (e0 < dim) ? qr[e0] : TQ(0);
                xv[8 * u + 2 * g] = (e0 < dim) ? xr[e0] : TX(0);
                qv[8 * u + 2 * g + 1] = (e1 < dim) ? qr[e1] : TQ(0);
                xv[8 * u + 2 * g + 1] = (e1 < dim) ? xr[e1] : TX(0);
            }
#pragma unroll
        for (int i = 0; i < 8 * U; i++) { keep(qv[i]); keep(xv[i]); }
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const double d0 = static_cast<double>(qv[8 * u + 2 * g]) - static_cast<double>(xv[8 * u + 2 * g]);
                const double d1 = static_cast<double>(qv[8 * u + 2 * g + 1]) - static_cast<double>(xv[8 * u + 2 * g + 1]);
                a[g][0] = fma(d0, d0, a[g][0]);      // out-of-range elements are 0 - 0: they add nothing
                a[g][1] = fma(d1, d1, a[g][1]);
            }
    }
    const double w0 = warp_sum(a[0][0] + a[0][1]), w1 = warp_sum(a[1][0] + a[1][1]);
    const double w2 = warp_sum(a[2][0] + a[2][1]), w3 = warp_sum(a[3][0] + a[3][1]);
    return ((w0 + w1) + w2) + w3;
}

// Row-sharded pools with host-resident queries: every rank uploads only its slice of a chunk's ORIGINAL query rows; the
// exact re-rank reads the row of a query straight from its owner's buffer over NVLink (peer loads) — and only for the
// queries that still have a local survivor after the global pruning, 5x less link traffic than broadcasting the rows.
// All buffers share one layout, so row q sits at the same offset on whichever rank owns it.
struct QueryPull {
    const char *base[16];          // every rank's exchange buffer as mapped here (base[0] == nullptr: rows are local, use qmat)
    size_t off;                    // byte offset of the chunk's row block in a buffer
    int slice_rows;                // rows [r * slice_rows, (r + 1) * slice_rows) of the chunk were uploaded by rank r
};
template <typename TQ>
__device__ __forceinline__ const TQ *query_row(const QueryPull &qp, const TQ *qmat, int64_t ld_q, int q) {
    if (qp.base[0] == nullptr) return qmat + static_cast<int64_t>(q) * ld_q;
    return reinterpret_cast<const TQ *>(qp.base[q / qp.slice_rows] + qp.off) + static_cast<int64_t>(q) * ld_q;
}

struct RerankParams {
    const float *cand_s;
    const int *cand_i;
    int max_slots;                 // row stride of cand_* in shortlists
    const int *slots_per_qtile;    // [query tiles] shortlists actually written for the rows of that tile
    int qtile_rows;                // query rows per tile (BM * CG)
    int dim;
    int64_t ld_x, ld_q;
    int n;                         // pool rows in the shard
    int kk;                        // neighbours to emit (<= C)
    int64_t index_base;
    unsigned flags;                // B200KNN_FLAG_*
    const float *qnorm_bf;         // [nq] ||q~||^2 (fp32, of rounded values)
    const float *q_err;            // [nq] ||q - q~|| rounded up
    const unsigned int *max_xnorm_bf_bits;   // device scalars (pool): max ||x~||^2, max ||x - x~||
    const unsigned int *max_x_err_bits;
    int kp;                        // padded K of the BF16 operands (accumulation length)
    int32_t *out_idx;              // [nq][kk]
    double *out_dist;              // [nq][kk]
    int *uncert_count;             // number of uncertified queries
    int *uncert_list;              // their row numbers
    float *uncert_thr;             // score threshold for the collection pass, per list slot
    int uncert_q_base;             // added to the row numbers queued in uncert_list (a call that runs one second pass over all its chunks)
    // Row-sharded pools (one rank per GPU): ext_bounds[r * ext_stride + q] is rank r's upper bound on the distance of
    // its kk-th nearest row to query q (+inf if it has fewer than kk rows), written into this rank's buffer by the
    // peers' bound_publish_kernel.  The global kk-th distance is at most their minimum, so a local candidate whose
    // lower bound exceeds it cannot be in the global answer: only the global survivors are evaluated exactly, and the
    // local list may hold fewer than kk entries (padded with (-1, DBL_MAX), which the merge skips).
    const float *ext_bounds;       // nullptr: single shard
    int ext_world;
    int64_t ext_stride;
    QueryPull qpull;               // where the original query rows live (zero-initialised: in qmat)
    const float *min_score;        // with ext_bounds: the query's smallest shortlist score (from bound_publish_kernel)
    // precision tier of the tensor pass that produced the scores (tiers.cuh) and its accumulation geometry
    int tier;                      // 0 bf16, 1 bf16x3 (hi/lo split, three segments), 2 tf32
    int k_unit;                    // products accumulated by the tensor core into one accumulator (<= k_total)
    int n_units;                   // accumulation units per score (partial sums added in fp32 by the epilogue when > 1)
    const float *q_lonorm;         // tier 1: ||q_lo|| per query
    const unsigned int *max_x_lonorm_bits;   // tier 1: max ||x_lo|| over the pool
};

// wait (bounded) until every rank's flag shows `step`; executed by the first `world` threads of a block
__device__ __forceinline__ void wait_peer_flags(const unsigned int *flags, int world, unsigned int step, int tag = 0) {
    if (static_cast<int>(threadIdx.x) < world) {
        const unsigned int *f = flags + threadIdx.x * 32;
        const uint64_t t0 = global_timer_ns();
        while (static_cast<int>(ld_acquire_sys(f) - step) < 0) {      // counters only grow; a peer that is ahead is fine
            __nanosleep(200);
            if (global_timer_ns() - t0 > 10000000000ull) {            // 10 s: a peer died
                printf("[b200knn] flag wait timed out: kind %d, rank %d shows %u, want %u (block %d)\n", tag, static_cast<int>(threadIdx.x),
                       ld_acquire_sys(f), step, blockIdx.x);
                assert(0 && "b200knn: peer flag wait timed out");
                __trap();
            }
        }
        __threadfence_system();
    }
}
__device__ __forceinline__ double ext_bound_of(const RerankParams &p, int q) {
    double u = DBL_MAX;
    if (p.ext_bounds)
        for (int r = 0; r < p.ext_world; r++) u = fmin(u, static_cast<double>(__ldcg(p.ext_bounds + r * p.ext_stride + q)));
    return u;
}

// Error model shared by the pruning rule, the certificate and the second-pass threshold.  With q~, x~ the BF16
// roundings:  s~ + ||q~||^2 = ||q~ - x~||^2 up to fp32 accumulation error eps_acc, and
// | ||q - x|| - ||q~ - x~|| | <= ||q - q~|| + ||x - x~|| =: eta   (triangle inequality; both norms are computed
// exactly by convert_norm_kernel, the pool side as a maximum over rows).
struct ErrModel {
    double qn_bf, eps_acc, eta;
    __device__ __forceinline__ double lower(double s) const {   // lower bound on the true distance, given score s
        const double v = s + qn_bf - eps_acc;
        return (v > 0.0 ? sqrt(v) : 0.0) - eta;
    }
    __device__ __forceinline__ double upper(double s) const {   // upper bound on the true distance
        const double v = s + qn_bf + eps_acc;
        return (v > 0.0 ? sqrt(v) : 0.0) + eta;
    }
};
__device__ __forceinline__ ErrModel make_err_model(const RerankParams &p, int q) {
    ErrModel m;
    m.qn_bf = static_cast<double>(p.qnorm_bf[q]);
    const double xn_bf = static_cast<double>(__uint_as_float(*p.max_xnorm_bf_bits));
    // magnitude of the operand rows as the tensor core sees them: plain tiers ||q^|| ||x^||; split tier the three
    // segments [hi | hi | lo] . [hi | lo | hi], with ||hi|| <= ||hi + lo|| + ||lo||
    double QN = m.qn_bf, XN = xn_bf, extra = 0.0;
    if (p.tier == 1) {
        const double ql = static_cast<double>(p.q_lonorm[q]), xl = static_cast<double>(__uint_as_float(*p.max_x_lonorm_bits));
        const double qh = sqrt(m.qn_bf) + ql, xh = sqrt(xn_bf) + xl;
        QN = 2.0 * qh * qh + ql * ql;
        XN = 2.0 * xh * xh + xl * xl;
        extra = 2.0 * ql * xl * (1.0 + 1e-6);          // the q_lo . x_lo products the three MMAs leave out
    }
    const double mag = sqrt(QN * XN);
    // fp32 accumulation inside the tensor core: k_unit products of magnitude <= mag in total per unit (x2 for the -2 factor,
    // truncating adds assumed: EMPIRICALLY VALIDATED model, tests/test_gpu_parity.py::test_tensor_scores_within_the_certified_error_model);
    // summed over the units by Cauchy-Schwarz the total still scales with ONE unit's length.  Then: the fp32 adds of the
    // partial sums (IEEE round-to-nearest), the norm sums, and forming s~ in fp32.
    m.eps_acc = (static_cast<double>(p.k_unit) + 8.0) * 2.4e-7 * mag * 1.001 + static_cast<double>(p.n_units) * 1.2e-7 * mag + extra;
    if (p.tier == 0) m.eps_acc += (static_cast<double>(p.kp) / 16.0 + 8.0) * 1.2e-7 * (xn_bf + m.qn_bf);
    else m.eps_acc += 4.8e-7 * (xn_bf + m.qn_bf + 2.0 * mag);     // norms summed in float64, rounded once
    m.eta = (static_cast<double>(p.q_err[q]) + static_cast<double>(__uint_as_float(*p.max_x_err_bits))) * (1.0 + 1e-6) + 1e-30;
    return m;
}

// Final step of a re-rank, executed by ONE full warp: rank the C exact squared distances by (d2, row), emit the best kk,
// and certify the answer (or queue the query for the second pass).  keysC: the C best-scored shortlist entries,
// ascending by (score, row); d2s: their exact squared distances (DBL_MAX = pruned / empty).
template <int C>
__device__ __forceinline__ void rerank_finish(const RerankParams &p, int q, const unsigned long long *keys, const double *d2s, int lane, double u_ext) {
    // rank the exact distances by (d2, index); each lane owns candidates lane, lane + 32 (C <= 64)
    constexpr int H = (C + 31) / 32;
    double myd[H];
    uint32_t myi[H];
    int rank[H];
#pragma unroll
    for (int h = 0; h < H; h++) {
        const int c = lane + 32 * h;
        myd[h] = (c < C) ? d2s[c] : DBL_MAX;
        myi[h] = (c < C) ? static_cast<uint32_t>(keys[c]) : 0xffffffffu;
        rank[h] = 0;
    }
#pragma unroll
    for (int g = 0; g < H; g++) {
#pragma unroll
        for (int o = 0; o < 32; o++) {
            const double od = __shfl_sync(0xffffffffu, myd[g], o);
            const uint32_t oi = __shfl_sync(0xffffffffu, myi[g], o);
#pragma unroll
            for (int h = 0; h < H; h++) rank[h] += (od < myd[h] || (od == myd[h] && oi < myi[h])) ? 1 : 0;
        }
    }
    double dk2 = DBL_MAX;
    unsigned mk = 0;
#pragma unroll
    for (int h = 0; h < H; h++) {
        const bool valid = (lane + 32 * h) < C;
        if (valid && rank[h] < p.kk) {
            const bool real = myd[h] < DBL_MAX;      // pruned / empty entries rank last: (-1, DBL_MAX) = no entry
            p.out_idx[static_cast<int64_t>(q) * p.kk + rank[h]] = real ? static_cast<int32_t>(p.index_base + myi[h]) : -1;
            p.out_dist[static_cast<int64_t>(q) * p.kk + rank[h]] = !real ? DBL_MAX : ((p.flags & 1u) ? myd[h] : sqrt(myd[h]));
        }
        // k-th exact distance (rank kk-1), broadcast
        const unsigned mh = __ballot_sync(0xffffffffu, valid && rank[h] == p.kk - 1);
        const double dh = __shfl_sync(0xffffffffu, myd[h], mh ? (__ffs(mh) - 1) : 0);
        if (mh) { dk2 = dh; mk = mh; }
    }
    if (lane == 0 && !(p.flags & 2u)) {
        // ---- certificate: every pool row NOT among the C kept has score >= tau (the C-th kept score), hence
        // true distance >= lower(tau).  The answer is exact when the kk-th exact distance is below that.
        bool certified = true;
        const unsigned long long kc = keys[C - 1];
        const ErrModel em = make_err_model(p, q);
        // bound on the (global) kk-th distance: the kk-th exact local distance if there is one, the peers' bound otherwise
        const bool have_dk = (mk != 0) && (dk2 < DBL_MAX);
        const double dk = fmin(have_dk ? sqrt(dk2) : DBL_MAX, u_ext);
        if (p.n > C && kc != ~0ull) {
            const double lb = em.lower(static_cast<double>(float_from_order_bits(static_cast<uint32_t>(kc >> 32))));
            certified = (dk < DBL_MAX) && (lb > 0.0) && (dk < lb);
        }                                           // else: every row of the shard is in the shortlist, nothing was dropped
        if (!certified) {
            // second pass collects every row with score <= thr: any x with d(q,x) <= dk has
            // ||q~ - x~|| <= dk + eta, i.e. s~ <= (dk + eta)^2 - ||q~||^2 + eps_acc.
            const double t = (dk + em.eta) * (dk + em.eta) - em.qn_bf + em.eps_acc;
            const int slot = atomicAdd(p.uncert_count, 1);
            p.uncert_list[slot] = q + p.uncert_q_base;
            p.uncert_thr[slot] = __double2float_ru(t + 1e-6 * fabs(t));
        }
    }
}

// Progressive pruning (executed by the whole block, uniform call): kk exact distances are known, the kk-th true distance
// is at most their maximum — a far tighter limit than the a-priori upper bound.  Scores are sorted, so the survivors
// stay a prefix; returns its new length.
__device__ __forceinline__ int tighten_survivors(const RerankParams &p, int q, const unsigned long long *keys, const double *d2s, int m,
                                                 int *m_s, int warp, int lane, double u_ext) {
    __syncthreads();                    // d2s[0..kk) were written by single threads, possibly without a barrier since
    if (warp == 0) {
        double mx = 0.0;
        for (int i = lane; i < p.kk; i += 32) mx = fmax(mx, d2s[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        int cnt = 0;
        if (mx < DBL_MAX) {
            const ErrModel em = make_err_model(p, q);
            const double dk = fmin(sqrt(mx), u_ext);
            for (int i = p.kk + lane; i < m; i += 32)
                cnt += (keys[i] != ~0ull && em.lower(static_cast<double>(float_from_order_bits(static_cast<uint32_t>(keys[i] >> 32)))) <= dk) ? 1 : 0;
        } else {
            cnt = (lane == 0) ? m - p.kk : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) *m_s = p.kk + cnt;
    }
    __syncthreads();
    return *m_s;
}

template <typename TX, typename TQ, int C, int NT>
__global__ void __launch_bounds__(NT)
rerank_kernel(const TX *__restrict__ x, const TQ *__restrict__ qmat, const RerankParams p) {
    extern __shared__ unsigned long long keys[];   // next_pow2(max_slots * C) entries (host-sized, <= MAX_KEYS)
    __shared__ double d2s[C];
    __shared__ int m_s;
    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    if (p.ext_bounds && p.min_score) {
        // Row-sharded pools: on most ranks a query has NO candidate that survives the global bound (its neighbour lives on
        // another shard).  Then nothing is evaluated, the list is empty, and the answer is certified as it stands: every
        // row of this shard scores at least the shortlist's minimum, whose lower bound already exceeds the global bound.
        const double u = ext_bound_of(p, q);
        const float ms = __ldg(p.min_score + q);
        if (u < DBL_MAX && p.n > C && ms < FLT_MAX && make_err_model(p, q).lower(static_cast<double>(ms)) > u) {     // uniform across the block
            for (int r = tid; r < p.kk; r += blockDim.x) {
                p.out_idx[static_cast<int64_t>(q) * p.kk + r] = -1;
                p.out_dist[static_cast<int64_t>(q) * p.kk + r] = DBL_MAX;
            }
            return;
        }
    }
    const int total = __ldg(p.slots_per_qtile + q / p.qtile_rows) * C;
    int P = 1;
    while (P < total) P <<= 1;

    for (int i = tid; i < P; i += blockDim.x) {
        unsigned long long key = ~0ull;
        if (i < total) {
            const int64_t o = static_cast<int64_t>(q) * p.max_slots * C + i;
            const int idx = p.cand_i[o];
            if (idx >= 0) key = (static_cast<unsigned long long>(float_order_bits(p.cand_s[o])) << 32) | static_cast<uint32_t>(idx);
        }
        keys[i] = key;
    }
    __syncthreads();
    // bitonic sort, ascending by (score, index)
    for (int k2 = 2; k2 <= P; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = keys[i], b = keys[l];
                    const bool up = (i & k2) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[l] = a; }
                }
            }
            __syncthreads();
        }
    }
    // Pruning: candidate c (ascending score) can be among the true top-kk only if its distance lower bound does not
    // exceed the kk-th smallest distance upper bound.  Scores are sorted, so the survivors are a prefix of length m.
    const double u_ext = ext_bound_of(p, q);       // DBL_MAX on a single shard
    if (tid == 0) {
        const ErrModel em = make_err_model(p, q);
        double u = u_ext;
        if (keys[p.kk - 1] != ~0ull)
            u = fmin(u, em.upper(static_cast<double>(float_from_order_bits(static_cast<uint32_t>(keys[p.kk - 1] >> 32)))));
        int m = 0;
        if (u < DBL_MAX) {
            // (single shard: the first kk entries always pass, their lower bounds are below the kk-th upper bound)
            while (m < C && keys[m] != ~0ull &&
                   em.lower(static_cast<double>(float_from_order_bits(static_cast<uint32_t>(keys[m] >> 32)))) <= u) m++;
        } else {
            m = C;                                  // fewer than kk candidates anywhere: evaluate them all
        }
        if (!p.ext_bounds) m = max(m, min(C, p.kk));
        m_s = m;
    }
    __syncthreads();
    int m = m_s;
    // exact float64 distances of the surviving candidates (the arithmetic of util.c:62-69, tree-summed).  The whole
    // block works on one candidate at a time: every thread owns a strided slice of the dimensions, keeps its slice
    // of the query row in registers across candidates, and issues its loads of the pool row back to back.
    const int warp = tid >> 5, lane = tid & 31;
    const TQ *qr = query_row(p.qpull, qmat, p.ld_q, q);
    if constexpr (NT > 128) {
        // Few queries, many shortlists (the trainer's 24-row calls, config 1): the device is far from full, so the eight
        // 128-thread groups of the block each take a candidate (named barrier per group), every candidate still summed
        // by 128 lanes in the canonical order.
        constexpr int NG = NT / 128;
        __shared__ double gpart[NG][4];
        const int grp = tid >> 7, gt = tid & 127;
        bool tightened = false;
        for (int c0 = 0; c0 < C; c0 += NG) {
            if (!tightened && c0 >= p.kk && c0 < m) {      // uniform across the block
                m = tighten_survivors(p, q, keys, d2s, m, &m_s, warp, lane, u_ext);
                tightened = true;
            }
            const int c = c0 + grp;
            if (c < C) {                                     // uniform across the group
                const unsigned long long key = keys[c];
                if (c >= m || key == ~0ull) {
                    if (gt == 0) d2s[c] = DBL_MAX;
                } else {
                    double a0 = 0.0, a1 = 0.0;
                    canon_d2_lanes<(sizeof(TX) + sizeof(TQ) >= 16) ? 8 : 16>(x + static_cast<int64_t>(static_cast<uint32_t>(key)) * p.ld_x, qr, p.dim, gt, a0, a1);   // 64-register budget
                    const double w = warp_sum(a0 + a1);
                    if (lane == 0) gpart[grp][warp & 3] = w;
                    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
                    if (gt == 0) d2s[c] = ((gpart[grp][0] + gpart[grp][1]) + gpart[grp][2]) + gpart[grp][3];
                }
            }
            __syncthreads();
        }
        if (warp == 0) rerank_finish<C>(p, q, keys, d2s, lane, u_ext);
        return;
    }
    constexpr int RQ = 24;                      // dims per thread held in registers (128 threads x 24 = 3072)
    constexpr int nth = 128;                    // the canonical 128 lanes (see canon_d2)
    const bool act = tid < nth;
    __shared__ double partial[4];
    double qreg[RQ];
    const bool fits = p.dim <= RQ * nth;
    if (fits && act && m > 0) {     // (row-sharded pools: most queries have no survivor on most ranks — their row is never read)
        // raw loads first, conversions after: a float->double conversion placed right behind its load would make the
        // in-order warp wait for that load before issuing the next one (24 serialized DRAM round trips)
        TQ qraw[RQ];
#pragma unroll
        for (int i = 0; i < RQ; i++) {
            const int e = tid + i * nth;
            qraw[i] = (e < p.dim) ? qr[e] : TQ(0);
        }
#pragma unroll
        for (int i = 0; i < RQ; i++) keep(qraw[i]);
#pragma unroll
        for (int i = 0; i < RQ; i++) qreg[i] = static_cast<double>(qraw[i]);
    }
    // (Tried in round 2 and dropped: two candidates per step for float32 pools — 164 registers cost a resident block and the
    // config-4 re-rank went from 3.3 to 4.1 ms.  ncu: the kernel is latency-bound per block (sort, prune, one candidate
    // after the other), not DRAM-bound: 4 % of DRAM throughput at config 4, 35 % at config 3.)
    for (int c = 0; c < C; c++) {
        if (c == p.kk && c < m) m = tighten_survivors(p, q, keys, d2s, m, &m_s, warp, lane, u_ext);   // uniform across the block
        const unsigned long long key = keys[c];
        if (c >= m || key == ~0ull) {           // uniform across the block
            if (tid == 0) d2s[c] = DBL_MAX;
            continue;
        }
        const TX *xr = x + static_cast<int64_t>(static_cast<uint32_t>(key)) * p.ld_x;
        if (!fits) {
            const double tot = canon_d2(xr, qr, p.dim, tid, partial);
            if (tid == 0) d2s[c] = tot;
            continue;
        }
        if (act) {                               // same order as canon_d2, query slice already in registers
            double a0 = 0.0, a1 = 0.0;
            TX xraw[RQ];
#pragma unroll
            for (int i = 0; i < RQ; i++) {
                const int e = tid + i * nth;
                xraw[i] = (e < p.dim) ? xr[e] : TX(0);
            }
#pragma unroll
            for (int i = 0; i < RQ; i++) keep(xraw[i]);
#pragma unroll
            for (int i = 0; i < RQ; i += 2) {
                const double d0 = qreg[i] - static_cast<double>(xraw[i]), d1 = qreg[i + 1] - static_cast<double>(xraw[i + 1]);
                a0 = fma(d0, d0, a0);
                a1 = fma(d1, d1, a1);
            }
            const double w = warp_sum(a0 + a1);
            if (lane == 0) partial[warp] = w;
        }
        __syncthreads();
        if (tid == 0) d2s[c] = ((partial[0] + partial[1]) + partial[2]) + partial[3];
        __syncthreads();
    }
    __syncthreads();
    if (warp == 0) rerank_finish<C>(p, q, keys, d2s, lane, u_ext);
}

// The same re-rank with ONE WARP per query (throughput flavour, many queries).  The block flavour above keeps one query
// per 128 threads and 124-157 registers: 3-4 queries resident per SM, each a chain of dependent latencies (sort, prune,
// one candidate after the other) — ncu: 22 % occupancy, 35 % (config 3) / 4 % (config 4) of DRAM throughput.  Here a
// query costs one warp and ~70 registers, so 24 queries are resident per SM and their latency chains overlap; a
// candidate row is streamed by canon_d2_warp (sixteen 8-byte loads in flight per lane, bit-identical to the block
// flavour's sum).  Same pruning, tightening, certificate and output code: results are bit-identical to rerank_kernel.
__device__ __forceinline__ int tighten_survivors_warp(const RerankParams &p, int q, const unsigned long long *keys, const double *d2s, int m,
                                                      int lane, double u_ext) {
    __syncwarp();
    double mx = 0.0;
    for (int i = lane; i < p.kk; i += 32) mx = fmax(mx, d2s[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    int cnt = 0;
    if (mx < DBL_MAX) {
        const ErrModel em = make_err_model(p, q);
        const double dk = fmin(sqrt(mx), u_ext);
        for (int i = p.kk + lane; i < m; i += 32)
            cnt += (keys[i] != ~0ull && em.lower(static_cast<double>(float_from_order_bits(static_cast<uint32_t>(keys[i] >> 32)))) <= dk) ? 1 : 0;
    } else {
        cnt = (lane == 0) ? m - p.kk : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    return p.kk + cnt;
}

// one query, one warp; keys: P entries of this warp's shared memory, d2s: C entries
template <typename TX, typename TQ, int C>
__device__ __forceinline__ void rerank_one_query_warp(const TX *__restrict__ x, const TQ *__restrict__ qmat, const RerankParams &p, int q, int lane,
                                                      unsigned long long *keys, double *d2s) {
    if (p.ext_bounds && p.min_score) {                 // row-sharded pools: no candidate survives the global bound (see rerank_kernel)
        const double u = ext_bound_of(p, q);
        const float ms = __ldg(p.min_score + q);
        if (u < DBL_MAX && p.n > C && ms < FLT_MAX && make_err_model(p, q).lower(static_cast<double>(ms)) > u) {
            for (int r = lane; r < p.kk; r += 32) {
                p.out_idx[static_cast<int64_t>(q) * p.kk + r] = -1;
                p.out_dist[static_cast<int64_t>(q) * p.kk + r] = DBL_MAX;
            }
            return;
        }
    }
    const int total = __ldg(p.slots_per_qtile + q / p.qtile_rows) * C;
    int Pq = 1;
    while (Pq < total) Pq <<= 1;
    for (int i = lane; i < Pq; i += 32) {
        unsigned long long key = ~0ull;
        if (i < total) {
            const int64_t o = static_cast<int64_t>(q) * p.max_slots * C + i;
            const int idx = p.cand_i[o];
            if (idx >= 0) key = (static_cast<unsigned long long>(float_order_bits(p.cand_s[o])) << 32) | static_cast<uint32_t>(idx);
        }
        keys[i] = key;
    }
    for (int c = lane; c < C; c += 32) d2s[c] = DBL_MAX;
    __syncwarp();
    for (int k2 = 2; k2 <= Pq; k2 <<= 1) {             // bitonic sort, ascending by (score, index)
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < Pq; i += 32) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = keys[i], b = keys[l];
                    const bool up = (i & k2) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[l] = a; }
                }
            }
            __syncwarp();
        }
    }
    const double u_ext = ext_bound_of(p, q);
    int m = 0;
    if (lane == 0) {                                   // the pruning rule of rerank_kernel, word for word
        const ErrModel em = make_err_model(p, q);
        double u = u_ext;
        if (keys[p.kk - 1] != ~0ull)
            u = fmin(u, em.upper(static_cast<double>(float_from_order_bits(static_cast<uint32_t>(keys[p.kk - 1] >> 32)))));
        if (u < DBL_MAX) {
            while (m < C && keys[m] != ~0ull &&
                   em.lower(static_cast<double>(float_from_order_bits(static_cast<uint32_t>(keys[m] >> 32)))) <= u) m++;
        } else {
            m = C;
        }
        if (!p.ext_bounds) m = max(m, min(C, p.kk));
    }
    m = __shfl_sync(0xffffffffu, m, 0);
    const TQ *qr = query_row(p.qpull, qmat, p.ld_q, q);
    for (int c = 0; c < m; c++) {
        if (c == p.kk) m = tighten_survivors_warp(p, q, keys, d2s, m, lane, u_ext);
        if (c >= m) break;
        const unsigned long long key = keys[c];
        if (key == ~0ull) continue;                    // (d2s[c] stays DBL_MAX)
        const double d2 = canon_d2_warp(x + static_cast<int64_t>(static_cast<uint32_t>(key)) * p.ld_x, qr, p.dim, lane);
        if (lane == 0) d2s[c] = d2;
    }
    __syncwarp();
    rerank_finish<C>(p, q, keys, d2s, lane, u_ext);
    __syncwarp();                                      // keys / d2s are reused by this warp's next query
}

// Persistent warps: every warp draws its next query from a device counter (zeroed by the pass's plan_pass_kernel), so a
// query with sixteen surviving candidates does not hold back seven warps that drew easy ones (ncu of the first version,
// 8 queries per block, one block per 8 queries: achieved occupancy 21.7 % of a theoretical 37.5 %).
template <typename TX, typename TQ, int C, int WPB>
__global__ void __launch_bounds__(WPB * 32, 3)
rerank_warp_kernel(const TX *__restrict__ x, const TQ *__restrict__ qmat, const RerankParams p, int nq, int P /* pow2 >= max_slots * C */,
                   unsigned int *__restrict__ next_query) {
    extern __shared__ unsigned long long keys_all[];   // [WPB][P]
    __shared__ double d2s_all[WPB][C];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long *keys = keys_all + static_cast<size_t>(warp) * P;
    double *d2s = d2s_all[warp];
    for (;;) {                                         // (the warps of a block never meet at a barrier)
        unsigned int q = 0;
        if (lane == 0) q = atomicAdd(next_query, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= static_cast<unsigned int>(nq)) break;
        rerank_one_query_warp<TX, TQ, C>(x, qmat, p, static_cast<int>(q), lane, keys, d2s);
    }
}

// Second pass, part 2: exact re-rank of the collected lists.  The number of lists is read from device memory (the
// first pass's uncertified counter): the host enqueues the whole second pass without synchronising, and with a count
// of zero every kernel of it returns at once.  Blocks loop over the lists.
// Lists longer than the capacity (or, on a single shard, shorter than kk) are handed to the exact scan via the
// overflow list, which the host inspects once, at the end of the call.
constexpr int COLLECT_CAP = 1024;
struct CollectRerankParams {
    const int *count_dev;        // number of lists (uncertified queries of this pass)
    const int *uncert_list;      // [count] query rows
    const int *coll_count;       // [count]
    const int *coll_idx;         // [count][COLLECT_CAP]
    int dim;
    int64_t ld_x, ld_q;
    int kk;
    int64_t index_base;
    unsigned flags;
    int allow_short;             // row-sharded pool: a list holds every local row within the GLOBAL bound, possibly fewer than kk
    int q_offset;                // first query row of this pass within the call (overflow entries are call-global rows)
    QueryPull qpull;             // where the original query rows live (zero-initialised: in qmat)
    int32_t *out_idx;
    double *out_dist;
    int *overflow_count;
    int *overflow_list;          // call-global query rows that need the exact scan
};

template <typename TX, typename TQ, int NW>      // NW warps per block
__global__ void __launch_bounds__(NW * 32)
rerank_collect_kernel(const TX *__restrict__ x, const TQ *__restrict__ qmat, const CollectRerankParams p) {
    __shared__ double d2[COLLECT_CAP];
    __shared__ int idx[COLLECT_CAP];
    __shared__ double sd[NW];
    __shared__ int si[NW];
    __shared__ double last_d_s;
    __shared__ int last_i_s;
    const int nlists = *p.count_dev;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int slot = blockIdx.x; slot < nlists; slot += gridDim.x) {
        const int q = p.uncert_list[slot];
        const int cnt = p.coll_count[slot];
        if (cnt > COLLECT_CAP || (cnt < p.kk && !p.allow_short)) {      // uniform across the block
            if (threadIdx.x == 0) {
                const int o = atomicAdd(p.overflow_count, 1);
                p.overflow_list[o] = p.q_offset + q;
            }
            continue;
        }
        const TQ *qr = query_row(p.qpull, qmat, p.ld_q, q);
        for (int c = warp; c < cnt; c += NW) {       // one candidate per warp, canonical summation order
            const int j = p.coll_idx[static_cast<int64_t>(slot) * COLLECT_CAP + c];
            const double a0 = canon_d2_warp(x + static_cast<int64_t>(j) * p.ld_x, qr, p.dim, lane);
            if (lane == 0) { d2[c] = a0; idx[c] = j; }
        }
        if (threadIdx.x == 0) { last_d_s = -1.0; last_i_s = -1; }
        __syncthreads();
        for (int r = 0; r < p.kk; r++) {
            const double ld = last_d_s;
            const int li = last_i_s;
            double bd = DBL_MAX;
            int bi = 0x7fffffff;
            for (int c = threadIdx.x; c < cnt; c += blockDim.x) {
                const double d = d2[c];
                const int j = idx[c];
                const bool after = (d > ld) || (d == ld && j > li);
                if (after && (d < bd || (d == bd && j < bi))) { bd = d; bi = j; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
            __syncthreads();
            if (lane == 0) { sd[warp] = bd; si[warp] = bi; }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int w = 1; w < NW; w++)
                    if (sd[w] < bd || (sd[w] == bd && si[w] < bi)) { bd = sd[w]; bi = si[w]; }
                last_d_s = bd;
                last_i_s = bi;
                const bool real = bi != 0x7fffffff;      // the list ran out (allow_short): (-1, DBL_MAX) = no entry
                p.out_idx[static_cast<int64_t>(q) * p.kk + r] = real ? static_cast<int32_t>(p.index_base + bi) : -1;
                p.out_dist[static_cast<int64_t>(q) * p.kk + r] = !real ? DBL_MAX : ((p.flags & 1u) ? bd : sqrt(bd));
            }
            __syncthreads();
        }
    }
}

// gather converted query rows (16-byte units) of the uncertified queries into a compact matrix for the collection pass
__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint4 *__restrict__ src, const int *__restrict__ list, const int *__restrict__ nsel_dev, int vec_per_row,
                   uint4 *__restrict__ dst) {
    const int nsel = *nsel_dev;        // device-side count: zero -> nothing to do
    const int64_t total = static_cast<int64_t>(nsel) * vec_per_row;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / vec_per_row), c = static_cast<int>(i % vec_per_row);
        dst[static_cast<int64_t>(r) * vec_per_row + c] = src[static_cast<int64_t>(list[r]) * vec_per_row + c];
    }
}

}  // namespace b200
