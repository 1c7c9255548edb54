// Multi-GPU: k-way merge of per-shard results and the NVLink peer-memory exchange.
#pragma once
#include "common.cuh"
#include "rerank.cuh"

namespace b200 {
// ------------------------------------------------------------------------------------------------
// k-way merge of per-shard results:  in [G][nq][kk] ascending  ->  out [nq][kk]; ties -> lower index.
// ------------------------------------------------------------------------------------------------
constexpr int MERGE_MAX_LISTS = 16;
__global__ void __launch_bounds__(128)
merge_topk_kernel(const int32_t *__restrict__ idx, const double *__restrict__ dist, int G, int64_t nq, int kk,
                  int32_t *__restrict__ out_idx, double *__restrict__ out_dist) {
    const int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (q >= nq) return;
    int head[MERGE_MAX_LISTS];
#pragma unroll
    for (int g = 0; g < MERGE_MAX_LISTS; g++) head[g] = 0;
    for (int r = 0; r < kk; r++) {
        double bd = DBL_MAX;
        int32_t bi = 0x7fffffff;
        int bg = -1;
#pragma unroll
        for (int g = 0; g < MERGE_MAX_LISTS; g++) {
            if (g < G && head[g] < kk) {
                const int64_t o = (static_cast<int64_t>(g) * nq + q) * kk + head[g];
                const double d = dist[o];
                const int32_t i = idx[o];
                if (i >= 0 && (d < bd || (d == bd && i < bi))) { bd = d; bi = i; bg = g; }
            }
        }
#pragma unroll
        for (int g = 0; g < MERGE_MAX_LISTS; g++)
            if (g == bg) head[g]++;
        out_idx[q * kk + r] = (bg >= 0) ? bi : -1;
        out_dist[q * kk + r] = (bg >= 0) ? bd : DBL_MAX;
    }
}

// A shard that holds fewer than kk rows answers with ks = its row count per query; pad its lists to kk entries with
// (-1, DBL_MAX), which the merge kernels skip (dci.py:278-279: num_neighbours=-1 means all points, however the pool is sharded).
__global__ void __launch_bounds__(256)
pad_topk_kernel(const int32_t *__restrict__ idx, const double *__restrict__ dist, int64_t nq, int ks, int kk,
                int32_t *__restrict__ out_idx, double *__restrict__ out_dist) {
    const int64_t total = nq * kk;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t q = i / kk;
        const int r = static_cast<int>(i % kk);
        out_idx[i] = (r < ks) ? idx[q * ks + r] : -1;
        out_dist[i] = (r < ks) ? dist[q * ks + r] : DBL_MAX;
    }
}

// ------------------------------------------------------------------------------------------------
// NVLink exchange for row-sharded pools, one process per GPU: all-gather by peer stores + merge, no NCCL.
//   publish_topk_kernel : every rank writes its local [nq][kk] (index, distance) lists straight into slot `rank` of EVERY
//                         peer's gather buffer (P2P stores over NVLink / NVSwitch, 16-byte vectors), fences, and the last
//                         block to finish raises the step flag in every peer's buffer.
//   merge_wait_kernel   : waits until all `world` flags in the LOCAL buffer show this step, then k-way merges the lists.
// Buffers are double-buffered by step parity: a rank can be at most one step ahead of a peer that is still merging.
// ------------------------------------------------------------------------------------------------
constexpr int EXCH_MAX_WORLD = 16;
struct ExchPeers {
    int32_t *idx[EXCH_MAX_WORLD];          // peer p's gather buffer for indices   [2][world][max_items]
    double *dist[EXCH_MAX_WORLD];          //                      for distances  [2][world][max_items]
    unsigned int *flags[EXCH_MAX_WORLD];   // peer p's flags [world] (one 128-byte line each)
};

__global__ void __launch_bounds__(256)
publish_topk_kernel(const int32_t *__restrict__ idx, const double *__restrict__ dist, int64_t items, int64_t max_items, int rank,
                    int world, unsigned int step, ExchPeers peers, unsigned int *__restrict__ done_counter,
                    const int *__restrict__ local_overflow /* nullable: count of queries this rank still owes an exact scan */) {
    const int64_t par = step & 1u;
    const int64_t slot = (par * world + rank) * max_items;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int p = 0; p < world; p++) {
        int32_t *di = peers.idx[p] + slot;
        double *dd = peers.dist[p] + slot;
        for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < items; i += stride) {
            di[i] = idx[i];
            dd[i] = dist[i];
        }
    }
    __threadfence_system();                  // my stores are visible to every GPU before the flag can be
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        last = (atomicAdd(done_counter, 1u) == gridDim.x - 1u);
        if (last) *done_counter = 0u;        // every block of this launch has arrived; the next launch starts clean
    }
    __syncthreads();
    if (last && threadIdx.x < world) {
        // word 1 of the flag line: "this rank has overflowed second-pass lists in this call" (every rank must learn it,
        // so that all of them repeat the exchange after the exact scan); ordered before the flag by the release store
        peers.flags[threadIdx.x][rank * 32 + 1] = local_overflow ? static_cast<unsigned int>(*local_overflow) : 0u;
        __threadfence_system();
        st_release_sys(peers.flags[threadIdx.x] + rank * 32, step);
    }
}

__global__ void __launch_bounds__(128)
merge_wait_kernel(const int32_t *__restrict__ gidx, const double *__restrict__ gdist, const unsigned int *__restrict__ flags,
                  int world, unsigned int step, int64_t max_items, int64_t nq, int kk, int32_t *__restrict__ out_idx,
                  double *__restrict__ out_dist, unsigned int *__restrict__ global_overflow /* nullable: sum over ranks */) {
    if (threadIdx.x < world) {
        // system-scope acquire: pairs with the publisher's __threadfence_system() + flag store on another GPU
        const unsigned int *f = flags + threadIdx.x * 32;
        const uint64_t t0 = global_timer_ns();
        while (ld_acquire_sys(f) < step) {   // steps only grow; a peer one step ahead is fine
            __nanosleep(200);
            if (global_timer_ns() - t0 > 10000000000ull) {           // 10 s: a peer died
                printf("[b200knn] merge wait timed out: rank %d shows %u, want %u\n", static_cast<int>(threadIdx.x), ld_acquire_sys(f), step);
                assert(0 && "b200knn: merge flag wait timed out");
                __trap();
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (global_overflow && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned int t = 0;
        for (int r = 0; r < world; r++) t += __ldcg(flags + r * 32 + 1);
        *global_overflow = t;
    }
    const int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (q >= nq) return;
    const int64_t base = static_cast<int64_t>(step & 1u) * world * max_items;
    int head[EXCH_MAX_WORLD];
#pragma unroll
    for (int g = 0; g < EXCH_MAX_WORLD; g++) head[g] = 0;
    for (int r = 0; r < kk; r++) {
        double bd = DBL_MAX;
        int32_t bi = 0x7fffffff;
        int bg = -1;
#pragma unroll
        for (int g = 0; g < EXCH_MAX_WORLD; g++) {
            if (g < world && head[g] < kk) {
                const int64_t o = base + g * max_items + q * kk + head[g];
                const double d = __ldcg(gdist + o);
                const int32_t i = __ldcg(gidx + o);
                if (i >= 0 && (d < bd || (d == bd && i < bi))) { bd = d; bi = i; bg = g; }
            }
        }
#pragma unroll
        for (int g = 0; g < EXCH_MAX_WORLD; g++)
            if (g == bg) head[g]++;
        out_idx[q * kk + r] = (bg >= 0) ? bi : -1;
        out_dist[q * kk + r] = (bg >= 0) ? bd : DBL_MAX;
    }
}

// ------------------------------------------------------------------------------------------------
// Row-sharded query protocol (b200knn_exchange_query*): flag lines, the bound exchange, broadcast plumbing.
// Every rank owns one buffer (CUDA IPC / peer mapped into the others); a flag "line" is 128 bytes, one per
// (kind, source rank); counters only grow, so a reader waits for `>= step`.
// ------------------------------------------------------------------------------------------------
struct PeerPtrs {
    char *base[EXCH_MAX_WORLD];            // every rank's buffer as mapped into THIS process (own entry: the local buffer)
};

// raise flag line (kind offset `flag_off`, slot `rank`) to `step` in every rank's buffer; everything the stream did
// before this kernel (copy-engine broadcasts, kernels) is ordered before the flag
__global__ void __launch_bounds__(32)
raise_flags_kernel(PeerPtrs peers, int world, size_t flag_off, int rank, unsigned int step) {
    if (static_cast<int>(threadIdx.x) < world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned int *>(peers.base[threadIdx.x] + flag_off) + rank * 32, step);
    }
}

// hold the stream until every rank's flag line of this kind (in the LOCAL buffer) shows `step`
__global__ void __launch_bounds__(32)
wait_flags_kernel(const unsigned int *__restrict__ flags, int world, unsigned int step, int tag) {
    wait_peer_flags(flags, world, step, tag);
}

// Broadcast of up to four byte ranges of THIS rank's buffer to the same offsets of every peer's buffer, by peer stores
// from the SMs (NVLink).  The copy engines cannot do it without stalling: peer copies and host-to-device uploads share
// an engine on this part, and a broadcast queued behind the next chunk's 30 MB upload arrives a millisecond late
// (measured).  A few dozen blocks are enough to fill the links; they share the SMs with the persistent distance kernel
// of the previous chunk for a fraction of a millisecond.  Every thread fences its stores at system scope before it
// exits, so the flag kernel that follows in the stream publishes complete data.
struct BcastSeg { size_t off; size_t bytes; };     // byte offset in the buffer, length (multiples of 4)
struct BcastParams {
    PeerPtrs peers;
    int world, rank, nseg;
    BcastSeg seg[4];
};
__global__ void __launch_bounds__(512)
broadcast_segments_kernel(const BcastParams p) {
    const size_t tid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
    const char *src = p.peers.base[p.rank];
    for (int sgi = 0; sgi < p.nseg; sgi++) {
        const size_t off = p.seg[sgi].off, bytes = p.seg[sgi].bytes;
        if (((off | bytes) & 15u) == 0) {
            const size_t n = bytes >> 4;
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src + off);
            for (size_t i = tid; i < n; i += 4 * nthreads) {      // four 16-byte loads in flight per thread, each stored world-1 times
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (i + u * nthreads < n) v[u] = __ldcs(s4 + i + u * nthreads);
                for (int r = 0; r < p.world; r++) {
                    if (r == p.rank) continue;
                    uint4 *d4 = reinterpret_cast<uint4 *>(p.peers.base[r] + off);
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (i + u * nthreads < n) d4[i + u * nthreads] = v[u];
                }
            }
        } else {
            const size_t n = bytes >> 2;
            const uint32_t *s1 = reinterpret_cast<const uint32_t *>(src + off);
            for (size_t i = tid; i < n; i += nthreads) {
                const uint32_t v = s1[i];
                for (int r = 0; r < p.world; r++)
                    if (r != p.rank) reinterpret_cast<uint32_t *>(p.peers.base[r] + off)[i] = v;
            }
        }
    }
    __threadfence_system();
}

// Rare path (k > 32 / forced scan with host-resident queries): the exact scan wants the whole chunk's original rows in
// local memory — pull the other ranks' slices over NVLink (peer loads, 16-byte units when the rows allow it).
__global__ void __launch_bounds__(256)
pull_rows_kernel(QueryPull qp, int rank, int64_t rows, size_t row_bytes, char *__restrict__ dst) {
    const size_t units = row_bytes >> 2;
    const size_t total = static_cast<size_t>(rows) * units;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int q = static_cast<int>(i / units);
        const int owner = q / qp.slice_rows;
        if (owner == rank) continue;
        reinterpret_cast<uint32_t *>(dst)[i] = reinterpret_cast<const uint32_t *>(qp.base[owner] + qp.off)[i];
    }
}

// Upper bound on the distance of this shard's kk-th nearest row, per query, from the tensor-pass shortlists
// (kk-th smallest score over the query's slots -> ErrModel::upper), stored straight into slot `rank` of EVERY rank's
// bound buffer (peer stores over NVLink); the last block raises the step flag everywhere.  One warp per query.
struct BoundParams {
    const float *cand_s;           // [nq][max_slots][C]
    const int *cand_i;
    int max_slots, c;
    const int *slots_per_qtile;
    int qtile_rows;
    int nq, kk;
    RerankParams em;               // error-model inputs (qnorm_bf, q_err, pool maxima, kp)
    PeerPtrs peers;
    int world, rank;
    size_t bounds_off;             // byte offset of the bound array [2][world][stride] in every buffer
    size_t flag_off;
    int64_t stride;
    unsigned int step;
    unsigned int *done_counter;
    float *min_score;              // [nq] local: the smallest shortlist score of the query (the re-rank's early exit reads it)
};
__global__ void __launch_bounds__(256)
bound_publish_kernel(const BoundParams p) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int64_t par = p.step & 1u;
    for (int q = blockIdx.x * wpb + (threadIdx.x >> 5); q < p.nq; q += gridDim.x * wpb) {
        const int total = __ldg(p.slots_per_qtile + q / p.qtile_rows) * p.c;
        const float *cs = p.cand_s + static_cast<int64_t>(q) * p.max_slots * p.c;
        const int *ci = p.cand_i + static_cast<int64_t>(q) * p.max_slots * p.c;
        // kk-th smallest valid score: kk rounds of "smallest score above the previous pick" (scores of distinct rows may
        // tie: picks are ordered by (score, position))
        float last_s = -FLT_MAX, first_s = FLT_MAX;
        int last_pos = -1;
        bool ok = true;
        for (int r = 0; r < p.kk && ok; r++) {
            float bs = FLT_MAX;
            int bp = 0x7fffffff;
            for (int i = lane; i < total; i += 32) {
                if (ci[i] < 0) continue;
                const float sc = cs[i];
                const bool after = (sc > last_s) || (sc == last_s && i > last_pos);
                if (after && (sc < bs || (sc == bs && i < bp))) { bs = sc; bp = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float os = __shfl_xor_sync(0xffffffffu, bs, o);
                const int op = __shfl_xor_sync(0xffffffffu, bp, o);
                if (os < bs || (os == bs && op < bp)) { bs = os; bp = op; }
            }
            ok = bp != 0x7fffffff;
            last_s = bs;
            last_pos = bp;
            if (r == 0) first_s = bs;          // FLT_MAX when the query has no candidate at all
        }
        if (lane == 0) p.min_score[q] = first_s;
        float u = __int_as_float(0x7f800000);      // +inf: fewer than kk rows in this shard's shortlists
        if (ok) {
            const ErrModel em = make_err_model(p.em, q);
            u = __double2float_ru(em.upper(static_cast<double>(last_s)) * (1.0 + 1e-7));
        }
        if (lane < p.world)
            reinterpret_cast<float *>(p.peers.base[lane] + p.bounds_off)[(par * p.world + p.rank) * p.stride + q] = u;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        last = (atomicAdd(p.done_counter, 1u) == gridDim.x - 1u);
        if (last) *p.done_counter = 0u;
    }
    __syncthreads();
    if (last && static_cast<int>(threadIdx.x) < p.world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned int *>(p.peers.base[threadIdx.x] + p.flag_off) + p.rank * 32, p.step);
    }
}

// Column sums of every shard -> global column means (the centring vector must be the SAME on every rank: a query
// row converted by one rank is compared with pool rows converted by another).  sums: [world][dim + 1] doubles in the
// local buffer (slot r written by rank r; last entry = its row count); ranks are summed in rank order.
__global__ void __launch_bounds__(256)
publish_colsum_kernel(const double *__restrict__ colsum, double rows, int dim, PeerPtrs peers, int world, int rank, size_t sums_off) {
    for (int p = 0; p < world; p++) {
        double *dst = reinterpret_cast<double *>(peers.base[p] + sums_off) + static_cast<int64_t>(rank) * (dim + 1);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= dim; i += gridDim.x * blockDim.x) dst[i] = (i < dim) ? colsum[i] : rows;
    }
}
__global__ void __launch_bounds__(256)
global_mean_kernel(const double *__restrict__ sums, int world, int dim, double *__restrict__ mean) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim) return;
    double t = 0.0, n = 0.0;
    for (int r = 0; r < world; r++) {
        t += __ldcg(sums + static_cast<int64_t>(r) * (dim + 1) + i);
        n += __ldcg(sums + static_cast<int64_t>(r) * (dim + 1) + dim);
    }
    mean[i] = n > 0.0 ? t / n : 0.0;
}

}  // namespace b200
