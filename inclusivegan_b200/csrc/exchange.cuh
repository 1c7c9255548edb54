// Multi-GPU: k-way merge of per-shard results and the NVLink peer-memory exchange.
#pragma once
#include "common.cuh"

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
                    int world, unsigned int step, ExchPeers peers, unsigned int *__restrict__ done_counter) {
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
        __threadfence_system();
        *reinterpret_cast<volatile unsigned int *>(peers.flags[threadIdx.x] + rank * 32) = step;
    }
}

__global__ void __launch_bounds__(128)
merge_wait_kernel(const int32_t *__restrict__ gidx, const double *__restrict__ gdist, const unsigned int *__restrict__ flags,
                  int world, unsigned int step, int64_t max_items, int64_t nq, int kk, int32_t *__restrict__ out_idx,
                  double *__restrict__ out_dist) {
    if (threadIdx.x < world) {
        // system-scope acquire: pairs with the publisher's __threadfence_system() + flag store on another GPU
        const unsigned int *f = flags + threadIdx.x * 32;
        const uint64_t t0 = global_timer_ns();
        while (ld_acquire_sys(f) < step) {   // steps only grow; a peer one step ahead is fine
            __nanosleep(200);
            if (global_timer_ns() - t0 > 20000000000ull) __trap();   // 20 s: a peer died
        }
        __threadfence_system();
    }
    __syncthreads();
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

}  // namespace b200
