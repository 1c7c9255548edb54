// Device code of libb200knn: the kernels of the IMLE matching path.
//
//   convert_norm_kernel   HBM-bound.  rows of f64/f32 -> BF16 rows (TMA-friendly pitch) + ||x~||^2 (fp32, of the
//                         rounded values) + ||x - x~||, the exact rounding perturbation (for the exactness certificate).
//   dist_topc_kernel      tensor-bound.  Q x N distance scores  s~ = ||x~||^2 - 2 q~.x~  as a BF16 GEMM on tcgen05
//                         (TMA -> 4-stage smem ring -> tcgen05.mma, fp32 accumulators double-buffered in TMEM) with
//                         a fused per-row top-C selection in the epilogue: the Q x N matrix never reaches HBM.
//   rerank_kernel         merges the per-chunk shortlists, recomputes the C survivors exactly (float64 accumulation
//                         of (q-x)^2 over the ORIGINAL f64/f32 rows, the arithmetic of the reference's
//                         compute_dist, dci_code/src/util.c:62-69), selects k, and CERTIFIES the answer: a query
//                         is exact when its k-th exact distance is below a rigorous lower bound on the true distance
//                         of every point the BF16 pass dropped.
//   scan_*                exact float64 CUDA-core scan: second pass for uncertified queries and path for k > 16.
//   merge_topk_kernel     k-way merge of per-shard results (multi-GPU row sharding).
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

// Opaque use of a register: everything loaded into the arguments must be issued before the compiler may start
// consuming them (it otherwise re-fuses "load all, then convert all" into load/convert pairs that reuse three
// registers, i.e. three loads in flight instead of twenty-four).
__device__ __forceinline__ void keep(float &v) { asm volatile("" : "+f"(v)); }
__device__ __forceinline__ void keep(double &v) { asm volatile("" : "+d"(v)); }

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
    const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
    const int64_t r1 = min(r0 + rows_per_block, n);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
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
// vec != 0 requires: dim % 8 == 0 (so kp == dim), src rows 16-byte aligned.
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
            // two 8-element groups per lane per step: all loads of a step are issued before the first use
            int g = lane;
            for (; g + 32 < groups; g += 64) {
                double v[2][8];
                load8<T>(s + (g << 3), v[0]);
                load8<T>(s + ((g + 32) << 3), v[1]);
                if (mu) {
                    double m0[8], m1[8];
                    load8_cached(mu + (g << 3), m0);
                    load8_cached(mu + ((g + 32) << 3), m1);
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        v[0][i] -= m0[i];
                        v[1][i] -= m1[i];
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    __nv_bfloat162 b[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        b[i] = __floats2bfloat162_rn(static_cast<float>(v[h][2 * i]), static_cast<float>(v[h][2 * i + 1]));
                        const float lo = __low2float(b[i]), hi = __high2float(b[i]);
                        acc = fmaf(lo, lo, acc);
                        acc = fmaf(hi, hi, acc);
                        const double e0 = v[h][2 * i] - static_cast<double>(lo), e1 = v[h][2 * i + 1] - static_cast<double>(hi);
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
        if (mx_bf > 0.f) atomicMax(max_norm_bf_bits, __float_as_uint(mx_bf));
        if (mx_er > 0.f) atomicMax(max_err_bits, __float_as_uint(mx_er));
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel 2: BF16 distance GEMM on tcgen05 with fused top-C epilogue.
// ------------------------------------------------------------------------------------------------
// Work schedule (built on the host, Shard::plan): `nrounds` rounds of `workers` items; worker w (a CTA, or a CTA pair
// for cta_group::2) takes items[r * workers + w] in round r.  An item is one query tile swept over a contiguous range
// of pool tiles.  All items of a round have the same length (+-1 tile) and the query tiles of a round form a group
// whose BF16 rows fit in L2 next to the pool tiles being streamed, so the workers that share pool tiles stay in
// lockstep and each pool tile is fetched from HBM once per round; a grid barrier separates rounds.
struct WorkItem { int qtile, t0, t1, slot; };   // qtile < 0: idle in this round; slot: bits 0-15 = shortlist slot of the query
                                                // rows (= pool-tile stream of the round), bits 16-23 = workers sharing the stream,
                                                // bits 24-31 = 0, or the number of active workers of a round that runs in
                                                // round-wide lockstep (long K: see Shard::plan)

struct DistParams {
    const float *xnorm;      // [n] ||x~||^2
    int n;                   // pool rows in this shard
    int nq;                  // query rows
    int num_kb;              // ceil(kp / BK)
    const WorkItem *items;   // [nrounds][workers]
    int nrounds;
    int workers;
    unsigned int *round_counter;   // grid barrier between rounds (zeroed by the host before the launch)
    unsigned int *stream_sync;     // [nrounds][max_slots] lockstep counters of the workers sharing a pool-tile stream (zeroed)
    int sync_tiles;                // the sharers of a stream re-align every sync_tiles tiles (0 = never)
    unsigned int sync_timeout_ns;  // bound on one lockstep wait (~2 tile times)
    int max_slots;           // shortlists per query row in cand_* (row stride)
    float *cand_s;           // [nq][max_slots][C] approximate scores, ascending
    int *cand_i;             // [nq][max_slots][C] shard-local row index (-1 = empty slot)
    // collect mode (second pass): every pool row whose score is <= thr[row] is appended to the row's list
    const float *thr;        // [nq]
    int *coll_count;         // [nq] running count (may exceed coll_cap: overflow)
    int *coll_idx;           // [nq][coll_cap]
    int coll_cap;
    unsigned opt;            // tuning switches (A/B measurements): bit2 grid barrier between rounds
};

__device__ __forceinline__ WorkItem load_item(const DistParams &p, int round, int worker) {
    const int4 v = __ldg(reinterpret_cast<const int4 *>(p.items) + static_cast<int64_t>(round) * p.workers + worker);
    WorkItem w;
    w.qtile = v.x; w.t0 = v.y; w.t1 = v.z; w.slot = v.w;
    return w;
}

// sorted-ascending register list; precondition for insert: s < v[C-1]
template <int C>
__device__ __forceinline__ void topc_insert(float (&v)[C], int (&id)[C], float s, int idx) {
    v[C - 1] = s;
    id[C - 1] = idx;
#pragma unroll
    for (int i = C - 1; i > 0; --i) {
        const bool sw = v[i] < v[i - 1];
        const float a = v[i], b = v[i - 1];
        const int ia = id[i], ib = id[i - 1];
        v[i] = sw ? b : a;
        v[i - 1] = sw ? a : b;
        id[i] = sw ? ib : ia;
        id[i - 1] = sw ? ia : ib;
    }
}

// COLLECT == false: per-(query row, chunk) top-C shortlist.   COLLECT == true: threshold collection (second pass).
// CG == 1: one CTA computes a 128(query) x 256(pool) tile per step.
// CG == 2: a cluster of two CTAs (one SM pair) computes a 256 x 256 tile with tcgen05.mma.cta_group::2: each CTA
//          stages its own 128 query rows (A half) and 128 of the 256 pool rows (B half), so per SM the operand
//          traffic into and out of shared memory drops by a third; CTA rank 0 issues every MMA, both CTAs run
//          the TMA producer and the epilogue for their own 128 query rows (their own TMEM lanes).
template <int CG>
struct DistCfg {
    static constexpr int STAGES = (CG == 1) ? 4 : 6;
    static constexpr int B_ROWS = BN / CG;                       // pool rows staged per CTA
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = B_ROWS * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SCRATCH_BYTES = 32 * 128 * 4;            // epilogue: one 32-score slab per thread (rare path)
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + 2 * BN * 4 + 256 + SCRATCH_BYTES;
};

template <int C, bool COLLECT, int CG>
__global__ void __launch_bounds__(DIST_THREADS, 1)
dist_topc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x, const DistParams p) {
    using Cfg = DistCfg<CG>;
    constexpr int NSTAGE = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles must sit on 1024-byte boundaries (identical carve-up in both CTAs of a pair)
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t smem_a = base;
    const uint32_t smem_b = base + NSTAGE * Cfg::A_BYTES;
    float *xn_s = reinterpret_cast<float *>(gen + NSTAGE * Cfg::STAGE_BYTES);   // [2][BN]
    const uint32_t bars = base + NSTAGE * Cfg::STAGE_BYTES + 2 * BN * 4;
    const uint32_t bar_full = bars;                       // [NSTAGE]  TMA -> MMA      (the leader's copy is used)
    const uint32_t bar_empty = bars + 8 * NSTAGE;         // [NSTAGE]  MMA -> TMA      (one per CTA)
    const uint32_t bar_tfull = bars + 16 * NSTAGE;        // [2]       MMA -> epilogue (one per CTA)
    const uint32_t bar_tempty = bars + 16 * NSTAGE + 16;  // [2]       epilogue -> MMA (the leader's copy is used)
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + NSTAGE * Cfg::STAGE_BYTES + 2 * BN * 4 + 16 * NSTAGE + 32);
    float *scratch = reinterpret_cast<float *>(gen + NSTAGE * Cfg::STAGE_BYTES + 2 * BN * 4 + 256);   // [32][128]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int worker = (CG == 2) ? (blockIdx.x >> 1) : blockIdx.x;          // cluster (or CTA) index

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_x);
        for (int s = 0; s < NSTAGE; s++) {
            mbar_init(bar_full + 8 * s, 1);       // the leader's arrive.expect_tx covers the bytes of BOTH CTAs' loads
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 4 * CG);   // one arrival per epilogue warp of every CTA
        }
        fence_mbar_init();
    }
    if (CG == 2) cluster_sync_all();   // barriers of both CTAs initialised before anyone allocates / arrives remotely
    if (warp == 1) {
        tmem_alloc<CG>(smem_u32(const_cast<uint32_t *>(tmem_slot)), TMEM_COLS);
        tmem_relinquish<CG>();
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (one lane per CTA) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t full_remote = 0;   // cluster address of the leader's full barriers
            if (CG == 2) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(full_remote) : "r"(bar_full), "r"(0));
            for (int round = 0; round < p.nrounds; round++) {
                const WorkItem w = load_item(p, round, worker);
                if (w.qtile >= 0) {
                    const int q0 = w.qtile * (BM * CG) + static_cast<int>(cta_rank) * BM;
                    const unsigned int sw = static_cast<unsigned int>(w.slot);
                    const int wide = static_cast<int>(sw >> 24), per_stream = static_cast<int>((sw >> 16) & 0xffu);
                    const int sharers = wide ? wide : per_stream;
                    // round-wide lockstep covers the tiles every stream of the round has (chunks differ by one tile)
                    const int sync_lim = wide ? ((p.n + BN - 1) / BN) / (wide / per_stream) : 0x7fffffff;
                    unsigned int *sync = p.stream_sync + static_cast<int64_t>(round) * p.max_slots + (wide ? 0u : (sw & 0xffffu));
                    for (int t = w.t0; t < w.t1; t++) {
                        // lockstep: the workers streaming the same pool tiles re-align every sync_tiles tiles, so a tile
                        // fetched from HBM by the first of them is still in L2 when the last one asks for it
                        if (p.sync_tiles > 0 && sharers > 1 && (CG == 1 || leader) && t > w.t0 && (t - w.t0) < sync_lim && (t - w.t0) % p.sync_tiles == 0) {
                            const unsigned int target = static_cast<unsigned int>((t - w.t0) / p.sync_tiles) * sharers;
                            atomicAdd(sync, 1u);
                            // bounded: if a sharer is not resident (a foreign kernel holds its SM) we go on alone after
                            // ~2 tile times — only L2 sharing is lost, never progress
                            const uint64_t t_start = global_timer_ns();
                            while (*reinterpret_cast<volatile unsigned int *>(sync) < target) {
                                __nanosleep(128);
                                if (global_timer_ns() - t_start > p.sync_timeout_ns) break;
                            }
                        }
                        const int n0 = t * BN + static_cast<int>(cta_rank) * Cfg::B_ROWS;
                        for (int kb = 0; kb < p.num_kb; kb++) {
                            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                            if (CG == 1) {
                                mbar_expect_tx(bar_full + 8 * stage, Cfg::STAGE_BYTES);
                                tma_load_2d(smem_a + stage * Cfg::A_BYTES, &tmap_q, bar_full + 8 * stage, kb * BK, q0);
                                tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tmap_x, bar_full + 8 * stage, kb * BK, n0);
                            } else {
                                // all four loads of the pair (2 x A half, 2 x B half) signal the LEADER's barrier; the
                                // peer never arrives there: its loads only complete_tx (a remote arrive per K block
                                // would cost a cluster-scope fence each time)
                                if (leader) mbar_expect_tx(bar_full + 8 * stage, 2 * Cfg::STAGE_BYTES);
                                const uint32_t fb = leader ? (bar_full + 8 * stage) : (full_remote + 8 * stage);
                                tma_load_2d_cg2(smem_a + stage * Cfg::A_BYTES, &tmap_q, fb, kb * BK, q0);
                                tma_load_2d_cg2(smem_b + stage * Cfg::B_BYTES, &tmap_x, fb, kb * BK, n0);
                            }
                            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                        }
                    }
                }
                // grid barrier: nobody starts streaming the next round's pool tiles before everyone is done issuing
                // this round's loads (keeps the workers that share pool tiles in lockstep)
                if ((p.opt & 4u) && round + 1 < p.nrounds) {
                    __threadfence();
                    atomicAdd(p.round_counter, 1u);
                    const unsigned int target = static_cast<unsigned int>(round + 1) * gridDim.x;
                    const uint64_t t_start = global_timer_ns();
                    while (*reinterpret_cast<volatile unsigned int *>(p.round_counter) < target) {
                        __nanosleep(256);
                        if (global_timer_ns() - t_start > 2000000ull) break;   // bounded (2 ms): never a deadlock
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        // The whole warp runs the loop converged (uniform control flow, uniform registers); one elected lane issues.
        // The issuing thread is on the critical path: per K block it must spend less than the 512 tensor cycles the
        // four MMAs take.
        if (leader) {
            constexpr uint32_t idesc = make_idesc_bf16(BM * CG, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int round = 0; round < p.nrounds; round++) {
                const WorkItem w = load_item(p, round, worker);
                if (w.qtile < 0) continue;
                for (int t = w.t0; t < w.t1; t++) {
                    mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);     // epilogues have drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * BN;
                    for (int kb = 0; kb < p.num_kb; kb++) {
                        mbar_wait(bar_full + 8 * stage, phase);               // TMA bytes (of both CTAs) have landed
                        tc_fence_after();
                        const uint32_t cur = stage;
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                        if (elect_one()) {
                            const uint64_t da = make_smem_desc_sw128(smem_a + cur * Cfg::A_BYTES);
                            const uint64_t db = make_smem_desc_sw128(smem_b + cur * Cfg::B_BYTES);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; k++) {
                                // +32 bytes per K slice inside the 128-byte swizzle row: +2 in the (>>4) address field
                                umma_bf16<CG>(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                            }
                            // frees the smem slot (in both CTAs) when the MMAs retire
                            if (CG == 1) umma_commit(bar_empty + 8 * cur);
                            else umma_commit_cg2(bar_empty + 8 * cur, 0x3);
                            if (kb == p.num_kb - 1) {                      // accumulator complete -> epilogue(s)
                                if (CG == 1) umma_commit(bar_tfull + 8 * acc);
                                else umma_commit_cg2(bar_tfull + 8 * acc, 0x3);
                            }
                        }
                        __syncwarp();
                    }
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue: 4 warps, thread <-> query row =====================
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may read
        const int row_in_tile = quarter * 32 + lane;
        const int et = threadIdx.x - 64;               // 0..127
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int round = 0; round < p.nrounds; round++) {
            const WorkItem w = load_item(p, round, worker);
            if (w.qtile < 0) continue;
            const int t0 = w.t0, t1 = w.t1;
            float v[C];
            int id[C];
            const int q = w.qtile * (BM * CG) + static_cast<int>(cta_rank) * BM + row_in_tile;
            float thr = -FLT_MAX;
            int local_hits = 0;
            if constexpr (COLLECT) {
                if (q < p.nq) thr = __ldg(p.thr + q);
            } else {
#pragma unroll
                for (int i = 0; i < C; i++) { v[i] = FLT_MAX; id[i] = -1; }
            }
            for (int t = t0; t < t1; t++) {
                const int n0 = t * BN;
                // stage ||x~||^2 of this tile; rows past the end of the pool can never be selected
                float *xs = xn_s + acc * BN;
                {
                    const int c0 = n0 + et, c1 = n0 + et + 128;
                    xs[et] = (c0 < p.n) ? __ldg(p.xnorm + c0) : FLT_MAX;
                    xs[et + 128] = (c1 < p.n) ? __ldg(p.xnorm + c1) : FLT_MAX;
                }
                named_bar_sync(1, 128);
                mbar_wait(bar_tfull + 8 * acc, acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
                // TMEM -> registers in 32-column slabs.  Hot path per score: FFMA + compare + predicated OR into a hit
                // mask (no branches, compact code: the issuing warps share the SM's instruction cache with this loop).
                // Slabs with hits park their 32 scores in shared memory and replay only the hit positions through ONE
                // copy of the sorted-insert code.
#pragma unroll 1
                for (int c = 0; c < BN / 32; c++) {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + c * 32, r);
                    tmem_ld_wait();
                    if constexpr (COLLECT) {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const float sc = fmaf(-2.f, __uint_as_float(r[j]), xs[c * 32 + j]);
                            if (sc <= thr && local_hits <= p.coll_cap) {   // a row that filled its list from here stops counting
                                local_hits++;
                                const int pos = atomicAdd(p.coll_count + q, 1);
                                if (pos < p.coll_cap) p.coll_idx[static_cast<int64_t>(q) * p.coll_cap + pos] = n0 + c * 32 + j;
                            }
                        }
                    } else {
                        const float worst = v[C - 1];
                        uint32_t hits = 0;
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const float sc = fmaf(-2.f, __uint_as_float(r[j]), xs[c * 32 + j]);
                            r[j] = __float_as_uint(sc);
                            hits |= (sc < worst) ? (1u << j) : 0u;
                        }
                        if (hits) {
#pragma unroll
                            for (int j = 0; j < 32; j++) scratch[j * 128 + et] = __uint_as_float(r[j]);
                            do {
                                const int j = __ffs(hits) - 1;
                                hits &= hits - 1;
                                const float sc = scratch[j * 128 + et];
                                if (sc < v[C - 1]) topc_insert<C>(v, id, sc, n0 + c * 32 + j);
                            } while (hits);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {                       // this warp is done with the accumulator
                    if (CG == 1 || leader) mbar_arrive(bar_tempty + 8 * acc);
                    else mbar_arrive_cluster(bar_tempty + 8 * acc, 0);
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
            if constexpr (!COLLECT) {
                if (q < p.nq) {
                    float *cs = p.cand_s + (static_cast<int64_t>(q) * p.max_slots + (w.slot & 0xffff)) * C;
                    int *ci = p.cand_i + (static_cast<int64_t>(q) * p.max_slots + (w.slot & 0xffff)) * C;
#pragma unroll
                    for (int i = 0; i < C; i += 4) {
                        *reinterpret_cast<float4 *>(cs + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        *reinterpret_cast<int4 *>(ci + i) = make_int4(id[i], id[i + 1], id[i + 2], id[i + 3]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();   // the peer's smem / TMEM stay alive until the leader's last MMA has retired
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<CG>(tmem_base, TMEM_COLS);
    }
}

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
template <typename TX, typename TQ>
__device__ __forceinline__ double canon_d2(const TX *__restrict__ xr, const TQ *__restrict__ qr, int dim, int tid, double *partial4) {
    if (tid < 128) {
        double a0 = 0.0, a1 = 0.0;
        constexpr int RB = 16;                   // elements per lane and step: 32 loads in flight before any arithmetic
        for (int base = 0; base < dim; base += RB * 128) {
            TQ qraw[RB];
            TX xraw[RB];
#pragma unroll
            for (int i = 0; i < RB; i++) {
                const int e = base + tid + i * 128;
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
        const double w = warp_sum(a0 + a1);
        if ((tid & 31) == 0) partial4[tid >> 5] = w;
    }
    __syncthreads();
    const double tot = ((partial4[0] + partial4[1]) + partial4[2]) + partial4[3];
    __syncthreads();
    return tot;
}

// bit-identical to canon_d2, no block barrier — lets the warps of a block work on different candidates.
template <typename TX, typename TQ>
__device__ __forceinline__ double canon_d2_warp(const TX *__restrict__ xr, const TQ *__restrict__ qr, int dim, int lane) {
    double a[4][2] = {};
    for (int base = 0; base < dim; base += 256) {
        // all sixteen loads of the step first (see keep()), then the arithmetic
        TQ qv[8];
        TX xv[8];
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const int e0 = base + lane + 32 * g, e1 = e0 + 128;
            qv[2 * g] = (e0 < dim) ? qr[e0] : TQ(0);
            xv[2 * g] = (e0 < dim) ? xr[e0] : TX(0);
            qv[2 * g + 1] = (e1 < dim) ? qr[e1] : TQ(0);
            xv[2 * g + 1] = (e1 < dim) ? xr[e1] : TX(0);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) { keep(qv[i]); keep(xv[i]); }
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const double d0 = static_cast<double>(qv[2 * g]) - static_cast<double>(xv[2 * g]);
            const double d1 = static_cast<double>(qv[2 * g + 1]) - static_cast<double>(xv[2 * g + 1]);
            a[g][0] = fma(d0, d0, a[g][0]);      // out-of-range elements are 0 - 0: they add nothing
            a[g][1] = fma(d1, d1, a[g][1]);
        }
    }
    const double w0 = warp_sum(a[0][0] + a[0][1]), w1 = warp_sum(a[1][0] + a[1][1]);
    const double w2 = warp_sum(a[2][0] + a[2][1]), w3 = warp_sum(a[3][0] + a[3][1]);
    return ((w0 + w1) + w2) + w3;
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
};

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
    const double K = static_cast<double>(p.kp);
    // fp32 accumulation error of the MMA (K terms of magnitude <= ||q~|| ||x~||, x2 for the -2 factor, truncating
    // adds assumed), of the fp32 norm sums, and of forming s~ in fp32
    m.eps_acc = (K + 8.0) * 2.4e-7 * sqrt(m.qn_bf * xn_bf) * 1.001 + (K / 16.0 + 8.0) * 1.2e-7 * (xn_bf + m.qn_bf);
    m.eta = (static_cast<double>(p.q_err[q]) + static_cast<double>(__uint_as_float(*p.max_x_err_bits))) * (1.0 + 1e-6) + 1e-30;
    return m;
}

// Final step of a re-rank, executed by ONE full warp: rank the C exact squared distances by (d2, row), emit the best kk,
// and certify the answer (or queue the query for the second pass).  keysC: the C best-scored shortlist entries,
// ascending by (score, row); d2s: their exact squared distances (DBL_MAX = pruned / empty).
template <int C>
__device__ __forceinline__ void rerank_finish(const RerankParams &p, int q, const unsigned long long *keys, const double *d2s, int lane) {
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
            p.out_idx[static_cast<int64_t>(q) * p.kk + rank[h]] = static_cast<int32_t>(p.index_base + myi[h]);
            p.out_dist[static_cast<int64_t>(q) * p.kk + rank[h]] = (p.flags & 1u) ? myd[h] : sqrt(myd[h]);
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
        const double dk = sqrt(dk2);
        if (p.n > C && kc != ~0ull && mk != 0) {
            const double lb = em.lower(static_cast<double>(float_from_order_bits(static_cast<uint32_t>(kc >> 32))));
            certified = (lb > 0.0) && (dk < lb);
        } else if (mk == 0) {
            certified = (p.n <= C);
        }
        if (!certified) {
            // second pass collects every row with score <= thr: any x with d(q,x) <= dk has
            // ||q~ - x~|| <= dk + eta, i.e. s~ <= (dk + eta)^2 - ||q~||^2 + eps_acc.
            const double t = (dk + em.eta) * (dk + em.eta) - em.qn_bf + em.eps_acc;
            const int slot = atomicAdd(p.uncert_count, 1);
            p.uncert_list[slot] = q;
            p.uncert_thr[slot] = __double2float_ru(t + 1e-6 * fabs(t));
        }
    }
}

template <typename TX, typename TQ, int C, int NT>
__global__ void __launch_bounds__(NT)
rerank_kernel(const TX *__restrict__ x, const TQ *__restrict__ qmat, const RerankParams p) {
    extern __shared__ unsigned long long keys[];   // next_pow2(max_slots * C) entries (host-sized, <= MAX_KEYS)
    __shared__ double d2s[C];
    __shared__ int m_s;
    const int q = blockIdx.x;
    const int tid = threadIdx.x;
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
    if (tid == 0) {
        int m = min(C, p.kk);
        if (keys[p.kk - 1] != ~0ull) {
            const ErrModel em = make_err_model(p, q);
            const double u = em.upper(static_cast<double>(float_from_order_bits(static_cast<uint32_t>(keys[p.kk - 1] >> 32))));
            while (m < C && keys[m] != ~0ull &&
                   em.lower(static_cast<double>(float_from_order_bits(static_cast<uint32_t>(keys[m] >> 32)))) <= u) m++;
        } else {
            m = C;
        }
        m_s = m;
    }
    __syncthreads();
    int m = m_s;
    // exact float64 distances of the surviving candidates (the arithmetic of util.c:62-69, tree-summed).  The whole
    // block works on one candidate at a time: every thread owns a strided slice of the dimensions, keeps its slice
    // of the query row in registers across candidates, and issues its loads of the pool row back to back.
    const int warp = tid >> 5, lane = tid & 31;
    const TQ *qr = qmat + static_cast<int64_t>(q) * p.ld_q;
    constexpr int RQ = 24;                      // dims per thread held in registers (128 threads x 24 = 3072)
    constexpr int nth = 128;                    // the canonical 128 lanes (see canon_d2); extra threads of the 1024-thread
    const bool act = tid < nth;                 // flavour only help with the merge sort above
    __shared__ double partial[4];
    double qreg[RQ];
    const bool fits = p.dim <= RQ * nth;
    if (fits && act) {
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
    for (int c = 0; c < C; c++) {
        if (c == p.kk && c < m) {               // uniform across the block
            // kk exact distances are known: the kk-th true distance is at most their maximum, which is a far tighter
            // pruning limit than the a-priori upper bound (the survivors stay a prefix: scores are sorted)
            __syncthreads();                    // d2s[0..kk) were written by thread 0, possibly without a barrier since
            if (warp == 0) {
                double mx = 0.0;
                for (int i = lane; i < p.kk; i += 32) mx = fmax(mx, d2s[i]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                int cnt = 0;
                if (mx < DBL_MAX) {
                    const ErrModel em = make_err_model(p, q);
                    const double dk = sqrt(mx);
                    for (int i = p.kk + lane; i < m; i += 32)
                        cnt += (keys[i] != ~0ull && em.lower(static_cast<double>(float_from_order_bits(static_cast<uint32_t>(keys[i] >> 32)))) <= dk) ? 1 : 0;
                } else {
                    cnt = (lane == 0) ? m - p.kk : 0;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                if (lane == 0) m_s = p.kk + cnt;
            }
            __syncthreads();
            m = m_s;
        }
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
    if (warp == 0) rerank_finish<C>(p, q, keys, d2s, lane);
}

// Second pass, part 2: exact re-rank of the collected lists.  One block per uncertified query (list slot).
// Lists longer than the capacity (or shorter than kk) are handed to the exact scan via the overflow list.
constexpr int COLLECT_CAP = 1024;
struct CollectRerankParams {
    const int *uncert_list;      // [nun] query rows
    const int *coll_count;       // [nun]
    const int *coll_idx;         // [nun][COLLECT_CAP]
    int dim;
    int64_t ld_x, ld_q;
    int kk;
    int64_t index_base;
    unsigned flags;
    int32_t *out_idx;
    double *out_dist;
    int *overflow_count;
    int *overflow_list;          // query rows that need the exact scan
};

template <typename TX, typename TQ>
__global__ void __launch_bounds__(256)
rerank_collect_kernel(const TX *__restrict__ x, const TQ *__restrict__ qmat, const CollectRerankParams p) {
    __shared__ double d2[COLLECT_CAP];
    __shared__ int idx[COLLECT_CAP];
    __shared__ double sd[8];
    __shared__ int si[8];
    __shared__ double last_d_s;
    __shared__ int last_i_s;
    const int slot = blockIdx.x;
    const int q = p.uncert_list[slot];
    const int cnt = p.coll_count[slot];
    if (cnt > COLLECT_CAP || cnt < p.kk) {
        if (threadIdx.x == 0) {
            const int o = atomicAdd(p.overflow_count, 1);
            p.overflow_list[o] = q;
        }
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TQ *qr = qmat + static_cast<int64_t>(q) * p.ld_q;
    for (int c = warp; c < cnt; c += 8) {       // one candidate per warp, canonical summation order
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
            for (int w = 1; w < 8; w++)
                if (sd[w] < bd || (sd[w] == bd && si[w] < bi)) { bd = sd[w]; bi = si[w]; }
            last_d_s = bd;
            last_i_s = bi;
            p.out_idx[static_cast<int64_t>(q) * p.kk + r] = static_cast<int32_t>(p.index_base + bi);
            p.out_dist[static_cast<int64_t>(q) * p.kk + r] = (p.flags & 1u) ? bd : sqrt(bd);
        }
        __syncthreads();
    }
}

// gather BF16 query rows of the uncertified queries into a compact matrix for the collection pass
__global__ void __launch_bounds__(256)
gather_rows_kernel(const __nv_bfloat16 *__restrict__ src, const int *__restrict__ list, int nsel, int kp, __nv_bfloat16 *__restrict__ dst) {
    const int vec_per_row = kp >> 3;   // kp is a multiple of 8: 16-byte chunks
    const int64_t total = static_cast<int64_t>(nsel) * vec_per_row;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / vec_per_row), c = static_cast<int>(i % vec_per_row);
        reinterpret_cast<uint4 *>(dst + static_cast<int64_t>(r) * kp)[c] =
            reinterpret_cast<const uint4 *>(src + static_cast<int64_t>(list[r]) * kp)[c];
    }
}

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
        const volatile unsigned int *f = flags + threadIdx.x * 32;
        const uint64_t t0 = global_timer_ns();
        while (*f < step) {                  // steps only grow; a peer one step ahead is fine
            __nanosleep(200);
            if (global_timer_ns() - t0 > 20000000000ull) __trap();   // 20 s: a peer died
        }
    }
    __syncthreads();
    __threadfence();
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
