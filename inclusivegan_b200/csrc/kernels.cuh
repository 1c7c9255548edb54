// Device code of libb200knn: the kernels of the IMLE matching path.
//
//   convert_norm_kernel   HBM-bound.  rows of f64/f32 -> BF16 rows (TMA-friendly pitch) + ||x~||^2 (fp32, of the
//                         rounded values) + ||x||^2 of the unrounded values (for the exactness certificate).
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
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KB
constexpr int TMEM_COLS = 512;               // two 128 x 256 fp32 accumulators
constexpr int DIST_THREADS = 192;            // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int DIST_SMEM_BYTES = 1024 /*align slack*/ + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 2 * BN * 4 + 256;
constexpr int MAX_CHUNKS = 64;               // shortlists per query the rerank kernel can merge

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

// ------------------------------------------------------------------------------------------------
// Kernel 1: convert + norms.  One warp per row, 8 elements (one 16-byte BF16 store) per lane per step.
// Algorithmic bytes per row: dim * (sizeof(T) + 2) + 8.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void load8(const T *p, float (&f)[8], double &ex);

template <>
__device__ __forceinline__ void load8<double>(const double *p, float (&f)[8], double &ex) {
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
    double2 v0 = __ldcs(p2), v1 = __ldcs(p2 + 1), v2 = __ldcs(p2 + 2), v3 = __ldcs(p2 + 3);
    double d[8] = {v0.x, v0.y, v1.x, v1.y, v2.x, v2.y, v3.x, v3.y};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        ex = fma(d[i], d[i], ex);
        f[i] = static_cast<float>(d[i]);
    }
}
template <>
__device__ __forceinline__ void load8<float>(const float *p, float (&f)[8], double &ex) {
    const float4 *p4 = reinterpret_cast<const float4 *>(p);
    float4 v0 = __ldcs(p4), v1 = __ldcs(p4 + 1);
    f[0] = v0.x; f[1] = v0.y; f[2] = v0.z; f[3] = v0.w;
    f[4] = v1.x; f[5] = v1.y; f[6] = v1.z; f[7] = v1.w;
#pragma unroll
    for (int i = 0; i < 8; i++) ex = fma(static_cast<double>(f[i]), static_cast<double>(f[i]), ex);
}

// vec != 0 requires: dim % 8 == 0 (so kp == dim), src rows 16-byte aligned.
template <typename T>
__global__ void __launch_bounds__(256)
convert_norm_kernel(const T *__restrict__ src, int64_t n, int64_t ld, int dim, int kp, int vec,
                    __nv_bfloat16 *__restrict__ dst, float *__restrict__ norm_bf, float *__restrict__ norm_ex,
                    unsigned int *__restrict__ max_norm_bf_bits, unsigned int *__restrict__ max_norm_ex_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_per_grid = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
    float mx_bf = 0.f, mx_ex = 0.f;
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps_per_grid) {
        const T *s = src + row * ld;
        __nv_bfloat16 *d = dst + row * kp;
        float acc = 0.f;      // sum of squares of the ROUNDED values (exact products, fp32 accumulation)
        double ex = 0.0;      // sum of squares of the unrounded values
        if (vec) {
            const int groups = dim >> 3;
            for (int g = lane; g < groups; g += 32) {
                float f[8];
                load8<T>(s + (g << 3), f, ex);
                __nv_bfloat162 b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    b[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                    const float lo = __low2float(b[i]), hi = __high2float(b[i]);
                    acc = fmaf(lo, lo, acc);
                    acc = fmaf(hi, hi, acc);
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
                float f = 0.f;
                if (e < dim) {
                    const T v = s[e];
                    ex = fma(static_cast<double>(v), static_cast<double>(v), ex);
                    f = static_cast<float>(v);
                }
                const __nv_bfloat16 b = __float2bfloat16_rn(f);
                const float fb = __bfloat162float(b);
                acc = fmaf(fb, fb, acc);
                d[e] = b;
            }
        }
        acc = warp_sum(acc);
        ex = warp_sum(ex);
        const float exf = __double2float_ru(ex);
        if (lane == 0) {
            norm_bf[row] = acc;
            norm_ex[row] = exf;
        }
        mx_bf = fmaxf(mx_bf, acc);
        mx_ex = fmaxf(mx_ex, exf);
    }
    if (lane == 0) {   // non-negative floats order like their bit patterns
        if (mx_bf > 0.f) atomicMax(max_norm_bf_bits, __float_as_uint(mx_bf));
        if (mx_ex > 0.f) atomicMax(max_norm_ex_bits, __float_as_uint(mx_ex));
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel 2: BF16 distance GEMM on tcgen05 with fused top-C epilogue.
// ------------------------------------------------------------------------------------------------
struct DistParams {
    const float *xnorm;      // [n] ||x~||^2
    int n;                   // pool rows in this shard
    int nq;                  // query rows
    int num_kb;              // ceil(kp / BK)
    int num_qtiles;          // ceil(nq / BM)
    int num_ntiles;          // ceil(n / BN)
    int tiles_per_chunk;     // N tiles swept per work item
    int num_chunks;          // ceil(num_ntiles / tiles_per_chunk)  (<= MAX_CHUNKS)
    int qgroup;              // query tiles scheduled together (L2 working-set control)
    float *cand_s;           // [nq][num_chunks][C] approximate scores, ascending
    int *cand_i;             // [nq][num_chunks][C] shard-local row index (-1 = empty slot)
};

struct WorkItem { int qtile, chunk; };
__device__ __forceinline__ WorkItem decode_item(int it, const DistParams &p) {
    const int per_group = p.qgroup * p.num_chunks;
    const int g = it / per_group;
    const int r = it - g * per_group;
    const int gsize = min(p.qgroup, p.num_qtiles - g * p.qgroup);
    WorkItem w;
    w.chunk = r / gsize;
    w.qtile = g * p.qgroup + (r - w.chunk * gsize);
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

template <int C>
__global__ void __launch_bounds__(DIST_THREADS, 1)
dist_topc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x, const DistParams p) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles must sit on 1024-byte boundaries
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t smem_a = base;
    const uint32_t smem_b = base + STAGES * A_STAGE_BYTES;
    float *xn_s = reinterpret_cast<float *>(gen + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));   // [2][BN]
    const uint32_t bars = base + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 2 * BN * 4;
    const uint32_t bar_full = bars;                       // [STAGES]  TMA -> MMA
    const uint32_t bar_empty = bars + 8 * STAGES;         // [STAGES]  MMA -> TMA
    const uint32_t bar_tfull = bars + 16 * STAGES;        // [2]       MMA -> epilogue
    const uint32_t bar_tempty = bars + 16 * STAGES + 16;  // [2]       epilogue -> MMA
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 2 * BN * 4 + 16 * STAGES + 32);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_x);
        for (int s = 0; s < STAGES; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 128);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc<1>(smem_u32(const_cast<uint32_t *>(tmem_slot)), TMEM_COLS);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_items = p.num_qtiles * p.num_chunks;

    if (warp == 0) {
        // ===================== TMA producer (one lane) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
                const WorkItem w = decode_item(it, p);
                const int q0 = w.qtile * BM;
                const int t0 = w.chunk * p.tiles_per_chunk;
                const int t1 = min(t0 + p.tiles_per_chunk, p.num_ntiles);
                for (int t = t0; t < t1; t++) {
                    const int n0 = t * BN;
                    for (int kb = 0; kb < p.num_kb; kb++) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        mbar_expect_tx(bar_full + 8 * stage, A_STAGE_BYTES + B_STAGE_BYTES);
                        tma_load_2d(smem_a + stage * A_STAGE_BYTES, &tmap_q, bar_full + 8 * stage, kb * BK, q0);
                        tma_load_2d(smem_b + stage * B_STAGE_BYTES, &tmap_x, bar_full + 8 * stage, kb * BK, n0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one lane) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
                const WorkItem w = decode_item(it, p);
                const int t0 = w.chunk * p.tiles_per_chunk;
                const int t1 = min(t0 + p.tiles_per_chunk, p.num_ntiles);
                for (int t = t0; t < t1; t++) {
                    mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);     // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * BN;
                    for (int kb = 0; kb < p.num_kb; kb++) {
                        mbar_wait(bar_full + 8 * stage, phase);        // TMA bytes have landed
                        tc_fence_after();
                        const uint64_t da = make_smem_desc_sw128(smem_a + stage * A_STAGE_BYTES);
                        const uint64_t db = make_smem_desc_sw128(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; k++) {
                            // +32 bytes per K slice inside the 128-byte swizzle row: +2 in the (>>4) address field
                            umma_bf16<1>(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        }
                        umma_commit(bar_empty + 8 * stage);            // frees the smem slot when the MMAs retire
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(bar_tfull + 8 * acc);                  // accumulator complete -> epilogue
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
        for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
            const WorkItem w = decode_item(it, p);
            const int t0 = w.chunk * p.tiles_per_chunk;
            const int t1 = min(t0 + p.tiles_per_chunk, p.num_ntiles);
            float v[C];
            int id[C];
#pragma unroll
            for (int i = 0; i < C; i++) { v[i] = FLT_MAX; id[i] = -1; }
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
#pragma unroll 1
                for (int c = 0; c < BN / 32; c++) {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float s = fmaf(-2.f, __uint_as_float(r[j]), xs[c * 32 + j]);
                        if (s < v[C - 1]) topc_insert<C>(v, id, s, n0 + c * 32 + j);
                    }
                }
                tc_fence_before();
                mbar_arrive(bar_tempty + 8 * acc);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
            const int q = w.qtile * BM + row_in_tile;
            if (q < p.nq) {
                float *cs = p.cand_s + (static_cast<int64_t>(q) * p.num_chunks + w.chunk) * C;
                int *ci = p.cand_i + (static_cast<int64_t>(q) * p.num_chunks + w.chunk) * C;
#pragma unroll
                for (int i = 0; i < C; i += 4) {
                    *reinterpret_cast<float4 *>(cs + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    *reinterpret_cast<int4 *>(ci + i) = make_int4(id[i], id[i + 1], id[i + 2], id[i + 3]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<1>(tmem_base, TMEM_COLS);
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

struct RerankParams {
    const float *cand_s;
    const int *cand_i;
    int num_chunks;
    int dim;
    int64_t ld_x, ld_q;
    int n;                         // pool rows in the shard
    int kk;                        // neighbours to emit (<= C)
    int64_t index_base;
    unsigned flags;                // B200KNN_FLAG_*
    const float *qnorm_bf;         // [nq] ||q~||^2 (fp32, of rounded values)
    const float *qnorm_ex;         // [nq] ||q||^2 rounded up
    const unsigned int *max_xnorm_bf_bits;   // device scalars (pool)
    const unsigned int *max_xnorm_ex_bits;
    int kp;                        // padded K of the BF16 operands (accumulation length)
    int32_t *out_idx;              // [nq][kk]
    double *out_dist;              // [nq][kk]
    int *uncert_count;             // number of uncertified queries
    int *uncert_list;              // their row numbers
};

template <typename TX, typename TQ, int C>
__global__ void __launch_bounds__(128)
rerank_kernel(const TX *__restrict__ x, const TQ *__restrict__ qmat, const RerankParams p) {
    constexpr int MAXP = MAX_CHUNKS * C;
    __shared__ unsigned long long keys[MAXP];
    __shared__ double d2s[C];
    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    const int total = p.num_chunks * C;
    int P = 1;
    while (P < total) P <<= 1;

    for (int i = tid; i < P; i += blockDim.x) {
        unsigned long long key = ~0ull;
        if (i < total) {
            const int64_t o = static_cast<int64_t>(q) * total + i;
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
    // exact float64 distances of the C best-scored candidates
    const int warp = tid >> 5, lane = tid & 31;
    const TQ *qr = qmat + static_cast<int64_t>(q) * p.ld_q;
    for (int c = warp; c < C; c += 4) {
        const unsigned long long key = keys[c];
        double acc = 0.0;
        if (key != ~0ull) {
            const TX *xr = x + static_cast<int64_t>(static_cast<uint32_t>(key)) * p.ld_x;
            for (int e = lane; e < p.dim; e += 32) {
                const double diff = static_cast<double>(qr[e]) - static_cast<double>(xr[e]);
                acc = fma(diff, diff, acc);
            }
            acc = warp_sum(acc);
        } else {
            acc = DBL_MAX;
        }
        if (lane == 0) d2s[c] = acc;
    }
    __syncthreads();
    if (warp == 0) {
        // rank the C exact distances by (d2, index); C <= 32
        double myd = DBL_MAX;
        uint32_t myi = 0xffffffffu;
        if (lane < C) { myd = d2s[lane]; myi = static_cast<uint32_t>(keys[lane]); }
        int rank = 0;
#pragma unroll
        for (int o = 0; o < C; o++) {
            const double od = __shfl_sync(0xffffffffu, myd, o);
            const uint32_t oi = __shfl_sync(0xffffffffu, myi, o);
            rank += (od < myd || (od == myd && oi < myi)) ? 1 : 0;
        }
        if (lane < C && rank < p.kk) {
            p.out_idx[static_cast<int64_t>(q) * p.kk + rank] = static_cast<int32_t>(p.index_base + myi);
            p.out_dist[static_cast<int64_t>(q) * p.kk + rank] = (p.flags & 1u) ? myd : sqrt(myd);
        }
        // k-th exact distance (rank kk-1), broadcast to lane 0
        const unsigned m = __ballot_sync(0xffffffffu, lane < C && rank == p.kk - 1);
        const double dk2 = __shfl_sync(0xffffffffu, myd, m ? (__ffs(m) - 1) : 0);
        if (lane == 0 && !(p.flags & 2u)) {
            // ---- certificate --------------------------------------------------------------------
            // every pool row NOT among the C kept has approximate score >= tau (the C-th kept score).
            // In exact arithmetic  s~ + ||q~||^2 = ||q~ - x~||^2; rounding to BF16 moves a vector by at
            // most 2^-9 of its norm, so  d(q,x) >= ||q~ - x~|| - 2^-9 (||q|| + ||x||).
            bool certified = true;
            const unsigned long long kc = keys[C - 1];
            if (p.n > C && kc != ~0ull && m != 0) {
                const double tau = static_cast<double>(float_from_order_bits(static_cast<uint32_t>(kc >> 32)));
                const double qn_bf = static_cast<double>(p.qnorm_bf[q]);
                const double qn_ex = static_cast<double>(p.qnorm_ex[q]);
                const double xn_bf = static_cast<double>(__uint_as_float(*p.max_xnorm_bf_bits));
                const double xn_ex = static_cast<double>(__uint_as_float(*p.max_xnorm_ex_bits));
                const double K = static_cast<double>(p.kp);
                // fp32 accumulation error of the MMA (K terms, magnitude <= ||q~|| ||x~||, x2 for the -2 factor),
                // of the fp32 norm sums, and of forming s~ in fp32
                const double eps_acc = (K + 8.0) * 2.4e-7 * sqrt(qn_bf * xn_bf) * 1.001
                                     + (K / 16.0 + 8.0) * 1.2e-7 * (xn_bf + qn_bf);
                const double lb2 = tau + qn_bf - eps_acc;
                const double eta = (1.0 / 512.0) * 1.0001 * (sqrt(qn_ex) + sqrt(xn_ex)) + 1e-30;
                const double lb = (lb2 > 0.0 ? sqrt(lb2) : 0.0) - eta;
                certified = (lb > 0.0) && (sqrt(dk2) < lb);
            } else if (m == 0) {
                certified = (p.n <= C);   // fewer than kk exact candidates can only happen for tiny pools
            }
            if (!certified) {
                const int slot = atomicAdd(p.uncert_count, 1);
                p.uncert_list[slot] = q;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Exact float64 scan (CUDA cores).  Used for uncertified queries and for k > 16.
//   scan_dist_kernel : d2[s][j] = sum_e (q[list[s]][e] - x[j][e])^2     (32 queries x 64 pool rows per block)
//   scan_select_kernel: kk passes of lexicographic (d2, index) arg-min  (kk <= 32)
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_TQ = 32, SCAN_TX = 64, SCAN_TK = 16;

template <typename TX, typename TQ>
__global__ void __launch_bounds__(256)
scan_dist_kernel(const TX *__restrict__ x, int64_t ld_x, int n, const TQ *__restrict__ qmat, int64_t ld_q,
                 const int *__restrict__ qlist, int nsub, int dim, double *__restrict__ d2) {
    __shared__ double qs[SCAN_TQ][SCAN_TK + 1];
    __shared__ double xs[SCAN_TX][SCAN_TK + 1];
    const int tx = threadIdx.x & 15;    // 16 column groups x 4 pool rows
    const int ty = threadIdx.x >> 4;    // 16 row groups x 2 queries
    const int x0 = blockIdx.x * SCAN_TX;
    const int s0 = blockIdx.y * SCAN_TQ;
    double acc[2][4] = {};
    for (int k0 = 0; k0 < dim; k0 += SCAN_TK) {
        for (int i = threadIdx.x; i < SCAN_TQ * SCAN_TK; i += 256) {
            const int r = i / SCAN_TK, c = i % SCAN_TK;
            const int s = s0 + r, e = k0 + c;
            double v = 0.0;
            if (s < nsub && e < dim) {
                const int qrow = qlist ? qlist[s] : s;
                v = static_cast<double>(qmat[static_cast<int64_t>(qrow) * ld_q + e]);
            }
            qs[r][c] = v;
        }
        for (int i = threadIdx.x; i < SCAN_TX * SCAN_TK; i += 256) {
            const int r = i / SCAN_TK, c = i % SCAN_TK;
            const int j = x0 + r, e = k0 + c;
            xs[r][c] = (j < n && e < dim) ? static_cast<double>(x[static_cast<int64_t>(j) * ld_x + e]) : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < SCAN_TK; c++) {
            const double q0 = qs[ty * 2][c], q1 = qs[ty * 2 + 1][c];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const double xv = xs[tx * 4 + b][c];
                const double d0 = q0 - xv, d1 = q1 - xv;
                acc[0][b] = fma(d0, d0, acc[0][b]);
                acc[1][b] = fma(d1, d1, acc[1][b]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 2; a++) {
        const int s = s0 + ty * 2 + a;
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

}  // namespace b200
