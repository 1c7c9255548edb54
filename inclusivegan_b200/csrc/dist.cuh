// Kernel 2 — BF16 distance GEMM on tcgen05/TMEM fed by TMA, fused top-C / threshold-collect epilogue (tensor-bound).
#pragma once
#include "common.cuh"

namespace b200 {
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
    const int *nq_dev;       // non-null: the number of valid query rows is read from device memory (second pass: the
                             // uncertified count of the first pass; the host does not synchronise to learn it)
    // Precision tiers.  Split BF16 (bf16x3): every operand row is hi + lo (two BF16 rows); the kernel runs the
    // K loop over THREE segments of nkb_seg blocks — [q_hi | q_hi | q_lo] . [x_hi | x_lo | x_hi] = q_hi.x_hi + q_hi.x_lo +
    // q_lo.x_hi — by pointing the TMA loads of a segment at the hi or lo tensor map; MMA and epilogue do not change.
    int nkb_seg;             // K blocks per segment (== num_kb: plain; num_kb == 3 * nkb_seg: split)
    // K-chunked accumulation: the tensor core accumulates at most kb_per_unit K blocks into one TMEM accumulator; the
    // epilogue warps add the partial sums in fp32 registers (IEEE round-to-nearest) into a running accumulator kept in the
    // other half of TMEM.  The certificate's bound on the accumulation error then scales with the unit length instead
    // of the row length (Cauchy-Schwarz over the units) — what keeps long rows (d = 49152) on the tensor path.
    int kb_per_unit;         // >= num_kb: one unit (no chunking)
};

// Device-side planner of a pass (same layout as the host's Shard::plan_schedule, which sizes the buffers): per round a
// group of up to `group` query tiles x rc pool-tile streams, chunk-major (workers sharing a stream are neighbours).
// The schedule is WRITTEN on the device instead of copied to it: a host-to-device copy of a few KB would queue behind
// the multi-megabyte query uploads on the same copy engine (measured: 1 ms stalls per chunk at 8 GPUs).  For the
// second (collection) pass the number of query rows is only known on the device (count_dev).  The kernel also zeroes
// every counter the pass uses, so a pass needs no memset either.  One block.
struct PlanParams {
    const int *count_dev;          // non-null: number of query rows (second pass: the uncertified count); else count
    int count;
    int qrows, nt, workers, group, max_rounds, wide, max_slots_cap;
    WorkItem *items;               // [max_rounds][workers]
    int *slots_per_qtile;          // nullable [ceil(count / qrows)]
    unsigned int *round_counter;   // zeroed
    unsigned int *stream_sync;     // [sync_entries] zeroed
    int sync_entries;
    unsigned int *zero_a;          // nullable single counters to zero (first pass: the uncertified count)
    unsigned int *zero_b;          // (first pass: the work counter of the persistent re-rank warps)
    int *zero_list;                // nullable [zero_list_n] (second pass: coll_count)
    int zero_list_n;
    unsigned int *total_uncertified;   // nullable: += count (second pass accounting)
};
__global__ void __launch_bounds__(256)
plan_pass_kernel(const PlanParams p) {
    const int cnt = p.count_dev ? *p.count_dev : p.count;
    if (threadIdx.x == 0) {
        if (p.total_uncertified && cnt > 0) atomicAdd(p.total_uncertified, static_cast<unsigned int>(cnt));
        *p.round_counter = 0u;
        if (p.zero_a) *p.zero_a = 0u;
        if (p.zero_b) *p.zero_b = 0u;
    }
    for (int i = threadIdx.x; i < p.sync_entries; i += blockDim.x) p.stream_sync[i] = 0u;
    for (int i = threadIdx.x; i < p.zero_list_n; i += blockDim.x) p.zero_list[i] = 0;
    const int qt = (cnt + p.qrows - 1) / p.qrows;
    for (int i = threadIdx.x; i < p.max_rounds * p.workers; i += blockDim.x) {
        const int round = i / p.workers, w = i % p.workers;
        const int q0 = round * p.group;
        const int gs = min(p.group, qt - q0);
        WorkItem it{-1, 0, 0, 0};
        if (gs > 0) {
            int rc = max(1, min(min(p.workers / gs, p.nt), p.max_slots_cap));
            if (p.wide && gs * rc > 255) rc = max(1, 255 / gs);
            const int c = w / gs, g = w % gs;
            if (c < rc) {
                const int wide_bits = (p.wide && gs > 1 && rc > 1) ? static_cast<int>(static_cast<unsigned int>(gs * rc) << 24) : 0;
                it.qtile = q0 + g;
                it.t0 = static_cast<int>(static_cast<int64_t>(c) * p.nt / rc);
                it.t1 = static_cast<int>(static_cast<int64_t>(c + 1) * p.nt / rc);
                it.slot = c | (gs << 16) | wide_bits;
                if (c == 0 && p.slots_per_qtile) p.slots_per_qtile[q0 + g] = rc;
            }
        }
        p.items[i] = it;
    }
}

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

template <int C, bool COLLECT, int CG, bool TF32 = false>
__global__ void __launch_bounds__(DIST_THREADS, 1)
dist_topc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                 const __grid_constant__ CUtensorMap tmap_qlo, const __grid_constant__ CUtensorMap tmap_xlo, const DistParams p) {
    using Cfg = DistCfg<CG>;
    constexpr int NSTAGE = Cfg::STAGES;
    constexpr int KELEMS = TF32 ? BK / 2 : BK;      // elements per 128-byte stage row: 64 bf16 or 32 tf32 (fp32 containers)
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
        tma_prefetch_desc(&tmap_qlo);
        tma_prefetch_desc(&tmap_xlo);
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
                        int seg = 0, kk = 0;                  // split tiers: segment 0 = hi.hi, 1 = q_hi.x_lo, 2 = q_lo.x_hi
                        for (int kb = 0; kb < p.num_kb; kb++) {
                            const CUtensorMap *ma = (seg == 2) ? &tmap_qlo : &tmap_q;
                            const CUtensorMap *mb = (seg == 1) ? &tmap_xlo : &tmap_x;
                            const int kcoord = kk * KELEMS;
                            if (++kk == p.nkb_seg) { kk = 0; seg++; }
                            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                            if (CG == 1) {
                                mbar_expect_tx(bar_full + 8 * stage, Cfg::STAGE_BYTES);
                                tma_load_2d(smem_a + stage * Cfg::A_BYTES, ma, bar_full + 8 * stage, kcoord, q0);
                                tma_load_2d(smem_b + stage * Cfg::B_BYTES, mb, bar_full + 8 * stage, kcoord, n0);
                            } else {
                                // all four loads of the pair (2 x A half, 2 x B half) signal the LEADER's barrier; the
                                // peer never arrives there: its loads only complete_tx (a remote arrive per K block
                                // would cost a cluster-scope fence each time)
                                if (leader) mbar_expect_tx(bar_full + 8 * stage, 2 * Cfg::STAGE_BYTES);
                                const uint32_t fb = leader ? (bar_full + 8 * stage) : (full_remote + 8 * stage);
                                tma_load_2d_cg2(smem_a + stage * Cfg::A_BYTES, ma, fb, kcoord, q0);
                                tma_load_2d_cg2(smem_b + stage * Cfg::B_BYTES, mb, fb, kcoord, n0);
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
            constexpr uint32_t idesc = TF32 ? make_idesc_tf32(BM * CG, BN) : make_idesc_bf16(BM * CG, BN);
            const bool chunked = p.kb_per_unit < p.num_kb;
            int stage = 0;
            uint32_t phase = 0;
            int tile_par = 0;                    // one unit per tile: the two accumulators alternate, the epilogue of t overlaps the MMAs of t+1
            uint32_t uses[2] = {0u, 0u};         // how often each accumulator has been handed to the epilogue (barrier parity)
            for (int round = 0; round < p.nrounds; round++) {
                const WorkItem w = load_item(p, round, worker);
                if (w.qtile < 0) continue;
                for (int t = w.t0; t < w.t1; t++) {
                    // chunked: unit 0 accumulates into buffer 0 (the running sum), every later unit into buffer 1 (a partial
                    // sum the epilogue folds into buffer 0)
                    int kb = 0;
                    for (int unit = 0; kb < p.num_kb; unit++) {
                        const int acc = chunked ? (unit ? 1 : 0) : tile_par;
                        const int kb_end = min(p.num_kb, kb + p.kb_per_unit);
                        mbar_wait(bar_tempty + 8 * acc, (uses[acc] & 1u) ^ 1u);     // epilogues have drained this accumulator
                        uses[acc]++;
                        tc_fence_after();
                        const uint32_t tmem_d = tmem_base + acc * BN;
                        for (bool first = true; kb < kb_end; kb++, first = false) {
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
                                    // (16 bf16 or 8 tf32 per slice: the same 32 bytes)
                                    if constexpr (TF32) umma_tf32<CG>(tmem_d, da + 2 * k, db + 2 * k, idesc, !(first && k == 0));
                                    else umma_bf16<CG>(tmem_d, da + 2 * k, db + 2 * k, idesc, !(first && k == 0));
                                }
                                // frees the smem slot (in both CTAs) when the MMAs retire
                                if (CG == 1) umma_commit(bar_empty + 8 * cur);
                                else umma_commit_cg2(bar_empty + 8 * cur, 0x3);
                                if (kb == kb_end - 1) {                        // unit complete -> epilogue(s)
                                    if (CG == 1) umma_commit(bar_tfull + 8 * acc);
                                    else umma_commit_cg2(bar_tfull + 8 * acc, 0x3);
                                }
                            }
                            __syncwarp();
                        }
                    }
                    tile_par ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue: 4 warps, thread <-> query row =====================
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may read
        const int row_in_tile = quarter * 32 + lane;
        const int nq_eff = p.nq_dev ? min(__ldcg(p.nq_dev), p.nq) : p.nq;
        const int et = threadIdx.x - 64;               // 0..127
        const bool chunked = p.kb_per_unit < p.num_kb;
        const int nunits = (p.num_kb + p.kb_per_unit - 1) / p.kb_per_unit;
        int tile_par = 0;
        uint32_t uses[2] = {0u, 0u};
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
                if (q < nq_eff) thr = __ldg(p.thr + q);
            } else {
#pragma unroll
                for (int i = 0; i < C; i++) { v[i] = FLT_MAX; id[i] = -1; }
            }
            for (int t = t0; t < t1; t++) {
                const int n0 = t * BN;
                const int acc = chunked ? 0 : tile_par;
                // stage ||x~||^2 of this tile; rows past the end of the pool can never be selected
                float *xs = xn_s + tile_par * BN;       // (two staging buffers whichever accumulator the tile uses)
                {
                    const int c0 = n0 + et, c1 = n0 + et + 128;
                    xs[et] = (c0 < p.n) ? __ldg(p.xnorm + c0) : FLT_MAX;
                    xs[et + 128] = (c1 < p.n) ? __ldg(p.xnorm + c1) : FLT_MAX;
                }
                named_bar_sync(1, 128);
                mbar_wait(bar_tfull + 8 * acc, uses[acc] & 1u);
                uses[acc]++;
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
                if (chunked) {
                    // fold the partial sums of units 1.. (buffer 1) into the running sum (buffer 0): fp32 adds in registers
                    const uint32_t paddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + BN;
                    for (int unit = 1; unit < nunits; unit++) {
                        mbar_wait(bar_tfull + 8, uses[1] & 1u);
                        uses[1]++;
                        tc_fence_after();
#pragma unroll 1
                        for (int c = 0; c < BN / 16; c++) {
                            uint32_t a[16], b[16];
                            tmem_ld_32x16(taddr + c * 16, a);
                            tmem_ld_32x16(paddr + c * 16, b);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; j++) a[j] = __float_as_uint(__fadd_rn(__uint_as_float(a[j]), __uint_as_float(b[j])));
                            tmem_st_32x16(taddr + c * 16, a);
                        }
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {                   // the partial accumulator is free for the next unit
                            if (CG == 1 || leader) mbar_arrive(bar_tempty + 8);
                            else mbar_arrive_cluster(bar_tempty + 8, 0);
                        }
                    }
                }
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
                tile_par ^= 1;
            }
            if constexpr (!COLLECT) {
                if (q < nq_eff) {
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

}  // namespace b200
