// One GPU's share of an index: the pool shard (originals, BF16 rows, norms), the work schedule of the distance kernel
// (Shard::plan), and the kernel sequence that answers a batch of device-resident queries (Shard::query_device).
#pragma once
#include "host_util.cuh"

namespace {

enum Kind { K_CONVERT = 0, K_DISTANCE = 1, K_RERANK = 2, K_SCAN = 3, K_WAIT = 4, K_NKINDS = 5 };

struct Shard {
    int device = 0;
    int num_sms = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;   // the one work is launched on (own or user-provided)
    bool ready = false;

    // ---- pool ----
    const void *x_raw = nullptr;     // original rows (f64 or f32), ld_x elements apart
    void *x_owned = nullptr;         // set when the library owns the copy
    DevBuf<unsigned char> x_store;   // the library's copy of the original rows: kept across clear() (the trainer re-adds a pool of
                                     // the same size at every refresh; freeing and re-allocating GBs costs tens of ms), freed by destroy()
    bool release_on_clear = false;   // $B200KNN_RELEASE_ON_CLEAR=1: give the memory back at clear()/reset()
    int x_dtype = B200KNN_F64;
    int64_t n = 0, ld_x = 0, index_base = 0;
    DevBuf<__nv_bfloat16> x_bf;
    DevBuf<float> xnorm_bf, x_err;
    // precision tier of the tensor pass (tiers.cuh): 0 bf16, 1 bf16x3 (split), 2 tf32; the tier's own operand copies
    int tier = 0;                    // $B200KNN_PRECISION / b200knn_set_precision
    int kc_elems = 4096;             // $B200KNN_KC: K elements the tensor core accumulates before the epilogue folds the partial sum (0 = never)
    DevBuf<__nv_bfloat16> x_lo, q_lo, q_lo2;
    DevBuf<float> x_tf, q_tf, q_tf2;
    DevBuf<float> xnorm_t, x_err_t, x_lonorm, qnorm_t, q_err_t, q_lonorm, qnorm_t2, q_err_t2, q_lonorm2;
    CUtensorMap tmap_xlo, tmap_xlo128, tmap_xtf, tmap_xtf128;
    DevBuf<double> col_mean;         // pool column means (subtracted from pool and queries before BF16 rounding)
    bool centered = false;
    bool use_centering = true;       // $B200KNN_CENTER=0 disables
    DevBuf<unsigned int> scalars;    // [0] max ||x~||^2 bits, [1] max ||x - x~|| bits, [2],[3] same for queries (unused), [4] uncertified count of
                                     // the pass in flight, [5] overflow count of the call, [6] grid-barrier counter of the distance kernel,
                                     // [7] max (r_j + e_j) bits (ball membership), [8] uncertified queries since the last stats reset,
                                     // [9] done counter of bound_publish_kernel, [10..12] tier maxima of the pool (||x^||^2, ||x - x^||,
                                     // ||x_lo||), [13..15] the same for the queries of the pass (unused), [16] next query of the
                                     // persistent re-rank warps
    CUtensorMap tmap_x;              // pool, 256-row box (1-CTA kernel)
    CUtensorMap tmap_x128;           // pool, 128-row box (each CTA of a pair stages half of the 256-row tile)
    int forced_cg = 0;               // $B200KNN_CTA_GROUP=1|2 pins the kernel flavour (A/B measurements)
    unsigned opt_flags = 0;          // $B200KNN_OPT: kernel tuning switches (see DistParams::opt)
    int a_budget_mb = 64;            // $B200KNN_A_BUDGET_MB: L2 budget for the query tiles of one round
    int wide_mode = 1;               // $B200KNN_WIDE: 0 never, 1 when cheaper in HBM traffic, 2 always (A/B measurements)
    int sync_tiles = -1;             // $B200KNN_SYNC_TILES: lockstep interval of the workers sharing a pool-tile stream (-1: by tile length)
    int max_pairs = 74;              // CTA pairs that can be co-resident (cudaOccupancyMaxActiveClusters)
    int rerank_warp_mode = 1;        // $B200KNN_RERANK_WARP: 0 block per query always, 1 warp per query for >= 2048 queries, 2 always

    // ---- query workspace ----
    DevBuf<__nv_bfloat16> q_bf;
    DevBuf<float> qnorm_bf, q_err;
    DevBuf<__nv_bfloat16> q_bf2;     // second pass: BF16 rows of the uncertified queries
    DevBuf<float> uncert_thr, min_score;
    DevBuf<int> coll_count, coll_idx, overflow_list;
    DevBuf<WorkItem> sched_items;    // first-pass schedule (cached across calls of the same shape: the trainer's 24-row loop)
    DevBuf<WorkItem> sched_items2;   // second pass / membership schedules
    DevBuf<double> radius2;          // ball membership: squared radii of this shard's rows
    DevBuf<float> colterm, rowthr;
    DevBuf<unsigned char> member;
    DevBuf<unsigned int> stream_sync;
    DevBuf<int> sched_slots;
    DevBuf<float> cand_s;
    DevBuf<int> cand_i;
    DevBuf<int> uncert_list;
    DevBuf<double> scan_d2;
    DevBuf<unsigned long long> scan_key;   // k > 32: (key, index) scratch of scan_topk_kernel, [queries][2][kk]
    DevBuf<int> scan_idx;
    DevBuf<double> projector;        // random projection (b200knn_set_projector): this device's copy of the projector [in_dim][dim],
    DevBuf<unsigned char> proj_stage;   // staging for unprojected rows, projected query rows
    DevBuf<double> proj_rows;
    DevBuf<unsigned char> q_stage;   // host API: device copy of the caller's query rows
    DevBuf<unsigned char> q_stage2;  // second buffer: the upload of chunk i+1 overlaps the compute of chunk i
    int64_t host_chunk_key[8] = {-1, -1, -1, -1, -1, -1, -1, -1};      // b200knn_query: cached upload ramp (plan_host_chunks) of the last call shape
    std::vector<std::pair<int64_t, int64_t>> host_chunks;
    cudaStream_t copy_stream = nullptr;
    int copy_threads = 8;            // $B200KNN_COPY_THREADS
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    DevBuf<int32_t> out_idx, pad_idx;
    DevBuf<double> out_dist, pad_dist;
    int *h_count = nullptr;          // pinned
    const __nv_bfloat16 *cur_q_bf = nullptr;   // BF16 query rows of the tensor pass in flight (second pass gathers from them)
    const __nv_bfloat16 *cur_q_lo = nullptr;   // ... their lo parts (split tier) / TF32 rows, and the tier that pass ran in
    const float *cur_q_tf = nullptr;
    int cur_tier = 0;
    int64_t last_nq = 0;             // geometry of the last tensor pass (b200knn_debug_shortlists)
    int last_slots = 0, last_c = 0;

    // ---- stats ----
    b200knn_stats stats{};
    bool profiling = false;
    struct Ev { cudaEvent_t a, b; int kind; double flops; };
    std::vector<Ev> events;

    int init(int dev) {
        device = dev;
        CU_TRY(cudaSetDevice(device));
        cudaDeviceProp prop;
        CU_TRY(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            return fail(B200KNN_ENODEVICE, "device %d (%s, sm_%d%d) is not an sm_100 (B200) GPU; libb200knn has no other code path",
                        device, prop.name, prop.major, prop.minor);
        num_sms = prop.multiProcessorCount;
        CU_TRY(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
        stream = own_stream;
        CU_TRY(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CU_TRY(cudaEventCreateWithFlags(&ev_copied[i], cudaEventDisableTiming));
            CU_TRY(cudaEventCreateWithFlags(&ev_consumed[i], cudaEventDisableTiming));
        }
        TRY(scalars.ensure(32));
        CU_TRY(cudaMemsetAsync(scalars.p, 0, 32 * sizeof(unsigned int), stream));
        CU_TRY(cudaMallocHost(reinterpret_cast<void **>(&h_count), 4 * sizeof(int)));
        TRY(set_kernel_attrs());
        ready = true;
        return B200KNN_OK;
    }
    // CUDA loads kernels lazily, at their first launch, and loading may synchronise the whole context.  The multi-GPU
    // protocol has kernels that spin until a peer (or another stream) has made progress: a first launch that lands behind
    // such a kernel would wait for it, and everything submitted after that waits too — a deadlock.  So every kernel of the
    // library is loaded here, once per process and device.
    int preload_kernels() {
        cudaFuncAttributes a;
#define B200_PRELOAD(k) CU_TRY(cudaFuncGetAttributes(&a, k))
#define B200_PRELOAD_T2(k) B200_PRELOAD((k<double, double>)); B200_PRELOAD((k<double, float>)); B200_PRELOAD((k<float, double>)); B200_PRELOAD((k<float, float>))
#define B200_PRELOAD_RR(C) B200_PRELOAD((rerank_kernel<double, double, C, 128>)); B200_PRELOAD((rerank_kernel<double, float, C, 128>)); \
        B200_PRELOAD((rerank_kernel<float, double, C, 128>)); B200_PRELOAD((rerank_kernel<float, float, C, 128>));                       \
        B200_PRELOAD((rerank_kernel<double, double, C, 1024>)); B200_PRELOAD((rerank_kernel<double, float, C, 1024>));                   \
        B200_PRELOAD((rerank_kernel<float, double, C, 1024>)); B200_PRELOAD((rerank_kernel<float, float, C, 1024>))
#define B200_PRELOAD_RW(C) B200_PRELOAD((rerank_warp_kernel<double, double, C, RR_WPB>)); B200_PRELOAD((rerank_warp_kernel<double, float, C, RR_WPB>)); \
        B200_PRELOAD((rerank_warp_kernel<float, double, C, RR_WPB>)); B200_PRELOAD((rerank_warp_kernel<float, float, C, RR_WPB>))
        B200_PRELOAD(colsum_kernel<double>); B200_PRELOAD(colsum_kernel<float>); B200_PRELOAD(scale_kernel);
        B200_PRELOAD(convert_norm_kernel<double>); B200_PRELOAD(convert_norm_kernel<float>);
        B200_PRELOAD(plan_pass_kernel); B200_PRELOAD(gather_rows_kernel);
        B200_PRELOAD((dist_topc_kernel<16, false, 1>)); B200_PRELOAD((dist_topc_kernel<32, false, 1>)); B200_PRELOAD((dist_topc_kernel<64, false, 1>));
        B200_PRELOAD((dist_topc_kernel<16, false, 2>)); B200_PRELOAD((dist_topc_kernel<32, false, 2>)); B200_PRELOAD((dist_topc_kernel<64, false, 2>));
        B200_PRELOAD((dist_topc_kernel<16, true, 1>)); B200_PRELOAD((dist_topc_kernel<16, true, 2>));
        B200_PRELOAD((dist_topc_kernel<16, false, 1, true>)); B200_PRELOAD((dist_topc_kernel<32, false, 1, true>)); B200_PRELOAD((dist_topc_kernel<64, false, 1, true>));
        B200_PRELOAD((dist_topc_kernel<16, false, 2, true>)); B200_PRELOAD((dist_topc_kernel<32, false, 2, true>)); B200_PRELOAD((dist_topc_kernel<64, false, 2, true>));
        B200_PRELOAD((dist_topc_kernel<16, true, 1, true>)); B200_PRELOAD((dist_topc_kernel<16, true, 2, true>));
        B200_PRELOAD((convert_tier_kernel<double, 1>)); B200_PRELOAD((convert_tier_kernel<float, 1>));
        B200_PRELOAD((convert_tier_kernel<double, 2>)); B200_PRELOAD((convert_tier_kernel<float, 2>));
        B200_PRELOAD_RR(16); B200_PRELOAD_RR(32); B200_PRELOAD_RR(64);
        B200_PRELOAD_RW(16); B200_PRELOAD_RW(32); B200_PRELOAD_RW(64);
        B200_PRELOAD((rerank_collect_kernel<double, double, 32>)); B200_PRELOAD((rerank_collect_kernel<double, float, 32>));
        B200_PRELOAD((rerank_collect_kernel<float, double, 32>)); B200_PRELOAD((rerank_collect_kernel<float, float, 32>));
        B200_PRELOAD_T2(scan_dist_kernel); B200_PRELOAD(scan_select_kernel); B200_PRELOAD(scan_topk_kernel);
        B200_PRELOAD(merge_topk_kernel); B200_PRELOAD(pad_topk_kernel); B200_PRELOAD(publish_topk_kernel); B200_PRELOAD(merge_wait_kernel);
        B200_PRELOAD(raise_flags_kernel); B200_PRELOAD(broadcast_segments_kernel); B200_PRELOAD(pull_rows_kernel); B200_PRELOAD(wait_flags_kernel); B200_PRELOAD(bound_publish_kernel);
        B200_PRELOAD(publish_colsum_kernel); B200_PRELOAD(global_mean_kernel);
        B200_PRELOAD(ball_colterm_kernel); B200_PRELOAD(ball_rowthr_kernel); B200_PRELOAD_T2(ball_member_kernel); B200_PRELOAD(scan_member_kernel);
        B200_PRELOAD(project_kernel<double>); B200_PRELOAD(project_kernel<float>);
#undef B200_PRELOAD_RR
#undef B200_PRELOAD_RW
#undef B200_PRELOAD_T2
#undef B200_PRELOAD
        return B200KNN_OK;
    }
    int set_kernel_attrs() {
        CU_TRY(cudaFuncSetAttribute(dist_topc_kernel<16, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<1>::SMEM_BYTES));
        CU_TRY(cudaFuncSetAttribute(dist_topc_kernel<32, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<1>::SMEM_BYTES));
        CU_TRY(cudaFuncSetAttribute(dist_topc_kernel<16, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<1>::SMEM_BYTES));
        CU_TRY(cudaFuncSetAttribute(dist_topc_kernel<16, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<2>::SMEM_BYTES));
        CU_TRY(cudaFuncSetAttribute(dist_topc_kernel<32, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<2>::SMEM_BYTES));
        CU_TRY(cudaFuncSetAttribute(dist_topc_kernel<16, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<2>::SMEM_BYTES));
        CU_TRY(cudaFuncSetAttribute(dist_topc_kernel<64, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<1>::SMEM_BYTES));
        CU_TRY(cudaFuncSetAttribute(dist_topc_kernel<64, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<2>::SMEM_BYTES));
        CU_TRY(cudaFuncSetAttribute((dist_topc_kernel<16, false, 1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<1>::SMEM_BYTES));
        CU_TRY(cudaFuncSetAttribute((dist_topc_kernel<32, false, 1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<1>::SMEM_BYTES));
        CU_TRY(cudaFuncSetAttribute((dist_topc_kernel<64, false, 1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<1>::SMEM_BYTES));
        CU_TRY(cudaFuncSetAttribute((dist_topc_kernel<16, true, 1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<1>::SMEM_BYTES));
        CU_TRY(cudaFuncSetAttribute((dist_topc_kernel<16, false, 2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<2>::SMEM_BYTES));
        CU_TRY(cudaFuncSetAttribute((dist_topc_kernel<32, false, 2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<2>::SMEM_BYTES));
        CU_TRY(cudaFuncSetAttribute((dist_topc_kernel<64, false, 2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<2>::SMEM_BYTES));
        CU_TRY(cudaFuncSetAttribute((dist_topc_kernel<16, true, 2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, DistCfg<2>::SMEM_BYTES));
        TRY(preload_kernels());
        if (const char *o = getenv("B200KNN_OPT")) opt_flags = static_cast<unsigned>(atoi(o));
        if (const char *o = getenv("B200KNN_A_BUDGET_MB")) a_budget_mb = std::max(1, atoi(o));
        if (const char *o = getenv("B200KNN_SYNC_TILES")) sync_tiles = std::max(0, atoi(o));
        if (const char *o = getenv("B200KNN_WIDE")) wide_mode = atoi(o);
        if (const char *o = getenv("B200KNN_COPY_THREADS")) copy_threads = std::max(1, atoi(o));
        if (const char *o = getenv("B200KNN_CENTER")) use_centering = atoi(o) != 0;
        if (const char *o = getenv("B200KNN_RELEASE_ON_CLEAR")) release_on_clear = atoi(o) != 0;
        if (const char *o = getenv("B200KNN_KC")) kc_elems = std::max(0, atoi(o));
        if (const char *o = getenv("B200KNN_RERANK_WARP")) rerank_warp_mode = std::max(0, std::min(2, atoi(o)));
        if (const char *o = getenv("B200KNN_PRECISION")) {
            if (!strcmp(o, "bf16x3") || !strcmp(o, "1")) tier = 1;
            else if (!strcmp(o, "tf32") || !strcmp(o, "2")) tier = 2;
        }
        copy_threads = std::min<int>(copy_threads, std::max(1u, std::thread::hardware_concurrency()));
        const char *e = getenv("B200KNN_CTA_GROUP");
        if (e && (e[0] == '1' || e[0] == '2')) forced_cg = e[0] - '0';
        // a persistent, statically-strided grid must be fully co-resident: ask how many CTA pairs fit at once
        {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(num_sms / 2 * 2);
            cfg.blockDim = dim3(DIST_THREADS);
            cfg.dynamicSmemBytes = DistCfg<2>::SMEM_BYTES;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int nc = 0;
            if (cudaOccupancyMaxActiveClusters(&nc, dist_topc_kernel<16, false, 2>, &cfg) == cudaSuccess && nc > 0) max_pairs = std::min(nc, num_sms / 2);
            else { cudaGetLastError(); max_pairs = num_sms / 2; }
            const char *g = getenv("B200KNN_MAX_PAIRS");
            if (g && atoi(g) > 0) max_pairs = atoi(g);
            if (getenv("B200KNN_VERBOSE")) fprintf(stderr, "[b200knn] device %d: %d SMs, %d co-resident CTA pairs (occupancy query %d)\n", device, num_sms, max_pairs, nc);
        }
        return B200KNN_OK;
    }
    void prof_begin(int kind, double flops = 0.0) {
        stats.kernel_launches++;
        if (!profiling) return;
        Ev e;
        cudaEventCreate(&e.a);
        cudaEventCreate(&e.b);
        e.kind = kind;
        e.flops = flops;
        cudaEventRecord(e.a, stream);
        events.push_back(e);
    }
    void prof_end() {
        if (!profiling) return;
        cudaEventRecord(events.back().b, stream);
    }
    void drain_events() {
        for (auto &e : events) {
            cudaEventSynchronize(e.b);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e.a, e.b);
            switch (e.kind) {
                case K_CONVERT: stats.ms_convert += ms; break;
                case K_DISTANCE: stats.ms_distance += ms; stats.distance_launches++; stats.distance_flops += e.flops; break;
                case K_RERANK: stats.ms_rerank += ms; break;
                case K_WAIT: stats.ms_wait += ms; break;
                default: stats.ms_scan += ms; break;
            }
            cudaEventDestroy(e.a);
            cudaEventDestroy(e.b);
        }
        events.clear();
    }
    void clear_pool() {
        if (!ready) return;
        cudaSetDevice(device);
        cudaStreamSynchronize(stream);
        if (x_owned && x_owned != x_store.p) cudaFree(x_owned);
        x_owned = nullptr;
        x_raw = nullptr;
        n = 0;
        sched1.key_nq = -1;
        sched2.key_nq = -1;
        tmap_q2_ptr = nullptr;
        if (release_on_clear || final_release) {       // (grow-only buffers otherwise: the next add() reuses them)
            x_store.release();
            x_bf.release();
            xnorm_bf.release();
            x_err.release();
            x_lo.release(); x_tf.release(); xnorm_t.release(); x_err_t.release(); x_lonorm.release();
        }
        centered = false;
    }
    bool final_release = false;
    void destroy() {
        if (!ready) return;
        final_release = true;
        clear_pool();
        drain_events();
        q_bf.release(); qnorm_bf.release(); q_err.release(); q_bf2.release(); uncert_thr.release(); min_score.release(); coll_count.release(); coll_idx.release(); overflow_list.release(); sched_items.release(); sched_items2.release(); sched_slots.release(); stream_sync.release(); radius2.release(); colterm.release(); rowthr.release(); member.release(); cand_s.release(); cand_i.release(); uncert_list.release();
        scan_d2.release(); scan_key.release(); scan_idx.release();
        projector.release(); proj_stage.release(); proj_rows.release();
        q_stage.release(); q_stage2.release(); out_idx.release(); out_dist.release(); pad_idx.release(); pad_dist.release(); scalars.release();
        if (h_count) cudaFreeHost(h_count);
        if (own_stream) cudaStreamDestroy(own_stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        for (int i = 0; i < 2; i++) {
            if (ev_copied[i]) cudaEventDestroy(ev_copied[i]);
            if (ev_consumed[i]) cudaEventDestroy(ev_consumed[i]);
        }
        ready = false;
    }

    // ------------------------------------------------------------------ host -> device rows
    // rows x row_bytes, source pitch src_pitch, destination packed.  Pinned / registered sources are DMA'd directly.
    // Pageable sources (what NumPy hands over) go through the pinned ring: several host threads memcpy a 32 MB piece
    // into a ring slot, the slot is DMA'd asynchronously, and the next piece is being filled meanwhile.
    int upload_rows(void *dst, const char *src, int64_t rows, size_t row_bytes, size_t src_pitch, cudaStream_t st) {
        if (rows <= 0) return B200KNN_OK;
        cudaPointerAttributes attr;
        bool pinned = false;
        if (cudaPointerGetAttributes(&attr, src) == cudaSuccess) pinned = (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
        else cudaGetLastError();
        const size_t total = static_cast<size_t>(rows) * row_bytes;
        if (pinned || total <= (8u << 20)) {   // small pageable copies: the driver's own staging is faster than spawning threads
            if (src_pitch == row_bytes) CU_TRY(cudaMemcpyAsync(dst, src, total, cudaMemcpyHostToDevice, st));
            else CU_TRY(cudaMemcpy2DAsync(dst, row_bytes, src, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice, st));
            return B200KNN_OK;
        }
        PinnedRing &rg = g_rings[device & 63];
        std::lock_guard<std::mutex> lock(rg.mu);     // one upload at a time per device
        for (int i = 0; i < PinnedRing::RING; i++) {
            if (!rg.buf[i]) {
                CU_TRY(cudaMallocHost(reinterpret_cast<void **>(&rg.buf[i]), PinnedRing::BYTES));
                CU_TRY(cudaEventCreateWithFlags(&rg.done[i], cudaEventDisableTiming));
            }
        }
        const size_t RING_BYTES = PinnedRing::BYTES;
        const int64_t piece_rows = std::max<int64_t>(1, static_cast<int64_t>(RING_BYTES / row_bytes));
        if (row_bytes > RING_BYTES) {   // absurdly wide rows: let the driver stage them
            CU_TRY(cudaMemcpy2DAsync(dst, row_bytes, src, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice, st));
            return B200KNN_OK;
        }
        for (int64_t r0 = 0; r0 < rows; r0 += piece_rows) {
            const int64_t pr = std::min(piece_rows, rows - r0);
            const int slot = rg.next;
            rg.next = (rg.next + 1) % PinnedRing::RING;
            if (rg.used[slot]) CU_TRY(cudaEventSynchronize(rg.done[slot]));   // its previous DMA has drained
            unsigned char *buf = rg.buf[slot];
            const char *sp = src + static_cast<size_t>(r0) * src_pitch;
            const int nt = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(copy_threads, pr), static_cast<int64_t>(pr * row_bytes) >> 21)));
            auto work = [&](int t) {
                const int64_t a = pr * t / nt, b = pr * (t + 1) / nt;
                if (src_pitch == row_bytes) {
                    std::memcpy(buf + static_cast<size_t>(a) * row_bytes, sp + static_cast<size_t>(a) * src_pitch, static_cast<size_t>(b - a) * row_bytes);
                } else {
                    for (int64_t r = a; r < b; r++) std::memcpy(buf + static_cast<size_t>(r) * row_bytes, sp + static_cast<size_t>(r) * src_pitch, row_bytes);
                }
            };
            if (nt == 1) {
                work(0);
            } else {
                std::vector<std::thread> th;
                th.reserve(nt - 1);
                for (int t = 1; t < nt; t++) th.emplace_back(work, t);
                work(0);
                for (auto &x : th) x.join();
            }
            CU_TRY(cudaMemcpyAsync(static_cast<char *>(dst) + static_cast<size_t>(r0) * row_bytes, buf, static_cast<size_t>(pr) * row_bytes,
                                   cudaMemcpyHostToDevice, st));
            CU_TRY(cudaEventRecord(rg.done[slot], st));
            rg.used[slot] = true;
        }
        return B200KNN_OK;
    }

    // ------------------------------------------------------------------ pool mean (centering)
    int compute_mean(const void *d_rows, int dtype, int64_t rows, int64_t ld, int dim) {
        centered = false;
        if (!use_centering || rows <= 0) return B200KNN_OK;
        TRY(col_mean.ensure(dim));
        CU_TRY(cudaMemsetAsync(col_mean.p, 0, static_cast<size_t>(dim) * sizeof(double), stream));
        dim3 grid(static_cast<unsigned>((rows + 255) / 256), static_cast<unsigned>((dim + 255) / 256));
        prof_begin(K_CONVERT);
        if (dtype == B200KNN_F64) colsum_kernel<double><<<grid, 256, 0, stream>>>(static_cast<const double *>(d_rows), rows, ld, dim, col_mean.p);
        else colsum_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float *>(d_rows), rows, ld, dim, col_mean.p);
        prof_end();
        prof_begin(K_CONVERT);
        scale_kernel<<<(dim + 255) / 256, 256, 0, stream>>>(col_mean.p, dim, 1.0 / static_cast<double>(rows));
        prof_end();
        CU_TRY(cudaGetLastError());
        centered = true;
        return B200KNN_OK;
    }

    // ------------------------------------------------------------------ kernels: convert
    // on_stream: launch there instead of the compute stream (the row-sharded protocol converts a rank's slice of the
    // query rows on its upload stream, ahead of the compute of earlier chunks); not profiled then
    int launch_convert(const void *src, int dtype, int64_t rows, int64_t ld, int dim, int kp, __nv_bfloat16 *dst, float *nbf,
                       float *nex, unsigned int *maxbits /* [2] */, cudaStream_t on_stream = nullptr) {
        const double *mu = centered ? col_mean.p : nullptr;
        if (rows <= 0) return B200KNN_OK;
        const size_t esz = dtype == B200KNN_F64 ? 8 : 4;
        int vec = (dim % 8 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0) && ((ld * esz) % 16 == 0);
        if (vec && reinterpret_cast<uintptr_t>(src) % 32 == 0 && (ld * esz) % 32 == 0) vec = 2;      // 256-bit loads
        // one row per warp and no grid-stride cap below ~half a million rows: the hardware block scheduler balances the
        // tail (a capped grid gave some warps 4 rows and others 3 at 30 000 query rows: 20 % of the launch idle)
        const int warps_per_block = 8;
        int64_t blocks = (rows + warps_per_block - 1) / warps_per_block;
        blocks = std::min<int64_t>(blocks, 65535);
        cudaStream_t st = on_stream ? on_stream : stream;
        if (on_stream) stats.kernel_launches++;
        else prof_begin(K_CONVERT);
        if (dtype == B200KNN_F64)
            convert_norm_kernel<double><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
                static_cast<const double *>(src), mu, rows, ld, dim, kp, vec, dst, nbf, nex, maxbits, maxbits + 1);
        else
            convert_norm_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
                static_cast<const float *>(src), mu, rows, ld, dim, kp, vec, dst, nbf, nex, maxbits, maxbits + 1);
        if (!on_stream) prof_end();
        CU_TRY(cudaGetLastError());
        return B200KNN_OK;
    }

    // operands of the split / tf32 tiers (tiers.cuh); hi == nullptr: the BF16 tier has already written the hi rows
    int launch_convert_tier(int t, const void *src, int dtype, int64_t rows, int64_t ld, int dim, int kp, __nv_bfloat16 *hi, __nv_bfloat16 *lo,
                            float *tf, float *norm, float *err, float *lonorm, unsigned int *maxbits /* [3] */) {
        const double *mu = centered ? col_mean.p : nullptr;
        if (rows <= 0) return B200KNN_OK;
        const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((rows + 7) / 8, static_cast<int64_t>(num_sms) * 8));
        prof_begin(K_CONVERT);
        if (t == 1) {
            if (dtype == B200KNN_F64) convert_tier_kernel<double, 1><<<blocks, 256, 0, stream>>>(static_cast<const double *>(src), mu, rows, ld, dim, kp, hi, lo, nullptr, norm, err, lonorm, maxbits);
            else convert_tier_kernel<float, 1><<<blocks, 256, 0, stream>>>(static_cast<const float *>(src), mu, rows, ld, dim, kp, hi, lo, nullptr, norm, err, lonorm, maxbits);
        } else {
            if (dtype == B200KNN_F64) convert_tier_kernel<double, 2><<<blocks, 256, 0, stream>>>(static_cast<const double *>(src), mu, rows, ld, dim, kp, nullptr, nullptr, tf, norm, err, nullptr, maxbits);
            else convert_tier_kernel<float, 2><<<blocks, 256, 0, stream>>>(static_cast<const float *>(src), mu, rows, ld, dim, kp, nullptr, nullptr, tf, norm, err, nullptr, maxbits);
        }
        prof_end();
        CU_TRY(cudaGetLastError());
        return B200KNN_OK;
    }
    // pool operands of the handle's tier; called after the BF16 conversion of add() (which every tier keeps: the BF16
    // rows are the hi part of the split tier, and ball membership always runs in BF16)
    int convert_pool_tier(int dim, int kp) {
        if (tier == 0 || n <= 0) return B200KNN_OK;
        TRY(xnorm_t.ensure(n));
        TRY(x_err_t.ensure(n));
        CU_TRY(cudaMemsetAsync(scalars.p + 10, 0, 3 * sizeof(unsigned int), stream));
        if (tier == 1) {
            TRY(x_lo.ensure(static_cast<size_t>(n) * kp));
            TRY(x_lonorm.ensure(n));
            TRY(launch_convert_tier(1, x_raw, x_dtype, n, ld_x, dim, kp, nullptr, x_lo.p, nullptr, xnorm_t.p, x_err_t.p, x_lonorm.p, scalars.p + 10));
            TRY(make_tmap(&tmap_xlo, x_lo.p, n, kp, BN));
            TRY(make_tmap(&tmap_xlo128, x_lo.p, n, kp, BN / 2));
        } else {
            TRY(x_tf.ensure(static_cast<size_t>(n) * kp));
            TRY(launch_convert_tier(2, x_raw, x_dtype, n, ld_x, dim, kp, nullptr, nullptr, x_tf.p, xnorm_t.p, x_err_t.p, nullptr, scalars.p + 10));
            TRY(make_tmap(&tmap_xtf, x_tf.p, n, kp, BN, true));
            TRY(make_tmap(&tmap_xtf128, x_tf.p, n, kp, BN / 2, true));
        }
        return B200KNN_OK;
    }

    // ------------------------------------------------------------------ pool
    int attach_pool(const void *d_rows, bool owned, int dtype, int64_t rows, int64_t ld, int dim, int kp, int64_t base) {
        x_raw = d_rows;
        x_owned = owned ? const_cast<void *>(d_rows) : nullptr;
        x_dtype = dtype;
        n = rows;
        ld_x = ld;
        index_base = base;
        TRY(x_bf.ensure(static_cast<size_t>(rows) * kp));
        TRY(xnorm_bf.ensure(rows));
        TRY(x_err.ensure(rows));
        CU_TRY(cudaMemsetAsync(scalars.p, 0, 2 * sizeof(unsigned int), stream));
        TRY(make_tmap(&tmap_x, x_bf.p, rows, kp, BN));
        TRY(make_tmap(&tmap_x128, x_bf.p, rows, kp, BN / 2));
        return B200KNN_OK;
    }

    // ------------------------------------------------------------------ schedule
    // ------------------------------------------------------------------ schedule
    // One round per group of query tiles: the group's `gs` tiles x `rc` chunks of the pool are processed side by side
    // (gs * rc <= workers), every worker sweeping NT/rc pool tiles (+-1).  The group's BF16 query rows (gs tiles) stay
    // in L2 for the whole round while `rc` pool-tile streams pass through once, each shared by `gs` workers that run
    // in lockstep: HBM traffic per round is ~ one pass over the pool instead of one per worker.
    struct Sched {
        int cg, qt, nt, workers, nrounds, max_slots, grid, qg;
        bool wide = false;                  // rounds run in round-wide lockstep (long K)
        int64_t key_nq = -1, key_n = -1;
        int key_kp = -1, key_slots = -1;
        bool matches(int64_t nq_, int64_t n_, int kp_, int slots_) const { return key_nq == nq_ && key_n == n_ && key_kp == kp_ && key_slots == slots_; }
        std::vector<WorkItem> items;        // [nrounds][workers]
        std::vector<int> slots_per_qtile;   // [qt]
    };
    Sched sched1;   // cached first-pass schedule
    int plan(Sched &s, int64_t nq, int kp, int max_slots_allowed) const {
        return plan_schedule(s, n, nq, kp, max_slots_allowed, forced_cg, max_pairs, num_sms, a_budget_mb, wide_mode);
    }
    // pure host function (no device needed): also reachable through b200knn_debug_plan for the CPU tests
    static int plan_schedule(Sched &s, int64_t n, int64_t nq, int kp, int max_slots_allowed, int forced_cg, int max_pairs, int num_sms,
                             int a_budget_mb, int wide_mode) {
        s.key_nq = nq;
        s.key_n = n;
        s.key_kp = kp;
        s.key_slots = max_slots_allowed;
        s.cg = forced_cg ? forced_cg : (nq > BM ? 2 : 1);
        s.workers = std::max(1, s.cg == 2 ? max_pairs : num_sms);
        const int W = s.workers;
        const int qrows = BM * s.cg;
        s.qt = static_cast<int>((nq + qrows - 1) / qrows);
        s.nt = static_cast<int>((n + BN - 1) / BN);
        // group size: as many query tiles as the L2 budget for the A operand allows, preferring sizes that tile the
        // worker count exactly
        const int64_t a_tile_bytes = static_cast<int64_t>(qrows) * kp * 2;
        const int g_cap = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(W, (static_cast<int64_t>(a_budget_mb) << 20) / std::max<int64_t>(a_tile_bytes, 1))));
        // Cost of a group size g in tile-times per worker: full rounds sweep nt / rc tiles each, the ragged last round has
        // its own rc.  Within 10 % of the cheapest, the LARGEST group wins: fewer rounds, and fewer pool streams per query
        // tile means fewer shortlists to publish, sort and bound per query (a 9-tile chunk split 2 x 37 instead of 9 x 8
        // cost 3x its tensor time in shortlist handling, measured in the multi-GPU chunk pipeline).
        int qg = 1;
        {
            auto rc_of = [&](int g) { return std::max(1, std::min(std::min(W / g, s.nt), max_slots_allowed)); };
            auto cost_of = [&](int g) {
                double c = static_cast<double>(s.qt / g) * s.nt / rc_of(g);
                if (s.qt % g) c += static_cast<double>(s.nt) / rc_of(s.qt % g);
                return c;
            };
            double best = 1e300;
            for (int g = 1; g <= std::min(g_cap, s.qt); g++) best = std::min(best, cost_of(g));
            for (int g = 1; g <= std::min(g_cap, s.qt); g++)
                if (cost_of(g) <= best * 1.10 + 1e-9) qg = g;
        }
        // Long K: only a few query tiles fit in L2 and each pool tile would be re-streamed from HBM for every small
        // group.  Alternative: a gs x rc grid of (query tile, pool stream) workers that advance through K together
        // (round-wide lockstep at every tile): a K block of a query tile is fetched once for its rc users and a K block
        // of a pool tile once for its gs users, nothing has to stay resident.  HBM rows fetched per output tile:
        // resident BN / gs, grid (gs * qrows + rc * BN) / (gs * rc); the cheaper one wins.
        int wide_g = 0;
        if (wide_mode != 0) {
            const int rc_res = std::max(1, std::min(std::min(W / qg, s.nt), max_slots_allowed));
            const double res_cost = static_cast<double>(BN) / qg / std::min(1.0, static_cast<double>(qg) * rc_res / W);
            double best_cost = 1e30;
            int bg = 0;
            for (int g = 2; g <= std::min(s.qt, 255); g++) {
                const int rc = std::min(std::min(W / g, s.nt), max_slots_allowed);
                if (rc < 2 || g * rc > 255) continue;
                const double util = static_cast<double>(g) * rc / W;
                const double cost = (static_cast<double>(g) * qrows + static_cast<double>(rc) * BN) / (static_cast<double>(g) * rc) / util;
                if (cost < best_cost) { best_cost = cost; bg = g; }
            }
            if (bg && (wide_mode == 2 || best_cost < 0.85 * res_cost)) wide_g = bg;
            if (wide_g) qg = wide_g;
        }
        s.wide = wide_g != 0;
        s.qg = qg;
        s.items.clear();
        s.slots_per_qtile.assign(s.qt, 0);
        s.nrounds = 0;
        s.max_slots = 1;
        for (int q0 = 0; q0 < s.qt; q0 += qg) {
            const int gs = std::min(qg, s.qt - q0);
            int rc = std::max(1, std::min(std::min(W / gs, s.nt), max_slots_allowed));
            if (s.wide && gs * rc > 255) rc = 255 / gs;
            const int wide_bits = (s.wide && gs > 1 && rc > 1) ? static_cast<int>(static_cast<unsigned int>(gs * rc) << 24) : 0;
            s.max_slots = std::max(s.max_slots, rc);
            s.items.resize(static_cast<size_t>(s.nrounds + 1) * W, WorkItem{-1, 0, 0, 0});
            WorkItem *row = s.items.data() + static_cast<size_t>(s.nrounds) * W;
            // chunk-major: workers sharing a pool-tile stream are neighbours
            for (int c = 0; c < rc; c++) {
                const int t0 = static_cast<int>(static_cast<int64_t>(c) * s.nt / rc);
                const int t1 = static_cast<int>(static_cast<int64_t>(c + 1) * s.nt / rc);
                for (int g = 0; g < gs; g++) row[c * gs + g] = WorkItem{q0 + g, t0, t1, c | (gs << 16) | wide_bits};
            }
            for (int g = 0; g < gs; g++) s.slots_per_qtile[q0 + g] = rc;
            s.nrounds++;
        }
        s.grid = s.cg * W;
        return B200KNN_OK;
    }
    // upload the schedule (small: a few KB) and reset the round barrier / lockstep counters
    int upload_schedule(const Sched &s, DevBuf<WorkItem> &items_dst, bool with_slots, bool cached) {
        if (!cached) {
            TRY(items_dst.ensure(s.items.size()));
            // (copies from pageable host memory are staged before cudaMemcpyAsync returns: no synchronisation needed)
            CU_TRY(cudaMemcpyAsync(items_dst.p, s.items.data(), s.items.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, stream));
            if (with_slots) {
                TRY(sched_slots.ensure(s.slots_per_qtile.size()));
                CU_TRY(cudaMemcpyAsync(sched_slots.p, s.slots_per_qtile.data(), s.slots_per_qtile.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
            }
        }
        CU_TRY(cudaMemsetAsync(scalars.p + 6, 0, sizeof(unsigned int), stream));
        TRY(stream_sync.ensure(static_cast<size_t>(s.nrounds) * s.max_slots));
        CU_TRY(cudaMemsetAsync(stream_sync.p, 0, static_cast<size_t>(s.nrounds) * s.max_slots * sizeof(unsigned int), stream));
        return B200KNN_OK;
    }

    // tier 0: (q, x, q, x); tier 1: (q_hi, x_hi, q_lo, x_lo); tier 2: the kind::tf32 flavour on (q_tf, x_tf)
    template <int C, bool COLLECT, bool TF32>
    int launch_dist_t(const Sched &s, const CUtensorMap &mq, const CUtensorMap &mx, const CUtensorMap &mx128, const CUtensorMap &mqlo,
                      const CUtensorMap &mxlo, const CUtensorMap &mxlo128, const DistParams &dp) {
        if (s.cg == 2) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(s.grid);
            cfg.blockDim = dim3(DIST_THREADS);
            cfg.dynamicSmemBytes = DistCfg<2>::SMEM_BYTES;
            cfg.stream = stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            CU_TRY(cudaLaunchKernelEx(&cfg, dist_topc_kernel<C, COLLECT, 2, TF32>, mq, mx128, mqlo, mxlo128, dp));
        } else {
            dist_topc_kernel<C, COLLECT, 1, TF32><<<s.grid, DIST_THREADS, DistCfg<1>::SMEM_BYTES, stream>>>(mq, mx, mqlo, mxlo, dp);
        }
        CU_TRY(cudaGetLastError());
        return B200KNN_OK;
    }
    // t: tier of this launch; tmap_q / tmap_qlo: the query operands of that tier (lo only for tier 1)
    template <int C, bool COLLECT>
    int launch_dist(const Sched &s, const CUtensorMap &tmap_q, const DistParams &dp, int t = 0, const CUtensorMap *tmap_qlo = nullptr) {
        if (t == 1) return launch_dist_t<C, COLLECT, false>(s, tmap_q, tmap_x, tmap_x128, *tmap_qlo, tmap_xlo, tmap_xlo128, dp);
        if (t == 2) return launch_dist_t<C, COLLECT, true>(s, tmap_q, tmap_xtf, tmap_xtf128, tmap_q, tmap_xtf, tmap_xtf128, dp);
        return launch_dist_t<C, COLLECT, false>(s, tmap_q, tmap_x, tmap_x128, tmap_q, tmap_x, tmap_x128, dp);
    }

    // ------------------------------------------------------------------ exact scan of a query subset
    // query s of the subset reads row d_inlist[s] of d_query (nullptr: row s) and writes row d_outlist[s] of the outputs
    // (nullptr: row s)
    template <typename TX, typename TQ>
    int scan_typed(const TQ *d_query, int64_t ld_q, const int *d_inlist, const int *d_outlist, int nsub, int dim, int kk, unsigned flags,
                   int32_t *d_out_idx, double *d_out_dist) {
        const TX *x = static_cast<const TX *>(x_raw);
        // sub-batches bounded to ~1.5 GB of scratch
        int64_t per_q = n * 8 + (kk > 32 ? static_cast<int64_t>(kk) * 24 : 0);
        int batch = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(nsub, (1536ll << 20) / std::max<int64_t>(per_q, 1))));
        batch = std::max(1, std::min(batch, 4096));
        TRY(scan_d2.ensure(static_cast<size_t>(batch) * n));
        for (int s0 = 0; s0 < nsub; s0 += batch) {
            const int ns = std::min(batch, nsub - s0);
            const TQ *qbase = d_inlist ? d_query : d_query + static_cast<int64_t>(s0) * ld_q;
            const int *qin = d_inlist ? d_inlist + s0 : nullptr;
            const int *ql = d_outlist ? d_outlist + s0 : nullptr;
            dim3 grid(static_cast<unsigned>((n + SCAN_TX - 1) / SCAN_TX), static_cast<unsigned>((ns + SCAN_TQ - 1) / SCAN_TQ));
            prof_begin(K_SCAN);
            scan_dist_kernel<TX, TQ><<<grid, 256, 0, stream>>>(x, ld_x, static_cast<int>(n), qbase, ld_q, qin, ns, dim, scan_d2.p);
            prof_end();
            CU_TRY(cudaGetLastError());
            int32_t *oi = d_outlist ? d_out_idx : d_out_idx + static_cast<int64_t>(s0) * kk;
            double *od = d_outlist ? d_out_dist : d_out_dist + static_cast<int64_t>(s0) * kk;
            if (kk <= 32) {
                prof_begin(K_SCAN);
                scan_select_kernel<<<ns, 256, 0, stream>>>(scan_d2.p, static_cast<int>(n), ql, kk, index_base, flags, oi, od);
                prof_end();
                CU_TRY(cudaGetLastError());
            } else {
                // kk > 32 / all points: radix select + stable radix sort of the kk selected, one block per scanned query
                TRY(scan_key.ensure(static_cast<size_t>(ns) * 2 * kk));
                TRY(scan_idx.ensure(static_cast<size_t>(ns) * 2 * kk));
                prof_begin(K_SCAN);
                scan_topk_kernel<<<ns, TOPK_THREADS, 0, stream>>>(scan_d2.p, static_cast<int>(n), ql, kk, index_base, flags, scan_key.p, scan_idx.p, oi, od);
                prof_end();
                CU_TRY(cudaGetLastError());
            }
        }
        return B200KNN_OK;
    }
    int scan(const void *d_query, int q_dtype, int64_t ld_q, const int *d_inlist, const int *d_outlist, int nsub, int dim, int kk, unsigned flags,
             int32_t *d_out_idx, double *d_out_dist) {
        if (x_dtype == B200KNN_F64 && q_dtype == B200KNN_F64)
            return scan_typed<double, double>(static_cast<const double *>(d_query), ld_q, d_inlist, d_outlist, nsub, dim, kk, flags, d_out_idx, d_out_dist);
        if (x_dtype == B200KNN_F64 && q_dtype == B200KNN_F32)
            return scan_typed<double, float>(static_cast<const float *>(d_query), ld_q, d_inlist, d_outlist, nsub, dim, kk, flags, d_out_idx, d_out_dist);
        if (x_dtype == B200KNN_F32 && q_dtype == B200KNN_F64)
            return scan_typed<float, double>(static_cast<const double *>(d_query), ld_q, d_inlist, d_outlist, nsub, dim, kk, flags, d_out_idx, d_out_dist);
        return scan_typed<float, float>(static_cast<const float *>(d_query), ld_q, d_inlist, d_outlist, nsub, dim, kk, flags, d_out_idx, d_out_dist);
    }

    // ------------------------------------------------------------------ rerank dispatch
    template <int C, int NT>
    void launch_rerank_nt(const void *d_query, int q_dtype, unsigned g, size_t sm, const RerankParams &rp) {
        if (x_dtype == B200KNN_F64 && q_dtype == B200KNN_F64)
            rerank_kernel<double, double, C, NT><<<g, NT, sm, stream>>>(static_cast<const double *>(x_raw), static_cast<const double *>(d_query), rp);
        else if (x_dtype == B200KNN_F64 && q_dtype == B200KNN_F32)
            rerank_kernel<double, float, C, NT><<<g, NT, sm, stream>>>(static_cast<const double *>(x_raw), static_cast<const float *>(d_query), rp);
        else if (x_dtype == B200KNN_F32 && q_dtype == B200KNN_F64)
            rerank_kernel<float, double, C, NT><<<g, NT, sm, stream>>>(static_cast<const float *>(x_raw), static_cast<const double *>(d_query), rp);
        else
            rerank_kernel<float, float, C, NT><<<g, NT, sm, stream>>>(static_cast<const float *>(x_raw), static_cast<const float *>(d_query), rp);
    }
    // throughput flavour: one warp per query, RR_WPB queries per block (rerank.cuh)
    static constexpr int RR_WPB = 8;
    template <int C>
    void launch_rerank_warp(const void *d_query, int q_dtype, int64_t nq, int pk, const RerankParams &rp) {
        // persistent: as many blocks as can be resident (3 per SM at 80 registers), every warp draws queries from scalars[16]
        const unsigned g = static_cast<unsigned>(std::min<int64_t>((nq + RR_WPB - 1) / RR_WPB, static_cast<int64_t>(num_sms) * 3));
        const size_t sm = static_cast<size_t>(RR_WPB) * pk * sizeof(unsigned long long);
        const int inq = static_cast<int>(nq);
        unsigned int *ctr = scalars.p + 16;
        if (x_dtype == B200KNN_F64 && q_dtype == B200KNN_F64)
            rerank_warp_kernel<double, double, C, RR_WPB><<<g, RR_WPB * 32, sm, stream>>>(static_cast<const double *>(x_raw), static_cast<const double *>(d_query), rp, inq, pk, ctr);
        else if (x_dtype == B200KNN_F64 && q_dtype == B200KNN_F32)
            rerank_warp_kernel<double, float, C, RR_WPB><<<g, RR_WPB * 32, sm, stream>>>(static_cast<const double *>(x_raw), static_cast<const float *>(d_query), rp, inq, pk, ctr);
        else if (x_dtype == B200KNN_F32 && q_dtype == B200KNN_F64)
            rerank_warp_kernel<float, double, C, RR_WPB><<<g, RR_WPB * 32, sm, stream>>>(static_cast<const float *>(x_raw), static_cast<const double *>(d_query), rp, inq, pk, ctr);
        else
            rerank_warp_kernel<float, float, C, RR_WPB><<<g, RR_WPB * 32, sm, stream>>>(static_cast<const float *>(x_raw), static_cast<const float *>(d_query), rp, inq, pk, ctr);
    }
    template <int C>
    int launch_rerank(const void *d_query, int q_dtype, int64_t nq, const RerankParams &rp) {
        prof_begin(K_RERANK);
        const unsigned g = static_cast<unsigned>(nq);
        int pk = 1;
        while (pk < rp.max_slots * C) pk <<= 1;
        const size_t sm = static_cast<size_t>(pk) * sizeof(unsigned long long);
        // Three flavours, bit-identical results.  Many queries: one warp per query (24 resident per SM, their latency
        // chains overlap).  Few queries: a block per query (128 lanes stream a candidate row at once: lowest latency) —
        // and when those few queries merge many shortlists (many pool streams), a latency-bound sort, 32 warps per query.
        // (long rows — config 5's 196 KB — keep the block flavour: its 128-lane sweep was tuned and measured there, the warp flavour was not)
        const bool warp_flavour = rerank_warp_mode == 2 || (rerank_warp_mode == 1 && nq >= 2048 && rp.dim <= 8192);
        if (warp_flavour && pk <= 512) launch_rerank_warp<C>(d_query, q_dtype, nq, pk, rp);
        else if (pk >= 1024 && nq <= 4096) launch_rerank_nt<C, 1024>(d_query, q_dtype, g, sm, rp);
        else launch_rerank_nt<C, 128>(d_query, q_dtype, g, sm, rp);
        prof_end();
        CU_TRY(cudaGetLastError());
        return B200KNN_OK;
    }

    // accumulation geometry of a tier: K blocks per segment, total, per unit (chunked accumulation only pays for long rows:
    // at least two full units), and the (k_unit, n_units) the error model charges
    struct AccGeom { int nkb_seg, num_kb, kb_per_unit, k_unit, n_units; };
    AccGeom acc_geom(int kp, int t) const {
        AccGeom g{};
        const int kelems = t == 2 ? BK / 2 : BK;
        g.nkb_seg = (kp + kelems - 1) / kelems;
        g.num_kb = t == 1 ? 3 * g.nkb_seg : g.nkb_seg;
        const int kbu = kc_elems > 0 ? std::max(1, kc_elems / kelems) : 0;
        g.kb_per_unit = (kbu > 0 && g.num_kb >= 2 * kbu) ? kbu : g.num_kb;
        g.n_units = (g.num_kb + g.kb_per_unit - 1) / g.kb_per_unit;
        g.k_unit = std::min(g.kb_per_unit, g.num_kb) * kelems;
        return g;
    }
    DistParams base_dist_params(int64_t nq, int kp, const Sched &s, const WorkItem *items, int t = 0) const {
        DistParams dp{};
        const AccGeom g = acc_geom(kp, t);
        dp.xnorm = t ? xnorm_t.p : xnorm_bf.p;
        dp.n = static_cast<int>(n);
        dp.nq = static_cast<int>(nq);
        dp.num_kb = g.num_kb;
        dp.nkb_seg = g.nkb_seg;
        dp.kb_per_unit = g.kb_per_unit;
        dp.items = items;
        dp.nrounds = s.nrounds;
        dp.workers = s.workers;
        dp.round_counter = scalars.p + 6;
        dp.stream_sync = stream_sync.p;
        // sharers drift by time, not by tiles: about one re-alignment per 16 tiles of K = 3072
        dp.sync_tiles = sync_tiles >= 0 ? sync_tiles : (s.wide ? 1 : std::max(1, 16 * 3072 / std::max(kp, 1)));
        dp.sync_timeout_ns = 40000u * static_cast<unsigned>(std::max(1, kp / 3072));     // ~2 tile times
        dp.max_slots = s.max_slots;
        dp.opt = opt_flags;
        return dp;
    }

    // ------------------------------------------------------------------ call bracket: overflow bookkeeping
    // A call (one C-ABI query) may span several passes (query chunks).  Queries whose second-pass list overflowed are
    // appended to overflow_list as call-global rows; the host looks at the counter ONCE, after the last pass.
    int begin_call(int64_t total_nq) {
        CU_TRY(cudaSetDevice(device));
        TRY(overflow_list.ensure(static_cast<size_t>(std::max<int64_t>(total_nq, 1))));
        CU_TRY(cudaMemsetAsync(scalars.p + 5, 0, sizeof(unsigned int), stream));
        return B200KNN_OK;
    }
    // enqueue the read-back of the overflow counter (the caller synchronises the stream, then reads *h_count)
    int enqueue_overflow_readback() {
        CU_TRY(cudaMemcpyAsync(h_count, scalars.p + 5, sizeof(int), cudaMemcpyDeviceToHost, stream));
        return B200KNN_OK;
    }
    // exact scan of the overflowed queries; d_query_all holds every query row of the call on the device
    int fix_overflow_device(const void *d_query_all, int q_dtype, int64_t ld_q, int nov, int dim, int kk, unsigned flags,
                            int32_t *d_out_idx_all, double *d_out_dist_all) {
        if (nov <= 0) return B200KNN_OK;
        stats.exact_scanned += nov;
        return scan(d_query_all, q_dtype, ld_q, overflow_list.p, overflow_list.p, nov, dim, kk, flags, d_out_idx_all, d_out_dist_all);
    }
    // the same when the query rows live on the HOST (chunks already recycled on the device): re-upload just those rows
    int fix_overflow_host(const void *h_query_all, int q_dtype, int64_t ld_q, int nov, int dim, int kk, unsigned flags,
                          int32_t *d_out_idx_all, double *d_out_dist_all) {
        if (nov <= 0) return B200KNN_OK;
        stats.exact_scanned += nov;
        const size_t esz = q_dtype == B200KNN_F64 ? 8 : 4;
        std::vector<int> rows(nov);
        CU_TRY(cudaMemcpyAsync(rows.data(), overflow_list.p, static_cast<size_t>(nov) * sizeof(int), cudaMemcpyDeviceToHost, stream));
        CU_TRY(cudaStreamSynchronize(stream));
        std::vector<unsigned char> packed(static_cast<size_t>(nov) * dim * esz);
        for (int i = 0; i < nov; i++)
            std::memcpy(packed.data() + static_cast<size_t>(i) * dim * esz,
                        static_cast<const char *>(h_query_all) + static_cast<size_t>(rows[i]) * ld_q * esz, static_cast<size_t>(dim) * esz);
        TRY(q_stage.ensure(packed.size()));
        CU_TRY(cudaMemcpyAsync(q_stage.p, packed.data(), packed.size(), cudaMemcpyHostToDevice, stream));
        CU_TRY(cudaStreamSynchronize(stream));          // `packed` is a stack-lifetime host buffer
        return scan(q_stage.p, q_dtype, dim, nullptr, overflow_list.p, nov, dim, kk, flags, d_out_idx_all, d_out_dist_all);
    }

    // hook of the row-sharded query protocol (one rank per GPU, b200knn_exchange_query*): after the tensor pass every
    // rank publishes an upper bound on its kk-th nearest distance per query; the re-rank prunes against their minimum
    struct ShardHook {
        PeerPtrs peers;
        int world = 1, rank = 0;
        size_t bounds_off = 0, bound_flag_off = 0;
        int64_t bounds_stride = 0;
        unsigned int bound_step = 0;
        const char *local_base = nullptr;
        int kk_global = 0;              // neighbours of the GLOBAL answer (a shard with fewer rows publishes +inf)
        QueryPull qpull{};              // host-row queries: the original rows are read from their owners' buffers (peer loads)
    };
    // a single-pass call may postpone the second pass until the host has seen the uncertified count at the call's one
    // synchronisation (the common small call has none: nothing is launched for it)
    struct Deferred {
        bool armed = false;
        const void *d_query = nullptr;
        int q_dtype = 0, dim = 0, kp = 0, kk = 0, q_offset = 0;
        int64_t ld_q = 0, nq = 0;
        unsigned flags = 0;
        int32_t *out_idx = nullptr;
        double *out_dist = nullptr;
    } deferred;
    bool defer_second_pass = false;     // set by the caller around query_device()
    // A call whose chunks all stay in HBM (b200knn_query on one device: the query rows of the whole call fit a device
    // buffer) runs ONE second pass at its end instead of one per chunk: every collection pass sweeps the whole BF16 shard
    // (0.26-0.44 ms at config 3) whether it serves 10 uncertified queries or 10 000.  The re-rank of every chunk appends to
    // the same queue (call-global row numbers); the converted rows of all chunks are kept (whole-call buffers).
    struct CallAccum {
        bool on = false, first = true;
        int64_t rows = 0;               // query rows of the call
        int64_t off = 0;                // first row of the chunk being enqueued
        const QuerySide *pre_base = nullptr;   // converted rows handed in by the caller (self-kNN): row 0 of the call
    } accum;
    // the call's one second pass; d_query_all / outputs: whole-call arrays
    int finish_accumulated_call(const void *d_query_all, int q_dtype, int64_t ld_q, int dim, int kp, int kk, unsigned flags,
                                int32_t *d_out_idx_all, double *d_out_dist_all) {
        const int64_t rows = accum.rows;
        accum = CallAccum{};
        if (kk > 32 || (flags & (B200KNN_FLAG_FORCE_SCAN | B200KNN_FLAG_NO_CERTIFY)) || rows <= 0) return B200KNN_OK;
        return enqueue_second_pass(d_query_all, q_dtype, ld_q, rows, dim, kp, kk, flags, d_out_idx_all, d_out_dist_all, 0, false);
    }
    int enqueue_uncertified_readback() {     // h_count[1] <- uncertified count of the last pass
        CU_TRY(cudaMemcpyAsync(h_count + 1, scalars.p + 4, sizeof(int), cudaMemcpyDeviceToHost, stream));
        return B200KNN_OK;
    }
    int run_deferred_second_pass() {
        if (!deferred.armed) return B200KNN_OK;
        deferred.armed = false;
        return enqueue_second_pass(deferred.d_query, deferred.q_dtype, deferred.ld_q, deferred.nq, deferred.dim, deferred.kp, deferred.kk,
                                   deferred.flags, deferred.out_idx, deferred.out_dist, deferred.q_offset, false);
    }

    // Second pass for the queries the certificate rejected: the same tcgen05 GEMM, but the epilogue collects every pool
    // row whose score is within the query's error margin; those short lists are re-ranked exactly.  Everything is
    // enqueued without knowing how many such queries there are (the count lives in scalars[4]; the schedule is planned
    // on the device): no host synchronisation, and with a count of zero every kernel returns at once.
    // Lists that overflow go to the exact CUDA-core scan at the end of the call (fix_overflow_*).
    Sched sched2;   // cached worst-case plan of the second pass (group size, kernel flavour)
    CUtensorMap tmap_q2;
    const void *tmap_q2_ptr = nullptr;
    int64_t tmap_q2_rows = -1;
    int tmap_q2_kp = -1;
    int enqueue_second_pass(const void *d_query, int q_dtype, int64_t ld_q, int64_t nq_cap, int dim, int kp, int kk, unsigned flags,
                            int32_t *d_out_idx, double *d_out_dist, int q_offset, bool allow_short, const QueryPull *qpull = nullptr) {
        const int t = cur_tier;                // the collection pass runs in the tier of the pass that found the query uncertified
        const int kp_plan = t ? 2 * kp : kp;   // operand bytes per row, in BF16-element units (L2 budget of the schedule)
        if (t != 2) TRY(q_bf2.ensure(static_cast<size_t>(nq_cap) * kp));
        if (t == 1) TRY(q_lo2.ensure(static_cast<size_t>(nq_cap) * kp));
        if (t == 2) TRY(q_tf2.ensure(static_cast<size_t>(nq_cap) * kp));
        TRY(coll_count.ensure(nq_cap));
        TRY(coll_idx.ensure(static_cast<size_t>(nq_cap) * COLLECT_CAP));
        if (!sched2.matches(nq_cap, n, kp_plan, 1 << 20)) TRY(plan(sched2, nq_cap, kp_plan, 1 << 20));
        const Sched &s = sched2;
        const int max_rounds = (s.qt + s.qg - 1) / s.qg;
        TRY(sched_items2.ensure(static_cast<size_t>(max_rounds) * s.workers));
        CUtensorMap tmap_q2lo;
        if (t == 0) {
            if (tmap_q2_ptr != q_bf2.p || tmap_q2_rows != nq_cap || tmap_q2_kp != kp) {
                TRY(make_tmap(&tmap_q2, q_bf2.p, nq_cap, kp, BM));
                tmap_q2_ptr = q_bf2.p; tmap_q2_rows = nq_cap; tmap_q2_kp = kp;
            }
        } else {
            tmap_q2_ptr = nullptr;
            if (t == 1) {
                TRY(make_tmap(&tmap_q2, q_bf2.p, nq_cap, kp, BM));
                TRY(make_tmap(&tmap_q2lo, q_lo2.p, nq_cap, kp, BM));
            } else {
                TRY(make_tmap(&tmap_q2, q_tf2.p, nq_cap, kp, BM, true));
            }
        }
        const int *count_dev = reinterpret_cast<const int *>(scalars.p + 4);
        TRY(stream_sync.ensure(static_cast<size_t>(max_rounds) * s.workers));
        PlanParams pp{};
        pp.count_dev = count_dev;
        pp.qrows = BM * s.cg; pp.nt = s.nt; pp.workers = s.workers; pp.group = s.qg; pp.max_rounds = max_rounds; pp.wide = s.wide ? 1 : 0;
        pp.max_slots_cap = 1 << 20;
        pp.items = sched_items2.p;
        pp.round_counter = scalars.p + 6;
        pp.stream_sync = stream_sync.p;
        pp.sync_entries = max_rounds * s.workers;
        pp.zero_list = coll_count.p;
        pp.zero_list_n = static_cast<int>(nq_cap);
        pp.total_uncertified = scalars.p + 8;
        prof_begin(K_SCAN);
        plan_pass_kernel<<<1, 256, 0, stream>>>(pp);
        prof_end();
        prof_begin(K_SCAN);
        if (t != 2)
            gather_rows_kernel<<<num_sms * 4, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(cur_q_bf), uncert_list.p, count_dev, kp / 8,
                                                                reinterpret_cast<uint4 *>(q_bf2.p));
        if (t == 1)
            gather_rows_kernel<<<num_sms * 4, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(cur_q_lo), uncert_list.p, count_dev, kp / 8,
                                                                reinterpret_cast<uint4 *>(q_lo2.p));
        if (t == 2)
            gather_rows_kernel<<<num_sms * 4, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(cur_q_tf), uncert_list.p, count_dev, kp / 4,
                                                                reinterpret_cast<uint4 *>(q_tf2.p));
        prof_end();
        CU_TRY(cudaGetLastError());
        const AccGeom ag = acc_geom(kp, t);
        DistParams dp{};
        dp.xnorm = t ? xnorm_t.p : xnorm_bf.p;
        dp.n = static_cast<int>(n);
        dp.nq = static_cast<int>(nq_cap);
        dp.nq_dev = count_dev;
        dp.num_kb = ag.num_kb;
        dp.nkb_seg = ag.nkb_seg;
        dp.kb_per_unit = ag.kb_per_unit;
        dp.items = sched_items2.p;
        dp.nrounds = max_rounds;
        dp.workers = s.workers;
        dp.round_counter = scalars.p + 6;
        dp.stream_sync = stream_sync.p;
        dp.sync_tiles = sync_tiles >= 0 ? sync_tiles : (s.wide ? 1 : std::max(1, 16 * 3072 / std::max(kp, 1)));
        dp.sync_timeout_ns = 40000u * static_cast<unsigned>(std::max(1, kp / 3072));
        dp.max_slots = s.workers;
        dp.opt = opt_flags;
        dp.thr = uncert_thr.p;
        dp.coll_count = coll_count.p;
        dp.coll_idx = coll_idx.p;
        dp.coll_cap = COLLECT_CAP;
        prof_begin(K_SCAN);
        const int rc2 = launch_dist<16, true>(s, tmap_q2, dp, t, &tmap_q2lo);
        prof_end();
        TRY(rc2);
        CollectRerankParams cp{};
        cp.count_dev = count_dev;
        cp.uncert_list = uncert_list.p;
        cp.coll_count = coll_count.p;
        cp.coll_idx = coll_idx.p;
        cp.dim = dim;
        cp.ld_x = ld_x;
        cp.ld_q = ld_q;
        cp.kk = kk;
        cp.index_base = index_base;
        cp.flags = flags;
        cp.allow_short = allow_short ? 1 : 0;
        cp.q_offset = q_offset;
        if (qpull) cp.qpull = *qpull;
        cp.out_idx = d_out_idx;
        cp.out_dist = d_out_dist;
        cp.overflow_count = reinterpret_cast<int *>(scalars.p + 5);
        cp.overflow_list = overflow_list.p;
        prof_begin(K_SCAN);
        const unsigned g = static_cast<unsigned>(std::min<int64_t>(nq_cap, static_cast<int64_t>(num_sms) * 2));
        if (x_dtype == B200KNN_F64 && q_dtype == B200KNN_F64)
            rerank_collect_kernel<double, double, 32><<<g, 1024, 0, stream>>>(static_cast<const double *>(x_raw), static_cast<const double *>(d_query), cp);
        else if (x_dtype == B200KNN_F64 && q_dtype == B200KNN_F32)
            rerank_collect_kernel<double, float, 32><<<g, 1024, 0, stream>>>(static_cast<const double *>(x_raw), static_cast<const float *>(d_query), cp);
        else if (x_dtype == B200KNN_F32 && q_dtype == B200KNN_F64)
            rerank_collect_kernel<float, double, 32><<<g, 1024, 0, stream>>>(static_cast<const float *>(x_raw), static_cast<const double *>(d_query), cp);
        else
            rerank_collect_kernel<float, float, 32><<<g, 1024, 0, stream>>>(static_cast<const float *>(x_raw), static_cast<const float *>(d_query), cp);
        prof_end();
        CU_TRY(cudaGetLastError());
        return B200KNN_OK;
    }

    template <int C>
    int tensor_pass(const void *d_query, int q_dtype, int64_t nq, int64_t ld_q, int dim, int kp, int kk, unsigned flags,
                    int32_t *d_out_idx, double *d_out_dist, const QuerySide *pre, int q_offset, const ShardHook *hook) {
        // query side: BF16 rows + norms, either converted now or already there (self-kNN: the pool's own, converted by
        // add(); row-sharded protocol: broadcast by the ranks that converted them)
        // The pass runs in the handle's tier when the query operands of that tier are at hand (converted here, or the
        // pool's own for self-kNN); BF16 slices broadcast by the row-sharded protocol run in BF16.
        const int t = (tier != 0 && (!pre || (tier == 1 ? pre->lo != nullptr : pre->tf != nullptr))) ? tier : 0;
        cur_tier = t;
        const int kp_plan = t ? 2 * kp : kp;
        const __nv_bfloat16 *qb = nullptr, *qlo = nullptr;
        const float *qtf = nullptr, *qn, *qe, *qln = nullptr;
        // (a call that postpones its second pass keeps the converted rows of ALL its chunks: whole-call buffers, this
        // chunk's rows at row `off`)
        const bool acc = accum.on && !hook && (!pre || accum.pre_base);
        const int64_t cap = acc ? accum.rows : nq, off = acc ? accum.off : 0;
        if (pre) {
            qb = pre->bf; qn = pre->norm; qe = pre->err; qlo = pre->lo; qtf = pre->tf; qln = pre->lonorm;
        } else if (t == 0) {
            TRY(q_bf.ensure(static_cast<size_t>(cap) * kp));
            TRY(qnorm_bf.ensure(cap));
            TRY(q_err.ensure(cap));
            qb = q_bf.p + static_cast<size_t>(off) * kp; qn = qnorm_bf.p + off; qe = q_err.p + off;
            TRY(launch_convert(d_query, q_dtype, nq, ld_q, dim, kp, q_bf.p + static_cast<size_t>(off) * kp, qnorm_bf.p + off, q_err.p + off, scalars.p + 2));
        } else {
            TRY(qnorm_t.ensure(cap));
            TRY(q_err_t.ensure(cap));
            if (t == 1) {
                TRY(q_bf.ensure(static_cast<size_t>(cap) * kp));
                TRY(q_lo.ensure(static_cast<size_t>(cap) * kp));
                TRY(q_lonorm.ensure(cap));
                TRY(launch_convert_tier(1, d_query, q_dtype, nq, ld_q, dim, kp, q_bf.p + static_cast<size_t>(off) * kp, q_lo.p + static_cast<size_t>(off) * kp, nullptr,
                                        qnorm_t.p + off, q_err_t.p + off, q_lonorm.p + off, scalars.p + 13));
                qb = q_bf.p + static_cast<size_t>(off) * kp; qlo = q_lo.p + static_cast<size_t>(off) * kp; qln = q_lonorm.p + off;
            } else {
                TRY(q_tf.ensure(static_cast<size_t>(cap) * kp));
                TRY(launch_convert_tier(2, d_query, q_dtype, nq, ld_q, dim, kp, nullptr, nullptr, q_tf.p + static_cast<size_t>(off) * kp, qnorm_t.p + off, q_err_t.p + off,
                                        nullptr, scalars.p + 13));
                qtf = q_tf.p + static_cast<size_t>(off) * kp;
            }
            qn = qnorm_t.p + off; qe = q_err_t.p + off;
        }
        // what the second pass gathers from: rows are addressed by the numbers the re-rank queues (chunk-local, or
        // call-global when the pass is postponed to the end of the call)
        if (acc && pre) {
            cur_q_bf = accum.pre_base->bf; cur_q_lo = accum.pre_base->lo; cur_q_tf = accum.pre_base->tf;
        } else {
            cur_q_bf = acc ? (qb ? q_bf.p : nullptr) : qb;
            cur_q_lo = acc ? (qlo ? q_lo.p : nullptr) : qlo;
            cur_q_tf = acc ? (qtf ? q_tf.p : nullptr) : qtf;
        }
        CUtensorMap tmap_q, tmap_qlo;
        if (t == 2) TRY(make_tmap(&tmap_q, qtf, nq, kp, BM, true));
        else TRY(make_tmap(&tmap_q, qb, nq, kp, BM));
        if (t == 1) TRY(make_tmap(&tmap_qlo, qlo, nq, kp, BM));
        if (!sched1.matches(nq, n, kp_plan, MAX_KEYS / C)) TRY(plan(sched1, nq, kp_plan, MAX_KEYS / C));
        const Sched &s = sched1;
        {   // the schedule is written on the device (and every counter of the pass zeroed) by one tiny kernel: no
            // host-to-device copy, no memset — nothing of a pass can queue behind a query upload on a copy engine
            TRY(sched_items.ensure(static_cast<size_t>(s.nrounds) * s.workers));
            TRY(sched_slots.ensure(s.qt));
            TRY(stream_sync.ensure(static_cast<size_t>(s.nrounds) * s.max_slots));
            PlanParams pp{};
            pp.count = static_cast<int>(nq);
            pp.qrows = BM * s.cg; pp.nt = s.nt; pp.workers = s.workers; pp.group = s.qg; pp.max_rounds = s.nrounds; pp.wide = s.wide ? 1 : 0;
            pp.max_slots_cap = MAX_KEYS / C;
            pp.items = sched_items.p;
            pp.slots_per_qtile = sched_slots.p;
            pp.round_counter = scalars.p + 6;
            pp.stream_sync = stream_sync.p;
            pp.sync_entries = s.nrounds * s.max_slots;
            pp.zero_a = (acc && !accum.first) ? nullptr : scalars.p + 4;      // (postponed second pass: the queue grows over the chunks)
            pp.zero_b = scalars.p + 16;
            stats.kernel_launches++;
            plan_pass_kernel<<<1, 256, 0, stream>>>(pp);
            CU_TRY(cudaGetLastError());
        }
        TRY(cand_s.ensure(static_cast<size_t>(nq) * s.max_slots * C));
        TRY(cand_i.ensure(static_cast<size_t>(nq) * s.max_slots * C));
        DistParams dp = base_dist_params(nq, kp, s, sched_items.p, t);
        dp.cand_s = cand_s.p;
        dp.cand_i = cand_i.p;
        prof_begin(K_DISTANCE, 2.0 * static_cast<double>(nq) * static_cast<double>(n) * dim);
        const int rc1 = launch_dist<C, false>(s, tmap_q, dp, t, &tmap_qlo);
        prof_end();
        TRY(rc1);

        last_nq = nq;
        last_slots = s.max_slots;
        last_c = C;
        TRY(uncert_list.ensure(cap));
        TRY(uncert_thr.ensure(cap));
        RerankParams rp{};
        rp.uncert_q_base = static_cast<int>(off);
        rp.cand_s = cand_s.p;
        rp.cand_i = cand_i.p;
        rp.max_slots = s.max_slots;
        rp.slots_per_qtile = sched_slots.p;
        rp.qtile_rows = BM * s.cg;
        rp.dim = dim;
        rp.ld_x = ld_x;
        rp.ld_q = ld_q;
        rp.n = static_cast<int>(n);
        rp.kk = kk;
        rp.index_base = index_base;
        rp.flags = flags;
        rp.qnorm_bf = qn;
        rp.q_err = qe;
        rp.max_xnorm_bf_bits = t ? scalars.p + 10 : scalars.p;
        rp.max_x_err_bits = t ? scalars.p + 11 : scalars.p + 1;
        rp.kp = kp;
        {
            const AccGeom ag = acc_geom(kp, t);
            rp.tier = t;
            rp.k_unit = ag.k_unit;
            rp.n_units = ag.n_units;
            rp.q_lonorm = qln;
            rp.max_x_lonorm_bits = scalars.p + 12;
        }
        rp.out_idx = d_out_idx;
        rp.out_dist = d_out_dist;
        rp.uncert_count = reinterpret_cast<int *>(scalars.p + 4);
        rp.uncert_list = uncert_list.p;
        rp.uncert_thr = uncert_thr.p;
        if (hook) {
            // every rank tells every rank how close its kk-th candidate is at most (peer stores + step flag) ...
            BoundParams bp{};
            bp.cand_s = cand_s.p;
            bp.cand_i = cand_i.p;
            bp.max_slots = s.max_slots;
            bp.c = C;
            bp.slots_per_qtile = sched_slots.p;
            bp.qtile_rows = BM * s.cg;
            bp.nq = static_cast<int>(nq);
            bp.kk = hook->kk_global > 0 ? hook->kk_global : kk;
            bp.em = rp;
            bp.peers = hook->peers;
            bp.world = hook->world;
            bp.rank = hook->rank;
            bp.bounds_off = hook->bounds_off;
            bp.flag_off = hook->bound_flag_off;
            bp.stride = hook->bounds_stride;
            bp.step = hook->bound_step;
            bp.done_counter = scalars.p + 9;
            TRY(min_score.ensure(nq));
            bp.min_score = min_score.p;
            rp.min_score = min_score.p;
            prof_begin(K_RERANK);
            bound_publish_kernel<<<static_cast<unsigned>(std::min<int64_t>(num_sms * 4, (nq + 7) / 8)), 256, 0, stream>>>(bp);
            prof_end();
            CU_TRY(cudaGetLastError());
            // ... and the re-rank prunes against the minimum (it waits for the peers' flags itself)
            rp.ext_bounds = reinterpret_cast<const float *>(hook->local_base + hook->bounds_off) +
                            static_cast<int64_t>(hook->bound_step & 1u) * hook->world * hook->bounds_stride;
            rp.ext_world = hook->world;
            rp.ext_stride = hook->bounds_stride;
            rp.qpull = hook->qpull;
            // the peers' bounds must have arrived: one tiny spinning block, so that the skew between ranks is not billed
            // to the re-rank
            prof_begin(K_WAIT);
            wait_flags_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<const unsigned int *>(hook->local_base + hook->bound_flag_off), hook->world,
                                                    hook->bound_step, 101);
            prof_end();
            CU_TRY(cudaGetLastError());
        }
        TRY(launch_rerank<C>(d_query, q_dtype, nq, rp));
        deferred.armed = false;
        if (acc) {                   // the caller enqueues ONE second pass for the whole call (finish_accumulated_call)
            accum.first = false;
            return B200KNN_OK;
        }
        if (!(flags & B200KNN_FLAG_NO_CERTIFY)) {
            if (defer_second_pass && !hook) {
                deferred.armed = true;
                deferred.d_query = d_query; deferred.q_dtype = q_dtype; deferred.ld_q = ld_q; deferred.nq = nq; deferred.dim = dim;
                deferred.kp = kp; deferred.kk = kk; deferred.flags = flags; deferred.out_idx = d_out_idx; deferred.out_dist = d_out_dist;
                deferred.q_offset = q_offset;
            } else {
                TRY(enqueue_second_pass(d_query, q_dtype, ld_q, nq, dim, kp, kk, flags, d_out_idx, d_out_dist, q_offset, hook != nullptr,
                                        hook ? &hook->qpull : nullptr));
            }
        }
        return B200KNN_OK;
    }

    // ------------------------------------------------------------------ ball membership (precision/recall metric)
    // d_out[i] |= 1 when query i lies inside any ball B(x_j, sqrt(radius2[j])) of this shard.  radius2 is already in
    // `radius2` (device).  Tensor-core filter in collect mode + exact float64 decision; overflowed lists without a
    // witness go to the exact scan.
    static constexpr int MEMBER_CAP = 256;
    int ball_membership(const void *d_query, int q_dtype, int64_t nq, int64_t ld_q, int dim, int kp, unsigned char *d_out) {
        if (nq <= 0 || n <= 0) return B200KNN_OK;
        CU_TRY(cudaSetDevice(device));
        stats.queries += nq;
        TRY(q_bf.ensure(static_cast<size_t>(nq) * kp));
        TRY(qnorm_bf.ensure(nq));
        TRY(q_err.ensure(nq));
        TRY(launch_convert(d_query, q_dtype, nq, ld_q, dim, kp, q_bf.p, qnorm_bf.p, q_err.p, scalars.p + 2));
        TRY(colterm.ensure(n));
        TRY(rowthr.ensure(nq));
        CU_TRY(cudaMemsetAsync(scalars.p + 7, 0, sizeof(unsigned int), stream));
        prof_begin(K_SCAN);
        ball_colterm_kernel<<<std::min<int64_t>(num_sms * 4, (n + 255) / 256), 256, 0, stream>>>(xnorm_bf.p, x_err.p, radius2.p, static_cast<int>(n),
                                                                                          colterm.p, scalars.p + 7);
        prof_end();
        prof_begin(K_SCAN);
        ball_rowthr_kernel<<<static_cast<unsigned>((nq + 255) / 256), 256, 0, stream>>>(qnorm_bf.p, q_err.p, scalars.p, scalars.p + 7, kp,
                                                                                        static_cast<int>(nq), rowthr.p);
        prof_end();
        CU_TRY(cudaGetLastError());
        TRY(coll_count.ensure(nq));
        TRY(coll_idx.ensure(static_cast<size_t>(nq) * MEMBER_CAP));
        TRY(overflow_list.ensure(nq));
        CU_TRY(cudaMemsetAsync(coll_count.p, 0, static_cast<size_t>(nq) * sizeof(int), stream));
        CU_TRY(cudaMemsetAsync(scalars.p + 5, 0, sizeof(unsigned int), stream));
        CUtensorMap tmap_q;
        TRY(make_tmap(&tmap_q, q_bf.p, nq, kp, BM));
        Sched s;
        TRY(plan(s, nq, kp, 1 << 20));
        TRY(upload_schedule(s, sched_items2, false, false));
        DistParams dp = base_dist_params(nq, kp, s, sched_items2.p);
        dp.xnorm = colterm.p;
        dp.thr = rowthr.p;
        dp.coll_count = coll_count.p;
        dp.coll_idx = coll_idx.p;
        dp.coll_cap = MEMBER_CAP;
        prof_begin(K_DISTANCE, 2.0 * static_cast<double>(nq) * static_cast<double>(n) * dim);
        const int rc = launch_dist<16, true>(s, tmap_q, dp);
        prof_end();
        TRY(rc);
        MemberParams mp{};
        mp.coll_count = coll_count.p;
        mp.coll_idx = coll_idx.p;
        mp.cap = MEMBER_CAP;
        mp.radius2 = radius2.p;
        mp.dim = dim;
        mp.ld_x = ld_x;
        mp.ld_q = ld_q;
        mp.out_member = d_out;
        mp.overflow_count = reinterpret_cast<int *>(scalars.p + 5);
        mp.overflow_list = overflow_list.p;
        prof_begin(K_RERANK);
        const unsigned g = static_cast<unsigned>(nq);
        if (x_dtype == B200KNN_F64 && q_dtype == B200KNN_F64)
            ball_member_kernel<double, double><<<g, 128, 0, stream>>>(static_cast<const double *>(x_raw), static_cast<const double *>(d_query), mp);
        else if (x_dtype == B200KNN_F64 && q_dtype == B200KNN_F32)
            ball_member_kernel<double, float><<<g, 128, 0, stream>>>(static_cast<const double *>(x_raw), static_cast<const float *>(d_query), mp);
        else if (x_dtype == B200KNN_F32 && q_dtype == B200KNN_F64)
            ball_member_kernel<float, double><<<g, 128, 0, stream>>>(static_cast<const float *>(x_raw), static_cast<const double *>(d_query), mp);
        else
            ball_member_kernel<float, float><<<g, 128, 0, stream>>>(static_cast<const float *>(x_raw), static_cast<const float *>(d_query), mp);
        prof_end();
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(h_count, scalars.p + 5, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CU_TRY(cudaStreamSynchronize(stream));
        const int nov = *h_count;
        if (nov > 0) {
            stats.exact_scanned += nov;
            TRY(scan_members(d_query, q_dtype, ld_q, nov, dim, d_out));
        }
        return B200KNN_OK;
    }
    template <typename TX, typename TQ>
    int scan_members_typed(const TQ *d_query, int64_t ld_q, int nsub, int dim, unsigned char *d_out) {
        const int batch = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min(nsub, 4096), (1536ll << 20) / std::max<int64_t>(n * 8, 1))));
        TRY(scan_d2.ensure(static_cast<size_t>(batch) * n));
        for (int s0 = 0; s0 < nsub; s0 += batch) {
            const int ns = std::min(batch, nsub - s0);
            dim3 grid(static_cast<unsigned>((n + SCAN_TX - 1) / SCAN_TX), static_cast<unsigned>((ns + SCAN_TQ - 1) / SCAN_TQ));
            prof_begin(K_SCAN);
            scan_dist_kernel<TX, TQ><<<grid, 256, 0, stream>>>(static_cast<const TX *>(x_raw), ld_x, static_cast<int>(n), d_query, ld_q,
                                                               overflow_list.p + s0, ns, dim, scan_d2.p);
            prof_end();
            prof_begin(K_SCAN);
            scan_member_kernel<<<ns, 256, 0, stream>>>(scan_d2.p, static_cast<int>(n), overflow_list.p + s0, radius2.p, d_out);
            prof_end();
            CU_TRY(cudaGetLastError());
        }
        return B200KNN_OK;
    }
    int scan_members(const void *d_query, int q_dtype, int64_t ld_q, int nsub, int dim, unsigned char *d_out) {
        if (x_dtype == B200KNN_F64 && q_dtype == B200KNN_F64) return scan_members_typed<double, double>(static_cast<const double *>(d_query), ld_q, nsub, dim, d_out);
        if (x_dtype == B200KNN_F64 && q_dtype == B200KNN_F32) return scan_members_typed<double, float>(static_cast<const float *>(d_query), ld_q, nsub, dim, d_out);
        if (x_dtype == B200KNN_F32 && q_dtype == B200KNN_F64) return scan_members_typed<float, double>(static_cast<const double *>(d_query), ld_q, nsub, dim, d_out);
        return scan_members_typed<float, float>(static_cast<const float *>(d_query), ld_q, nsub, dim, d_out);
    }

    // Allocate every per-pass buffer a tensor pass over `nq` rows will touch (same formulas as tensor_pass /
    // enqueue_second_pass).  Growing a buffer frees the old one, and cudaFree synchronises the context: inside the
    // multi-GPU protocol that must not happen while flag-waiting kernels are in flight, so the protocol reserves for all
    // its chunk sizes before it enqueues anything.
    int reserve_pass(int64_t nq, int kp, int kk, bool pre) {
        if (nq <= 0 || kk > 32) return B200KNN_OK;
        const int C = kk <= 4 ? 16 : (kk <= 16 ? 32 : 64);
        const int t = pre ? 0 : tier;
        const int kp_plan = t ? 2 * kp : kp;
        if (!pre) {
            if (t != 2) TRY(q_bf.ensure(static_cast<size_t>(nq) * kp));
            TRY(qnorm_bf.ensure(nq));
            TRY(q_err.ensure(nq));
            if (t) { TRY(qnorm_t.ensure(nq)); TRY(q_err_t.ensure(nq)); }
            if (t == 1) { TRY(q_lo.ensure(static_cast<size_t>(nq) * kp)); TRY(q_lonorm.ensure(nq)); TRY(q_lo2.ensure(static_cast<size_t>(nq) * kp)); }
            if (t == 2) { TRY(q_tf.ensure(static_cast<size_t>(nq) * kp)); TRY(q_tf2.ensure(static_cast<size_t>(nq) * kp)); }
        }
        Sched a;
        TRY(plan(a, nq, kp_plan, MAX_KEYS / C));
        TRY(sched_items.ensure(a.items.size()));
        TRY(sched_slots.ensure(a.slots_per_qtile.size()));
        TRY(stream_sync.ensure(static_cast<size_t>(a.nrounds) * a.max_slots));
        TRY(cand_s.ensure(static_cast<size_t>(nq) * a.max_slots * C));
        TRY(cand_i.ensure(static_cast<size_t>(nq) * a.max_slots * C));
        TRY(uncert_list.ensure(nq));
        TRY(uncert_thr.ensure(nq));
        TRY(min_score.ensure(nq));
        TRY(q_bf2.ensure(static_cast<size_t>(nq) * kp));
        TRY(coll_count.ensure(nq));
        TRY(coll_idx.ensure(static_cast<size_t>(nq) * COLLECT_CAP));
        Sched b;
        TRY(plan(b, nq, kp_plan, 1 << 20));
        const int max_rounds = (b.qt + b.qg - 1) / b.qg;
        TRY(sched_items2.ensure(static_cast<size_t>(max_rounds) * b.workers));
        TRY(stream_sync.ensure(static_cast<size_t>(max_rounds) * b.workers));
        return B200KNN_OK;
    }
    // the pinned ring of upload_rows (pageable sources): allocated before the protocol starts, for the same reason
    int reserve_upload_ring() {
        PinnedRing &rg = g_rings[device & 63];
        std::lock_guard<std::mutex> lock(rg.mu);
        for (int i = 0; i < PinnedRing::RING; i++) {
            if (!rg.buf[i]) {
                CU_TRY(cudaMallocHost(reinterpret_cast<void **>(&rg.buf[i]), PinnedRing::BYTES));
                CU_TRY(cudaEventCreateWithFlags(&rg.done[i], cudaEventDisableTiming));
            }
        }
        return B200KNN_OK;
    }

    // a whole call in one pass, device-resident queries: bracket + pass + overflow fix-up; synchronises the stream once
    int query_device_sync(const void *d_query, int q_dtype, int64_t nq, int64_t ld_q, int dim, int kp, int k, unsigned flags,
                          int32_t *d_out_idx, double *d_out_dist, const QuerySide *pre = nullptr) {
        if (nq <= 0) return B200KNN_OK;
        const int kk = static_cast<int>(std::min<int64_t>(k, n));
        TRY(begin_call(nq));
        TRY(query_device(d_query, q_dtype, nq, ld_q, dim, kp, k, flags, d_out_idx, d_out_dist, pre, 0, nullptr));
        TRY(enqueue_overflow_readback());
        CU_TRY(cudaStreamSynchronize(stream));
        return fix_overflow_device(d_query, q_dtype, ld_q, *h_count, dim, kk, flags, d_out_idx, d_out_dist);
    }

    // One pass: queries and outputs on this device; nq bounded by the caller's chunking.  Everything is ENQUEUED on the
    // stream; the caller brackets the passes of a call with begin_call() ... enqueue_overflow_readback() + synchronise +
    // fix_overflow_*().  q_offset: first row of this pass within the call.
    int query_device(const void *d_query, int q_dtype, int64_t nq, int64_t ld_q, int dim, int kp, int k, unsigned flags,
                     int32_t *d_out_idx, double *d_out_dist, const QuerySide *pre = nullptr, int q_offset = 0, const ShardHook *hook = nullptr) {
        if (nq <= 0) return B200KNN_OK;
        CU_TRY(cudaSetDevice(device));
        const int kk = static_cast<int>(std::min<int64_t>(k, n));
        stats.queries += nq;
        if (kk > 32 || (flags & B200KNN_FLAG_FORCE_SCAN))
            return scan(d_query, q_dtype, ld_q, nullptr, nullptr, static_cast<int>(nq), dim, kk, flags, d_out_idx, d_out_dist);
        if (kk <= 4) return tensor_pass<16>(d_query, q_dtype, nq, ld_q, dim, kp, kk, flags, d_out_idx, d_out_dist, pre, q_offset, hook);
        if (kk <= 16) return tensor_pass<32>(d_query, q_dtype, nq, ld_q, dim, kp, kk, flags, d_out_idx, d_out_dist, pre, q_offset, hook);
        return tensor_pass<64>(d_query, q_dtype, nq, ld_q, dim, kp, kk, flags, d_out_idx, d_out_dist, pre, q_offset, hook);
    }
};

constexpr int64_t QUERY_CHUNK = 32768;   // query rows per device pass (bounds workspace; 256 query tiles)
// largest call that runs ONE second pass over all its passes (Shard::CallAccum): the collection lists are sized for the
// worst case, every row uncertified (4 KB per row: 1 GB here), as are the whole-call copies of the converted rows
constexpr int64_t WHOLE_CALL_MAX_ROWS = 262144;

}  // namespace
