// libb200knn — host side and C ABI (include/b200knn.h) of the B200-native exact kNN engine.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC (see build.py).
// No CPU fallback: every compute entry point needs an sm_100 device and fails loudly without one.
#include "shard.cuh"

struct b200knn_index {
    int dim = 0;
    int kp = 0;
    std::vector<int> device_ids;
    std::vector<Shard> shards;
    int64_t n_total = 0;
    bool devices_ready = false;
    void *user_stream = nullptr;
    bool profiling = false;
    // multi-device gather buffers on shard 0
    DevBuf<int32_t> g_idx;
    DevBuf<double> g_dist;
    // random projection: every shard keeps the projector [proj_in_dim][dim], a staging buffer for unprojected rows and one
    // for projected queries (Shard::projector / proj_stage / proj_rows)
    int64_t proj_in_dim = 0;
    // multi-device handles: one exchange per non-empty shard, wired together in-process (b200knn_exchange_connect_local);
    // add() and query() then run the same collective protocol a torchrun job runs, one host thread per shard
    std::vector<b200knn_exchange *> exch;
    std::vector<int> exch_shards;       // shard index of rank i

    int ensure_devices() {
        if (devices_ready) return B200KNN_OK;
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count <= 0) {
            cudaGetLastError();
            return fail(B200KNN_ENODEVICE, "no CUDA device available (%s); libb200knn has no CPU fallback",
                        e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        }
        if (device_ids.empty()) {
            int cur = 0;
            cudaGetDevice(&cur);
            device_ids.push_back(cur);
        }
        for (int d : device_ids)
            if (d < 0 || d >= count) return fail(B200KNN_EINVAL, "device id %d out of range (%d devices)", d, count);
        shards.resize(device_ids.size());
        for (size_t i = 0; i < shards.size(); i++) {
            TRY(shards[i].init(device_ids[i]));
            shards[i].profiling = profiling;
            if (user_stream && shards.size() == 1) shards[i].stream = static_cast<cudaStream_t>(user_stream);
        }
        for (size_t i = 0; i < shards.size(); i++)
            for (size_t j = 0; j < shards.size(); j++)
                if (i != j) {
                    cudaSetDevice(device_ids[i]);
                    cudaDeviceEnablePeerAccess(device_ids[j], 0);
                    cudaGetLastError();
                }
        devices_ready = true;
        return B200KNN_OK;
    }
};

#include "exchange_host.cuh"

namespace {

void drop_group(b200knn_index *ix) {
    for (auto *e : ix->exch) ex_destroy(e);
    ix->exch.clear();
    ix->exch_shards.clear();
}

// one exchange per non-empty shard, capacity bounded to ~2 GiB of original query rows per device
int ensure_group(b200knn_index *ix, const std::vector<int> &active) {
    if (ix->exch_shards == active && ix->exch.size() == active.size()) return B200KNN_OK;
    drop_group(ix);
    const int W = static_cast<int>(active.size());
    int64_t max_nq = (int64_t(1) << 31) / (int64_t(2) * ix->dim * 8);
    max_nq = std::max<int64_t>(BM * 2, std::min<int64_t>(QUERY_CHUNK, max_nq / (BM * 2) * (BM * 2)));
    for (int i = 0; i < W; i++) {
        b200knn_exchange *e = nullptr;
        const int rc = ex_create(ix->shards[active[i]].device, i, W, ix->dim, max_nq, 32, &e);
        if (rc != B200KNN_OK) { drop_group(ix); return rc; }
        ix->exch.push_back(e);
    }
    ix->exch_shards = active;
    for (int i = 0; i < W; i++) {
        b200knn_exchange *e = ix->exch[i];
        if (cudaSetDevice(e->device) != cudaSuccess) { drop_group(ix); return fail(B200KNN_ECUDA, "cudaSetDevice failed"); }
        for (int p = 0; p < W; p++) {
            if (p == i) continue;
            if (ix->exch[p]->device != e->device) {
                cudaError_t ce = cudaDeviceEnablePeerAccess(ix->exch[p]->device, 0);
                if (ce != cudaSuccess && ce != cudaErrorPeerAccessAlreadyEnabled) {
                    drop_group(ix);
                    return fail(B200KNN_ECUDA, "no peer access between devices %d and %d: %s", e->device, ix->exch[p]->device, cudaGetErrorString(ce));
                }
                cudaGetLastError();
            }
            e->peer_base[p] = ix->exch[p]->base;
        }
        e->connected = true;
    }
    return B200KNN_OK;
}

// run fn(rank) on one host thread per rank; the first failure (if any) is reported
template <typename Fn>
int for_each_rank(int W, Fn fn) {
    std::vector<int> rcs(W, B200KNN_OK);
    std::vector<std::string> errs(W);
    std::vector<std::thread> th;
    for (int i = 1; i < W; i++)
        th.emplace_back([&, i]() {
            rcs[i] = fn(i);
            if (rcs[i] != B200KNN_OK) errs[i] = g_last_error;
        });
    rcs[0] = fn(0);
    if (rcs[0] != B200KNN_OK) errs[0] = g_last_error;
    for (auto &t : th) t.join();
    for (int i = 0; i < W; i++)
        if (rcs[i] != B200KNN_OK) return fail(rcs[i], "%s", errs[i].c_str());
    return B200KNN_OK;
}

// a failed add leaves no half-built pool behind (device memory released, handle reusable)
int cleanup_failed_add(b200knn_index *ix, int rc) {
    if (rc != B200KNN_OK && rc != B200KNN_ESTATE && ix && ix->n_total == 0) {
        const std::string msg = g_last_error;
        for (auto &s : ix->shards) s.clear_pool();
        cudaGetLastError();
        g_last_error = msg;
    }
    return rc;
}

// Row-shard n rows over the handle's devices and index them.  fill(shard, d_rows, r0, rows) brings rows [r0, r0 + rows)
// of the caller's matrix into the shard's own store (packed, dim elements per row), asynchronously on the shard's stream:
// an upload from host memory, a peer copy of device rows, or a projection of unprojected rows.  One host thread per
// shard (the PCIe links of the GPUs are independent); then, collectively over the non-empty shards, the global column
// means over peer memory and one convert pass.
template <typename Fill>
int add_sharded(b200knn_index *ix, int64_t n, int dtype, Fill fill) {
    const size_t esz = dtype == B200KNN_F64 ? 8 : 4;
    const int G = static_cast<int>(ix->shards.size());
    const int64_t per = (n + G - 1) / G;
    auto shard_add = [&](int g) -> int {
        Shard &s = ix->shards[g];
        const int64_t r0 = std::min<int64_t>(n, per * g), r1 = std::min<int64_t>(n, per * (g + 1));
        const int64_t rows = r1 - r0;
        if (rows <= 0) { s.n = 0; return B200KNN_OK; }
        CU_TRY(cudaSetDevice(s.device));
        TRY(s.x_store.ensure(static_cast<size_t>(rows) * ix->dim * esz));
        void *d_rows = s.x_store.p;
        int r = s.attach_pool(d_rows, true, dtype, rows, ix->dim, ix->dim, ix->kp, r0);
        if (r != B200KNN_OK) return r;
        TRY(fill(s, d_rows, r0, rows));
        if (G > 1) return B200KNN_OK;        // phase 2 below: the centring vector is the GLOBAL column mean
        TRY(s.compute_mean(d_rows, dtype, rows, ix->dim, ix->dim));
        TRY(s.launch_convert(d_rows, dtype, rows, ix->dim, ix->dim, ix->kp, s.x_bf.p, s.xnorm_bf.p, s.x_err.p, s.scalars.p));
        TRY(s.convert_pool_tier(ix->dim, ix->kp));
        return B200KNN_OK;
    };
    if (G == 1) {
        TRY(shard_add(0));
    } else {
        std::vector<int> rcs(G, B200KNN_OK);
        std::vector<std::string> errs(G);
        const int saved_threads = ix->shards[0].copy_threads;
        for (auto &s : ix->shards) s.copy_threads = std::max(1, saved_threads * 2 / G);
        std::vector<std::thread> th;
        for (int g = 0; g < G; g++)
            th.emplace_back([&, g]() {
                rcs[g] = shard_add(g);
                if (rcs[g] != B200KNN_OK) errs[g] = g_last_error;
            });
        for (auto &t : th) t.join();
        for (auto &s : ix->shards) s.copy_threads = saved_threads;
        for (int g = 0; g < G; g++)
            if (rcs[g] != B200KNN_OK) return fail(rcs[g], "%s", errs[g].c_str());
        // phase 2 (collective over the non-empty shards; every rank got this far): column sums gathered over peer memory ->
        // global means -> BF16 convert + norms
        std::vector<int> active;
        for (int g = 0; g < G; g++)
            if (ix->shards[g].n > 0) active.push_back(g);
        TRY(ensure_group(ix, active));
        TRY(for_each_rank(static_cast<int>(active.size()), [&](int i) -> int {
            Shard &s = ix->shards[active[i]];
            CU_TRY(cudaSetDevice(s.device));
            return ex_finish_add(ix->exch[i], s, ix->dim, ix->kp);
        }));
    }
    for (auto &s : ix->shards) {
        CU_TRY(cudaSetDevice(s.device));
        CU_TRY(cudaStreamSynchronize(s.stream));   // the caller may free / overwrite its rows after return
    }
    ix->n_total = n;
    return B200KNN_OK;
}

// Self-kNN on a multi-device handle: the rows of every shard, chunk by chunk, are the queries of one collective call
// (b200knn_exchange_query_device's protocol, one host thread per shard).  The chunk's ORIGINAL rows are replicated to the
// other shards over NVLink first (peer copies into their stage buffers: the protocol takes replicated device rows; every
// shard then converts 1/G of them and broadcasts the BF16 slice).  Results are bit-identical to a single-device handle.
int query_self_sharded(b200knn_index *ix, int k, unsigned flags, int32_t *out_idx, double *out_dist, int *out_kk) {
    std::vector<int> active;
    for (size_t g = 0; g < ix->shards.size(); g++)
        if (ix->shards[g].n > 0) active.push_back(static_cast<int>(g));
    const int W = static_cast<int>(active.size());
    if (ix->exch.empty() || ix->exch_shards != active) return fail(B200KNN_ESTATE, "the handle's device group is not set up (add() does that)");
    const int kk = static_cast<int>(std::min<int64_t>(k, ix->n_total));
    if (out_kk) *out_kk = kk;
    if (kk > ix->exch[0]->max_kk) return fail(B200KNN_EINVAL, "query_self on a multi-device handle serves k <= %d (single-device handles: any k)", ix->exch[0]->max_kk);
    const int dim = ix->dim;
    const int64_t step_rows = std::min<int64_t>(ix->exch[0]->max_nq, QUERY_CHUNK);
    Shard &s0 = ix->shards[active[0]];
    // Every allocation of every rank FIRST, for every chunk size that will occur: once a rank's flag-waiting kernels are in
    // flight no rank may allocate (with peer access enabled, cudaMalloc / cudaFree touch the peers' address spaces and can
    // wait for a kernel that itself waits for this rank — the rule of ex_query_host_prepare)
    {
        std::vector<int64_t> sizes;
        for (int r = 0; r < W; r++)
            for (int64_t q0 = 0; q0 < ix->shards[active[r]].n; q0 += step_rows) {
                const int64_t cq = std::min(step_rows, ix->shards[active[r]].n - q0);
                if (std::find(sizes.begin(), sizes.end(), cq) == sizes.end()) sizes.push_back(cq);
            }
        const int64_t max_cq = *std::max_element(sizes.begin(), sizes.end());
        for (int i = 0; i < W; i++) {
            Shard &s = ix->shards[active[i]];
            CU_TRY(cudaSetDevice(s.device));
            TRY(s.out_idx.ensure(static_cast<size_t>(max_cq) * kk));
            TRY(s.out_dist.ensure(static_cast<size_t>(max_cq) * kk));
            TRY(s.q_stage.ensure(static_cast<size_t>(max_cq) * dim * 8));
            ExCall scratch;
            TRY(ex_begin_query(ix->exch[i], s, max_cq, k, flags, scratch));
            for (int64_t cq : sizes) TRY(s.reserve_pass(cq, ix->kp, scratch.kk_l, W > 1 && s.tier == 0));
        }
    }
    for (int r = 0; r < W; r++) {
        Shard &src = ix->shards[active[r]];
        const size_t esz = src.x_dtype == B200KNN_F64 ? 8 : 4;
        for (int64_t q0 = 0; q0 < src.n; q0 += step_rows) {
            const int64_t cq = std::min(step_rows, src.n - q0);
            const char *rows_src = static_cast<const char *>(src.x_raw) + static_cast<size_t>(q0) * src.ld_x * esz;
            for (int i = 0; i < W; i++) {           // replicate the chunk's original rows (peer copies on the receivers' streams)
                Shard &s = ix->shards[active[i]];
                if (i == r) continue;
                CU_TRY(cudaSetDevice(s.device));
                CU_TRY(cudaMemcpy2DAsync(s.q_stage.p, static_cast<size_t>(dim) * esz, rows_src, static_cast<size_t>(src.ld_x) * esz, static_cast<size_t>(dim) * esz,
                                         static_cast<size_t>(cq), cudaMemcpyDefault, s.stream));
            }
            TRY(for_each_rank(W, [&](int i) -> int {
                Shard &s = ix->shards[active[i]];
                CU_TRY(cudaSetDevice(s.device));
                const void *dq = (i == r) ? static_cast<const void *>(rows_src) : static_cast<const void *>(s.q_stage.p);
                return ex_query_device(ix->exch[i], s, dim, ix->kp, dq, src.x_dtype, cq, (i == r) ? src.ld_x : dim, k, flags, s.out_idx.p, s.out_dist.p, nullptr);
            }));
            CU_TRY(cudaSetDevice(s0.device));        // every rank holds the merged lists; shard 0's copy goes to the caller
            const int64_t g0 = src.index_base + q0;
            CU_TRY(cudaMemcpyAsync(out_idx + g0 * kk, s0.out_idx.p, static_cast<size_t>(cq) * kk * sizeof(int32_t), cudaMemcpyDeviceToHost, s0.stream));
            CU_TRY(cudaMemcpyAsync(out_dist + g0 * kk, s0.out_dist.p, static_cast<size_t>(cq) * kk * sizeof(double), cudaMemcpyDeviceToHost, s0.stream));
            CU_TRY(cudaStreamSynchronize(s0.stream));
        }
    }
    return B200KNN_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Chunks of a single-device host-row call in whole-call mode (every chunk has its own slice of one device buffer, the
// uploads run back to back on the copy stream, the call has ONE second pass): which cut hides most of the upload?
// The first chunk's upload is the only one nothing can hide, so it should be small; every later chunk must have arrived
// before the chunk before it has been computed, so chunks may grow by about the ratio of the compute to the upload time
// per row (2.8 at config 3 from pinned memory) — a ramp.  Not every size is a good one: a chunk of h query tiles runs as
// rounds of the distance kernel's schedule (Shard::plan_schedule), and h x (pool streams) should fill the workers
// (8 x 9, 12 x 6, 24 x 3, 37 x 2 of 74).  So: keep the tail of full groups, cut the head (one group plus the ragged
// remainder) into up to four parts by exhaustive search over a timeline model — upload time per tile from the link speed,
// compute time from the ACTUAL schedule of that chunk size — and take the cut that finishes first.
// Pure host arithmetic (b200knn_debug_host_chunks exposes it to the CPU tests).
struct HostChunkModel {
    double upload_gbs = 55.0;        // pinned host memory over PCIe 5 x16 (measured 52-60 GB/s); pageable through the ring: ~40
    double tile_us = 20.8;           // one 256 x 256 x 3072 BF16 tile pair on the tcgen05 pipe, power-capped clocks (bench.py, config 3)
    double chunk_overhead_ms = 0.30; // convert + plan + re-rank launches, pipeline fill and the round barrier tail of a chunk
    int min_part_tiles = 6;          // no chunk below this many query tiles: with fewer, every query has dozens of pool streams, i.e.
                                     // shortlists to write, sort and bound (the first model allowed 1-tile chunks and lost 1 ms to them)
};

double sched_cost_tiles(const Shard::Sched &s) {     // pool tiles swept by the busiest worker, summed over the rounds
    double total = 0.0;
    for (int r = 0; r < s.nrounds; r++) {
        int mx = 0;
        for (int w = 0; w < s.workers; w++) {
            const WorkItem &it = s.items[static_cast<size_t>(r) * s.workers + w];
            if (it.qtile >= 0) mx = std::max(mx, it.t1 - it.t0);
        }
        total += mx;
    }
    return total;
}

int plan_host_chunks(int64_t n, int64_t nq, int kp_plan, int dim, size_t esz, int max_slots, int forced_cg, int max_pairs, int num_sms,
                     int a_budget_mb, int wide_mode, int64_t cap_rows, const HostChunkModel &m, int tier_mmas,
                     std::vector<std::pair<int64_t, int64_t>> &chunks) {
    chunks.clear();
    Shard::Sched full;
    TRY(Shard::plan_schedule(full, n, nq, kp_plan, max_slots, forced_cg, max_pairs, num_sms, a_budget_mb, wide_mode));
    const int64_t qrows = static_cast<int64_t>(BM) * full.cg;
    const int qt = static_cast<int>((nq + qrows - 1) / qrows);
    int G = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(full.qg, std::max<int64_t>(1, cap_rows / qrows))));
    if (qt < 4 || full.wide) {                       // tiny calls, long rows (grid schedule): the ragged remainder, then whole groups
        const int64_t group_rows = std::max<int64_t>(qrows, std::min<int64_t>(static_cast<int64_t>(full.qg) * qrows, cap_rows / qrows * qrows));
        if (nq <= group_rows + group_rows / 4) {
            chunks.emplace_back(0, nq);
        } else {
            const int64_t rem = nq % group_rows;
            int64_t q0 = 0;
            if (rem > 0) { chunks.emplace_back(0, rem); q0 = rem; }
            for (; q0 < nq; q0 += group_rows) chunks.emplace_back(q0, std::min(group_rows, nq - q0));
        }
        return B200KNN_OK;
    }
    const double up_ms_per_tile = static_cast<double>(qrows) * dim * static_cast<double>(esz) / (m.upload_gbs * 1e6);
    const double tile_ms = m.tile_us * 1e-3 * (static_cast<double>(kp_plan) / 3072.0) * tier_mmas;
    // compute time of a chunk of h tiles, from its real schedule
    const int k_tail = std::max(0, qt / G - 1);      // full groups kept at the end
    const int H = qt - k_tail * G;                   // head: G .. 2G-1 tiles (or the whole call when it is shorter than 2 groups)
    std::vector<double> comp(static_cast<size_t>(std::max(H, G)) + 1, 0.0);
    for (int h = 1; h <= std::max(H, G); h++) {
        Shard::Sched sc;
        TRY(Shard::plan_schedule(sc, n, static_cast<int64_t>(h) * qrows, kp_plan, max_slots, forced_cg, max_pairs, num_sms, a_budget_mb, wide_mode));
        comp[h] = sched_cost_tiles(sc) * tile_ms + m.chunk_overhead_ms;
    }
    auto finish_time = [&](const int *parts, int np) {
        double up = 0.0, done = 0.0;
        for (int i = 0; i < np; i++) {
            up += parts[i] * up_ms_per_tile;
            done = std::max(done, up) + comp[parts[i]];
        }
        for (int i = 0; i < k_tail; i++) {
            up += G * up_ms_per_tile;
            done = std::max(done, up) + comp[G];
        }
        return done;
    };
    int best[4] = {H, 0, 0, 0}, nbest = 1;
    double best_t = finish_time(best, 1);
    const int max_part = static_cast<int>(std::max<int64_t>(1, cap_rows / qrows));
    const int mp = std::min(m.min_part_tiles, H);
    for (int a = mp; a <= H; a++) {
        if (a > max_part) break;
        for (int b = 0; a + b <= H; b++) {
            if (b > max_part) break;
            if (b != 0 && b < mp) continue;
            if (b == 0) {
                if (a != H) continue;
                const int p[1] = {a};
                const double t = finish_time(p, 1);
                if (t < best_t - 1e-9) { best_t = t; nbest = 1; best[0] = a; }
                continue;
            }
            for (int c = 0; a + b + c <= H; c++) {
                if (c > max_part) break;
                if (c != 0 && c < mp) continue;
                const int d = H - a - b - c;
                if (d > max_part) continue;
                if (c == 0 && d != 0) continue;
                if (d != 0 && d < mp) continue;
                int p[4] = {a, b, c, d};
                const int np = c == 0 ? 2 : (d == 0 ? 3 : 4);
                const double t = finish_time(p, np);
                if (t < best_t - 1e-9) { best_t = t; nbest = np; for (int i = 0; i < 4; i++) best[i] = p[i]; }
            }
        }
    }
    int64_t q0 = 0;
    auto push = [&](int tiles) {
        const int64_t rows = std::min<int64_t>(static_cast<int64_t>(tiles) * qrows, nq - q0);
        if (rows > 0) { chunks.emplace_back(q0, rows); q0 += rows; }
    };
    for (int i = 0; i < nbest; i++) push(best[i]);
    for (int i = 0; i < k_tail; i++) push(G);
    if (q0 != nq) return fail(B200KNN_EINVAL, "internal: host chunk plan covers %lld of %lld rows", (long long)q0, (long long)nq);
    return B200KNN_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char *b200knn_last_error(void) { return g_last_error.c_str(); }
int b200knn_abi_version(void) { return B200KNN_ABI_VERSION; }

int b200knn_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int d = 0; d < count; d++) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) ok++;
    }
    return ok;
}

int b200knn_create(int dim, int n_devices, const int *device_ids, b200knn_index **out) {
    if (!out) return fail(B200KNN_EINVAL, "out is NULL");
    *out = nullptr;
    if (dim <= 0) return fail(B200KNN_EINVAL, "dim must be positive (got %d)", dim);
    if (n_devices > MERGE_MAX_LISTS) return fail(B200KNN_EINVAL, "at most %d devices per handle", MERGE_MAX_LISTS);
    b200knn_index *ix = new (std::nothrow) b200knn_index();
    if (!ix) return fail(B200KNN_ENOMEM, "out of host memory");
    ix->dim = dim;
    ix->kp = (dim + 7) / 8 * 8;   // BF16 row pitch must be a multiple of 16 bytes for TMA
    for (int i = 0; i < n_devices; i++) ix->device_ids.push_back(device_ids ? device_ids[i] : i);
    *out = ix;
    return B200KNN_OK;
}

int b200knn_destroy(b200knn_index *ix) {
    if (!ix) return B200KNN_OK;
    drop_group(ix);
    for (auto &s : ix->shards) {
        if (s.ready) cudaSetDevice(s.device);
        s.destroy();
    }
    if (ix->devices_ready && !ix->shards.empty()) {
        cudaSetDevice(ix->shards[0].device);
        ix->g_idx.release();
        ix->g_dist.release();
    }
    delete ix;
    return B200KNN_OK;
}

int b200knn_clear(b200knn_index *ix) {
    if (!ix) return fail(B200KNN_EINVAL, "index is NULL");
    for (auto &s : ix->shards) s.clear_pool();
    ix->n_total = 0;
    return B200KNN_OK;
}

int64_t b200knn_num_points(const b200knn_index *ix) { return ix ? ix->n_total : 0; }
int b200knn_dim(const b200knn_index *ix) { return ix ? ix->dim : 0; }

int b200knn_set_stream(b200knn_index *ix, void *stream) {
    if (!ix) return fail(B200KNN_EINVAL, "index is NULL");
    if (ix->device_ids.size() > 1) return fail(B200KNN_EINVAL, "set_stream is only valid for single-device handles");
    ix->user_stream = stream;
    for (auto &s : ix->shards) s.stream = stream ? static_cast<cudaStream_t>(stream) : s.own_stream;
    return B200KNN_OK;
}

int b200knn_set_precision(b200knn_index *ix, int tier) {
    if (!ix) return fail(B200KNN_EINVAL, "index is NULL");
    if (tier < 0 || tier > 2) return fail(B200KNN_EINVAL, "precision tier must be B200KNN_TIER_BF16 / BF16X3 / TF32 (got %d)", tier);
    TRY(ix->ensure_devices());
    for (auto &s : ix->shards) {
        if (s.tier == tier) continue;
        s.tier = tier;
        if (s.n > 0) {               // a pool is indexed already: build the tier's copy of its operands now
            CU_TRY(cudaSetDevice(s.device));
            TRY(s.convert_pool_tier(ix->dim, ix->kp));
            CU_TRY(cudaStreamSynchronize(s.stream));
        }
    }
    return B200KNN_OK;
}

int b200knn_set_profiling(b200knn_index *ix, int profiling) {
    if (!ix) return fail(B200KNN_EINVAL, "index is NULL");
    ix->profiling = profiling != 0;
    for (auto &s : ix->shards) s.profiling = ix->profiling;
    return B200KNN_OK;
}

int b200knn_get_stats(b200knn_index *ix, b200knn_stats *out) {
    if (!ix || !out) return fail(B200KNN_EINVAL, "NULL argument");
    std::memset(out, 0, sizeof(*out));
    for (auto &s : ix->shards) {
        if (!s.ready) continue;
        cudaSetDevice(s.device);
        cudaStreamSynchronize(s.stream);
        s.drain_events();
        unsigned int unc = 0;      // counted on the device (the host never learns the per-pass count)
        if (cudaMemcpy(&unc, s.scalars.p + 8, sizeof(unc), cudaMemcpyDeviceToHost) != cudaSuccess) cudaGetLastError();
        s.stats.uncertified = unc;
        out->kernel_launches += s.stats.kernel_launches;
        out->queries = std::max(out->queries, s.stats.queries);
        out->uncertified += s.stats.uncertified;
        out->exact_scanned += s.stats.exact_scanned;
        out->ms_convert += s.stats.ms_convert;
        out->ms_distance += s.stats.ms_distance;
        out->ms_rerank += s.stats.ms_rerank;
        out->ms_scan += s.stats.ms_scan;
        out->ms_wait += s.stats.ms_wait;
        out->distance_launches += s.stats.distance_launches;
        out->distance_flops += s.stats.distance_flops;
    }
    return B200KNN_OK;
}

int b200knn_reset_stats(b200knn_index *ix) {
    if (!ix) return fail(B200KNN_EINVAL, "index is NULL");
    for (auto &s : ix->shards) {
        if (s.ready) { cudaSetDevice(s.device); cudaStreamSynchronize(s.stream); s.drain_events(); cudaMemset(s.scalars.p + 8, 0, sizeof(unsigned int)); }
        s.stats = b200knn_stats{};
    }
    return B200KNN_OK;
}

static int check_matrix_args(const b200knn_index *ix, const void *p, int dtype, int64_t rows, int64_t ld, const char *what) {
    if (!ix) return fail(B200KNN_EINVAL, "index is NULL");
    if (!p && rows > 0) return fail(B200KNN_EINVAL, "%s pointer is NULL", what);
    if (dtype != B200KNN_F64 && dtype != B200KNN_F32) return fail(B200KNN_EINVAL, "%s dtype %d is not B200KNN_F64/F32", what, dtype);
    if (rows < 0) return fail(B200KNN_EINVAL, "%s row count is negative", what);
    if (ld < ix->dim) return fail(B200KNN_EINVAL, "%s leading dimension %lld < dim %d", what, (long long)ld, ix->dim);
    if (rows > 0x7fffffffll - 512) return fail(B200KNN_EINVAL, "%s has too many rows (%lld)", what, (long long)rows);
    return B200KNN_OK;
}

int b200knn_add_device(b200knn_index *ix, const void *d_data, int dtype, int64_t n, int64_t ld, int64_t index_base) {
    TRY(check_matrix_args(ix, d_data, dtype, n, ld, "data"));
    if (ix->n_total > 0) return fail(B200KNN_ESTATE, "index already holds %lld points; clear() it first", (long long)ix->n_total);
    if (n == 0) return B200KNN_OK;
    TRY(ix->ensure_devices());
    if (ix->shards.size() > 1) {
        // multi-device handle: the rows (on any device of the process) are COPIED into the shards, slice by slice over
        // NVLink, then indexed like a host add (global column means over peer memory, convert)
        if (index_base != 0) return fail(B200KNN_EINVAL, "index_base must be 0 on a multi-device handle (it shards the rows itself)");
        const int rc = add_sharded(ix, n, dtype, [&](Shard &s, void *d_rows, int64_t r0, int64_t rows) -> int {
            const size_t esz = dtype == B200KNN_F64 ? 8 : 4;
            CU_TRY(cudaMemcpy2DAsync(d_rows, static_cast<size_t>(ix->dim) * esz, static_cast<const char *>(d_data) + static_cast<size_t>(r0) * ld * esz,
                                     static_cast<size_t>(ld) * esz, static_cast<size_t>(ix->dim) * esz, static_cast<size_t>(rows), cudaMemcpyDefault, s.stream));
            return B200KNN_OK;
        });
        return cleanup_failed_add(ix, rc);
    }
    Shard &s = ix->shards[0];
    CU_TRY(cudaSetDevice(s.device));
    TRY(s.attach_pool(d_data, false, dtype, n, ld, ix->dim, ix->kp, index_base));
    TRY(s.compute_mean(d_data, dtype, n, ld, ix->dim));
    TRY(s.launch_convert(d_data, dtype, n, ld, ix->dim, ix->kp, s.x_bf.p, s.xnorm_bf.p, s.x_err.p, s.scalars.p));
    TRY(s.convert_pool_tier(ix->dim, ix->kp));
    ix->n_total = n;
    return B200KNN_OK;
}

static int add_impl(b200knn_index *ix, const void *data, int dtype, int64_t n, int64_t ld) {
    TRY(check_matrix_args(ix, data, dtype, n, ld, "data"));
    if (ix->n_total > 0) return fail(B200KNN_ESTATE, "index already holds %lld points; clear() it first", (long long)ix->n_total);
    if (n == 0) return B200KNN_OK;
    TRY(ix->ensure_devices());
    const size_t esz = dtype == B200KNN_F64 ? 8 : 4;
    return add_sharded(ix, n, dtype, [&](Shard &s, void *d_rows, int64_t r0, int64_t rows) -> int {
        // upload (pinned ring for pageable sources)
        return s.upload_rows(d_rows, static_cast<const char *>(data) + static_cast<size_t>(r0) * ld * esz, rows, ix->dim * esz, ld * esz, s.stream);
    });
}

int b200knn_add(b200knn_index *ix, const void *data, int dtype, int64_t n, int64_t ld) {
    return cleanup_failed_add(ix, add_impl(ix, data, dtype, n, ld));
}

int b200knn_debug_plan(int64_t n, int64_t nq, int kp, int num_sms, int cta_group, int max_slots, int a_budget_mb, int wide_mode,
                       int32_t *items, int64_t capacity, int32_t *geometry) {
    if (n <= 0 || nq <= 0 || kp <= 0 || num_sms <= 0 || max_slots <= 0 || a_budget_mb <= 0) return fail(B200KNN_EINVAL, "plan arguments must be positive");
    if (cta_group < 0 || cta_group > 2 || !geometry) return fail(B200KNN_EINVAL, "bad cta_group or NULL geometry");
    Shard::Sched s;
    TRY(Shard::plan_schedule(s, n, nq, kp, max_slots, cta_group, std::max(1, num_sms / 2), num_sms, a_budget_mb, wide_mode));
    const int32_t g[8] = {s.cg, s.workers, s.nrounds, s.qt, s.nt, s.max_slots, s.qg, s.wide ? 1 : 0};
    std::memcpy(geometry, g, sizeof(g));
    const int64_t count = static_cast<int64_t>(s.items.size());
    if (items) {
        if (capacity < count) return fail(B200KNN_EINVAL, "items capacity %lld < %lld", (long long)capacity, (long long)count);
        static_assert(sizeof(WorkItem) == 4 * sizeof(int32_t), "WorkItem is four ints");
        std::memcpy(items, s.items.data(), static_cast<size_t>(count) * sizeof(WorkItem));
    }
    return B200KNN_OK;
}

int b200knn_debug_chunks(int64_t n, int64_t nq, int kp, int num_sms, int64_t cap_rows, int world, int64_t *out, int64_t capacity, int64_t *count) {
    if (n <= 0 || nq <= 0 || kp <= 0 || num_sms <= 0 || cap_rows <= 0 || world <= 0 || !count) return fail(B200KNN_EINVAL, "chunk arguments must be positive");
    Shard::Sched sch;
    TRY(Shard::plan_schedule(sch, n, nq, kp, 64, 0, std::max(1, num_sms / 2), num_sms, 64, 1));
    std::vector<std::pair<int64_t, int64_t>> chunks;
    TRY(ex_chunks_from_group(nq, static_cast<int64_t>(sch.qg) * BM * sch.cg, cap_rows, chunks));
    *count = static_cast<int64_t>(chunks.size());
    if (out) {
        if (capacity < *count * (2 + 2 * world)) return fail(B200KNN_EINVAL, "capacity too small");
        int64_t *o = out;
        for (auto &c : chunks) {
            *o++ = c.first;
            *o++ = c.second;
            const int64_t slice = (c.second + world - 1) / world;          // the rule of ex_query_host_run / ex_query_device
            for (int r = 0; r < world; r++) {
                *o++ = std::min(c.second, slice * r);
                *o++ = std::min(c.second, slice * (r + 1));
            }
        }
    }
    return B200KNN_OK;
}

int b200knn_debug_host_chunks(int64_t n, int64_t nq, int kp, int dim, int elem_bytes, int k, int num_sms, int pinned, int64_t *out, int64_t capacity,
                              int64_t *count) {
    if (n <= 0 || nq <= 0 || kp <= 0 || dim <= 0 || num_sms <= 0 || k <= 0 || k > 32 || !count || (elem_bytes != 4 && elem_bytes != 8))
        return fail(B200KNN_EINVAL, "host chunk arguments out of range");
    const int C = k <= 4 ? 16 : (k <= 16 ? 32 : 64);
    HostChunkModel m;
    if (!pinned) m.upload_gbs = 40.0;
    std::vector<std::pair<int64_t, int64_t>> chunks;
    const int64_t cap_rows = std::max<int64_t>(BM * 2, std::min<int64_t>(QUERY_CHUNK, (512ll << 20) / (static_cast<int64_t>(dim) * elem_bytes) / (BM * 2) * (BM * 2)));
    TRY(plan_host_chunks(n, nq, kp, dim, static_cast<size_t>(elem_bytes), MAX_KEYS / C, 0, std::max(1, num_sms / 2), num_sms, 64, 1, cap_rows, m, 1, chunks));
    *count = static_cast<int64_t>(chunks.size());
    if (out) {
        if (capacity < *count * 2) return fail(B200KNN_EINVAL, "capacity too small");
        for (size_t i = 0; i < chunks.size(); i++) { out[2 * i] = chunks[i].first; out[2 * i + 1] = chunks[i].second; }
    }
    return B200KNN_OK;
}

int b200knn_debug_shortlists(b200knn_index *ix, float *scores, int32_t *rows, int64_t capacity, int64_t *nq, int *slots, int *c) {
    if (!ix || !scores || !rows || !nq || !slots || !c) return fail(B200KNN_EINVAL, "NULL argument");
    if (ix->shards.size() != 1 || !ix->shards[0].ready) return fail(B200KNN_ESTATE, "needs a single-device handle that has answered a query");
    Shard &s = ix->shards[0];
    *nq = s.last_nq;
    *slots = s.last_slots;
    *c = s.last_c;
    const int64_t total = s.last_nq * s.last_slots * s.last_c;
    if (total <= 0) return fail(B200KNN_ESTATE, "no tensor pass has run on this handle");
    if (capacity < total) return fail(B200KNN_EINVAL, "capacity %lld < %lld entries", (long long)capacity, (long long)total);
    CU_TRY(cudaSetDevice(s.device));
    CU_TRY(cudaStreamSynchronize(s.stream));
    CU_TRY(cudaMemcpy(scores, s.cand_s.p, static_cast<size_t>(total) * sizeof(float), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(rows, s.cand_i.p, static_cast<size_t>(total) * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return B200KNN_OK;
}

int b200knn_merge_topk_device(const int32_t *d_idx, const double *d_dist, int n_lists, int64_t nq, int kk, int32_t *d_out_idx,
                              double *d_out_dist, void *stream) {
    if (!d_idx || !d_dist || !d_out_idx || !d_out_dist) return fail(B200KNN_EINVAL, "NULL buffer");
    if (n_lists <= 0 || n_lists > MERGE_MAX_LISTS) return fail(B200KNN_EINVAL, "n_lists must be in 1..%d", MERGE_MAX_LISTS);
    if (nq < 0 || kk <= 0) return fail(B200KNN_EINVAL, "bad nq/kk");
    if (nq == 0) return B200KNN_OK;
    const unsigned blocks = static_cast<unsigned>((nq + 127) / 128);
    merge_topk_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_idx, d_dist, n_lists, nq, kk, d_out_idx, d_out_dist);
    CU_TRY(cudaGetLastError());
    return B200KNN_OK;
}

// ------------------------------------------------------------------------------------------------ NVLink exchange
int b200knn_exchange_create(int device, int rank, int world, int64_t max_nq, int max_kk, b200knn_exchange **out) {
    return ex_create(device, rank, world, 0, max_nq, max_kk, out);
}
int b200knn_exchange_create_for_queries(int device, int rank, int world, int dim, int64_t max_nq, int max_kk, b200knn_exchange **out) {
    if (dim <= 0) return fail(B200KNN_EINVAL, "dim must be positive (got %d)", dim);
    return ex_create(device, rank, world, dim, max_nq, max_kk, out);
}

int b200knn_exchange_handle(b200knn_exchange *ex, void *out_bytes) {
    if (!ex || !out_bytes) return fail(B200KNN_EINVAL, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == B200KNN_IPC_BYTES, "IPC handle size");
    CU_TRY(cudaSetDevice(ex->device));
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, ex->base));
    std::memcpy(out_bytes, &h, sizeof(h));
    return B200KNN_OK;
}

int b200knn_exchange_connect(b200knn_exchange *ex, const void *all_handles) {
    if (!ex || !all_handles) return fail(B200KNN_EINVAL, "NULL argument");
    CU_TRY(cudaSetDevice(ex->device));
    for (int p = 0; p < ex->world; p++) {
        if (p == ex->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char *>(all_handles) + static_cast<size_t>(p) * B200KNN_IPC_BYTES, sizeof(h));
        CU_TRY(cudaIpcOpenMemHandle(&ex->peer_base[p], h, cudaIpcMemLazyEnablePeerAccess));
    }
    ex->ipc_mapped = true;
    ex->connected = true;
    return B200KNN_OK;
}

int b200knn_exchange_connect_local(b200knn_exchange *ex, b200knn_exchange *const *all) {
    if (!ex || !all) return fail(B200KNN_EINVAL, "NULL argument");
    CU_TRY(cudaSetDevice(ex->device));
    for (int p = 0; p < ex->world; p++) {
        if (p == ex->rank) continue;
        if (!all[p] || all[p]->world != ex->world || all[p]->rank != p || all[p]->bytes != ex->bytes)
            return fail(B200KNN_EINVAL, "exchange %d of the group does not match (same world, sizes and rank order required)", p);
        if (all[p]->device != ex->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(all[p]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(B200KNN_ECUDA, "no peer access from device %d to device %d: %s", ex->device, all[p]->device, cudaGetErrorString(e));
            cudaGetLastError();
        }
        ex->peer_base[p] = all[p]->base;
    }
    ex->connected = true;
    return B200KNN_OK;
}

int b200knn_exchange_allgather_merge(b200knn_exchange *ex, const int32_t *d_idx, const double *d_dist, int64_t nq, int kk,
                                     int32_t *d_out_idx, double *d_out_dist, void *stream) {
    if (!ex || !d_idx || !d_dist || !d_out_idx || !d_out_dist) return fail(B200KNN_EINVAL, "NULL argument");
    if (!ex->connected) return fail(B200KNN_ESTATE, "exchange is not connected to its peers");
    CU_TRY(cudaSetDevice(ex->device));
    return ex_allgather_merge(ex, d_idx, d_dist, nq, kk, d_out_idx, d_out_dist, static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

int b200knn_exchange_destroy(b200knn_exchange *ex) {
    ex_destroy(ex);
    return B200KNN_OK;
}

static int ex_index_args(b200knn_exchange *ex, b200knn_index *ix, const char *what) {
    if (!ex || !ix) return fail(B200KNN_EINVAL, "NULL argument");
    if (ix->device_ids.size() > 1) return fail(B200KNN_EINVAL, "%s takes a single-device handle (one rank per GPU)", what);
    TRY(ix->ensure_devices());
    return ex_check_pair(ex, ix->shards[0], ix->dim, what);
}

int b200knn_exchange_add_device(b200knn_exchange *ex, b200knn_index *ix, const void *d_data, int dtype, int64_t n, int64_t ld, int64_t index_base) {
    TRY(check_matrix_args(ix, d_data, dtype, n, ld, "data"));
    TRY(ex_index_args(ex, ix, "b200knn_exchange_add_device"));
    if (ix->n_total > 0) return fail(B200KNN_ESTATE, "index already holds %lld points; clear() it first", (long long)ix->n_total);
    if (n <= 0) return fail(B200KNN_EINVAL, "every rank must hold at least one pool row");
    Shard &s = ix->shards[0];
    CU_TRY(cudaSetDevice(s.device));
    TRY(s.attach_pool(d_data, false, dtype, n, ld, ix->dim, ix->kp, index_base));
    TRY(ex_finish_add(ex, s, ix->dim, ix->kp));
    ix->n_total = n;
    return B200KNN_OK;
}

int b200knn_exchange_add(b200knn_exchange *ex, b200knn_index *ix, const void *data, int dtype, int64_t n, int64_t ld, int64_t index_base) {
    TRY(check_matrix_args(ix, data, dtype, n, ld, "data"));
    TRY(ex_index_args(ex, ix, "b200knn_exchange_add"));
    if (ix->n_total > 0) return fail(B200KNN_ESTATE, "index already holds %lld points; clear() it first", (long long)ix->n_total);
    if (n <= 0) return fail(B200KNN_EINVAL, "every rank must hold at least one pool row");
    Shard &s = ix->shards[0];
    CU_TRY(cudaSetDevice(s.device));
    const size_t esz = dtype == B200KNN_F64 ? 8 : 4;
    TRY(s.x_store.ensure(static_cast<size_t>(n) * ix->dim * esz));
    void *d_rows = s.x_store.p;
    int rc = s.attach_pool(d_rows, true, dtype, n, ix->dim, ix->dim, ix->kp, index_base);      // the shard owns d_rows from here
    if (rc == B200KNN_OK) rc = s.upload_rows(d_rows, static_cast<const char *>(data), n, ix->dim * esz, ld * esz, s.stream);
    if (rc == B200KNN_OK) rc = ex_finish_add(ex, s, ix->dim, ix->kp);
    if (rc != B200KNN_OK) {
        const std::string msg = g_last_error;
        s.clear_pool();
        cudaGetLastError();
        g_last_error = msg;
        return rc;
    }
    ix->n_total = n;
    return B200KNN_OK;
}

int b200knn_exchange_query_device(b200knn_exchange *ex, b200knn_index *ix, const void *d_query, int dtype, int64_t nq, int64_t ld, int k,
                                  unsigned flags, int32_t *d_out_idx, double *d_out_dist, int *out_kk) {
    TRY(check_matrix_args(ix, d_query, dtype, nq, ld, "query"));
    TRY(ex_index_args(ex, ix, "b200knn_exchange_query_device"));
    if (k <= 0) return fail(B200KNN_EINVAL, "k must be positive (got %d)", k);
    if (ix->n_total <= 0 || ex->n_global <= 0) return fail(B200KNN_ESTATE, "query on an empty index (use b200knn_exchange_add*)");
    if (nq > 0 && (!d_out_idx || !d_out_dist)) return fail(B200KNN_EINVAL, "output buffer is NULL");
    if (nq == 0) { if (out_kk) *out_kk = static_cast<int>(std::min<int64_t>(k, ex->n_global)); return B200KNN_OK; }
    return ex_query_device(ex, ix->shards[0], ix->dim, ix->kp, d_query, dtype, nq, ld, k, flags, d_out_idx, d_out_dist, out_kk);
}

int b200knn_exchange_query(b200knn_exchange *ex, b200knn_index *ix, const void *query, int dtype, int64_t nq, int64_t ld, int k, unsigned flags,
                           int32_t *out_idx, double *out_dist, int *out_kk) {
    TRY(check_matrix_args(ix, query, dtype, nq, ld, "query"));
    TRY(ex_index_args(ex, ix, "b200knn_exchange_query"));
    if (k <= 0) return fail(B200KNN_EINVAL, "k must be positive (got %d)", k);
    if (ix->n_total <= 0 || ex->n_global <= 0) return fail(B200KNN_ESTATE, "query on an empty index (use b200knn_exchange_add*)");
    if (nq > 0 && (!out_idx || !out_dist)) return fail(B200KNN_EINVAL, "output buffer is NULL");
    if (nq == 0) { if (out_kk) *out_kk = static_cast<int>(std::min<int64_t>(k, ex->n_global)); return B200KNN_OK; }
    return ex_query_host(ex, ix->shards[0], ix->dim, ix->kp, query, dtype, nq, ld, k, flags, out_idx, out_dist, out_kk);
}

int b200knn_query_device(b200knn_index *ix, const void *d_query, int dtype, int64_t nq, int64_t ld, int k, unsigned flags,
                         int32_t *d_out_idx, double *d_out_dist, int *out_kk) {
    TRY(check_matrix_args(ix, d_query, dtype, nq, ld, "query"));
    if (k <= 0) return fail(B200KNN_EINVAL, "k must be positive (got %d)", k);
    if (ix->n_total <= 0) return fail(B200KNN_ESTATE, "query on an empty index");
    if (ix->shards.size() != 1) return fail(B200KNN_EINVAL, "query_device is only valid for single-device handles");
    if (!d_out_idx || !d_out_dist) return fail(B200KNN_EINVAL, "output buffer is NULL");
    Shard &s = ix->shards[0];
    const int kk = static_cast<int>(std::min<int64_t>(k, s.n));
    if (out_kk) *out_kk = kk;
    const size_t esz = dtype == B200KNN_F64 ? 8 : 4;
    if (nq == 0) return B200KNN_OK;
    TRY(s.begin_call(nq));
    const bool whole = nq > QUERY_CHUNK && nq <= WHOLE_CALL_MAX_ROWS;       // several passes: one second pass for all of them, at the end
    s.accum = Shard::CallAccum{};
    for (int64_t q0 = 0; q0 < nq; q0 += QUERY_CHUNK) {
        const int64_t cq = std::min(QUERY_CHUNK, nq - q0);
        if (whole) { s.accum.on = true; s.accum.rows = nq; s.accum.off = q0; }
        const int rc = s.query_device(static_cast<const char *>(d_query) + static_cast<size_t>(q0) * ld * esz, dtype, cq, ld, ix->dim, ix->kp, k, flags,
                                      d_out_idx + q0 * kk, d_out_dist + q0 * kk, nullptr, static_cast<int>(q0));
        if (rc != B200KNN_OK) { s.accum = Shard::CallAccum{}; return rc; }
    }
    if (whole) TRY(s.finish_accumulated_call(d_query, dtype, ld, ix->dim, ix->kp, kk, flags, d_out_idx, d_out_dist));
    // one synchronisation per call: did any second-pass list overflow?  (then: exact scan of those queries)
    TRY(s.enqueue_overflow_readback());
    CU_TRY(cudaStreamSynchronize(s.stream));
    return s.fix_overflow_device(d_query, dtype, ld, *s.h_count, ix->dim, kk, flags, d_out_idx, d_out_dist);
}

int b200knn_query_self(b200knn_index *ix, int k, unsigned flags, int32_t *out_idx, double *out_dist, int *out_kk) {
    if (!ix) return fail(B200KNN_EINVAL, "index is NULL");
    if (k <= 0) return fail(B200KNN_EINVAL, "k must be positive (got %d)", k);
    if (ix->n_total <= 0) return fail(B200KNN_ESTATE, "query on an empty index");
    if (!out_idx || !out_dist) return fail(B200KNN_EINVAL, "output buffer is NULL");
    if (ix->shards.size() != 1) return query_self_sharded(ix, k, flags, out_idx, out_dist, out_kk);
    Shard &s = ix->shards[0];
    const int kk = static_cast<int>(std::min<int64_t>(k, s.n));
    if (out_kk) *out_kk = kk;
    CU_TRY(cudaSetDevice(s.device));
    TRY(s.out_idx.ensure(static_cast<size_t>(s.n) * kk));
    TRY(s.out_dist.ensure(static_cast<size_t>(s.n) * kk));
    const size_t esz = s.x_dtype == B200KNN_F64 ? 8 : 4;
    TRY(s.begin_call(s.n));
    auto pool_side = [&](int64_t q0) {       // the pool's own operands (of the handle's precision tier) from row q0 on
        QuerySide pre{s.x_bf.p + static_cast<size_t>(q0) * ix->kp, s.xnorm_bf.p + q0, s.x_err.p + q0};
        if (s.tier != 0) {
            pre.norm = s.xnorm_t.p + q0;
            pre.err = s.x_err_t.p + q0;
            if (s.tier == 1) { pre.lo = s.x_lo.p + static_cast<size_t>(q0) * ix->kp; pre.lonorm = s.x_lonorm.p + q0; }
            else pre.tf = s.x_tf.p + static_cast<size_t>(q0) * ix->kp;
        }
        return pre;
    };
    const QuerySide base = pool_side(0);
    const bool whole = s.n > QUERY_CHUNK && s.n <= WHOLE_CALL_MAX_ROWS;       // several passes: one second pass for all of them, at the end
    s.accum = Shard::CallAccum{};
    for (int64_t q0 = 0; q0 < s.n; q0 += QUERY_CHUNK) {
        const int64_t cq = std::min(QUERY_CHUNK, s.n - q0);
        const QuerySide pre = pool_side(q0);
        if (whole) { s.accum.on = true; s.accum.rows = s.n; s.accum.off = q0; s.accum.pre_base = &base; }
        const int rc = s.query_device(static_cast<const char *>(s.x_raw) + static_cast<size_t>(q0) * s.ld_x * esz, s.x_dtype, cq, s.ld_x, ix->dim, ix->kp, k,
                                      flags, s.out_idx.p + q0 * kk, s.out_dist.p + q0 * kk, &pre, static_cast<int>(q0));
        if (rc != B200KNN_OK) { s.accum = Shard::CallAccum{}; return rc; }
    }
    if (whole) TRY(s.finish_accumulated_call(s.x_raw, s.x_dtype, s.ld_x, ix->dim, ix->kp, kk, flags, s.out_idx.p, s.out_dist.p));
    auto copy_out = [&]() -> int {
        CU_TRY(cudaMemcpyAsync(out_idx, s.out_idx.p, static_cast<size_t>(s.n) * kk * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
        CU_TRY(cudaMemcpyAsync(out_dist, s.out_dist.p, static_cast<size_t>(s.n) * kk * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
        return B200KNN_OK;
    };
    TRY(copy_out());
    TRY(s.enqueue_overflow_readback());
    CU_TRY(cudaStreamSynchronize(s.stream));
    if (*s.h_count > 0) {      // rare: overflowed second-pass lists are answered by the exact scan, results copied again
        TRY(s.fix_overflow_device(s.x_raw, s.x_dtype, s.ld_x, *s.h_count, ix->dim, kk, flags, s.out_idx.p, s.out_dist.p));
        TRY(copy_out());
        CU_TRY(cudaStreamSynchronize(s.stream));
    }
    return B200KNN_OK;
}

// ---- random projection on the device ----
static int check_projected_args(b200knn_index *ix, const void *p, int dtype, int64_t rows, int64_t ld, const char *what) {
    if (!ix) return fail(B200KNN_EINVAL, "index is NULL");
    if (ix->proj_in_dim <= 0) return fail(B200KNN_ESTATE, "no projector set (b200knn_set_projector)");
    if (!p && rows > 0) return fail(B200KNN_EINVAL, "%s pointer is NULL", what);
    if (dtype != B200KNN_F64 && dtype != B200KNN_F32) return fail(B200KNN_EINVAL, "%s dtype %d is not B200KNN_F64/F32", what, dtype);
    if (rows < 0) return fail(B200KNN_EINVAL, "%s row count is negative", what);
    if (ld < ix->proj_in_dim) return fail(B200KNN_EINVAL, "%s leading dimension %lld < projector input dim %lld", what, (long long)ld, (long long)ix->proj_in_dim);
    if (rows > 0x7fffffffll - 512) return fail(B200KNN_EINVAL, "%s has too many rows (%lld)", what, (long long)rows);
    return B200KNN_OK;
}

// rows [n][in_dim] on the HOST -> d_out [n][dim] float64 on shard s's device, in row chunks through its staging buffer
// (every shard holds its own copy of the projector)
static int project_host_rows(b200knn_index *ix, Shard &s, const void *rows, int dtype, int64_t n, int64_t ld, double *d_out) {
    const size_t esz = dtype == B200KNN_F64 ? 8 : 4;
    const int64_t in_dim = ix->proj_in_dim;
    const int64_t chunk = std::max<int64_t>(PJ_T, std::min<int64_t>((n + PJ_T - 1) / PJ_T * PJ_T, (512ll << 20) / (in_dim * static_cast<int64_t>(esz)) / PJ_T * PJ_T));
    TRY(s.proj_stage.ensure(static_cast<size_t>(std::min(chunk, n)) * in_dim * esz));
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t cr = std::min(chunk, n - r0);
        TRY(s.upload_rows(s.proj_stage.p, static_cast<const char *>(rows) + static_cast<size_t>(r0) * ld * esz, cr, static_cast<size_t>(in_dim) * esz,
                          static_cast<size_t>(ld) * esz, s.stream));
        const dim3 grid(static_cast<unsigned>((ix->dim + PJ_T - 1) / PJ_T), static_cast<unsigned>((cr + PJ_T - 1) / PJ_T));
        s.stats.kernel_launches++;
        if (dtype == B200KNN_F64)
            project_kernel<double><<<grid, 256, 0, s.stream>>>(reinterpret_cast<const double *>(s.proj_stage.p), in_dim, static_cast<int>(cr), s.projector.p,
                                                               static_cast<int>(in_dim), ix->dim, d_out + r0 * ix->dim, ix->dim);
        else
            project_kernel<float><<<grid, 256, 0, s.stream>>>(reinterpret_cast<const float *>(s.proj_stage.p), in_dim, static_cast<int>(cr), s.projector.p,
                                                              static_cast<int>(in_dim), ix->dim, d_out + r0 * ix->dim, ix->dim);
        CU_TRY(cudaGetLastError());
        // the next chunk's upload reuses the staging buffer: its DMA is enqueued on the same stream, behind this kernel
    }
    return B200KNN_OK;
}

int b200knn_set_projector(b200knn_index *ix, const double *projector, int64_t in_dim, int64_t ld) {
    if (!ix) return fail(B200KNN_EINVAL, "index is NULL");
    if (!projector) return fail(B200KNN_EINVAL, "projector pointer is NULL");
    if (in_dim <= 0 || in_dim > 0x7fffffffll) return fail(B200KNN_EINVAL, "projector input dim %lld out of range", (long long)in_dim);
    if (ld < ix->dim) return fail(B200KNN_EINVAL, "projector leading dimension %lld < dim %d", (long long)ld, ix->dim);
    TRY(ix->ensure_devices());
    ix->proj_in_dim = 0;
    for (auto &s : ix->shards) {       // every device of the handle keeps a copy (a few MB: in_dim x dim float64)
        CU_TRY(cudaSetDevice(s.device));
        CU_TRY(cudaStreamSynchronize(s.stream));          // nothing in flight may still read the previous projector
        TRY(s.projector.ensure(static_cast<size_t>(in_dim) * ix->dim));
        TRY(s.upload_rows(s.projector.p, reinterpret_cast<const char *>(projector), in_dim, static_cast<size_t>(ix->dim) * 8, static_cast<size_t>(ld) * 8, s.stream));
        CU_TRY(cudaStreamSynchronize(s.stream));
    }
    ix->proj_in_dim = in_dim;
    return B200KNN_OK;
}

int b200knn_project_rows(b200knn_index *ix, const void *rows, int dtype, int64_t n, int64_t ld, double *out) {
    TRY(check_projected_args(ix, rows, dtype, n, ld, "rows"));
    if (n == 0) return B200KNN_OK;
    if (!out) return fail(B200KNN_EINVAL, "output buffer is NULL");
    Shard &s = ix->shards[0];
    CU_TRY(cudaSetDevice(s.device));
    TRY(s.proj_rows.ensure(static_cast<size_t>(n) * ix->dim));
    TRY(project_host_rows(ix, s, rows, dtype, n, ld, s.proj_rows.p));
    CU_TRY(cudaMemcpyAsync(out, s.proj_rows.p, static_cast<size_t>(n) * ix->dim * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CU_TRY(cudaStreamSynchronize(s.stream));
    return B200KNN_OK;
}

int b200knn_add_projected(b200knn_index *ix, const void *rows, int dtype, int64_t n, int64_t ld) {
    TRY(check_projected_args(ix, rows, dtype, n, ld, "rows"));
    if (ix->n_total > 0) return fail(B200KNN_ESTATE, "index already holds %lld points; clear() it first", (long long)ix->n_total);
    if (n == 0) return B200KNN_OK;
    const size_t esz = dtype == B200KNN_F64 ? 8 : 4;
    // every shard projects its own slice of the rows straight into its store (float64), then the usual indexing
    return cleanup_failed_add(ix, add_sharded(ix, n, B200KNN_F64, [&](Shard &s, void *d_rows, int64_t r0, int64_t cnt) -> int {
        return project_host_rows(ix, s, static_cast<const char *>(rows) + static_cast<size_t>(r0) * ld * esz, dtype, cnt, ld, static_cast<double *>(d_rows));
    }));
}

int b200knn_query_projected(b200knn_index *ix, const void *rows, int dtype, int64_t nq, int64_t ld, int k, unsigned flags,
                            int32_t *out_idx, double *out_dist, int *out_kk) {
    TRY(check_projected_args(ix, rows, dtype, nq, ld, "query"));
    if (k <= 0) return fail(B200KNN_EINVAL, "k must be positive (got %d)", k);
    if (ix->n_total <= 0) return fail(B200KNN_ESTATE, "query on an empty index");
    if (nq > 0 && (!out_idx || !out_dist)) return fail(B200KNN_EINVAL, "output buffer is NULL");
    const int kk = static_cast<int>(std::min<int64_t>(k, ix->n_total));
    if (out_kk) *out_kk = kk;
    if (nq == 0) return B200KNN_OK;
    Shard &s = ix->shards[0];
    const size_t esz = dtype == B200KNN_F64 ? 8 : 4;
    CU_TRY(cudaSetDevice(s.device));
    TRY(s.proj_rows.ensure(static_cast<size_t>(std::min(nq, QUERY_CHUNK)) * ix->dim));
    if (ix->shards.size() > 1) {
        // multi-device handle: shard 0 projects a chunk, the projected rows go through pinned host memory into the
        // collective query (every shard uploads 1/G of them over its own link) — same answers as on one device
        const int64_t step = std::min(nq, QUERY_CHUNK);
        double *h_rows = nullptr;
        CU_TRY(cudaMallocHost(reinterpret_cast<void **>(&h_rows), static_cast<size_t>(step) * ix->dim * sizeof(double)));
        int rc = B200KNN_OK;
        for (int64_t q0 = 0; q0 < nq && rc == B200KNN_OK; q0 += step) {
            const int64_t cq = std::min(step, nq - q0);
            rc = project_host_rows(ix, s, static_cast<const char *>(rows) + static_cast<size_t>(q0) * ld * esz, dtype, cq, ld, s.proj_rows.p);
            if (rc == B200KNN_OK && (cudaMemcpyAsync(h_rows, s.proj_rows.p, static_cast<size_t>(cq) * ix->dim * sizeof(double), cudaMemcpyDeviceToHost, s.stream) != cudaSuccess ||
                                     cudaStreamSynchronize(s.stream) != cudaSuccess))
                rc = fail(B200KNN_ECUDA, "reading back the projected rows failed: %s", cudaGetErrorString(cudaGetLastError()));
            if (rc == B200KNN_OK) rc = b200knn_query(ix, h_rows, B200KNN_F64, cq, ix->dim, k, flags, out_idx + q0 * kk, out_dist + q0 * kk, nullptr);
            cudaSetDevice(s.device);
        }
        const std::string msg = g_last_error;
        cudaSetDevice(s.device);
        cudaFreeHost(h_rows);
        g_last_error = msg;
        return rc;
    }
    TRY(s.out_idx.ensure(static_cast<size_t>(std::min(nq, QUERY_CHUNK)) * kk));
    TRY(s.out_dist.ensure(static_cast<size_t>(std::min(nq, QUERY_CHUNK)) * kk));
    for (int64_t q0 = 0; q0 < nq; q0 += QUERY_CHUNK) {
        const int64_t cq = std::min(QUERY_CHUNK, nq - q0);
        TRY(project_host_rows(ix, s, static_cast<const char *>(rows) + static_cast<size_t>(q0) * ld * esz, dtype, cq, ld, s.proj_rows.p));
        TRY(s.query_device_sync(s.proj_rows.p, B200KNN_F64, cq, ix->dim, ix->dim, ix->kp, k, flags, s.out_idx.p, s.out_dist.p));
        CU_TRY(cudaMemcpyAsync(out_idx + q0 * kk, s.out_idx.p, static_cast<size_t>(cq) * kk * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
        CU_TRY(cudaMemcpyAsync(out_dist + q0 * kk, s.out_dist.p, static_cast<size_t>(cq) * kk * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
        CU_TRY(cudaStreamSynchronize(s.stream));
    }
    return B200KNN_OK;
}

int b200knn_ball_membership(b200knn_index *ix, const void *query, int dtype, int64_t nq, int64_t ld, const double *radius2,
                            unsigned char *out_member) {
    TRY(check_matrix_args(ix, query, dtype, nq, ld, "query"));
    if (ix->n_total <= 0) return fail(B200KNN_ESTATE, "membership query on an empty index");
    if (!radius2 || (nq > 0 && !out_member)) return fail(B200KNN_EINVAL, "NULL buffer");
    if (nq == 0) return B200KNN_OK;
    const size_t esz = dtype == B200KNN_F64 ? 8 : 4;
    const int dim = ix->dim;
    const int64_t chunk = std::max<int64_t>(BM, std::min<int64_t>(QUERY_CHUNK, (512ll << 20) / (static_cast<int64_t>(dim) * esz) / BM * BM));
    std::vector<unsigned char> tmp;
    std::memset(out_member, 0, static_cast<size_t>(nq));
    int64_t row0 = 0;
    for (auto &s : ix->shards) {
        if (s.n <= 0) continue;
        CU_TRY(cudaSetDevice(s.device));
        TRY(s.radius2.ensure(s.n));
        CU_TRY(cudaMemcpyAsync(s.radius2.p, radius2 + row0, static_cast<size_t>(s.n) * sizeof(double), cudaMemcpyHostToDevice, s.stream));
        row0 += s.n;
        for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
            const int64_t cq = std::min(chunk, nq - q0);
            TRY(s.q_stage.ensure(static_cast<size_t>(cq) * dim * esz));
            TRY(s.member.ensure(cq));
            CU_TRY(cudaMemsetAsync(s.member.p, 0, static_cast<size_t>(cq), s.stream));
            TRY(s.upload_rows(s.q_stage.p, static_cast<const char *>(query) + static_cast<size_t>(q0) * ld * esz, cq, static_cast<size_t>(dim) * esz,
                              static_cast<size_t>(ld) * esz, s.stream));
            TRY(s.ball_membership(s.q_stage.p, dtype, cq, dim, dim, ix->kp, s.member.p));
            tmp.resize(cq);
            CU_TRY(cudaMemcpyAsync(tmp.data(), s.member.p, static_cast<size_t>(cq), cudaMemcpyDeviceToHost, s.stream));
            CU_TRY(cudaStreamSynchronize(s.stream));
            for (int64_t i = 0; i < cq; i++) out_member[q0 + i] |= tmp[i];
        }
    }
    return B200KNN_OK;
}

int b200knn_query(b200knn_index *ix, const void *query, int dtype, int64_t nq, int64_t ld, int k, unsigned flags, int32_t *out_idx,
                  double *out_dist, int *out_kk) {
    TRY(check_matrix_args(ix, query, dtype, nq, ld, "query"));
    if (k <= 0) return fail(B200KNN_EINVAL, "k must be positive (got %d)", k);
    if (ix->n_total <= 0) return fail(B200KNN_ESTATE, "query on an empty index");
    if (nq > 0 && (!out_idx || !out_dist)) return fail(B200KNN_EINVAL, "output buffer is NULL");
    const int kk = static_cast<int>(std::min<int64_t>(k, ix->n_total));
    if (out_kk) *out_kk = kk;
    if (nq == 0) return B200KNN_OK;
    const size_t esz = dtype == B200KNN_F64 ? 8 : 4;
    const int G = static_cast<int>(ix->shards.size());
    const int dim = ix->dim;
    auto upload = [&](Shard &s, void *dst, const char *src, int64_t rows, cudaStream_t st) -> int {
        return s.upload_rows(dst, src, rows, static_cast<size_t>(dim) * esz, static_cast<size_t>(ld) * esz, st);
    };
    if (G == 1) {
        // ---- single device: double-buffered pipeline, the upload of chunk i+1 overlaps the compute of chunk i ----
        Shard &s = ix->shards[0];
        CU_TRY(cudaSetDevice(s.device));
        // Chunking follows the kernel's schedule: one chunk = one full group of query tiles (a round that keeps every
        // SM busy); the ragged remainder goes FIRST, so the first upload - the only one nothing can hide - is short.
        // Whole-call mode: when the call's query rows fit a device buffer ($B200KNN_CALL_BUFFER_MB, default 4096) every
        // chunk is uploaded into its own slice of it — no stage buffer is recycled, so no upload ever waits for a compute
        // pass — and the call runs ONE second pass at its end instead of one per chunk (Shard::CallAccum).
        const int64_t call_buffer_mb = []() { const char *e = getenv("B200KNN_CALL_BUFFER_MB"); return e ? std::max<int64_t>(0, atoll(e)) : 4096ll; }();
        // $B200KNN_UPLOAD_RAMP: 0 never, 1 for pageable sources only, 2 (default) for every source.  Measured at config 3 on one
        // GPU, A/B on one box (bench.py e2e / e2e_pageable, ms per 30k-query call): page-locked rows 39.81 -> 39.28 (gap to the
        // device-resident step 1.76 -> 0.95 ms), pageable NumPy rows 42.36 -> 40.34.  A first version of the model allowed chunks
        // of 1-4 query tiles and lost what the earlier start bought (40.4 against 39.9): such chunks have dozens of pool streams
        // per query tile - shortlists to write, sort and bound - and a pipeline fill each; hence min_part_tiles.
        const int ramp_mode = []() { const char *e = getenv("B200KNN_UPLOAD_RAMP"); return e ? atoi(e) : 2; }();
        const bool can_whole = kk <= 32 && !(flags & B200KNN_FLAG_FORCE_SCAN) && nq <= WHOLE_CALL_MAX_ROWS &&
                               static_cast<int64_t>(nq) * dim * static_cast<int64_t>(esz) <= (call_buffer_mb << 20);
        std::vector<std::pair<int64_t, int64_t>> chunks;   // (first row, rows)
        const int64_t cap_rows = std::max<int64_t>(BM * 2, std::min<int64_t>(QUERY_CHUNK, (512ll << 20) / (static_cast<int64_t>(dim) * esz) / (BM * 2) * (BM * 2)));
        cudaPointerAttributes attr;
        const bool pinned = cudaPointerGetAttributes(&attr, query) == cudaSuccess && (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
        cudaGetLastError();
        if (can_whole && (ramp_mode >= 2 || (ramp_mode == 1 && !pinned)) && nq >= 4 * BM * 2) {
            // a ramp of growing chunks sized on the kernel's schedule and the link speed (plan_host_chunks)
            HostChunkModel hm;
            if (!pinned) hm.upload_gbs = 40.0;
            const int C = kk <= 4 ? 16 : (kk <= 16 ? 32 : 64);
            const int64_t key[8] = {s.n, nq, ix->kp, dim, static_cast<int64_t>(esz), C, pinned ? 1 : 0, s.tier};
            if (std::memcmp(key, s.host_chunk_key, sizeof(key)) != 0 || s.host_chunks.empty()) {      // (the search costs ~0.1 ms of host time: cached per shape)
                TRY(plan_host_chunks(s.n, nq, s.tier ? 2 * ix->kp : ix->kp, dim, esz, MAX_KEYS / C, s.forced_cg, s.max_pairs, s.num_sms, s.a_budget_mb, s.wide_mode,
                                     cap_rows, hm, s.tier == 1 ? 3 : (s.tier == 2 ? 2 : 1), s.host_chunks));
                std::memcpy(s.host_chunk_key, key, sizeof(key));
            }
            chunks = s.host_chunks;
        } else {
            Shard::Sched sch;
            TRY(s.plan(sch, nq, ix->kp, 64));
            int64_t group_rows = static_cast<int64_t>(sch.qg) * BM * sch.cg;
            group_rows = std::max<int64_t>(BM * 2, std::min(group_rows, cap_rows));
            if (nq <= group_rows + group_rows / 4) {
                chunks.emplace_back(0, nq);
            } else {
                const int64_t rem = nq % group_rows;
                int64_t q0 = 0;
                if (rem > 0) { chunks.emplace_back(0, rem); q0 = rem; }
                for (; q0 < nq; q0 += group_rows) chunks.emplace_back(q0, std::min(group_rows, nq - q0));
            }
        }
        int64_t max_rows = 0;
        for (auto &c : chunks) max_rows = std::max(max_rows, c.second);
        const int64_t nchunks = static_cast<int64_t>(chunks.size());
        const bool whole = can_whole && nchunks > 1;
        if (whole) {
            TRY(s.q_stage.ensure(static_cast<size_t>(nq) * dim * esz));
        } else {
            TRY(s.q_stage.ensure(static_cast<size_t>(max_rows) * dim * esz));
            if (nchunks > 1) TRY(s.q_stage2.ensure(static_cast<size_t>(max_rows) * dim * esz));
        }
        TRY(s.out_idx.ensure(static_cast<size_t>(nq) * kk));
        TRY(s.out_dist.ensure(static_cast<size_t>(nq) * kk));
        unsigned char *stage[2] = {s.q_stage.p, s.q_stage2.p};
        auto stage_of = [&](int64_t c) { return whole ? s.q_stage.p + static_cast<size_t>(chunks[c].first) * dim * esz : stage[c & 1]; };
        const char *src = static_cast<const char *>(query);
        TRY(s.begin_call(nq));
        auto copy_out = [&]() -> int {
            CU_TRY(cudaMemcpyAsync(out_idx, s.out_idx.p, static_cast<size_t>(nq) * kk * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
            CU_TRY(cudaMemcpyAsync(out_dist, s.out_dist.p, static_cast<size_t>(nq) * kk * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
            return B200KNN_OK;
        };
        // after the call's single synchronisation: overflowed second-pass lists (rare) -> exact scan of those queries,
        // their rows re-uploaded from the caller's buffer (the stage buffers have been recycled), results copied again
        auto finish = [&](Shard &sh) -> int {
            const int nov = *sh.h_count;
            if (nov <= 0) return B200KNN_OK;
            if (whole) TRY(sh.fix_overflow_device(sh.q_stage.p, dtype, dim, nov, dim, kk, flags, sh.out_idx.p, sh.out_dist.p));
            else TRY(sh.fix_overflow_host(query, dtype, ld, nov, dim, kk, flags, sh.out_idx.p, sh.out_dist.p));
            TRY(copy_out());
            CU_TRY(cudaStreamSynchronize(sh.stream));
            return B200KNN_OK;
        };
        // The uploads run on their own host thread (a pageable source keeps that thread busy with memcpy) and their
        // own stream; this thread enqueues the compute.  uploaded / consumed count chunks; the CUDA events order the
        // streams, the counters order the host threads.
        if (nchunks == 1) {   // small calls (the trainer's 24-row loop): no helper thread, everything on one stream
            TRY(upload(s, stage[0], src, nq, s.stream));
            s.defer_second_pass = true;       // decided after the call's one synchronisation: usually there is nothing to do
            const int rcq = s.query_device(stage[0], dtype, nq, dim, dim, ix->kp, k, flags, s.out_idx.p, s.out_dist.p);
            s.defer_second_pass = false;
            TRY(rcq);
            TRY(copy_out());
            TRY(s.enqueue_uncertified_readback());
            TRY(s.enqueue_overflow_readback());
            CU_TRY(cudaStreamSynchronize(s.stream));      // the only synchronisation of a small call
            if (s.deferred.armed && s.h_count[1] > 0) {
                TRY(s.run_deferred_second_pass());
                TRY(copy_out());
                TRY(s.enqueue_overflow_readback());
                CU_TRY(cudaStreamSynchronize(s.stream));
            }
            s.deferred.armed = false;
            return finish(s);
        }
        std::atomic<int64_t> uploaded{0}, consumed{0};
        std::atomic<int> up_rc{B200KNN_OK};
        std::string up_err;
        std::thread uploader([&]() {
            cudaSetDevice(s.device);
            for (int64_t c = 0; c < nchunks; c++) {
                const int b = static_cast<int>(c & 1);
                if (c >= 2) {
                    // chunk c-2 enqueued: its wait on ev_copied[b] is in the stream, the event may be recorded again
                    while (consumed.load(std::memory_order_acquire) < c - 1) std::this_thread::yield();
                    if (!whole) cudaStreamWaitEvent(s.copy_stream, s.ev_consumed[b], 0);      // (whole-call mode recycles no buffer)
                }
                int rc = upload(s, stage_of(c), src + static_cast<size_t>(chunks[c].first) * ld * esz, chunks[c].second, s.copy_stream);
                if (rc == B200KNN_OK && cudaEventRecord(s.ev_copied[b], s.copy_stream) != cudaSuccess) rc = B200KNN_ECUDA;
                if (rc != B200KNN_OK) {
                    up_err = g_last_error;      // thread-local in the uploader: hand it over
                    up_rc.store(rc);
                    uploaded.store(nchunks, std::memory_order_release);
                    return;
                }
                uploaded.store(c + 1, std::memory_order_release);
            }
        });
        int rc_main = B200KNN_OK;
        s.accum = Shard::CallAccum{};
        for (int64_t c = 0; c < nchunks && rc_main == B200KNN_OK; c++) {
            const int b = static_cast<int>(c & 1);
            const int64_t q0 = chunks[c].first, cq = chunks[c].second;
            while (uploaded.load(std::memory_order_acquire) < c + 1) std::this_thread::yield();
            if (up_rc.load() != B200KNN_OK) break;
            if (cudaStreamWaitEvent(s.stream, s.ev_copied[b], 0) != cudaSuccess) { rc_main = fail(B200KNN_ECUDA, "cudaStreamWaitEvent failed"); break; }
            if (whole) { s.accum.on = true; s.accum.rows = nq; s.accum.off = q0; }
            rc_main = s.query_device(stage_of(c), dtype, cq, dim, dim, ix->kp, k, flags, s.out_idx.p + q0 * kk, s.out_dist.p + q0 * kk, nullptr, static_cast<int>(q0));
            if (rc_main == B200KNN_OK && cudaEventRecord(s.ev_consumed[b], s.stream) != cudaSuccess) rc_main = fail(B200KNN_ECUDA, "cudaEventRecord failed");
            consumed.store(c + 1, std::memory_order_release);
        }
        consumed.store(nchunks + 2, std::memory_order_release);   // never leave the uploader waiting
        uploader.join();
        if (up_rc.load() != B200KNN_OK || rc_main != B200KNN_OK) s.accum = Shard::CallAccum{};
        if (up_rc.load() != B200KNN_OK) return fail(up_rc.load(), "%s", up_err.c_str());
        if (rc_main != B200KNN_OK) return rc_main;
        if (whole) TRY(s.finish_accumulated_call(s.q_stage.p, dtype, dim, dim, ix->kp, kk, flags, s.out_idx.p, s.out_dist.p));
        TRY(copy_out());
        TRY(s.enqueue_overflow_readback());
        CU_TRY(cudaStreamSynchronize(s.stream));
        return finish(s);
    }
    // ---- multi-device handle: every chunk is uploaded ONCE (to shard 0, on its own host thread and stream, double
    // buffered) and broadcast to the other shards over NVLink; every shard answers on its own host thread (the
    // per-shard call synchronises its stream); results are gathered on shard 0 with peer copies and merged there ----
    std::vector<int> active;
    for (int g = 0; g < G; g++)
        if (ix->shards[g].n > 0) active.push_back(g);
    const int lists = static_cast<int>(active.size());
    if (kk <= 32 && ix->exch_shards == active && !ix->exch.empty()) {
        // The collective protocol (exchange_host.cuh), one host thread per shard: every shard uploads 1/lists of every chunk,
        // BF16 rows + norms and (behind the tensor pass) the original rows are broadcast over NVLink by the copy engines,
        // bound exchange + globally pruned exact re-rank, list exchange + merge; no host synchronisation between chunks.
        std::vector<ExCall> calls(lists);
        for (int i = 0; i < lists; i++)       // phase 1: geometry + every allocation, nothing collective yet
            TRY(ex_query_host_prepare(ix->exch[i], ix->shards[active[i]], ix->kp, nq, k, flags, calls[i]));
        return for_each_rank(lists, [&](int i) -> int {
            return ex_query_host_run(ix->exch[i], ix->shards[active[i]], dim, ix->kp, query, dtype, nq, ld, k, flags,
                                     i == 0 ? out_idx : nullptr, i == 0 ? out_dist : nullptr, calls[i]);
        });
    }
    // k > 32 (exact scan per shard, lists of any length): every chunk uploaded once and broadcast, per-shard answers
    // gathered on shard 0 and merged there
    Shard &s0 = ix->shards[active[0]];
    // chunks of one query-tile group of the per-shard kernel (see the single-device path)
    std::vector<std::pair<int64_t, int64_t>> chunks;
    {
        Shard::Sched sch;
        TRY(s0.plan(sch, nq, ix->kp, 64));
        int64_t group_rows = static_cast<int64_t>(sch.qg) * BM * sch.cg;
        const int64_t cap_rows = std::max<int64_t>(BM * 2, (512ll << 20) / (static_cast<int64_t>(dim) * esz) / (BM * 2) * (BM * 2));
        group_rows = std::max<int64_t>(BM * 2, std::min(group_rows, cap_rows));
        if (nq <= group_rows + group_rows / 4) {
            chunks.emplace_back(0, nq);
        } else {
            const int64_t rem = nq % group_rows;
            int64_t q0 = 0;
            if (rem > 0) { chunks.emplace_back(0, rem); q0 = rem; }
            for (; q0 < nq; q0 += group_rows) chunks.emplace_back(q0, std::min(group_rows, nq - q0));
        }
    }
    const int64_t nchunks = static_cast<int64_t>(chunks.size());
    int64_t max_rows = 0;
    for (auto &c : chunks) max_rows = std::max(max_rows, c.second);
    for (int g : active) {
        Shard &s = ix->shards[g];
        CU_TRY(cudaSetDevice(s.device));
        TRY(s.q_stage.ensure(static_cast<size_t>(max_rows) * dim * esz));
        if (nchunks > 1) TRY(s.q_stage2.ensure(static_cast<size_t>(max_rows) * dim * esz));
        TRY(s.out_idx.ensure(static_cast<size_t>(max_rows) * kk));
        TRY(s.out_dist.ensure(static_cast<size_t>(max_rows) * kk));
        if (s.n < kk) {   // a shard with fewer than kk rows answers min(kk, rows) per query; its lists are padded to kk below
            TRY(s.pad_idx.ensure(static_cast<size_t>(max_rows) * kk));
            TRY(s.pad_dist.ensure(static_cast<size_t>(max_rows) * kk));
        }
    }
    CU_TRY(cudaSetDevice(s0.device));
    TRY(ix->g_idx.ensure(static_cast<size_t>(lists) * max_rows * kk));
    TRY(ix->g_dist.ensure(static_cast<size_t>(lists) * max_rows * kk));
    std::vector<cudaEvent_t> done(lists, nullptr);
    for (int i = 0; i < lists; i++) {
        CU_TRY(cudaSetDevice(ix->shards[active[i]].device));
        CU_TRY(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    }
    std::atomic<int64_t> uploaded{0}, consumed{0};
    std::atomic<int> up_rc{B200KNN_OK};
    std::string up_err;
    std::thread uploader([&]() {
        for (int64_t c = 0; c < nchunks; c++) {
            const int b = static_cast<int>(c & 1);
            while (consumed.load(std::memory_order_acquire) < c - 1) std::this_thread::yield();   // chunk c-2 fully done (host-synchronised)
            cudaSetDevice(s0.device);
            unsigned char *dst0 = b ? s0.q_stage2.p : s0.q_stage.p;
            int rc = upload(s0, dst0, static_cast<const char *>(query) + static_cast<size_t>(chunks[c].first) * ld * esz, chunks[c].second, s0.copy_stream);
            if (rc == B200KNN_OK && cudaEventRecord(s0.ev_copied[b], s0.copy_stream) != cudaSuccess) rc = B200KNN_ECUDA;
            const size_t qbytes = static_cast<size_t>(chunks[c].second) * dim * esz;
            for (int i = 1; i < lists && rc == B200KNN_OK; i++) {      // NVLink broadcast on each receiver's copy stream
                Shard &s = ix->shards[active[i]];
                cudaSetDevice(s.device);
                unsigned char *dst = b ? s.q_stage2.p : s.q_stage.p;
                if (cudaStreamWaitEvent(s.copy_stream, s0.ev_copied[b], 0) != cudaSuccess ||
                    cudaMemcpyPeerAsync(dst, s.device, dst0, s0.device, qbytes, s.copy_stream) != cudaSuccess ||
                    cudaEventRecord(s.ev_copied[b], s.copy_stream) != cudaSuccess)
                    rc = B200KNN_ECUDA;
            }
            if (rc != B200KNN_OK) {
                up_err = g_last_error.empty() ? std::string("query upload / broadcast failed") : g_last_error;
                up_rc.store(rc);
                uploaded.store(nchunks, std::memory_order_release);
                return;
            }
            uploaded.store(c + 1, std::memory_order_release);
        }
    });
    int rc_all = B200KNN_OK;
    std::string err_all;
    for (int64_t c = 0; c < nchunks && rc_all == B200KNN_OK; c++) {
        const int b = static_cast<int>(c & 1);
        const int64_t q0 = chunks[c].first, cq = chunks[c].second;
        while (uploaded.load(std::memory_order_acquire) < c + 1) std::this_thread::yield();
        if (up_rc.load() != B200KNN_OK) break;
        std::vector<int> rcs(lists, B200KNN_OK);
        std::vector<std::string> errs(lists);
        auto work = [&](int i) {
            Shard &s = ix->shards[active[i]];
            int rc = B200KNN_OK;
            cudaSetDevice(s.device);
            if (cudaStreamWaitEvent(s.stream, s.ev_copied[b], 0) != cudaSuccess) rc = fail(B200KNN_ECUDA, "cudaStreamWaitEvent failed");
            unsigned char *qp = b ? s.q_stage2.p : s.q_stage.p;
            if (rc == B200KNN_OK) rc = s.query_device_sync(qp, dtype, cq, dim, dim, ix->kp, kk, flags, s.out_idx.p, s.out_dist.p);
            if (rc == B200KNN_OK) {
                const int32_t *src_i = s.out_idx.p;
                const double *src_d = s.out_dist.p;
                cudaError_t e = cudaSuccess;
                if (s.n < kk) {
                    s.stats.kernel_launches++;
                    pad_topk_kernel<<<static_cast<unsigned>(std::min<int64_t>(s.num_sms * 4, (cq * kk + 255) / 256)), 256, 0, s.stream>>>(
                        s.out_idx.p, s.out_dist.p, cq, static_cast<int>(s.n), kk, s.pad_idx.p, s.pad_dist.p);
                    e = cudaGetLastError();
                    src_i = s.pad_idx.p;
                    src_d = s.pad_dist.p;
                }
                if (e == cudaSuccess)
                    e = cudaMemcpyPeerAsync(ix->g_idx.p + static_cast<size_t>(i) * cq * kk, s0.device, src_i, s.device,
                                            static_cast<size_t>(cq) * kk * sizeof(int32_t), s.stream);
                if (e == cudaSuccess)
                    e = cudaMemcpyPeerAsync(ix->g_dist.p + static_cast<size_t>(i) * cq * kk, s0.device, src_d, s.device,
                                            static_cast<size_t>(cq) * kk * sizeof(double), s.stream);
                if (e == cudaSuccess) e = cudaEventRecord(done[i], s.stream);
                if (e != cudaSuccess) rc = fail(B200KNN_ECUDA, "gathering shard results failed: %s", cudaGetErrorString(e));
            }
            rcs[i] = rc;
            if (rc != B200KNN_OK) errs[i] = g_last_error;
        };
        std::vector<std::thread> th;
        for (int i = 1; i < lists; i++) th.emplace_back(work, i);
        work(0);
        for (auto &t : th) t.join();
        for (int i = 0; i < lists; i++)
            if (rcs[i] != B200KNN_OK && rc_all == B200KNN_OK) { rc_all = rcs[i]; err_all = errs[i]; }
        if (rc_all != B200KNN_OK) break;
        cudaSetDevice(s0.device);
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < lists && e == cudaSuccess; i++) e = cudaStreamWaitEvent(s0.stream, done[i], 0);
        if (e == cudaSuccess) {
            int rm = b200knn_merge_topk_device(ix->g_idx.p, ix->g_dist.p, lists, cq, kk, s0.out_idx.p, s0.out_dist.p, s0.stream);
            if (rm != B200KNN_OK) { rc_all = rm; err_all = g_last_error; break; }
            s0.stats.kernel_launches++;
            e = cudaMemcpyAsync(out_idx + q0 * kk, s0.out_idx.p, static_cast<size_t>(cq) * kk * sizeof(int32_t), cudaMemcpyDeviceToHost, s0.stream);
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_dist + q0 * kk, s0.out_dist.p, static_cast<size_t>(cq) * kk * sizeof(double), cudaMemcpyDeviceToHost, s0.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s0.stream);      // also: every shard is done with stage buffer b
        if (e != cudaSuccess) { rc_all = B200KNN_ECUDA; err_all = std::string("multi-device merge failed: ") + cudaGetErrorString(e); break; }
        consumed.store(c + 1, std::memory_order_release);
    }
    consumed.store(nchunks + 2, std::memory_order_release);   // never leave the uploader waiting
    uploader.join();
    for (auto ev : done) cudaEventDestroy(ev);
    if (up_rc.load() != B200KNN_OK) return fail(up_rc.load(), "%s", up_err.c_str());
    if (rc_all != B200KNN_OK) return fail(rc_all, "%s", err_all.c_str());
    return B200KNN_OK;
}

}  // extern "C"
