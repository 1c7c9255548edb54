// Row-sharded pools, one rank per GPU: the peer-memory exchange object and the collective add / query protocol
// (b200knn_exchange_* in include/b200knn.h).  A "rank" is a (Shard, exchange) pair: a process of a torchrun job (buffers
// mapped into the peers with CUDA IPC) or one device of a single-process multi-device handle (buffers reached through
// plain peer access).  No collective library, no host synchronisation between the steps of a query:
//
//   add     every rank sums the columns of its shard, stores the sums into every rank's buffer and raises a flag; the
//           global column means (the centring vector MUST be the same everywhere: a query row converted by one rank is
//           compared with pool rows converted by another) follow from the gathered sums in rank order.
//   query   per chunk of query rows: every rank uploads and converts 1/world of the rows and broadcasts the BF16 rows +
//           norms (what the tensor pass needs) over NVLink by peer stores; tensor pass on all rows against the local shard;
//           every rank publishes an upper bound on its k-th nearest distance per query; the exact re-rank evaluates only
//           candidates that survive the minimum of those bounds (the replicated re-rank of round 1 becomes 1/world of it
//           per rank) and reads the ORIGINAL row of such a query straight from the rank that uploaded it (peer loads:
//           the float64 rows are never broadcast); the exact local lists are exchanged and merged.
#pragma once
#include <chrono>
#include "shard.cuh"

enum ExFlag { F_TOPK = 0, F_BOUND, F_QBF, F_CONSUMED, F_MEAN, F_NKINDS };

struct b200knn_exchange {
    int device = 0, rank = 0, world = 1;
    int64_t max_items = 0;       // result entries per rank and step (max_nq * max_kk)
    int64_t max_nq = 0;          // query rows per pass
    int max_kk = 0;
    int dim = 0, kp = 0;         // 0: result exchange only (no query / bound / mean buffers)
    void *base = nullptr;        // ONE cudaMalloc, mapped into every peer
    size_t bytes = 0;
    size_t off_flags[F_NKINDS] = {};   // per kind: world lines of 128 bytes (word 0: step counter; F_TOPK word 1: overflow count)
    size_t off_done = 0, off_idx = 0, off_dist = 0, off_bounds = 0, off_sums = 0, off_qbf = 0, off_qnorm = 0, off_qerr = 0, off_qraw = 0;
    void *peer_base[EXCH_MAX_WORLD] = {};
    bool connected = false, ipc_mapped = false;
    unsigned int step = 0;       // F_TOPK
    unsigned int qstep = 0;      // F_QBF / F_CONSUMED: one id per chunk of a host-row query
    unsigned int bstep = 0;      // F_BOUND: one id per tensor pass of any collective query
    unsigned int mstep = 0;      // F_MEAN
    int64_t n_global = 0;        // rows of the whole pool (known after add)
    cudaStream_t up_stream = nullptr;    // host-to-device uploads + conversion of this rank's slices, ahead of the compute stream
    cudaStream_t bc_stream = nullptr;    // NVLink broadcasts of the slices: the next upload (PCIe) overlaps them
    cudaEvent_t ev_conv[2] = {nullptr, nullptr}, ev_bf[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};   // per buffer parity
    DevBuf<int32_t> loc_idx, pad_idx;    // this rank's exact lists of a call, before the merge
    DevBuf<double> loc_dist, pad_dist;
    DevBuf<unsigned int> gl_overflow;    // [1] overflowed second-pass lists over ALL ranks (written by the last merge of a call)
    unsigned int *h_glovf = nullptr;     // pinned
    double *h_counts = nullptr;          // pinned [world]

    PeerPtrs peers() const {
        PeerPtrs p{};
        for (int i = 0; i < world; i++) p.base[i] = static_cast<char *>(peer_base[i]);
        return p;
    }
    char *local() const { return static_cast<char *>(base); }
    const unsigned int *local_flags(int kind) const { return reinterpret_cast<const unsigned int *>(local() + off_flags[kind]); }
};

namespace {

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

// $B200KNN_VERBOSE: host-side trace of the collective protocol (one line per stage, wall-clock ms since the first line)
void ex_trace(const b200knn_exchange *ex, const char *fmt, ...) {
    static const bool on = getenv("B200KNN_VERBOSE") != nullptr;
    if (!on) return;
    static const auto t0 = std::chrono::steady_clock::now();
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    fprintf(stderr, "[b200knn %8.2f ms rank %d] %s\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), ex->rank, buf);
}

int ex_create(int device, int rank, int world, int dim, int64_t max_nq, int max_kk, b200knn_exchange **out) {
    if (!out) return fail(B200KNN_EINVAL, "out is NULL");
    *out = nullptr;
    if (world < 1 || world > EXCH_MAX_WORLD || rank < 0 || rank >= world || max_nq <= 0 || max_kk <= 0 || dim < 0)
        return fail(B200KNN_EINVAL, "bad rank/world/size (world <= %d)", EXCH_MAX_WORLD);
    CU_TRY(cudaSetDevice(device));
    b200knn_exchange *ex = new (std::nothrow) b200knn_exchange();
    if (!ex) return fail(B200KNN_ENOMEM, "out of host memory");
    ex->device = device;
    ex->rank = rank;
    ex->world = world;
    if (dim > 0) max_nq = std::max<int64_t>(max_nq, BM * 2);      // a pass is at least one 256-row query tile
    ex->max_nq = max_nq;
    ex->max_kk = max_kk;
    ex->dim = dim;
    ex->kp = (dim + 7) / 8 * 8;
    ex->max_items = (max_nq * max_kk + 3) / 4 * 4;
    size_t o = 0;
    for (int k = 0; k < F_NKINDS; k++) { ex->off_flags[k] = o; o += static_cast<size_t>(world) * 128; }
    ex->off_done = o; o += 128;
    const size_t zero_bytes = o;
    ex->off_idx = o = align256(o);
    o += static_cast<size_t>(2) * world * ex->max_items * sizeof(int32_t);
    ex->off_dist = o = align256(o);
    o += static_cast<size_t>(2) * world * ex->max_items * sizeof(double);
    if (dim > 0) {
        ex->off_bounds = o = align256(o);
        o += static_cast<size_t>(2) * world * max_nq * sizeof(float);
        ex->off_sums = o = align256(o);
        o += static_cast<size_t>(2) * world * (dim + 1) * sizeof(double);
        ex->off_qnorm = o = align256(o);
        o += static_cast<size_t>(2) * max_nq * sizeof(float);
        ex->off_qerr = o = align256(o);
        o += static_cast<size_t>(2) * max_nq * sizeof(float);
        ex->off_qbf = o = align256(o);
        o += static_cast<size_t>(2) * max_nq * ex->kp * sizeof(__nv_bfloat16);
        ex->off_qraw = o = align256(o);
        o += static_cast<size_t>(2) * max_nq * dim * sizeof(double);
    }
    ex->bytes = align256(o);
    cudaError_t e = cudaMalloc(&ex->base, ex->bytes);
    if (e != cudaSuccess) {
        const size_t b = ex->bytes;
        delete ex;
        return fail(B200KNN_ENOMEM, "cudaMalloc(%zu) for the exchange buffer failed: %s", b, cudaGetErrorString(e));
    }
    cudaMemset(ex->base, 0, zero_bytes);
    cudaDeviceSynchronize();   // flags are zero before any peer can map the buffer and publish into it
    ex->peer_base[rank] = ex->base;
    if (world == 1) ex->connected = true;
    if (dim > 0) {
        if (cudaStreamCreateWithFlags(&ex->up_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&ex->bc_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaMallocHost(reinterpret_cast<void **>(&ex->h_glovf), sizeof(unsigned int)) != cudaSuccess ||
            cudaMallocHost(reinterpret_cast<void **>(&ex->h_counts), sizeof(double) * EXCH_MAX_WORLD) != cudaSuccess ||
            ex->gl_overflow.ensure(1) != B200KNN_OK) {
            cudaFree(ex->base);
            delete ex;
            return fail(B200KNN_ECUDA, "creating the exchange's stream / pinned words failed");
        }
        cudaMemset(ex->gl_overflow.p, 0, sizeof(unsigned int));
        for (int i = 0; i < 2; i++) {
            cudaEventCreateWithFlags(&ex->ev_conv[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&ex->ev_bf[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&ex->ev_consumed[i], cudaEventDisableTiming);
        }
    }
    *out = ex;
    return B200KNN_OK;
}

// all-gather by peer stores + k-way merge (two launches, asynchronous on `st`)
int ex_allgather_merge(b200knn_exchange *ex, const int32_t *d_idx, const double *d_dist, int64_t nq, int kk, int32_t *d_out_idx,
                       double *d_out_dist, cudaStream_t st, const int *local_overflow, unsigned int *global_overflow) {
    const int64_t items = nq * kk;
    if (items <= 0 || items > ex->max_items)
        return fail(B200KNN_EINVAL, "nq*kk = %lld exceeds the exchange capacity %lld", (long long)items, (long long)ex->max_items);
    ex->step++;
    ExchPeers peers{};
    for (int p = 0; p < ex->world; p++) {
        char *b = static_cast<char *>(ex->peer_base[p]);
        peers.flags[p] = reinterpret_cast<unsigned int *>(b + ex->off_flags[F_TOPK]);
        peers.idx[p] = reinterpret_cast<int32_t *>(b + ex->off_idx);
        peers.dist[p] = reinterpret_cast<double *>(b + ex->off_dist);
    }
    char *lb = ex->local();
    const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(64, (items + 255) / 256)));
    publish_topk_kernel<<<blocks, 256, 0, st>>>(d_idx, d_dist, items, ex->max_items, ex->rank, ex->world, ex->step, peers,
                                                reinterpret_cast<unsigned int *>(lb + ex->off_done), local_overflow);
    CU_TRY(cudaGetLastError());
    merge_wait_kernel<<<static_cast<unsigned>((nq + 127) / 128), 128, 0, st>>>(
        reinterpret_cast<const int32_t *>(lb + ex->off_idx), reinterpret_cast<const double *>(lb + ex->off_dist),
        ex->local_flags(F_TOPK), ex->world, ex->step, ex->max_items, nq, kk, d_out_idx, d_out_dist, global_overflow);
    CU_TRY(cudaGetLastError());
    return B200KNN_OK;
}

int ex_check_pair(const b200knn_exchange *ex, const Shard &s, int dim, const char *what) {
    if (!ex->connected) return fail(B200KNN_ESTATE, "exchange is not connected to its peers");
    if (ex->dim <= 0) return fail(B200KNN_ESTATE, "%s needs an exchange created with b200knn_exchange_create_for_queries", what);
    if (ex->dim != dim) return fail(B200KNN_EINVAL, "exchange dim %d != index dim %d", ex->dim, dim);
    if (s.device != ex->device) return fail(B200KNN_EINVAL, "exchange lives on device %d, the index shard on device %d", ex->device, s.device);
    return B200KNN_OK;
}

// ---- add: the rows of this rank's shard are already attached (s.x_raw); column sums -> global means -> convert ----
int ex_finish_add(b200knn_exchange *ex, Shard &s, int dim, int kp) {
    const int64_t rows = s.n;
    TRY(s.col_mean.ensure(dim));
    CU_TRY(cudaMemsetAsync(s.col_mean.p, 0, static_cast<size_t>(dim) * sizeof(double), s.stream));
    dim3 grid(static_cast<unsigned>((rows + 255) / 256), static_cast<unsigned>((dim + 255) / 256));
    s.prof_begin(K_CONVERT);
    if (s.x_dtype == B200KNN_F64) colsum_kernel<double><<<grid, 256, 0, s.stream>>>(static_cast<const double *>(s.x_raw), rows, s.ld_x, dim, s.col_mean.p);
    else colsum_kernel<float><<<grid, 256, 0, s.stream>>>(static_cast<const float *>(s.x_raw), rows, s.ld_x, dim, s.col_mean.p);
    s.prof_end();
    CU_TRY(cudaGetLastError());
    ex->mstep++;
    const size_t par_off = ex->off_sums + static_cast<size_t>(ex->mstep & 1u) * ex->world * (dim + 1) * sizeof(double);
    s.stats.kernel_launches += 4;
    publish_colsum_kernel<<<std::max(1, std::min(64, (dim + 256) / 256)), 256, 0, s.stream>>>(s.col_mean.p, static_cast<double>(rows), dim, ex->peers(),
                                                                                          ex->world, ex->rank, par_off);
    raise_flags_kernel<<<1, 32, 0, s.stream>>>(ex->peers(), ex->world, ex->off_flags[F_MEAN], ex->rank, ex->mstep);
    wait_flags_kernel<<<1, 32, 0, s.stream>>>(ex->local_flags(F_MEAN), ex->world, ex->mstep, F_MEAN);
    const double *sums = reinterpret_cast<const double *>(ex->local() + par_off);
    global_mean_kernel<<<(dim + 255) / 256, 256, 0, s.stream>>>(sums, ex->world, dim, s.col_mean.p);
    CU_TRY(cudaGetLastError());
    s.centered = s.use_centering;
    TRY(s.launch_convert(s.x_raw, s.x_dtype, rows, s.ld_x, dim, kp, s.x_bf.p, s.xnorm_bf.p, s.x_err.p, s.scalars.p));
    TRY(s.convert_pool_tier(dim, kp));
    for (int r = 0; r < ex->world; r++)
        CU_TRY(cudaMemcpyAsync(ex->h_counts + r, sums + static_cast<size_t>(r) * (dim + 1) + dim, sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CU_TRY(cudaStreamSynchronize(s.stream));
    ex->n_global = 0;
    for (int r = 0; r < ex->world; r++) ex->n_global += static_cast<int64_t>(ex->h_counts[r]);
    return B200KNN_OK;
}

// chunks of a call: (first row, rows); one chunk = one group of query tiles of the per-shard kernel (a round that keeps
// every SM busy), bounded by the exchange's capacity; the ragged remainder goes FIRST (the only upload nothing hides)
int ex_chunks_from_group(int64_t nq, int64_t group_rows, int64_t cap_rows, std::vector<std::pair<int64_t, int64_t>> &chunks);
int ex_chunks(const Shard &s, int64_t nq, int kp, int64_t cap_rows, std::vector<std::pair<int64_t, int64_t>> &chunks) {
    Shard::Sched sch;
    TRY(s.plan(sch, nq, kp, 64));
    return ex_chunks_from_group(nq, static_cast<int64_t>(sch.qg) * BM * sch.cg, cap_rows, chunks);
}
// pure host arithmetic (also reachable through b200knn_debug_chunks for the CPU tests)
int ex_chunks_from_group(int64_t nq, int64_t group_rows, int64_t cap_rows, std::vector<std::pair<int64_t, int64_t>> &chunks) {
    cap_rows = std::max<int64_t>(BM * 2, cap_rows / (BM * 2) * (BM * 2));
    group_rows = std::max<int64_t>(BM * 2, std::min(group_rows, cap_rows));
    // Every chunk ends in two rank-wide exchanges (bounds, lists), i.e. it costs the skew between the ranks twice.  Long rows
    // have small groups (8 query tiles at d = 49152: 15 chunks of 19 ms each, ~2 ms of skew per exchange measured at 8 GPUs):
    // a chunk is then several groups, about a quarter of the call.
    if (nq / 4 > group_rows) group_rows = std::min(cap_rows / group_rows * group_rows, (nq / 4 + group_rows - 1) / group_rows * group_rows);
    if (nq <= std::min(cap_rows, group_rows + group_rows / 4)) {
        chunks.emplace_back(0, nq);
        return B200KNN_OK;
    }
    // The ragged remainder first (the only upload nothing hides), then whole groups.  (A ramp of growing chunks was tried
    // at 8 ranks to start the tensor cores earlier: every extra chunk costs two rank-wide exchanges and a second-pass
    // sweep, and the small chunks ran at half efficiency next to the uploads — 8.4 ms against 7.7 without it.)
    const int64_t rem = nq % group_rows;
    int64_t q0 = 0;
    auto push = [&](int64_t rows) { if (rows > 0) { chunks.emplace_back(q0, rows); q0 += rows; } };
    push(rem);
    while (q0 < nq) push(std::min(group_rows, nq - q0));
    return B200KNN_OK;
}

struct ExCall {       // geometry of one collective query call
    int kk_g = 0, kk_l = 0;
    std::vector<std::pair<int64_t, int64_t>> chunks;
};

int ex_begin_query(b200knn_exchange *ex, Shard &s, int64_t nq, int k, unsigned flags, ExCall &c) {
    c.kk_g = static_cast<int>(std::min<int64_t>(k, ex->n_global));
    c.kk_l = static_cast<int>(std::min<int64_t>(k, s.n));
    if (c.kk_g > ex->max_kk) return fail(B200KNN_EINVAL, "k = %d exceeds the exchange's max_kk %d", c.kk_g, ex->max_kk);
    CU_TRY(cudaSetDevice(s.device));
    TRY(ex->loc_idx.ensure(static_cast<size_t>(nq) * c.kk_l));
    TRY(ex->loc_dist.ensure(static_cast<size_t>(nq) * c.kk_l));
    if (c.kk_l < c.kk_g) {
        TRY(ex->pad_idx.ensure(static_cast<size_t>(nq) * c.kk_g));
        TRY(ex->pad_dist.ensure(static_cast<size_t>(nq) * c.kk_g));
    }
    TRY(s.begin_call(nq));
    (void)flags;
    return B200KNN_OK;
}

Shard::ShardHook ex_hook(b200knn_exchange *ex, int kk_g) {
    Shard::ShardHook h{};
    h.peers = ex->peers();
    h.world = ex->world;
    h.rank = ex->rank;
    h.bounds_off = ex->off_bounds;
    h.bound_flag_off = ex->off_flags[F_BOUND];
    h.bounds_stride = ex->max_nq;
    h.bound_step = ++ex->bstep;
    h.local_base = ex->local();
    h.kk_global = kk_g;
    return h;
}

// exchange + merge of one chunk's exact local lists (padded to kk_g when the shard holds fewer rows)
int ex_merge_chunk(b200knn_exchange *ex, Shard &s, const ExCall &c, int64_t q0, int64_t cq, int32_t *d_out_idx, double *d_out_dist, bool last) {
    const int32_t *src_i = ex->loc_idx.p + q0 * c.kk_l;
    const double *src_d = ex->loc_dist.p + q0 * c.kk_l;
    if (c.kk_l < c.kk_g) {
        s.stats.kernel_launches++;
        pad_topk_kernel<<<static_cast<unsigned>(std::min<int64_t>(s.num_sms * 4, (cq * c.kk_g + 255) / 256)), 256, 0, s.stream>>>(
            src_i, src_d, cq, c.kk_l, c.kk_g, ex->pad_idx.p + q0 * c.kk_g, ex->pad_dist.p + q0 * c.kk_g);
        CU_TRY(cudaGetLastError());
        src_i = ex->pad_idx.p + q0 * c.kk_g;
        src_d = ex->pad_dist.p + q0 * c.kk_g;
    }
    s.stats.kernel_launches += 2;
    return ex_allgather_merge(ex, src_i, src_d, cq, c.kk_g, d_out_idx + q0 * c.kk_g, d_out_dist + q0 * c.kk_g, s.stream,
                              last ? reinterpret_cast<const int *>(s.scalars.p + 5) : nullptr, last ? ex->gl_overflow.p : nullptr);
}

// After the call's one synchronisation: some rank's second-pass lists overflowed (pathological data) -> every rank answers
// its own overflowed queries with the exact scan and the exchange is repeated for every chunk.  fix() is the rank-local
// scan (device- or host-resident query rows).
template <typename Fix>
int ex_finish_query(b200knn_exchange *ex, Shard &s, const ExCall &c, int32_t *d_out_idx, double *d_out_dist, Fix fix) {
    if (*ex->h_glovf == 0) return B200KNN_OK;
    TRY(fix(*s.h_count));
    CU_TRY(cudaMemsetAsync(s.scalars.p + 5, 0, sizeof(unsigned int), s.stream));     // the repeated exchange reports a clean state
    for (size_t i = 0; i < c.chunks.size(); i++)
        TRY(ex_merge_chunk(ex, s, c, c.chunks[i].first, c.chunks[i].second, d_out_idx, d_out_dist, i + 1 == c.chunks.size()));
    return B200KNN_OK;
}

// ---- query, rows replicated in every rank's HBM ----
int ex_query_device(b200knn_exchange *ex, Shard &s, int dim, int kp, const void *d_query, int dtype, int64_t nq, int64_t ld, int k, unsigned flags,
                    int32_t *d_out_idx, double *d_out_dist, int *out_kk) {
    ExCall c;
    TRY(ex_begin_query(ex, s, nq, k, flags, c));
    if (out_kk) *out_kk = c.kk_g;
    const size_t esz = dtype == B200KNN_F64 ? 8 : 4;
    const int64_t step_rows = std::min<int64_t>(ex->max_nq, QUERY_CHUNK);
    for (int64_t q0 = 0; q0 < nq; q0 += step_rows) c.chunks.emplace_back(q0, std::min(step_rows, nq - q0));
    for (auto &ch : c.chunks) TRY(s.reserve_pass(ch.second, kp, c.kk_l, ex->world > 1 && s.tier == 0));      // no (re)allocation once flag-waiting kernels are in flight
    const int W = ex->world, R = ex->rank;
    char *lb = ex->local();
    const bool tensor_path = c.kk_l <= 32 && !(flags & B200KNN_FLAG_FORCE_SCAN);
    for (size_t i = 0; i < c.chunks.size(); i++) {
        const int64_t q0 = c.chunks[i].first, cq = c.chunks[i].second;
        const char *qsrc = static_cast<const char *>(d_query) + static_cast<size_t>(q0) * ld * esz;
        const Shard::ShardHook hook = ex_hook(ex, c.kk_g);
        if (W > 1 && tensor_path && s.tier == 0) {     // (other precision tiers: every rank converts all rows itself, below)
            // The rows are replicated, the conversion need not be: every rank converts 1/world of the chunk to BF16 + norms
            // and broadcasts that slice by peer stores (same buffers, flags and step counter as the host-row protocol; all
            // on the compute stream here — there is no upload to overlap).
            const unsigned int step = ++ex->qstep;
            const unsigned par = step & 1u;
            __nv_bfloat16 *qb = reinterpret_cast<__nv_bfloat16 *>(lb + ex->off_qbf) + static_cast<size_t>(par) * ex->max_nq * kp;
            float *qn = reinterpret_cast<float *>(lb + ex->off_qnorm) + static_cast<size_t>(par) * ex->max_nq;
            float *qe = reinterpret_cast<float *>(lb + ex->off_qerr) + static_cast<size_t>(par) * ex->max_nq;
            const int64_t slice = (cq + W - 1) / W;
            const int64_t a = std::min(cq, slice * R), b = std::min(cq, slice * (R + 1));
            s.stats.kernel_launches += 4;
            if (step > 2) wait_flags_kernel<<<1, 32, 0, s.stream>>>(ex->local_flags(F_CONSUMED), W, step - 2, F_CONSUMED);
            if (b > a) {
                TRY(s.launch_convert(qsrc + static_cast<size_t>(a) * ld * esz, dtype, b - a, ld, dim, kp, qb + a * kp, qn + a, qe + a, s.scalars.p + 2));
                BcastParams bp{};
                bp.peers = ex->peers(); bp.world = W; bp.rank = R; bp.nseg = 3;
                bp.seg[0] = {static_cast<size_t>(reinterpret_cast<char *>(qb + a * kp) - lb), static_cast<size_t>(b - a) * kp * sizeof(__nv_bfloat16)};
                bp.seg[1] = {static_cast<size_t>(reinterpret_cast<char *>(qn + a) - lb), static_cast<size_t>(b - a) * sizeof(float)};
                bp.seg[2] = {static_cast<size_t>(reinterpret_cast<char *>(qe + a) - lb), static_cast<size_t>(b - a) * sizeof(float)};
                broadcast_segments_kernel<<<static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>(148, ((bp.seg[0].bytes >> 4) + 2047) / 2048))), 512, 0, s.stream>>>(bp);
            }
            raise_flags_kernel<<<1, 32, 0, s.stream>>>(ex->peers(), W, ex->off_flags[F_QBF], R, step);
            s.prof_begin(K_WAIT);
            wait_flags_kernel<<<1, 32, 0, s.stream>>>(ex->local_flags(F_QBF), W, step, F_QBF);
            s.prof_end();
            CU_TRY(cudaGetLastError());
            const QuerySide pre{qb, qn, qe};
            TRY(s.query_device(qsrc, dtype, cq, ld, dim, kp, k, flags, ex->loc_idx.p + q0 * c.kk_l, ex->loc_dist.p + q0 * c.kk_l, &pre,
                               static_cast<int>(q0), &hook));
            TRY(ex_merge_chunk(ex, s, c, q0, cq, d_out_idx, d_out_dist, i + 1 == c.chunks.size()));
            raise_flags_kernel<<<1, 32, 0, s.stream>>>(ex->peers(), W, ex->off_flags[F_CONSUMED], R, step);
            CU_TRY(cudaGetLastError());
            continue;
        }
        TRY(s.query_device(qsrc, dtype, cq, ld, dim, kp, k, flags, ex->loc_idx.p + q0 * c.kk_l, ex->loc_dist.p + q0 * c.kk_l, nullptr,
                           static_cast<int>(q0), &hook));
        TRY(ex_merge_chunk(ex, s, c, q0, cq, d_out_idx, d_out_dist, i + 1 == c.chunks.size()));
    }
    TRY(s.enqueue_overflow_readback());
    CU_TRY(cudaMemcpyAsync(ex->h_glovf, ex->gl_overflow.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, s.stream));
    CU_TRY(cudaStreamSynchronize(s.stream));
    return ex_finish_query(ex, s, c, d_out_idx, d_out_dist, [&](int nov) {
        return s.fix_overflow_device(d_query, dtype, ld, nov, dim, c.kk_l, flags, ex->loc_idx.p, ex->loc_dist.p);
    });
}

// ---- query, rows in HOST memory (every rank sees the same matrix; each uploads 1/world of every chunk) ----
// Phase 1 of a host-row query: geometry and EVERY allocation, nothing collective.  A rank that fails here has not touched
// the protocol yet (a single-process group checks all its ranks before any of them enters phase 2).
int ex_query_host_prepare(b200knn_exchange *ex, Shard &s, int kp, int64_t nq, int k, unsigned flags, ExCall &c) {
    TRY(ex_begin_query(ex, s, nq, k, flags, c));
    TRY(ex_chunks(s, nq, kp, std::min<int64_t>(ex->max_nq, QUERY_CHUNK), c.chunks));
    TRY(s.out_idx.ensure(static_cast<size_t>(nq) * c.kk_g));
    TRY(s.out_dist.ensure(static_cast<size_t>(nq) * c.kk_g));
    for (auto &ch : c.chunks) TRY(s.reserve_pass(ch.second, kp, c.kk_l, true));       // no (re)allocation once flag-waiting kernels are in flight
    TRY(s.reserve_upload_ring());
    return B200KNN_OK;
}

// $B200KNN_VERBOSE: device-side timeline of a call (CUDA events on both streams, printed after the call's synchronisation)
struct ExTimeline {
    bool on = false;
    cudaEvent_t t0 = nullptr;
    struct Mark { cudaEvent_t ev; const char *what; long long chunk; };
    std::vector<Mark> marks;
    std::mutex mu;
    void begin(cudaStream_t st) {
        on = getenv("B200KNN_VERBOSE") != nullptr;
        if (!on) return;
        cudaEventCreate(&t0);
        cudaEventRecord(t0, st);
    }
    void mark(cudaStream_t st, const char *what, long long chunk) {
        if (!on) return;
        Mark m{nullptr, what, chunk};
        cudaEventCreate(&m.ev);
        cudaEventRecord(m.ev, st);
        std::lock_guard<std::mutex> lock(mu);
        marks.push_back(m);
    }
    void print(const b200knn_exchange *ex) {
        if (!on) return;
        for (auto &m : marks) {
            float ms = 0.f;
            cudaEventSynchronize(m.ev);
            cudaEventElapsedTime(&ms, t0, m.ev);
            if (ex->rank == 0 || ex->rank == ex->world - 1) fprintf(stderr, "[b200knn timeline rank %d] %8.3f ms  chunk %lld  %s\n", ex->rank, ms, m.chunk, m.what);
            cudaEventDestroy(m.ev);
        }
        cudaEventDestroy(t0);
        marks.clear();
    }
};

// Phase 2: the pipeline.  h_out_* may be NULL (a rank of a single-process group whose copy of the result nobody needs).
int ex_query_host_run(b200knn_exchange *ex, Shard &s, int dim, int kp, const void *h_query, int dtype, int64_t nq, int64_t ld, int k, unsigned flags,
                      int32_t *h_out_idx, double *h_out_dist, ExCall &c) {
    CU_TRY(cudaSetDevice(s.device));
    const size_t esz = dtype == B200KNN_F64 ? 8 : 4;
    const size_t row_bytes = static_cast<size_t>(dim) * esz;
    const int64_t nchunks = static_cast<int64_t>(c.chunks.size());
    const unsigned int step0 = ex->qstep;
    ExTimeline tl;
    tl.begin(s.stream);
    ex->qstep += static_cast<unsigned int>(nchunks);
    const int W = ex->world, R = ex->rank;
    char *lb = ex->local();
    auto q_raw = [&](char *b, unsigned par) { return b + ex->off_qraw + static_cast<size_t>(par) * ex->max_nq * dim * sizeof(double); };
    auto q_bf = [&](char *b, unsigned par) { return reinterpret_cast<__nv_bfloat16 *>(b + ex->off_qbf) + static_cast<size_t>(par) * ex->max_nq * kp; };
    auto q_norm = [&](char *b, unsigned par) { return reinterpret_cast<float *>(b + ex->off_qnorm) + static_cast<size_t>(par) * ex->max_nq; };
    auto q_err = [&](char *b, unsigned par) { return reinterpret_cast<float *>(b + ex->off_qerr) + static_cast<size_t>(par) * ex->max_nq; };

    // ---- uploader: its own host thread (a pageable source keeps it busy with memcpy) and stream ----
    // Ordering rules.  Across ranks: flags in peer memory.  Between this rank's two streams: CUDA events — a kernel that
    // spins on a flag is only ever launched when everything it waits for is guaranteed to run without it finishing
    // (the own-rank part of every dependency has completed, or is ordered by an event, before the spin starts).
    // uploaded / consumed count chunks whose events have been RECORDED (host-side hand-over between the two threads).
    std::atomic<int64_t> uploaded{0}, consumed{0};
    std::atomic<int> up_rc{B200KNN_OK};
    std::string up_err;
    std::thread uploader([&]() {
        cudaSetDevice(s.device);
        cudaStream_t up = ex->up_stream;
        for (int64_t i = 0; i < nchunks; i++) {
            const unsigned int step = step0 + static_cast<unsigned int>(i) + 1;
            const unsigned par = step & 1u;
            const int64_t q0 = c.chunks[i].first, cq = c.chunks[i].second;
            const int64_t slice = (cq + W - 1) / W;                      // rank r uploads rows [r * slice, (r + 1) * slice) of the chunk
            const int64_t a = std::min(cq, slice * R), b = std::min(cq, slice * (R + 1));
            int rc = B200KNN_OK;
            ex_trace(ex, "uploader: chunk %lld step %u rows [%lld,%lld) slice [%lld,%lld)", (long long)i, step, (long long)q0, (long long)(q0 + cq), (long long)a, (long long)b);
            // the buffers of this parity are free once EVERY rank (this one included) has consumed chunk step-2
            if (i >= 2) {          // own compute of chunk i-2: event; the peers': their flags
                while (consumed.load(std::memory_order_acquire) < i - 1) std::this_thread::yield();
                if (cudaStreamWaitEvent(up, ex->ev_consumed[par], 0) != cudaSuccess) rc = fail(B200KNN_ECUDA, "cudaStreamWaitEvent failed");
            }
            if (step > 2) wait_flags_kernel<<<1, 32, 0, up>>>(ex->local_flags(F_CONSUMED), W, step - 2, F_CONSUMED);
            cudaStream_t bc = ex->bc_stream;
            if (b > a && rc == B200KNN_OK) {
                char *raw = q_raw(lb, par) + static_cast<size_t>(a) * row_bytes;
                tl.mark(up, "up: start (consumed wait passed)", i);
                rc = s.upload_rows(raw, static_cast<const char *>(h_query) + static_cast<size_t>(q0 + a) * ld * esz, b - a, row_bytes,
                                   static_cast<size_t>(ld) * esz, up);
                tl.mark(up, "up: H2D done", i);
                if (rc == B200KNN_OK)
                    rc = s.launch_convert(raw, dtype, b - a, dim, dim, kp, q_bf(lb, par) + a * kp, q_norm(lb, par) + a, q_err(lb, par) + a,
                                          s.scalars.p + 2, up);
            }
            // the broadcasts (NVLink copy engines) run on their own stream, so that the next chunk's upload (PCIe) overlaps them
            if (rc == B200KNN_OK && (cudaEventRecord(ex->ev_conv[par], up) != cudaSuccess || cudaStreamWaitEvent(bc, ex->ev_conv[par], 0) != cudaSuccess))
                rc = fail(B200KNN_ECUDA, "event hand-over to the broadcast stream failed");
            if (rc == B200KNN_OK) {
                // first what the tensor pass needs (BF16 rows, norms, rounding errors) ...
                if (b > a && W > 1) {
                    BcastParams bp{};
                    bp.peers = ex->peers(); bp.world = W; bp.rank = R; bp.nseg = 3;
                    bp.seg[0] = {static_cast<size_t>(reinterpret_cast<char *>(q_bf(lb, par) + a * kp) - lb), static_cast<size_t>(b - a) * kp * sizeof(__nv_bfloat16)};
                    bp.seg[1] = {static_cast<size_t>(reinterpret_cast<char *>(q_norm(lb, par) + a) - lb), static_cast<size_t>(b - a) * sizeof(float)};
                    bp.seg[2] = {static_cast<size_t>(reinterpret_cast<char *>(q_err(lb, par) + a) - lb), static_cast<size_t>(b - a) * sizeof(float)};
                    const size_t units = bp.seg[0].bytes >> 4;
                    s.stats.kernel_launches++;
                    broadcast_segments_kernel<<<static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>(32, (units + 2047) / 2048))), 512, 0, bc>>>(bp);
                }
                raise_flags_kernel<<<1, 32, 0, bc>>>(ex->peers(), W, ex->off_flags[F_QBF], R, step);
                if (cudaEventRecord(ex->ev_bf[par], bc) != cudaSuccess) rc = fail(B200KNN_ECUDA, "cudaEventRecord failed");
                tl.mark(bc, "bc: BF16 slice broadcast", i);
                if (cudaGetLastError() != cudaSuccess) rc = fail(B200KNN_ECUDA, "broadcast / flag kernel launch failed");
                // (the original rows stay where they were uploaded: the exact re-rank of any rank reads them from here)
            }
            if (rc != B200KNN_OK) {
                up_err = g_last_error;      // thread-local in the uploader: hand it over
                up_rc.store(rc);
                uploaded.store(nchunks, std::memory_order_release);      // never leave the compute thread waiting
                ex_trace(ex, "uploader: FAILED at chunk %lld: %s", (long long)i, up_err.c_str());
                return;
            }
            uploaded.store(i + 1, std::memory_order_release);
            ex_trace(ex, "uploader: chunk %lld enqueued", (long long)i);
        }
    });
    // ---- compute: this thread, the shard's stream; ordered against the uploads by the flags alone ----
    int rc_main = B200KNN_OK;
    for (int64_t i = 0; i < nchunks && rc_main == B200KNN_OK; i++) {
        const unsigned int step = step0 + static_cast<unsigned int>(i) + 1;
        const unsigned par = step & 1u;
        const int64_t q0 = c.chunks[i].first, cq = c.chunks[i].second;
        while (uploaded.load(std::memory_order_acquire) < i + 1) std::this_thread::yield();
        if (up_rc.load() != B200KNN_OK) break;
        // this rank's slice: event (it has been converted and sent before the spin below starts); the peers': their flags
        if (cudaStreamWaitEvent(s.stream, ex->ev_bf[par], 0) != cudaSuccess) { rc_main = fail(B200KNN_ECUDA, "cudaStreamWaitEvent failed"); break; }
        wait_flags_kernel<<<1, 32, 0, s.stream>>>(ex->local_flags(F_QBF), W, step, F_QBF);      // every rank's BF16 slice is here
        tl.mark(s.stream, "compute: all BF16 slices here", i);
        const QuerySide pre{q_bf(lb, par), q_norm(lb, par), q_err(lb, par)};
        Shard::ShardHook hook = ex_hook(ex, c.kk_g);
        // original rows: read from the rank that uploaded them (its QBF flag also says its slice of them is in place)
        for (int r = 0; r < W; r++) hook.qpull.base[r] = static_cast<const char *>(ex->peer_base[r]);
        hook.qpull.off = static_cast<size_t>(q_raw(lb, par) - lb);
        hook.qpull.slice_rows = static_cast<int>((cq + W - 1) / W);
        s.stats.kernel_launches += 2;
        if (W > 1 && (c.kk_l > 32 || (flags & B200KNN_FLAG_FORCE_SCAN))) {      // the exact scan wants the whole chunk's rows locally: pull them
            s.stats.kernel_launches++;
            pull_rows_kernel<<<s.num_sms * 2, 256, 0, s.stream>>>(hook.qpull, R, cq, row_bytes, q_raw(lb, par));
        }
        rc_main = s.query_device(q_raw(lb, par), dtype, cq, dim, dim, kp, k, flags, ex->loc_idx.p + q0 * c.kk_l, ex->loc_dist.p + q0 * c.kk_l, &pre,
                                 static_cast<int>(q0), &hook);
        tl.mark(s.stream, "compute: tensor pass + re-rank + second pass done", i);
        if (rc_main == B200KNN_OK) rc_main = ex_merge_chunk(ex, s, c, q0, cq, s.out_idx.p, s.out_dist.p, i + 1 == nchunks);
        tl.mark(s.stream, "compute: lists exchanged + merged", i);
        // every reader of this parity's query buffers (re-rank, second pass) is enqueued before this flag
        raise_flags_kernel<<<1, 32, 0, s.stream>>>(ex->peers(), W, ex->off_flags[F_CONSUMED], R, step);
        if (rc_main == B200KNN_OK && cudaGetLastError() != cudaSuccess) rc_main = fail(B200KNN_ECUDA, "flag kernel launch failed");
        if (rc_main == B200KNN_OK && cudaEventRecord(ex->ev_consumed[par], s.stream) != cudaSuccess) rc_main = fail(B200KNN_ECUDA, "cudaEventRecord failed");
        consumed.store(i + 1, std::memory_order_release);
        ex_trace(ex, "compute: chunk %lld step %u enqueued (rc %d)", (long long)i, step, rc_main);
    }
    consumed.store(nchunks + 2, std::memory_order_release);   // never leave the uploader waiting
    uploader.join();
    ex_trace(ex, "call enqueued: %lld chunks, nq %lld, k %d", (long long)nchunks, (long long)nq, k);
    if (up_rc.load() != B200KNN_OK) return fail(up_rc.load(), "%s", up_err.c_str());
    if (rc_main != B200KNN_OK) return rc_main;
    auto copy_out = [&]() -> int {
        if (!h_out_idx || !h_out_dist) return B200KNN_OK;
        CU_TRY(cudaMemcpyAsync(h_out_idx, s.out_idx.p, static_cast<size_t>(nq) * c.kk_g * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
        CU_TRY(cudaMemcpyAsync(h_out_dist, s.out_dist.p, static_cast<size_t>(nq) * c.kk_g * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
        return B200KNN_OK;
    };
    TRY(copy_out());
    TRY(s.enqueue_overflow_readback());
    CU_TRY(cudaMemcpyAsync(ex->h_glovf, ex->gl_overflow.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, s.stream));
    CU_TRY(cudaStreamSynchronize(s.stream));
    ex_trace(ex, "call done: overflow local %d global %u", *s.h_count, *ex->h_glovf);
    tl.print(ex);
    if (*ex->h_glovf == 0) return B200KNN_OK;
    TRY(ex_finish_query(ex, s, c, s.out_idx.p, s.out_dist.p, [&](int nov) {
        return s.fix_overflow_host(h_query, dtype, ld, nov, dim, c.kk_l, flags, ex->loc_idx.p, ex->loc_dist.p);
    }));
    TRY(copy_out());
    CU_TRY(cudaStreamSynchronize(s.stream));
    return B200KNN_OK;
}

int ex_query_host(b200knn_exchange *ex, Shard &s, int dim, int kp, const void *h_query, int dtype, int64_t nq, int64_t ld, int k, unsigned flags,
                  int32_t *h_out_idx, double *h_out_dist, int *out_kk) {
    ExCall c;
    TRY(ex_query_host_prepare(ex, s, kp, nq, k, flags, c));
    if (out_kk) *out_kk = c.kk_g;
    return ex_query_host_run(ex, s, dim, kp, h_query, dtype, nq, ld, k, flags, h_out_idx, h_out_dist, c);
}

void ex_destroy(b200knn_exchange *ex) {
    if (!ex) return;
    cudaSetDevice(ex->device);
    cudaDeviceSynchronize();
    if (ex->ipc_mapped)
        for (int p = 0; p < ex->world; p++)
            if (p != ex->rank && ex->peer_base[p]) cudaIpcCloseMemHandle(ex->peer_base[p]);
    ex->loc_idx.release(); ex->loc_dist.release(); ex->pad_idx.release(); ex->pad_dist.release(); ex->gl_overflow.release();
    if (ex->h_glovf) cudaFreeHost(ex->h_glovf);
    if (ex->h_counts) cudaFreeHost(ex->h_counts);
    if (ex->up_stream) cudaStreamDestroy(ex->up_stream);
    if (ex->bc_stream) cudaStreamDestroy(ex->bc_stream);
    for (int i = 0; i < 2; i++) {
        if (ex->ev_conv[i]) cudaEventDestroy(ex->ev_conv[i]);
        if (ex->ev_bf[i]) cudaEventDestroy(ex->ev_bf[i]);
        if (ex->ev_consumed[i]) cudaEventDestroy(ex->ev_consumed[i]);
    }
    if (ex->base) cudaFree(ex->base);
    delete ex;
}

}  // namespace
