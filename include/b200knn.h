/*
 * b200knn — C ABI of the B200-native exact k-nearest-neighbour engine that replaces the DCI
 * native core of ningyu1991/InclusiveGAN on the IMLE matching path.
 *
 * Plain C, plain pointers and sizes; no Python, NumPy, torch or TensorFlow types.  The shared
 * library (inclusivegan_b200/libb200knn.so) is what a reference-side FFI binds; the host-side
 * mirror of the reference's Python API (inclusivegan_b200/dci.py, class DCI) calls it via ctypes.
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Each entry point cites the reference interface it replaces
 * (paths relative to the reference root, ningyu1991/InclusiveGAN):
 *     C core      dci_code/include/dci.h:76-89   dci_init / dci_add / dci_query / dci_clear / dci_reset / dci_free
 *     extension   dci_code/src/py_dci.c:311-321  _dci.new / add / query / clear / reset / get_num_points / ...
 *
 * Conventions
 *   - every function returns B200KNN_OK (0) or a negative B200KNN_E* code; b200knn_last_error()
 *     returns a thread-local human-readable message for the last failure on the calling thread
 *     (the reference's C core has no error channel: it assert()s and abort()s, dci.c:122-123).
 *   - matrices are row-major: `rows x dim` elements with a leading dimension `ld` (elements between
 *     consecutive rows, ld >= dim).  This is the NumPy layout the reference takes
 *     (py_dci.c:106-107: "py_data->data is N x D").
 *   - distances are Euclidean (sqrt of the sum of squared differences) like util.c:62-69, unless
 *     B200KNN_FLAG_SQUARED is set (metrics/precision_recall.py:20-57 works on squared L2).
 *   - results are exact: the approximation knobs of dci_query_config (dci.h:62-74) have no
 *     counterpart here.  Per query exactly kk = min(k, num_points) neighbours are produced,
 *     ascending by (distance, index); ties resolve to the lower index.
 *   - there is NO CPU fallback: with no usable CUDA device every compute entry point fails with
 *     B200KNN_ENODEVICE.
 *   - not re-entrant per handle (same as the reference, which holds no locks: SURVEY.md §8b).
 */
#ifndef B200KNN_H
#define B200KNN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200KNN_ABI_VERSION 1

/* status codes */
#define B200KNN_OK          0
#define B200KNN_EINVAL     -1   /* bad argument */
#define B200KNN_ESTATE     -2   /* e.g. add() on a non-empty index (dci.py:228-229), query on an empty one */
#define B200KNN_ENODEVICE  -3   /* no CUDA device / driver, or device is not sm_100 */
#define B200KNN_ECUDA      -4   /* a CUDA runtime / driver call failed */
#define B200KNN_ENOMEM     -5   /* host or device allocation failed */

/* element types of caller buffers */
#define B200KNN_F64 0           /* float64 — the only dtype the reference accepts (dci.py:116-117) */
#define B200KNN_F32 1           /* float32 — extension: halves PCIe bytes, skips the f64 blow-up */

/* query flags */
#define B200KNN_FLAG_SQUARED      1u   /* return squared distances */
#define B200KNN_FLAG_NO_CERTIFY   2u   /* debug: skip the exactness certificate + exact second pass */
#define B200KNN_FLAG_FORCE_SCAN   4u   /* debug: answer with the exact CUDA-core scan only (no tensor-core pass) */

typedef struct b200knn_index b200knn_index;

/* ---- life cycle --------------------------------------------------------------------------- */

/* Replaces dci_init (dci.h:76, dci.c:73-93) / _dci.new (py_dci.c:66-83).
 * dim: feature dimensionality.  n_devices / device_ids: GPUs the pool is row-sharded over
 * (n_devices <= 0: the current device only; device_ids may be NULL for 0..n_devices-1).
 * No device work happens here, so a handle can be created (and argument errors tested) on a
 * machine without a GPU; the device is touched on the first add(). */
int b200knn_create(int dim, int n_devices, const int *device_ids, b200knn_index **out);

/* Replaces dci_free (dci.h:89) / the capsule destructor py_dci_free (py_dci.c:53-64). */
int b200knn_destroy(b200knn_index *index);

/* Replaces dci_clear and dci_reset (dci.h:83-86; py_dci.c:214-259): drop the pool, keep the handle.
 * (dci_reset also redraws DCI's random projections — exact search has none.) */
int b200knn_clear(b200knn_index *index);

/* Replaces _dci.get_num_points (py_dci.c:262-272). */
int64_t b200knn_num_points(const b200knn_index *index);
int b200knn_dim(const b200knn_index *index);

/* ---- host-buffer entry points (what the DCI Python class calls) ---------------------------- */

/* Replaces dci_add (dci.h:79, dci.c:108-337) / _dci.add (py_dci.c:86-128).
 * data: HOST pointer, n x dim row-major (ld elements per row), dtype B200KNN_F64|F32.
 * The rows are copied to the device(s) (the reference instead borrows the caller's buffer,
 * py_dci.c:118-123), converted to BF16 + norms there, and row-sharded across the handle's GPUs.
 * Fails with B200KNN_ESTATE when the index already holds points (dci.py:228-229). */
int b200knn_add(b200knn_index *index, const void *data, int dtype, int64_t n, int64_t ld);

/* Replaces dci_query (dci.h:81, dci.c:788-828) / _dci.query (py_dci.c:130-211).
 * query: HOST pointer, nq x dim row-major.  out_idx (int32) / out_dist (float64): caller-allocated
 * HOST buffers of nq x kk, kk = min(k, num_points) (rectangular — the reference malloc()s ragged
 * per-query arrays the caller must free, dci.c:812-821).  Indices are row positions in the array
 * passed to add() (py_dci.c:185).  *out_kk receives kk (may be NULL). */
int b200knn_query(b200knn_index *index, const void *query, int dtype, int64_t nq, int64_t ld, int k,
                  unsigned flags, int32_t *out_idx, double *out_dist, int *out_kk);

/* Self-kNN: every pool row queried against the pool it belongs to (row i's own entry, distance 0, comes first unless
 * duplicates tie).  The reference's precision/recall metric does exactly this to get the k-th neighbour radii
 * (metrics/precision_recall.py:74-90); the rows, their BF16 copies and norms are already on the device from add(), so
 * nothing is uploaded or converted.  out_idx / out_dist: HOST, num_points x kk.
 * Multi-device handles (k <= 32): the rows of every shard, chunk by chunk, are the queries of one collective call — the
 * chunk's original rows are replicated to the other shards over NVLink, every shard converts 1/G of them, broadcasts the
 * BF16 slice and answers against its own rows; bit-identical to a single-device handle. */
int b200knn_query_self(b200knn_index *index, int k, unsigned flags, int32_t *out_idx, double *out_dist, int *out_kk);

/* Ball membership for the k-NN precision/recall metric (reference metrics/precision_recall.py:96-134, the
 * `np.any(distance_batch[..., None] <= self.D, axis=1)` test): out_member[i] = 1 iff query row i lies inside at least
 * one ball B(x_j, sqrt(radius2[j])) around a pool row (squared Euclidean distance <= radius2[j]; the reference works on
 * squared distances, precision_recall.py:32).  radius2: HOST float64 [num_points], indexed like the rows passed to
 * add().  Exact (float64 decision); the reference decides in float16.  out_member: HOST uint8 [nq]. */
int b200knn_ball_membership(b200knn_index *index, const void *query, int dtype, int64_t nq, int64_t ld, const double *radius2,
                            unsigned char *out_member);

/* ---- random projection on the device ----
 * The trainer's optional `--init-proj-dim` path multiplies every 256-row chunk of generated images and every 24-row
 * chunk of real images by a fixed Gaussian matrix in NumPy float64 on the host before handing them to DCI
 * (training/training_loop.py:205-212 builds `projector`, :362-365 and :379-381 do `np.matmul(rows.astype(float64),
 * projector)`).  These entry points take the UNPROJECTED rows (float32 as the generator emits them, or float64) and
 * do that product on the device in float64 (one sequential FMA chain per output element: independent of chunking),
 * so neither the float64 blow-up of the images nor the projected matrix crosses PCIe.
 * projector: HOST float64 [in_dim][dim] row-major (leading dimension ld), dim = the handle's dim; may be replaced at
 * any time, applies to later calls.  rows: HOST [n][in_dim].  Results exactly as add() / query() on the projected rows.
 * Multi-device handles: every device keeps a copy of the projector; add_projected projects each shard's slice of the rows
 * on that shard's device (its own PCIe link), query_projected projects on the first device and feeds the row-sharded
 * collective query. */
int b200knn_set_projector(b200knn_index *index, const double *projector, int64_t in_dim, int64_t ld);
int b200knn_add_projected(b200knn_index *index, const void *rows, int dtype, int64_t n, int64_t ld);
int b200knn_query_projected(b200knn_index *index, const void *rows, int dtype, int64_t nq, int64_t ld, int k, unsigned flags,
                            int32_t *out_idx, double *out_dist, int *out_kk);
/* Debug / parity: the projected rows themselves.  out: HOST float64 [n][dim]. */
int b200knn_project_rows(b200knn_index *index, const void *rows, int dtype, int64_t n, int64_t ld, double *out);

/* ---- device-buffer entry points (inputs already resident in HBM) ---- */

/* Launch all work of this handle on `stream` (a cudaStream_t passed as void*, NULL = the
 * library's own stream).  Lets a host that owns streams (e.g. torch) time and order the work. */
int b200knn_set_stream(b200knn_index *index, void *stream);

/* As b200knn_add, but `data` is a DEVICE pointer on the handle's device; the rows are NOT copied:
 * the index borrows the buffer until clear/destroy (the reference's ownership model, dci.h:78).
 * index_base is added to every returned index (row offset of this shard in a pool that is
 * row-sharded across processes; py_dci.c:185's data_idx_offset).
 * Multi-device handles: d_data may live on any device of the process (peer access); the handle shards the rows itself,
 * COPYING each slice into its shard over NVLink (index_base must be 0), then indexes them like b200knn_add. */
int b200knn_add_device(b200knn_index *index, const void *d_data, int dtype, int64_t n, int64_t ld,
                       int64_t index_base);

/* As b200knn_query with DEVICE query / output buffers (single-device handles; the row-sharded form is
 * b200knn_exchange_query_device); synchronises the handle's stream once, at the end.  d_out_idx: int32 nq x kk,
 * d_out_dist: float64 nq x kk. */
int b200knn_query_device(b200knn_index *index, const void *d_query, int dtype, int64_t nq, int64_t ld, int k,
                         unsigned flags, int32_t *d_out_idx, double *d_out_dist, int *out_kk);

/* k-way merge of `n_lists` per-shard results (each nq x kk, ascending) into the global nq x kk result.
 * d_idx / d_dist: DEVICE buffers laid out [n_lists][nq][kk] (what an NCCL all-gather of the local
 * results produces).  Ties -> lower index.  Runs on `stream`. */
int b200knn_merge_topk_device(const int32_t *d_idx, const double *d_dist, int n_lists, int64_t nq, int kk,
                              int32_t *d_out_idx, double *d_out_dist, void *stream);

/* ---- NVLink exchange: all-gather + merge over peer memory, one process per GPU, no NCCL -------------------
 * The multi-GPU step of the path: every rank holds an exact local top-k of its pool shard; the global
 * answer is the k-way merge of the `world` lists.  Instead of an NCCL all-gather followed by a merge, each rank
 * STORES its lists directly into every peer's gather buffer over NVLink (CUDA IPC peer mappings) and raises a step
 * flag there; the merge kernel on each rank waits for the `world` flags in its own buffer and merges.
 * Two kernel launches per step, no host synchronisation, no collective library.
 *   create  : allocate this rank's buffer (capacity max_nq * max_kk entries per rank, double-buffered).
 *   handle  : B200KNN_IPC_BYTES opaque bytes to hand to the other ranks (e.g. torch.distributed.all_gather_object).
 *   connect : all_handles = world * B200KNN_IPC_BYTES bytes in rank order (own entry ignored).
 *   allgather_merge : d_idx/d_dist local [nq][kk] (device) -> d_out_* global [nq][kk]; asynchronous on `stream`.
 *     Every rank must call it the same number of times with the same nq, kk. */
#define B200KNN_IPC_BYTES 64
typedef struct b200knn_exchange b200knn_exchange;
int b200knn_exchange_create(int device, int rank, int world, int64_t max_nq, int max_kk, b200knn_exchange **out);
int b200knn_exchange_handle(b200knn_exchange *ex, void *out_bytes);
int b200knn_exchange_connect(b200knn_exchange *ex, const void *all_handles);
int b200knn_exchange_allgather_merge(b200knn_exchange *ex, const int32_t *d_idx, const double *d_dist, int64_t nq, int kk,
                                     int32_t *d_out_idx, double *d_out_dist, void *stream);
int b200knn_exchange_destroy(b200knn_exchange *ex);

/* ---- the whole multi-GPU step behind the library (north_star: "queries are broadcast, each GPU keeps a
 * local top-k, ... a k-way merge kernel produces the global result") ----------------------------------------------
 * One rank per GPU (a process of a torchrun job, or one device of an in-process group); rank r holds pool rows
 * [index_base, index_base + n) in a single-device handle `index`.  All calls below are COLLECTIVE: every rank makes the
 * same sequence of calls with the same nq / k / flags.  Nothing but these kernels and the copy engines touches NVLink.
 *   create_for_queries : as create, plus per-rank buffers for one pass of max_nq query rows of `dim` columns (bounds,
 *                        BF16 rows + norms, original rows; double-buffered).  handle/connect as above; connect_local
 *                        wires the exchanges of ONE process together (all[] in rank order) through plain peer access.
 *   add / add_device   : b200knn_add / b200knn_add_device of this rank's shard, except that the centring vector is the
 *                        GLOBAL column mean (the shards' column sums are gathered over peer memory), so BF16 query rows
 *                        converted by one rank are valid on every rank.
 *   query_device       : d_query = ALL nq query rows, resident on every rank.  Tensor pass against the local shard; the
 *                        ranks exchange an upper bound on their k-th nearest distance per query, and the exact re-rank
 *                        evaluates only candidates that survive the minimum of the bounds (1/world of the re-rank work
 *                        per rank instead of all of it); exact local lists -> all-gather by peer stores -> k-way merge.
 *                        d_out_*: [nq][kk] on every rank, kk = min(k, rows of the WHOLE pool).
 *   query              : the same with HOST rows (every rank passes the same matrix): per chunk every rank uploads and
 *                        converts only 1/world of the rows, the copy engines broadcast the BF16 rows + norms and, behind
 *                        the tensor pass, the original rows; uploads of chunk i+1 overlap the compute of chunk i. */
int b200knn_exchange_create_for_queries(int device, int rank, int world, int dim, int64_t max_nq, int max_kk, b200knn_exchange **out);
int b200knn_exchange_connect_local(b200knn_exchange *ex, b200knn_exchange *const *all);
int b200knn_exchange_add(b200knn_exchange *ex, b200knn_index *index, const void *data, int dtype, int64_t n, int64_t ld, int64_t index_base);
int b200knn_exchange_add_device(b200knn_exchange *ex, b200knn_index *index, const void *d_data, int dtype, int64_t n, int64_t ld,
                                int64_t index_base);
int b200knn_exchange_query(b200knn_exchange *ex, b200knn_index *index, const void *query, int dtype, int64_t nq, int64_t ld, int k,
                           unsigned flags, int32_t *out_idx, double *out_dist, int *out_kk);
int b200knn_exchange_query_device(b200knn_exchange *ex, b200knn_index *index, const void *d_query, int dtype, int64_t nq, int64_t ld, int k,
                                  unsigned flags, int32_t *d_out_idx, double *d_out_dist, int *out_kk);

/* ---- introspection -------------------------------------------------------------------------- */

typedef struct b200knn_stats {
    int64_t kernel_launches;       /* kernels of this library launched by this handle since creation */
    int64_t queries;               /* query rows answered */
    int64_t uncertified;           /* query rows whose tensor-core shortlist could not be certified exact and
                                      were re-answered by the second (threshold-collection) tensor pass */
    double  ms_convert;            /* accumulated device time per kernel family, profiling mode only */
    double  ms_distance;
    double  ms_rerank;
    double  ms_scan;               /* second pass (gather + collection GEMM + list re-rank) and exact scan */
    int64_t distance_launches;     /* launches of the tcgen05 distance kernel counted in ms_distance */
    double  distance_flops;        /* 2*nq*n*dim summed over those launches */
    int64_t exact_scanned;         /* query rows whose second-pass list overflowed: answered by the exact CUDA-core scan */
    double  ms_wait;               /* row-sharded pools: time the stream spent waiting for the peers' flags (skew between ranks) */
} b200knn_stats;

/* ---- precision tiers of the tensor pass ----------------------------------------------------------
 * Results are exact in every tier (the exact float64 re-rank and the certificate do not change); a tier decides how
 * sharp the tensor-core scores are, i.e. how many candidates must be re-ranked and how many queries need the second
 * pass.  BF16 (default): one tcgen05 kind::f16 MMA per product.  BF16X3: every operand row is split into hi + lo BF16
 * rows and the kernel runs three K segments (hi.hi + hi.lo + lo.hi; the lo.lo term is bounded by ||q_lo|| ||x_lo||) — 3x
 * the tensor work, rounding error 2^-17 instead of 2^-9 per element: for feature spaces whose nearest-neighbour gaps
 * are far below BF16 resolution.  TF32: tcgen05 kind::tf32 on fp32 containers (half the BF16 rate, 2^-12).
 * Set before add() (or after: the pool's operands of the tier are then built at once).  The row-sharded host-row protocol
 * (b200knn_exchange_query) always runs BF16; b200knn_exchange_query_device honours the tier.  $B200KNN_PRECISION sets the
 * default (bf16 | bf16x3 | tf32). */
#define B200KNN_TIER_BF16   0
#define B200KNN_TIER_BF16X3 1
#define B200KNN_TIER_TF32   2
int b200knn_set_precision(b200knn_index *index, int tier);

/* profiling != 0: bracket every kernel with CUDA events on the launching stream; read with get_stats. */
int b200knn_set_profiling(b200knn_index *index, int profiling);
int b200knn_get_stats(b200knn_index *index, b200knn_stats *out);   /* synchronises the handle's stream(s) */
int b200knn_reset_stats(b200knn_index *index);

/* Test hook (pure host code, no device needed): the work schedule the distance kernel would run for a pool of n rows,
 * nq query rows and padded row length kp on a GPU with num_sms SMs.  cta_group 0 = automatic; wide_mode as
 * $B200KNN_WIDE (0 never / 1 when cheaper in HBM traffic / 2 always the long-row grid schedule).  items receives up to
 * `capacity` work items as 4 ints each (query tile, first pool tile, end pool tile, slot word: bits 0-15 shortlist slot,
 * 16-23 workers sharing the pool stream, 24-31 workers in round-wide lockstep or 0), row-major [rounds][workers];
 * geometry[8] receives {cta_group, workers, rounds, query tiles, pool tiles, max slots, query-tile group size, wide}. */
int b200knn_debug_plan(int64_t n, int64_t nq, int kp, int num_sms, int cta_group, int max_slots, int a_budget_mb, int wide_mode,
                       int32_t *items, int64_t capacity, int32_t *geometry);

/* Test hook (pure host code): how a host-row collective query of nq rows (b200knn_exchange_query) is cut into chunks on a rank
 * holding n pool rows, and which rows of every chunk each of `world` ranks uploads.  out receives, per chunk,
 * {first row, rows, then per rank: slice begin, slice end (chunk-relative)}; *count = number of chunks. */
int b200knn_debug_chunks(int64_t n, int64_t nq, int kp, int num_sms, int64_t cap_rows, int world, int64_t *out, int64_t capacity, int64_t *count);

/* Test hook (pure host code): how b200knn_query cuts a single-device host-row call of nq rows into chunks (an upload ramp
 * sized on the distance kernel's schedule and the host link speed; csrc/b200knn.cu plan_host_chunks).  out receives
 * {first row, rows} per chunk; *count = number of chunks.  pinned: the query rows are in page-locked host memory. */
int b200knn_debug_host_chunks(int64_t n, int64_t nq, int kp, int dim, int elem_bytes, int k, int num_sms, int pinned, int64_t *out, int64_t capacity,
                              int64_t *count);

/* Test hook: copy out the BF16-pass shortlists of the LAST tensor pass of a single-device handle (the last query
 * chunk): scores[nq][slots][C] (s~ = ||x~||^2 - 2 q~.x~ as computed on the tensor cores) and rows[nq][slots][C]
 * (shard-local pool row, -1 = empty).  *nq, *slots, *c receive the geometry; the HOST buffers must hold `capacity`
 * entries each (B200KNN_EINVAL if too small).  Used by the tests that validate the certificate's error model. */
int b200knn_debug_shortlists(b200knn_index *index, float *scores, int32_t *rows, int64_t capacity, int64_t *nq, int *slots, int *c);

/* Thread-local message for the last failing call on this thread ("" if none). */
const char *b200knn_last_error(void);
int b200knn_abi_version(void);
/* Number of usable sm_100 devices (0 on a machine without a GPU; never fails). */
int b200knn_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200KNN_H */
