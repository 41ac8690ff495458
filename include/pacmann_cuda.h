/*
 * pacmann_cuda.h -- C-ABI of libpacmann_cuda.so: the B200 (sm_100a) implementation of Pacmann's
 * data-parallel hot path.  This is the drop-in boundary: the entry points below are what the
 * reference's Go packages bind over cgo in place of their Plan-9 assembler stubs and inner loops
 * (INTEGRATION.md shows the cgo side).  Plain pointers and sizes only; no C++ or torch types.
 *
 * Reference seam each entry point replaces (paths relative to the wuwuz/Pacmann tree):
 *   pm_expand_key      expandKeyAsm / GetLongKey          pianopir/util.go:117,167-171, aes_amd64.s:87-126
 *   pm_prf_batch       PRFEvalWithLongKeyAndTag/aes128MMO pianopir/util.go:157-165,     aes_amd64.s:51-82
 *   pm_xor_slices      xorSlices / EntryXor               pianopir/util.go:173, pir.go:258-265, aes_amd64.s:133-157
 *   pm_hintgen*        PianoPIRClient.Preprocessing +     pianopir/pir.go:267-352 (hint parities),
 *                      UpdatePreprocessing, batch fan-out  pianopir/batch-pir.go:119-155
 *   pm_gather_rows     replacement values                 pianopir/pir.go:345-349
 *   pm_answer_batch*   PianoPIRServer.PrivateQuery,       pianopir/pir.go:65-88,
 *                      SimpleBatchPianoPIR.Query fan-out   pianopir/batch-pir.go:189-216
 *   pm_client_*        PianoPIRClient.{Initialization,    pianopir/pir.go:203-255, 267-352, 354-471
 *                      Preprocessing,Query} resident form
 *   pm_l2_pairs/_batch L2Dist / L2DistanceSIMD            graphann/build_graph.go:119-134, l2_distance_amd64.s:4-36
 *   pm_l2_idpairs      the L2Dist calls of robustPrune     graphann/build_graph.go:169-236
 *   pm_ip_u32_scan     InnerProduct + scan loop           graphann/l2_distance_amd64.s:39-68, graphann_test.go:268-273
 *
 * Conventions
 *   - every function returns 0 (PM_OK) or a negative PM_ERR_* code; pm_last_error() returns a
 *     thread-local message for the last failure on the calling thread.  There is no CPU fallback:
 *     without a usable CUDA device every compute entry point fails with PM_ERR_CUDA.
 *   - functions without a `_dev` suffix take HOST pointers and include the host<->device copies;
 *     `_dev` variants take DEVICE pointers (on the handle's device), enqueue on `stream`
 *     (a cudaStream_t passed as void*, NULL = the handle's own stream) and do not synchronise.
 *   - the library is re-entrant per handle; every entry point sets the CUDA device itself, so it may
 *     be called from goroutines that migrate between OS threads.  C never retains a caller pointer
 *     after the call returns; device-resident copies are owned by the opaque handle.
 *   - bit-exactness: integer/byte results are bit-identical to the reference algorithm, including
 *     xorSlices' "len%4 tail is not xored" behaviour; pm_l2_* reproduces the reference's 8-lane fp32
 *     evaluation order (no FMA contraction), so distances are bit-identical as well.
 */
#ifndef PACMANN_CUDA_H
#define PACMANN_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM_OK 0
#define PM_ERR_ARG (-1)         /* bad argument (null pointer, size mismatch, out-of-range index) */
#define PM_ERR_CUDA (-2)        /* CUDA runtime failure, or no usable device */
#define PM_ERR_UNSUPPORTED (-3) /* shape outside what the kernels implement */
#define PM_ERR_NOMEM (-4)

#define PM_NO_SKIP (-1)

/* Device-resident flat row-major table of uint64: rows[n_rows][entry_u64].  The PIR database
 * (rawDB, pianopir/pir.go:28-39) and, through its prefix of `dim` fp32 per row, the vector table
 * used for distances (private-search.go:371-397 wire format).  A [N][D] uint32 matrix with even D
 * is the same thing with entry_u64 = D/2. */
typedef struct pm_db pm_db;

const char *pm_version(void);
const char *pm_last_error(void);
int pm_device_count(int *count);
/* number of kernel launches issued by this library in this process (for accounting) */
uint64_t pm_launch_count(void);

/* Launch-time tuning knobs (tests and profiling force the launch variants the heuristics would pick at other sizes).
 * Knobs: "hg_sync" (round barrier of the hint kernel: -1 auto, 0, 1), "hg_warps" (CTA width in warps: 0 auto, 1..16),
 * "hg_ntab" (AES T-tables: 0 auto, 1, 4), "hg_tail_split" (shared last round: -1 auto, 0 off, 1 on), "hg_serpentine"
 * (0, 1), "hg_xbytes" (chunk-id bytes the hoisted PRF rounds treat as varying: 0 auto, 2, 4), "ans_split" (CTAs per
 * sub-query: 0 auto, 1..8), "hg_d2h_groups" (launch groups of pm_hintgen: 0 auto, 1..16), "search_ans_stream" (read by
 * pm_search_create: 1 = the answer kernel of a search step runs on a low-priority stream of its own, 0 = one stream).  Initial values come from the
 * environment (PM_HG_SYNC, PM_HG_WARPS, ...).  Results never depend on a knob. */
int pm_tuning_set(const char *name, int value);
int pm_tuning_get(const char *name, int *value);

int pm_db_create(const uint64_t *rows_host, uint64_t n_rows, uint64_t entry_u64, int device, pm_db **out);
int pm_db_create_empty(uint64_t n_rows, uint64_t entry_u64, int device, pm_db **out);
/* borrow an existing device allocation (16-byte aligned, on `device`); the caller keeps ownership */
int pm_db_wrap(void *device_rows, uint64_t n_rows, uint64_t entry_u64, int device, pm_db **out);
int pm_db_upload(pm_db *db, uint64_t row0, uint64_t n_rows, const uint64_t *rows_host);
int pm_db_info(const pm_db *db, uint64_t *n_rows, uint64_t *entry_u64, int *device, void **device_ptr);
int pm_db_destroy(pm_db *db);
/* blocks until everything enqueued on the handle's own streams has finished */
int pm_db_sync(pm_db *db);

/* Plain device buffers shareable between the per-GPU processes of one box (CUDA IPC).  Multi-GPU hint generation:
 * rank 0 allocates the full parity table, the other ranks map it (pm_buf_ipc_open) and pass addresses inside it as
 * pm_hint_job.parity_out to pm_hintgen_dev, so every kernel stores its shard straight into rank 0's HBM over NVLink. */
int pm_buf_alloc(uint64_t bytes, int device, void **dev_ptr);
int pm_buf_free(void *dev_ptr, int device);
int pm_buf_ipc_export(void *dev_ptr, int device, uint8_t handle[64]);
int pm_buf_ipc_open(const uint8_t handle[64], int device, void **dev_ptr);
int pm_buf_ipc_close(void *dev_ptr, int device);
/* synchronous copies / clearing of such buffers (a Go caller has no cudaMemcpy of its own) */
int pm_buf_upload(void *dev_ptr, const void *host, uint64_t bytes, int device);
int pm_buf_download(void *host, const void *dev_ptr, uint64_t bytes, int device);
int pm_buf_zero(void *dev_ptr, uint64_t bytes, int device);
/* asynchronous device-to-device copy on `stream` (either side may be a peer GPU's IPC-mapped buffer: copy engines over NVLink) */
int pm_buf_copy_dev(void *dst, const void *src, uint64_t bytes, int device, void *stream);
/* Completion flags for the peer-memory exchange.  `flags` points at (n_flags + 1) 128-byte lines in device memory
 * (usually inside the consumer's IPC-shared buffer, zeroed once): line i holds rank i's counter in its first uint32,
 * line n_flags a timeout marker.  pm_flag_signal_dev adds 1 to ONE counter (system-scope release) once everything
 * enqueued before it on `stream` has finished; pm_flag_wait_dev blocks `stream` until every one of the n_flags counters
 * has reached `target` (wrap-safe compare), or sets the marker to 1 after `timeout_ms` and lets the stream go on.
 * timeout_ms = 0: unbounded wait as stream memory operations (no polling kernel, no SM occupied -- needed when the waiting
 * GPU launches cooperative one-CTA-per-SM kernels meanwhile). */
int pm_flag_signal_dev(void *flag, int device, void *stream);
int pm_flag_wait_dev(void *flags, uint32_t n_flags, uint32_t target, uint32_t timeout_ms, int device, void *stream);

/* A1: FIPS-197 AES-128 key schedule, 11 round keys as 44 little-endian uint32 (raw 16-byte blocks). */
int pm_expand_key(const uint8_t key[16], uint32_t rk[44]);
int pm_expand_key_batch(const uint8_t *keys /* [n][16] */, uint64_t n, uint32_t *rk /* [n][44] */);
/* A2: out[i] = LE64((AES128_rk(B) xor B)[0:8]),  B = LE64((tags[i] << 35) + xs[i]) || 0^64. */
int pm_prf_batch(const uint32_t rk[44], const uint64_t *tags, const uint64_t *xs, uint64_t n, uint64_t *out);
/* A3: dst[0 : 4*(len_src/4)] ^= src[...] ; the len_src % 4 tail is left untouched (reference behaviour). */
int pm_xor_slices(uint64_t *dst, const uint64_t *src, uint64_t len_src);

/* A5/A7: one hint-generation job = (a range of) the hints of one PianoPIR instance over its DB slice.
 *   parity[i] = XOR over chunks c in [0,set_size), c != skip_i, row = c*chunk_size + (PRF(rk,tag_i,c) & (chunk_size-1)),
 *               row < n_rows  of  db[row0 + row]          (rows past n_rows are the reference's zero padding)
 * Hints are numbered as PianoPIRClient.Initialization numbers them (pir.go:220-251): hint h has
 *   tag_h  = tags ? tags[i] : h,                       h = hint_begin + i
 *   skip_h = skip_chunk ? skip_chunk[i] : (h < n_primary ? PM_NO_SKIP : (h - n_primary) / backup_group)
 * so the initial table needs no tag upload, while refreshed / custom tables pass explicit arrays.
 * chunk_size must be a power of two; entry words beyond 4*(entry_u64/4) stay zero (A3 behaviour). */
typedef struct pm_hint_job {
    uint64_t row0, n_rows;         /* DB slice of this PianoPIR instance */
    uint64_t chunk_size, set_size; /* pir.go:487-494 */
    uint32_t rk[44];               /* longKey */
    uint64_t hint_begin, n_hints;  /* hints [hint_begin, hint_begin + n_hints) */
    uint64_t n_primary;            /* primaryHintNum */
    uint64_t backup_group;         /* maxQueryPerChunk (0 = no backup hints) */
    const uint64_t *tags;          /* [n_hints] or NULL */
    const int32_t *skip_chunk;     /* [n_hints] or NULL */
    uint64_t *parity_out;          /* [n_hints][entry_u64] */
    uint16_t *offsets_out;         /* optional (device pointer, pm_hintgen_dev only; chunk_size <= 65536):
                                      [n_hints][(set_size + 7) & ~7] the offset PRF(rk, tag_i, c) & (chunk_size-1) of every
                                      (hint, chunk) -- the kernel evaluates them all anyway; the resident client keeps them
                                      as its offset index so that the online hint search needs no AES */
} pm_hint_job;

int pm_hintgen(pm_db *db, const pm_hint_job *jobs, uint64_t n_jobs);
int pm_hintgen_dev(pm_db *db, const pm_hint_job *jobs, uint64_t n_jobs, void *stream);

/* out[i] = db[row0 + idx[i]] if idx[i] < n_rows else zeros  (replacementVal, pir.go:345-349) */
int pm_gather_rows(pm_db *db, uint64_t row0, uint64_t n_rows, const uint64_t *idx, uint64_t n, uint64_t *out);

/* A6/A8: q server answers in one launch.  Sub-query i runs against the PianoPIR instance
 * (row0[i], n_rows[i], chunk_size[i], set_size[i]) with offsets[i*offsets_stride + c], c < set_size[i]:
 *   out[i] = XOR over c with idx = offsets[..c] + c*chunk_size < n_rows[i]  of  db[row0[i] + idx]. */
int pm_answer_batch(pm_db *db, const uint64_t *row0, const uint64_t *n_rows, const uint32_t *chunk_size,
                    const uint32_t *set_size, const uint32_t *offsets, uint64_t offsets_stride, uint64_t q,
                    uint64_t *out);
int pm_answer_batch_dev(pm_db *db, const uint64_t *row0, const uint64_t *n_rows, const uint32_t *chunk_size,
                        const uint32_t *set_size, const uint32_t *offsets, uint64_t offsets_stride, uint64_t q,
                        uint64_t *out, void *stream);

/* ---- GPU-resident client ------------------------------------------------------------
 * The hint tables of every sub-PIR of one SimpleBatchPianoPIR stay in HBM: Preprocessing ships no parities to
 * the host, and the online Query's hint search / set expansion / refresh (pianopir/pir.go:354-471) run on the
 * GPU around the server answer.  Results and the client state are bit-identical to the sequential reference. */
typedef struct pm_client pm_client;
typedef struct pm_client_part {
    uint64_t row0, n_rows, chunk_size, set_size;
    uint64_t n_primary;     /* primaryHintNum */
    uint64_t backup_group;  /* maxQueryPerChunk */
    uint64_t max_query_num; /* MaxQueryNum */
} pm_client_part;
typedef struct pm_client_query {
    uint32_t part;                  /* sub-PIR index */
    uint32_t kind;                  /* 0 = dummy query (pir.go:363-371), 1 = real query */
    uint64_t idx;                   /* index inside the sub-PIR (real queries) */
    uint64_t dummy_seed, dummy_ctr; /* dummy: offsets[c] = mix64(dummy_seed, dummy_ctr + c) & (chunk_size-1) */
} pm_client_query;

int pm_client_create(pm_db *db, const pm_client_part *parts, uint64_t n_parts, pm_client **out);
int pm_client_destroy(pm_client *c);
/* Initialization (pir.go:203-255) + Preprocessing (pir.go:267-352) of the listed parts, entirely on the device.
 * rk = [n][44] long keys; replacement offset of (chunk c, slot j) = mix64(repl_seed[i], c*mqpc + j) & (chunk_size-1).
 * skip_prep != 0 is DummyPreprocessing (pir.go:520-523). */
int pm_client_preprocess(pm_client *c, const uint32_t *part_ids, uint64_t n, const uint32_t *rk, const uint64_t *repl_seed,
                         int skip_prep);
/* q client queries in one call.  Queries of the same part are processed in array order (each real query consumes
 * and refreshes a hint).  out[i] = the entry (zeros for dummy / failed queries); status[i] = 0 ok, 2 query budget
 * exceeded, 3 too many queries in the chunk, 4 no hit hint (pir.go:386-419).  The caller keeps the local cache
 * (pir.go:381-383) and therefore never sends an index twice between two preprocessings. */
int pm_client_query_batch(pm_client *c, const pm_client_query *queries, uint64_t q, uint64_t *out, int32_t *status);
/* same, and dist_out[i] = L2Dist(first `dim` fp32 of out[i], query_vec) computed on the device right behind the
 * answers (the per-step L2Dist call site of SearchKNN, graphann/search.go:204), saving a second round trip */
int pm_client_query_batch_l2(pm_client *c, const pm_client_query *queries, uint64_t q, uint64_t *out, int32_t *status,
                             const float *query_vec, uint64_t dim, float *dist_out);
/* same for a call that carries the sub-queries of several searches (several independent client instances kept as
 * groups of parts of one pm_client, driven in lock step): dist_out[i] = L2Dist(out[i], query_vecs[vec_id[i]]),
 * query_vecs = [n_vecs][dim] */
int pm_client_query_batch_l2m(pm_client *c, const pm_client_query *queries, uint64_t q, uint64_t *out, int32_t *status,
                              const float *query_vecs, uint64_t n_vecs, const uint32_t *vec_id, uint64_t dim, float *dist_out);
/* copy one table of one part to the host (tests / checkpointing): 0 primaryShortTag, 1 primaryParity,
 * 2 primaryProgramPoint, 3 replacementIdx, 4 replacementVal, 5 backupShortTag, 6 backupParity, 7 QueryHistogram,
 * 8 FinishedQueryNum */
int pm_client_download(pm_client *c, uint32_t part, int table, uint64_t *out, uint64_t cap_words);

/* ---- GPU-resident lock-step search -----------------------------------------------------
 * graphann.SearchKNN's frontier (explore heap, known set, reach steps, final re-rank: graphann/search.go:114-234) and
 * SimpleBatchPianoPIR.Query's bookkeeping around the fetch (bucketing, drops, dummy padding, the client's local cache:
 * pianopir/batch-pir.go:170-248, pir.go:381-383,468) for `lanes` independent clients of one pm_client (lane l = parts
 * [l*partition_num, (l+1)*partition_num)), all resident in HBM.  A round runs one SearchKNN per listed lane, in lock
 * step: pm_search_begin, then max_step times pm_search_fetch (which only ENQUEUES the step: next batch from the neighbour
 * lists in HBM -> client prepare -> server answer -> client finish + distances), then pm_search_finish (final ranking,
 * one small copy back).  No entry, neighbour list or distance visits the host.  Every lane returns exactly what its
 * client returns searching alone (start ranking by (distance, position), container/heap explore queue, final ranking by
 * (distance, id)); pm_client_preprocess of a part also clears that part's cache here. */
typedef struct pm_search pm_search;
typedef struct pm_search_config {
    uint64_t n, dim, m;                       /* vertices, vector dimension, degree: an entry is dim f32 || m u32 */
    uint64_t max_step, parallel;              /* SearchKNN parameters (search.go:114) */
    uint64_t lanes;                           /* independent clients */
    uint64_t partition_num, partition_size;   /* SimpleBatchPianoPIR geometry (batch-pir.go:62-64) */
    uint64_t n_start;                         /* start vertices per lane (search.go:51-65: sqrt(n)) */
    uint64_t cache_entries;                   /* local-cache capacity per sub-PIR (MaxQueryNum) */
} pm_search_config;
int pm_search_create(pm_client *c, const pm_search_config *cfg, pm_search **out);
int pm_search_destroy(pm_search *s);
/* start vertices of one lane: ids [n_start], vectors [n_start][dim], neighbours [n_start][m] (plaintext copies, private-search.go:508-531) */
int pm_search_set_start(pm_search *s, uint32_t lane, const int64_t *ids, const float *vectors, const int32_t *neighbors);
/* seeds of the lane's dummy-query offset streams, one per sub-PIR (pir.go:363-371) */
int pm_search_set_dummy_seed(pm_search *s, uint32_t lane, const uint64_t *seeds);
/* start a round: lanes[a] searches queries[a]; rand_seeds[a] seeds the "random vertex" branch (search.go:155-159) */
int pm_search_begin(pm_search *s, const uint32_t *lanes, uint64_t act, const float *queries, const uint64_t *rand_seeds, uint64_t k,
                    int benchmarking);
int pm_search_fetch(pm_search *s, int apply_previous);
int pm_search_apply(pm_search *s);
/* ret / step_ret: [act][k] (-1 padded); stats: [act][3] = fetched entries, entries equal to the true row, server sub-queries;
 * finished: [act][partition_num] = FinishedQueryNum of every sub-PIR after the round */
int pm_search_finish(pm_search *s, int apply_previous, int64_t *ret, int64_t *step_ret, uint64_t *stats, uint64_t *finished);
/* the cached (index, entry) pairs of one sub-PIR of one lane, in insertion order; idx_out / entries_out may be NULL to
 * query the count */
int pm_search_cache_download(pm_search *s, uint32_t lane, uint32_t part, uint64_t *idx_out, uint64_t *entries_out, uint64_t cap,
                             uint64_t *count);

/* Page-locked host memory for callers that want result buffers the GPU can write directly (any host pointer works
 * everywhere; a pm_host_alloc'ed `out` of pm_client_query_batch* just saves one host-side copy of the answers). */
int pm_host_alloc(void **out, uint64_t bytes);
int pm_host_free(void *p);
/* page-lock an existing host range (e.g. a shared-memory mapping that several per-GPU processes copy parities into) */
int pm_host_register(void *host, uint64_t bytes);
int pm_host_unregister(void *host);

/* A9: squared L2 in the reference's exact fp32 order.  out[i] = L2Dist(a[i], b[i]), rows of `dim` floats. */
int pm_l2_pairs(const float *a, const float *b, uint64_t n, uint64_t dim, float *out, int device);
/* out[i] = L2Dist(vecs[i], query): one query against n host vectors (SearchKNN's per-step and re-rank call sites) */
int pm_l2_query(const float *vecs, uint64_t n, uint64_t dim, const float *query, float *out, int device);
/* out[q*k + j] = L2Dist(first `dim` fp32 of db row ids[q*k + j], queries[q]);  ids outside [0, n_rows) give +inf. */
int pm_l2_batch(pm_db *db, uint64_t dim, const float *queries, uint64_t n_queries, const int64_t *ids, uint64_t k,
                float *out);
int pm_l2_batch_dev(pm_db *db, uint64_t dim, const float *queries, uint64_t n_queries, const int64_t *ids, uint64_t k,
                    float *out, void *stream);
/* out[p] = L2Dist(first `dim` fp32 of row ids_a[p], of row ids_b[p]); a row id outside [0, n_rows) gives +inf.  The
 * distance matrices of graph construction (robustPrune, graphann/build_graph.go:169-236) in one launch. */
int pm_l2_idpairs(pm_db *db, uint64_t dim, const int64_t *ids_a, const int64_t *ids_b, uint64_t n, float *out);
int pm_l2_idpairs_dev(pm_db *db, uint64_t dim, const int64_t *ids_a, const int64_t *ids_b, uint64_t n, float *out, void *stream);

/* A11: db viewed as rows[n_rows][dim] uint32 (dim = 2*entry_u64, dim % 16 == 0 as the reference requires).
 * checksum_out[t] = sum_i InnerProduct(row_i, queries[t]) mod 2^32.  If ip_out != NULL it also receives
 * every per-row product, ip_out[t*n_rows + i]. */
int pm_ip_u32_scan(pm_db *db, uint64_t dim, const uint32_t *queries, uint64_t n_queries, uint32_t *checksum_out,
                   uint32_t *ip_out);
int pm_ip_u32_scan_dev(pm_db *db, uint64_t dim, const uint32_t *queries, uint64_t n_queries, uint32_t *checksum_out,
                       uint32_t *ip_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PACMANN_CUDA_H */
