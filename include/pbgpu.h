/*
 * pbgpu.h -- C ABI of libpbgpu.so, the B200-native interval-join engine.
 *
 * This is the drop-in boundary for the reference's hot path:
 * the three providers that polars-bio's plan builder constructs from the un-vendored crate
 * datafusion-bio-function-ranges v0.11.0,
 *     OverlapProvider::new_with_output_mode   /root/reference/src/operation.rs:253-263
 *     NearestProvider::new                    /root/reference/src/operation.rs:146-158
 *     CountOverlapsProvider::new              /root/reference/src/operation.rs:331-340
 * and the PyO3 entry points above them,
 *     range_operation_frame / _lazy / _scan   /root/reference/src/lib.rs:79-88,154-166,216-228.
 * Plain pointers and sizes only: no torch / Arrow C++ / STL types cross this boundary.
 * Every function returns 0 on success or a PBGPU_E* code; pbgpu_last_error() has the text.
 * Nothing here aborts or throws (the reference's .unwrap() panics, operation.rs:106-107,269,
 * are not reproduced).
 *
 * Two levels:
 *   (1) device level  -- columns already resident in HBM (int32 contig code, start, end);
 *                        what a long-lived Rust ExecutionPlan or torch caller drives, and what
 *                        bench.py's `value` times.
 *   (2) Arrow level   -- ArrowArrayStream in, ArrowArrayStream out (host buffers, H2D/D2H
 *                        inside); what range_operation_frame binds; bench.py's `e2e`.
 *
 * Enum values mirror /root/reference/src/option.rs:89-112.
 */
#ifndef PBGPU_H
#define PBGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBGPU_API __attribute__((visibility("default")))

/* ---- status codes ---------------------------------------------------------------------- */
#define PBGPU_OK 0
#define PBGPU_EINVAL 1   /* bad argument / unsupported option                                 */
#define PBGPU_ECUDA 2    /* CUDA runtime error (text in pbgpu_last_error)                      */
#define PBGPU_ENOMEM 3   /* device or host allocation failed                                  */
#define PBGPU_ERANGE 4   /* coordinate outside the int32 domain (reference limit,             */
                         /* docs/features/operations.md:37) or > 2^32-1 rows                  */
#define PBGPU_ESCHEMA 5  /* Arrow schema problem: missing column, unsupported type            */
#define PBGPU_ESTREAM 6  /* an input ArrowArrayStream callback failed                         */

/* ---- option enums (src/option.rs) ------------------------------------------------------ */
enum { PBGPU_FILTER_WEAK = 0, PBGPU_FILTER_STRICT = 1 };             /* option.rs:96-99   */
enum { PBGPU_OP_OVERLAP = 0, PBGPU_OP_COMPLEMENT = 1, PBGPU_OP_CLUSTER = 2, PBGPU_OP_NEAREST = 3, PBGPU_OP_COVERAGE = 4,
       PBGPU_OP_SUBTRACT = 5, PBGPU_OP_COUNT_OVERLAPS_NAIVE = 6, PBGPU_OP_MERGE = 7 };  /* option.rs:103-112; the unary
       sweeps (1, 2, 5, 7) run through pbgpu_range_op like the binary operations (and pbgpu_merge / _cluster /
       _subtract at the device level)                                                                              */
enum { PBGPU_OUT_JOIN = 0, PBGPU_OUT_LEFT = 1, PBGPU_OUT_LEFT_DISTINCT = 2 }; /* operation.rs:229-233 */

#define PBGPU_NO_PARTNER 0xFFFFFFFFu

PBGPU_API const char *pbgpu_last_error(void);          /* thread-local, never NULL            */
PBGPU_API const char *pbgpu_version(void);
PBGPU_API int pbgpu_device_count(int *count);
/* number of this library's kernels launched by the calling process so far (bench evidence) */
PBGPU_API uint64_t pbgpu_launch_count(void);

/* =========================================================================================
 * (1) Device level.  All d_* pointers are device memory on the current CUDA device; `stream`
 * is a cudaStream_t passed as void* (NULL = legacy default stream).  Work is enqueued on
 * `stream`; functions that return a host scalar synchronise that stream before returning.
 * Rows whose contig code is < 0 or >= n_contigs are null-keyed: they never match.
 * Row ids are uint32 (at most 2^32-2 rows per table).
 * ========================================================================================= */

/* Search structure over the indexed ("build") table -- replaces the per-contig COITrees that
 * OverlapProvider / NearestProvider / CountOverlapsProvider build from their indexed side.
 * Layout in HBM: rows radix-partitioned by contig and sorted by start (st, en, row), the
 * running maximum of `en` per contig (window lower bound), and the ends sorted per contig
 * (rank identity, nearest upstream candidate).                                              */
typedef struct pbgpu_index pbgpu_index;

PBGPU_API int pbgpu_index_build(const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end,
                                int64_t m, int32_t n_contigs, void *stream, pbgpu_index **out);
/* The same with an explicit id column: indexed row i is reported as d_row_ids[i] (uint32) wherever the plain build
 * reports i -- pair buffers, nearest partners.  Multi-GPU: the global row ids that travelled with the rows, so results
 * need no translation pass.  d_row_ids == NULL: pbgpu_index_build. */
PBGPU_API int pbgpu_index_build_ids(const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end,
                                    const uint32_t *d_row_ids, int64_t m, int32_t n_contigs, void *stream, pbgpu_index **out);
PBGPU_API void pbgpu_index_free(pbgpu_index *ix);
/* Stream-ordered release: the index memory becomes reusable only after everything enqueued on `stream` so far (the
 * kernels that read the index) has run.  The right call for users of non-blocking streams. */
PBGPU_API void pbgpu_index_free_async(pbgpu_index *ix, void *stream);
PBGPU_API int64_t pbgpu_index_rows(const pbgpu_index *ix);     /* rows kept (non-null keys)   */
PBGPU_API size_t pbgpu_index_bytes(const pbgpu_index *ix);     /* HBM held by the index       */

/* CountOverlapsProvider (operation.rs:331-340), coverage=false: for every iterated row the
 * number of indexed rows on the same contig satisfying the FilterOp predicate
 * (docs/developers.md:549-552).  d_counts: int64[n] (the reference's `count` column type,
 * range_op_helpers.py:315-316).                                                             */
PBGPU_API int pbgpu_count_overlaps(const pbgpu_index *ix, const int32_t *d_contig, const int32_t *d_start,
                                   const int32_t *d_end, int64_t n, int filter_op, int64_t *d_counts,
                                   void *stream);

/* CountOverlapsProvider, coverage=true: positions of every iterated row covered by the union
 * of the indexed rows (Strict: half-open lengths; Weak: closed lengths).  int64[n].         */
PBGPU_API int pbgpu_coverage(const pbgpu_index *ix, const int32_t *d_contig, const int32_t *d_start,
                             const int32_t *d_end, int64_t n, int filter_op, int64_t *d_coverage,
                             void *stream);

/* OverlapProvider (operation.rs:253-263) as two passes so the pair buffer is exact-sized:
 *   pass 1  pbgpu_overlap_count  -> plan + total number of pairs (host scalar; syncs stream)
 *   pass 2  pbgpu_overlap_emit   -> (probe_row, build_row) uint32 pairs, SoA, ordered by
 *                                   probe row, then by (start, row) of the indexed partner.
 * The plan borrows the probe columns and the index: keep them alive until the plan is freed. */
typedef struct pbgpu_overlap_plan pbgpu_overlap_plan;

PBGPU_API int pbgpu_overlap_count(const pbgpu_index *ix, const int32_t *d_contig, const int32_t *d_start,
                                  const int32_t *d_end, int64_t n, int filter_op, void *stream,
                                  pbgpu_overlap_plan **plan, int64_t *total_pairs);
/* pbgpu_overlap_count with an id column for the probes: probe row i is reported as d_probe_ids[i] (NULL: i). */
PBGPU_API int pbgpu_overlap_count_ids(const pbgpu_index *ix, const int32_t *d_contig, const int32_t *d_start,
                                      const int32_t *d_end, const uint32_t *d_probe_ids, int64_t n, int filter_op,
                                      void *stream, pbgpu_overlap_plan **plan, int64_t *total_pairs);
PBGPU_API int pbgpu_overlap_emit(const pbgpu_overlap_plan *plan, uint32_t *d_probe_rows,
                                 uint32_t *d_build_rows, void *stream);
/* per-probe pair counts of pass 1 (uint32[n], device; valid until the plan is freed) -- the
 * Left / LeftDistinct output modes (operation.rs:229-233) are filters over these.           */
PBGPU_API const uint32_t *pbgpu_overlap_plan_counts(const pbgpu_overlap_plan *plan);
PBGPU_API void pbgpu_overlap_plan_free(pbgpu_overlap_plan *plan);
/* same, with the plan's device scratch released in stream order on `stream` (the stream pass 2 was enqueued on):
 * no host synchronisation is needed between pbgpu_overlap_emit and the free                                     */
PBGPU_API void pbgpu_overlap_plan_free_async(pbgpu_overlap_plan *plan, void *stream);
/* Streaming sink (BASELINE config 5: ~1e9 pairs through a bounded buffer).  Pass 1 leaves the
 * exclusive pair offset of every 256-probe block; pass 2 can then be run over any block range [blk_lo, blk_hi)
 * into a buffer of offsets[blk_hi] - offsets[blk_lo] pairs, so a consumer sizes each chunk to its ring slot
 * (reference: the output stream of IntervalJoinExec yields bounded RecordBatches, range_op_io.py:148-161). */
PBGPU_API int64_t pbgpu_overlap_plan_blocks(const pbgpu_overlap_plan *plan);
PBGPU_API int pbgpu_overlap_plan_block_offsets(const pbgpu_overlap_plan *plan, uint64_t *h_offsets /* [blocks+1], host */,
                                               void *stream);
PBGPU_API int pbgpu_overlap_emit_blocks(const pbgpu_overlap_plan *plan, int64_t blk_lo, int64_t blk_hi,
                                        uint32_t *d_probe_rows, uint32_t *d_build_rows, void *stream);

/* NearestProvider (operation.rs:146-158): for every iterated row up to k indexed rows of the
 * same contig ordered by (overlapping first when include_overlaps, distance, start, row);
 * distance = 0 for overlapping partners else max(b.start-a.end, a.start-b.end) (>= 0)
 * (tests/_expected.py:162).  d_partner: uint32[n*k] (PBGPU_NO_PARTNER = none),
 * d_distance: int64[n*k] (-1 = none) or NULL when compute_distance is off.                  */
PBGPU_API int pbgpu_nearest(const pbgpu_index *ix, const int32_t *d_contig, const int32_t *d_start,
                            const int32_t *d_end, int64_t n, int filter_op, int64_t k, int include_overlaps,
                            uint32_t *d_partner, int64_t *d_distance, void *stream);

/* Multi-GPU plumbing: bucket rows by owning rank before the NCCL all-to-all.
 * d_owner: int32[n_contigs] contig -> rank table.  Writes per-rank row counts into
 * d_rank_counts (int64[n_ranks]) and, stably grouped by destination rank, 16-byte records
 * (contig, start, end, row_id_base + row) into d_packed (int32[4*n]; only the first
 * sum(d_rank_counts) records are meaningful -- null-keyed rows sort behind them).           */
PBGPU_API int pbgpu_pack_by_owner(const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end,
                                  int64_t n, const int32_t *d_owner, int32_t n_contigs, int32_t n_ranks,
                                  uint32_t row_id_base, int32_t *d_packed, int64_t *d_rank_counts,
                                  void *stream);

/* d_out[i] = d_src[d_rows[i]] (0 where d_rows[i] == PBGPU_NO_PARTNER): materialises the key columns
 * (contig code, start, end) of result rows on the device, where they already live, so the host does not have
 * to gather them from the input tables (the first step of payload materialisation).   */
PBGPU_API int pbgpu_gather_i32(const int32_t *d_src, const uint32_t *d_rows, int64_t n, int32_t *d_out, void *stream);
/* Rows per contig, ADDED onto d_hist (int64[n_contigs]; zero it first); null keys are ignored.  Feeds the
 * owner table (LPT bin packing) of the multi-GPU exchange.                                               */
PBGPU_API int pbgpu_contig_histogram(const int32_t *d_contig, int64_t n, int32_t n_contigs, int64_t *d_hist,
                                     void *stream);
/* After the all-to-all: split received 16-byte records back into columns (d_row = global row ids).     */
PBGPU_API int pbgpu_unpack_records(const int32_t *d_packed, int64_t n, int32_t *d_contig, int32_t *d_start,
                                   int32_t *d_end, uint32_t *d_row, void *stream);
/* Pair buffers of a shard hold positions in the shard's received columns; map them to global row ids:
 * d_out[i] = d_global_of_local[d_local[i]]  (PBGPU_NO_PARTNER passes through).  d_out may alias d_local. */
PBGPU_API int pbgpu_translate_rows(const uint32_t *d_local, int64_t n, const uint32_t *d_global_of_local,
                                   uint32_t *d_out, void *stream);

/* The same exchange over NVLink peer memory, without NCCL on the data path (one process per GPU on one node; there is
 * no reference counterpart -- the reference is single-process).  Every rank owns two receive ARENAS
 * (they alternate between steps) and one CONTROL block (pbgpu_peer_alloc: cudaMalloc + CUDA IPC handle), maps those of
 * its peers (pbgpu_peer_open) and stores each row straight into the column arrays of the rank that owns its contig.
 *
 * Arena layout: table t at byte 16*sum(cap_rows[<t]); columns contig | start | end | row, each cap_rows[t] x 4 bytes
 * (cap_rows: multiples of 64).  Control block (pbgpu_peer_ctl_bytes, zeroed before the handles are exchanged): flag
 * words written remotely with st.release.sys and polled locally with a bounded spin ($PBGPU_PEER_TIMEOUT_MS, default
 * 60 s: a dead peer becomes status 2 in d_result[3*n_tables] / *d_status, never a hung GPU) and one histogram slot per
 * source rank.  `step` counts exchange steps from 1 and must be the same on every rank.
 *
 * One step = pbgpu_peer_begin + one pbgpu_peer_table per table:
 *   histograms    rows per contig of this rank's slice of every table (+ the slice sizes), one launch; the block that
 *                 finishes last publishes them into every control block and raises this rank's flag A there
 *   plan          waits for flag A of every source; then, on the device and identically on every rank: contig -> owner
 *                 (LPT bin packing), this rank's region in every destination's arena (regions in source-rank order, so
 *                 received rows are ordered by global row id as after the stable NCCL exchange).  Outputs: d_owner
 *                 int32[n_contigs], d_dst [n_tables][world] records of 4 pointers, d_result int64[3*n_tables+1] =
 *                 received rows per table | global row id base per table | largest region any rank needs per table |
 *                 flag (1 = some region exceeds cap_rows: nothing is scattered; 2 = a peer did not publish in time).
 *                 h_result (page-locked host int64[3*n_tables+2], or NULL): the kernel stores the same words there and
 *                 then the step number into the last one, so the host spins on that word instead of copying.
 *   scatter       one table: block destination counts, their scan, and the scatter itself
 *   signal, wait  flag B[table] raised at every destination after the scatter (same stream); wait until flag B[table]
 *                 of every source shows `step`: the table is complete on this rank
 * pbgpu_peer_table may run on a stream of its own ordered after pbgpu_peer_begin (work on one table overlaps the
 * transfer of the next).  ctl_base == NULL: no control blocks -- begin stops after the histograms (all-gather d_hist into
 * int64 [world][n_tables][n_contigs+1], call pbgpu_peer_plan with it), table only scatters, and the caller orders "all
 * peers have written" before "I read" with a tiny all_reduce.                                                       */
PBGPU_API int pbgpu_peer_alloc(size_t bytes, void **d_ptr, unsigned char *handle_out /* [64] */);
PBGPU_API int pbgpu_peer_free(void *d_ptr);
PBGPU_API int pbgpu_peer_open(const unsigned char *handle /* [64] */, void **d_ptr);
PBGPU_API int pbgpu_peer_close(void *d_ptr);
PBGPU_API size_t pbgpu_peer_ctl_bytes(int32_t world, int32_t n_tables, int32_t n_contigs);
/* d_own_ctl != NULL: wait for flag A, histograms from the control block (d_gathered ignored); NULL: plain d_gathered.
 * arena_base / cap_rows: HOST arrays [world] / [n_tables].                                                          */
PBGPU_API int pbgpu_peer_plan(const int64_t *d_gathered, const void *d_own_ctl, uint64_t step, int32_t world, int32_t rank,
                              int32_t n_tables, int32_t n_contigs, const uint64_t *arena_base, const int64_t *cap_rows,
                              int32_t *d_owner, void *d_dst, int64_t *d_result, int64_t *h_result, void *stream);
/* d_row_id_base, d_dst (+ table * world records) and d_flag (= d_result + 3*n_tables) are read on the device.      */
PBGPU_API int pbgpu_peer_scatter(const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, int64_t n,
                                 const int32_t *d_owner, int32_t n_contigs, int32_t n_ranks, const int64_t *d_row_id_base,
                                 const void *d_dst, const int64_t *d_flag, void *stream);
#define PBGPU_PEER_HIST 1     /* pbgpu_peer_begin: histograms (+ publication)                */
#define PBGPU_PEER_PLAN 2     /* pbgpu_peer_begin: plan                                      */
#define PBGPU_PEER_SCATTER 4  /* pbgpu_peer_table: scatter                                   */
#define PBGPU_PEER_SIGNAL 8   /* pbgpu_peer_table: raise this rank's flag at every peer      */
#define PBGPU_PEER_WAIT 16    /* pbgpu_peer_table: wait for every peer's flag                */
typedef struct {
  int32_t world, rank, n_tables, n_contigs;
  int32_t phases;                /* 0 = everything the call does; else a PBGPU_PEER_* mask (tests drive the  */
                                 /* phases of several simulated ranks one by one on a single stream)        */
  int32_t reserved;
  uint64_t step;                 /* counts from 1, the same on every rank                                   */
  const int32_t *contig[4];      /* this rank's slice of every table (device)                               */
  const int32_t *start[4];
  const int32_t *end[4];
  int64_t rows[4];
  const uint64_t *arena_base;    /* HOST [world]: every rank's arena of this step, addresses in this process */
  const uint64_t *ctl_base;      /* HOST [world]: control blocks, or NULL                                    */
  const int64_t *cap_rows;       /* HOST [n_tables]                                                          */
  int64_t *d_hist;               /* device int64 [n_tables*(n_contigs+1) + 1]                                */
  int32_t *d_owner;              /* device int32 [n_contigs]                                                 */
  void *d_dst;                   /* device, 32 bytes x n_tables x world                                      */
  int64_t *d_result;             /* device int64 [3*n_tables+1]                                              */
  int64_t *h_result;             /* page-locked host int64 [3*n_tables+2], or NULL                           */
  int64_t *d_status;             /* device int64 [1]: set to 2 by a wait that timed out                      */
  void *d_scratch;               /* device scratch of pbgpu_peer_scratch_bytes(step) bytes kept by the caller, */
  uint64_t scratch_bytes;        /* or NULL / too small: the calls allocate stream-ordered scratch themselves */
} pbgpu_peer_step;
PBGPU_API size_t pbgpu_peer_scratch_bytes(const pbgpu_peer_step *step);
PBGPU_API int pbgpu_peer_begin(const pbgpu_peer_step *step, void *stream);
PBGPU_API int pbgpu_peer_table(const pbgpu_peer_step *step, int32_t table, void *stream);
/* ---- unary sweeps: they reuse the contig partition + start sort of the index build.
 * Reference: MergeProvider / ClusterProvider / ComplementProvider / SubtractProvider constructed at
 * /root/reference/src/operation.rs:352-380, 382-430, 432-461, 463-510; behaviour pinned by tests/_expected.py:174-181,
 * tests/test_coordinate_system_metadata.py:1032-1054, tests/test_partitioned_range_operation_regressions.py:24-59.
 * A row joins the running interval iff  start < reach + min_dist  (Strict, 0-based half-open) /
 * start <= reach + min_dist  (Weak, 1-based closed), reach = largest end so far; min_dist >= 0.
 * Results are device tables owned by the library (exact-sized; the call synchronises `stream` once to size them). */
typedef struct pbgpu_intervals pbgpu_intervals;
PBGPU_API int64_t pbgpu_intervals_rows(const pbgpu_intervals *t);
/* device columns of a result (any out pointer may be NULL): merge -> contig, start, end, count (n_intervals);
 * subtract -> row (uint32 row of the left table), start, end; columns a result does not have read NULL          */
PBGPU_API int pbgpu_intervals_columns(const pbgpu_intervals *t, const int32_t **d_contig, const uint32_t **d_row,
                                      const int32_t **d_start, const int32_t **d_end, const int64_t **d_count);
/* copy columns into caller-owned device buffers of pbgpu_intervals_rows() entries (NULL = skip), on `stream`     */
PBGPU_API int pbgpu_intervals_copy(const pbgpu_intervals *t, int32_t *d_contig, uint32_t *d_row, int32_t *d_start,
                                   int32_t *d_end, int64_t *d_count, void *stream);
PBGPU_API void pbgpu_intervals_free(pbgpu_intervals *t, void *stream);   /* stream-ordered release on `stream` */
/* merge: one row per merged interval, ordered by (contig code, start)                                            */
PBGPU_API int pbgpu_merge(const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, int64_t m,
                          int32_t n_contigs, int filter_op, int64_t min_dist, void *stream, pbgpu_intervals **out);
/* cluster: for every input row the id of its merged interval (numbered from 0 in (contig code, start) order; -1 for
 * null-keyed rows) and that interval's start / end.  d_cluster int64[m], d_cluster_start / d_cluster_end int32[m]. */
PBGPU_API int pbgpu_cluster(const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, int64_t m,
                            int32_t n_contigs, int filter_op, int64_t min_dist, int64_t *d_cluster,
                            int32_t *d_cluster_start, int32_t *d_cluster_end, int64_t *n_clusters, void *stream);
/* subtract: for every left row the pieces no right row of the same contig covers, ordered by (left row, start); an
 * untouched row passes through whole, a covered one leaves nothing.  Right rows with start > end cover nothing, left
 * rows with start > end pass through.  Weak (closed coordinates): [s,e] minus [s2,e2] leaves [s,s2-1] and [e2+1,e].
 * complement (operation.rs:432-461) is this call with the view table on the left.                                 */
PBGPU_API int pbgpu_subtract(const int32_t *l_contig, const int32_t *l_start, const int32_t *l_end, int64_t n,
                             const int32_t *r_contig, const int32_t *r_start, const int32_t *r_end, int64_t m,
                             int32_t n_contigs, int filter_op, void *stream, pbgpu_intervals **out);

/* CUDA-event durations (ns) of the most recent index build / provider kernels issued by the calling
 * thread, measured on the stream they were launched on (events are recorded on every call; this function
 * waits for the last one).  A stage that has not run yet reads 0.                              */
typedef struct {
  uint64_t partition_sort_ns;  /* pbgpu_index_build: contig partition + start sort + aux arrays + directories */
  uint64_t count_ns;           /* overlap pass 1 (incl. the probe partition when the index is beyond the L2) */
  uint64_t scan_ns;            /* pair-offset scan of pass 1                                   */
  uint64_t emit_ns;            /* overlap pass 2 kernel                                        */
  uint64_t count_overlaps_ns;  /* pbgpu_count_overlaps: every kernel of the call (partition + count + un-binning when used) */
  uint64_t bin_ns;             /* most recent probe partition (histogram + one radix pass; csrc/bins.cuh), 0 if none;    */
                               /* it is INSIDE count_ns / count_overlaps_ns of the call that ran it                      */
  uint64_t unbin_ns;           /* pbgpu_count_overlaps: counts back from bin order to row order                          */
} pbgpu_stage_times;
PBGPU_API int pbgpu_last_stage_times(pbgpu_stage_times *out);

/* =========================================================================================
 * (2) Arrow level -- the binding target of range_operation_frame / _lazy (src/lib.rs:79-88,
 * 154-166) and of a DataFusion ExecutionPlan adapter.
 * Arrow C Data / C Stream Interface structs as defined by the Arrow specification.
 * ========================================================================================= */
struct ArrowArrayStream;

typedef struct {                 /* mirrors RangeOptions, src/option.rs:8-41                  */
  int32_t range_op;              /* PBGPU_OP_*                                                */
  int32_t filter_op;             /* PBGPU_FILTER_*                                            */
  int32_t output_mode;           /* PBGPU_OUT_* (overlap only)                                */
  int32_t emit;                  /* 0 = materialised rows (reference column contract,        */
                                 /*     operation.rs:170-195,272-299); 1 = index pairs:       */
                                 /*     overlap -> (left_row u32, right_row u32)              */
                                 /*     nearest -> (left_row u32, right_row u32 nullable,     */
                                 /*                 distance i64 nullable)                    */
  const char *cols1[3];          /* contig, start, end column names of `left`  (columns_1)    */
  const char *cols2[3];          /* contig, start, end column names of `right` (columns_2)    */
  const char *suffixes[2];       /* NULL -> "_1","_2"                                         */
  uint64_t nearest_k;            /* 0 -> 1                                                    */
  int32_t include_overlaps;      /* nearest                                                   */
  int32_t compute_distance;      /* nearest                                                   */
  uint64_t limit;                /* 0 = none (src/lib.rs:120-131)                             */
  uint32_t max_batch_rows;       /* 0 -> 1<<20; low_memory / batch_size cap (range_op.py:168) */
  int32_t device;                /* CUDA device ordinal, -1 = current                         */
  uint64_t sink_pairs;           /* overlap: pairs per ring slot of the streaming sink; a larger */
                                 /* result is emitted chunk by chunk from out->get_next with the */
                                 /* device state kept alive.  0 -> $PBGPU_SINK_PAIRS -> 1<<24    */
  int64_t min_dist;              /* merge / cluster: rows closer than this are joined (>= 0;     */
                                 /* range_op.py:599-700 -> operation.rs:352-430)                 */
} PbRangeOptions;

/* `left` / `right` are df1 / df2 exactly as the Python facade passes them to
 * range_operation_frame (so for count_overlaps / coverage the facade has already swapped
 * them, range_op.py:407-409,511; for nearest the engine swaps roles internally the way
 * do_nearest does, operation.rs:143-158).
 * Ownership: the callee MOVES left/right (calls their release exactly once, on the calling
 * thread, before returning -- the GIL rule of src/lib.rs:63-72); the caller owns `out` and
 * calls out->release.  out->get_next may be called from another thread, not concurrently.
 * Unary sweeps (do_merge / do_cluster / do_complement / do_subtract, operation.rs:352-510): `left` is the input table
 * (cols1); merge and cluster ignore `right` (may be NULL); subtract takes the second table as `right` (cols2); complement
 * takes the view table as `right` (cols2) or NULL = every contig present spans [0, INT64_MAX).  Output: merge -> contig,
 * start, end (Int64, named like cols1), n_intervals; cluster -> every input column + cluster, cluster_start, cluster_end
 * (Int64); complement -> contig, start, end; subtract -> every `left` column with the interval columns holding the
 * remaining pieces (Int64).  Contigs are ordered by name.                                                              */
PBGPU_API int pbgpu_range_op(struct ArrowArrayStream *left, struct ArrowArrayStream *right,
                             const PbRangeOptions *opts, struct ArrowArrayStream *out);

/* Build once, probe many: the streamed counterpart of pbgpu_range_op -- what range_operation_lazy / _scan
 * (src/lib.rs:154-166, 216-228) and the probe-side batch streams of src/scan.rs:103-139, 320-357 do in the reference:
 * the indexed side is collected and indexed ONCE, the iterated side arrives in chunks (one pbgpu_range_probe call per
 * chunk of record batches; concurrent calls from several partitions / threads are allowed), and every call returns the
 * result rows of its chunk with the same column contract as pbgpu_range_op.  Host staging is sized by the chunk, not by
 * the table.  Which input is the indexed one follows the operation (see pbgpu_range_op): overlap / nearest index
 * `right` (df2, cols2) and iterate `left`; count_overlaps / coverage index `left` (cols1) and iterate `right`.
 * With emit = 1 the iterated-side row ids of a probe are relative to its chunk.  opts->limit applies per probe call.
 * Ownership: both calls MOVE their input stream (released before returning); the caller owns `out` and the session. */
typedef struct pbgpu_range_session pbgpu_range_session;
PBGPU_API int pbgpu_range_open(struct ArrowArrayStream *indexed, const PbRangeOptions *opts, pbgpu_range_session **out);
PBGPU_API int pbgpu_range_probe(pbgpu_range_session *session, struct ArrowArrayStream *iterated, struct ArrowArrayStream *out);
PBGPU_API void pbgpu_range_close(pbgpu_range_session *session);
/* Page-locked host staging the Arrow level holds right now and its high-water mark since the last reset (bytes): what a
 * streamed join keeps bounded by its chunk size.  Diagnostics; any pointer may be NULL. */
PBGPU_API void pbgpu_pinned_stats(uint64_t *busy_bytes, uint64_t *peak_bytes, int reset_peak);

#ifdef __cplusplus
}
#endif
#endif /* PBGPU_H */
