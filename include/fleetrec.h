/* fleetrec.h -- C ABI of the B200-native FleetRec inference hot path.
 *
 * One call chain replaces three reference boundaries:
 *   B1  FPGA lookup kernel launch   embedding_47_krnl(...)            FPGA/kernel/user_krnl/embedding_47_krnl/src/hls/embedding_47_krnl.hpp:32-71
 *       + its host bring-up         host.cpp setArg/enqueueTask        FPGA/host/embedding_47_krnl/host.cpp:691-761
 *   B2  wire format                 item-major LE fp32, INPUT_SIZE/item embedding_47_krnl.cpp:774, GPU/.../constant.h:37-38
 *   B3  GPU MLP worker              thread_consume(CUDA_thread_info*)  GPU/final_network_cublasLt_1_node_no_FIFO_scatter/cuda_server.c:91-101
 *
 * Conventions: plain C types only (no CUDA/torch types in signatures); every call
 * returns an fr_status (0 = ok) and leaves a message retrievable through
 * fr_last_error(); nothing ever calls exit().  The engine owns all device memory.
 * One engine drives ONE GPU (one process per GPU); table sharding across
 * processes is configured with the fr_shard_* calls.  There is no CPU fallback:
 * fr_create fails with FR_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef FLEETREC_H
#define FLEETREC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int fr_status;
enum {
  FR_OK = 0,
  FR_ERR_INVALID = 1,     /* bad argument / descriptor                       */
  FR_ERR_CUDA = 2,        /* CUDA runtime or driver error (message has it)   */
  FR_ERR_OOM = 3,         /* device or host allocation failed                */
  FR_ERR_STATE = 4,       /* call order: table/layer not loaded, not sharded */
  FR_ERR_UNSUPPORTED = 5  /* shape the kernels do not cover                  */
};

/* Memory tier a table had in the reference (constants.hpp); informational on B200
 * (everything is HBM3e) except that FR_TIER_PLRAM tables are replicated when sharding. */
enum { FR_TIER_HBM = 0, FR_TIER_DDR = 1, FR_TIER_PLRAM = 2, FR_TIER_CPU = 3 };

/* MLP semantics.  LINEAR is what cuda_server.c:468-491 executes (4 GEMMs, alpha=1,
 * beta=0, no bias, no activation).  BIAS_RELU_SIGMOID is what constant.h:1-17
 * documents (W*x+B) plus ReLU between layers and a final sigmoid (north star). */
enum { FR_MLP_LINEAR = 0, FR_MLP_BIAS_RELU_SIGMOID = 1 };

/* Arithmetic of the three hidden GEMMs.
 *   TF32   tcgen05.mma kind::tf32, operands rounded to TF32 (rna), FP32 accumulate in TMEM
 *   FP32   SIMT FFMA kernels, FP32 operands and accumulate (reference's CUBLAS_COMPUTE_32F) */
enum { FR_PREC_TF32 = 0, FR_PREC_FP32 = 1 };

/* Storage type of the embedding tables in HBM.  The reference stores
 * fp32 (constants.hpp:4,11: one axi word = 4 floats).  F16 / BF16 halve the table bytes and the
 * lookup's row traffic, FP8 (E4M3) quarters them; rows are converted with round-to-nearest-even when loaded and widened
 * exactly to fp32 by the lookup, so the concat vector is bit-exact AFTER THE STATED DEQUANT:
 * concat == float32(float16(row)) element for element.  Everything after the lookup is unchanged. */
enum { FR_TABLE_F32 = 0, FR_TABLE_F16 = 1, FR_TABLE_BF16 = 2, FR_TABLE_FP8 = 3 /* E4M3, saturating: |x| > 448 -> +-448 */ };

typedef struct fr_table_desc {
  int tier;          /* FR_TIER_*                                              */
  int tier_index;    /* i of TABLE_SIZE_<tier>_<i>                             */
  int bank;          /* i mod <tier>_BANK_NUM                                  */
  int round;         /* i div <tier>_BANK_NUM                                  */
  int64_t rows;      /* TABLE_SIZE_<tier>_<i>                                  */
  int dim;           /* DATA_SIZE_<tier>_<i> == 4*AXI_PADDED_SIZE (floats)     */
} fr_table_desc;

/* One contiguous piece of the per-item concat vector: floats
 * [dst, dst+len) = table[idx[table]][col, col+len).  The list restates
 * gather_embeddings() of each kernel, including the medium model's duplicate pad. */
typedef struct fr_segment_desc {
  int dst, table, col, len;
} fr_segment_desc;

typedef struct fr_model_desc {
  const char* name;
  int n_tables;
  const fr_table_desc* tables;      /* idx[b][t] addresses tables[t]           */
  int n_segments;
  const fr_segment_desc* segments;
  int concat_floats;                /* INPUT_SIZE (multiple of 16)             */
  int hidden[4];                    /* HIDDEN_SIZE1..3, OUTPUT_SIZE (=1)       */
  int mlp_mode;                     /* FR_MLP_*                                */
  int precision;                    /* FR_PREC_*                               */
  int max_batch;                    /* largest B a stream will be given        */
  int table_dtype;                  /* FR_TABLE_* (0 = fp32, the reference)    */
} fr_model_desc;

typedef struct fr_engine fr_engine;
/* A worker context: one CUDA stream plus the activation workspaces for one batch
 * in flight.  Mirrors one thread_consume() worker (cuda_server.c:101-183). */
typedef struct fr_stream_s* fr_stream;

/* ---- catalogue ---------------------------------------------------------- */
/* Fill *out with a built-in catalogue: "small" (47 tables, 352), "medium" (98,
 * 880), "large_half" (188, 1952), "large" (377, 3968).  Pointers inside stay
 * valid for the life of the library.  mlp_mode/precision/max_batch get defaults
 * (BIAS_RELU_SIGMOID, TF32, 16384) the caller may overwrite. */
fr_status fr_model_builtin(const char* name, fr_model_desc* out);

/* ---- engine ------------------------------------------------------------- */
/* n_gpus must be 1 (one process per GPU); device_ids[0] is the CUDA ordinal. */
fr_status fr_create(const fr_model_desc* desc, int n_gpus, const int* device_ids, fr_engine** out);
void fr_destroy(fr_engine* e);
/* Message of the last failing call on this engine (or of the last failing
 * fr_create / fr_model_builtin on this thread when e == NULL). */
const char* fr_last_error(const fr_engine* e);

/* Override the row count of a table before it is loaded/filled (tests and
 * memory-capped runs use the real dims with fewer rows). */
fr_status fr_set_table_rows(fr_engine* e, int table_id, int64_t rows);

/* Copy a host table image [rows][dim] fp32 into HBM (converted to the engine's table_dtype on the
 * device, round-to-nearest-even); caller keeps ownership.
 * Replaces host.cpp:324-423,592-750 (vector alloc + init + migrate). */
fr_status fr_load_table(fr_engine* e, int table_id, const float* host_rows, int64_t rows, int dim);
/* Device-side fills, bit-identical to the oracle's:
 *   reference: even rows 1.0f, odd rows 0.0f (host.cpp:66-88, embedding_47_krnl.cpp:871-897);
 *              debug_rows > 0 reproduces the `#define DEBUG` truncation (first debug_rows rows).
 *   hash:      every float a distinct finite normal derived from (seed, table, row, col). */
fr_status fr_fill_table_reference(fr_engine* e, int table_id, int64_t debug_rows);
fr_status fr_fill_table_hash(fr_engine* e, int table_id, uint32_t seed);
/* Read rows back (tests): copies [n_rows][dim] starting at first_row to host, widened to fp32. */
fr_status fr_read_table(fr_engine* e, int table_id, int64_t first_row, int64_t n_rows, float* host_out);

/* 1 when the library was built with -DFR_EXPERIMENTS (`make exp`): the measured-slower kernel variants of DESIGN.md
 * section 4 (one-launch MLP chain, multicast clusters, cp.async A loader, zero-copy index staging, programmatic
 * dependent launch) and the environment variables that select them exist only there.  A release build reads two
 * environment variables, both test hooks, once per fr_create: FR_TC_TILES=N1,N2,N3[,ctas] (pin the tcgen05 tile width
 * of layers 1..3 and the CTAs per tile) and FR_TC_MAX_CLUSTERS=n (cap the persistent grids). */
int fr_build_has_experiments(void);

/* ---- engine options ------------------------------------------------------- */
/* Set at any time between calls; a change synchronises the device and drops the worker streams' cached CUDA graphs.
 *   FR_OPT_CUDA_GRAPHS    1 (default): a (buffers, B) combination seen before on a worker is replayed as one CUDA
 *                         graph; 0: every step is issued as plain launches.
 *   FR_OPT_CHECK_INDICES  1: the lookup kernels compare every index with its table's row count; offenders read row 0
 *                         and the next fr_sync returns FR_ERR_INVALID naming one of them.  0 (default): unchecked,
 *                         like the reference (embedding_47_krnl.cpp:925-934).  fr_ingest validates FR_INGEST_INDICES
 *                         blocks on the host either way (they come off a socket).
 *   FR_OPT_FUSE_LOOKUP    1: fr_infer gathers straight into layer 1's shared-memory A tile (no concat vector in
 *                         global memory; TF32, fp32 tables, single GPU).  0 (default): separate lookup kernel.
 *   FR_OPT_TILE_HINT      FR_HINT_LATENCY: one batch at a time -- narrow tcgen05 tiles spread over as many SMs as the
 *                         batch allows; FR_HINT_THROUGHPUT: many batches in flight on several workers -- wide tiles,
 *                         few CTAs per launch; FR_HINT_AUTO (default): latency when the engine has at most two
 *                         worker streams (the reference's THREAD_NUM, cuda_server.c:554-556), else throughput.
 *   FR_OPT_F16_OPERANDS   FR_F16_GUARDED: the tcgen05 path computes on fp16 operands and activations (same 11-bit
 *                         significand as TF32, half the bytes, fp32 accumulate) IF a range analysis of the loaded
 *                         tables and weights proves that no operand can leave fp16's normal range (largest table
 *                         magnitude, |W|^T.bound + |b| layer by layer); otherwise, and for fr_mlp_only / fr_layer_only
 *                         whose inputs it cannot bound, TF32.  FR_F16_OFF (default): TF32.  fr_f16_report says
 *                         which one an engine runs and why.
 *   FR_OPT_INDEX_FORMAT   FR_IDX_I32 (default): index rows are int32 [n_tables], what the reference streams
 *                         (load_access_idx, embedding_47_krnl.cpp:899-914).  FR_IDX_PACKED: a transport format for the
 *                         host -> device hop, which is what bounds the end-to-end rate once the kernels are fast:
 *                         every `idx` argument of the hot-path calls then points to packed rows -- the columns of tables
 *                         with more than 65536 rows as int32, in table order, followed by the columns of the smaller
 *                         tables as uint16, the row padded to a multiple of 4 bytes (small model: 120 instead of 188
 *                         bytes per item).  fr_index_layout describes the rows; the lookup kernels read either format
 *                         through the same per-piece byte offsets.  Not with FR_OPT_FUSE_LOOKUP or FR_INGEST_INDICES. */
enum { FR_OPT_CUDA_GRAPHS = 0, FR_OPT_CHECK_INDICES = 1, FR_OPT_FUSE_LOOKUP = 2, FR_OPT_TILE_HINT = 3,
       FR_OPT_F16_OPERANDS = 4, FR_OPT_INDEX_FORMAT = 5 };
enum { FR_IDX_I32 = 0, FR_IDX_PACKED = 1 };
enum { FR_HINT_AUTO = 0, FR_HINT_LATENCY = 1, FR_HINT_THROUGHPUT = 2 };
enum { FR_F16_OFF = 0, FR_F16_GUARDED = 1 };
fr_status fr_set_option(fr_engine* e, int option, int value);
/* Layout of one index row under the engine's FR_OPT_INDEX_FORMAT.  which = 2: the full rows of fr_infer / fr_infer_many /
 * fr_shard_infer (all tables); which = 0 / 1: the column-sliced blocks of fr_shard_infer_sliced (owned / replicated
 * tables, the lists of fr_shard_tables).  byte_offset[i], width[i] (2 or 4): where column i of that list sits in a row
 * (either may be null); *n = columns, *row_bytes = bytes per row (a multiple of 4). */
fr_status fr_index_layout(fr_engine* e, int which, int32_t* byte_offset, int32_t* width, int* n, int* row_bytes);
/* Outcome of the FR_F16_GUARDED range analysis (run by the first fr_infer after tables / weights / the option
 * changed; this call runs it if it is due): *active = 1 when fr_infer computes on fp16 operands; bounds[0..2] =
 * upper bounds of |concat element|, |H1 element|, |H2 element| (must stay below 60000); bounds[3] = smallest
 * non-zero table magnitude (must be fp16-normal, >= 2^-14); bounds[4] = largest share of a unit's weight mass that
 * fp16 would represent inexactly (must stay below 1e-6). */
fr_status fr_f16_report(fr_engine* e, int* active, float* bounds5);

/* Layer k in 0..3.  W is the reference's layout: column-major out_k x in_k with
 * ld = out_k (cuda_server.c:215,253,291,329), i.e. row-major [in_k][out_k].
 * bias [out_k] may be NULL (treated as zeros; ignored in LINEAR mode). */
fr_status fr_load_mlp(fr_engine* e, int layer, const float* W_in_major, const float* bias);
fr_status fr_set_mlp_mode(fr_engine* e, int mlp_mode);
fr_status fr_set_precision(fr_engine* e, int precision);

/* ---- worker streams ----------------------------------------------------- */
fr_status fr_stream_create(fr_engine* e, fr_stream* out);
void fr_stream_destroy(fr_engine* e, fr_stream s);
/* Raw cudaStream_t of a worker (so callers holding CUDA code can order against it). */
void* fr_stream_cuda(fr_stream s);

/* ---- hot path ----------------------------------------------------------- */
/* idx: [B][n_tables] int32, row index per table per item; scores: [B] fp32.
 * Both may be host or device pointers (detected); host buffers are copied
 * inside the call chain on the worker's stream (pin them for true async).
 * stream == NULL uses the engine's default worker.  Asynchronous: results are
 * valid after fr_sync().  Out-of-range indices are a caller error, detected only
 * with FR_OPT_CHECK_INDICES (the reference never checks).  A (idx, scores, B)
 * combination a worker has seen before is replayed as a CUDA graph: the buffers
 * must stay allocated as what they were (fr_graph_flush forgets them). */
fr_status fr_infer(fr_engine* e, const int32_t* idx, int B, float* scores, fr_stream s);
/* n batches of B items (B <= max_batch) in one call on one worker: idx [n][B][n_tables] and scores [n][B]
 * contiguous.  Every batch runs through the same kernels as fr_infer, back to back; host buffers travel in ONE copy
 * each way, which is what the call is for: a copy occupies the copy engine for ~4 us on top of its bytes, so
 * per-batch copies of the reference's staging loop (read() -> cudaMemcpyAsync per batch, cuda_server.c:425-461)
 * cap the end-to-end rate well below what the kernels sustain. */
fr_status fr_infer_many(fr_engine* e, const int32_t* idx, int n, int B, float* scores, fr_stream s);
/* Parity hook: the concat vectors [B][concat_floats] exactly as the FPGA would
 * put them on the wire (embedding_47_krnl.cpp:774 byte stream, item-major). */
fr_status fr_gather_only(fr_engine* e, const int32_t* idx, int B, float* concat, fr_stream s);
/* B3 alone: x is what cuda_server.c:425-461 receives, [B][concat_floats] fp32. */
fr_status fr_mlp_only(fr_engine* e, const float* x, int B, float* scores, fr_stream s);
/* One launch of the MLP chain in isolation (unit-test hook): step k of the current
 * precision's chain on x [B][in_k].  TF32: k = 0,1 -> y [B][out_k] (tf32-rounded
 * activations), k = 2 -> y [B] scores (layer 3 with the output layer folded in).
 * FP32: k = 0..2 -> y [B][out_k], k = 3 -> y [B] scores. */
fr_status fr_layer_only(fr_engine* e, int k, const float* x, int B, float* y, fr_stream s);
fr_status fr_sync(fr_engine* e, fr_stream s);

/* ---- introspection ------------------------------------------------------ */
/* Number of kernels this library has launched on this engine so far. */
int64_t fr_launch_count(const fr_engine* e);
/* CUDA-graph cache of the hot-path calls: steps replayed from a graph, steps that were captured (and then launched as
 * a graph) for the first time, steps issued as plain launches (graphs off, first call of a batch size on the
 * engine, pageable buffers).  A steady-state loop shows only `replayed` growing. */
fr_status fr_graph_stats(const fr_engine* e, int64_t* replayed, int64_t* captured, int64_t* direct);
/* Forget the graphs cached on a worker (its caller buffers are about to be freed or re-used differently). */
fr_status fr_graph_flush(fr_engine* e, fr_stream s);
/* Device bytes currently held by tables / by everything. */
int64_t fr_table_bytes(const fr_engine* e);
/* Device time (ms) between two marks on a worker stream: fr_mark(s, 0) ...
 * fr_mark(s, 1); fr_elapsed_ms() syncs on mark 1.  CUDA events on the stream the
 * kernels are launched on. */
fr_status fr_mark(fr_engine* e, fr_stream s, int which);
fr_status fr_elapsed_ms(fr_engine* e, fr_stream s, float* ms);
/* Per-kernel device time for one batch: every kernel of the step is launched
 * `reps` times back to back, alone, between two CUDA events on the worker's own
 * stream (after one untimed pass).  ms5 = average ms per launch of {gather,
 * layer 1, layer 2, layer 3 (+output layer when fused), output layer (FP32 path
 * only, else 0)}.  Measurement hook for the roofline report (bench.py). */
fr_status fr_time_kernels(fr_engine* e, const int32_t* idx, int B, int reps, fr_stream s, float* ms5);

/* ---- table sharding across processes (one engine per GPU) ---------------- */
/* owner[t] in [0, world) = rank holding table t, or -1 = replicated on every
 * rank.  Must be called before tables are loaded; non-owned tables then take
 * no memory and fr_load_table / fr_fill_* on them are no-ops returning FR_OK. */
fr_status fr_shard_init(fr_engine* e, int rank, int world, const int* owner);
/* Exchange buffers: each rank exports one CUDA-IPC handle (64 bytes) for its
 * receive buffer, the host gathers all of them (any transport) and hands the
 * [world][64] array to every rank. */
fr_status fr_shard_export(fr_engine* e, void* handle64);
fr_status fr_shard_import(fr_engine* e, const void* handles /* [world][64] */);
/* In-process variant (tests, single-process multi-GPU): peers[r] = engine of rank r. */
fr_status fr_shard_attach_local(fr_engine* e, fr_engine* const* peers);
/* Sharded step, phase 1: gather the locally owned tables for the GLOBAL batch
 * idx [B_global][n_tables] and push every row piece straight into the concat
 * buffer of the rank that owns the item (rank r owns items [r*B_global/world,
 * (r+1)*B_global/world)) through NVLink peer stores.  Phase 2 (after the host
 * has barriered all ranks): MLP over the local items, scores [B_global/world]. */
fr_status fr_shard_gather_push(fr_engine* e, const int32_t* idx, int B_global, fr_stream s);
fr_status fr_shard_mlp(fr_engine* e, int B_global, float* scores_local, fr_stream s);
/* The same step as ONE asynchronous call with device-side synchronisation: push ->
 * publish a per-rank step flag into every peer's exchange region (by the push kernel's last block) ->
 * the first MLP kernel polls, right before its first load of the concat buffer, until all
 * ranks have published this step -> MLP.  No host barrier, no NCCL on the data path.
 * Every rank must issue the same sequence of calls with the same global batch.
 * world <= 32.  A peer that never arrives makes the wait give up after ~2 s (4e9 SM cycles): the step's
 * scores are then invalid, and fr_sync and every later sharded call return FR_ERR_STATE. */
fr_status fr_shard_infer(fr_engine* e, const int32_t* idx, int B_global, float* scores_local, fr_stream s);

/* The same step fed with column-sliced index blocks -- what every FPGA of the reference receives: only the indices
 * of its own tables (load_access_idx is instantiated per bank, embedding_47_krnl.cpp:899-914).  fr_shard_tables
 * returns, ascending (= the column order of the blocks), which = 0: the tables this rank owns, for which it needs
 * the indices of ALL B_global items; which = 1: the replicated tables, for which it needs the indices of ITS
 * B_global / world items only (ids may be null to query n).  A rank then uploads B_global * n_owned +
 * B_local * n_repl indices per step instead of B_global * T.  When idx_repl starts right behind idx_owned in one
 * host buffer (at the next multiple of 4 ints) both blocks travel in one copy. */
fr_status fr_shard_tables(fr_engine* e, int which, int32_t* ids, int* n);
fr_status fr_shard_infer_sliced(fr_engine* e, const int32_t* idx_owned /* [B_global][n_owned] */,
                                const int32_t* idx_repl /* [B_global / world][n_repl] */, int B_global,
                                float* scores_local, fr_stream s);
/* n consecutive sharded steps in one call on one worker: idx_owned [n][B_global][n_owned], idx_repl [n][B_global / world]
 * [n_repl], scores_local [n][B_global / world], each contiguous; host blocks travel in one copy each way (the same
 * adjacency rule).  Every rank must call it with the same n.  What fr_infer_many is to fr_infer. */
fr_status fr_shard_infer_sliced_many(fr_engine* e, const int32_t* idx_owned, const int32_t* idx_repl, int n, int B_global,
                                     float* scores_local, fr_stream s);
/* Local concat buffer after the exchange (parity hook), [B_global/world][concat_floats]. */
fr_status fr_shard_read_concat(fr_engine* e, int B_global, float* concat_local, fr_stream s);

/* ---- Cartesian-merged tables (MicroRec), SURVEY.md section 8c(b) ---------- */
/* merged[iA*rowsB + iB] = A[iA] || B[iB]; index of the merged row, in int64. */
int64_t fr_merge_index(int64_t iA, int64_t iB, int64_t rowsB);
/* Build the merged table on the device from two loaded tables and install it
 * as table `dst_table` (whose desc must have rows = rowsA*rowsB, dim = dimA+dimB). */
fr_status fr_merge_tables(fr_engine* e, int table_a, int table_b, int dst_table);

/* ---- request-driven batching front-end, SURVEY.md section 8(f)2 -------------- */
/* Replaces the reference's fixed batch hand-out (cuda_server.c:23-25,406-417: THREAD_NUM workers
 * pull batch numbers off a mutex-guarded counter and block in read() until BATCH_SIZE items have
 * arrived, cuda_server.c:425-461) with a batch former: requests of any size from any thread are
 * packed into pinned staging buffers; a batch is dispatched to a worker stream when it holds
 * max_batch items or when its oldest request has waited max_delay_us.  Host threads only; the
 * device work is fr_infer's.  Not available on a table-sharded engine. */
typedef struct fr_batcher fr_batcher;
typedef struct fr_batcher_config {
  int max_batch;     /* items per dispatched batch, <= the engine's max_batch        */
  int max_delay_us;  /* the open batch is dispatched once its oldest request has waited this long    */
  int n_workers;     /* worker threads = fr_streams = batches in flight (THREAD_NUM) */
} fr_batcher_config;
typedef struct fr_batcher_stats {
  int64_t batches, items, requests, closed_by_deadline;
  float latency_p50_us, latency_p99_us;   /* submit -> scores written, last <= 65536 request parts */
} fr_batcher_stats;
fr_status fr_batcher_create(fr_engine* e, const fr_batcher_config* cfg, fr_batcher** out);
/* idx [n][n_tables] is copied before the call returns; scores_out[n] is written by a worker
 * thread and valid once fr_batcher_wait(ticket) returns.  n may exceed max_batch (split). */
fr_status fr_batcher_submit(fr_batcher* b, const int32_t* idx, int n, float* scores_out, uint64_t* ticket);
fr_status fr_batcher_wait(fr_batcher* b, uint64_t ticket);
/* Dispatch the open batch now, whatever it holds. */
fr_status fr_batcher_flush(fr_batcher* b);
fr_status fr_batcher_get_stats(fr_batcher* b, fr_batcher_stats* out);
/* Scores everything already submitted, then stops the threads and frees the staging buffers. */
void fr_batcher_destroy(fr_batcher* b);

/* ---- B2-compatible streaming ingest, SURVEY.md section 8(f)3 ------------------ */
/* Accepts the byte stream the reference's senders produce (sendData() embedding_47_krnl.cpp:45-147,
 * multiple_connections_network_client_sender.c:55-100): connection i on base_port + i, raw
 * little-endian fp32, exactly batch * concat_floats * 4 bytes per batch, no header -- what
 * thread_consume() read()s (cuda_server.c:360-461, constant.h:33-41) -- and runs fr_mlp_only on every
 * block; batch numbers come off one counter shared by all connections (global_batch_count).
 * FR_INGEST_INDICES takes batch * n_tables int32 per block instead and runs fr_infer. */
enum { FR_INGEST_CONCAT = 0, FR_INGEST_INDICES = 1 };
typedef struct fr_ingest fr_ingest;
typedef struct fr_ingest_config {
  int base_port;                 /* PORT (constant.h:39)                                        */
  int n_conn;                    /* THREAD_NUM: listeners, one accepted connection each         */
  int batch;                     /* BATCH_SIZE                                                  */
  int payload;                   /* FR_INGEST_*                                                 */
  int64_t total_batches;         /* TOTAL_BATCH_NUM over all connections; 0 = until senders close */
  int listen_any;                /* 0 (default): bind 127.0.0.1 only; 1: INADDR_ANY like the reference
                                    (cuda_server.c:379) -- the socket is unauthenticated            */
  float* scores_out;             /* optional sink [n_conn][max_batches_per_conn][batch]: scores of
                                    connection c's k-th block (the reference discards all but the
                                    first outputs, cuda_server.c:499-502)                       */
  int64_t max_batches_per_conn;
} fr_ingest_config;
typedef struct fr_ingest_stats {
  int64_t batches, bytes;
  double seconds;                /* start -> last connection finished                           */
  int connections;
} fr_ingest_stats;
/* Binds and listens on every port before returning, then accepts/receives on its own threads. */
fr_status fr_ingest_start(fr_engine* e, const fr_ingest_config* cfg, fr_ingest** out);
/* Blocks until every connection has ended (sender closed, or total_batches reached). */
fr_status fr_ingest_wait(fr_ingest* g, fr_ingest_stats* stats);
fr_status fr_ingest_last_scores(fr_ingest* g, int conn, float* scores /* [batch] */, int64_t* batch_no);
void fr_ingest_destroy(fr_ingest* g);

#ifdef __cplusplus
}
#endif
#endif /* FLEETREC_H */
