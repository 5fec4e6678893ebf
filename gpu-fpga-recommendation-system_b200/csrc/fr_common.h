// fr_common.h -- internal engine structures shared by the CUDA translation units.
// Nothing here is part of the ABI (include/fleetrec.h is).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/fleetrec.h"

#define FR_MAX_LAYERS 4

// One 16-byte piece of an item's concat vector: out4[b][c] = base[idx[b][table]*stride4 + col4].
// (32 bytes, 16-byte aligned: the lookup kernel fetches a descriptor with two 128-bit loads)
struct alignas(16) FrChunk {
  const float4* base;  // table base (device), NULL when the table is not resident on this rank
  int table;           // column of idx[][] to read
  int stride4;         // row pitch in float4 (= dim/4)
  int col4;            // float4 offset inside the row
  int rows;            // rows of the table (clamped to INT_MAX): bound of the optional index check
  int idx_off;         // byte offset of the table's index inside an index row; bit 31 set: the index is a uint16
  int pad_;
};
constexpr int kIdx16 = (int)0x80000000;
static_assert(sizeof(FrChunk) == 32, "FrChunk is two 16-byte words");

// One index out of an index row (`row` = its first int32 word; rows are whole int32 words in either format).  ONE aligned
// 32-bit load whatever the width -- a uint16 column sits at an even byte offset, i.e. in the low or high half of its word
// -- so a warp whose pieces mix both widths still issues a single load instruction (the lookup is bound by load/store
// instructions on the small model: two predicated loads per index cost 4.2 -> 5.2 us per batch of 2048).
#ifdef __CUDACC__
__host__ __device__ __forceinline__
#else
inline
#endif
int64_t fr_index_at(const int32_t* row, int idx_off) {
  const int off = idx_off & 0x7FFFFFFF;
  const int32_t* p = row + (off >> 2);
#ifdef __CUDA_ARCH__
  const int32_t w = __ldg(p);
#else
  const int32_t w = *p;
#endif
  return idx_off < 0 ? (int64_t)(((uint32_t)w >> ((off & 2) * 8)) & 0xFFFFu) : (int64_t)w;
}

// The same piece, compact (16 bytes), for the lookup fused into layer 1 (staged in shared memory).
struct FrFuseChunk {
  const float4* base;
  int table;
  int stride4_col4;    // (row pitch in float4) << 8 | float4 offset inside the row
};

struct FrTable {
  float* d = nullptr;    // device image; elements are fp32, or 2-byte f16 / bf16 when the engine's table_dtype says so
  int64_t rows = 0;
  int dim = 0;
  int tier = 0;
  bool resident = true;  // false: owned by another rank
  bool loaded = false;
  // max |x| and min non-zero |x| of the image (fr_precision.cu), valid until the table is written again
  float maxabs = 0.f, minabs = 0.f;
  bool range_valid = false;
};

struct fr_stream_s {
  cudaStream_t stream = nullptr;
  int32_t* d_idx = nullptr;   // [max_batch][T] (grown by the *_many calls: idx_cap ints)
  size_t idx_cap = 0, scores_cap = 0;   // capacity of d_idx (ints) / d_scores (floats)
  float* d_x = nullptr;       // [max_batch][D]      concat activations
  std::vector<void*> retired;  // staging buffers outgrown by a *_many call (freed with the worker)
  float* d_h[3] = {nullptr, nullptr, nullptr};  // [max_batch][hidden k]
  float* d_scores = nullptr;  // [max_batch]
  cudaEvent_t ev[2] = {nullptr, nullptr};
  // One batch = H2D? + gather + 3..4 GEMM launches + D2H?: replayed as ONE cudaGraphLaunch once a
  // (idx, scores, B, mode) combination has been seen twice on this worker (launch-bound otherwise).
  struct Graph {
    const void* idx;
    const void* idx2;       // second index block of the column-sliced sharded step (else null)
    const void* scores;
    int B, mode, prec;
    int variant;            // FR_GV_*: which entry point (and how many sub-batches) the graph replays
    int flavours;           // 1, or 2 for the sharded steps (one graph per exchange-buffer parity, captured together)
    bool failed;            // capture / instantiation failed once: this combination runs un-graphed
    int launches;           // kernels inside one graph
    uint64_t last_use;      // LRU stamp
    cudaGraphExec_t exec[2];
  };
  std::vector<Graph> graphs;
  uint64_t graph_clock = 0;
  // table-sharded steps: which exchange slot this worker owns (creation order, identical on every
  // rank) and how many sharded steps it has issued (parity of the concat buffer = step & 1)
  int slot = 0;
  int shard_step = 0;
  bool f16 = false;   // the step being enqueued runs the tcgen05 MLP on fp16 operands (set by the entry point)
};

struct FrPeer {
  float* concat = nullptr;  // peer's exchange buffer (mapped into this process)
  bool ipc = false;
};

// Tuning / test knobs of one engine, read ONCE from the environment by fr_create (fr_read_knobs, fr_api.cu).  A
// release build reads only the two test hooks that pin the tcgen05 tile shapes; everything else -- the measured-slower
// kernel variants and their switches -- exists only in a library built with -DFR_EXPERIMENTS
// (`make exp` -> libfleetrec_exp.so).
struct FrKnobs {
  // test hooks (release and experiments builds)
  bool tiles_pinned = false;      // FR_TC_TILES=N1,N2,N3[,ctas]: tile width of layers 1..3, CTAs per tile
  int tiles[3] = {256, 256, 256};
  int tile_ctas = 2;
  int max_clusters = 0;           // FR_TC_MAX_CLUSTERS: cap the persistent grids (several tiles per cluster)
  // tuning
  int min_kb = 32;                // K slices of main loop a cluster should get before short tiles stop being ganged
  // experiments (FR_EXPERIMENTS builds only; the defaults below are what a release build runs)
  int pdl_mask = 0;               // FR_PDL: programmatic dependent launch edges (bit 0 GEMM->GEMM, 1 lookup->layer 1, 2 layer 3->lookup)
  int zero_copy_pct = 0;          // FR_ZEROCOPY: share of a pinned index batch the SMs fetch over PCIe themselves
  bool mcast = false;             // FR_TC_MCAST: 4-CTA clusters multicasting the weight slices
  bool a_lsu = false;             // FR_TC_ALSU: A operand through cp.async instead of TMA
  bool chain = false;             // FR_CHAIN: the whole MLP as one persistent launch
  bool chain_prof = false;        // FR_CHAIN_PROF: phase timeline of the chain kernel's CTA 0
  // FR_SHARD_FOLD=1: the first MLP kernel of a sharded step polls the peers' flags itself instead of following a
  // 1-warp wait kernel.  Off, and not offered in release builds: a persistent tcgen05 grid that spins holds its SMs
  // (shared memory), so with several worker streams in flight two ranks that happen to start different workers'
  // layer-1 kernels first wait for each other's exchange kernels behind launches that cannot get an SM -- measured:
  // 15.7 instead of 13.x us per step at two ranks (small model), time-outs on the large model (148-CTA grids).
  int shard_fold_wait = 0;
  int dbg_nostore = 0;            // FR_TC_NOSTORE: storing epilogues skip their stores (timing experiments; results are garbage)
  bool tc_prof = false;           // FR_TC_PROF: cycle counters of the per-layer kernel's pipelines (tools/tc_prof.py)
};

struct fr_engine {
  int device = 0;
  FrKnobs knobs;
  int sm_count = 0;
  std::string name;
  std::vector<fr_table_desc> tdesc;
  std::vector<fr_segment_desc> segs;
  int D = 0;  // concat floats
  int dims[FR_MAX_LAYERS + 1] = {0, 0, 0, 0, 0};
  int mlp_mode = FR_MLP_BIAS_RELU_SIGMOID;
  int precision = FR_PREC_TF32;
  int table_dtype = FR_TABLE_F32;
  int max_batch = 0;
  bool use_graphs = true;  // FR_OPT_CUDA_GRAPHS
  std::vector<FrTable> tables;
  FrChunk* d_chunks = nullptr;  // [D/4]
  FrFuseChunk* d_fchunks = nullptr;  // [D/4]
  // FR_OPT_FUSE_LOOKUP: fr_infer gathers straight into layer 1's A tile (no concat in global memory).  Off by default:
  // parity-green but slower on B200.
  bool fuse_lookup = false;
  bool chunks_dirty = true;

  float* d_W[FR_MAX_LAYERS] = {nullptr, nullptr, nullptr, nullptr};    // [in][out] fp32 (reference layout)
  float* d_Wt[FR_MAX_LAYERS] = {nullptr, nullptr, nullptr, nullptr};   // [out][in] tf32-rounded (tcgen05 B operand)
  void* d_Wt16[FR_MAX_LAYERS] = {nullptr, nullptr, nullptr, nullptr};  // the same as fp16 (tc_f16)
  bool tc_f16 = false;            // the decision: fr_infer computes on fp16 operands
  int f16_mode = FR_F16_OFF;      // FR_OPT_F16_OPERANDS
  bool f16_dirty = true;          // tables / weights / the option changed since the range analysis last ran
  float f16_bounds[5] = {0, 0, 0, 0, 0};
  float* d_bias[FR_MAX_LAYERS] = {nullptr, nullptr, nullptr, nullptr};
  bool layer_loaded[FR_MAX_LAYERS] = {false, false, false, false};
  void* tc_state = nullptr;  // tensor maps etc., owned by fr_mlp_tc.cu
  int tc_last_ctas = 0;      // grid of the last tcgen05 GEMM launch, and of the last launch of every MLP step
  int tc_layer_ctas[4] = {0, 0, 0, 0};

  fr_stream_s* default_stream = nullptr;
  std::vector<fr_stream_s*> streams;
  std::mutex mu;

  // sharding
  int rank = 0, world = 1;
  std::vector<int> owner;        // [T], -1 replicated
  float* d_xchg = nullptr;       // exchange region: n_slots x { concat[2][max_batch/world][D], flags[world] }
  int n_slots = 0;               // in-flight sharded steps (one per worker stream)
  int next_slot = 0;             // slot handed to the next stream created
  int* d_step = nullptr;         // [n_slots] device-side step counters (the flag kernel increments its slot's)
  std::vector<FrPeer> peers;     // [world]
  float** d_peer_ptrs = nullptr; // device copy of peers[].concat
  int* d_owned_ids = nullptr;    // concat pieces this rank produces for every item
  int n_owned = 0;
  int* d_repl_ids = nullptr;     // pieces of replicated tables (local items only)
  int n_repl = 0;
  // column-sliced index blocks (fr_shard_infer_sliced): the tables this rank owns / the replicated ones, ascending
  // = the column order of the caller's blocks, and the piece descriptors with `table` renumbered to those columns
  std::vector<int> owned_tables, repl_tables;
  FrChunk* d_chunks_sliced = nullptr;
  bool shard_lists_built = false;
  int* h_shard_err = nullptr;    // pinned+mapped: set to 1 by the wait kernel on time-out
  int* h_watch = nullptr;        // pinned+mapped int[8]: which barrier wait a tcgen05 kernel gave up on before trapping

  std::atomic<int64_t> launches{0};
  // CUDA-graph cache introspection (fr_graph_stats): steps replayed from a graph, steps that had to be captured
  // first, steps issued as plain launches (graphs off, cache bypass, un-capturable buffers)
  std::atomic<int64_t> graph_hits{0}, graph_captures{0}, graph_direct{0};
  std::vector<std::pair<int, int>> warmed;   // (entry point, B) pairs that have run un-captured once (guarded by mu)
  // fr_set_check_indices: the lookup kernels compare every index with its table's row count, count the offenders in
  // h_idx_err (pinned + mapped: {count, table column, value, item}) and read row 0 instead; fr_sync reports them
  // FR_OPT_INDEX_FORMAT: layout of the index rows the hot-path calls are given (fr_index_rows() below)
  int index_format = FR_IDX_I32;
  std::vector<int> idx_off_full, idx_off_owned, idx_off_repl;   // per column: byte offset | kIdx16
  int ipr_full = 0, ipr_owned = 0, ipr_repl = 0;                  // int32 words per row (rows are padded to 4 bytes)
  int tile_hint = 0;              // FR_HINT_*: latency- or throughput-oriented tcgen05 tiles (fr_set_option)
  bool check_indices = false;
  int* h_idx_err = nullptr;
  mutable std::string err;
};

fr_status fr_fail(const fr_engine* e, fr_status code, const char* fmt, ...);
// fr_infer with the CUDA-graph cache bypassed on request (one-shot buffer / batch-size combinations)
fr_status fr_infer_opts(fr_engine* e, const int32_t* idx, int B, float* scores, fr_stream s, bool no_graph);
// the tcgen05 path runs on fp16 operands (FR_TC_F16=1 and the engine's precision is the tensor-core one)
inline bool fr_tc_f16(const fr_engine* e) {
  return e->tc_f16 && e->f16_mode == FR_F16_GUARDED && e->precision == FR_PREC_TF32 && e->world == 1;
}
#define FR_CUDA(e, call)                                                                          \
  do {                                                                                            \
    cudaError_t err__ = (call);                                                                   \
    if (err__ != cudaSuccess)                                                                     \
      return fr_fail((e), err__ == cudaErrorMemoryAllocation ? FR_ERR_OOM : FR_ERR_CUDA,          \
                     "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

// Host->device upload that is COMPLETE on return.  A plain cudaMemcpy from pageable memory may
// return once the bytes sit in the driver's staging buffer (the DMA is then ordered on the legacy
// stream only), and the worker streams are cudaStreamNonBlocking -- a kernel launched right after
// could read the old contents.  So every upload is enqueued on the engine's own stream and waited for.
inline cudaError_t fr_h2d(fr_engine* e, void* dst, const void* src, size_t bytes) {
  cudaError_t err = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, e->default_stream->stream);
  if (err != cudaSuccess) return err;
  return cudaStreamSynchronize(e->default_stream->stream);
}

// ---- kernels (each returns after enqueueing; bumps e->launches) -----------
fr_status frk_upload_chunks(fr_engine* e);
fr_status frk_gather(fr_engine* e, const int32_t* d_idx, int B, float* d_out, bool round_tf32, cudaStream_t st,
                     bool out_f16 = false);
// fp32 -> fp16 (round to nearest even), n elements (n % 4 == 0)
fr_status frk_to_f16(fr_engine* e, const float* src, void* dst, int64_t n, cudaStream_t st);
// fp32 -> tf32 (round to nearest, ties away: cvt.rna), n elements (n % 4 == 0); dst may alias src
fr_status frk_round_tf32(fr_engine* e, const float* src, float* dst, int64_t n, cudaStream_t st);
// copy `bytes` (a multiple of 16) of indices from a mapped page-locked host buffer into device memory with SM loads
fr_status frk_stage_idx(fr_engine* e, const void* mapped_src, int32_t* d_dst, size_t bytes, cudaStream_t st);
// The exchange of a table-sharded step: this rank's owned pieces of every item of the global batch are stored into
// the concat buffer (slot, parity) of the rank that owns the item, the replicated pieces of its own items into its
// own, then a one-warp kernel publishes the step to every peer and (wait = true) waits for theirs.  idx_owned
// [B_global][T_owned], idx_repl [B_global / world][T_repl] (row 0 = this rank's first item); chunks[].table names the
// column in those blocks.
fr_status frk_shard_exchange(fr_engine* e, const FrChunk* chunks, const int32_t* d_idx_owned, int T_owned,
                             const int32_t* d_idx_repl, int T_repl, int B_global, int slot, int parity, bool wait, cudaStream_t st);
const FrChunk* frk_sliced_chunks(fr_engine* e);   // descriptors whose `table` is the column of a column-sliced block
void fr_shard_table_lists(fr_engine* e);   // fills owned_tables / repl_tables from owner[] (idempotent)
// (re)computes idx_off_* / ipr_* for the engine's index format and table lists
void fr_index_rows(fr_engine* e);
// wait (on the device) until every rank has published the slot's current step; the tcgen05 path does this inside
// its first kernel instead (FrPeerWait)
fr_status frk_shard_wait(fr_engine* e, int slot, cudaStream_t st);
// a peer that has not published after this many SM cycles (~2 s) is given up on: h_shard_err is set, fr_sync reports it
constexpr long long kShardTimeoutCycles = 4000000000ll;
// what the first MLP kernel of a sharded step polls before it touches the concat buffer
struct FrPeerWait {
  const int* flags;   // this rank's flag block of the slot: flags[r] = last step rank r has pushed
  const int* step;    // the slot's step counter (already bumped by this rank's exchange kernel)
  int world;
  int* err;           // device alias of h_shard_err
};

// exchange-region geometry (floats): one slot = two concat buffers + a flag block
inline size_t fr_xchg_buf_floats(const fr_engine* e) { return (size_t)(e->max_batch / e->world) * e->D; }
inline size_t fr_xchg_slot_floats(const fr_engine* e) { return 2 * fr_xchg_buf_floats(e) + 64; }
inline size_t fr_xchg_concat_off(const fr_engine* e, int slot, int parity) {
  return (size_t)slot * fr_xchg_slot_floats(e) + (size_t)parity * fr_xchg_buf_floats(e);
}
inline size_t fr_xchg_flags_off(const fr_engine* e, int slot) {
  return (size_t)slot * fr_xchg_slot_floats(e) + 2 * fr_xchg_buf_floats(e);
}
// element size of the table storage type
inline size_t fr_table_esize(const fr_engine* e) {
  return e->table_dtype == FR_TABLE_F32 ? 4 : (e->table_dtype == FR_TABLE_FP8 ? 1 : 2);
}
fr_status frk_fill_reference(fr_engine* e, float* d, int64_t rows, int dim, int64_t debug_rows, cudaStream_t st);
fr_status frk_fill_hash(fr_engine* e, float* d, uint32_t seed, int table, int64_t rows, int dim, cudaStream_t st);
// fp32 -> table storage type (round to nearest even) and back (exact), n elements
fr_status frk_quantize(fr_engine* e, const float* src, void* dst, int64_t n, cudaStream_t st);
fr_status frk_dequantize(fr_engine* e, const void* src, float* dst, int64_t n, cudaStream_t st);
fr_status frk_merge(fr_engine* e, const float* A, int64_t rowsA, int dimA, const float* B, int64_t rowsB, int dimB,
                    float* M, cudaStream_t st);
fr_status frk_transpose_round_tf32(fr_engine* e, const float* W, int in, int out, float* Wt, cudaStream_t st);

// FP32 SIMT path: Y[B][N] = act(X[B][K] . W[K][N] + bias)
fr_status frk_sgemm_bias_act(fr_engine* e, const float* X, const float* W, const float* bias, float* Y, int B, int K,
                             int N, bool relu, cudaStream_t st);
// final 1-wide layer: scores[b] = (sigmoid?)(dot(H[b][0..K), w) + bias0)
fr_status frk_final_dot(fr_engine* e, const float* H, const float* w, const float* bias, float* scores, int B, int K,
                        bool sigmoid, cudaStream_t st);

// TF32 tcgen05 path (fr_mlp_tc.cu)
fr_status frtc_prepare(fr_engine* e);                // builds tensor maps for weights; idempotent
fr_status frtc_prepare_f16(fr_engine* e);            // tensor maps of the fp16 weight copies (after frtc_prepare)
// FR_F16_GUARDED: bound every operand of the MLP from the loaded tables and weights, decide e->tc_f16 (fr_precision.cu)
fr_status fr_f16_analyse(fr_engine* e);
void frtc_destroy(fr_engine* e);
fr_status frtc_layer(fr_engine* e, fr_stream_s* s, int k, const float* in, int B, float* d_scores,
                     const FrPeerWait* wait = nullptr);   // wait: poll the peers' step flags before the first load of `in`
// the whole MLP (3 GEMMs + output layer) as ONE persistent launch: in [B][dims[0]] -> d_scores [B]; uses s->d_h[0..1]
bool frtc_can_chain(const fr_engine* e, int B);
fr_status frtc_chain(fr_engine* e, fr_stream_s* s, const float* in, int B, float* d_scores);
// lookup fused into layer 1: d_idx [B][T] -> s->d_h[0]; frtc_can_fuse() says whether this engine's shapes allow it
bool frtc_can_fuse(const fr_engine* e);
fr_status frtc_fused_layer1(fr_engine* e, fr_stream_s* s, const int32_t* d_idx, int B);
