// fr_api.cu -- the C ABI of include/fleetrec.h: engine lifetime, table and weight
// residency in HBM, worker streams, and the hot-path entry points.
//
// Host-side shape of the reference this replaces:
//   FPGA/host/embedding_47_krnl/host.cpp:324-761  (allocate + init tables, migrate, launch)
//   GPU/final_network_cublasLt_1_node_no_FIFO_scatter/cuda_server.c:101-183,346-354,406-495
//                                                  (per-worker buffers + stream, weights H2D, batch loop)
#include <limits.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "fr_common.h"

// ---------------------------------------------------------------------------
// built-in catalogues (generated from the reference's constants.hpp)
struct fr_builtin_model {
  const char* name;
  int n_tables;
  const fr_table_desc* tables;
  int n_segments;
  const fr_segment_desc* segments;
  int concat_floats;
  int hidden[4];
};
#include "fr_catalogue_data.inc"

static thread_local std::string g_tls_err;

// The ONLY place the library reads the environment (once per fr_create).
static void fr_read_knobs(FrKnobs* k) {
  if (const char* env = getenv("FR_TC_TILES")) {
    sscanf(env, "%d,%d,%d,%d", &k->tiles[0], &k->tiles[1], &k->tiles[2], &k->tile_ctas);
    k->tiles[2] = 256;   // layer 3 keeps its whole row (the output layer is folded into its epilogue)
    k->tiles_pinned = true;
  }
  if (const char* env = getenv("FR_TC_MAX_CLUSTERS")) k->max_clusters = atoi(env);
#ifdef FR_EXPERIMENTS
  auto num = [](const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; };
  k->min_kb = num("FR_TC_MIN_KB", k->min_kb) > 0 ? num("FR_TC_MIN_KB", k->min_kb) : 1;
  k->pdl_mask = num("FR_PDL", 0) & 7;
  const int zc = num("FR_ZEROCOPY", 0);
  k->zero_copy_pct = zc == 1 ? 100 : (zc < 0 ? 0 : (zc > 100 ? 100 : zc));
  k->mcast = num("FR_TC_MCAST", 0) != 0;
  k->a_lsu = num("FR_TC_ALSU", 0) != 0;
  k->chain = num("FR_CHAIN", 0) != 0;
  k->chain_prof = num("FR_CHAIN_PROF", 0) != 0;
  k->tc_prof = num("FR_TC_PROF", 0) != 0;
  k->dbg_nostore = num("FR_TC_NOSTORE", 0);
  k->shard_fold_wait = num("FR_SHARD_FOLD", 0);
#endif
}

// 1 when the library was built with -DFR_EXPERIMENTS (the measured-slower kernel variants and their switches exist)
extern "C" int fr_build_has_experiments(void) {
#ifdef FR_EXPERIMENTS
  return 1;
#else
  return 0;
#endif
}

fr_status fr_fail(const fr_engine* e, fr_status code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (e) e->err = buf;
  g_tls_err = buf;
  return code;
}

extern "C" const char* fr_last_error(const fr_engine* e) { return e ? e->err.c_str() : g_tls_err.c_str(); }

extern "C" fr_status fr_model_builtin(const char* name, fr_model_desc* out) {
  if (!name || !out) return fr_fail(nullptr, FR_ERR_INVALID, "fr_model_builtin: null argument");
  for (const fr_builtin_model& m : k_builtin_models)
    if (strcmp(m.name, name) == 0) {
      out->name = m.name;
      out->n_tables = m.n_tables;
      out->tables = m.tables;
      out->n_segments = m.n_segments;
      out->segments = m.segments;
      out->concat_floats = m.concat_floats;
      for (int k = 0; k < 4; k++) out->hidden[k] = m.hidden[k];
      out->mlp_mode = FR_MLP_BIAS_RELU_SIGMOID;
      out->precision = FR_PREC_TF32;
      out->max_batch = 16384;
      out->table_dtype = FR_TABLE_F32;
      return FR_OK;
    }
  return fr_fail(nullptr, FR_ERR_INVALID, "fr_model_builtin: unknown model '%s' (small|medium|large_half|large)", name);
}

// ---------------------------------------------------------------------------
static fr_status validate_desc(const fr_model_desc* d) {
  if (!d || !d->tables || !d->segments) return fr_fail(nullptr, FR_ERR_INVALID, "fr_create: null descriptor");
  if (d->n_tables <= 0 || d->n_segments <= 0) return fr_fail(nullptr, FR_ERR_INVALID, "fr_create: empty model");
  if (d->concat_floats <= 0 || d->concat_floats % 16)
    return fr_fail(nullptr, FR_ERR_INVALID, "fr_create: concat_floats=%d must be a positive multiple of 16 (one 512-bit "
                   "network word, constants.hpp:9)", d->concat_floats);
  if (d->max_batch <= 0) return fr_fail(nullptr, FR_ERR_INVALID, "fr_create: max_batch must be > 0");
  if (d->table_dtype < FR_TABLE_F32 || d->table_dtype > FR_TABLE_FP8)
    return fr_fail(nullptr, FR_ERR_INVALID, "fr_create: table_dtype %d (FR_TABLE_F32 | F16 | BF16 | FP8)", d->table_dtype);
  for (int t = 0; t < d->n_tables; t++) {
    const fr_table_desc& td = d->tables[t];
    if (td.dim <= 0 || td.dim % 4) return fr_fail(nullptr, FR_ERR_INVALID, "table %d: dim=%d must be a multiple of 4 "
                                                  "floats (one 128-bit axi word)", t, td.dim);
    if (td.rows <= 0) return fr_fail(nullptr, FR_ERR_INVALID, "table %d: rows must be > 0", t);
  }
  for (int s = 0; s < d->n_segments; s++) {
    const fr_segment_desc& sg = d->segments[s];
    if (sg.table < 0 || sg.table >= d->n_tables) return fr_fail(nullptr, FR_ERR_INVALID, "segment %d: bad table", s);
    if (sg.len <= 0 || sg.len % 4 || sg.col % 4 || sg.dst % 4 || sg.col < 0 || sg.dst < 0 ||
        sg.col + sg.len > d->tables[sg.table].dim || sg.dst + sg.len > d->concat_floats)
      return fr_fail(nullptr, FR_ERR_INVALID, "segment %d: (dst=%d col=%d len=%d) not 4-float aligned or out of range", s,
                     sg.dst, sg.col, sg.len);
  }
  if (d->hidden[3] != 1) return fr_fail(nullptr, FR_ERR_UNSUPPORTED, "output layer must be 1 wide (OUTPUT_SIZE)");
  for (int k = 0; k < 3; k++)
    if (d->hidden[k] <= 0 || d->hidden[k] % 128)
      return fr_fail(nullptr, FR_ERR_UNSUPPORTED, "hidden[%d]=%d must be a multiple of 128", k, d->hidden[k]);
  return FR_OK;
}

static void free_stream(fr_stream_s* s);
static fr_status alloc_stream_parts(fr_engine* e, fr_stream_s* s) {
  const size_t mb = (size_t)e->max_batch;
  FR_CUDA(e, cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  FR_CUDA(e, cudaMalloc(&s->d_idx, mb * e->tables.size() * sizeof(int32_t)));
  FR_CUDA(e, cudaMalloc(&s->d_x, mb * e->D * sizeof(float)));
  for (int k = 0; k < 3; k++) FR_CUDA(e, cudaMalloc(&s->d_h[k], mb * e->dims[k + 1] * sizeof(float)));
  FR_CUDA(e, cudaMalloc(&s->d_scores, mb * sizeof(float)));
  s->idx_cap = mb * e->tables.size();
  s->scores_cap = mb;
  FR_CUDA(e, cudaEventCreate(&s->ev[0]));
  FR_CUDA(e, cudaEventCreate(&s->ev[1]));
  return FR_OK;
}

static fr_status alloc_stream(fr_engine* e, fr_stream_s** out) {
  fr_stream_s* s = new fr_stream_s();
  const fr_status st = alloc_stream_parts(e, s);
  if (st != FR_OK) {   // whatever was allocated before the failing call goes back
    free_stream(s);
    return st;
  }
  s->slot = e->next_slot++;   // creation order: identical on every rank of a sharded job
  *out = s;
  return FR_OK;
}

static void free_stream(fr_stream_s* s) {
  if (!s) return;
  if (s->stream) cudaStreamSynchronize(s->stream);
  for (fr_stream_s::Graph& g : s->graphs)
    for (int f = 0; f < 2; f++)
      if (g.exec[f]) cudaGraphExecDestroy(g.exec[f]);
  cudaFree(s->d_idx);
  cudaFree(s->d_x);
  for (void* p : s->retired) cudaFree(p);
  for (int k = 0; k < 3; k++) cudaFree(s->d_h[k]);
  cudaFree(s->d_scores);
  if (s->ev[0]) cudaEventDestroy(s->ev[0]);
  if (s->ev[1]) cudaEventDestroy(s->ev[1]);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

extern "C" fr_status fr_create(const fr_model_desc* desc, int n_gpus, const int* device_ids, fr_engine** out) {
  if (!out) return fr_fail(nullptr, FR_ERR_INVALID, "fr_create: out is null");
  *out = nullptr;
  fr_status st = validate_desc(desc);
  if (st != FR_OK) return st;
  if (n_gpus != 1 || !device_ids)
    return fr_fail(nullptr, FR_ERR_INVALID, "fr_create: one engine drives one GPU (n_gpus must be 1); shard across "
                   "processes with fr_shard_init");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fr_fail(nullptr, FR_ERR_CUDA, "fr_create: no CUDA device (%s); there is no CPU fallback",
                   ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0");
  const int dev = device_ids[0];
  if (dev < 0 || dev >= ndev) return fr_fail(nullptr, FR_ERR_INVALID, "fr_create: device %d of %d", dev, ndev);
  cudaDeviceProp prop;
  if ((ce = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess)
    return fr_fail(nullptr, FR_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(ce));
  if (prop.major != 10)
    return fr_fail(nullptr, FR_ERR_CUDA, "fr_create: device %d is sm_%d%d; this library is built for sm_100a only", dev,
                   prop.major, prop.minor);
  if ((ce = cudaSetDevice(dev)) != cudaSuccess)
    return fr_fail(nullptr, FR_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(ce));

  fr_engine* e = new fr_engine();
  e->device = dev;
  e->sm_count = prop.multiProcessorCount;
  e->name = desc->name ? desc->name : "custom";
  e->tdesc.assign(desc->tables, desc->tables + desc->n_tables);
  e->segs.assign(desc->segments, desc->segments + desc->n_segments);
  e->D = desc->concat_floats;
  e->dims[0] = e->D;
  for (int k = 0; k < 4; k++) e->dims[k + 1] = desc->hidden[k];
  e->mlp_mode = desc->mlp_mode;
  e->precision = desc->precision;
  e->table_dtype = desc->table_dtype;
  e->max_batch = desc->max_batch;
  fr_read_knobs(&e->knobs);
  e->tables.resize(desc->n_tables);
  for (int t = 0; t < desc->n_tables; t++) {
    e->tables[t].rows = desc->tables[t].rows;
    e->tables[t].dim = desc->tables[t].dim;
    e->tables[t].tier = desc->tables[t].tier;
  }
  fr_index_rows(e);
  st = alloc_stream(e, &e->default_stream);
  if (st != FR_OK) {
    g_tls_err = e->err;
    fr_destroy(e);
    return st;
  }
  *out = e;
  return FR_OK;
}

extern "C" void fr_destroy(fr_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  frtc_destroy(e);
  for (fr_stream_s* s : e->streams) free_stream(s);
  free_stream(e->default_stream);
  for (FrTable& t : e->tables) cudaFree(t.d);
  for (int k = 0; k < FR_MAX_LAYERS; k++) {
    cudaFree(e->d_W[k]);
    cudaFree(e->d_Wt[k]);
    cudaFree(e->d_Wt16[k]);
    cudaFree(e->d_bias[k]);
  }
  cudaFree(e->d_chunks);
  cudaFree(e->d_fchunks);
  cudaFree(e->d_owned_ids);
  cudaFree(e->d_repl_ids);
  cudaFree(e->d_chunks_sliced);
  for (int r = 0; r < (int)e->peers.size(); r++)
    if (e->peers[r].ipc && e->peers[r].concat) cudaIpcCloseMemHandle(e->peers[r].concat);
  cudaFree(e->d_peer_ptrs);
  cudaFree(e->d_xchg);
  cudaFree(e->d_step);
  if (e->h_shard_err) cudaFreeHost(e->h_shard_err);
  if (e->h_idx_err) cudaFreeHost(e->h_idx_err);
  if (e->h_watch) cudaFreeHost(e->h_watch);
  delete e;
}

// ---------------------------------------------------------------------------
static fr_status check_table(fr_engine* e, int t) {
  if (!e) return fr_fail(nullptr, FR_ERR_INVALID, "null engine");
  if (t < 0 || t >= (int)e->tables.size()) return fr_fail(e, FR_ERR_INVALID, "table id %d out of range", t);
  return FR_OK;
}

static fr_status ensure_table_mem(fr_engine* e, int t) {
  FrTable& tb = e->tables[t];
  if (tb.d) return FR_OK;
  FR_CUDA(e, cudaSetDevice(e->device));
  FR_CUDA(e, cudaMalloc(&tb.d, (size_t)tb.rows * tb.dim * fr_table_esize(e)));
  e->chunks_dirty = true;
  return FR_OK;
}

extern "C" fr_status fr_set_table_rows(fr_engine* e, int table_id, int64_t rows) {
  fr_status st = check_table(e, table_id);
  if (st != FR_OK) return st;
  if (rows <= 0) return fr_fail(e, FR_ERR_INVALID, "rows must be > 0");
  if (e->tables[table_id].d) return fr_fail(e, FR_ERR_STATE, "table %d already resident; set rows before loading", table_id);
  e->tables[table_id].rows = rows;
  return FR_OK;
}

extern "C" fr_status fr_load_table(fr_engine* e, int table_id, const float* host_rows, int64_t rows, int dim) {
  fr_status st = check_table(e, table_id);
  if (st != FR_OK) return st;
  FrTable& tb = e->tables[table_id];
  if (!host_rows) return fr_fail(e, FR_ERR_INVALID, "fr_load_table: null rows");
  if (dim != tb.dim) return fr_fail(e, FR_ERR_INVALID, "table %d: dim %d given, catalogue says %d", table_id, dim, tb.dim);
  if (!tb.resident) return FR_OK;
  if (tb.d && rows != tb.rows) return fr_fail(e, FR_ERR_STATE, "table %d: resident with %lld rows", table_id, (long long)tb.rows);
  if (rows <= 0) return fr_fail(e, FR_ERR_INVALID, "rows must be > 0");
  tb.rows = rows;
  if ((st = ensure_table_mem(e, table_id)) != FR_OK) return st;
  if (e->table_dtype == FR_TABLE_F32) {
    FR_CUDA(e, fr_h2d(e, tb.d, host_rows, (size_t)rows * dim * sizeof(float)));
  } else {
    // fp32 image -> 2- or 1-byte rows on the device, through a bounded fp32 staging buffer
    const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)(64 << 20) / ((int64_t)dim * 4));
    float* stage = nullptr;
    FR_CUDA(e, cudaMalloc(&stage, (size_t)std::min(chunk_rows, rows) * dim * sizeof(float)));
    for (int64_t r0 = 0; r0 < rows; r0 += chunk_rows) {
      const int64_t n = std::min(chunk_rows, rows - r0) * dim;
      cudaError_t ce = fr_h2d(e, stage, host_rows + r0 * dim, (size_t)n * sizeof(float));
      if (ce == cudaSuccess) {
        st = frk_quantize(e, stage, reinterpret_cast<char*>(tb.d) + (size_t)r0 * dim * fr_table_esize(e), n, e->default_stream->stream);
        ce = cudaStreamSynchronize(e->default_stream->stream);
      }
      if (ce != cudaSuccess || st != FR_OK) {
        cudaFree(stage);
        if (st != FR_OK) return st;
        return fr_fail(e, FR_ERR_CUDA, "fr_load_table: %s", cudaGetErrorString(ce));
      }
    }
    cudaFree(stage);
  }
  tb.loaded = true;
  tb.range_valid = false;
  e->f16_dirty = true;
  return FR_OK;
}

extern "C" fr_status fr_fill_table_reference(fr_engine* e, int table_id, int64_t debug_rows) {
  fr_status st = check_table(e, table_id);
  if (st != FR_OK) return st;
  FrTable& tb = e->tables[table_id];
  if (!tb.resident) return FR_OK;
  if ((st = ensure_table_mem(e, table_id)) != FR_OK) return st;
  if ((st = frk_fill_reference(e, tb.d, tb.rows, tb.dim, debug_rows, e->default_stream->stream)) != FR_OK) return st;
  FR_CUDA(e, cudaStreamSynchronize(e->default_stream->stream));
  tb.loaded = true;
  tb.range_valid = false;
  e->f16_dirty = true;
  return FR_OK;
}

extern "C" fr_status fr_fill_table_hash(fr_engine* e, int table_id, uint32_t seed) {
  fr_status st = check_table(e, table_id);
  if (st != FR_OK) return st;
  FrTable& tb = e->tables[table_id];
  if (!tb.resident) return FR_OK;
  if ((st = ensure_table_mem(e, table_id)) != FR_OK) return st;
  if ((st = frk_fill_hash(e, tb.d, seed, table_id, tb.rows, tb.dim, e->default_stream->stream)) != FR_OK) return st;
  FR_CUDA(e, cudaStreamSynchronize(e->default_stream->stream));
  tb.loaded = true;
  tb.range_valid = false;
  e->f16_dirty = true;
  return FR_OK;
}

extern "C" fr_status fr_read_table(fr_engine* e, int table_id, int64_t first_row, int64_t n_rows, float* host_out) {
  fr_status st = check_table(e, table_id);
  if (st != FR_OK) return st;
  FrTable& tb = e->tables[table_id];
  if (!tb.d || !tb.loaded) return fr_fail(e, FR_ERR_STATE, "table %d not loaded", table_id);
  if (first_row < 0 || n_rows < 0 || first_row + n_rows > tb.rows) return fr_fail(e, FR_ERR_INVALID, "row range");
  if (e->table_dtype == FR_TABLE_F32) {
    FR_CUDA(e, cudaMemcpy(host_out, tb.d + first_row * tb.dim, (size_t)n_rows * tb.dim * sizeof(float),
                          cudaMemcpyDeviceToHost));
    return FR_OK;
  }
  if (n_rows == 0) return FR_OK;
  float* tmp = nullptr;
  FR_CUDA(e, cudaMalloc(&tmp, (size_t)n_rows * tb.dim * sizeof(float)));
  st = frk_dequantize(e, reinterpret_cast<const char*>(tb.d) + (size_t)first_row * tb.dim * fr_table_esize(e), tmp, n_rows * tb.dim,
                      e->default_stream->stream);
  cudaError_t ce = st == FR_OK ? cudaStreamSynchronize(e->default_stream->stream) : cudaSuccess;
  if (st == FR_OK && ce == cudaSuccess)
    ce = cudaMemcpy(host_out, tmp, (size_t)n_rows * tb.dim * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(tmp);
  if (st != FR_OK) return st;
  if (ce != cudaSuccess) return fr_fail(e, FR_ERR_CUDA, "fr_read_table: %s", cudaGetErrorString(ce));
  return FR_OK;
}

extern "C" fr_status fr_load_mlp(fr_engine* e, int layer, const float* W, const float* bias) {
  if (!e) return fr_fail(nullptr, FR_ERR_INVALID, "null engine");
  if (layer < 0 || layer >= FR_MAX_LAYERS) return fr_fail(e, FR_ERR_INVALID, "layer %d out of range", layer);
  if (!W) return fr_fail(e, FR_ERR_INVALID, "fr_load_mlp: null weights");
  const int in = e->dims[layer], out = e->dims[layer + 1];
  const size_t nw = (size_t)in * out;
  FR_CUDA(e, cudaSetDevice(e->device));
  if (!e->d_W[layer]) FR_CUDA(e, cudaMalloc(&e->d_W[layer], nw * sizeof(float)));
  if (!e->d_Wt[layer]) FR_CUDA(e, cudaMalloc(&e->d_Wt[layer], nw * sizeof(float)));
  if (!e->d_bias[layer]) FR_CUDA(e, cudaMalloc(&e->d_bias[layer], (size_t)out * sizeof(float)));
  FR_CUDA(e, fr_h2d(e, e->d_W[layer], W, nw * sizeof(float)));
  if (bias) FR_CUDA(e, fr_h2d(e, e->d_bias[layer], bias, (size_t)out * sizeof(float)));
  else FR_CUDA(e, cudaMemsetAsync(e->d_bias[layer], 0, (size_t)out * sizeof(float), e->default_stream->stream));
  fr_status st = frk_transpose_round_tf32(e, e->d_W[layer], in, out, e->d_Wt[layer], e->default_stream->stream);
  if (st != FR_OK) return st;
  FR_CUDA(e, cudaStreamSynchronize(e->default_stream->stream));
  e->layer_loaded[layer] = true;
  e->f16_dirty = true;
  return FR_OK;
}

extern "C" fr_status fr_set_mlp_mode(fr_engine* e, int mode) {
  if (!e) return fr_fail(nullptr, FR_ERR_INVALID, "null engine");
  if (mode != FR_MLP_LINEAR && mode != FR_MLP_BIAS_RELU_SIGMOID) return fr_fail(e, FR_ERR_INVALID, "bad mlp_mode %d", mode);
  e->mlp_mode = mode;
  return FR_OK;
}

extern "C" fr_status fr_set_precision(fr_engine* e, int p) {
  if (!e) return fr_fail(nullptr, FR_ERR_INVALID, "null engine");
  if (p != FR_PREC_TF32 && p != FR_PREC_FP32) return fr_fail(e, FR_ERR_INVALID, "bad precision %d", p);
  e->precision = p;
  e->f16_dirty = true;
  return FR_OK;
}

// ---------------------------------------------------------------------------
extern "C" fr_status fr_stream_create(fr_engine* e, fr_stream* out) {
  if (!e || !out) return fr_fail(e, FR_ERR_INVALID, "fr_stream_create: null argument");
  FR_CUDA(e, cudaSetDevice(e->device));
  fr_stream_s* s = nullptr;
  fr_status st = alloc_stream(e, &s);
  if (st != FR_OK) return st;
  std::lock_guard<std::mutex> g(e->mu);
  e->streams.push_back(s);
  *out = s;
  return FR_OK;
}

extern "C" void fr_stream_destroy(fr_engine* e, fr_stream s) {
  if (!e || !s) return;
  {
    std::lock_guard<std::mutex> g(e->mu);
    for (size_t i = 0; i < e->streams.size(); i++)
      if (e->streams[i] == s) { e->streams.erase(e->streams.begin() + i); break; }
  }
  cudaSetDevice(e->device);
  free_stream(s);
}

extern "C" void* fr_stream_cuda(fr_stream s) { return s ? (void*)s->stream : nullptr; }

static bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// The device-side alias of a page-locked, mapped host buffer (cudaHostAlloc / cudaHostRegister under UVA),
// null for pageable or device memory.
static void* mapped_host_alias(const fr_engine* e, const void* p) {
  if (e->knobs.zero_copy_pct <= 0 || !p) return nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

static fr_status prep(fr_engine* e, fr_stream* s, int B, bool need_tables, bool need_mlp) {
  if (!e) return fr_fail(nullptr, FR_ERR_INVALID, "null engine");
  if (B < 0 || B > e->max_batch) return fr_fail(e, FR_ERR_INVALID, "B=%d outside [0, max_batch=%d]", B, e->max_batch);
  if (!*s) *s = e->default_stream;
  FR_CUDA(e, cudaSetDevice(e->device));
  if (need_tables) {
    for (size_t t = 0; t < e->tables.size(); t++)
      if (e->tables[t].resident && !e->tables[t].loaded)
        return fr_fail(e, FR_ERR_STATE, "table %d not loaded (fr_load_table / fr_fill_table_*)", (int)t);
    if (e->chunks_dirty) {
      std::lock_guard<std::mutex> g(e->mu);
      if (e->chunks_dirty) {
        fr_status st = frk_upload_chunks(e);
        if (st != FR_OK) return st;
      }
    }
  }
  if (need_mlp) {
    for (int k = 0; k < FR_MAX_LAYERS; k++)
      if (!e->layer_loaded[k]) return fr_fail(e, FR_ERR_STATE, "MLP layer %d not loaded (fr_load_mlp)", k);
    if (e->precision == FR_PREC_TF32) {
      fr_status st = frtc_prepare(e);
      if (st != FR_OK) return st;
    }
    if (need_tables && e->f16_dirty && e->f16_mode != FR_F16_OFF) {   // tables and weights are all loaded here
      std::lock_guard<std::mutex> g(e->mu);
      if (e->f16_dirty) {
        fr_status st = fr_f16_analyse(e);
        if (st != FR_OK) return st;
      }
    }
  }
  return FR_OK;
}

static fr_status stage_idx(fr_engine* e, fr_stream_s* s, const int32_t* idx, int B, const int32_t** d_idx) {
  if (!idx && B > 0) return fr_fail(e, FR_ERR_INVALID, "null idx");
  if (B == 0 || is_device_ptr(idx)) {
    *d_idx = idx;
    return FR_OK;
  }
  const size_t bytes = (size_t)B * e->ipr_full * sizeof(int32_t);   // (ipr_full: int32 words per index row in the engine's format)
  *d_idx = s->d_idx;
  // page-locked caller buffer: the copy engine moves the head of the batch, the SMs fetch the tail over PCIe
  // (zero_copy_pct of it, in 16-byte units); pageable buffer: the driver's staged copy
  const char* alias = static_cast<const char*>(mapped_host_alias(e, idx));
  size_t head = bytes;
  if (alias && (reinterpret_cast<uintptr_t>(alias) & 15) == 0 && bytes % 16 == 0)
    head = (bytes / 16) * (size_t)(100 - e->knobs.zero_copy_pct) / 100 * 16;
  if (head > 0) FR_CUDA(e, cudaMemcpyAsync(s->d_idx, idx, head, cudaMemcpyHostToDevice, s->stream));
  if (head < bytes)
    return frk_stage_idx(e, alias + head, reinterpret_cast<int32_t*>(reinterpret_cast<char*>(s->d_idx) + head), bytes - head,
                         s->stream);
  return FR_OK;
}

// Where the last MLP kernel writes the scores: the caller's buffer when the device can address it (device
// memory, or page-locked host memory written over PCIe by the epilogue itself), else the worker's buffer,
// which emit_scores() then copies out.
static float* score_target(const fr_engine* e, fr_stream_s* s, float* scores, int B) {
  if (B <= 0 || !scores) return s->d_scores;
  if (is_device_ptr(scores)) return scores;
  void* alias = mapped_host_alias(e, scores);
  return alias ? static_cast<float*>(alias) : s->d_scores;
}

// Launch `step` of the MLP chain on a worker stream (cuda_server.c:468-491).  TF32: 3 launches
// (layer 3 carries the output layer); FP32: 4 launches.  Returns the step's output buffer in *out.
static int mlp_steps(const fr_engine* e) { return e->precision == FR_PREC_TF32 ? 3 : 4; }

static fr_status run_mlp_step(fr_engine* e, fr_stream_s* s, int step, const float* in, int B, float* d_scores,
                              const float** out, const FrPeerWait* wait = nullptr) {
  const bool act = (e->mlp_mode == FR_MLP_BIAS_RELU_SIGMOID);
  if (e->precision == FR_PREC_TF32) {
    *out = step < 2 ? s->d_h[step] : d_scores;
    return frtc_layer(e, s, step, in, B, d_scores, wait);
  }
  if (step < 3) {
    *out = s->d_h[step];
    return frk_sgemm_bias_act(e, in, e->d_W[step], act ? e->d_bias[step] : nullptr, s->d_h[step], B, e->dims[step],
                              e->dims[step + 1], act, s->stream);
  }
  *out = d_scores;
  return frk_final_dot(e, in, e->d_W[3], act ? e->d_bias[3] : nullptr, d_scores, B, e->dims[3], act, s->stream);
}

// wait_slot >= 0 (FR_SHARD_FOLD, experiments build): the exchange published without waiting; d_x is the slot's concat
// buffer, complete once every rank has published the slot's current step -- the tcgen05 path polls the flags inside
// its first kernel, the others follow a wait kernel.
static fr_status run_mlp(fr_engine* e, fr_stream_s* s, const float* d_x, int B, float* d_scores, int wait_slot = -1) {
  if (B == 0) return wait_slot >= 0 ? frk_shard_wait(e, wait_slot, s->stream) : FR_OK;
  FrPeerWait pw = {nullptr, nullptr, 0, nullptr};
  const FrPeerWait* wait = nullptr;
  if (wait_slot >= 0) {
    if (e->precision == FR_PREC_TF32 && !frtc_can_chain(e, B)) {
      pw.flags = reinterpret_cast<const int*>(e->d_xchg + fr_xchg_flags_off(e, wait_slot));
      pw.step = e->d_step + wait_slot;
      pw.world = e->world;
      FR_CUDA(e, cudaHostGetDevicePointer(&pw.err, e->h_shard_err, 0));
      wait = &pw;
    } else {
      fr_status st = frk_shard_wait(e, wait_slot, s->stream);
      if (st != FR_OK) return st;
    }
  }
  if (frtc_can_chain(e, B)) return frtc_chain(e, s, d_x, B, d_scores);   // all layers in one persistent launch
  const float* in = d_x;
  for (int k = 0; k < mlp_steps(e); k++) {
    const float* out = nullptr;
    fr_status st = run_mlp_step(e, s, k, in, B, d_scores, &out, k == 0 ? wait : nullptr);
    if (st != FR_OK) return st;
    in = out;
  }
  return FR_OK;
}

static fr_status emit_scores(fr_engine* e, fr_stream_s* s, float* scores, int B, float* d_scores) {
  if (d_scores == s->d_scores && d_scores != scores && B > 0)
    FR_CUDA(e, cudaMemcpyAsync(scores, d_scores, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
  return FR_OK;
}

// Enqueue one batch on the worker's stream: [H2D idx] -> gather -> MLP launches -> [D2H scores].
static fr_status infer_enqueue(fr_engine* e, fr_stream_s* s, const int32_t* idx, int B, float* scores) {
  s->f16 = fr_tc_f16(e);
  const int32_t* d_idx = nullptr;
  fr_status st = stage_idx(e, s, idx, B, &d_idx);
  if (st != FR_OK) return st;
  float* d_scores = score_target(e, s, scores, B);
  if (B > 0 && frtc_can_fuse(e)) {
    // lookup fused into layer 1: the concat vectors never exist in global memory (3 launches)
    if ((st = frtc_fused_layer1(e, s, d_idx, B)) != FR_OK) return st;
    const float* in = s->d_h[0];
    for (int k = 1; k < 3; k++) {
      const float* out = nullptr;
      if ((st = run_mlp_step(e, s, k, in, B, d_scores, &out)) != FR_OK) return st;
      in = out;
    }
    return emit_scores(e, s, scores, B, d_scores);
  }
  if ((st = frk_gather(e, d_idx, B, s->d_x, e->precision == FR_PREC_TF32, s->stream, fr_tc_f16(e))) != FR_OK) return st;
  if ((st = run_mlp(e, s, s->d_x, B, d_scores)) != FR_OK) return st;
  return emit_scores(e, s, scores, B, d_scores);
}

// Do two addresses lie in ONE allocation (a single copy may span them)?  Two separately pinned buffers can sit back to
// back in the address space; a copy across the seam is an invalid argument.
static bool same_allocation(const void* a, const void* b) {
  typedef int (*PFN_attr)(void*, int, unsigned long long);   // CUresult cuPointerGetAttribute(void*, CUpointer_attribute, CUdeviceptr)
  static PFN_attr fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    cudaGetLastError();
    return reinterpret_cast<PFN_attr>(f);
  }();
  if (!fn) return false;
  unsigned long long sa = 0, sb = 0;
  constexpr int kRangeStart = 11;   // CU_POINTER_ATTRIBUTE_RANGE_START_ADDR
  if (fn(&sa, kRangeStart, (unsigned long long)(uintptr_t)a) != 0 || fn(&sb, kRangeStart, (unsigned long long)(uintptr_t)b) != 0) return false;
  return sa != 0 && sa == sb;
}

// device memory or page-locked host memory: the only buffers a captured memcpy node may reference
static bool is_capturable_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged || a.type == cudaMemoryTypeHost;
}

// The batch is 4-6 tiny launches; issued one by one the host (~4 us per launch) is the bottleneck, so a
// (buffers, B, entry point) combination is captured into a CUDA graph the first time a worker sees it and
// replayed as ONE cudaGraphLaunch from then on.  `enqueue(f)` issues the step's work on s->stream (called
// directly, or under capture) for flavour f: the table-sharded steps alternate between two exchange
// buffers, so their graphs come in pairs that are captured TOGETHER -- which graph a call replays is then
// independent of how many steps the worker has issued before (no per-parity warm-up).  The very first call
// of an (entry point, B) pair on an engine runs un-captured: it sets function attributes and fills the
// tensor-map cache.  The cache holds kMaxGraphs combinations per worker, least recently used evicted.
enum { FR_GV_INFER = 0, FR_GV_SHARD = 1, FR_GV_SHARD_SLICED = 2, FR_GV_MANY = 3 };   // | sub-batches << 4
constexpr size_t kMaxGraphs = 64;

static bool engine_warm(fr_engine* e, int variant, int B) {
  std::lock_guard<std::mutex> g(e->mu);
  for (const std::pair<int, int>& w : e->warmed)
    if (w.first == variant && w.second == B) return true;
  e->warmed.push_back({variant, B});
  return false;
}

static void drop_graph(fr_stream_s::Graph& g) {
  for (int f = 0; f < 2; f++)
    if (g.exec[f]) cudaGraphExecDestroy(g.exec[f]);
  g.exec[0] = g.exec[1] = nullptr;
}

template <class F>
static fr_status run_or_replay(fr_engine* e, fr_stream_s* s, const void* idx, const void* scores, int B, int variant,
                               int flavours, int flavour, F&& enqueue, const void* idx2 = nullptr, bool bypass = false) {
  if (!e->use_graphs || bypass) {
    e->graph_direct++;
    return enqueue(flavour);
  }
  fr_stream_s::Graph* g = nullptr;
  const int prec_key = e->precision | (fr_tc_f16(e) ? 16 : 0);
  for (fr_stream_s::Graph& c : s->graphs)
    if (c.idx == idx && c.idx2 == idx2 && c.scores == scores && c.B == B && c.mode == e->mlp_mode &&
        c.prec == prec_key && c.variant == variant) {
      g = &c;
      break;
    }
  if (g && g->exec[flavour]) {
    FR_CUDA(e, cudaGraphLaunch(g->exec[flavour], s->stream));
    g->last_use = ++s->graph_clock;
    e->launches += g->launches;
    e->graph_hits++;
    return FR_OK;
  }
  if ((g && g->failed) || !is_capturable_ptr(idx) || !is_capturable_ptr(scores) || (idx2 && !is_capturable_ptr(idx2)) ||
      !engine_warm(e, variant, B)) {
    e->graph_direct++;
    return enqueue(flavour);
  }
  if (!g) {
    if (s->graphs.size() >= kMaxGraphs) {   // evict the least recently used combination
      size_t lru = 0;
      for (size_t i = 1; i < s->graphs.size(); i++)
        if (s->graphs[i].last_use < s->graphs[lru].last_use) lru = i;
      drop_graph(s->graphs[lru]);
      s->graphs.erase(s->graphs.begin() + lru);
    }
    s->graphs.push_back({idx, idx2, scores, B, e->mlp_mode, prec_key, variant, flavours, false, 0, 0, {nullptr, nullptr}});
    g = &s->graphs.back();
  }
  g->last_use = ++s->graph_clock;
  for (int f = 0; f < flavours && !g->failed; f++) {
    const int64_t l0 = e->launches.load();
    cudaError_t ce = cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal);
    if (ce != cudaSuccess) {
      cudaGetLastError();
      g->failed = true;
      break;
    }
    const fr_status st = enqueue(f);
    cudaGraph_t graph = nullptr;
    ce = cudaStreamEndCapture(s->stream, &graph);
    g->launches = (int)(e->launches.load() - l0);
    e->launches -= g->launches;  // nothing ran yet
    if (st == FR_OK && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&g->exec[f], graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (st != FR_OK || ce != cudaSuccess || !g->exec[f]) {
      cudaGetLastError();
      g->failed = true;
    }
  }
  if (g->failed) {
    drop_graph(*g);
    e->graph_direct++;
    return enqueue(flavour);
  }
  FR_CUDA(e, cudaGraphLaunch(g->exec[flavour], s->stream));
  e->launches += g->launches;
  e->graph_captures++;
  return FR_OK;
}

// Drop every cached graph of a worker (its buffers are about to be freed or re-used for something else).
extern "C" fr_status fr_graph_flush(fr_engine* e, fr_stream s) {
  if (!e) return fr_fail(nullptr, FR_ERR_INVALID, "null engine");
  if (!s) s = e->default_stream;
  FR_CUDA(e, cudaSetDevice(e->device));
  FR_CUDA(e, cudaStreamSynchronize(s->stream));
  for (fr_stream_s::Graph& g : s->graphs) drop_graph(g);
  s->graphs.clear();
  return FR_OK;
}

extern "C" fr_status fr_graph_stats(const fr_engine* e, int64_t* replayed, int64_t* captured, int64_t* direct) {
  if (!e) return fr_fail(nullptr, FR_ERR_INVALID, "null engine");
  if (replayed) *replayed = e->graph_hits.load();
  if (captured) *captured = e->graph_captures.load();
  if (direct) *direct = e->graph_direct.load();
  return FR_OK;
}

// The *_many calls stage the indices of several batches with one copy: grow the worker's index / score staging buffers
// to what the call needs (the activation buffers stay one batch deep).  Growing synchronises the worker and drops its
// graphs, which reference the old buffers; it happens once per worker.
static fr_status ensure_group_capacity(fr_engine* e, fr_stream_s* s, size_t idx_ints, size_t n_scores) {
  if (idx_ints <= s->idx_cap && n_scores <= s->scores_cap) return FR_OK;
  FR_CUDA(e, cudaStreamSynchronize(s->stream));
  for (fr_stream_s::Graph& g : s->graphs) drop_graph(g);
  s->graphs.clear();
  // (the outgrown buffers are kept until the worker is destroyed: cudaFree synchronises the whole device, and another
  // engine of this process may have a sharded step spinning there that waits for THIS engine's next launch)
  if (idx_ints > s->idx_cap) {
    s->retired.push_back(s->d_idx);
    s->d_idx = nullptr;
    s->idx_cap = 0;
    FR_CUDA(e, cudaMalloc(&s->d_idx, idx_ints * sizeof(int32_t)));
    s->idx_cap = idx_ints;
  }
  if (n_scores > s->scores_cap) {
    s->retired.push_back(s->d_scores);
    s->d_scores = nullptr;
    s->scores_cap = 0;
    FR_CUDA(e, cudaMalloc(&s->d_scores, n_scores * sizeof(float)));
    s->scores_cap = n_scores;
  }
  return FR_OK;
}

fr_status fr_infer_opts(fr_engine* e, const int32_t* idx, int B, float* scores, fr_stream s, bool no_graph) {
  fr_status st = prep(e, &s, B, true, true);
  if (st != FR_OK) return st;
  if (B == 0) return FR_OK;
  if (!idx || !scores) return fr_fail(e, FR_ERR_INVALID, "null idx/scores");
  if (e->world > 1) return fr_fail(e, FR_ERR_STATE, "engine is table-sharded over %d ranks: use fr_shard_infer", e->world);
  return run_or_replay(e, s, idx, scores, B, FR_GV_INFER, 1, 0, [&](int) { return infer_enqueue(e, s, idx, B, scores); },
                       nullptr, no_graph);
}

extern "C" fr_status fr_infer(fr_engine* e, const int32_t* idx, int B, float* scores, fr_stream s) {
  return fr_infer_opts(e, idx, B, scores, s, false);
}

// n batches of B items in ONE call on one worker: idx [n][B][T] and scores [n][B] contiguous.  Host buffers
// travel in one copy each way (a copy costs the engine ~4 us on top of its bytes, whatever its size), the
// batches then run back to back on the worker's stream, each through the same kernels as fr_infer.
extern "C" fr_status fr_infer_many(fr_engine* e, const int32_t* idx, int n, int B, float* scores, fr_stream s) {
  if (n < 0 || n > 4095 || B < 0 || (int64_t)n * B > INT_MAX) return fr_fail(e, FR_ERR_INVALID, "fr_infer_many: n=%d B=%d", n, B);
  fr_status st = prep(e, &s, B, true, true);
  if (st != FR_OK) return st;
  if (n == 0 || B == 0) return FR_OK;
  if (!idx || !scores) return fr_fail(e, FR_ERR_INVALID, "null idx/scores");
  if (e->world > 1) return fr_fail(e, FR_ERR_STATE, "engine is table-sharded over %d ranks: use fr_shard_infer", e->world);
  if ((st = ensure_group_capacity(e, s, (size_t)n * B * e->ipr_full, (size_t)n * B)) != FR_OK) return st;
  return run_or_replay(e, s, idx, scores, B, FR_GV_MANY | (n << 4), 1, 0, [&](int) {
    s->f16 = fr_tc_f16(e);
    const int32_t* d_idx = nullptr;
    fr_status r = stage_idx(e, s, idx, n * B, &d_idx);
    if (r != FR_OK) return r;
    float* d_scores = score_target(e, s, scores, n * B);
    for (int i = 0; i < n; i++) {
      const int32_t* bi = d_idx + (size_t)i * B * e->ipr_full;
      if ((r = frk_gather(e, bi, B, s->d_x, e->precision == FR_PREC_TF32, s->stream, fr_tc_f16(e))) != FR_OK) return r;
      if ((r = run_mlp(e, s, s->d_x, B, d_scores + (size_t)i * B)) != FR_OK) return r;
    }
    return emit_scores(e, s, scores, n * B, d_scores);
  });
}

extern "C" fr_status fr_gather_only(fr_engine* e, const int32_t* idx, int B, float* concat, fr_stream s) {
  fr_status st = prep(e, &s, B, true, false);
  if (st != FR_OK) return st;
  if (B > 0 && !concat) return fr_fail(e, FR_ERR_INVALID, "null concat");
  if (e->world > 1) return fr_fail(e, FR_ERR_STATE, "engine is table-sharded over %d ranks: use fr_shard_gather_push", e->world);
  const int32_t* d_idx = nullptr;
  if ((st = stage_idx(e, s, idx, B, &d_idx)) != FR_OK) return st;
  if (B == 0) return FR_OK;
  const bool dev = is_device_ptr(concat);
  float* d_out = dev ? concat : s->d_x;
  if ((st = frk_gather(e, d_idx, B, d_out, false, s->stream)) != FR_OK) return st;
  if (!dev)
    FR_CUDA(e, cudaMemcpyAsync(concat, d_out, (size_t)B * e->D * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
  return FR_OK;
}

extern "C" fr_status fr_mlp_only(fr_engine* e, const float* x, int B, float* scores, fr_stream s) {
  fr_status st = prep(e, &s, B, false, true);
  if (st != FR_OK) return st;
  if (B == 0) return FR_OK;
  if (!x || !scores) return fr_fail(e, FR_ERR_INVALID, "null x/scores");
  // (caller data: the fp16 range analysis cannot bound it, so this entry point always runs TF32 / FP32)
  s->f16 = false;
  const float* d_x = x;
  if (!is_device_ptr(x)) {
    FR_CUDA(e, cudaMemcpyAsync(s->d_x, x, (size_t)B * e->D * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    d_x = s->d_x;
  }
  // kind::tf32 TRUNCATES fp32 operands; the header promises round-to-nearest (what the lookup does on the
  // fr_infer path), so the same input gives the same scores through either entry point
  if (e->precision == FR_PREC_TF32) {
    if ((st = frk_round_tf32(e, d_x, s->d_x, (int64_t)B * e->D, s->stream)) != FR_OK) return st;
    d_x = s->d_x;
  }
  float* d_scores = score_target(e, s, scores, B);
  if ((st = run_mlp(e, s, d_x, B, d_scores)) != FR_OK) return st;
  return emit_scores(e, s, scores, B, d_scores);
}

extern "C" fr_status fr_layer_only(fr_engine* e, int k, const float* x, int B, float* y, fr_stream s) {
  fr_status st = prep(e, &s, B, false, true);
  if (st != FR_OK) return st;
  if (k < 0 || k >= mlp_steps(e)) return fr_fail(e, FR_ERR_INVALID, "fr_layer_only: step %d of %d", k, mlp_steps(e));
  s->f16 = false;
  if (B == 0) return FR_OK;
  if (!x || !y) return fr_fail(e, FR_ERR_INVALID, "null x/y");
  const bool last = (k == mlp_steps(e) - 1);
  const int in_w = (e->precision == FR_PREC_FP32 && k == 3) ? e->dims[3] : e->dims[k];
  // stage the input in the buffer the previous step would have written
  float* d_in = (k == 0) ? s->d_x : s->d_h[k - 1];
  FR_CUDA(e, cudaMemcpyAsync(d_in, x, (size_t)B * in_w * sizeof(float), cudaMemcpyDefault, s->stream));
  const float* out = nullptr;
  if ((st = run_mlp_step(e, s, k, d_in, B, s->d_scores, &out)) != FR_OK) return st;
  const size_t n = last ? (size_t)B : (size_t)B * e->dims[k + 1];
  FR_CUDA(e, cudaMemcpyAsync(y, out, n * sizeof(float), cudaMemcpyDefault, s->stream));
  return FR_OK;
}

// Debug hook (bench.py binds it by name): enqueue MLP step k alone on a worker, on whatever its activation
// buffers hold, with no copies and no synchronisation -- so a bench can run one kernel on all workers at once and
// see what it sustains at the occupancy of the real step.
extern "C" fr_status frdbg_enqueue_layer(fr_engine* e, int k, int B, fr_stream s) {
  fr_status st = prep(e, &s, B, false, true);
  if (st != FR_OK) return st;
  if (k < 0 || k >= mlp_steps(e) || B <= 0) return fr_fail(e, FR_ERR_INVALID, "frdbg_enqueue_layer: bad argument");
  const float* out = nullptr;
  return run_mlp_step(e, s, k, k == 0 ? s->d_x : s->d_h[k - 1], B, s->d_scores, &out);
}

extern "C" fr_status fr_sync(fr_engine* e, fr_stream s) {
  if (!e) return fr_fail(nullptr, FR_ERR_INVALID, "null engine");
  if (!s) s = e->default_stream;
  cudaError_t ce = cudaStreamSynchronize(s->stream);
  if (ce != cudaSuccess) {
    const int* w = e->h_watch;
    if (w && w[0])
      return fr_fail(e, FR_ERR_CUDA, "cudaStreamSynchronize: %s; tcgen05 kernel watchdog: wait %d (1 smem slot free, 2 TMEM stage "
                     "free, 3 smem slot full, 4 TMEM stage full, 8 peer ranks' rows) gave up in CTA %d of %d, warp %d, parity %d, "
                     "counters %d/%d", cudaGetErrorString(ce), w[1], w[2], w[5], w[3], w[4], w[6], w[7]);
    return fr_fail(e, FR_ERR_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(ce));
  }
  // a sharded step that gave up waiting for a peer rank ran its MLP on an incomplete concat buffer
  if (e->h_shard_err && *reinterpret_cast<volatile int*>(e->h_shard_err))
    return fr_fail(e, FR_ERR_STATE, "a sharded step timed out waiting for a peer rank's rows (~2 s); its scores are invalid");
  if (e->h_idx_err) {
    volatile int* ie = e->h_idx_err;
    if (ie[0]) {
      const int n = ie[0], col = ie[1], val = ie[2], item = ie[3];
      ie[0] = 0;
      return fr_fail(e, FR_ERR_INVALID, "%d out-of-range indices since the last fr_sync (e.g. item %d, index column %d: %d); "
                     "row 0 was read instead", n, item, col, val);
    }
  }
  return FR_OK;
}

static void drop_all_graphs(fr_engine* e) {
  std::lock_guard<std::mutex> g(e->mu);
  std::vector<fr_stream_s*> all = e->streams;
  all.push_back(e->default_stream);
  for (fr_stream_s* s : all) {
    for (fr_stream_s::Graph& gr : s->graphs) drop_graph(gr);
    s->graphs.clear();
  }
}

// Engine options (include/fleetrec.h FR_OPT_*).  Changing one synchronises the device and drops the cached CUDA
// graphs, which were captured under the old setting.
extern "C" fr_status fr_set_option(fr_engine* e, int option, int value) {
  if (!e) return fr_fail(nullptr, FR_ERR_INVALID, "null engine");
  FR_CUDA(e, cudaSetDevice(e->device));
  switch (option) {
    case FR_OPT_CUDA_GRAPHS: case FR_OPT_CHECK_INDICES: case FR_OPT_FUSE_LOOKUP:
      if (value != 0 && value != 1) return fr_fail(e, FR_ERR_INVALID, "fr_set_option(%d): value must be 0 or 1", option);
      break;
    case FR_OPT_TILE_HINT:
      if (value < FR_HINT_AUTO || value > FR_HINT_THROUGHPUT) return fr_fail(e, FR_ERR_INVALID, "FR_OPT_TILE_HINT: FR_HINT_*");
      break;
    case FR_OPT_INDEX_FORMAT:
      if (value != FR_IDX_I32 && value != FR_IDX_PACKED) return fr_fail(e, FR_ERR_INVALID, "FR_OPT_INDEX_FORMAT: FR_IDX_*");
      break;
    case FR_OPT_F16_OPERANDS:
      if (value < FR_F16_OFF || value > FR_F16_GUARDED) return fr_fail(e, FR_ERR_INVALID, "FR_OPT_F16_OPERANDS: FR_F16_*");
      if (value != FR_F16_OFF && e->world > 1)
        return fr_fail(e, FR_ERR_UNSUPPORTED, "fp16 operands are not available on a table-sharded engine");
      break;
    default:
      return fr_fail(e, FR_ERR_INVALID, "fr_set_option: unknown option %d", option);
  }
  if (option == FR_OPT_CHECK_INDICES && value && !e->h_idx_err) {
    FR_CUDA(e, cudaHostAlloc(&e->h_idx_err, 4 * sizeof(int), cudaHostAllocMapped));
    memset(e->h_idx_err, 0, 4 * sizeof(int));
  }
  FR_CUDA(e, cudaDeviceSynchronize());
  drop_all_graphs(e);
  switch (option) {
    case FR_OPT_CUDA_GRAPHS: e->use_graphs = value != 0; break;
    case FR_OPT_CHECK_INDICES: e->check_indices = value != 0; break;
    case FR_OPT_FUSE_LOOKUP: e->fuse_lookup = value != 0; break;
    case FR_OPT_TILE_HINT: e->tile_hint = value; break;
    case FR_OPT_INDEX_FORMAT:   // the piece descriptors carry the index offsets: rebuilt before the next lookup
      e->index_format = value;
      e->chunks_dirty = true;
      e->shard_lists_built = false;
      fr_index_rows(e);
      break;
    case FR_OPT_F16_OPERANDS:
      e->f16_mode = value;
      e->f16_dirty = true;
      if (value == FR_F16_OFF) e->tc_f16 = false;
      break;
  }
  return FR_OK;
}

extern "C" fr_status fr_index_layout(fr_engine* e, int which, int32_t* byte_offset, int32_t* width, int* n, int* row_bytes) {
  if (!e || which < 0 || which > 2) return fr_fail(e, FR_ERR_INVALID, "fr_index_layout: bad argument");
  if (which < 2 && e->owner.empty() && e->world > 1) return fr_fail(e, FR_ERR_STATE, "fr_shard_init first");
  fr_index_rows(e);
  const std::vector<int>& off = which == 2 ? e->idx_off_full : (which == 0 ? e->idx_off_owned : e->idx_off_repl);
  const int words = which == 2 ? e->ipr_full : (which == 0 ? e->ipr_owned : e->ipr_repl);
  for (size_t i = 0; i < off.size(); i++) {
    if (byte_offset) byte_offset[i] = off[i] & 0x7FFFFFFF;
    if (width) width[i] = off[i] < 0 ? 2 : 4;
  }
  if (n) *n = (int)off.size();
  if (row_bytes) *row_bytes = 4 * words;
  return FR_OK;
}

extern "C" fr_status fr_f16_report(fr_engine* e, int* active, float* bounds5) {
  fr_stream s = nullptr;
  fr_status st = prep(e, &s, 0, true, true);   // runs the analysis if it is due (tables and weights must be loaded)
  if (st != FR_OK) return st;
  if (active) *active = fr_tc_f16(e) ? 1 : 0;
  if (bounds5)
    for (int i = 0; i < 5; i++) bounds5[i] = e->f16_bounds[i];
  return FR_OK;
}

extern "C" int64_t fr_launch_count(const fr_engine* e) { return e ? e->launches.load() : 0; }

extern "C" int64_t fr_table_bytes(const fr_engine* e) {
  if (!e) return 0;
  int64_t n = 0;
  for (const FrTable& t : e->tables)
    if (t.d) n += t.rows * t.dim * (int64_t)fr_table_esize(e);
  return n;
}

extern "C" fr_status fr_mark(fr_engine* e, fr_stream s, int which) {
  if (!e || which < 0 || which > 1) return fr_fail(e, FR_ERR_INVALID, "fr_mark: bad argument");
  if (!s) s = e->default_stream;
  FR_CUDA(e, cudaEventRecord(s->ev[which], s->stream));
  return FR_OK;
}

extern "C" fr_status fr_elapsed_ms(fr_engine* e, fr_stream s, float* ms) {
  if (!e || !ms) return fr_fail(e, FR_ERR_INVALID, "fr_elapsed_ms: bad argument");
  if (!s) s = e->default_stream;
  FR_CUDA(e, cudaEventSynchronize(s->ev[1]));
  FR_CUDA(e, cudaEventElapsedTime(ms, s->ev[0], s->ev[1]));
  return FR_OK;
}

extern "C" fr_status fr_time_kernels(fr_engine* e, const int32_t* idx, int B, int reps, fr_stream s, float* ms5) {
  fr_status st = prep(e, &s, B, true, true);
  if (st != FR_OK) return st;
  if (!ms5 || reps <= 0 || B <= 0) return fr_fail(e, FR_ERR_INVALID, "fr_time_kernels: bad argument");
  const int32_t* d_idx = nullptr;
  if ((st = stage_idx(e, s, idx, B, &d_idx)) != FR_OK) return st;
  for (int i = 0; i < 5; i++) ms5[i] = 0.f;
  cudaEvent_t e0 = s->ev[0], e1 = s->ev[1];
  const bool round = e->precision == FR_PREC_TF32;
  s->f16 = fr_tc_f16(e);
  // every kernel is timed alone, back to back `reps` times, events on its own stream
  for (int pass = 0; pass < 2; pass++) {  // pass 0 = warm-up
    const float* in = s->d_x;
    const bool fused = frtc_can_fuse(e);
    if (fused) {
      // no stand-alone lookup in the step: slot 0 stays 0, slot 1 is lookup + layer 1 in one kernel
      FR_CUDA(e, cudaEventRecord(e0, s->stream));
      for (int r = 0; r < reps; r++)
        if ((st = frtc_fused_layer1(e, s, d_idx, B)) != FR_OK) return st;
      FR_CUDA(e, cudaEventRecord(e1, s->stream));
      FR_CUDA(e, cudaEventSynchronize(e1));
      FR_CUDA(e, cudaEventElapsedTime(&ms5[1], e0, e1));
      in = s->d_h[0];
    } else if (e->world == 1) {
      FR_CUDA(e, cudaEventRecord(e0, s->stream));
      for (int r = 0; r < reps; r++)
        if ((st = frk_gather(e, d_idx, B, s->d_x, round, s->stream, fr_tc_f16(e))) != FR_OK) return st;
      FR_CUDA(e, cudaEventRecord(e1, s->stream));
      FR_CUDA(e, cudaEventSynchronize(e1));
      FR_CUDA(e, cudaEventElapsedTime(&ms5[0], e0, e1));
    } else {
      // sharded: the lookup needs every rank; time the MLP on this worker's last exchanged batch
      in = e->d_xchg + fr_xchg_concat_off(e, s->slot < e->n_slots ? s->slot : 0, s->shard_step & 1);
    }
    if (!fused && frtc_can_chain(e, B)) {
      // the MLP of this batch size is ONE launch: slot 1 carries it, slots 2..4 stay 0
      FR_CUDA(e, cudaEventRecord(e0, s->stream));
      for (int r = 0; r < reps; r++)
        if ((st = frtc_chain(e, s, in, B, s->d_scores)) != FR_OK) return st;
      FR_CUDA(e, cudaEventRecord(e1, s->stream));
      FR_CUDA(e, cudaEventSynchronize(e1));
      FR_CUDA(e, cudaEventElapsedTime(&ms5[1], e0, e1));
      continue;
    }
    for (int k = fused ? 1 : 0; k < mlp_steps(e); k++) {
      const float* out = nullptr;
      FR_CUDA(e, cudaEventRecord(e0, s->stream));
      for (int r = 0; r < reps; r++)
        if ((st = run_mlp_step(e, s, k, in, B, s->d_scores, &out)) != FR_OK) return st;
      FR_CUDA(e, cudaEventRecord(e1, s->stream));
      FR_CUDA(e, cudaEventSynchronize(e1));
      FR_CUDA(e, cudaEventElapsedTime(&ms5[1 + k], e0, e1));
      in = out;
    }
  }
  for (int i = 0; i < 5; i++) ms5[i] /= (float)reps;
  return FR_OK;
}

// ---------------------------------------------------------------------------
// Sharding: table-wise model parallel, push all-to-all over NVLink.
//
// Exchange region of one rank (ONE cudaMalloc, exported through CUDA IPC), n_slots times:
//   [ concat buffer 0 | concat buffer 1 | flags: int32[world] (padded to 256 B) ]
// A slot belongs to one worker stream (same creation order on every rank), so several sharded
// steps are in flight at once, one per worker.  Within a slot, concat buffer p holds this rank's
// B_global/world items of the step with parity p; flags[r] is the last step of the slot rank r
// has finished pushing for (written by rank r over NVLink).
extern "C" fr_status fr_shard_init(fr_engine* e, int rank, int world, const int* owner) {
  if (!e) return fr_fail(nullptr, FR_ERR_INVALID, "null engine");
  if (world < 1 || rank < 0 || rank >= world || !owner) return fr_fail(e, FR_ERR_INVALID, "fr_shard_init: rank/world/owner");
  for (const FrTable& t : e->tables)
    if (t.d) return fr_fail(e, FR_ERR_STATE, "fr_shard_init must precede table loading");
  if (e->max_batch % world) return fr_fail(e, FR_ERR_INVALID, "max_batch %d not divisible by world %d", e->max_batch, world);
  if (e->d_xchg) return fr_fail(e, FR_ERR_STATE, "fr_shard_init called twice");
  e->rank = rank;
  e->world = world;
  e->owner.assign(owner, owner + e->tables.size());
  for (size_t t = 0; t < e->tables.size(); t++) {
    if (owner[t] < -1 || owner[t] >= world) return fr_fail(e, FR_ERR_INVALID, "owner[%d]=%d", (int)t, owner[t]);
    e->tables[t].resident = (owner[t] == -1 || owner[t] == rank);
  }
  if (world > 32) return fr_fail(e, FR_ERR_UNSUPPORTED, "world %d > 32 (one warp publishes and polls the flags)", world);
  FR_CUDA(e, cudaSetDevice(e->device));
  e->n_slots = 33;   // the default worker + 32 created workers
  const size_t bytes = (size_t)e->n_slots * fr_xchg_slot_floats(e) * sizeof(float);
  FR_CUDA(e, cudaMalloc(&e->d_xchg, bytes));
  FR_CUDA(e, cudaMemsetAsync(e->d_xchg, 0, bytes, e->default_stream->stream));
  FR_CUDA(e, cudaMalloc(&e->d_step, sizeof(int) * e->n_slots));
  FR_CUDA(e, cudaMemsetAsync(e->d_step, 0, sizeof(int) * e->n_slots, e->default_stream->stream));
  FR_CUDA(e, cudaHostAlloc(&e->h_shard_err, sizeof(int), cudaHostAllocMapped));
  *e->h_shard_err = 0;
  FR_CUDA(e, cudaStreamSynchronize(e->default_stream->stream));
  e->peers.assign(world, FrPeer());
  e->peers[rank].concat = e->d_xchg;
  e->chunks_dirty = true;
  e->owned_tables.clear();
  e->repl_tables.clear();
  fr_index_rows(e);
  return FR_OK;
}

static fr_status upload_peers(fr_engine* e) {
  std::vector<float*> p(e->world);
  for (int r = 0; r < e->world; r++) {
    if (!e->peers[r].concat) return fr_fail(e, FR_ERR_STATE, "peer %d not attached", r);
    p[r] = e->peers[r].concat;
  }
  if (!e->d_peer_ptrs) FR_CUDA(e, cudaMalloc(&e->d_peer_ptrs, sizeof(float*) * e->world));
  FR_CUDA(e, fr_h2d(e, e->d_peer_ptrs, p.data(), sizeof(float*) * e->world));
  return FR_OK;
}

extern "C" fr_status fr_shard_export(fr_engine* e, void* handle64) {
  if (!e || !handle64) return fr_fail(e, FR_ERR_INVALID, "fr_shard_export: null argument");
  if (!e->d_xchg) return fr_fail(e, FR_ERR_STATE, "fr_shard_init first");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  FR_CUDA(e, cudaSetDevice(e->device));
  cudaIpcMemHandle_t h;
  FR_CUDA(e, cudaIpcGetMemHandle(&h, e->d_xchg));
  memcpy(handle64, &h, 64);
  return FR_OK;
}

extern "C" fr_status fr_shard_import(fr_engine* e, const void* handles) {
  if (!e || !handles) return fr_fail(e, FR_ERR_INVALID, "fr_shard_import: null argument");
  if (!e->d_xchg) return fr_fail(e, FR_ERR_STATE, "fr_shard_init first");
  FR_CUDA(e, cudaSetDevice(e->device));
  for (int r = 0; r < e->world; r++) {
    if (r == e->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * r, 64);
    void* p = nullptr;
    FR_CUDA(e, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    e->peers[r].concat = (float*)p;
    e->peers[r].ipc = true;
  }
  return upload_peers(e);
}

extern "C" fr_status fr_shard_attach_local(fr_engine* e, fr_engine* const* peers) {
  if (!e || !peers) return fr_fail(e, FR_ERR_INVALID, "fr_shard_attach_local: null argument");
  if (!e->d_xchg) return fr_fail(e, FR_ERR_STATE, "fr_shard_init first");
  FR_CUDA(e, cudaSetDevice(e->device));
  for (int r = 0; r < e->world; r++) {
    if (r == e->rank) continue;
    if (!peers[r] || !peers[r]->d_xchg) return fr_fail(e, FR_ERR_STATE, "peer %d has no exchange buffer", r);
    if (peers[r]->device != e->device) {
      int can = 0;
      FR_CUDA(e, cudaDeviceCanAccessPeer(&can, e->device, peers[r]->device));
      if (!can) return fr_fail(e, FR_ERR_CUDA, "device %d cannot access peer %d", e->device, peers[r]->device);
      cudaError_t ce = cudaDeviceEnablePeerAccess(peers[r]->device, 0);
      if (ce != cudaSuccess && ce != cudaErrorPeerAccessAlreadyEnabled)
        return fr_fail(e, FR_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(ce));
      cudaGetLastError();
    }
    e->peers[r].concat = peers[r]->d_xchg;
  }
  return upload_peers(e);
}

static fr_status shard_check(fr_engine* e, const fr_stream_s* s, int B_global) {
  if (!e->d_xchg) return fr_fail(e, FR_ERR_STATE, "fr_shard_init first");
  if (s->slot >= e->n_slots)
    return fr_fail(e, FR_ERR_UNSUPPORTED, "worker %d has no exchange slot (%d slots per engine)", s->slot, e->n_slots);
  if (!e->d_peer_ptrs) return fr_fail(e, FR_ERR_STATE, "exchange buffers not attached (fr_shard_import)");
  if (B_global % e->world) return fr_fail(e, FR_ERR_INVALID, "B_global %d not divisible by world %d", B_global, e->world);
  return FR_OK;
}

// Two-phase form (the host barriers all ranks between the calls); parity 0 buffer only.
extern "C" fr_status fr_shard_gather_push(fr_engine* e, const int32_t* idx, int B_global, fr_stream s) {
  fr_status st = prep(e, &s, B_global, true, false);
  if (st != FR_OK) return st;
  if ((st = shard_check(e, s, B_global)) != FR_OK) return st;
  const int32_t* d_idx = nullptr;
  if ((st = stage_idx(e, s, idx, B_global, &d_idx)) != FR_OK) return st;
  if (B_global == 0) return FR_OK;
  const int T = e->ipr_full;
  return frk_shard_exchange(e, e->d_chunks, d_idx, T, d_idx + (size_t)e->rank * (B_global / e->world) * T, T, B_global,
                            s->slot, 0, false, s->stream);   // (the host barriers the ranks before fr_shard_mlp)
}

extern "C" fr_status fr_shard_mlp(fr_engine* e, int B_global, float* scores_local, fr_stream s) {
  fr_status st = prep(e, &s, B_global, false, true);
  if (st != FR_OK) return st;
  if ((st = shard_check(e, s, B_global)) != FR_OK) return st;
  const int Bl = B_global / e->world;
  if (Bl == 0) return FR_OK;
  if (!scores_local) return fr_fail(e, FR_ERR_INVALID, "null scores");
  s->f16 = false;
  float* d_scores = is_device_ptr(scores_local) ? scores_local : s->d_scores;
  if ((st = run_mlp(e, s, e->d_xchg + fr_xchg_concat_off(e, s->slot, 0), Bl, d_scores)) != FR_OK) return st;
  return emit_scores(e, s, scores_local, Bl, d_scores);
}

extern "C" fr_status fr_shard_read_concat(fr_engine* e, int B_global, float* concat_local, fr_stream s) {
  if (!e || !concat_local) return fr_fail(e, FR_ERR_INVALID, "fr_shard_read_concat: null argument");
  if (!e->d_xchg) return fr_fail(e, FR_ERR_STATE, "fr_shard_init first");
  if (!s) s = e->default_stream;
  FR_CUDA(e, cudaSetDevice(e->device));
  if (s->slot >= e->n_slots) return fr_fail(e, FR_ERR_UNSUPPORTED, "worker has no exchange slot");
  if (*reinterpret_cast<volatile int*>(e->h_shard_err))
    return fr_fail(e, FR_ERR_STATE, "a sharded step timed out waiting for a peer rank's rows; the concat buffer is incomplete");
  const int Bl = B_global / e->world;
  const float* src = e->d_xchg + fr_xchg_concat_off(e, s->slot, s->shard_step & 1);  // last step's parity
  FR_CUDA(e, cudaMemcpyAsync(concat_local, src, (size_t)Bl * e->D * sizeof(float), cudaMemcpyDefault, s->stream));
  return FR_OK;
}

// One-call sharded step with device-side synchronisation (no host barrier, no NCCL on the data
// path): push my tables' pieces for the global batch into the owners' buffers of this step's
// parity -> publish "rank r finished step n of this slot" into every peer's flag block -> wait
// until all ranks have published step n -> MLP over my B_global/world items.  Every rank must issue
// the same sequence of calls per worker with the same global batch.  The step number lives on the
// device, so the step is replayed as a CUDA graph (one per buffer parity).
static fr_status shard_infer_enqueue(fr_engine* e, fr_stream_s* s, const int32_t* idx, int B_global, float* scores,
                                     int parity) {
  s->f16 = false;
  const int32_t* d_idx = nullptr;
  fr_status st = stage_idx(e, s, idx, B_global, &d_idx);
  if (st != FR_OK) return st;
  const int T = e->ipr_full;
  const int Bl = B_global / e->world;
  const bool fold = e->knobs.shard_fold_wait != 0;
  if ((st = frk_shard_exchange(e, e->d_chunks, d_idx, T, d_idx + (size_t)e->rank * Bl * T, T, B_global, s->slot, parity,
                               !fold, s->stream)) != FR_OK) return st;
  float* d_scores = score_target(e, s, scores, Bl);
  const float* x = e->d_xchg + fr_xchg_concat_off(e, s->slot, parity);
  if ((st = run_mlp(e, s, x, Bl, d_scores, fold ? s->slot : -1)) != FR_OK) return st;
  return emit_scores(e, s, scores, Bl, d_scores);
}

extern "C" fr_status fr_shard_infer(fr_engine* e, const int32_t* idx, int B_global, float* scores_local, fr_stream s) {
  fr_status st = prep(e, &s, B_global, true, true);
  if (st != FR_OK) return st;
  if ((st = shard_check(e, s, B_global)) != FR_OK) return st;
  if (B_global == 0) return FR_OK;
  if (!idx || !scores_local) return fr_fail(e, FR_ERR_INVALID, "null idx/scores");
  if (*e->h_shard_err) return fr_fail(e, FR_ERR_STATE, "a previous sharded step timed out waiting for a peer rank");
  const int parity = (++s->shard_step) & 1;
  return run_or_replay(e, s, idx, scores_local, B_global, FR_GV_SHARD, 2, parity,
                       [&](int par) { return shard_infer_enqueue(e, s, idx, B_global, scores_local, par); });
}


// Which tables this rank needs indices for, ascending (= the column order of the sliced blocks): which = 0 the tables
// it owns (indices of ALL items of the global batch), which = 1 the replicated tables (indices of ITS items only).
extern "C" fr_status fr_shard_tables(fr_engine* e, int which, int32_t* ids, int* n) {
  if (!e || !n || which < 0 || which > 1) return fr_fail(e, FR_ERR_INVALID, "fr_shard_tables: bad argument");
  if (e->owner.empty() && e->world > 1) return fr_fail(e, FR_ERR_STATE, "fr_shard_init first");
  fr_shard_table_lists(e);
  const std::vector<int>& v = which == 0 ? e->owned_tables : e->repl_tables;
  *n = (int)v.size();
  if (ids)
    for (size_t i = 0; i < v.size(); i++) ids[i] = v[i];
  return FR_OK;
}

// fr_shard_infer fed the way the reference feeds its FPGAs (each receives only its own tables' indices,
// embedding_47_krnl.cpp:899-914 per bank): idx_owned [B_global][n_owned] for the tables of fr_shard_tables(0),
// idx_repl [B_global / world][n_repl] for the replicated tables of this rank's own items.  A rank then uploads
// B_global * n_owned + B_local * n_repl indices per step instead of B_global * T.
static fr_status shard_infer_sliced_enqueue(fr_engine* e, fr_stream_s* s, const int32_t* idx_owned, const int32_t* idx_repl,
                                            int n, int B_global, float* scores, int parity0) {
  s->f16 = false;
  const int Bl = B_global / e->world;
  const size_t o1 = (size_t)B_global * e->ipr_owned, r1 = (size_t)Bl * e->ipr_repl;   // int32 words per batch
  const size_t n_o = o1 * n, n_r = r1 * n;
  const int32_t* d_o = idx_owned;
  const int32_t* d_r = idx_repl;
  int32_t* stage_r = s->d_idx + (n_o + 3) / 4 * 4;
  if (n_o && n_r && idx_repl == idx_owned + (n_o + 3) / 4 * 4 && !is_device_ptr(idx_owned) && same_allocation(idx_owned, idx_repl)) {
    // one host buffer, the replicated block right behind the owned one (16-byte aligned): ONE copy -- every copy
    // costs the engine ~4 us on top of its bytes
    FR_CUDA(e, cudaMemcpyAsync(s->d_idx, idx_owned, ((n_o + 3) / 4 * 4 + n_r) * sizeof(int32_t), cudaMemcpyHostToDevice,
                               s->stream));
    d_o = s->d_idx;
    d_r = stage_r;
  } else if (n_o && !is_device_ptr(idx_owned)) {
    FR_CUDA(e, cudaMemcpyAsync(s->d_idx, idx_owned, n_o * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
    d_o = s->d_idx;
  }
  if (n_r && d_r == idx_repl && !is_device_ptr(idx_repl)) {
    FR_CUDA(e, cudaMemcpyAsync(stage_r, idx_repl, n_r * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
    d_r = stage_r;
  }
  const FrChunk* chunks = frk_sliced_chunks(e);
  if (!chunks) return FR_ERR_CUDA;   // (message left by the failing upload)
  const bool fold = e->knobs.shard_fold_wait != 0;
  float* d_scores = score_target(e, s, scores, n * Bl);
  for (int i = 0; i < n; i++) {   // consecutive steps of this worker's slot: the exchange buffers alternate
    const int parity = (parity0 + i) & 1;
    fr_status st = frk_shard_exchange(e, chunks, d_o + i * o1, e->ipr_owned, d_r + i * r1, e->ipr_repl, B_global, s->slot,
                                      parity, !fold, s->stream);
    if (st != FR_OK) return st;
    const float* x = e->d_xchg + fr_xchg_concat_off(e, s->slot, parity);
    if ((st = run_mlp(e, s, x, Bl, d_scores + (size_t)i * Bl, fold ? s->slot : -1)) != FR_OK) return st;
  }
  return emit_scores(e, s, scores, n * Bl, d_scores);
}

// n consecutive sharded steps in ONE call on one worker: idx_owned [n][B_global][n_owned], idx_repl [n][B_global / world]
// [n_repl], scores_local [n][B_global / world], all contiguous.  Host blocks travel in one copy each way (two when
// idx_repl does not start right behind idx_owned in the same allocation).
extern "C" fr_status fr_shard_infer_sliced_many(fr_engine* e, const int32_t* idx_owned, const int32_t* idx_repl, int n,
                                                int B_global, float* scores_local, fr_stream s) {
  if (n < 0 || n > 4095) return fr_fail(e, FR_ERR_INVALID, "fr_shard_infer_sliced_many: n=%d", n);
  fr_status st = prep(e, &s, B_global, true, true);
  if (st != FR_OK) return st;
  if ((st = shard_check(e, s, B_global)) != FR_OK) return st;
  if (B_global == 0 || n == 0) return FR_OK;
  fr_shard_table_lists(e);
  if (!scores_local || (!idx_owned && !e->owned_tables.empty()) || (!idx_repl && !e->repl_tables.empty()))
    return fr_fail(e, FR_ERR_INVALID, "null idx/scores");
  if (*e->h_shard_err) return fr_fail(e, FR_ERR_STATE, "a previous sharded step timed out waiting for a peer rank");
  const int Bl = B_global / e->world;
  fr_index_rows(e);
  const size_t ints = ((size_t)n * B_global * e->ipr_owned + 3) / 4 * 4 + (size_t)n * Bl * e->ipr_repl;
  if ((st = ensure_group_capacity(e, s, ints, (size_t)n * Bl)) != FR_OK) return st;
  const int parity0 = (s->shard_step + 1) & 1;   // the first step's exchange buffer; the graphs come in pairs by it
  s->shard_step += n;
  const void* key = idx_owned ? (const void*)idx_owned : (const void*)idx_repl;
  return run_or_replay(e, s, key, scores_local, B_global, FR_GV_SHARD_SLICED | (n << 4), 2, parity0,
                       [&](int par) { return shard_infer_sliced_enqueue(e, s, idx_owned, idx_repl, n, B_global, scores_local, par); },
                       idx_owned ? idx_repl : nullptr);
}

extern "C" fr_status fr_shard_infer_sliced(fr_engine* e, const int32_t* idx_owned, const int32_t* idx_repl, int B_global,
                                           float* scores_local, fr_stream s) {
  return fr_shard_infer_sliced_many(e, idx_owned, idx_repl, 1, B_global, scores_local, s);
}


// ---------------------------------------------------------------------------
extern "C" int64_t fr_merge_index(int64_t iA, int64_t iB, int64_t rowsB) { return iA * rowsB + iB; }

extern "C" fr_status fr_merge_tables(fr_engine* e, int a, int b, int dst) {
  fr_status st;
  if ((st = check_table(e, a)) != FR_OK || (st = check_table(e, b)) != FR_OK || (st = check_table(e, dst)) != FR_OK)
    return st;
  FrTable &A = e->tables[a], &B = e->tables[b], &M = e->tables[dst];
  if (!A.loaded || !B.loaded) return fr_fail(e, FR_ERR_STATE, "merge sources not loaded");
  if (e->table_dtype != FR_TABLE_F32) return fr_fail(e, FR_ERR_UNSUPPORTED, "fr_merge_tables builds fp32 tables only");
  if (M.dim != A.dim + B.dim || M.rows != A.rows * B.rows)
    return fr_fail(e, FR_ERR_INVALID, "merged table %d must be (%lld rows, dim %d)", dst, (long long)(A.rows * B.rows),
                   A.dim + B.dim);
  if ((st = ensure_table_mem(e, dst)) != FR_OK) return st;
  if ((st = frk_merge(e, A.d, A.rows, A.dim, B.d, B.rows, B.dim, M.d, e->default_stream->stream)) != FR_OK) return st;
  FR_CUDA(e, cudaStreamSynchronize(e->default_stream->stream));
  M.loaded = true;
  M.range_valid = false;
  e->f16_dirty = true;
  return FR_OK;
}
