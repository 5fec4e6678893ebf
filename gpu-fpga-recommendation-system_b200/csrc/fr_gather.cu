// fr_gather.cu -- multi-table embedding lookup + concat (rows L1-L4 of SURVEY.md 8a).
//
// Replaces, on one B200, what the FPGA does with 30 HBM/DDR AXI masters and 17-44
// on-chip tables: load_single_embedding_N_tables (embedding_47_krnl.cpp:916-935,
// embedding_98_krnl.cpp:1016-1041, embedding_377_krnl.cpp:1180-1291) followed by the
// group_* / gather_*_embedding_streams re-packers (47: 964-1217).  The concat order
// is data (FrChunk list built from fr_segment_desc), not code.
//
// Mapping: one thread owns one 16-byte piece (float4 = one reference `axi_t`) of the
// concat vector and walks ITEMS items with it, so the piece descriptor lives in
// registers, the ITEMS index loads are issued back to back, then the ITEMS row
// loads (128-bit, read-only path, no L1 allocation), then the ITEMS stores.
// Consecutive threads own consecutive pieces: stores are fully coalesced and
// lanes that share a table read one contiguous row.  HBM-bound integer/byte work:
// no tensor cores, no shared memory (nothing is reused).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include <stddef.h>
#include <string.h>

#include "fr_common.h"

namespace {

constexpr int kItems = 4;  // items per thread (independent loads in flight)

__device__ __forceinline__ float4 ld_row16(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// One 4-float piece of a row stored as 2-byte elements (FR_TABLE_F16 / FR_TABLE_BF16): 8 bytes in,
// widened exactly to fp32 (the stated dequant: concat == float32(float16(row))).
template <int DT>
__device__ __forceinline__ float4 ld_row8(const float4* base, int64_t piece) {
  uint2 w;
  const uint2* p = reinterpret_cast<const uint2*>(base) + piece;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v2.u32 {%0,%1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(p));
  float4 v;
  if (DT == FR_TABLE_F16) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&w.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
    v = make_float4(a.x, a.y, b.x, b.y);
  } else {   // bf16 is the upper half of an fp32
    v = make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u), __uint_as_float(w.y << 16),
                    __uint_as_float(w.y & 0xFFFF0000u));
  }
  return v;
}

// Optional bounds check of one index: an offender is counted in err[0] (the first one also leaves its index column,
// value and item in err[1..3]) and row 0 is read in its place.
__device__ __forceinline__ int64_t checked_row(int64_t row, const FrChunk& ch, int item, int* err) {
  if ((uint64_t)row < (uint64_t)ch.rows) return row;
  if (ch.col4 == 0 && atomicAdd(err, 1) == 0) {   // one report per (item, table), by the piece at the row's head
    err[1] = ch.table;
    err[2] = (int)row;
    err[3] = item;
  }
  return 0;
}

// FR_TABLE_FP8: one 4-float piece is 4 bytes of E4M3, widened exactly (through fp16) to fp32.
__device__ __forceinline__ float fp8_to_float(uint32_t byte) {
  const __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)byte, __NV_E4M3);
  return __half2float(*reinterpret_cast<const __half*>(&h));
}
__device__ __forceinline__ float4 ld_row4(const float4* base, int64_t piece) {
  uint32_t w;
  const uint32_t* p = reinterpret_cast<const uint32_t*>(base) + piece;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(w) : "l"(p));
  return make_float4(fp8_to_float(w & 0xFFu), fp8_to_float((w >> 8) & 0xFFu), fp8_to_float((w >> 16) & 0xFFu), fp8_to_float(w >> 24));
}

// One index of item b: rows are row_words int32 words long; the piece's descriptor says where its table's index sits in
// the row and whether it is an int32 or (FR_IDX_PACKED, tables of at most 65536 rows) a uint16 (fr_index_at, fr_common.h).
__device__ __forceinline__ int64_t ld_index(const int32_t* __restrict__ idx, size_t b, int row_words, int idx_off) {
  return fr_index_at(idx + b * (size_t)row_words, idx_off);
}

// A piece descriptor in two 128-bit loads (the compiler splits the struct copy into three or four).
static_assert(offsetof(FrChunk, table) == 8 && offsetof(FrChunk, stride4) == 12 && offsetof(FrChunk, col4) == 16 &&
              offsetof(FrChunk, rows) == 20 && offsetof(FrChunk, idx_off) == 24, "ld_chunk unpacks this layout");
__host__ __device__ __forceinline__ FrChunk unpack_chunk(const uint4 lo, const uint4 hi) {
  FrChunk ch;
  ch.base = reinterpret_cast<const float4*>((uint64_t)lo.x | ((uint64_t)lo.y << 32));
  ch.table = (int)lo.z;
  ch.stride4 = (int)lo.w;
  ch.col4 = (int)hi.x;
  ch.rows = (int)hi.y;
  ch.idx_off = (int)hi.z;
  ch.pad_ = 0;
  return ch;
}
__device__ __forceinline__ FrChunk ld_chunk(const FrChunk* __restrict__ p) {
  return unpack_chunk(__ldg(reinterpret_cast<const uint4*>(p)), __ldg(reinterpret_cast<const uint4*>(p) + 1));
}

// PUSH = false: out4 is the local [B][C] buffer.
// PUSH = true : peer_out[r] is rank r's exchange buffer; item b lands on rank
//               b / items_per_rank at local row b % items_per_rank (NVLink peer stores).
template <bool ROUND, bool PUSH, int DT, bool OUT16 = false>
__global__ void __launch_bounds__(256) gather_concat_kernel(const FrChunk* __restrict__ chunks,
                                                            const int* __restrict__ chunk_ids, int n_chunks,
                                                            const int32_t* __restrict__ idx, int T /* int32 words per index row */, int b_begin,
                                                            int b_end, float4* __restrict__ out4,
                                                            float4* const* __restrict__ peer_out, int C,
                                                            int items_per_rank, long long peer_off4,
                                                            int* __restrict__ idx_err) {
  // one thread = one piece x kItems consecutive items; block = (pieces rounded to 32, <= 128) x item groups.
  // (A flattened (piece, item group) space, which keeps every lane busy when a rank's piece subset is not a
  // multiple of 32, was measured SLOWER for the sharded push: 13.4 against 12.0 us per step at two ranks.)
  const int ci = blockIdx.x * blockDim.x + threadIdx.x;
  if (ci >= n_chunks) return;
  const int b0 = b_begin + (blockIdx.y * blockDim.y + threadIdx.y) * kItems;
  const int c = chunk_ids ? chunk_ids[ci] : ci;
  const FrChunk ch = ld_chunk(chunks + c);

  int64_t row[kItems];
#pragma unroll
  for (int i = 0; i < kItems; i++) {
    const int b = b0 + i;
    row[i] = (b < b_end) ? ld_index(idx, (size_t)b, T, ch.idx_off) : 0;
  }
  if (idx_err) {   // fr_set_check_indices: the reference never checks (embedding_47_krnl.cpp:925-934)
#pragma unroll
    for (int i = 0; i < kItems; i++) row[i] = checked_row(row[i], ch, b0 + i, idx_err);
  }
  float4 v[kItems];
#pragma unroll
  for (int i = 0; i < kItems; i++)  // 64-bit addressing: 100 M rows x 128 B = 12.8 GB tables
    v[i] = DT == FR_TABLE_F32 ? ld_row16(ch.base + row[i] * ch.stride4 + ch.col4)
           : DT == FR_TABLE_FP8 ? ld_row4(ch.base, row[i] * ch.stride4 + ch.col4)
                                : ld_row8<DT == FR_TABLE_BF16 ? FR_TABLE_BF16 : FR_TABLE_F16>(ch.base, row[i] * ch.stride4 + ch.col4);
  // Programmatic dependent launch: indices and tables are not written by the kernels of the
  // preceding batch, so everything above overlaps its tail; the concat buffer is (the first MLP
  // layer of the previous batch read it), so the stores wait for the grid dependency.  Both
  // instructions are no-ops when the launch carries no programmatic attribute.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
  for (int i = 0; i < kItems; i++) {
    const int b = b0 + i;
    if (b >= b_end) break;
    float4 o = v[i];
    if (ROUND) {
      o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w);
    }
    if (PUSH) {
      const int r = b / items_per_rank;
      peer_out[r][peer_off4 + (size_t)(b - r * items_per_rank) * C + c] = o;
    } else if (OUT16) {   // fp16 concat vectors (tc_f16): the piece is 4 halves
      const __half2 lo = __floats2half2_rn(o.x, o.y), hi = __floats2half2_rn(o.z, o.w);
      uint2 w;
      w.x = *reinterpret_cast<const uint32_t*>(&lo);
      w.y = *reinterpret_cast<const uint32_t*>(&hi);
      reinterpret_cast<uint2*>(out4)[(size_t)b * C + c] = w;
    } else {
      out4[(size_t)b * C + c] = o;
    }
  }
}

__device__ __forceinline__ uint16_t quant16(float x, int dt) {
  return dt == FR_TABLE_F16 ? __half_as_ushort(__float2half_rn(x)) : __bfloat16_as_ushort(__float2bfloat16_rn(x));
}
__device__ __forceinline__ float dequant16(uint16_t h, int dt) {
  return dt == FR_TABLE_F16 ? __half2float(__ushort_as_half(h)) : __uint_as_float((uint32_t)h << 16);
}
__global__ void quantize_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t n, int dt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = quant16(src[i], dt);
}
__global__ void dequantize_kernel(const uint16_t* __restrict__ src, float* __restrict__ dst, int64_t n, int dt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = dequant16(src[i], dt);
}
// FP8 (E4M3): round to nearest even, saturating at +-448 (NaN stays NaN)
__device__ __forceinline__ uint8_t quant8(float x) { return (uint8_t)__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3); }
__global__ void quantize8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = quant8(src[i]);
}
__global__ void dequantize8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = fp8_to_float(src[i]);
}
__global__ void fill_reference8_kernel(uint8_t* __restrict__ t, int64_t n, int dim, int64_t filled_rows) {
  const uint8_t one = quant8(1.f);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / dim;
    t[i] = (r < filled_rows && (r & 1) == 0) ? one : (uint8_t)0;
  }
}
// 2-byte variants of the two device fills: the fp32 value of the fp32 fill, rounded to nearest even
__global__ void fill_reference16_kernel(uint16_t* __restrict__ t, int64_t n, int dim, int64_t filled_rows, int dt) {
  const uint16_t one = quant16(1.f, dt);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / dim;
    t[i] = (r < filled_rows && (r & 1) == 0) ? one : (uint16_t)0;
  }
}

__global__ void fill_reference_kernel(float4* __restrict__ t, int64_t n4, int dim4, int64_t filled_rows) {
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / dim4;
    t[i] = (r < filled_rows && (r & 1) == 0) ? one : zero;
  }
}

__device__ __forceinline__ uint32_t hash_bits(uint32_t seed, uint32_t table, uint64_t row, uint32_t col) {
  uint64_t z = row * 0x9E3779B97F4A7C15ull + ((uint64_t)table << 40) + ((uint64_t)col << 28) + seed;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  const uint32_t h = (uint32_t)(z >> 16);
  const uint32_t expo = 118u + ((h >> 23) & 0xFFu) % 9u;
  return (h & 0x80000000u) | (expo << 23) | (h & 0x007FFFFFu);
}

__global__ void fill_hash_kernel(uint32_t* __restrict__ t, int64_t n, int dim, uint32_t seed, uint32_t table) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    t[i] = hash_bits(seed, table, (uint64_t)(i / dim), (uint32_t)(i % dim));
}

__global__ void fill_hash16_kernel(uint16_t* __restrict__ t, int64_t n, int dim, uint32_t seed, uint32_t table, int dt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    t[i] = quant16(__uint_as_float(hash_bits(seed, table, (uint64_t)(i / dim), (uint32_t)(i % dim))), dt);
}

__global__ void fill_hash8_kernel(uint8_t* __restrict__ t, int64_t n, int dim, uint32_t seed, uint32_t table) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    t[i] = quant8(__uint_as_float(hash_bits(seed, table, (uint64_t)(i / dim), (uint32_t)(i % dim))));
}

__global__ void merge_kernel(const float4* __restrict__ A, int dimA4, const float4* __restrict__ B, int64_t rowsB,
                             int dimB4, float4* __restrict__ M, int64_t n4) {
  const int dm4 = dimA4 + dimB4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / dm4;
    const int c = (int)(i - r * dm4);
    const int64_t ia = r / rowsB, ib = r - ia * rowsB;
    M[i] = (c < dimA4) ? A[ia * dimA4 + c] : B[ib * dimB4 + (c - dimA4)];
  }
}

// Wt[o][i] = rna_tf32(W[i][o]); 32x32 smem tile transpose.
__global__ void transpose_round_kernel(const float* __restrict__ W, int in, int out, float* __restrict__ Wt) {
  __shared__ float tile[32][33];
  const int i0 = blockIdx.y * 32, o0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r, o = o0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < in && o < out) ? W[(size_t)i * out + o] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int o = o0 + r, i = i0 + threadIdx.x;
    if (o < out && i < in) Wt[(size_t)o * in + i] = round_tf32(tile[threadIdx.x][r]);
  }
}

int grid_for(int64_t n, int block, int sm_count) {
  int64_t g = (n + block - 1) / block;
  const int64_t cap = (int64_t)sm_count * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <bool ROUND, bool PUSH>
void launch_gather(const fr_engine* e, const int* d_ids, int n_chunks, const int32_t* d_idx, int b_begin, int b_end,
                   float4* out4, float4* const* peers, int items_per_rank, cudaStream_t st, long long peer_off4 = 0,
                   const FrChunk* chunks = nullptr, int idx_cols = 0, bool out_f16 = false) {
  if (!chunks) {   // full index rows; else a column-sliced block with its own descriptors and row length
    chunks = e->d_chunks;
    idx_cols = e->ipr_full;
  }
  const int C = e->D / 4;
  const int n_items = b_end - b_begin;
  // pieces along x: rounded up to a multiple of 32 (at most 128), or -- a rank's share of a sharded model can be a
  // handful of pieces -- to 8 / 16, so that a warp covers several items instead of idling most of its lanes
  int bx = (n_chunks + 31) / 32 * 32;
  if (bx > 128) bx = 128;
  if (n_chunks <= 8) bx = 8;
  else if (n_chunks <= 16) bx = 16;
  const int by = 256 / bx;
  dim3 block(bx, by);
  dim3 grid((n_chunks + bx - 1) / bx, (n_items + by * kItems - 1) / (by * kItems));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ((e->knobs.pdl_mask & 4) && !PUSH) ? 1 : 0;
  auto kern = e->table_dtype == FR_TABLE_F32 ? gather_concat_kernel<ROUND, PUSH, FR_TABLE_F32>
              : e->table_dtype == FR_TABLE_F16 ? gather_concat_kernel<ROUND, PUSH, FR_TABLE_F16>
              : e->table_dtype == FR_TABLE_BF16 ? gather_concat_kernel<ROUND, PUSH, FR_TABLE_BF16>
                                                : gather_concat_kernel<ROUND, PUSH, FR_TABLE_FP8>;
  if (out_f16 && !PUSH)
    kern = e->table_dtype == FR_TABLE_F32 ? gather_concat_kernel<false, false, FR_TABLE_F32, true>
           : e->table_dtype == FR_TABLE_F16 ? gather_concat_kernel<false, false, FR_TABLE_F16, true>
           : e->table_dtype == FR_TABLE_BF16 ? gather_concat_kernel<false, false, FR_TABLE_BF16, true>
                                             : gather_concat_kernel<false, false, FR_TABLE_FP8, true>;
  int* d_err = nullptr;
  if (e->check_indices && e->h_idx_err) cudaHostGetDevicePointer(&d_err, e->h_idx_err, 0);
  cudaLaunchKernelEx(&cfg, kern, chunks, d_ids, n_chunks, d_idx, idx_cols, b_begin, b_end, out4, peers, C, items_per_rank,
                     peer_off4, d_err);
}

}  // namespace

// Index-row layouts for the engine's FR_OPT_INDEX_FORMAT: full rows (all tables) and, when sharded, the two
// column-sliced blocks.  FR_IDX_I32: column i at byte 4 i.  FR_IDX_PACKED: the int32 columns (tables of more than 65536
// rows) first, in list order, then the uint16 columns, the row padded to a multiple of 4 bytes.
static void index_row_layout_rows(const int64_t* rows, int n, int format, int* off, int* words) {
  int pos = 0;
  if (format == FR_IDX_I32) {
    for (int i = 0; i < n; i++, pos += 4) off[i] = pos;
  } else {
    for (int i = 0; i < n; i++)
      if (rows[i] > 65536) { off[i] = pos; pos += 4; }
    for (int i = 0; i < n; i++)
      if (rows[i] <= 65536) { off[i] = pos | kIdx16; pos += 2; }
  }
  *words = (pos + 3) / 4;
}
static void index_row_layout(const fr_engine* e, const std::vector<int>& tables, std::vector<int>* off, int* words) {
  std::vector<int64_t> rows(tables.size());
  for (size_t i = 0; i < tables.size(); i++) rows[i] = e->tables[tables[i]].rows;
  off->assign(tables.size(), 0);
  index_row_layout_rows(rows.data(), (int)rows.size(), e->index_format, off->data(), words);
}
// The layout rule on its own, for the CPU tests (off[i]: byte offset, bit 31 set for a uint16 column).
extern "C" int frdbg_index_layout(const int64_t* rows, int n, int format, int32_t* off, int* words) {
  if (!rows || !off || !words || n < 0 || (format != FR_IDX_I32 && format != FR_IDX_PACKED)) return -1;
  index_row_layout_rows(rows, n, format, off, words);
  return 0;
}

void fr_index_rows(fr_engine* e) {
  std::vector<int> all(e->tables.size());
  for (size_t t = 0; t < all.size(); t++) all[t] = (int)t;
  index_row_layout(e, all, &e->idx_off_full, &e->ipr_full);
  fr_shard_table_lists(e);
  index_row_layout(e, e->owned_tables, &e->idx_off_owned, &e->ipr_owned);
  index_row_layout(e, e->repl_tables, &e->idx_off_repl, &e->ipr_repl);
}

fr_status frk_upload_chunks(fr_engine* e) {
  const int C = e->D / 4;
  fr_index_rows(e);
  std::vector<FrChunk> h(C);
  std::vector<FrFuseChunk> hf(C);
  std::vector<char> covered(C, 0);
  for (const fr_segment_desc& s : e->segs) {
    const FrTable& t = e->tables[s.table];
    for (int k = 0; k < s.len / 4; k++) {
      FrChunk& c = h[s.dst / 4 + k];
      c.base = reinterpret_cast<const float4*>(t.d);
      c.table = s.table;
      c.stride4 = t.dim / 4;
      c.col4 = s.col / 4 + k;
      c.rows = t.rows > 0x7FFFFFFF ? 0x7FFFFFFF : (int)t.rows;
      c.idx_off = e->idx_off_full[s.table];
      c.pad_ = 0;
      hf[s.dst / 4 + k] = {c.base, c.table, (c.stride4 << 8) | c.col4};
      covered[s.dst / 4 + k] = 1;
    }
  }
  for (int i = 0; i < C; i++)
    if (!covered[i]) return fr_fail(e, FR_ERR_INVALID, "concat float %d is not covered by any segment", i * 4);
  if (!e->d_chunks) FR_CUDA(e, cudaMalloc(&e->d_chunks, sizeof(FrChunk) * C));
  FR_CUDA(e, fr_h2d(e, e->d_chunks, h.data(), sizeof(FrChunk) * C));
  if (!e->d_fchunks) FR_CUDA(e, cudaMalloc(&e->d_fchunks, sizeof(FrFuseChunk) * C));
  FR_CUDA(e, fr_h2d(e, e->d_fchunks, hf.data(), sizeof(FrFuseChunk) * C));
  e->chunks_dirty = false;
  return FR_OK;
}

fr_status frk_gather(fr_engine* e, const int32_t* d_idx, int B, float* d_out, bool round_tf32, cudaStream_t st, bool out_f16) {
  if (B == 0) return FR_OK;
  if (out_f16)
    launch_gather<false, false>(e, nullptr, e->D / 4, d_idx, 0, B, reinterpret_cast<float4*>(d_out), nullptr, 1, st, 0, nullptr, 0,
                                true);
  else if (round_tf32)
    launch_gather<true, false>(e, nullptr, e->D / 4, d_idx, 0, B, reinterpret_cast<float4*>(d_out), nullptr, 1, st);
  else
    launch_gather<false, false>(e, nullptr, e->D / 4, d_idx, 0, B, reinterpret_cast<float4*>(d_out), nullptr, 1, st);
  e->launches++;
  FR_CUDA(e, cudaGetLastError());
  return FR_OK;
}

// Index staging without the copy engine (replaces the read() + cudaMemcpyAsync of cuda_server.c:425-461 for
// page-locked caller buffers): the SMs read the caller's mapped host buffer over PCIe with 16-byte loads,
// every element exactly once, one load per thread so the whole batch is one round trip deep, and write it to
// the worker's device index buffer.  A memcpy node costs the copy engine ~2-3 us of set-up per batch and the
// batches of all workers queue on that one engine; kernels of different workers overlap.
// ld.cv: the same host buffer carries new indices on every replay -- never serve it from a cache.
__global__ void stage_idx_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n16) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = __ldcv(src + i);
}

__global__ void to_f16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 w;
    w.x = *reinterpret_cast<const uint32_t*>(&lo);
    w.y = *reinterpret_cast<const uint32_t*>(&hi);
    dst[i] = w;
  }
}

fr_status frk_to_f16(fr_engine* e, const float* src, void* dst, int64_t n, cudaStream_t st) {
  const int64_t n4 = n / 4;
  to_f16_kernel<<<grid_for(n4, 256, e->sm_count), 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(dst), n4);
  e->launches++;
  FR_CUDA(e, cudaGetLastError());
  return FR_OK;
}

__global__ void round_tf32_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = src[i];
    v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
    dst[i] = v;
  }
}

fr_status frk_round_tf32(fr_engine* e, const float* src, float* dst, int64_t n, cudaStream_t st) {
  const int64_t n4 = n / 4;
  round_tf32_kernel<<<grid_for(n4, 256, e->sm_count), 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), n4);
  e->launches++;
  FR_CUDA(e, cudaGetLastError());
  return FR_OK;
}

fr_status frk_stage_idx(fr_engine* e, const void* mapped_src, int32_t* d_dst, size_t bytes, cudaStream_t st) {
  const int n16 = (int)(bytes / 16);
  int blocks = (n16 + 255) / 256;
  if (blocks > 8 * e->sm_count) blocks = 8 * e->sm_count;
  stage_idx_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(mapped_src), reinterpret_cast<uint4*>(d_dst), n16);
  e->launches++;
  FR_CUDA(e, cudaGetLastError());
  return FR_OK;
}

fr_status frk_fill_reference(fr_engine* e, float* d, int64_t rows, int dim, int64_t debug_rows, cudaStream_t st) {
  // host.cpp:66-88: pairs (2i, 2i+1) for i < rows/2 (or < debug_rows/2 with DEBUG)
  int64_t pairs = rows / 2;
  if (debug_rows > 0 && debug_rows / 2 < pairs) pairs = debug_rows / 2;
  const int64_t n4 = rows * dim / 4;
  if (e->table_dtype == FR_TABLE_FP8)
    fill_reference8_kernel<<<grid_for(rows * dim, 256, e->sm_count), 256, 0, st>>>(reinterpret_cast<uint8_t*>(d), rows * dim, dim, pairs * 2);
  else if (e->table_dtype != FR_TABLE_F32)
    fill_reference16_kernel<<<grid_for(rows * dim, 256, e->sm_count), 256, 0, st>>>(reinterpret_cast<uint16_t*>(d), rows * dim,
                                                                                   dim, pairs * 2, e->table_dtype);
  else
    fill_reference_kernel<<<grid_for(n4, 256, e->sm_count), 256, 0, st>>>(reinterpret_cast<float4*>(d), n4, dim / 4,
                                                                          pairs * 2);
  e->launches++;
  FR_CUDA(e, cudaGetLastError());
  return FR_OK;
}

fr_status frk_fill_hash(fr_engine* e, float* d, uint32_t seed, int table, int64_t rows, int dim, cudaStream_t st) {
  const int64_t n = rows * dim;
  if (e->table_dtype == FR_TABLE_FP8)
    fill_hash8_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, st>>>(reinterpret_cast<uint8_t*>(d), n, dim, seed, (uint32_t)table);
  else if (e->table_dtype != FR_TABLE_F32)
    fill_hash16_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, st>>>(reinterpret_cast<uint16_t*>(d), n, dim, seed,
                                                                     (uint32_t)table, e->table_dtype);
  else
    fill_hash_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, st>>>(reinterpret_cast<uint32_t*>(d), n, dim, seed,
                                                                    (uint32_t)table);
  e->launches++;
  FR_CUDA(e, cudaGetLastError());
  return FR_OK;
}

fr_status frk_quantize(fr_engine* e, const float* src, void* dst, int64_t n, cudaStream_t st) {
  if (e->table_dtype == FR_TABLE_FP8) quantize8_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, st>>>(src, reinterpret_cast<uint8_t*>(dst), n);
  else quantize_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, st>>>(src, reinterpret_cast<uint16_t*>(dst), n, e->table_dtype);
  e->launches++;
  FR_CUDA(e, cudaGetLastError());
  return FR_OK;
}

fr_status frk_dequantize(fr_engine* e, const void* src, float* dst, int64_t n, cudaStream_t st) {
  if (e->table_dtype == FR_TABLE_FP8) dequantize8_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(src), dst, n);
  else dequantize_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, st>>>(reinterpret_cast<const uint16_t*>(src), dst, n, e->table_dtype);
  e->launches++;
  FR_CUDA(e, cudaGetLastError());
  return FR_OK;
}

fr_status frk_merge(fr_engine* e, const float* A, int64_t rowsA, int dimA, const float* B, int64_t rowsB, int dimB,
                    float* M, cudaStream_t st) {
  const int64_t n4 = rowsA * rowsB * (dimA + dimB) / 4;
  merge_kernel<<<grid_for(n4, 256, e->sm_count), 256, 0, st>>>(reinterpret_cast<const float4*>(A), dimA / 4,
                                                               reinterpret_cast<const float4*>(B), rowsB, dimB / 4,
                                                               reinterpret_cast<float4*>(M), n4);
  e->launches++;
  FR_CUDA(e, cudaGetLastError());
  return FR_OK;
}

fr_status frk_transpose_round_tf32(fr_engine* e, const float* W, int in, int out, float* Wt, cudaStream_t st) {
  dim3 grid((out + 31) / 32, (in + 31) / 32), block(32, 8);
  transpose_round_kernel<<<grid, block, 0, st>>>(W, in, out, Wt);
  e->launches++;
  FR_CUDA(e, cudaGetLastError());
  return FR_OK;
}

// Sharded step, phase 1: owned tables for the GLOBAL batch are pushed
// to the rank that owns each item with NVLink peer stores; replicated (on-chip
// class) tables are gathered locally for this rank's items only.
static fr_status build_shard_lists(fr_engine* e) {
  std::vector<int> owned, repl;
  const int C = e->D / 4;
  std::vector<int> table_of(C, -1);
  for (const fr_segment_desc& s : e->segs)
    for (int k = 0; k < s.len / 4; k++) table_of[s.dst / 4 + k] = s.table;
  for (int c = 0; c < C; c++) {
    const int o = e->owner.empty() ? e->rank : e->owner[table_of[c]];
    if (o < 0) repl.push_back(c);
    else if (o == e->rank) owned.push_back(c);
  }
  e->n_owned = (int)owned.size();
  e->n_repl = (int)repl.size();
  cudaFree(e->d_owned_ids);   // (rebuilt when the index format changes)
  cudaFree(e->d_repl_ids);
  e->d_owned_ids = e->d_repl_ids = nullptr;
  if (e->n_owned) {
    FR_CUDA(e, cudaMalloc(&e->d_owned_ids, sizeof(int) * owned.size()));
    FR_CUDA(e, fr_h2d(e, e->d_owned_ids, owned.data(), sizeof(int) * owned.size()));
  }
  if (e->n_repl) {
    FR_CUDA(e, cudaMalloc(&e->d_repl_ids, sizeof(int) * repl.size()));
    FR_CUDA(e, fr_h2d(e, e->d_repl_ids, repl.data(), sizeof(int) * repl.size()));
  }
  // column-sliced variant of the descriptors: `table` = the column of that table in the caller's sliced block
  fr_shard_table_lists(e);
  std::vector<FrChunk> ch(C);
  FR_CUDA(e, cudaMemcpy(ch.data(), e->d_chunks, sizeof(FrChunk) * C, cudaMemcpyDeviceToHost));
  std::vector<int> col_of(e->tables.size(), 0);
  for (size_t i = 0; i < e->owned_tables.size(); i++) col_of[e->owned_tables[i]] = (int)i;
  for (size_t i = 0; i < e->repl_tables.size(); i++) col_of[e->repl_tables[i]] = (int)i;
  fr_index_rows(e);
  for (int c = 0; c < C; c++) {
    const int t = table_of[c], col = col_of[t];
    ch[c].table = col;
    ch[c].idx_off = (e->owner.empty() ? e->rank : e->owner[t]) < 0 ? e->idx_off_repl[col] : e->idx_off_owned[col];
  }
  if (!e->d_chunks_sliced) FR_CUDA(e, cudaMalloc(&e->d_chunks_sliced, sizeof(FrChunk) * C));
  FR_CUDA(e, fr_h2d(e, e->d_chunks_sliced, ch.data(), sizeof(FrChunk) * C));
  e->shard_lists_built = true;
  return FR_OK;
}

void fr_shard_table_lists(fr_engine* e) {
  if (!e->owned_tables.empty() || !e->repl_tables.empty()) return;
  for (int t = 0; t < (int)e->tables.size(); t++) {
    const int o = e->owner.empty() ? e->rank : e->owner[t];
    if (o < 0) e->repl_tables.push_back(t);
    else if (o == e->rank) e->owned_tables.push_back(t);
  }
}

// ---- the exchange of a table-sharded step ---------------------------------------------------------
// 1. push: gather_concat_kernel<_, PUSH = true> looks up this rank's OWNED pieces for every item of the global batch
//    and stores them straight into the concat buffer of the rank that owns the item (128-bit NVLink peer stores:
//    lookup and all-to-all are one kernel);
// 2. the REPLICATED (on-chip class) pieces of this rank's own items, gathered locally into its own buffer;
// 3. shard_signal_wait_kernel (one warp): the kernel boundary has completed the pushes; a system fence, then lane t
//    publishes "rank r has pushed step n of this slot" into rank t's flag block (st.release.sys) and polls this rank's
//    own flag of rank t (ld.acquire.sys) until it shows step n.  The step number lives in device memory (bumped
//    here), so no launch has a per-step argument and the whole step replays as a CUDA graph.
// A waiting warp costs nothing, which matters: with a dozen worker streams in flight some step is always waiting.
// (Folding publish into the push kernel's last block -- every block then needs its own system fence -- and the wait
// into the first MLP kernel -- a spinning tcgen05 grid holds its SMs -- were both measured slower: 15.2-15.7 against
// 12-13 us per step at two ranks; the second can even deadlock two ranks that start different workers first.)
namespace {
__global__ void shard_signal_wait_kernel(float* const* __restrict__ peer_base, long long flags_off_floats, int rank,
                                         int world, int* step_counter, int* err, long long timeout_cycles, int do_wait) {
  const int t = threadIdx.x;
  const int step = *reinterpret_cast<volatile int*>(step_counter) + 1;   // kernels of one slot are stream-ordered
  __syncwarp();
  __threadfence_system();   // the push kernel(s) before this one are complete; make their stores visible system-wide
  if (t < world) {
    int* f = reinterpret_cast<int*>(peer_base[t] + flags_off_floats) + rank;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(f), "r"(step) : "memory");
  }
  if (t == 0) *step_counter = step;
  if (t < world && do_wait) {
    const int* mine = reinterpret_cast<const int*>(peer_base[rank] + flags_off_floats) + t;
    const long long t0 = clock64();
    int v;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (v < step && clock64() - t0 > timeout_cycles) {
        *reinterpret_cast<volatile int*>(err) = 1;   // reported by fr_sync and by the next sharded call
        break;
      }
    } while (v < step);
  }
}

// Wait only (the first MLP kernel's own poll, FR_SHARD_FOLD, has nothing to follow in the FP32 path).
__global__ void shard_wait_kernel(const int* __restrict__ flags, int world, const int* step_counter, int* err,
                                  long long timeout_cycles) {
  const int t = threadIdx.x;
  const int step = *reinterpret_cast<const volatile int*>(step_counter);
  if (t >= world) return;
  const long long t0 = clock64();
  int v;
  do {
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + t) : "memory");
    if (v < step && clock64() - t0 > timeout_cycles) {
      *reinterpret_cast<volatile int*>(err) = 1;
      break;
    }
  } while (v < step);
}
}  // namespace

static fr_status ensure_shard_lists(fr_engine* e) {
  if (e->shard_lists_built) return FR_OK;
  std::lock_guard<std::mutex> g(e->mu);
  if (e->shard_lists_built) return FR_OK;
  return build_shard_lists(e);
}

fr_status frk_shard_exchange(fr_engine* e, const FrChunk* chunks, const int32_t* d_idx_owned, int T_owned,
                             const int32_t* d_idx_repl, int T_repl, int B_global, int slot, int parity, bool wait, cudaStream_t st) {
  fr_status s = ensure_shard_lists(e);
  if (s != FR_OK) return s;
  const int per = B_global / e->world;
  const bool round = (e->precision == FR_PREC_TF32);
  // concat buffer `parity` of slot `slot` in every rank's exchange region (same layout on all ranks)
  const long long off4 = (long long)(fr_xchg_concat_off(e, slot, parity) / 4);
  float4* const* peers = reinterpret_cast<float4* const*>(e->d_peer_ptrs);
  if (e->n_owned) {
    if (round) launch_gather<true, true>(e, e->d_owned_ids, e->n_owned, d_idx_owned, 0, B_global, nullptr, peers, per, st, off4, chunks, T_owned);
    else launch_gather<false, true>(e, e->d_owned_ids, e->n_owned, d_idx_owned, 0, B_global, nullptr, peers, per, st, off4, chunks, T_owned);
    e->launches++;
  }
  if (e->n_repl) {
    // local items only, written into this rank's own buffer; the index block's row 0 is global item rank * per
    float4* own = reinterpret_cast<float4*>(e->d_xchg) + off4;
    if (round) launch_gather<true, false>(e, e->d_repl_ids, e->n_repl, d_idx_repl, 0, per, own, nullptr, 1, st, 0, chunks, T_repl);
    else launch_gather<false, false>(e, e->d_repl_ids, e->n_repl, d_idx_repl, 0, per, own, nullptr, 1, st, 0, chunks, T_repl);
    e->launches++;
  }
  FR_CUDA(e, cudaGetLastError());
  int* d_err = nullptr;
  FR_CUDA(e, cudaHostGetDevicePointer(&d_err, e->h_shard_err, 0));
  shard_signal_wait_kernel<<<1, 32, 0, st>>>(e->d_peer_ptrs, (long long)fr_xchg_flags_off(e, slot), e->rank, e->world,
                                             e->d_step + slot, d_err, kShardTimeoutCycles, wait ? 1 : 0);
  e->launches++;
  FR_CUDA(e, cudaGetLastError());
  return FR_OK;
}

const FrChunk* frk_sliced_chunks(fr_engine* e) {
  return ensure_shard_lists(e) == FR_OK ? e->d_chunks_sliced : nullptr;
}

fr_status frk_shard_wait(fr_engine* e, int slot, cudaStream_t st) {
  int* d_err = nullptr;
  FR_CUDA(e, cudaHostGetDevicePointer(&d_err, e->h_shard_err, 0));
  shard_wait_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const int*>(e->d_xchg + fr_xchg_flags_off(e, slot)), e->world,
                                      e->d_step + slot, d_err, kShardTimeoutCycles);
  e->launches++;
  FR_CUDA(e, cudaGetLastError());
  return FR_OK;
}

// Host-side evaluation of the lookup kernels' index decode (same inline function), for the CPU tests of the packed
// index rows: index of item b / descriptor offset idx_off out of rows of row_words int32 words.
extern "C" int64_t frdbg_index_at(const int32_t* rows, int64_t b, int row_words, int idx_off) {
  return fr_index_at(rows + b * (int64_t)row_words, idx_off);
}

// The lookup kernels' descriptor unpack evaluated on the host (CPU test: out == in for any descriptor).
extern "C" void frdbg_chunk_roundtrip(const void* chunk32, void* out32) {
  uint4 w[2];
  memcpy(w, chunk32, 32);
  const FrChunk ch = unpack_chunk(w[0], w[1]);
  memcpy(out32, &ch, 32);
}
