// tma_probe.cu -- how fast can an SM take operands in through TMA?  (measurement tool, not product code)
//
// The tcgen05 MLP kernels are bound by the rate at which a CTA's shared-memory ring is filled
// (DESIGN.md section 4).  This probe runs the SAME load pattern with no MMAs at all: every CTA walks
// (M tile, K slice) like tc_linear_kernel's producer, loading a [rows_a x 128 B] box of an activation
// matrix A[M][K] (distinct rows per CTA) and a [rows_b x 128 B] box of a weight matrix W[N][K] (the same rows
// for every CTA: L2-hot) per K slice into a STAGES-deep ring; a consumer thread frees each slot as soon as
// it is full.  Bytes landed / elapsed = the feed ceiling of this pattern, per SM, for any number of CTAs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu -lcuda
//   tools/tma_probe            (prints a table)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(smem)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

struct P { int M, N, K, stages, rows_a, rows_b, n_b_boxes, split_producers, iters, store_rows, stores_per_tile, ks; };

// smem: ring of stages x (rows_a + n_b_boxes * rows_b) x 128 B
__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb,
                                                       const __grid_constant__ CUtensorMap to, const P p,
                                                       unsigned long long* out_bytes) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = (p.rows_a + p.n_b_boxes * p.rows_b) * 128 * p.ks;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty = full + p.stages;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], p.split_producers ? 2 : 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int num_kb = p.K / 32 / p.ks;   // stages per tile
  const int m_tiles = p.M / p.rows_a;
  unsigned long long bytes = 0;
  if (warp == 0 && lane == 0) {          // producer (A and B, or A only when split)
    uint32_t kc = 0;
    for (int it = 0; it < p.iters; it++)
      for (int kb = 0, t = (blockIdx.x + it * gridDim.x) % m_tiles; kb < num_kb; kb++, kc++) {
        const int s = kc % p.stages;
        mbar_wait(&empty[s], ((kc / p.stages) & 1) ^ 1);
        uint8_t* dst = smem + s * stage_bytes;
        if (p.split_producers) {
          mbar_expect_tx(&full[s], p.rows_a * 128);
          tma_load_2d(&ta, &full[s], dst, kb * 32, t * p.rows_a);
        } else if (p.ks > 1) {   // 3-D maps {32 floats, rows, K / 32}: one box = `ks` K slices of the rows
          mbar_expect_tx(&full[s], stage_bytes);
          tma_load_3d(&ta, &full[s], dst, 0, t * p.rows_a, kb * p.ks);
          for (int h = 0; h < p.n_b_boxes; h++)
            tma_load_3d(&tb, &full[s], dst + (p.rows_a + h * p.rows_b) * 128 * p.ks, 0, (h * p.rows_b) % p.N, kb * p.ks);
        } else {
          mbar_expect_tx(&full[s], stage_bytes);
          tma_load_2d(&ta, &full[s], dst, kb * 32, t * p.rows_a);
          for (int h = 0; h < p.n_b_boxes; h++)
            tma_load_2d(&tb, &full[s], dst + (p.rows_a + h * p.rows_b) * 128, kb * 32, (h * p.rows_b) % p.N);
        }
        bytes += stage_bytes;
      }
  } else if (warp == 1 && lane == 0 && p.split_producers) {   // second producer thread: B only
    uint32_t kc = 0;
    for (int it = 0; it < p.iters; it++)
      for (int kb = 0, t = (blockIdx.x + it * gridDim.x) % m_tiles; kb < num_kb; kb++, kc++) {
        const int s = kc % p.stages;
        mbar_wait(&empty[s], ((kc / p.stages) & 1) ^ 1);
        uint8_t* dst = smem + s * stage_bytes;
        mbar_expect_tx(&full[s], p.n_b_boxes * p.rows_b * 128);
        for (int h = 0; h < p.n_b_boxes; h++)
          tma_load_2d(&tb, &full[s], dst + (p.rows_a + h * p.rows_b) * 128, kb * 32, (h * p.rows_b) % p.N);
      }
  } else if (warp == 3 && lane == 0 && p.store_rows) {   // epilogue stand-in: stores_per_tile TMA stores per M tile
    uint8_t* sbuf = smem + p.stages * stage_bytes + 1024;      // [2][store_rows x 128 B], contents irrelevant
    for (int it = 0; it < p.iters; it++) {
      const int t = (blockIdx.x + it * gridDim.x) % m_tiles;
      for (int c = 0; c < p.stores_per_tile; c++) {
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        tma_store_2d(&to, sbuf + (c & 1) * p.store_rows * 128, (c * 32) % p.N, (t * p.rows_a) % p.M);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (warp == 2 && lane == 0) {   // consumer: free the slot as soon as it is full
    uint32_t kc = 0;
    for (int it = 0; it < p.iters; it++)
      for (int kb = 0, t = (blockIdx.x + it * gridDim.x) % m_tiles; kb < num_kb; kb++, kc++) {
        const int s = kc % p.stages;
        mbar_wait(&full[s], (kc / p.stages) & 1);
        mbar_arrive(&empty[s]);
      }
  }
  __syncthreads();
  if (threadIdx.x == 0 && bytes) atomicAdd(out_bytes, bytes);
}

typedef CUresult (*PFN_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// row-major [rows][K] fp32 seen as {32 floats, rows, K / 32} with strides {K * 4, 128} bytes: a box of `ks` K slices
static CUtensorMap make_map3(PFN_encode enc, void* base, int rows, int K, int box_rows, int ks, int* rc) {
  CUtensorMap m;
  const cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)(K / 32)};
  const cuuint64_t strides[2] = {(cuuint64_t)K * 4, 128};
  const cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)ks};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  *rc = (int)r;
  return m;
}

static CUtensorMap make_map(PFN_encode enc, void* base, int rows, int K, int box_rows) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  PFN_encode enc = (PFN_encode)fn;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const int M = 16384, N = 1024;
  float *A, *W, *O;
  unsigned long long* d_bytes;
  CK(cudaMalloc(&A, (size_t)M * 1024 * 4));
  CK(cudaMalloc(&W, (size_t)N * 1024 * 4));
  CK(cudaMalloc(&d_bytes, 8));
  CK(cudaMalloc(&O, (size_t)M * N * 4));
  CK(cudaMemset(A, 0, (size_t)M * 1024 * 4));
  CK(cudaMemset(W, 0, (size_t)N * 1024 * 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  printf("# %s, %d SMs, %d MHz nominal.  A[M=%d][K] distinct rows per CTA, W[N=%d][K] shared by all CTAs; fp32, 128-byte swizzled boxes\n",
         prop.name, sms, clk_khz / 1000, M, N);
  printf("# %-4s %-6s %-6s %-7s %-7s %-6s %-5s %-9s | %-9s %-10s %-9s %-9s %-12s\n", "K", "rows_a", "rows_b", "b_boxes", "stages", "ctas", "split", "stores", "us", "GB/s", "GB/s/SM", "B/clk/SM", "clk/stage");
  struct Cfg { int K, rows_a, rows_b, n_b, stages, ctas, split, store_rows, stores_per_tile, ks; };
  const Cfg cfgs[] = {
      // the GEMM patterns: 256-wide pair tile (A 128 rows + B 128 rows per CTA), 512-wide pair tile (A 128 + B 2 x 128)
      {1024, 128, 128, 1, 5, sms, 0}, {1024, 128, 128, 2, 3, sms, 0}, {1024, 128, 128, 2, 4, sms, 0},
      // fewer CTAs: is the limit per SM or chip-wide?
      {1024, 128, 128, 1, 5, 16, 0}, {1024, 128, 128, 1, 5, 64, 0}, {1024, 128, 128, 2, 3, 16, 0}, {1024, 128, 128, 2, 3, 64, 0},
      // box shapes: many small boxes / fewer large boxes for the same bytes
      {1024, 128, 64, 2, 5, sms, 0}, {1024, 128, 256, 1, 3, sms, 0}, {1024, 64, 64, 2, 6, sms, 0}, {1024, 256, 128, 1, 3, sms, 0},
      // two producer threads (A and B issued by different warps)
      {1024, 128, 128, 1, 5, sms, 1}, {1024, 128, 128, 2, 3, sms, 1},
      // deeper rings with the same stage shape (latency or bandwidth?)
      {1024, 128, 128, 1, 2, sms, 0}, {1024, 128, 128, 1, 3, sms, 0}, {1024, 128, 128, 1, 6, sms, 0},
      // only weights (L2-hot, same for all CTAs) / only activations
      {1024, 128, 128, 0, 6, sms, 0},
      // short K (layer 1 of the small model: 352)
      {352, 128, 128, 1, 5, sms, 0},
      // cost of ONE instruction against its box size: a single box per stage
      {1024, 32, 128, 0, 6, sms, 0}, {1024, 64, 128, 0, 6, sms, 0}, {1024, 128, 128, 0, 6, sms, 0}, {1024, 256, 128, 0, 6, sms, 0},
      // loads of a 256-wide tile + the epilogue's stores of the same tile (128 rows x 256 fp32 per CTA and M tile):
      // 32 stores of 32 rows x 128 B (one per epilogue warp and chunk) against 8 stores of 128 rows x 128 B
      {1024, 128, 128, 1, 5, sms, 0, 32, 32}, {1024, 128, 128, 1, 5, sms, 0, 128, 8},
      {352, 128, 128, 1, 5, sms, 0, 32, 32}, {352, 128, 128, 1, 5, sms, 0, 128, 8},
      // 3-D maps over the same row-major matrices: one instruction fetches 2 (4) K slices of the rows
      {1024, 128, 128, 1, 3, sms, 0, 0, 0, 2}, {1024, 128, 128, 1, 2, sms, 0, 0, 0, 2}, {1024, 128, 128, 2, 2, sms, 0, 0, 0, 2},
      {1024, 128, 128, 1, 1, sms, 0, 0, 0, 4}, {1024, 128, 64, 1, 4, sms, 0, 0, 0, 2}, {1024, 128, 128, 0, 6, sms, 0, 0, 0, 2},
  };
  for (const Cfg& c : cfgs) {
    const int ks = c.ks ? c.ks : 1;
    P p = {M, N, c.K, c.stages, c.rows_a, c.rows_b, c.n_b, c.split, 8 * 1024 / c.K, c.store_rows, c.stores_per_tile, ks};
    CUtensorMap to = make_map(enc, O, M, N, c.store_rows ? c.store_rows : 32);
    CUtensorMap ta = make_map(enc, A, M, c.K, c.rows_a), tb = make_map(enc, W, N, c.K, c.rows_b);
    if (ks > 1) {
      int r1 = 0, r2 = 0;
      ta = make_map3(enc, A, M, c.K, c.rows_a, ks, &r1);
      tb = make_map3(enc, W, N, c.K, c.rows_b, ks, &r2);
      if (r1 || r2) { printf("  (3-D map with strides {K*4, 128} refused: CUresult %d / %d)\n", r1, r2); continue; }
    }
    const int smem = c.stages * (c.rows_a + c.n_b * c.rows_b) * 128 * ks + 1024 + 2 * c.store_rows * 128 + 1024 + 64;
    if (smem > 227 * 1024) { printf("  (skip: %d B of smem)\n", smem); continue; }
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    float best = 1e30f;
    unsigned long long bytes = 0;
    for (int rep = 0; rep < 6; rep++) {
      CK(cudaMemset(d_bytes, 0, 8));
      CK(cudaEventRecord(e0));
      probe_kernel<<<c.ctas, 128, smem>>>(ta, tb, to, p, d_bytes);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
      CK(cudaMemcpy(&bytes, d_bytes, 8, cudaMemcpyDeviceToHost));
    }
    const double gbs = bytes / (best * 1e-3) / 1e9;
    char st[32];
    snprintf(st, sizeof(st), "%dx%dr", c.stores_per_tile, c.store_rows);
    const double stages_per_cta = (double)bytes / c.ctas / ((c.rows_a + c.n_b * c.rows_b) * 128 * ks);
    printf("  %-4d %-6d %-6d %-7d %-7d %-6d %-5d %-9s | %-9.1f %-10.0f %-9.1f %-9.1f %-12.0f %s\n", c.K, c.rows_a, c.rows_b, c.n_b, c.stages, c.ctas,
           c.split, c.store_rows ? st : "-", best * 1e3, gbs, gbs / c.ctas, gbs / c.ctas / (clk_khz / 1e6),
           best * 1e-3 * clk_khz * 1e3 / stages_per_cta, ks > 1 ? (ks == 2 ? "2 K slices per box" : "4 K slices per box") : "");
  }
  return 0;
}
