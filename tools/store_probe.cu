// store_probe.cu -- how fast can an SM write a GEMM tile's output?  (measurement tool, not product code)
//
// tools/tc_prof.py shows the epilogue of tc_pair_kernel draining a 128-row x 256-column fp32 tile (128 KB per CTA) in
// ~4-5 us next to the main loop's operand reads, and 0.9 us with the stores switched off.  This probe writes the same
// tiles with nothing else going on: every CTA owns 128 rows of an [M][N] fp32 matrix and rewrites `tiles` column tiles
// of it, (a) with coalesced st.global.v4 (eight lanes per 128-byte line, the epilogue's pattern), (b) the same with an
// L2 evict-first / no-allocate hint, (c) with TMA bulk stores of 16 KB boxes from shared memory.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/store_probe tools/store_probe.cu -lcuda && tools/store_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: st.global.v4   1: st.global.cs.v4 (streaming)   2: st.global.L1::no_allocate.v4
template <int MODE>
__global__ void __launch_bounds__(256, 1) stg_kernel(float* out, int N, int tiles, int reps) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int row0 = blockIdx.x * 128 + (warp % 4) * 32;   // 8 warps: two per 32-row quarter, alternate 128-byte chunks
  const int half = warp / 4;
  const uint32_t v = threadIdx.x;
  for (int r = 0; r < reps; r++)
    for (int t = 0; t < tiles; t++)
      for (int c = half * 32; c < 256; c += 64)
        for (int hh = 0; hh < 2; hh++)
          for (int i = 0; i < 4; i++) {
            const int R = (lane >> 3) + 4 * i + 16 * hh, P = lane & 7;
            float* p = out + (size_t)(row0 + R) * N + (t * 256 + c) % N + P * 4;
            if (MODE == 0) asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
            else if (MODE == 1) asm volatile("st.global.cs.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
            else asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
          }
}

__global__ void __launch_bounds__(128, 1) tma_store_kernel(const __grid_constant__ CUtensorMap to, int N, int tiles, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0) {
    for (int r = 0; r < reps; r++)
      for (int t = 0; t < tiles; t++)
        for (int c = 0; c < 256; c += 32) {   // 8 boxes of 128 rows x 128 bytes per tile
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&to),
                       "r"(smem_u32(smem + ((c / 32) & 1) * 16384)), "r"((t * 256 + c) % N), "r"((int)blockIdx.x * 128) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

typedef CUresult (*PFN_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  PFN_encode enc = (PFN_encode)fn;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount, N = 1024, M = sms * 128;
  float* out;
  CK(cudaMalloc(&out, (size_t)M * N * 4));   // 148 x 128 rows x 4 KB = 77 MB: L2-resident
  CUtensorMap to;
  const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
  const cuuint64_t strides[1] = {(cuuint64_t)N * 4};
  const cuuint32_t box[2] = {32, 128}, estr[2] = {1, 1};
  if (enc(&to, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("encode failed\n");
    return 1;
  }
  CK(cudaFuncSetAttribute(tma_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  printf("# %s: every CTA rewrites 128 rows x 256 fp32 column tiles of a [%d][%d] matrix (%.0f MB)\n", prop.name, M, N, (double)M * N * 4 / 1e6);
  printf("# %-28s %-6s %-8s | %-9s %-10s %-9s %-12s\n", "how", "ctas", "tiles", "us", "GB/s", "GB/s/SM", "us per tile");
  const char* names[4] = {"st.global.v4 (8 warps)", "st.global.cs.v4", "st.global.L1::no_allocate.v4", "TMA store, 16 KB boxes"};
  for (int ctas : {16, 64, sms})
    for (int mode = 0; mode < 4; mode++) {
      const int tiles = 4, reps = 16;
      float best = 1e30f;
      for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0));
        if (mode == 0) stg_kernel<0><<<ctas, 256>>>(out, N, tiles, reps);
        else if (mode == 1) stg_kernel<1><<<ctas, 256>>>(out, N, tiles, reps);
        else if (mode == 2) stg_kernel<2><<<ctas, 256>>>(out, N, tiles, reps);
        else tma_store_kernel<<<ctas, 128, 34 * 1024>>>(to, N, tiles, reps);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
      }
      const double bytes = (double)ctas * tiles * reps * 128 * 256 * 4;
      printf("  %-28s %-6d %-8d | %-9.1f %-10.0f %-9.1f %-12.2f\n", names[mode], ctas, tiles * reps, best * 1e3, bytes / (best * 1e-3) / 1e9,
             bytes / (best * 1e-3) / 1e9 / ctas, best * 1e3 / (tiles * reps));
    }
  return 0;
}
