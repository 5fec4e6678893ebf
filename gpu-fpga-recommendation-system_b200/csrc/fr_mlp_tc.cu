// fr_mlp_tc.cu -- the MLP as tcgen05 / TMEM GEMMs (FR_PREC_TF32), sm_100a only.
//
// Replaces the four cublasLtMatmul calls of cuda_server.c:468-491 (FP32 SGEMM chain
//   R1 = W1.X, R2 = W2.R1, R3 = W3.R2, out = W4.R3, alpha=1 beta=0)
// with three hand-written kernels launches per batch:
//   layer 1, 2 : Y[B][N] = rna_tf32(act(A[B][K] . Wt[N][K]^T + bias))     (EPI_STORE)
//   layer 3+4  : score[b] = sig(b4 + sum_n w4[n] * act(A[b] . Wt3[n] + b3[n]))  (EPI_DOT)
// i.e. bias + ReLU live in the TMEM epilogue and the 256->1 output layer plus the
// sigmoid are folded into layer 3's epilogue, so R3 never exists in memory.
//
// One CTA computes one 128 x BLOCK_N output tile:
//   warp 0      TMA producer  : cp.async.bulk.tensor 2D loads of the A (128 x 32 fp32)
//                               and B (BLOCK_N x 32 fp32) K-slices into a STAGES-deep
//                               128B-swizzled smem ring, completion on mbarriers
//   warp 1      MMA issuer    : allocates TMEM, one elected lane issues
//                               tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BLOCK_N, K=8)
//                               x4 per K-slice, tcgen05.commit frees the smem slot
//   warps 2..5  epilogue      : tcgen05.ld 32x32b (one TMEM lane = one output row per
//                               thread), bias/ReLU/(dot), stores
// Operands are K-major in smem for both A and B: activations are row-major [B][K]
// as the reference has them (cuda_server.c:216), weights are transposed once at
// load time ([in][out] -> [out][in], rounded to TF32).  FP32 accumulate in TMEM.
// K tails (880 = 27.5 slices) and M tails rely on TMA zero fill of out-of-bounds
// box elements; stores are masked by the row count.
#include <cuda.h>
#include <string.h>

#include "fr_common.h"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;            // fp32 elements = 128 bytes = one swizzle atom
constexpr int UMMA_K = 8;              // tf32: 32 bytes of K per instruction
constexpr int kThreads = 192;
constexpr int kEpiWarp0 = 2;

enum { EPI_STORE = 0, EPI_DOT = 1 };

// ---- PTX wrappers -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Where a kernel that gave up waiting says so: pinned + mapped host memory, written once, then trap.
// {magic, code, blockIdx.x, warp, parity, gridDim.x, k-slices done, tile}
__device__ int* g_watch = nullptr;
constexpr long long kWatchCycles = 6000000000ll;   // ~3 s at 1.97 GHz: no legitimate wait comes close

__device__ __noinline__ void watchdog_fire(int code, uint32_t parity, uint32_t a, uint32_t b) {
  int* w = g_watch;
  if (w && atomicCAS(w, 0, 0x57415443) == 0) {
    w[1] = code; w[2] = blockIdx.x; w[3] = threadIdx.x / 32; w[4] = (int)parity; w[5] = gridDim.x; w[6] = (int)a; w[7] = (int)b;
    __threadfence_system();
  }
  __trap();
}

// Bounded spin: a wait that outlives kWatchCycles is a protocol bug (or a peer CTA that died); the
// kernel records which wait it was and traps, so the host sees an error instead of a hung stream.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int code = 0, uint32_t a = 0, uint32_t b = 0) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spins = 0; !done; spins++) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && (spins & 1023) == 1023) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kWatchCycles) watchdog_fire(code, parity, a, b);
    }
  }
}
// Same, acquiring at cluster scope: the data guarded by the barrier was written by the peer CTA's threads.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int code = 0, uint32_t a = 0, uint32_t b = 0) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spins = 0; !done; spins++) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && (spins & 1023) == 1023) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kWatchCycles) watchdog_fire(code, parity, a, b);
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* smem, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem] . B[smem]^T, tf32 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}


// fp32 -> tf32, round to nearest with ties away from zero (what cvt.rna.tf32.f32 computes for finite
// values): add half an ulp of the 10-bit mantissa to the magnitude and clear the 13 low bits.  Two
// full-rate integer instructions; the cvt runs at 16 / clk / SM, which made it the largest single cost
// of the store epilogue (128 x 512 conversions per tile).
__device__ __forceinline__ float round_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// Explicit shared-space accesses.  The dynamic smem base is aligned through an integer cast, after which
// the compiler no longer knows the pointers are shared: it emitted generic LD.E / ST.E and, unable to
// prove the bias loads independent of the staging stores, serialised them (load, wait, 4 FADDs, store,
// next load ...: ~720 cycles per 32-column chunk).  These keep the accesses LDS / STS, and the callers
// issue all loads of a chunk before its stores.
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// One epilogue chunk of the storing layers: 32 accumulator columns of this thread's row + bias -> ReLU ->
// TF32 rounding -> this warp's 128B-swizzled 4 KB staging buffer (16-byte piece j of row `lane` lives at
// piece j ^ (lane & 7), the layout the TMA store box expects).  bias_addr / buf_addr: shared-space addresses.
__device__ __forceinline__ void epi_chunk_to_smem(const uint32_t (&r)[32], uint32_t bias_addr, uint32_t buf_addr,
                                                  int lane, int relu) {
  float4 bv[8];
#pragma unroll
  for (int j = 0; j < 8; j++) bv[j] = lds128(bias_addr + j * 16);
  const uint32_t row = buf_addr + lane * 128;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    float4 o;
    o.x = __uint_as_float(r[4 * j + 0]) + bv[j].x;
    o.y = __uint_as_float(r[4 * j + 1]) + bv[j].y;
    o.z = __uint_as_float(r[4 * j + 2]) + bv[j].z;
    o.w = __uint_as_float(r[4 * j + 3]) + bv[j].w;
    if (relu) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w);
    sts128(row + ((j ^ (lane & 7)) << 4), o);
  }
}

// The same for fp16 activations (ELT = 2 kernels): 64 accumulator columns (two TMEM loads) + bias -> ReLU -> fp16
// (round to nearest even) -> one 128-byte row of the staging buffer, 8 pieces of 8 halves.
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void epi_chunk64_to_smem_f16(const uint32_t (&r0)[32], const uint32_t (&r1)[32], uint32_t bias_addr,
                                                        uint32_t buf_addr, int lane, int relu) {
  const uint32_t row = buf_addr + lane * 128;
#pragma unroll
  for (int j = 0; j < 8; j++) {   // piece j = columns 8j .. 8j+7
    const float4 b0 = lds128(bias_addr + j * 32), b1 = lds128(bias_addr + j * 32 + 16);
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int cidx = 8 * j + q;
      v[q] = __uint_as_float(cidx < 32 ? r0[cidx] : r1[cidx - 32]);
    }
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    if (relu) {
#pragma unroll
      for (int q = 0; q < 8; q++) v[q] = fmaxf(v[q], 0.f);
    }
    uint4 o;
    o.x = pack_f16x2(v[0], v[1]); o.y = pack_f16x2(v[2], v[3]); o.z = pack_f16x2(v[4], v[5]); o.w = pack_f16x2(v[6], v[7]);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + ((j ^ (lane & 7)) << 4)), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w)
                 : "memory");
  }
}

// One epilogue chunk of layer 3: dot += sum_j act(acc[j] + bias[j]) * w4[j] over 32 columns, in column order.
__device__ __forceinline__ float epi_chunk_dot(const uint32_t (&r)[32], uint32_t bias_addr, uint32_t w4_addr, int relu,
                                               float dot) {
  float4 bv[8], wv[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    bv[j] = lds128(bias_addr + j * 16);
    wv[j] = lds128(w4_addr + j * 16);
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    float v0 = __uint_as_float(r[4 * j + 0]) + bv[j].x, v1 = __uint_as_float(r[4 * j + 1]) + bv[j].y;
    float v2 = __uint_as_float(r[4 * j + 2]) + bv[j].z, v3 = __uint_as_float(r[4 * j + 3]) + bv[j].w;
    if (relu) {
      v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
    }
    dot = fmaf(v0, wv[j].x, dot);
    dot = fmaf(v1, wv[j].y, dot);
    dot = fmaf(v2, wv[j].z, dot);
    dot = fmaf(v3, wv[j].w, dot);
  }
  return dot;
}

// K-major, 128B-swizzled operand tile: rows are 128 bytes, 8-row groups are 1024 bytes
// apart (SBO), LBO unused for swizzled K-major; descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                              // leading byte offset (ignored), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                              // version, bits [46,48)
  d |= (uint64_t)2 << 61;                              // layout type SWIZZLE_128B, bits [61,64)
  return d;
}

// ---- cluster helpers (2-CTA pairs) --------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
// 2-CTA TMA load: data lands in THIS CTA's smem, completion bytes are posted on the mbarrier at
// cluster address `bar_cluster` (the pair leader's), which cta_group::2 permits.
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_cluster, void* smem, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
// The same, delivered to this CTA and to every CTA of `mask` at the same smem offset (the pairs of a 4-CTA
// cluster that multiply the same weight slice).  `bar_own` is THIS CTA's barrier address with the pair bit
// (bit 24 of a shared-window address) cleared: the byte count is posted on the barrier at that offset in the
// LEADER of each destination CTA's pair.
__device__ __forceinline__ void tma_load_2d_pair_mcast(const CUtensorMap* map, uint32_t bar_own, void* smem, int c0, int c1,
                                                       uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], "
      "[%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem)),
      "l"(map), "r"(bar_own), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the pair bit of a shared-window address: "the leader's"
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D (256 x N, rows 0..127 in the leader's TMEM, 128..255 in the peer's) (+)= A . B^T with A and B
// halves read from BOTH CTAs' smem at the same offsets; issued by one thread of the leader CTA.
__device__ __forceinline__ void umma_tf32_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same MMA on fp16 operands (kind::f16, K = 16 per instruction = the same 32 bytes of a swizzled row).
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once all prior MMAs of the pair retired) on the barrier at this offset in BOTH CTAs.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// The same, arriving in every CTA of `mask` (cluster ranks).
__device__ __forceinline__ void umma_commit_mask(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// ---- TMA store / programmatic dependent launch ----------------------------------
// smem (128B-swizzled box) -> global; completion tracked by this thread's bulk async-groups.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's store groups may still be READING their smem source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// PDL: block until the preceding kernel of the stream has completed and its writes are visible
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// PDL: the next kernel of the stream may start its prologue now
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}

// 16-byte asynchronous copy global -> shared through the LSU (LDGSTS), bypassing L1; src_bytes = 0 writes zeros
// (rows past the batch, pieces past K) without touching `src`.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
constexpr int kLoaderWarp0 = 6;      // A_LSU kernels: warps 6..9 load the A operand with cp.async
constexpr int kLoaderWarps = 4;
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int kMaxN = 2048;          // widest layer whose bias vector is staged in smem (large model H1)
constexpr int kStoreBoxRows = 32;    // TMA store box of the fused / chain kernels: one epilogue warp's 32 rows x 32 fp32 columns
constexpr int kStoreBufBytes = kStoreBoxRows * BLOCK_K * 4;   // 4 KB, 128B-swizzled
// tc_linear_kernel stores a whole CTA's 128 rows x 128 bytes per TMA instruction.  The TMA unit of an SM retires one
// instruction per ~340 cycles whatever its box size up to 32 KB (tools/tma_probe.cu: 16 KB boxes land at 48 B/clk/SM,
// 32 KB boxes at 72-96), so the 32 4-KB stores a 256-wide tile used to issue cost the unit as much as 32 K slices of
// loads -- more than the tile's MMAs.  Four epilogue warps now fill ONE 16 KB staging buffer (named barrier) and one
// thread stores it: 8 stores per tile.
constexpr int kCtaStoreBytes = BLOCK_M * 128;   // 16 KB
constexpr int kCtaStoreBufs = 3;                // store i-2 has released its buffer before chunk i+1 is staged
constexpr int kEpiBarrier = 1;                  // named barrier of the 4 epilogue warps

// CTAS = 1: one CTA owns 128 x BLOCK_N tiles.  CTAS = 2: a CTA pair (cluster of 2 along M) owns
// 256 x BLOCK_N tiles; each CTA stages its own 128 rows of A and its own BLOCK_N/2 rows of B, so per
// SM the smem fill per MMA cycle halves for B -- the point of cta_group::2 for 4-byte operands.
template <int BLOCK_N, int STAGES, int CTAS>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 4;
  static constexpr int kBRows = BLOCK_N / CTAS;
  static constexpr int kBBytes = kBRows * BLOCK_K * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStoreOff = STAGES * kStageBytes;          // kCtaStoreBufs staging buffers of 128 rows x 128 B
  static constexpr int kAuxOff = kStoreOff + kCtaStoreBufs * kCtaStoreBytes;   // bias[kMaxN], w4[256]
  static constexpr int kBarOff = kAuxOff + (kMaxN + 256) * 4;
  static constexpr int kNumBars = 3 * STAGES + 4;                 // full, empty, tmem_full[2], tmem_empty[2], a_full
  static constexpr int kTotal = kBarOff + kNumBars * 8 + 16;
  static constexpr int kDyn = kTotal + 1024;                      // slack for manual 1024-B alignment
  // BLOCK_N <= 256: two accumulator stages (epilogue of tile i under the MMAs of tile i+1).
  // BLOCK_N = 512: ONE stage filling all 512 TMEM columns, two N = 256 MMAs per K step sharing the A
  // slice -- 48 KB of operands per 2 x 4 MMAs instead of 32 KB per 4 (33 against 45 B/clk/SM at the
  // measured TF32 rate; L2 -> SM delivers ~40), at the price of an exposed epilogue.
  static constexpr int kAcc = BLOCK_N <= 256 ? 2 : 1;
  static constexpr int kTmemCols = kAcc * BLOCK_N;
  static constexpr int kMmaN = BLOCK_N <= 256 ? BLOCK_N : 256;    // columns per tcgen05.mma
  static constexpr int kNSub = BLOCK_N / kMmaN;                   // MMAs per K step
  static constexpr int kSubRows = kMmaN / CTAS;                   // rows of Wt one CTA stages per MMA
  static constexpr int kSubBytes = kSubRows * BLOCK_K * 4;
};

struct TcParams {
  const float* a;      // [M][K] activations (A_LSU kernels read them with cp.async; the others through tmap_a)
  const float* bias;   // [N] or null
  float* out;          // EPI_DOT: scores [M]  (tc_linear_kernel's EPI_STORE writes through tmap_out)
  void* out_act;       // tc_pair_kernel EPI_STORE: activations [M][N] (fp32 tf32-rounded, or fp16), row pitch N
  const float* w4;     // EPI_DOT: output-layer weights [N]
  const float* b4;     // EPI_DOT: output-layer bias [1] or null
  int M, N, K;
  int relu, sigmoid;
  int pdl;             // issue griddepcontrol.wait / launch_dependents (no-ops unless a launch in the chain
                       // carries the programmatic-serialization attribute)
  int latency;         // host-side hint: small batch, one tile per cluster (no ganging)
  // table-sharded step (layer 1 only): the A operand is this rank's exchange buffer, which the peer ranks fill over
  // NVLink.  The TMA producer polls wait_flags[0..wait_n) until every rank has published step *wait_step, right
  // before its first A load -- TMEM allocation, barrier set-up and bias staging of this launch, and every other
  // worker's kernels, run while the peers' rows are still in flight.  wait_flags == nullptr: no wait.
  const int* wait_flags;
  const int* wait_step;
  int wait_n;
  int* wait_err;
  // FR_EXPERIMENTS builds, FR_TC_PROF=1: where the three pipelines of the first 8 CTAs spend their cycles
  // (tools/tc_prof.py): prof[cta][16] = {producer total, waiting for a free slot, slices | MMA issuer total, waiting
  // for operands, waiting for a free accumulator, slices | epilogue warp total, waiting for an accumulator, waiting
  // for a free staging buffer + barrier, tiles}
  long long* prof;
  int dbg_nostore;     // FR_EXPERIMENTS, FR_TC_NOSTORE=1: the storing epilogues skip their global stores (timing experiments only)
};
#ifdef FR_EXPERIMENTS
#define FR_PROF_T0(v) const long long v = clock64()
#define FR_PROF_ADD(acc, v) acc += clock64() - (v)
#else
#define FR_PROF_T0(v) do { } while (0)
#define FR_PROF_ADD(acc, v) do { } while (0)
#endif

// Poll the peers' step flags (system-scope acquire: the rows were written by other GPUs); a peer that has not
// published after ~2 s is given up on -- *err is set (fr_sync reports it) and the kernel runs on what is there.
__device__ __forceinline__ void wait_for_peers(const TcParams& p) {
  const int want = *reinterpret_cast<const volatile int*>(p.wait_step);
  const long long t0 = clock64();
  for (int r = 0; r < p.wait_n; r++) {
    int v;
    for (;;) {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p.wait_flags + r) : "memory");
      if (v >= want) break;
      if (clock64() - t0 > 4000000000ll) {
        *reinterpret_cast<volatile int*>(p.wait_err) = 1;
        break;
      }
    }
  }
  asm volatile("fence.proxy.async;" ::: "memory");   // the TMA loads that follow read through the async proxy
}

// kind::tf32 instruction descriptor: D=f32, A=B=tf32, both K-major, M = 128*CTAS, N = n.
__host__ __device__ constexpr uint32_t make_idesc_tf32_m(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// kind::f16 instruction descriptor: D = f32, A = B = f16 (format 0), both K-major.
__host__ __device__ constexpr uint32_t make_idesc_f16_m(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Persistent: cluster c (one CTA or a CTA pair) walks tiles c, c + n_clusters, ... of the
// (M / (128*CTAS)) x (N / BLOCK_N) tile grid, n fastest.  Three pipelines run concurrently:
//   smem ring   full/empty        TMA producer  <-> MMA issuer
//   TMEM        tmem_full/empty   MMA issuer    <-> epilogue   (two accumulator stages, so the
//                                 epilogue of tile i overlaps the MMAs of tile i+1)
//   store bufs  bulk async-groups epilogue warp <-> TMA store  (two 4 KB buffers per warp)
// PAIRS = 2 (CTAS = 2 only): a cluster of FOUR CTAs, two MMA pairs working on adjacent 256-row tiles of the
// same N tile in lockstep.  Both pairs multiply the same weight slice, so every CTA loads only HALF of its
// share of it and TMA-multicasts that half to the CTA of the same parity in the other pair: per K slice a
// CTA issues 16 + 16 KB (512-wide) or 16 + 8 KB (256-wide) of loads instead of 16 + 32 / 16 + 16, and the L2
// serves every weight byte once per cluster.  The TMA unit of an SM sustains ~33 B/clk here while the MMAs of
// a 512-wide tile want 49, so issued bytes per SM are what bounds the main loop.  A smem slot may be refilled
// only when BOTH pairs have consumed it (the neighbour's load writes into it): the empty barrier counts two
// multicast commits.
//
// A_LSU: the A operand (activations) does not go through the TMA unit.  An SM takes in ~33 B/clk through TMA
// whatever the box shape, stage count, L2 promotion or multicast (all measured), and a 512-wide tile wants 49;
// but 16-byte LDGSTS copies through the LSU run beside it (a probe streaming 38 GB/s per SM next to the TMA
// feed slowed the main loop by 8 %).  So four more warps (6..9) copy the 128 x 32 A slice with cp.async.cg into
// the same 128B-swizzled slot layout TMA would have produced (16-byte piece j of row r at piece j ^ (r & 7));
// every loader thread posts cp.async.mbarrier.arrive.noinc on the slot's a_full barrier, so the arrival fires
// when its copies have landed and the loaders only ever wait for free slots (all slots can be in flight).
// The leader's MMA issuer waits for its own CTA's a_full; the peer CTA's otherwise idle warp 1 waits for the
// peer's and relays it as one remote arrival on the leader's full barrier.  TMA then carries the weights
// only: 32 KB per K slice of a 512-wide tile (0.49 us at the measured rate, the MMAs take 0.5).
//
// ELT = 2 (experimental, FR_TC_F16=1): fp16 operands and activations, kind::f16.  The shared-memory picture is the
// same in BYTES (128-byte swizzled rows, 32 bytes of K per MMA), a K slice is 64 elements instead of 32; fp16
// carries the same 11-bit significand TF32 keeps, in half the bytes, at twice the MMA rate -- at the price of
// fp16's range.
template <int BLOCK_N, int STAGES, int EPI, int CTAS, int PAIRS, bool A_LSU, int ELT>
__global__ void __launch_bounds__(kThreads + (A_LSU ? kLoaderWarps * 32 : 0), 1)
tc_linear_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_out, const TcParams p) {
  static_assert(!A_LSU || (CTAS == 2 && PAIRS == 1), "the cp.async A loader is built for plain CTA pairs");
  static_assert(ELT == 4 || (ELT == 2 && CTAS == 2 && PAIRS == 1 && !A_LSU), "fp16 operands: plain CTA pairs only");
  constexpr int BK = 128 / ELT;   // elements of K per slice (one 128-byte swizzle row)
  using L = SmemLayout<BLOCK_N, STAGES, CTAS>;
  extern __shared__ uint8_t smem_raw[];
  // the dynamic-smem base offset is identical in both CTAs of a pair, so is the aligned layout
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_bias = reinterpret_cast<float*>(smem + L::kAuxOff);
  float* s_w4 = s_bias + kMaxN;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2], the leader's are the ones used
  uint64_t* a_full_bar = tmem_empty_bar + 2;        // [STAGES], A_LSU: this CTA's A slice has landed (cp.async arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_full_bar + STAGES);

#ifdef FR_EXPERIMENTS
  // launch timeline of CTA 0 (FR_TC_PROF): prof[112 + 3 * (launch % 4) + {0, 1, 2}] = %globaltimer at kernel entry, after
  // the prologue, at exit; prof[127] counts launches
  long long* pf_tl = nullptr;
  if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) {
    const long long n = atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 127), 1ull);
    pf_tl = p.prof + 112 + 3 * (n & 3);
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    pf_tl[0] = t;
  }
#endif
  static_assert(PAIRS == 1 || CTAS == 2, "multicast clusters are made of CTA pairs");
  constexpr int CS = CTAS * PAIRS;   // CTAs per cluster
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t crank = (CTAS == 2) ? cluster_ctarank() : 0u;   // rank in the cluster
  const uint32_t rank = crank & 1u, pair = crank >> 1;            // rank in the MMA pair, pair in the cluster
  const bool leader = (rank == 0);
  const int num_kb = (p.K + BK - 1) / BK;
  const int n_tiles_n = p.N / BLOCK_N;
  const int n_tiles = ((p.M + BLOCK_M * CS - 1) / (BLOCK_M * CS)) * n_tiles_n;   // a tile = (BLOCK_M * CS) rows x BLOCK_N
  const int cluster_id = blockIdx.x / CS, n_clusters = gridDim.x / CS;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    if (EPI == EPI_STORE) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_out) : "memory");
    for (int s = 0; s < STAGES; s++) {
      // the leader's producer arrives once, bytes of both CTAs are expected; A_LSU: + the peer CTA's relay
      mbar_init(&full_bar[s], 1 + (A_LSU ? 1 : 0));
      mbar_init(&a_full_bar[s], kLoaderWarps * 32);   // A_LSU: one cp.async-completion arrival per loader thread
      mbar_init(&empty_bar[s], PAIRS);  // one (multicast) tcgen05.commit per use from every pair that reads the slot
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(&tmem_full_bar[a], 1);          // one (multicast) tcgen05.commit per tile
      mbar_init(&tmem_empty_bar[a], 4 * CTAS);  // every epilogue warp of every CTA of the pair
    }
    fence_barrier_init();
  } else if (warp == 1) {
    if (CTAS == 2) tmem_alloc_pair(tmem_ptr, L::kTmemCols);
    else tmem_alloc(tmem_ptr, L::kTmemCols);
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
    // weights, not produced by the preceding kernel: staged before the grid dependency resolves
    const int et = threadIdx.x - kEpiWarp0 * 32;  // 0..127
    for (int i = et; i < p.N; i += 128) s_bias[i] = p.bias ? p.bias[i] : 0.f;
    if (EPI == EPI_DOT)
      for (int i = et; i < BLOCK_N; i += 128) s_w4[i] = p.w4[i];
  }
  tc_fence_before();
  __syncthreads();                     // CTA-level order of the barrier inits, the staged bias and tcgen05.alloc's write of
  if (CTAS == 2) cluster_sync_all();   // tmem_ptr; peer barriers must exist before any remote complete_tx / commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (p.pdl) griddep_wait();           // activations of the previous kernel are complete and visible from here
#ifdef FR_EXPERIMENTS
  if (pf_tl) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    pf_tl[1] = t;
  }
#endif

  if (warp == 0) {
    // ===== TMA producer (both CTAs of a pair) =====
    if (lane == 0) {
      if (p.wait_flags) wait_for_peers(p);
      long long pf_wait = 0;
      FR_PROF_T0(pf_start);
      uint32_t kc = 0;   // k-slices issued so far (ring position runs on across tiles)
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
        const int m0 = (((tile / n_tiles_n) * PAIRS + (int)pair) * CTAS + (int)rank) * BLOCK_M;
        const int nb0 = (tile % n_tiles_n) * BLOCK_N + (int)rank * L::kSubRows;
        for (int kb = 0; kb < num_kb; kb++, kc++) {
          const int s = kc % STAGES;
          const uint32_t ph = (kc / STAGES) & 1;
          FR_PROF_T0(pf_t);
          mbar_wait(&empty_bar[s], ph ^ 1, 1, kc, tile);
          FR_PROF_ADD(pf_wait, pf_t);
          uint8_t* a_dst = smem + s * L::kStageBytes;
          uint8_t* b_dst = a_dst + L::kABytes;
          if (PAIRS == 2) {
            // this CTA's rows of Wt for the K step, in smem order: sub-tile h, row r <-> Wt row nb0 + h * kMmaN + r.
            // It loads part `pair` of them and multicasts it to the CTA of its parity in both pairs.
            constexpr int kPartRows = L::kNSub * L::kSubRows / PAIRS;
            const int r0 = (int)pair * kPartRows;
            const uint32_t bar = smem_u32(&full_bar[s]) & kPeerBitMask;   // "my pair leader's", in every destination
            if (leader) mbar_expect_tx(&full_bar[s], 2 * L::kStageBytes);
            tma_load_2d_pair(&tmap_a, bar, a_dst, kb * BK, m0);
            tma_load_2d_pair_mcast(&tmap_b, bar, b_dst + r0 * (BLOCK_K * 4), kb * BK,
                                   nb0 + (r0 / L::kSubRows) * L::kMmaN + r0 % L::kSubRows, (uint16_t)(0x5u << rank));
          } else if (CTAS == 2) {
            const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);   // leader's barrier
            if (leader) mbar_expect_tx(&full_bar[s], 2 * (L::kStageBytes - (A_LSU ? L::kABytes : 0)));
            if (!A_LSU) tma_load_2d_pair(&tmap_a, bar, a_dst, kb * BK, m0);
#pragma unroll
            for (int h = 0; h < L::kNSub; h++)   // this CTA's rows of Wt for MMA h of the K step
              tma_load_2d_pair(&tmap_b, bar, b_dst + h * L::kSubBytes, kb * BK, nb0 + h * L::kMmaN);
          } else {
            mbar_expect_tx(&full_bar[s], L::kStageBytes);
            tma_load_2d(&tmap_a, &full_bar[s], a_dst, kb * BK, m0);
#pragma unroll
            for (int h = 0; h < L::kNSub; h++)
              tma_load_2d(&tmap_b, &full_bar[s], b_dst + h * L::kSubBytes, kb * BK, nb0 + h * L::kMmaN);
          }
        }
      }
      if (p.pdl) griddep_launch();
#ifdef FR_EXPERIMENTS
      if (p.prof && blockIdx.x < 8) {
        long long* o = p.prof + blockIdx.x * 16;
        o[0] = clock64() - pf_start; o[1] = pf_wait; o[2] = kc;
      }
#endif
      (void)pf_wait;
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only for a pair) =====
    if (A_LSU && !leader) {
      // the peer CTA's warp 1 has no MMAs to issue: it relays "my A slice has landed" to the leader
      const uint32_t full_remote = mapa_u32(smem_u32(&full_bar[0]), 0);
      uint32_t kc = 0;
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters)
        for (int kb = 0; kb < num_kb; kb++, kc++) {
          const int s = kc % STAGES;
          mbar_wait(&a_full_bar[s], (kc / STAGES) & 1, 7, kc, tile);
          if (lane == 0) mbar_arrive_cluster(full_remote + s * 8);
          __syncwarp();
        }
    }
    if (leader) {
      constexpr uint32_t idesc = ELT == 2 ? make_idesc_f16_m(BLOCK_M * CTAS, L::kMmaN) : make_idesc_tf32_m(BLOCK_M * CTAS, L::kMmaN);
      uint32_t kc = 0, it = 0;
      long long pf_full = 0, pf_tmem = 0;
      FR_PROF_T0(pf_start);
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, it++) {
        const uint32_t as = it % L::kAcc;
        FR_PROF_T0(pf_t0);
        mbar_wait(&tmem_empty_bar[as], ((it / L::kAcc) & 1) ^ 1, 2, kc, tile);   // the epilogue has drained this accumulator stage
        FR_PROF_ADD(pf_tmem, pf_t0);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < num_kb; kb++, kc++) {
          const int s = kc % STAGES;
          const uint32_t ph = (kc / STAGES) & 1;
          if (A_LSU) {
            mbar_wait_cluster(&full_bar[s], ph, 3, kc, tile);   // weights of both CTAs + the peer's A slice (relayed)
            mbar_wait(&a_full_bar[s], ph, 7, kc, tile);         // this CTA's A slice
            // (no fence.proxy.async here: issued by the thread that has MMAs in flight it drains them, one K slice
            // at a time -- 0.9 us per slice; cp.async completion -> mbarrier -> tcgen05.mma is ordered as it is)
          } else {
            FR_PROF_T0(pf_t1);
            mbar_wait(&full_bar[s], ph, 3, kc, tile);
            FR_PROF_ADD(pf_full, pf_t1);
          }
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
            const uint64_t a_desc = make_smem_desc(a_addr);
            const uint64_t b_desc = make_smem_desc(a_addr + L::kABytes);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; k++) {
#pragma unroll
              for (int h = 0; h < L::kNSub; h++) {
                // advance 32 bytes of K inside the 128-byte swizzle atom: +2 in the (addr >> 4) field;
                // MMA h of the step reads Wt sub-tile h and accumulates into TMEM columns h * 256
                const uint64_t bd = b_desc + (uint64_t)(k * 2) + (uint64_t)(h * (L::kSubBytes >> 4));
                if (ELT == 2) umma_f16_pair(d_tmem + h * L::kMmaN, a_desc + (uint64_t)(k * 2), bd, idesc, (kb | k) != 0);
                else if (CTAS == 2) umma_tf32_pair(d_tmem + h * L::kMmaN, a_desc + (uint64_t)(k * 2), bd, idesc, (kb | k) != 0);
                else umma_tf32(d_tmem + h * L::kMmaN, a_desc + (uint64_t)(k * 2), bd, idesc, (kb | k) != 0);
              }
            }
            if (PAIRS == 2) {
              umma_commit_mask(&empty_bar[s], 0xF);                                        // one of the two releases, in all four CTAs
              if (kb == num_kb - 1) umma_commit_mask(&tmem_full_bar[as], (uint16_t)(0x3u << (2 * pair)));
            } else if (CTAS == 2) {
              umma_commit_pair(&empty_bar[s]);                            // frees the slot in both CTAs
              if (kb == num_kb - 1) umma_commit_pair(&tmem_full_bar[as]);  // accumulators complete in both CTAs
            } else {
              umma_commit(&empty_bar[s]);
              if (kb == num_kb - 1) umma_commit(&tmem_full_bar[as]);
            }
          }
          __syncwarp();
        }
      }
#ifdef FR_EXPERIMENTS
      if (p.prof && blockIdx.x < 8 && lane == 0) {
        long long* o = p.prof + blockIdx.x * 16 + 4;
        o[0] = clock64() - pf_start; o[1] = pf_full; o[2] = pf_tmem; o[3] = kc;
      }
#endif
      (void)pf_full; (void)pf_tmem;
    }
  } else if (A_LSU && warp >= kLoaderWarp0) {
    // ===== A loaders: thread t copies piece j = t % 8 of rows t / 8 + 16 i (i < 8) of every K slice =====
    const int t = threadIdx.x - kLoaderWarp0 * 32;   // 0..127
    const int j = t & 7, r0 = t >> 3;
    const float* A = p.a;
    uint32_t kc = 0;      // K slices issued
    for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
      const int m0 = ((tile / n_tiles_n) * CTAS + (int)rank) * BLOCK_M;
      for (int kb = 0; kb < num_kb; kb++, kc++) {
        const int s = kc % STAGES;
        mbar_wait(&empty_bar[s], ((kc / STAGES) & 1) ^ 1, 5, kc, tile);
        const uint32_t slot = smem_u32(smem + s * L::kStageBytes);
        const int col = kb * BK + j * 4;
        const bool col_ok = col < p.K;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int r = r0 + 16 * i, row = m0 + r;
          const bool ok = col_ok && row < p.M;
          cp_async16(slot + r * 128 + ((j ^ (r & 7)) << 4), ok ? A + (size_t)row * p.K + col : A, ok ? 16u : 0u);
        }
        // fires when this thread's copies above have landed
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&a_full_bar[s])) : "memory");
      }
    }
    cp_async_wait_all();   // nothing of this thread is still in flight when the CTA retires
  } else {
    // ===== epilogue warps: TMEM lane quarter = warp % 4 =====
    const int q = warp % 4;
    uint8_t* store_base = smem + L::kStoreOff;
    const bool issuer = (warp == kEpiWarp0 && lane == 0);   // the one thread that issues (and tracks) this CTA's TMA stores
    const uint32_t empty_remote = mapa_u32(smem_u32(&tmem_empty_bar[0]), crank & ~1u);   // my pair leader's tmem_empty_bar[0]
    uint32_t it = 0, sc = 0;   // tiles done, store chunks issued
    long long pf_acc = 0, pf_buf = 0;
    FR_PROF_T0(pf_start);
    for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, it++) {
      const uint32_t as = it % L::kAcc;
      const int m0 = (((tile / n_tiles_n) * PAIRS + (int)pair) * CTAS + (int)rank) * BLOCK_M;   // this CTA's first row
      const int row0 = m0 + q * 32;
      const int n0 = (tile % n_tiles_n) * BLOCK_N;
      FR_PROF_T0(pf_t0);
      mbar_wait(&tmem_full_bar[as], (it / L::kAcc) & 1, 4, it, tile);
      FR_PROF_ADD(pf_acc, pf_t0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N;
      float dot = 0.f;
      if (EPI == EPI_STORE) {
        // one chunk = 128 bytes of every row of the CTA (32 fp32 / 64 fp16 columns): each warp stages its 32 rows in
        // its quarter of the 16 KB buffer, one thread stores the buffer
        constexpr int kChunkCols = 128 / ELT;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += kChunkCols, sc++) {
          uint8_t* buf = store_base + (sc % kCtaStoreBufs) * kCtaStoreBytes;
          const uint32_t mine = smem_u32(buf) + q * (32 * 128);
          if (ELT == 2) {
            uint32_t r0[32], r1[32];
            tmem_ld32(taddr + c, r0);
            tmem_ld32(taddr + c + 32, r1);
            epi_chunk64_to_smem_f16(r0, r1, smem_u32(s_bias + n0 + c), mine, lane, p.relu);
          } else {
            uint32_t r[32];
            tmem_ld32(taddr + c, r);
            epi_chunk_to_smem(r, smem_u32(s_bias + n0 + c), mine, lane, p.relu);
          }
          fence_proxy_async();
          // the store issued two chunks ago has finished READING its buffer -- the one chunk sc + 1 will be staged in
          FR_PROF_T0(pf_t1);
          if (issuer) bulk_wait_read<1>();
          asm volatile("bar.sync %0, %1;" ::"n"(kEpiBarrier), "n"(128) : "memory");
          FR_PROF_ADD(pf_buf, pf_t1);
          if (issuer && m0 < p.M) {   // rows past M inside the box are clipped by the TMA unit
            tma_store_2d(&tmap_out, buf, n0 + c, m0);
            bulk_commit();
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c, r);
          dot = epi_chunk_dot(r, smem_u32(s_bias + c), smem_u32(s_w4 + c), p.relu, dot);
        }
      }
      // this warp's quarter of the accumulator stage has been read: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(empty_remote + as * 8);
      if (EPI == EPI_DOT && row0 + lane < p.M) {
        if (p.b4) dot += p.b4[0];
        p.out[row0 + lane] = p.sigmoid ? 1.f / (1.f + __expf(-dot)) : dot;
      }
    }
    if (EPI == EPI_STORE && issuer) bulk_wait_all<0>();   // stores complete before the CTA retires
#ifdef FR_EXPERIMENTS
    if (p.prof && blockIdx.x < 8 && issuer) {
      long long* o = p.prof + blockIdx.x * 16 + 8;
      o[0] = clock64() - pf_start; o[1] = pf_acc; o[2] = pf_buf; o[3] = it;
    }
#endif
    (void)pf_acc; (void)pf_buf;
    tc_fence_before();
  }
  if (CTAS == 2) cluster_sync_all();   // the peer's smem / TMEM must stay alive until the pair is done
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem_base, L::kTmemCols);
    else tmem_dealloc(tmem_base, L::kTmemCols);
  }
#ifdef FR_EXPERIMENTS
  if (pf_tl) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    pf_tl[2] = t;
  }
#endif
}

// ---- the production kernel: CTA pairs, multi-slice TMA boxes, eight epilogue warps ------------------------------
// tc_pair_kernel is what every layer of every batch runs through; tc_linear_kernel above survives only in
// FR_EXPERIMENTS builds (1-CTA tiles, multicast clusters, the cp.async A loader).  Measurements behind it
// (tools/tma_probe.cu, tools/tc_prof.py, profiles/r02_*):
//   * The TMA unit of an SM retires roughly one load instruction per ~270-340 cycles whatever the box holds, up to
//     32 KB: 16 KB boxes (one 128-row x 128-byte K slice) land at 48 B/clk/SM, two per slice = 683 cycles = 0.35 us,
//     which is what the main loop of the old kernel ran at (0.36 us per slice; the four MMAs of a slice take 0.25-0.28).
//     So a box here holds KS = 2 K slices: the row-major operand [rows][K] is described as a 3-D tensor
//     {128 bytes, rows, K / slice} with strides {K * elt, 128} and fetched in boxes {128 B, 128 rows, 2 slices} = 32 KB,
//     which land as two consecutive 128B-swizzled slice tiles: 86 B/clk/SM, the MMAs become the limit.
//   * The epilogue of a 256-wide tile took 4.2 us with four warps (one per TMEM lane quarter: ~170 dependent
//     instructions per 32-column chunk on a single warp per scheduler, a proxy fence, a named barrier and a TMA
//     store) -- longer than the 4.0 us main loop of a layer-1 tile (K = 352), so layer 1 was epilogue-bound and
//     every launch ended in a 2.5-4 us tail.  Now EIGHT epilogue warps (two per lane quarter, alternate 128-byte
//     column chunks), each on its own: TMEM -> registers (bias, ReLU, rounding) -> a private 2 KB shared-memory
//     buffer, 16 rows at a time (128-byte swizzle, conflict-free) -> read back transposed -> coalesced 16-byte
//     global stores (eight lanes write the 128 contiguous bytes of a row, four rows per instruction: whole lines).
//     No proxy fences, no named barriers, no TMA stores, 16 KB of staging instead of 48: the rest went to the
//     operand ring.  (Measured on the way: storing straight from registers -- a thread owns a row, so 32 lanes hit
//     32 different lines per instruction -- 7.8 us per tile; half-line pieces -- 4.9 us.)
//   * What is left at batch 16384 is the memory system: with the stores of the epilogue switched off
//     (FR_TC_NOSTORE=1, experiments build) the main loop runs at 0.26 us per slice (layer 2: 25 instead of 35 us),
//     i.e. 33-67 MB of activation writes per layer next to 12+ TB/s of operand reads out of L2 are what the
//     MMAs wait for.
// Pipelines: smem ring (full / empty) TMA producer <-> MMA issuer; TMEM (tmem_full / tmem_empty) MMA issuer <->
// epilogue, two accumulator stages for tiles up to 256 wide.
constexpr int kPairThreads = 320;   // warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue
constexpr int kPairEpiWarps = 8;
constexpr int kEpiBufBytes = 32 * 64;   // one epilogue warp's transpose buffer: 64 bytes of each of its 32 rows

template <int BLOCK_N, int STAGES, int KS>
struct PairLayout {
  static constexpr int kASlice = BLOCK_M * 128;                   // one K slice of this CTA's 128 rows of A
  static constexpr int kABytes = kASlice * KS;
  static constexpr int kMmaN = BLOCK_N <= 256 ? BLOCK_N : 256;    // columns per tcgen05.mma
  static constexpr int kNSub = BLOCK_N / kMmaN;                   // MMAs per K step
  static constexpr int kSubRows = kMmaN / 2;                      // rows of Wt this CTA stages per MMA
  static constexpr int kSubSlice = kSubRows * 128;
  static constexpr int kSubBytes = kSubSlice * KS;                // [KS][kSubRows][128 B]
  static constexpr int kStageBytes = kABytes + kNSub * kSubBytes;
  static constexpr int kStoreOff = STAGES * kStageBytes;          // 8 epilogue warps x 2 KB transpose buffer
  static constexpr int kAuxOff = kStoreOff + kPairEpiWarps * kEpiBufBytes;   // bias[kMaxN], w4[256], dot partials [2][128]
  static constexpr int kBarOff = kAuxOff + (kMaxN + 256 + 256) * 4;
  static constexpr int kNumBars = 2 * STAGES + 4;                 // full, empty, tmem_full[2], tmem_empty[2]
  static constexpr int kTotal = kBarOff + kNumBars * 8 + 16;
  static constexpr int kDyn = kTotal + 1024;                      // slack for manual 1024-B alignment
  static constexpr int kAcc = BLOCK_N <= 256 ? 2 : 1;
  static constexpr int kTmemCols = kAcc * BLOCK_N;
};

// 3-D variant of the pair load: coordinates {byte column (always 0), row, K slice}
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* map, uint32_t bar_cluster, void* smem, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void stg128(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int BLOCK_N, int STAGES, int EPI, int ELT, int KS>
__global__ void __launch_bounds__(kPairThreads, 1)
tc_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const TcParams p) {
  constexpr int BK = 128 / ELT;   // elements of K per slice (one 128-byte swizzle row)
  using L = PairLayout<BLOCK_N, STAGES, KS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_bias = reinterpret_cast<float*>(smem + L::kAuxOff);
  float* s_w4 = s_bias + kMaxN;
  float* s_dot = s_w4 + 256;                           // [2 accumulator stages][128 rows]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2], the leader's are the ones used
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int num_kb = (p.K + BK - 1) / BK;             // K slices
  const int num_st = (num_kb + KS - 1) / KS;          // ring stages per tile
  const int n_tiles_n = p.N / BLOCK_N;
  const int n_tiles = ((p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * n_tiles_n;   // a tile = 256 rows x BLOCK_N
  const int cluster_id = blockIdx.x / 2, n_clusters = gridDim.x / 2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full_bar[s], 1);    // the leader's producer arrives once, bytes of both CTAs are expected
      mbar_init(&empty_bar[s], 1);   // one multicast tcgen05.commit per use
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(&tmem_full_bar[a], 1);                    // one multicast tcgen05.commit per tile
      mbar_init(&tmem_empty_bar[a], 2 * kPairEpiWarps);   // every epilogue warp of both CTAs
    }
    fence_barrier_init();
  } else if (warp == 1) {
    tmem_alloc_pair(tmem_ptr, L::kTmemCols);
  } else if (warp >= kEpiWarp0) {
    // weights, not produced by the preceding kernel: staged before the grid dependency resolves
    const int et = threadIdx.x - kEpiWarp0 * 32;  // 0..255
    for (int i = et; i < p.N; i += kPairEpiWarps * 32) s_bias[i] = p.bias ? p.bias[i] : 0.f;
    if (EPI == EPI_DOT)
      for (int i = et; i < BLOCK_N; i += kPairEpiWarps * 32) s_w4[i] = p.w4[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // peer barriers must exist before any remote complete_tx / commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (p.pdl) griddep_wait();           // activations of the previous kernel are complete and visible from here

  if (warp == 0) {
    // ===== TMA producer (both CTAs of the pair) =====
    if (lane == 0) {
      if (p.wait_flags) wait_for_peers(p);
      long long pf_wait = 0;
      FR_PROF_T0(pf_start);
      const uint32_t bar0 = mapa_u32(smem_u32(&full_bar[0]), 0);   // the leader's full barriers
      uint32_t kc = 0;   // ring stages issued so far (the ring position runs on across tiles)
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
        const int m0 = ((tile / n_tiles_n) * 2 + (int)rank) * BLOCK_M;
        const int nb0 = (tile % n_tiles_n) * BLOCK_N + (int)rank * L::kSubRows;
        for (int st = 0; st < num_st; st++, kc++) {
          const int s = kc % STAGES;
          FR_PROF_T0(pf_t);
          mbar_wait(&empty_bar[s], ((kc / STAGES) & 1) ^ 1, 1, kc, tile);
          FR_PROF_ADD(pf_wait, pf_t);
          uint8_t* a_dst = smem + s * L::kStageBytes;
          uint8_t* b_dst = a_dst + L::kABytes;
          if (leader) mbar_expect_tx(&full_bar[s], 2 * L::kStageBytes);
          if (KS == 1) {
            tma_load_2d_pair(&tmap_a, bar0 + s * 8, a_dst, st * BK, m0);
#pragma unroll
            for (int h = 0; h < L::kNSub; h++) tma_load_2d_pair(&tmap_b, bar0 + s * 8, b_dst + h * L::kSubBytes, st * BK, nb0 + h * L::kMmaN);
          } else {   // K slices past K are zero-filled by the TMA unit (and their MMAs skipped)
            tma_load_3d_pair(&tmap_a, bar0 + s * 8, a_dst, 0, m0, st * KS);
#pragma unroll
            for (int h = 0; h < L::kNSub; h++) tma_load_3d_pair(&tmap_b, bar0 + s * 8, b_dst + h * L::kSubBytes, 0, nb0 + h * L::kMmaN, st * KS);
          }
        }
      }
      if (p.pdl) griddep_launch();
#ifdef FR_EXPERIMENTS
      if (p.prof && blockIdx.x < 8) {
        long long* o = p.prof + blockIdx.x * 16;
        o[0] = clock64() - pf_start; o[1] = pf_wait; o[2] = kc;
      }
#endif
      (void)pf_wait;
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (the pair leader) =====
    if (leader) {
      constexpr uint32_t idesc = ELT == 2 ? make_idesc_f16_m(2 * BLOCK_M, L::kMmaN) : make_idesc_tf32_m(2 * BLOCK_M, L::kMmaN);
      uint32_t kc = 0, it = 0;
      long long pf_full = 0, pf_tmem = 0;
      FR_PROF_T0(pf_start);
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, it++) {
        const uint32_t as = it % L::kAcc;
        FR_PROF_T0(pf_t0);
        mbar_wait(&tmem_empty_bar[as], ((it / L::kAcc) & 1) ^ 1, 2, kc, tile);   // the epilogue has drained this accumulator stage
        FR_PROF_ADD(pf_tmem, pf_t0);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int st = 0; st < num_st; st++, kc++) {
          const int s = kc % STAGES;
          FR_PROF_T0(pf_t1);
          mbar_wait(&full_bar[s], (kc / STAGES) & 1, 3, kc, tile);
          FR_PROF_ADD(pf_full, pf_t1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
            const uint32_t b_addr = a_addr + L::kABytes;
#pragma unroll
            for (int j = 0; j < KS; j++) {
              if (st * KS + j < num_kb) {
                const uint64_t a_desc = make_smem_desc(a_addr + j * L::kASlice);
#pragma unroll
                for (int k = 0; k < 4; k++) {   // 32 bytes of K per instruction: +2 in the (addr >> 4) field
#pragma unroll
                  for (int h = 0; h < L::kNSub; h++) {
                    const uint64_t b_desc = make_smem_desc(b_addr + h * L::kSubBytes + j * L::kSubSlice);
                    const uint32_t acc = (st | j | k) != 0;
                    if (ELT == 2) umma_f16_pair(d_tmem + h * L::kMmaN, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, acc);
                    else umma_tf32_pair(d_tmem + h * L::kMmaN, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, acc);
                  }
                }
              }
            }
            umma_commit_pair(&empty_bar[s]);                            // frees the slot in both CTAs
            if (st == num_st - 1) umma_commit_pair(&tmem_full_bar[as]);  // accumulators complete in both CTAs
          }
          __syncwarp();
        }
      }
#ifdef FR_EXPERIMENTS
      if (p.prof && blockIdx.x < 8 && lane == 0) {
        long long* o = p.prof + blockIdx.x * 16 + 4;
        o[0] = clock64() - pf_start; o[1] = pf_full; o[2] = pf_tmem; o[3] = kc;
      }
#endif
      (void)pf_full; (void)pf_tmem;
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter q = warp % 4, two warps per quarter on alternate chunks =====
    const int q = warp % 4, half = (warp - kEpiWarp0) / 4;
    const uint32_t empty_remote = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);   // the leader's tmem_empty_bar[0]
    uint32_t it = 0;
    long long pf_acc = 0;
    FR_PROF_T0(pf_start);
    for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, it++) {
      const uint32_t as = it % L::kAcc;
      const int row = ((tile / n_tiles_n) * 2 + (int)rank) * BLOCK_M + q * 32 + lane;
      const int n0 = (tile % n_tiles_n) * BLOCK_N;
      FR_PROF_T0(pf_t0);
      mbar_wait(&tmem_full_bar[as], (it / L::kAcc) & 1, 4, it, tile);
      FR_PROF_ADD(pf_acc, pf_t0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N;
      float dot = 0.f;
      if (EPI == EPI_STORE) {
        // a chunk = 128 bytes (one line) of every row of this warp: 32 fp32 / 64 fp16 output columns, staged 16 rows at
        // a time in the warp's 2 KB buffer (classic 128-byte swizzle: piece j of row r at piece j ^ (r & 7))
        constexpr int CH = 128 / ELT;
        const int row_base = row - lane;
        const uint32_t wbuf = smem_u32(smem + L::kStoreOff + (warp - kEpiWarp0) * kEpiBufBytes);
        const uint32_t st_row = wbuf + (lane & 15) * 128, st_sw = (uint32_t)(lane & 7);
        uint8_t* obase = reinterpret_cast<uint8_t*>(p.out_act) + (size_t)n0 * ELT;
#pragma unroll 1
        for (int c = half * CH; c < BLOCK_N; c += 2 * CH) {
          uint32_t o[32];
          const uint32_t ba = smem_u32(s_bias + n0 + c);
          if (ELT == 2) {
            uint32_t r0[32], r1[32];
            tmem_ld32(taddr + c, r0);
            tmem_ld32(taddr + c + 32, r1);
#pragma unroll
            for (int j = 0; j < 16; j++) {   // 4 columns -> 2 packed registers
              const float4 bv = lds128(ba + j * 16);
              const uint32_t* r = j < 8 ? r0 + 4 * j : r1 + 4 * (j - 8);
              float v0 = __uint_as_float(r[0]) + bv.x, v1 = __uint_as_float(r[1]) + bv.y;
              float v2 = __uint_as_float(r[2]) + bv.z, v3 = __uint_as_float(r[3]) + bv.w;
              if (p.relu) {
                v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
              }
              o[2 * j] = pack_f16x2(v0, v1);
              o[2 * j + 1] = pack_f16x2(v2, v3);
            }
          } else {
            uint32_t r[32];
            tmem_ld32(taddr + c, r);
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const float4 bv = lds128(ba + j * 16);
              float v0 = __uint_as_float(r[4 * j]) + bv.x, v1 = __uint_as_float(r[4 * j + 1]) + bv.y;
              float v2 = __uint_as_float(r[4 * j + 2]) + bv.z, v3 = __uint_as_float(r[4 * j + 3]) + bv.w;
              if (p.relu) {
                v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
              }
              o[4 * j] = __float_as_uint(round_tf32(v0)); o[4 * j + 1] = __float_as_uint(round_tf32(v1));
              o[4 * j + 2] = __float_as_uint(round_tf32(v2)); o[4 * j + 3] = __float_as_uint(round_tf32(v3));
            }
          }
#pragma unroll
          for (int hh = 0; hh < 2; hh++) {   // rows 0..15 of the warp, then rows 16..31
            __syncwarp();                    // the previous 16 rows have been read back
            if ((lane >> 4) == hh) {
#pragma unroll
              for (int j = 0; j < 8; j++)
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(st_row + ((j ^ st_sw) << 4)), "r"(o[4 * j]), "r"(o[4 * j + 1]),
                             "r"(o[4 * j + 2]), "r"(o[4 * j + 3]) : "memory");
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; i++) {    // eight lanes write the 128 contiguous bytes of a row, four rows per instruction
              const int R = (lane >> 3) + 4 * i, P = lane & 7;
              uint32_t v0, v1, v2, v3;
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                           : "r"(wbuf + R * 128 + ((P ^ (R & 7)) << 4)));
              const int grow = row_base + hh * 16 + R;
              if (grow < p.M && !p.dbg_nostore) stg128(obase + ((size_t)grow * p.N + c) * ELT + P * 16, v0, v1, v2, v3);
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c = half * 32; c < BLOCK_N; c += 64) {
          uint32_t r[32];
          tmem_ld32(taddr + c, r);
          dot = epi_chunk_dot(r, smem_u32(s_bias + c), smem_u32(s_w4 + c), p.relu, dot);
        }
      }
      // this warp's share of the accumulator stage has been read: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(empty_remote + as * 8);
      if (EPI == EPI_DOT) {
        // the two warps of a lane quarter hold the even / odd chunks' partial sums of the same rows: the second
        // hands its sum over in shared memory (double-buffered by accumulator stage), the first finishes the row
        float* slot = s_dot + as * BLOCK_M + q * 32 + lane;
        if (half == 1) *slot = dot;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(64) : "memory");
        if (half == 0 && row < p.M) {
          dot += *slot;
          if (p.b4) dot += p.b4[0];
          p.out[row] = p.sigmoid ? 1.f / (1.f + __expf(-dot)) : dot;
        }
      }
    }
#ifdef FR_EXPERIMENTS
    if (p.prof && blockIdx.x < 8 && warp == kEpiWarp0 && lane == 0) {
      long long* o = p.prof + blockIdx.x * 16 + 8;
      o[0] = clock64() - pf_start; o[1] = pf_acc; o[2] = 0; o[3] = it;
    }
#endif
    (void)pf_acc;
    tc_fence_before();
  }
  cluster_sync_all();   // the peer's smem / TMEM must stay alive until the pair is done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, L::kTmemCols);
  }
}

// ---- the whole MLP in one persistent launch ---------------------------------------
// tc_linear_kernel pays its fixed cost (TMEM allocation, barrier set-up, pipeline fill, the last
// epilogue, teardown: ~6 us) once per layer, and at batch 2048 a layer is only 8..32 tiles: a third
// to a half of every launch is that fixed cost (small model, 12 batches in flight: 47 % of the TF32
// peak for the whole step).  Here a CTA pair owns 256 ITEMS and walks them through all the layers
// (cuda_server.c:468-491, the four cublasLtMatmul calls, become the phases of one kernel):
//   layer 1   N1/512 phases   D[256 x 512] = X[256 x K0] . Wt1[chunk]^T    -> H1 (tf32-rounded, TMA store)
//   layer 2   N2/512 phases   D[256 x 512] = H1[256 x N1] . Wt2[chunk]^T   -> H2
//   layer 3   1 phase         D[256 x 256] = H2 . Wt3^T, output layer + sigmoid folded -> score
// The accumulator is ONE TMEM stage of 512 columns (two N = 256 MMAs per K step sharing the A slice,
// 33 B/clk/SM of operand fill), drained by EIGHT epilogue warps (two per TMEM lane quarter, half the
// columns each).  H1 / H2 make a round trip through global memory (L2-resident: 1 MB + 0.5 MB per
// pair) because 128 rows x 1024 fp32 do not fit an SM; the rows a CTA stores are the rows the same
// CTA loads again, so the dependency is CTA-local: an epilogue warp waits for its bulk stores to
// COMPLETE (cp.async.bulk.wait_group, not .read) and arrives on ready[layer][chunk]; the producer
// waits on that barrier only before the first K slice that reads the chunk.  Weight slices never
// wait: the producer runs ahead into the next phase while the epilogue drains, and layer 2's first
// 16 K slices read the H1 chunk that was stored one phase earlier, so the only exposed gaps are the
// TMEM hand-overs (~1 us each) and the H2 turn-around before layer 3.
constexpr int kChainThreads = 320;     // warp 0 producer, warp 1 MMA issuer, warps 2..9 epilogue
constexpr int kChainEpiWarps = 8;
constexpr int kChainMaxBias = 3072;    // N1 + N2 + N3 floats staged in smem (large model: 2048 + 512 + 256)
constexpr int kChainMaxChunks = 4;     // 512-wide chunks of the widest storing layer (2048)
constexpr int kChainW = 512;           // phase width of the storing layers

template <int STAGES>
struct ChainLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 4;          // 16 KB: this CTA's 128 items x 32 floats
  static constexpr int kBSubBytes = 128 * BLOCK_K * 4;           // 16 KB: this CTA's 128 rows of Wt for one N = 256 MMA
  static constexpr int kStageBytes = kABytes + 2 * kBSubBytes;   // 48 KB
  static constexpr int kStoreOff = STAGES * kStageBytes;
  static constexpr int kAuxOff = kStoreOff + kChainEpiWarps * 2 * kStoreBufBytes;   // bias[kChainMaxBias], w4[256]
  static constexpr int kBarOff = kAuxOff + (kChainMaxBias + 256) * 4;
  static constexpr int kNumBars = 2 * STAGES + 2 + 2 * kChainMaxChunks;   // full, empty, tmem_full, tmem_empty, ready[2][4]
  static constexpr int kTotal = kBarOff + kNumBars * 8 + 16;
  static constexpr int kDyn = kTotal + 1024;
};

struct ChainMaps {
  CUtensorMap a[3];   // X, H1, H2 as A operands: 128-row x 32-float boxes
  CUtensorMap o[2];   // H1, H2 as store targets: 32-row x 32-float boxes
  CUtensorMap w[3];   // Wt1, Wt2, Wt3: 128-row boxes (one CTA's half of an N = 256 MMA)
};

struct ChainParams {
  const float* bias[3];   // null in LINEAR mode
  const float* w4;        // output-layer weights [256]
  const float* b4;        // output-layer bias [1] or null
  float* out;             // scores [M]
  int M;
  int dims[4];            // K0, N1, N2, N3 (= 256)
  int relu, sigmoid;
  long long* prof;        // FR_CHAIN_PROF=1: clock64 stamps of CTA 0's first item tiles (kProf* below), else null
  const float* probe_src; // PROBE: buffer the probe warps stream (L2-resident)
  uint32_t probe_vecs;    // its size in float4
};

// timeline of CTA 0: prof[((item tile iteration * kProfPhases) + phase) * kProfSlots + slot]
constexpr int kProfIters = 4, kProfPhases = 8, kProfSlots = 8;
enum { PROF_MMA_TMEM_FREE = 0, PROF_MMA_FIRST_FULL, PROF_MMA_LAST_ISSUED, PROF_EPI_FULL_SEEN, PROF_EPI_TMEM_RELEASED,
       PROF_EPI_STORES_DONE, PROF_PROD_READY_WAIT, PROF_PROD_READY_OK };
__device__ __forceinline__ void prof_stamp(long long* prof, uint32_t it, int ph, int slot) {
  if (prof && it < (uint32_t)kProfIters && ph < kProfPhases) prof[((int)it * kProfPhases + ph) * kProfSlots + slot] = clock64();
}

// PROBE (experiment, FR_CHAIN_PROBE=1): four extra warps stream 16-byte L2 loads through the LSU while the TMA
// unit feeds the MMAs, to see whether the two paths into an SM share one ingest limit (tools/chain_timeline.py).
template <int STAGES, bool PROBE>
__global__ void __launch_bounds__(kChainThreads + (PROBE ? 128 : 0), 1)
tc_mlp_chain_kernel(const __grid_constant__ ChainMaps maps, const ChainParams p) {
  using L = ChainLayout<STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_bias = reinterpret_cast<float*>(smem + L::kAuxOff);
  float* s_w4 = s_bias + kChainMaxBias;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;
  uint64_t* ready_bar = tmem_empty_bar + 1;          // [2][kChainMaxChunks]: chunk c of layer l's output is in global memory
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(ready_bar + 2 * kChainMaxChunks);
  volatile uint32_t* probe_done = tmem_ptr + 1;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int n_tiles = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);   // item tiles of 256
  if (PROBE && threadIdx.x == 0) *probe_done = 0;
  const int cluster_id = blockIdx.x / 2, n_clusters = gridDim.x / 2;
  long long* const prof = (blockIdx.x == 0 && lane == 0) ? p.prof : nullptr;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 3; i++) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[i]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w[i]) : "memory");
      if (i < 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.o[i]) : "memory");
    }
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full_bar[s], 1);    // the leader's producer arrives once; bytes of both CTAs are expected
      mbar_init(&empty_bar[s], 1);   // one multicast tcgen05.commit per use
    }
    mbar_init(tmem_full_bar, 1);                        // one multicast tcgen05.commit per phase
    mbar_init(tmem_empty_bar, 2 * kChainEpiWarps);      // every epilogue warp of both CTAs
    for (int i = 0; i < 2 * kChainMaxChunks; i++) mbar_init(&ready_bar[i], kChainEpiWarps);   // this CTA's epilogue warps
    fence_barrier_init();
  } else if (warp == 1) {
    tmem_alloc_pair(tmem_ptr, 512);
  } else if (warp >= kEpiWarp0) {
    const int et = threadIdx.x - kEpiWarp0 * 32;   // 0..255
    int off = 0;
    for (int l = 0; l < 3; l++) {
      const int N = p.dims[l + 1];
      for (int i = et; i < N; i += kChainEpiWarps * 32) s_bias[off + i] = p.bias[l] ? p.bias[l][i] : 0.f;
      off += N;
    }
    for (int i = et; i < 256; i += kChainEpiWarps * 32) s_w4[i] = p.w4[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): weights run ahead, activations wait for their chunk =====
    if (lane == 0) {
      const uint32_t bar0 = mapa_u32(smem_u32(&full_bar[0]), 0);   // the leader's full barriers
      uint32_t kc = 0, it = 0;
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, it++) {
        const int m0 = (tile * 2 + (int)rank) * BLOCK_M;
        int ph = 0;
        for (int l = 0; l < 3; l++) {
          const int num_kb = (p.dims[l] + BLOCK_K - 1) / BLOCK_K;
          const int width = l < 2 ? kChainW : 256;
          const int nsub = width / 256;
          const int n_chunks = p.dims[l + 1] / width;
          int ready_upto = 0;   // chunks of layer l-1's output known to be in global memory (this item tile)
          for (int c = 0; c < n_chunks; c++, ph++) {
            const int nb0 = c * width + (int)rank * 128;
            for (int kb = 0; kb < num_kb; kb++, kc++) {
              const int s = kc % STAGES;
              mbar_wait(&empty_bar[s], ((kc / STAGES) & 1) ^ 1, 1, kc, tile);
              uint8_t* a_dst = smem + s * L::kStageBytes;
              uint8_t* b_dst = a_dst + L::kABytes;
              const uint32_t bar = bar0 + s * 8;
              if (leader) mbar_expect_tx(&full_bar[s], 2 * (L::kABytes + nsub * L::kBSubBytes));
              for (int h = 0; h < nsub; h++)
                tma_load_2d_pair(&maps.w[l], bar, b_dst + h * L::kBSubBytes, kb * BLOCK_K, nb0 + h * 256);
              if (l > 0) {
                const int need = (kb * BLOCK_K) / kChainW;   // chunk of the previous layer this K slice reads
                while (ready_upto <= need) {
                  prof_stamp(prof, it, ph, PROF_PROD_READY_WAIT);
                  mbar_wait(&ready_bar[(l - 1) * kChainMaxChunks + ready_upto], it & 1, 6, kc, tile);
                  prof_stamp(prof, it, ph, PROF_PROD_READY_OK);
                  ready_upto++;
                }
                // the rows were written by this CTA's bulk stores (async proxy) and ordered before us by
                // generic-proxy barrier operations: order our async-proxy reads after those
                asm volatile("fence.proxy.async;" ::: "memory");
              }
              tma_load_2d_pair(&maps.a[l], bar, a_dst, kb * BLOCK_K, m0);
            }
          }
        }
      }
      if (PROBE) *probe_done = 1;
    }
    __syncwarp();
  } else if (PROBE && warp >= kEpiWarp0 + kChainEpiWarps) {
    // ===== probe warps: 16 KB of L2 loads per iteration through the LSU until the producer is done =====
    const int t = threadIdx.x - (kEpiWarp0 + kChainEpiWarps) * 32;   // 0..127
    const float4* src = reinterpret_cast<const float4*>(p.probe_src);
    const uint32_t nvec = p.probe_vecs;
    uint32_t base = (blockIdx.x * 8191u) % nvec, iters = 0;
    float acc = 0.f;
    const long long t0 = clock64();
    while (*probe_done == 0) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        float4 v;
        const float4* a = src + (base + u * 128 + t) % nvec;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a));
        acc += v.x + v.y + v.z + v.w;
      }
      base = (base + 1024) % nvec;
      iters++;
    }
    if (p.prof && blockIdx.x == 0 && t == 0) {
      long long* a = p.prof + (kProfIters - 1) * kProfPhases * kProfSlots + 7 * kProfSlots;   // last row of the buffer
      a[0] = iters; a[1] = clock64() - t0; a[2] = (long long)acc;
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA) =====
    if (leader) {
      constexpr uint32_t idesc = make_idesc_tf32_m(2 * BLOCK_M, 256);
      uint32_t kc = 0, pc = 0, it = 0;   // K slices consumed, phases started, item tiles started
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, it++) {
        int ph = 0;
        for (int l = 0; l < 3; l++) {
          const int num_kb = (p.dims[l] + BLOCK_K - 1) / BLOCK_K;
          const int width = l < 2 ? kChainW : 256;
          const int nsub = width / 256;
          const int n_chunks = p.dims[l + 1] / width;
          for (int c = 0; c < n_chunks; c++, pc++, ph++) {
            mbar_wait(tmem_empty_bar, (pc & 1) ^ 1, 2, kc, tile);   // the epilogue has drained the accumulator
            tc_fence_after();
            prof_stamp(prof, it, ph, PROF_MMA_TMEM_FREE);
            for (int kb = 0; kb < num_kb; kb++, kc++) {
              const int s = kc % STAGES;
              mbar_wait(&full_bar[s], (kc / STAGES) & 1, 3, kc, tile);
              tc_fence_after();
              if (kb == 0) prof_stamp(prof, it, ph, PROF_MMA_FIRST_FULL);
              if (elect_one()) {
                const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
                const uint64_t a_desc = make_smem_desc(a_addr);
                const uint64_t b_desc = make_smem_desc(a_addr + L::kABytes);
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; k++) {
                  umma_tf32_pair(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                  if (nsub == 2)
                    umma_tf32_pair(tmem_base + 256, a_desc + (uint64_t)(k * 2),
                                   b_desc + (uint64_t)(k * 2) + (uint64_t)(L::kBSubBytes >> 4), idesc, (kb | k) != 0);
                }
                umma_commit_pair(&empty_bar[s]);
                if (kb == num_kb - 1) umma_commit_pair(tmem_full_bar);
              }
              __syncwarp();
            }
            prof_stamp(prof, it, ph, PROF_MMA_LAST_ISSUED);
          }
        }
      }
    }
  } else {
    // ===== epilogue: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
    const int q = warp % 4, half = (warp - kEpiWarp0) / 4;
    uint8_t* store_buf = smem + L::kStoreOff + (warp - kEpiWarp0) * 2 * kStoreBufBytes;
    const uint32_t empty_remote = mapa_u32(smem_u32(tmem_empty_bar), 0);   // the leader's
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t pc = 0, sc = 0, it = 0;
    long long* const eprof = (warp == kEpiWarp0) ? prof : nullptr;   // one epilogue warp keeps the timeline
    for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, it++) {
      const int row0 = (tile * 2 + (int)rank) * BLOCK_M + q * 32;
      int boff = 0, ph = 0;
      for (int l = 0; l < 2; l++) {
        const int n_chunks = p.dims[l + 1] / kChainW;
        for (int c = 0; c < n_chunks; c++, pc++, ph++) {
          mbar_wait(tmem_full_bar, pc & 1, 4, pc, tile);
          tc_fence_after();
          prof_stamp(eprof, it, ph, PROF_EPI_FULL_SEEN);
          long long t_ld = 0, t_wr = 0, t_st = 0, t_fe = 0, t_is = 0;   // FR_CHAIN_PROF: where this warp's epilogue time goes
#pragma unroll 1
          for (int cc = half * 256; cc < half * 256 + 256; cc += 32) {
            uint32_t r[32];
            const long long c0 = eprof ? clock64() : 0;
            tmem_ld32(taddr + cc, r);
            const long long c1 = eprof ? clock64() : 0;
            uint8_t* buf = store_buf + (sc & 1) * kStoreBufBytes;
            if (lane == 0) bulk_wait_read<1>();   // the store issued two chunks ago no longer reads `buf`
            __syncwarp();
            const long long c2 = eprof ? clock64() : 0;
            epi_chunk_to_smem(r, smem_u32(s_bias + boff + c * kChainW + cc), smem_u32(buf), lane, p.relu);
            const long long c2b = eprof ? clock64() : 0;
            fence_proxy_async();
            __syncwarp();
            const long long c3 = eprof ? clock64() : 0;
            if (lane == 0 && row0 < p.M) {
              tma_store_2d(&maps.o[l], buf, c * kChainW + cc, row0);
              bulk_commit();
            }
            sc++;
            if (eprof) {
              t_ld += c1 - c0; t_wr += c2 - c1; t_st += c2b - c2; t_fe += c3 - c2b; t_is += clock64() - c3;
            }
          }
          if (eprof && it == 0 && ph < 4) {   // rows 4..7 of iteration 0 carry the breakdown of phases 0..3
            long long* a = eprof + (4 + ph) * kProfSlots;
            a[0] = t_ld; a[1] = t_wr; a[2] = t_st; a[3] = t_is; a[4] = t_fe;
          }
          // accumulator read: hand TMEM back first, then wait for this warp's rows to be in global memory
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_cluster(empty_remote);
            prof_stamp(eprof, it, ph, PROF_EPI_TMEM_RELEASED);
            bulk_wait_all<0>();
            asm volatile("fence.proxy.async;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&ready_bar[l * kChainMaxChunks + c])) : "memory");
            prof_stamp(eprof, it, ph, PROF_EPI_STORES_DONE);
          }
        }
        boff += p.dims[l + 1];
      }
      // layer 3: the 256-wide row stays in TMEM; output layer + sigmoid folded (warps of column half 0)
      mbar_wait(tmem_full_bar, pc & 1, 4, pc, tile);
      tc_fence_after();
      prof_stamp(eprof, it, ph, PROF_EPI_FULL_SEEN);
      float dot = 0.f;
      if (half == 0) {
#pragma unroll 1
        for (int cc = 0; cc < 256; cc += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + cc, r);
          dot = epi_chunk_dot(r, smem_u32(s_bias + boff + cc), smem_u32(s_w4 + cc), p.relu, dot);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(empty_remote);
      prof_stamp(eprof, it, ph, PROF_EPI_TMEM_RELEASED);
      if (half == 0 && row0 + lane < p.M) {
        if (p.b4) dot += p.b4[0];
        p.out[row0 + lane] = p.sigmoid ? 1.f / (1.f + __expf(-dot)) : dot;
      }
      pc++;
    }
    tc_fence_before();
  }
  cluster_sync_all();   // the peer's smem / TMEM must stay alive until the pair is done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ---- lookup fused into layer 1 ---------------------------------------------------
// The north-star fusion: the multi-table lookup IS the A-operand producer of the first GEMM, so the
// concat vectors never exist in global memory.  One CTA pair owns 256 items x 512 hidden units:
//   warp 0        TMA producer for the weights: two 128-row boxes of Wt per CTA per K slice
//   warp 1        MMA issuer (leader): per K step two cta_group::2 MMAs (N = 256 each) that share
//                 the A tile and fill all 512 TMEM columns
//   warps 2-5     epilogue: bias + ReLU + TF32 rounding, TMA store of H1
//   warps 6-13    lookup producers: for every 32-float K slice, 8 lanes per item read the item's 8
//                 row pieces (index -> 16-byte row load, exactly the stand-alone lookup:
//                 embedding_47_krnl.cpp:916-935 + the concat order of gather_embeddings) and write
//                 them, TF32-rounded, into the 128B-swizzled A slot.  Two groups of 4 warps take
//                 alternate slices and each keeps the next slice's rows and the one after's indices
//                 in flight in registers, so four slices of random-access latency overlap the MMAs.
// Each 256 x 512 tile re-reads its items' rows once per 512 hidden units (2x for H1 = 1024, 4x for
// 2048): those re-reads hit L2, and the alternative -- X written to and re-read from global memory
// by 4 column tiles -- moved more bytes.  Per SM and K slice 16 KB of A + 32 KB of B feed 2 x 4 MMAs
// (33 B/clk/SM at the measured TF32 rate, against 45 for 256 x 256 tiles).
constexpr int kFuseThreads = 448;
constexpr int kFuseGatherWarp0 = 6;
constexpr int kFuseN = 512;
constexpr int kFuseMaxChunks = 1024;   // 16-byte pieces per item: large model 992

template <int STAGES>
struct FuseLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 4;        // 16 KB: this CTA's 128 items x 32 floats
  static constexpr int kBHalfBytes = 128 * BLOCK_K * 4;        // 16 KB: 128 rows of Wt (this CTA's half of an N=256 MMA)
  static constexpr int kBBytes = 2 * kBHalfBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;        // 48 KB
  static constexpr int kStoreOff = STAGES * kStageBytes;
  static constexpr int kAuxOff = kStoreOff + 4 * 2 * kStoreBufBytes;       // bias[kMaxN]
  static constexpr int kChunkOff = kAuxOff + kMaxN * 4;                    // FrFuseChunk[kFuseMaxChunks]
  static constexpr int kBarOff = kChunkOff + kFuseMaxChunks * 16;
  static constexpr int kNumBars = 2 * STAGES + 2;                          // full, empty, tmem_full, tmem_empty
  static constexpr int kTotal = kBarOff + kNumBars * 8 + 16;
  static constexpr int kDyn = kTotal + 1024;
};

struct FuseParams {
  const FrFuseChunk* chunks;   // [C] piece descriptors in concat (wire) order
  const int32_t* idx;          // [M][T]
  const float* bias;           // [N] or null
  int C, T, M, N;
  int relu;
};

template <int STAGES>
__global__ void __launch_bounds__(kFuseThreads, 1)
tc_gather_linear_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_out,
                        const FuseParams p) {
  using L = FuseLayout<STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_bias = reinterpret_cast<float*>(smem + L::kAuxOff);
  FrFuseChunk* s_chunks = reinterpret_cast<FrFuseChunk*>(smem + L::kChunkOff);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int num_kb = (p.C + 7) / 8;                 // 8 pieces = 32 floats per K slice; the tail slice is zero-padded
  const int n_tiles_n = p.N / kFuseN;
  const int n_tiles = ((p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * n_tiles_n;
  const int cluster_id = blockIdx.x / 2, n_clusters = gridDim.x / 2;
  const int my_tiles = cluster_id < n_tiles ? (n_tiles - cluster_id + n_clusters - 1) / n_clusters : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_out) : "memory");
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full_bar[s], 1 + 8);   // the leader's weight producer + 4 lookup warps of each CTA
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, 8);
    fence_barrier_init();
  } else if (warp == 1) {
    tmem_alloc_pair(tmem_ptr, 512);
  } else if (warp >= kFuseGatherWarp0) {
    for (int i = threadIdx.x - kFuseGatherWarp0 * 32; i < p.C; i += (kFuseThreads - kFuseGatherWarp0 * 32)) s_chunks[i] = p.chunks[i];
  } else {
    for (int i = threadIdx.x - kEpiWarp0 * 32; i < p.N; i += 128) s_bias[i] = p.bias ? p.bias[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== weight producer (both CTAs) =====
    if (lane == 0) {
      uint32_t kc = 0;
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
        const int n0 = (tile % n_tiles_n) * kFuseN + (int)rank * 128;
        for (int kb = 0; kb < num_kb; kb++, kc++) {
          const int s = kc % STAGES;
          mbar_wait(&empty_bar[s], ((kc / STAGES) & 1) ^ 1, 1, kc, tile);
          uint8_t* b_dst = smem + s * L::kStageBytes + L::kABytes;
          const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
          if (leader) mbar_expect_tx(&full_bar[s], 2 * L::kBBytes);
          tma_load_2d_pair(&tmap_b, bar, b_dst, kb * BLOCK_K, n0);
          tma_load_2d_pair(&tmap_b, bar, b_dst + L::kBHalfBytes, kb * BLOCK_K, n0 + 256);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (leader) =====
    if (leader) {
      constexpr uint32_t idesc = make_idesc_tf32_m(2 * BLOCK_M, 256);
      uint32_t kc = 0, it = 0;
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, it++) {
        mbar_wait(tmem_empty_bar, (it & 1) ^ 1, 2, kc, tile);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; kb++, kc++) {
          const int s = kc % STAGES;
          mbar_wait_cluster(&full_bar[s], (kc / STAGES) & 1, 3, kc, tile);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
            const uint64_t a_desc = make_smem_desc(a_addr);
            const uint64_t b0_desc = make_smem_desc(a_addr + L::kABytes);
            const uint64_t b1_desc = make_smem_desc(a_addr + L::kABytes + L::kBHalfBytes);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; k++) {
              umma_tf32_pair(tmem_base, a_desc + (uint64_t)(k * 2), b0_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
              umma_tf32_pair(tmem_base + 256, a_desc + (uint64_t)(k * 2), b1_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            }
            umma_commit_pair(&empty_bar[s]);
            if (kb == num_kb - 1) umma_commit_pair(tmem_full_bar);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < kFuseGatherWarp0) {
    // ===== epilogue =====
    const int q = warp % 4;
    uint8_t* store_buf = smem + L::kStoreOff + (warp - kEpiWarp0) * 2 * kStoreBufBytes;
    const uint32_t empty_remote = mapa_u32(smem_u32(tmem_empty_bar), 0);
    uint32_t it = 0, sc = 0;
    for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, it++) {
      const int row0 = ((tile / n_tiles_n) * 2 + (int)rank) * BLOCK_M + q * 32;
      const int n0 = (tile % n_tiles_n) * kFuseN;
      mbar_wait(tmem_full_bar, it & 1, 4, it, tile);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < kFuseN; c += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c, r);
        uint8_t* buf = store_buf + (sc & 1) * kStoreBufBytes;
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        epi_chunk_to_smem(r, smem_u32(s_bias + n0 + c), smem_u32(buf), lane, p.relu);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && row0 < p.M) {
          tma_store_2d(&tmap_out, buf, n0 + c, row0);
          bulk_commit();
        }
        sc++;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(empty_remote);
    }
    if (lane == 0) bulk_wait_all<0>();
    tc_fence_before();
  } else {
    // ===== lookup producers; group 0 / 1 (4 warps each) takes even / odd K slices =====
    // Within a slice a thread owns ONE of the 8 pieces (j) of 8 items (i0, i0+16, ...): the 8 lanes
    // of an item read neighbouring index columns of one index row and neighbouring pieces (often of
    // the same table row), so a warp-wide load touches a few cache lines instead of 32 -- the
    // L1 wavefront rate, not latency, is what bounds a gather feeding a GEMM at this rate.
    const int gt = threadIdx.x - kFuseGatherWarp0 * 32;   // 0..255
    const int group = gt >> 7, g = gt & 127;
    const int j = g & 7, i0 = g >> 3;                     // piece within the slice, first item (of 8, stride 16)
    const uint32_t full_remote = mapa_u32(smem_u32(&full_bar[0]), 0);
    const int total = my_tiles * num_kb;                  // K slices this cluster consumes, in order

    // Two-deep software pipeline per group, so neither the index load nor the dependent row load
    // is waited for in the iteration that issues it:
    //   iteration q:  rows of slice q+2 are requested (their indices were requested one iteration
    //                 ago), indices of slice q+4 are requested, slice q (rows requested one iteration
    //                 ago) is stored.
    auto load_idx = [&](int q, int (&ia)[8]) {
      const int tile = cluster_id + (q / num_kb) * n_clusters;
      const int c = (q % num_kb) * 8 + j;
      const int b0 = ((tile / n_tiles_n) * 2 + (int)rank) * BLOCK_M + i0;
      const int table = c < p.C ? s_chunks[c].table : -1;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int b = b0 + 16 * i;
        ia[i] = (table >= 0 && b < p.M) ? __ldg(p.idx + (size_t)b * p.T + table) : -1;
      }
    };
    auto load_rows = [&](int q, const int (&ia)[8], float4 (&out)[8]) {
      const int c = (q % num_kb) * 8 + j;
      const FrFuseChunk ch = s_chunks[c < p.C ? c : 0];
      const int stride4 = ch.stride4_col4 >> 8, col4 = ch.stride4_col4 & 255;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        out[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ia[i] >= 0) {   // K-tail pieces and items past the batch stay zero
          const float4* src = ch.base + (int64_t)ia[i] * stride4 + col4;
          asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=f"(out[i].x), "=f"(out[i].y), "=f"(out[i].z), "=f"(out[i].w)
                       : "l"(src));
        }
      }
    };

    float4 vn[8];
    int ia[8];
    int q = group;
    if (q < total) {
      load_idx(q, ia);
      load_rows(q, ia, vn);
      if (q + 2 < total) load_idx(q + 2, ia);
    }
    for (; q < total; q += 2) {
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = vn[i];
      if (q + 2 < total) load_rows(q + 2, ia, vn);
      if (q + 4 < total) load_idx(q + 4, ia);
      const int s = q % STAGES;
      mbar_wait(&empty_bar[s], ((q / STAGES) & 1) ^ 1, 5, q, g);
      uint8_t* dst = smem + s * L::kStageBytes;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int row = i0 + 16 * i;
        float4 o = v[i];
        o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w);
        // 128B swizzle, as TMA would have laid it out: piece j of row r lives at chunk j ^ (r & 7)
        *reinterpret_cast<float4*>(dst + row * 128 + ((j ^ (row & 7)) << 4)) = o;
      }
      fence_proxy_async();      // generic-proxy stores -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(full_remote + s * 8);
    }
  }
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ---- host side ----------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcLayerCfg {
  int block_n, ctas;   // (pair) tile width and CTAs per tile
};

struct TcState {
  bool auto_tiles = true;   // FR_TC_TILES unset: tile width per layer and batch chosen by pick_block_n()
  PFN_encodeTiled encode = nullptr;
  CUtensorMap w_map[3];
  CUtensorMap w_fuse_map;   // layer-1 weights in 128-row boxes, for the fused lookup + layer 1 kernel
  CUtensorMap w_map64[3];   // 64-row boxes: 128-wide pair tiles (small batches); a CTA's multicast half of a 256-wide tile
  // FR_TC_MCAST=1: 4-CTA multicast clusters for throughput-sized batches.  Off by default: parity-green, but not
  // faster -- halving the bytes every SM ISSUES changed nothing (layer 2, batch 2048: 28.8 us either way; batch
  // 16384: 37.8 against 34.7 us), so the ~33 B/clk an SM takes in is an ingest limit, and 4-CTA clusters pack worse.
  bool mcast = false;
  // FR_TC_ALSU=1: the A operand of throughput-sized batches through cp.async (LSU) instead of TMA.  Off by default:
  // parity-green (bit-identical) but slower -- 0.85 us per K slice whatever the stage count (layer 2, batch 2048:
  // 35.0 against 28.8 us; layer 3: 30.8 against 16.5), i.e. ~19 GB/s per SM of LDGSTS next to the TMA weight feed.
  bool a_lsu = false;
  CUtensorMap w_map128[3];  // every layer in 128-row boxes: the one-launch chain (a CTA's half of an N = 256 MMA)
  // FR_CHAIN=1: the whole MLP as one launch where the chain kernel applies.  Off by default: parity-green
  // (bit-identical to the per-layer kernels) but slower -- with ONE 512-column TMEM stage every epilogue is
  // exposed (3 x 5 us + 1 us per 256-item tile) and the TMA unit, which sustains ~33 B/clk/SM here, carries
  // the stores and the next phase's prefetch in that window: 65 us per tile, 192 against 223 M inferences/s.
  bool chain = false;
  long long* d_prof = nullptr;   // FR_CHAIN_PROF=1: the chain kernel's CTA 0 writes its phase timeline here
  long long* d_tc_prof = nullptr;   // FR_TC_PROF=1 (experiments build): pipeline cycle counters of tc_linear_kernel's first 8 CTAs
  TcLayerCfg cfg[3];
  bool ready = false;
  // cached activation maps keyed by (pointer, K, rows, box rows): 128-row boxes feed the A operand,
  // 32-row boxes are the epilogue's store boxes
  struct AMap {
    const void* ptr;
    int K, rows, box_rows, elt, ks;   // ks = 0: 2-D map, one K slice per box; ks > 1: 3-D map, ks slices per box
    CUtensorMap map;
  };
  CUtensorMap w_map16[3];   // tc_f16: fp16 weights in 128-row x 64-element boxes
  // 3-D views {128 bytes, rows, K slices} of the same weights, boxes of two K slices (tc_pair_kernel KS = 2): 128- and
  // 64-row boxes of the TF32 weights, 128-row boxes of the fp16 copies; w3_ok[k]: layer k's K is a whole number of slices
  CUtensorMap w3_128[3], w3_64[3], w3h_128[3];
  bool w3_ok[3] = {false, false, false}, w3h_ok[3] = {false, false, false};
  std::vector<AMap> a_maps;
  std::mutex mu;
};

fr_status encode_2d(fr_engine* e, TcState* st, CUtensorMap* map, const void* base, int rows, int K, int box_rows,
                    int elt = 4) {
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * elt};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / elt), (cuuint32_t)box_rows};   // one 128-byte swizzle row of K
  const cuuint32_t estr[2] = {1, 1};
  // (L2 promotion 256 B / 128 B / none measured identical to 0.1 us per phase of the chain timeline)
  const CUtensorMapL2promotion l2p = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  CUresult r = st->encode(map, elt == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                          const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2p,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fr_fail(e, FR_ERR_CUDA, "cuTensorMapEncodeTiled(rows=%d K=%d box=%d) failed: CUresult %d", rows, K, box_rows,
                   (int)r);
  return FR_OK;
}

// The same row-major operand [rows][K] seen as a 3-D tensor {one 128-byte K slice, rows, K slices} with byte strides
// {K * elt, 128}: a box {128 B, box_rows, ks} brings `ks` consecutive K slices of box_rows rows in ONE instruction and
// lands as ks consecutive 128B-swizzled slice tiles -- the shared-memory layout the per-slice 2-D boxes produce, at a
// quarter to half the TMA instructions per byte (tools/tma_probe.cu: 86 against 48 B/clk/SM).  K must be a whole
// number of slices (the unit cannot know where a row ends inside a slice); slices past K read as zeros.
fr_status encode_3d(fr_engine* e, TcState* st, CUtensorMap* map, const void* base, int rows, int K, int box_rows, int ks,
                    int elt = 4) {
  const int bk = 128 / elt;
  if (K % bk) return fr_fail(e, FR_ERR_INVALID, "encode_3d: K=%d is not a whole number of %d-element slices", K, bk);
  const cuuint64_t dims[3] = {(cuuint64_t)bk, (cuuint64_t)rows, (cuuint64_t)(K / bk)};
  const cuuint64_t strides[2] = {(cuuint64_t)K * elt, 128};
  const cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)box_rows, (cuuint32_t)ks};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = st->encode(map, elt == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                          const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fr_fail(e, FR_ERR_CUDA, "cuTensorMapEncodeTiled(3-D, rows=%d K=%d box=%d x %d slices) failed: CUresult %d", rows, K,
                   box_rows, ks, (int)r);
  return FR_OK;
}

// A cluster's fixed cost (pipeline fill, last epilogue, teardown: ~4 us) is only hidden behind further
// tiles, so short tiles are ganged: every cluster gets at least knobs.min_kb K-slices of main loop
// (small model layer 1 has 11 per tile -> three tiles per cluster, layer 3 has 16 -> two).
// Measured with 12 workers in flight, small model, batch 2048: 16 -> 214, 32 -> 223, 48 -> 212, 64 -> 201 M
// inferences/s (a longer gang leaves too few clusters per launch to keep 148 SMs busy).

template <int BLOCK_N, int STAGES, int EPI, int CTAS, int PAIRS = 1, bool A_LSU = false, int ELT = 4>
fr_status launch(fr_engine* e, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const TcParams& p,
                 bool pdl_attr, cudaStream_t st) {
  using L = SmemLayout<BLOCK_N, STAGES, CTAS>;
  constexpr int CS = CTAS * PAIRS;
  static_assert(L::kDyn <= 227 * 1024, "tile configuration exceeds the 227 KB shared memory of an SM");
  static_assert(L::kTmemCols == 256 || L::kTmemCols == 512, "TMEM allocation must be a power of two");
  static_assert(L::kNSub == 1 || CTAS == 2, "512-wide tiles are pair tiles");
  auto kern = tc_linear_kernel<BLOCK_N, STAGES, EPI, CTAS, PAIRS, A_LSU, ELT>;
  static std::atomic<uint64_t> attr_done{0};  // bit d: opt-in smem size set on device d for this instantiation
  const uint64_t bit = 1ull << (e->device & 63);
  if (!(attr_done.load() & bit)) {
    FR_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDyn));
    attr_done.fetch_or(bit);
  }
  // persistent: one cluster per tile up to one CTA per SM
  const int n_tiles = (p.M + BLOCK_M * CS - 1) / (BLOCK_M * CS) * (p.N / BLOCK_N);
  int max_clusters = e->sm_count / CS;
  if (e->knobs.max_clusters > 0 && e->knobs.max_clusters < max_clusters) max_clusters = e->knobs.max_clusters;   // test hook
  const int num_kb = (p.K + 128 / ELT - 1) / (128 / ELT);
  // (ganging only pays when the epilogue of one tile runs under the next tile's MMAs: two accumulator stages)
  // an fp16 K slice carries twice the K of a TF32 one in the same bytes: half as many slices make the same gang
  const int min_kb = ELT == 2 ? (e->knobs.min_kb + 1) / 2 : e->knobs.min_kb;
  // ... and only for launches that could not fill half the machine anyway: a launch with that many tiles is spread
  // over all the SMs it can use (layer 3 at batch 16384 is 64 tiles: ganged in twos it ran on 64 of the 148 SMs)
  const int gang = (num_kb >= min_kb || L::kAcc == 1 || p.latency || 2 * n_tiles >= max_clusters)
                       ? 1 : (min_kb + num_kb - 1) / num_kb;   // tiles per cluster wanted
  int n_clusters = (n_tiles + gang - 1) / gang;
  if (n_clusters > max_clusters) n_clusters = max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_clusters * CS, 1, 1);
  cfg.blockDim = dim3(kThreads + (A_LSU ? kLoaderWarps * 32 : 0), 1, 1);
  cfg.dynamicSmemBytes = L::kDyn;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_attr ? 2 : 1;
  FR_CUDA(e, cudaLaunchKernelEx(&cfg, kern, a, b, o, p));
  e->launches++;
  e->tc_last_ctas = n_clusters * CS;
  return FR_OK;
}

fr_status get_a_map(fr_engine* e, TcState* st, const void* ptr, int K, int rows, int box_rows, CUtensorMap* out,
                    int elt = 4, int ks = 0) {
  std::lock_guard<std::mutex> g(st->mu);
  for (const TcState::AMap& m : st->a_maps)
    if (m.ptr == ptr && m.K == K && m.rows == rows && m.box_rows == box_rows && m.elt == elt && m.ks == ks) {
      *out = m.map;
      return FR_OK;
    }
  TcState::AMap m;
  m.ptr = ptr;
  m.K = K;
  m.rows = rows;
  m.box_rows = box_rows;
  m.elt = elt;
  m.ks = ks;
  fr_status s = ks > 1 ? encode_3d(e, st, &m.map, ptr, rows, K, box_rows, ks, elt) : encode_2d(e, st, &m.map, ptr, rows, K, box_rows, elt);
  if (s != FR_OK) return s;
  if (st->a_maps.size() < 1024) st->a_maps.push_back(m);
  *out = m.map;
  return FR_OK;
}

// Launch of the production kernel: persistent, one CTA pair per tile up to one CTA per SM.
template <int BLOCK_N, int STAGES, int EPI, int ELT, int KS>
fr_status launch_pair(fr_engine* e, const CUtensorMap& a, const CUtensorMap& b, const TcParams& p, bool pdl_attr, cudaStream_t st) {
  using L = PairLayout<BLOCK_N, STAGES, KS>;
  static_assert(L::kDyn <= 227 * 1024, "tile configuration exceeds the 227 KB shared memory of an SM");
  static_assert(L::kTmemCols == 256 || L::kTmemCols == 512, "TMEM allocation must be a power of two");
  auto kern = tc_pair_kernel<BLOCK_N, STAGES, EPI, ELT, KS>;
  static std::atomic<uint64_t> attr_done{0};  // bit d: opt-in smem size set on device d for this instantiation
  const uint64_t bit = 1ull << (e->device & 63);
  if (!(attr_done.load() & bit)) {
    FR_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDyn));
    attr_done.fetch_or(bit);
  }
  const int n_tiles = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M) * (p.N / BLOCK_N);
  int max_clusters = e->sm_count / 2;
  if (e->knobs.max_clusters > 0 && e->knobs.max_clusters < max_clusters) max_clusters = e->knobs.max_clusters;   // test hook
  const int num_kb = (p.K + 128 / ELT - 1) / (128 / ELT);
  // Short tiles are ganged (several per cluster) so that a cluster's fixed cost is paid once -- only with two
  // accumulator stages (the epilogue of one tile under the MMAs of the next), never in latency mode, and only for
  // launches that could not fill half the machine anyway; an fp16 K slice carries twice the K of a TF32 one.
  const int min_kb = ELT == 2 ? (e->knobs.min_kb + 1) / 2 : e->knobs.min_kb;
  const int gang = (num_kb >= min_kb || L::kAcc == 1 || p.latency || 2 * n_tiles >= max_clusters)
                       ? 1 : (min_kb + num_kb - 1) / num_kb;
  int n_clusters = (n_tiles + gang - 1) / gang;
  if (n_clusters > max_clusters) n_clusters = max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_clusters * 2, 1, 1);
  cfg.blockDim = dim3(kPairThreads, 1, 1);
  cfg.dynamicSmemBytes = L::kDyn;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_attr ? 2 : 1;
  FR_CUDA(e, cudaLaunchKernelEx(&cfg, kern, a, b, p));
  e->launches++;
  e->tc_last_ctas = n_clusters * 2;
  return FR_OK;
}

// Automatic tile width of a storing layer (pair tiles).  A launch whose 256-wide tiles cannot fill the
// machine anyway is run with 512-wide tiles when its K loop is long: fewer, denser CTAs (33 instead of
// 45 B/clk/SM of operand traffic) leave more SMs to the other workers' kernels -- small model, batch
// 2048, 12 workers: 196 M against 178 M inferences/s, medium model 130 M against 111 M.  Short K loops
// (small model layer 1: 11 slices) keep 256-wide tiles, ganged two per cluster so the epilogue of one
// runs under the MMAs of the next; launches with enough tiles keep 256-wide tiles and two accumulator
// stages (batch 16384: 40 us against 42 us for layer 2).
// Small batches (one or two M tiles) are latency cases -- nothing else is in flight to fill the machine
// -- so they take 128-wide tiles, un-ganged: 2-4x more clusters, each with a quarter of the MMAs
// (medium model, batch 1: 77 -> ~45 us per fr_infer).
constexpr int kLatencyBatch = 512;
constexpr int kWideMinKb = 16;   // K slices from which a 512-wide single-stage tile beats two 256-wide ones
// Is this launch a latency case -- nothing else in flight to fill the machine?  FR_OPT_TILE_HINT says so explicitly;
// by default an engine with at most two worker streams (the reference's THREAD_NUM, cuda_server.c:554-556) is
// taken to serve one batch at a time, and batches of <= 512 items always are.
bool latency_mode(const fr_engine* e, int B) {
  if (e->tile_hint == FR_HINT_LATENCY) return true;
  if (e->tile_hint == FR_HINT_THROUGHPUT) return B <= kLatencyBatch;
  return B <= kLatencyBatch || e->streams.size() <= 2;
}
int pick_block_n(const fr_engine* e, int k, int B) {
  const int N = e->dims[k + 1], K = e->dims[k];
  const int m_tiles = (B + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int tiles256 = m_tiles * (N / 256);
  if (latency_mode(e, B)) {
    // the narrowest tiles that still run as ONE wave of clusters: most SMs busy, least work per SM
    if (B <= kLatencyBatch || m_tiles * (N / 128) <= e->sm_count / 2) return 128;
    if (tiles256 <= e->sm_count / 2) return 256;
  }
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  if (N % 512 == 0 && num_kb >= kWideMinKb && tiles256 < e->sm_count / 2) return 512;
  return 256;
}

}  // namespace

fr_status frtc_prepare(fr_engine* e) {
  TcState* st = static_cast<TcState*>(e->tc_state);
  if (st && st->ready) return FR_OK;
  std::lock_guard<std::mutex> g(e->mu);
  st = static_cast<TcState*>(e->tc_state);
  if (!st) {
    st = new TcState();
    e->tc_state = st;
  }
  if (st->ready) return FR_OK;
  if (!st->encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FR_CUDA(e, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
      return fr_fail(e, FR_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    st->encode = reinterpret_cast<PFN_encodeTiled>(fn);
  }
  if (e->dims[3] != 256)
    return fr_fail(e, FR_ERR_UNSUPPORTED, "TF32 path folds the output layer into layer 3 and needs hidden[2] == 256 "
                   "(got %d)", e->dims[3]);
  st->auto_tiles = !e->knobs.tiles_pinned;
  for (int k = 0; k < 3; k++) st->cfg[k] = {e->knobs.tiles[k], e->knobs.tile_ctas};
  if (!e->h_watch) {
    FR_CUDA(e, cudaHostAlloc(&e->h_watch, 8 * sizeof(int), cudaHostAllocMapped));
    memset(e->h_watch, 0, 8 * sizeof(int));
  }
  int* d_watch = nullptr;
  FR_CUDA(e, cudaHostGetDevicePointer(&d_watch, e->h_watch, 0));
  FR_CUDA(e, cudaMemcpyToSymbol(g_watch, &d_watch, sizeof(d_watch)));
  for (int k = 0; k < 3; k++) {
    const TcLayerCfg c = st->cfg[k];
    if ((c.block_n != 128 && c.block_n != 256 && c.block_n != 512) || (c.ctas != 1 && c.ctas != 2) ||
        (c.block_n == 512 && c.ctas != 2))
      return fr_fail(e, FR_ERR_UNSUPPORTED, "tile N %d / ctas %d not built (N in {128,256,512}, ctas in {1,2}, 512 only as a pair)",
                     c.block_n, c.ctas);
    if (e->dims[k + 1] % c.block_n)
      return fr_fail(e, FR_ERR_UNSUPPORTED, "hidden[%d]=%d not a multiple of tile N %d", k, e->dims[k + 1], c.block_n);
    if (e->dims[k + 1] > kMaxN)
      return fr_fail(e, FR_ERR_UNSUPPORTED, "hidden[%d]=%d wider than the %d-float bias staging area", k, e->dims[k + 1], kMaxN);
    fr_status s = encode_2d(e, st, &st->w_map[k], e->d_Wt[k], e->dims[k + 1], e->dims[k],
                            (c.block_n < 256 ? c.block_n : 256) / c.ctas);
    if (s != FR_OK) return s;
  }
  {
    fr_status s = encode_2d(e, st, &st->w_fuse_map, e->d_Wt[0], e->dims[1], e->dims[0], 128);
    if (s != FR_OK) return s;
    for (int k = 0; k < 3; k++)
      if ((s = encode_2d(e, st, &st->w_map64[k], e->d_Wt[k], e->dims[k + 1], e->dims[k], 64)) != FR_OK) return s;
    st->mcast = e->knobs.mcast;
    st->a_lsu = e->knobs.a_lsu;
    for (int k = 0; k < 3; k++)
      if ((s = encode_2d(e, st, &st->w_map128[k], e->d_Wt[k], e->dims[k + 1], e->dims[k], 128)) != FR_OK) return s;
    for (int k = 0; k < 3; k++) {
      st->w3_ok[k] = e->dims[k] % BLOCK_K == 0;   // (medium model layer 1: K = 880 = 27.5 slices -> one slice per box)
      if (!st->w3_ok[k]) continue;
      if ((s = encode_3d(e, st, &st->w3_128[k], e->d_Wt[k], e->dims[k + 1], e->dims[k], 128, 2)) != FR_OK) return s;
      if ((s = encode_3d(e, st, &st->w3_64[k], e->d_Wt[k], e->dims[k + 1], e->dims[k], 64, 2)) != FR_OK) return s;
    }
    st->chain = e->knobs.chain;
    if (e->knobs.tc_prof && !st->d_tc_prof) {
      FR_CUDA(e, cudaMalloc(&st->d_tc_prof, 8 * 16 * sizeof(long long)));
      FR_CUDA(e, cudaMemset(st->d_tc_prof, 0, 8 * 16 * sizeof(long long)));
    }
    if (e->knobs.chain_prof && !st->d_prof) {
      FR_CUDA(e, cudaMalloc(&st->d_prof, kProfIters * kProfPhases * kProfSlots * sizeof(long long)));
      FR_CUDA(e, cudaMemset(st->d_prof, 0, kProfIters * kProfPhases * kProfSlots * sizeof(long long)));
    }
  }
  st->ready = true;
  return FR_OK;
}

// Tensor maps of the fp16 weight copies (128-row x 64-element boxes); called by the range analysis once it has
// decided for fp16 operands and (re)built d_Wt16 -- the device addresses do not change afterwards.
fr_status frtc_prepare_f16(fr_engine* e) {
  TcState* st = static_cast<TcState*>(e->tc_state);
  if (!st || !st->ready) return fr_fail(e, FR_ERR_STATE, "frtc_prepare_f16 before frtc_prepare");
  for (int k = 0; k < 3; k++) {
    fr_status s = encode_2d(e, st, &st->w_map16[k], e->d_Wt16[k], e->dims[k + 1], e->dims[k], 128, 2);
    if (s != FR_OK) return s;
    st->w3h_ok[k] = e->dims[k] % 64 == 0;   // an fp16 K slice is 64 elements (small model layer 1: 352 = 5.5)
    if (st->w3h_ok[k] && (s = encode_3d(e, st, &st->w3h_128[k], e->d_Wt16[k], e->dims[k + 1], e->dims[k], 128, 2, 2)) != FR_OK) return s;
  }
  return FR_OK;
}

// Debug hook: CTAs (= SMs, one CTA per SM) of the last launch of MLP step k, so a bench can relate a kernel timed
// alone to the SMs it occupied.
extern "C" int frdbg_layer_ctas(const fr_engine* e, int k) { return (e && k >= 0 && k < 4) ? e->tc_layer_ctas[k] : 0; }

#ifdef FR_EXPERIMENTS
// Debug hook (not part of the ABI in include/fleetrec.h; tools/chain_timeline.py binds it by name): the clock64
// stamps the last chain launch's CTA 0 wrote, [kProfIters][kProfPhases][kProfSlots]; returns the number of values.
extern "C" int frdbg_chain_timeline(fr_engine* e, long long* out, int n) {
  TcState* st = e ? static_cast<TcState*>(e->tc_state) : nullptr;
  const int total = kProfIters * kProfPhases * kProfSlots;
  if (!st || !st->d_prof || !out || n < total) return 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return 0;
  if (cudaMemcpy(out, st->d_prof, total * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  return total;
}

#endif

#ifdef FR_EXPERIMENTS
// Debug hook (tools/tc_prof.py): the cycle counters the last tc_linear_kernel launch's first 8 CTAs wrote, [8][16].
extern "C" int frdbg_tc_prof(fr_engine* e, long long* out, int n) {
  TcState* st = e ? static_cast<TcState*>(e->tc_state) : nullptr;
  if (!st || !st->d_tc_prof || !out || n < 128) return 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return 0;
  if (cudaMemcpy(out, st->d_tc_prof, 128 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  cudaMemset(st->d_tc_prof, 0, 128 * sizeof(long long));
  return 128;
}
#endif

void frtc_destroy(fr_engine* e) {
  if (TcState* st = static_cast<TcState*>(e->tc_state)) {
    cudaFree(st->d_prof);
    cudaFree(st->d_tc_prof);
  }
  delete static_cast<TcState*>(e->tc_state);
  e->tc_state = nullptr;
}

bool frtc_can_fuse(const fr_engine* e) {
  const TcState* st = static_cast<const TcState*>(e->tc_state);
  // row pitch and in-row offset of a piece must fit the packed descriptor (dims up to 1020 floats)
  bool dims_ok = true;
  for (const FrTable& t : e->tables) dims_ok = dims_ok && t.dim / 4 < 256;
  return e->fuse_lookup && !fr_tc_f16(e) && st && st->ready && e->world == 1 && e->precision == FR_PREC_TF32 &&
         e->index_format == FR_IDX_I32 &&
         e->table_dtype == FR_TABLE_F32 && e->dims[1] % kFuseN == 0 &&
         e->dims[1] <= kMaxN && e->D / 4 <= kFuseMaxChunks && dims_ok;
}

fr_status frtc_fused_layer1(fr_engine* e, fr_stream_s* s, const int32_t* d_idx, int B) {
  TcState* st = static_cast<TcState*>(e->tc_state);
  constexpr int STAGES = 3;
  using L = FuseLayout<STAGES>;
  static_assert(L::kDyn <= 227 * 1024, "fused tile configuration exceeds the 227 KB shared memory of an SM");
  const bool act = (e->mlp_mode == FR_MLP_BIAS_RELU_SIGMOID);
  CUtensorMap o;
  fr_status r = get_a_map(e, st, s->d_h[0], e->dims[1], B, kStoreBoxRows, &o);
  if (r != FR_OK) return r;
  FuseParams p;
  p.chunks = e->d_fchunks;
  p.idx = d_idx;
  p.bias = act ? e->d_bias[0] : nullptr;
  p.C = e->D / 4;
  p.T = (int)e->tables.size();
  p.M = B;
  p.N = e->dims[1];
  p.relu = act ? 1 : 0;
  auto kern = tc_gather_linear_kernel<STAGES>;
  static std::atomic<uint64_t> attr_done{0};
  const uint64_t bit = 1ull << (e->device & 63);
  if (!(attr_done.load() & bit)) {
    FR_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDyn));
    attr_done.fetch_or(bit);
  }
  const int n_tiles = (B + 2 * BLOCK_M - 1) / (2 * BLOCK_M) * (p.N / kFuseN);
  int max_clusters = e->sm_count / 2;
  if (e->knobs.max_clusters > 0 && e->knobs.max_clusters < max_clusters) max_clusters = e->knobs.max_clusters;
  const int n_clusters = n_tiles < max_clusters ? n_tiles : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_clusters * 2, 1, 1);
  cfg.blockDim = dim3(kFuseThreads, 1, 1);
  cfg.dynamicSmemBytes = L::kDyn;
  cfg.stream = s->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // the weight map of layer 1 must carry 128-row boxes (a CTA's half of an N = 256 MMA)
  FR_CUDA(e, cudaLaunchKernelEx(&cfg, kern, st->w_fuse_map, o, p));
  e->launches++;
  return FR_OK;
}

// The whole MLP as one launch (tc_mlp_chain_kernel) applies to throughput-sized batches of models whose
// storing layers are multiples of 512 wide (small / medium: 1024-512-256, large: 2048-512-256) when the
// tile shapes are not pinned by FR_TC_TILES; batches <= kLatencyBatch keep the per-layer 128-wide tiles,
// which spread one small batch over more SMs.
#ifndef FR_EXPERIMENTS
bool frtc_can_chain(const fr_engine*, int) { return false; }
fr_status frtc_chain(fr_engine* e, fr_stream_s*, const float*, int, float*) {
  return fr_fail(e, FR_ERR_UNSUPPORTED, "the one-launch MLP chain exists in FR_EXPERIMENTS builds only");
}
#else
bool frtc_can_chain(const fr_engine* e, int B) {
  const TcState* st = static_cast<const TcState*>(e->tc_state);
  if (!st || !st->ready || !st->chain || !st->auto_tiles || e->precision != FR_PREC_TF32 || fr_tc_f16(e)) return false;
  if (B <= kLatencyBatch || e->dims[3] != 256) return false;
  for (int k = 1; k <= 2; k++)
    if (e->dims[k] % kChainW || e->dims[k] / kChainW > kChainMaxChunks) return false;
  return e->dims[1] + e->dims[2] + e->dims[3] <= kChainMaxBias;
}

template <int STAGES, bool PROBE = false>
static fr_status launch_chain(fr_engine* e, cudaStream_t stream, const ChainMaps& maps, const ChainParams& p) {
  using L = ChainLayout<STAGES>;
  static_assert(L::kDyn <= 227 * 1024, "chain configuration exceeds the 227 KB shared memory of an SM");
  auto kern = tc_mlp_chain_kernel<STAGES, PROBE>;
  static std::atomic<uint64_t> attr_done{0};
  const uint64_t bit = 1ull << (e->device & 63);
  if (!(attr_done.load() & bit)) {
    FR_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDyn));
    attr_done.fetch_or(bit);
  }
  const int n_tiles = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  int max_clusters = e->sm_count / 2;
  if (e->knobs.max_clusters > 0 && e->knobs.max_clusters < max_clusters) max_clusters = e->knobs.max_clusters;
  const int n_clusters = n_tiles < max_clusters ? n_tiles : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_clusters * 2, 1, 1);
  cfg.blockDim = dim3(kChainThreads + (PROBE ? 128 : 0), 1, 1);
  cfg.dynamicSmemBytes = L::kDyn;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FR_CUDA(e, cudaLaunchKernelEx(&cfg, kern, maps, p));
  e->launches++;
  return FR_OK;
}

fr_status frtc_chain(fr_engine* e, fr_stream_s* s, const float* in, int B, float* d_scores) {
  TcState* st = static_cast<TcState*>(e->tc_state);
  const bool act = (e->mlp_mode == FR_MLP_BIAS_RELU_SIGMOID);
  ChainMaps maps;
  const float* acts[3] = {in, s->d_h[0], s->d_h[1]};
  fr_status r;
  for (int l = 0; l < 3; l++) {
    if ((r = get_a_map(e, st, acts[l], e->dims[l], B, BLOCK_M, &maps.a[l])) != FR_OK) return r;
    if (l > 0 && (r = get_a_map(e, st, acts[l], e->dims[l], B, kStoreBoxRows, &maps.o[l - 1])) != FR_OK) return r;
    maps.w[l] = st->w_map128[l];
  }
  ChainParams p;
  for (int l = 0; l < 3; l++) p.bias[l] = act ? e->d_bias[l] : nullptr;
  p.w4 = e->d_W[3];
  p.b4 = act ? e->d_bias[3] : nullptr;
  p.out = d_scores;
  p.M = B;
  for (int l = 0; l < 4; l++) p.dims[l] = e->dims[l];
  p.relu = act ? 1 : 0;
  p.sigmoid = act ? 1 : 0;
  p.prof = st->d_prof;
  p.probe_src = in;
  p.probe_vecs = (uint32_t)((size_t)B * e->dims[0] / 4);
  return launch_chain<3>(e, s->stream, maps, p);
}

#endif   // FR_EXPERIMENTS

// One launch: layer k (0,1: store tf32-rounded activations into s->d_h[k]; 2: layer 3 with the
// output layer + sigmoid folded in, writes d_scores).
static fr_status frtc_layer_impl(fr_engine* e, fr_stream_s* s, int k, const float* in, int B, float* d_scores,
                                 const FrPeerWait* wait);
fr_status frtc_layer(fr_engine* e, fr_stream_s* s, int k, const float* in, int B, float* d_scores, const FrPeerWait* wait) {
  const fr_status r = frtc_layer_impl(e, s, k, in, B, d_scores, wait);
  if (r == FR_OK) e->tc_layer_ctas[k] = e->tc_last_ctas;
  return r;
}

#ifdef FR_EXPERIMENTS
// tc_linear_kernel and its variants (not instantiated in release builds): 1-CTA tiles, 4-CTA multicast clusters, the
// cp.async A loader; TMA-store epilogue with four warps.
static fr_status frtc_layer_legacy(fr_engine* e, fr_stream_s* s, TcState* st, int k, const float* in, int B, const TcParams& p,
                                   TcLayerCfg c, bool pa) {
  CUtensorMap a, o;
  fr_status r = get_a_map(e, st, in, e->dims[k], B, BLOCK_M, &a);
  if (r != FR_OK) return r;
  o = a;
  if (k < 2 && (r = get_a_map(e, st, s->d_h[k], e->dims[k + 1], B, BLOCK_M, &o)) != FR_OK) return r;
  const CUtensorMap& w = (st->auto_tiles && k < 2 && c.block_n == 128) ? st->w_map64[k] : st->w_map[k];
  cudaStream_t cs = s->stream;
  if (st->mcast && st->auto_tiles && c.ctas == 2 && c.block_n >= 256) {
    if (k < 2 && c.block_n == 512) return launch<512, 3, EPI_STORE, 2, 2>(e, a, st->w_map128[k], o, p, pa, cs);
    if (k < 2) return launch<256, 5, EPI_STORE, 2, 2>(e, a, st->w_map64[k], o, p, pa, cs);
    return launch<256, 5, EPI_DOT, 2, 2>(e, a, st->w_map64[k], o, p, pa, cs);
  }
  if (st->a_lsu && st->auto_tiles && c.ctas == 2 && c.block_n >= 256) {
    if (k < 2 && c.block_n == 512) return launch<512, 3, EPI_STORE, 2, 1, true>(e, a, w, o, p, pa, cs);
    if (k < 2) return launch<256, 5, EPI_STORE, 2, 1, true>(e, a, w, o, p, pa, cs);
    return launch<256, 5, EPI_DOT, 2, 1, true>(e, a, w, o, p, pa, cs);
  }
  if (k < 2) {
    if (c.block_n == 512) return launch<512, 3, EPI_STORE, 2>(e, a, w, o, p, pa, cs);
    if (c.ctas == 2) return c.block_n == 256 ? launch<256, 5, EPI_STORE, 2>(e, a, w, o, p, pa, cs) : launch<128, 7, EPI_STORE, 2>(e, a, w, o, p, pa, cs);
    return c.block_n == 256 ? launch<256, 3, EPI_STORE, 1>(e, a, w, o, p, pa, cs) : launch<128, 5, EPI_STORE, 1>(e, a, w, o, p, pa, cs);
  }
  return c.ctas == 2 ? launch<256, 5, EPI_DOT, 2>(e, a, w, o, p, pa, cs) : launch<256, 3, EPI_DOT, 1>(e, a, w, o, p, pa, cs);
}
#endif

static void fill_params(fr_engine* e, fr_stream_s* s, TcState* st, int k, const float* in, int B, float* d_scores,
                        const FrPeerWait* wait, TcParams* p) {
  const bool act = (e->mlp_mode == FR_MLP_BIAS_RELU_SIGMOID);
  p->a = in;
  p->bias = act ? e->d_bias[k] : nullptr;
  p->M = B;
  p->N = e->dims[k + 1];
  p->K = e->dims[k];
  p->relu = act ? 1 : 0;
  p->sigmoid = act ? 1 : 0;
  p->w4 = e->d_W[3];
  p->b4 = act ? e->d_bias[3] : nullptr;
  p->pdl = e->knobs.pdl_mask ? 1 : 0;
  p->out = d_scores;
  p->out_act = k < 2 ? s->d_h[k] : nullptr;
  p->latency = (st->auto_tiles && latency_mode(e, B)) ? 1 : 0;
  p->wait_flags = wait ? wait->flags : nullptr;
  p->wait_step = wait ? wait->step : nullptr;
  p->wait_n = wait ? wait->world : 0;
  p->wait_err = wait ? wait->err : nullptr;
  p->prof = st->d_tc_prof;
  p->dbg_nostore = e->knobs.dbg_nostore;
}

// fp16 operands (s->f16): `in` and s->d_h[k] hold fp16 [B][dim]; 512-wide tiles where K is long, else 256-wide
static fr_status frtc_layer_f16(fr_engine* e, fr_stream_s* s, int k, const float* in, int B, float* d_scores) {
  TcState* st = static_cast<TcState*>(e->tc_state);
  TcParams p;
  fill_params(e, s, st, k, in, B, d_scores, nullptr, &p);
  const bool ks2 = st->w3h_ok[k];
  CUtensorMap a;
  fr_status r = get_a_map(e, st, in, e->dims[k], B, BLOCK_M, &a, 2, ks2 ? 2 : 0);
  if (r != FR_OK) return r;
  const int N = p.N, num_kb = (p.K + 63) / 64;
  const int tiles256 = (B + 2 * BLOCK_M - 1) / (2 * BLOCK_M) * (N / 256);
  cudaStream_t cs = s->stream;
  if (k == 2) return ks2 ? launch_pair<256, 3, EPI_DOT, 2, 2>(e, a, st->w3h_128[k], p, false, cs)
                         : launch_pair<256, 6, EPI_DOT, 2, 1>(e, a, st->w_map16[k], p, false, cs);
  if (N % 512 == 0 && num_kb >= kWideMinKb && tiles256 < e->sm_count / 2 && !p.latency) {
    if (ks2) {   // (512-wide tiles take one slice per box: three instructions per 48 KB keep up with their eight MMAs)
      r = get_a_map(e, st, in, e->dims[k], B, BLOCK_M, &a, 2, 0);
      if (r != FR_OK) return r;
    }
    return launch_pair<512, 4, EPI_STORE, 2, 1>(e, a, st->w_map16[k], p, false, cs);
  }
  return ks2 ? launch_pair<256, 3, EPI_STORE, 2, 2>(e, a, st->w3h_128[k], p, false, cs)
             : launch_pair<256, 6, EPI_STORE, 2, 1>(e, a, st->w_map16[k], p, false, cs);
}

static fr_status frtc_layer_impl(fr_engine* e, fr_stream_s* s, int k, const float* in, int B, float* d_scores,
                                 const FrPeerWait* wait) {
  if (s->f16) return frtc_layer_f16(e, s, k, in, B, d_scores);   // (single-GPU engines only: never with a wait)
  TcState* st = static_cast<TcState*>(e->tc_state);
  TcParams p;
  fill_params(e, s, st, k, in, B, d_scores, wait, &p);
  const bool pa = (e->knobs.pdl_mask & (k == 0 ? 2 : 1)) != 0;   // may this layer start under the tail of its predecessor
  TcLayerCfg c = st->cfg[k];
  if (st->auto_tiles && k < 2) c.block_n = pick_block_n(e, k, B);
  cudaStream_t cs = s->stream;
#ifdef FR_EXPERIMENTS
  // the older kernel with its variants: 1-CTA tiles (FR_TC_TILES=...,1), multicast clusters, cp.async A loader
  if (c.ctas == 1 || ((st->mcast || st->a_lsu) && !wait && st->auto_tiles && B > kLatencyBatch && c.block_n >= 256))
    return frtc_layer_legacy(e, s, st, k, in, B, p, c, pa);
#else
  if (c.ctas != 2) return fr_fail(e, FR_ERR_UNSUPPORTED, "1-CTA tcgen05 tiles exist in FR_EXPERIMENTS builds only");
#endif
  // rows = B: TMA zero-fills the M tail on load; the epilogue masks its stores by the row count
  const bool ks2 = st->w3_ok[k] && c.block_n != 512;
  CUtensorMap a;
  fr_status r = get_a_map(e, st, in, e->dims[k], B, BLOCK_M, &a, 4, ks2 ? 2 : 0);
  if (r != FR_OK) return r;
  if (k == 2) return ks2 ? launch_pair<256, 3, EPI_DOT, 4, 2>(e, a, st->w3_128[k], p, pa, cs)
                         : launch_pair<256, 6, EPI_DOT, 4, 1>(e, a, st->w_map128[k], p, pa, cs);
  if (c.block_n == 512) return launch_pair<512, 4, EPI_STORE, 4, 1>(e, a, st->w_map128[k], p, pa, cs);
  if (c.block_n == 256) return ks2 ? launch_pair<256, 3, EPI_STORE, 4, 2>(e, a, st->w3_128[k], p, pa, cs)
                                   : launch_pair<256, 6, EPI_STORE, 4, 1>(e, a, st->w_map128[k], p, pa, cs);
  return ks2 ? launch_pair<128, 4, EPI_STORE, 4, 2>(e, a, st->w3_64[k], p, pa, cs)
             : launch_pair<128, 8, EPI_STORE, 4, 1>(e, a, st->w_map64[k], p, pa, cs);
}
