// fr_mlp_simt.cu -- FP32 (FFMA) MLP layers: the FR_PREC_FP32 arithmetic mode.
//
// Same arithmetic class as the reference's cublasLtMatmul with CUBLAS_COMPUTE_32F
// (cuda_server.c:211,468-491): FP32 operands, FP32 accumulate, k ascending inside
// a thread.  Layouts are the reference's, untouched: X row-major [B][K] (col-major
// K x B, ld=K), W row-major [K][N] (col-major N x K, ld=N), Y row-major [B][N].
// Bias + ReLU are fused into the store; the 1-wide output layer is a warp dot
// product with the sigmoid fused (frk_final_dot), also used by the TF32 path.
#include "fr_common.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8;  // 256 threads, 8x8 outputs each

// Y[M][N] = act(X[M][K] . W[K][N] + bias);  N % 128 == 0, K % 16 == 0, M arbitrary.
__global__ void __launch_bounds__(256) sgemm_bias_act_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                             const float* __restrict__ bias, float* __restrict__ Y,
                                                             int M, int K, int N, int relu) {
  __shared__ __align__(16) float As[2][BK][BM + 4];  // transposed: As[k][m]
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 thread grid, each 8 x 8

  // global -> smem mapping: A tile 128 x 16 floats = 512 float4, 2 per thread; B tile 16 x 128 = 512 float4
  const int a_row = tid / 4, a_k4 = tid % 4;        // rows a_row and a_row+64
  const int b_k = tid / 32, b_n4 = tid % 32;        // k rows b_k and b_k+8

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto load_g = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int r = m0 + a_row + h * 64;
      ra[h] = (r < M) ? *reinterpret_cast<const float4*>(X + (size_t)r * K + k0 + a_k4 * 4)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[h] = *reinterpret_cast<const float4*>(W + (size_t)(k0 + b_k + h * 8) * N + n0 + b_n4 * 4);
    }
  };
  auto store_s = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int r = a_row + h * 64;
      As[buf][a_k4 * 4 + 0][r] = ra[h].x;
      As[buf][a_k4 * 4 + 1][r] = ra[h].y;
      As[buf][a_k4 * 4 + 2][r] = ra[h].z;
      As[buf][a_k4 * 4 + 3][r] = ra[h].w;
      *reinterpret_cast<float4*>(&Bs[buf][b_k + h * 8][b_n4 * 4]) = rb[h];
    }
  };

  load_g(0);
  store_s(0);
  __syncthreads();
  const int nk = K / BK;
  for (int kt = 0; kt < nk; kt++) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_g((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float a[TM], b[TN];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_s(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < TM; i++) {
    const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; jh++) {
      const int c = n0 + jh * 64 + tx * 4;
      float4 o = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      if (bias) {
        const float4 bv = *reinterpret_cast<const float4*>(bias + c);
        o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
      }
      if (relu) {
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
      }
      *reinterpret_cast<float4*>(Y + (size_t)r * N + c) = o;
    }
  }
}

// scores[b] = act(sum_k H[b][k]*w[k] + bias[0]); one warp per item, K % 4 == 0.
__global__ void __launch_bounds__(256) final_dot_kernel(const float* __restrict__ H, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ scores,
                                                        int B, int K, int sigmoid) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= B) return;
  const float4* h4 = reinterpret_cast<const float4*>(H + (size_t)warp * K);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  float s = 0.f;
  for (int k = lane; k < K / 4; k += 32) {
    const float4 a = h4[k], b = __ldg(w4 + k);
    s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    if (bias) s += bias[0];
    scores[warp] = sigmoid ? 1.f / (1.f + __expf(-s)) : s;
  }
}

}  // namespace

fr_status frk_sgemm_bias_act(fr_engine* e, const float* X, const float* W, const float* bias, float* Y, int B, int K,
                             int N, bool relu, cudaStream_t st) {
  if (N % BN || K % BK)
    return fr_fail(e, FR_ERR_UNSUPPORTED, "FP32 GEMM needs N %% 128 == 0 and K %% 16 == 0 (N=%d K=%d)", N, K);
  if (B == 0) return FR_OK;
  dim3 grid(N / BN, (B + BM - 1) / BM);
  sgemm_bias_act_kernel<<<grid, 256, 0, st>>>(X, W, bias, Y, B, K, N, relu ? 1 : 0);
  e->launches++;
  FR_CUDA(e, cudaGetLastError());
  return FR_OK;
}

fr_status frk_final_dot(fr_engine* e, const float* H, const float* w, const float* bias, float* scores, int B, int K,
                        bool sigmoid, cudaStream_t st) {
  if (B == 0) return FR_OK;
  final_dot_kernel<<<(B * 32 + 255) / 256, 256, 0, st>>>(H, w, bias, scores, B, K, sigmoid ? 1 : 0);
  e->launches++;
  FR_CUDA(e, cudaGetLastError());
  return FR_OK;
}
