// fr_precision.cu -- range analysis behind FR_OPT_F16_OPERANDS = FR_F16_GUARDED.
//
// The tcgen05 MLP can run on fp16 operands and activations (kind::f16, fp32 accumulate): fp16 carries the
// 11-bit significand TF32 keeps, in half the bytes, so the feed-bound GEMMs run ~1.4x faster.  What fp16
// does not have is fp32's exponent range (TF32 does): above 65504 an operand is infinite -- the reference's
// own all-ones known answer passes 352 x 1024 at layer 2 (GPU/README.md:9) -- and below 2^-14 it loses
// significand bits.  So the path is chosen per engine from bounds computed ON THE DATA THAT IS LOADED:
//
//   |x_j|  <= max |table(j)|                       (x = lookup output; every element is a table element)
//   |h1_o| <= sum_j |W1[j][o]| * ub_x[j] + |b1_o|   (ReLU only shrinks; LINEAR mode has no bias: same bound)
//   |h2_o| <= sum_i |W2[i][o]| * ub_h1[i] + |b2_o|  (H3 never leaves TMEM: fp32)
//
// fp16 is used iff every bound (with 0.1 % slack for the roundings) stays below 60000, every weight is
// finite in fp16, the smallest non-zero table magnitude is fp16-normal, and the share of any unit's weight
// mass that fp16 represents inexactly (sub-normal weights) is below 1e-6.  fr_mlp_only / fr_layer_only take
// caller data the engine cannot bound and always run TF32.
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <math.h>
#include <string.h>

#include "fr_common.h"

namespace {

// max |x| and min non-zero |x| of a table image, as uint bit patterns (order-preserving for non-negative
// floats; Inf / NaN sort above every finite value, so they fail the bound).  out[0] = max, out[1] = min.
__global__ void table_range_kernel(const void* __restrict__ t, int64_t n, int dt, unsigned* __restrict__ out) {
  unsigned mx = 0u, mn = 0xFFFFFFFFu;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned b;
    if (dt == FR_TABLE_F32) b = reinterpret_cast<const unsigned*>(t)[i] & 0x7FFFFFFFu;
    else if (dt == FR_TABLE_F16) b = __float_as_uint(__half2float(__ushort_as_half(reinterpret_cast<const uint16_t*>(t)[i]))) & 0x7FFFFFFFu;
    else if (dt == FR_TABLE_FP8) {
      const __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)reinterpret_cast<const uint8_t*>(t)[i], __NV_E4M3);
      b = __float_as_uint(__half2float(*reinterpret_cast<const __half*>(&h))) & 0x7FFFFFFFu;
    } else b = ((unsigned)reinterpret_cast<const uint16_t*>(t)[i] << 16) & 0x7FFFFFFFu;
    mx = max(mx, b);
    if (b) mn = min(mn, b);
  }
  for (int o = 16; o; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, mx);
    atomicMin(out + 1, mn);
  }
}

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// One thread per output unit o of a layer W[in][out] (the reference layout, cuda_server.c:215):
//   ub_out[o]  = 1.001 * sum_i |w_io| ub_in[i] + |bias[o]|     w = the TF32-rounded weight the kernels multiply
//   inexact[o] = sum_i |fp16(w_io) - w_io| ub_in[i] / sum_i |w_io| ub_in[i]     (Inf when a weight overflows fp16)
__global__ void abs_matvec_kernel(const float* __restrict__ W, const float* __restrict__ ub_in, const float* __restrict__ bias,
                                  int in, int out, float* __restrict__ ub_out, float* __restrict__ inexact) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= out) return;
  float mass = 0.f, lost = 0.f;
  for (int i = 0; i < in; i++) {
    const float w = rna_tf32(W[(size_t)i * out + o]);
    const float u = ub_in[i];
    mass = fmaf(fabsf(w), u, mass);
    lost = fmaf(fabsf(__half2float(__float2half_rn(w)) - w), u, lost);
  }
  ub_out[o] = 1.001f * mass + (bias ? fabsf(bias[o]) : 0.f);
  inexact[o] = mass > 0.f ? lost / mass : (lost > 0.f ? INFINITY : 0.f);
}

}  // namespace

// Range of one resident table (cached until the table is written again).
static fr_status table_range(fr_engine* e, FrTable& tb, unsigned* d_tmp) {
  if (tb.range_valid) return FR_OK;
  cudaStream_t st = e->default_stream->stream;
  const unsigned init[2] = {0u, 0xFFFFFFFFu};
  FR_CUDA(e, cudaMemcpyAsync(d_tmp, init, sizeof(init), cudaMemcpyHostToDevice, st));
  const int64_t n = tb.rows * tb.dim;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)e->sm_count * 16) blocks = (int64_t)e->sm_count * 16;
  table_range_kernel<<<(int)blocks, 256, 0, st>>>(tb.d, n, e->table_dtype, d_tmp);
  e->launches++;
  FR_CUDA(e, cudaGetLastError());
  unsigned h[2];
  FR_CUDA(e, cudaMemcpyAsync(h, d_tmp, sizeof(h), cudaMemcpyDeviceToHost, st));
  FR_CUDA(e, cudaStreamSynchronize(st));
  memcpy(&tb.maxabs, &h[0], 4);
  if (h[1] == 0xFFFFFFFFu) tb.minabs = INFINITY;   // all zeros
  else memcpy(&tb.minabs, &h[1], 4);
  tb.range_valid = true;
  return FR_OK;
}

// Decide e->tc_f16 (called with e->mu held, tables and layers loaded).  Leaves the numbers in e->f16_bounds.
fr_status fr_f16_analyse(fr_engine* e) {
  e->f16_dirty = false;
  const bool was = e->tc_f16;
  e->tc_f16 = false;
  for (int i = 0; i < 5; i++) e->f16_bounds[i] = 0.f;
  if (e->f16_mode != FR_F16_GUARDED || e->precision != FR_PREC_TF32 || e->world != 1) return FR_OK;
  FR_CUDA(e, cudaSetDevice(e->device));
  cudaStream_t st = e->default_stream->stream;
  const int D = e->D;
  int widest = D;
  for (int k = 1; k <= 3; k++) widest = widest > e->dims[k] ? widest : e->dims[k];
  float* d_buf = nullptr;   // ub_in | ub_out | inexact | 2 x unsigned
  FR_CUDA(e, cudaMalloc(&d_buf, (size_t)(3 * widest + 4) * sizeof(float)));
  float *d_in = d_buf, *d_out = d_buf + widest, *d_lost = d_buf + 2 * widest;
  unsigned* d_tmp = reinterpret_cast<unsigned*>(d_buf + 3 * widest);
  fr_status rc = FR_OK;
  std::vector<float> h(widest), hl(widest);
  float min_nonzero = INFINITY, max_lost = 0.f;
  bool ok = true;
  do {
    // concat bound: every element of x is an element of its table
    for (FrTable& tb : e->tables)
      if ((rc = table_range(e, tb, d_tmp)) != FR_OK) break;
    if (rc != FR_OK) break;
    for (const fr_segment_desc& sg : e->segs) {
      const FrTable& tb = e->tables[sg.table];
      for (int j = 0; j < sg.len; j++) h[sg.dst + j] = tb.maxabs;
      if (tb.minabs < min_nonzero) min_nonzero = tb.minabs;
    }
    float bx = 0.f;
    for (int j = 0; j < D; j++) bx = h[j] > bx || h[j] != h[j] ? h[j] : bx;
    e->f16_bounds[0] = bx;
    e->f16_bounds[3] = min_nonzero;
    cudaError_t ce = cudaMemcpyAsync(d_in, h.data(), (size_t)D * sizeof(float), cudaMemcpyHostToDevice, st);
    // hidden layers 1 and 2 are stored as fp16; layer 3's weights must be representable, its output stays fp32
    for (int k = 0; k < 3 && ce == cudaSuccess; k++) {
      const int in = e->dims[k], out = e->dims[k + 1];
      abs_matvec_kernel<<<(out + 127) / 128, 128, 0, st>>>(e->d_W[k], d_in, e->mlp_mode == FR_MLP_BIAS_RELU_SIGMOID ? e->d_bias[k] : nullptr,
                                                          in, out, d_out, d_lost);
      e->launches++;
      if ((ce = cudaGetLastError()) != cudaSuccess) break;
      if ((ce = cudaMemcpyAsync(h.data(), d_out, (size_t)out * sizeof(float), cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
      if ((ce = cudaMemcpyAsync(hl.data(), d_lost, (size_t)out * sizeof(float), cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
      if ((ce = cudaStreamSynchronize(st)) != cudaSuccess) break;
      float b = 0.f;
      for (int o = 0; o < out; o++) {
        if (!(h[o] <= b)) b = h[o];                 // (NaN propagates)
        if (!(hl[o] <= max_lost)) max_lost = hl[o];
      }
      if (k < 2) e->f16_bounds[1 + k] = b;
      float* t = d_in; d_in = d_out; d_out = t;     // this layer's bound feeds the next
    }
    if (ce != cudaSuccess) {
      rc = fr_fail(e, FR_ERR_CUDA, "fp16 range analysis: %s", cudaGetErrorString(ce));
      break;
    }
    e->f16_bounds[4] = max_lost;
    for (int i = 0; i < 3; i++) ok = ok && e->f16_bounds[i] <= 60000.f;       // false for NaN too
    ok = ok && min_nonzero >= 6.103515625e-05f && max_lost <= 1e-6f;
  } while (0);
  cudaFree(d_buf);
  if (rc != FR_OK) return rc;
  if (ok) {
    // fp16 copies of the TF32-rounded K-major weights: exact (11-bit significands, in range -- just proven)
    for (int k = 0; k < 3; k++) {
      const size_t nw = (size_t)e->dims[k] * e->dims[k + 1];
      if (!e->d_Wt16[k]) FR_CUDA(e, cudaMalloc(&e->d_Wt16[k], nw * 2));
      if ((rc = frk_to_f16(e, e->d_Wt[k], e->d_Wt16[k], (int64_t)nw, st)) != FR_OK) return rc;
    }
    FR_CUDA(e, cudaStreamSynchronize(st));
    if ((rc = frtc_prepare_f16(e)) != FR_OK) return rc;
  }
  e->tc_f16 = ok;
  (void)was;
  return FR_OK;
}
