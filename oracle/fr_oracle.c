/* fr_oracle.c -- CPU restatement of the reference's lookup + concat + MLP.
 * TEST INFRASTRUCTURE ONLY (see fr_oracle.h for who may link this and for the
 * parity-pinning statement).  Plain C11 + OpenMP; build: oracle/Makefile. */
#include "fr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int fro_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* FPGA/host/embedding_47_krnl/host.cpp:66-88 (init_vectors): row 2i <- 1.0f in
 * every lane, row 2i+1 <- 0.0f; embedding_47_krnl.cpp:871-897 does the same for the
 * on-chip tables with the bit pattern 1065353216.  A trailing odd row (rows odd)
 * is left 0, as table_entry_num/2 pairs are written. */
void fro_fill_reference(float* table, int64_t rows, int dim, int64_t debug_rows) {
  memset(table, 0, (size_t)rows * dim * sizeof(float));
  int64_t pairs = rows / 2;
  if (debug_rows > 0 && debug_rows / 2 < pairs) pairs = debug_rows / 2;
  const uint32_t one = 1065353216u;
  for (int64_t i = 0; i < pairs; i++) {
    float* r = table + (2 * i) * dim;
    for (int j = 0; j < dim; j++) memcpy(r + j, &one, 4);
  }
}

uint32_t fro_hash_bits(uint32_t seed, uint32_t table, uint64_t row, uint32_t col) {
  uint64_t z = row * 0x9E3779B97F4A7C15ull + ((uint64_t)table << 40) + ((uint64_t)col << 28) + seed;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  uint32_t h = (uint32_t)(z >> 16);
  uint32_t sign = h & 0x80000000u;
  uint32_t mant = h & 0x007FFFFFu;
  uint32_t expo = 118u + ((h >> 23) & 0xFFu) % 9u; /* |v| in [2^-9, 1): finite normal */
  return sign | (expo << 23) | mant;
}

void fro_fill_hash(float* table, uint32_t seed, int table_id, int64_t rows, int dim) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; r++)
    for (int c = 0; c < dim; c++) {
      uint32_t b = fro_hash_bits(seed, (uint32_t)table_id, (uint64_t)r, (uint32_t)c);
      memcpy(table + r * dim + c, &b, 4);
    }
}

/* embedding_47_krnl.cpp:903-904 */
static const int k_idx_random[32] = {3,  99, 38, 72, 29, 57, 1,  72, 36, 76, 35, 50, 37, 57, 13, 66,
                                     26, 70, 41, 93, 48, 82, 44, 78, 25, 52, 3,  92, 36, 56, 46, 88};

void fro_idx_reference(int32_t* idx, int B, int T) {
  for (int b = 0; b < B; b++)
    for (int t = 0; t < T; t++) idx[(size_t)b * T + t] = k_idx_random[b % 32];
}

/* load_single_embedding_N_tables: base = START + idx*AXI_PADDED (axi units, long),
 * emit AXI_PADDED words (embedding_47_krnl.cpp:925-934).  Here every table has its
 * own base pointer (START folded in) and dims are in floats = 4*AXI_PADDED.
 * Addressing is 64-bit: the largest table is 100 M rows x 32 floats = 12.8 GB. */
void fro_gather(const float* const* tables, const int* dims, const fro_segment* segs, int n_segs,
                const int32_t* idx, int T, int B, int concat_floats, float* out, int threads) {
  if (threads <= 0) threads = fro_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int b = 0; b < B; b++) {
    float* o = out + (size_t)b * concat_floats;
    const int32_t* ib = idx + (size_t)b * T;
    for (int s = 0; s < n_segs; s++) {
      const fro_segment sg = segs[s];
      const float* row = tables[sg.table] + (int64_t)ib[sg.table] * dims[sg.table];
      memcpy(o + sg.dst, row + sg.col, (size_t)sg.len * sizeof(float));
    }
  }
}

/* One layer: Y[B][out] = X[B][in] . W[in][out] (+bias, relu) -- the reference's
 * cublasLtMatmul with col-major W(out x in, ld=out) and X(in x B, ld=in),
 * cuda_server.c:215-217,468-473.  k ascends, fp32 (or double) accumulate. */
static void layer_f32(const float* X, int B, int in, int out, const float* W, const float* bias, int relu,
                      float* Y, int threads) {
  /* register tile: RB rows x CB columns of Y stay in vector registers over the whole k loop
   * (k ascending per output, exactly as a scalar loop would); W is first re-packed into
   * contiguous [in][CB] column panels so the k loop streams memory linearly. */
  enum { RB = 6, CB = 16 };
  const int np = (out + CB - 1) / CB;
  float* Wp = (float*)aligned_alloc(64, (size_t)np * in * CB * sizeof(float));
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int p = 0; p < np; p++)
    for (int k = 0; k < in; k++)
      for (int j = 0; j < CB; j++)
        Wp[((size_t)p * in + k) * CB + j] = (p * CB + j < out) ? W[(size_t)k * out + p * CB + j] : 0.0f;
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int b0 = 0; b0 < B; b0 += RB) {
    const int nb = B - b0 < RB ? B - b0 : RB;
    const float* xr[RB];
    for (int r = 0; r < RB; r++) xr[r] = X + (size_t)(b0 + (r < nb ? r : 0)) * in;
    for (int p = 0; p < np; p++) {
      const int j0 = p * CB;
      const int nc = out - j0 < CB ? out - j0 : CB;
      const float* wp = Wp + (size_t)p * in * CB;
      float acc[RB][CB];
      for (int r = 0; r < RB; r++)
        for (int j = 0; j < CB; j++) acc[r][j] = 0.0f;
      for (int k = 0; k < in; k++) {
        const float* w = wp + (size_t)k * CB;
        for (int r = 0; r < RB; r++) {
          const float a = xr[r][k];
#pragma omp simd
          for (int j = 0; j < CB; j++) acc[r][j] += a * w[j];
        }
      }
      for (int r = 0; r < nb; r++) {
        float* y = Y + (size_t)(b0 + r) * out + j0;
        for (int j = 0; j < nc; j++) {
          float v = acc[r][j] + (bias ? bias[j0 + j] : 0.0f);
          y[j] = (relu && v < 0.0f) ? 0.0f : v;
        }
      }
    }
  }
  free(Wp);
}

static void layer_f64(const float* X, int B, int in, int out, const float* W, const float* bias, int relu,
                      float* Y, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int b = 0; b < B; b++) {
    double* acc = (double*)calloc((size_t)out, sizeof(double));
    for (int k = 0; k < in; k++) {
      const double a = X[(size_t)b * in + k];
      const float* w = W + (size_t)k * out;
      for (int j = 0; j < out; j++) acc[j] += a * (double)w[j];
    }
    for (int j = 0; j < out; j++) {
      double v = acc[j] + (bias ? (double)bias[j] : 0.0);
      if (relu && v < 0.0) v = 0.0;
      Y[(size_t)b * out + j] = (float)v;
    }
    free(acc);
  }
}

void fro_mlp(const float* x, int B, const int* dims, const float* const* W, const float* const* bias,
             int mode, int acc64, float* scores, int threads) {
  if (threads <= 0) threads = fro_max_threads();
  int maxd = 0;
  for (int k = 1; k <= 4; k++)
    if (dims[k] > maxd) maxd = dims[k];
  float* buf0 = (float*)malloc((size_t)B * maxd * sizeof(float));
  float* buf1 = (float*)malloc((size_t)B * maxd * sizeof(float));
  const float* in = x;
  float* outb = buf0;
  for (int k = 0; k < 4; k++) {
    const float* bk = (mode == 1 && bias) ? bias[k] : NULL;
    const int relu = (mode == 1 && k < 3);
    float* dst = (k == 3) ? scores : outb;
    if (acc64)
      layer_f64(in, B, dims[k], dims[k + 1], W[k], bk, relu, dst, threads);
    else
      layer_f32(in, B, dims[k], dims[k + 1], W[k], bk, relu, dst, threads);
    in = dst;
    outb = (outb == buf0) ? buf1 : buf0;
  }
  if (mode == 1)
    for (int b = 0; b < B; b++) scores[b] = 1.0f / (1.0f + expf(-scores[b]));
  free(buf0);
  free(buf1);
}

int64_t fro_merge_index(int64_t iA, int64_t iB, int64_t rowsB) { return iA * rowsB + iB; }

void fro_merge_tables(const float* A, int64_t rowsA, int dimA, const float* Bt, int64_t rowsB, int dimB,
                      float* M) {
  const int dm = dimA + dimB;
#pragma omp parallel for schedule(static)
  for (int64_t a = 0; a < rowsA; a++)
    for (int64_t b = 0; b < rowsB; b++) {
      float* m = M + (a * rowsB + b) * dm;
      memcpy(m, A + a * dimA, (size_t)dimA * sizeof(float));
      memcpy(m + dimA, Bt + b * dimB, (size_t)dimB * sizeof(float));
    }
}
