/* fr_oracle.h -- CPU restatement of the reference's lookup + concat + MLP.
 *
 * TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py; never by the product path.
 *
 * Parity pinning: the lookup/concat half is pinned by executing the reference's
 * own gather code (oracle/_ref, built by oracle/Makefile from the sources in
 * /root/reference against the ap_uint/hls::stream shim in oracle/shim/) and by the
 * committed golden vectors in tests/golden/ made from it; the MLP half is pinned
 * by the reference's README known-answer values (all-ones KAT).  Per-table
 * distinct indices, the Cartesian-merge remap and bias/ReLU/sigmoid are not
 * exercised by any reference code: parity unpinned for those.
 */
#ifndef FR_ORACLE_H
#define FR_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int dst, table, col, len; } fro_segment;

/* host.cpp:66-88 init_vectors / embedding_47_krnl.cpp:871-897 init_plram_t_1_table:
 * even rows all 1.0f, odd rows all 0.0f.  debug_rows > 0: only the first
 * debug_rows rows are written, the rest stay 0 (the `#define DEBUG` path). */
void fro_fill_reference(float* table, int64_t rows, int dim, int64_t debug_rows);
/* Position-encoding fill: float bits = hash(seed, table, row, col) with the
 * exponent forced into [1, 254] (finite, normal, never NaN/Inf/denormal). */
uint32_t fro_hash_bits(uint32_t seed, uint32_t table, uint64_t row, uint32_t col);
void fro_fill_hash(float* table, uint32_t seed, int table_id, int64_t rows, int dim);

/* embedding_47_krnl.cpp:899-914 load_access_idx: item j of every FPGA batch of
 * 32 reads row idx_random[j % 32] in EVERY table. */
void fro_idx_reference(int32_t* idx, int B, int T);

/* Rows L1-L4: per item, per segment, copy table[idx][col..col+len) to dst.
 * load_single_embedding_*_tables (embedding_47_krnl.cpp:916-935) +
 * gather_embeddings (47: 1097-1217, 98: 1331-1605, 377: 1665-1873). */
void fro_gather(const float* const* tables, const int* dims, const fro_segment* segs, int n_segs,
                const int32_t* idx, int T, int B, int concat_floats, float* out, int threads);

/* cuda_server.c:468-491: R1 = W1.X ... out = W4.R3, fp32.  W[k] row-major
 * [in_k][out_k]; mode 0 = LINEAR (no bias/activation), 1 = bias+ReLU, final
 * sigmoid.  acc64 != 0 accumulates in double (tolerance analysis).  dims =
 * {in, h1, h2, h3, 1}. */
void fro_mlp(const float* x, int B, const int* dims, const float* const* W, const float* const* bias,
             int mode, int acc64, float* scores, int threads);

/* MicroRec Cartesian merge: M[iA*rowsB + iB] = A[iA] || B[iB]. */
int64_t fro_merge_index(int64_t iA, int64_t iB, int64_t rowsB);
void fro_merge_tables(const float* A, int64_t rowsA, int dimA, const float* B, int64_t rowsB, int dimB,
                      float* M);

int fro_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
