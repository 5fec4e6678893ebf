// ref_harness.cpp -- runs the REFERENCE's own lookup kernel on the CPU.
//
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile three times (one per
// model) into oracle/_ref/libref_{small,medium,large_half}.so.  The reference
// source is #included from where it lies under /root/reference (never copied):
//   -DKRNL_CPP="<.../embedding_47_krnl.cpp>"  -DKRNL_TOP=embedding_47_krnl  -DN_PLRAM=17
// against the ap_uint / hls::stream stand-ins in oracle/shim/.  HLS DATAFLOW
// becomes sequential C simulation: every stage runs to completion into an
// unbounded FIFO, so the TCP offload engine is emulated by pre-loading its
// response streams (the same trick the reference's own testbenches use,
// FPGA/kernel/user_krnl/scatter_krnl/src/hls/test_scatter.cpp:36-167).
#include <pthread.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include KRNL_CPP

#define S4(A, i) A[i], A[i + 1], A[i + 2], A[i + 3]
#define H28(A) S4(A, 0), S4(A, 4), S4(A, 8), S4(A, 12), S4(A, 16), S4(A, 20), S4(A, 24)
#if N_PLRAM == 17
#define PALL(A) S4(A, 0), S4(A, 4), S4(A, 8), S4(A, 12), A[16]
#elif N_PLRAM == 19
#define PALL(A) S4(A, 0), S4(A, 4), S4(A, 8), S4(A, 12), A[16], A[17], A[18]
#elif N_PLRAM == 11
#define PALL(A) S4(A, 0), S4(A, 4), A[8], A[9], A[10]
#else
#error "N_PLRAM must be 17, 19 or 11"
#endif

namespace {

struct TopArgs {
  const float* const* banks;
  int batch_num, useConn, pkgWordCount;
  float* out;
  long cap, written;
};

void* run_top_thread(void* vp) {
  TopArgs* a = static_cast<TopArgs*>(vp);
  hls::stream<pkt512> udp_rx, udp_tx, tcp_rx_data, tcp_tx_data;
  hls::stream<pkt256> udp_rx_meta, udp_tx_meta;
  hls::stream<pkt16> listen_port, close_conn, rx_meta;
  hls::stream<pkt8> port_status;
  hls::stream<pkt64> open_conn, tx_status;
  hls::stream<pkt32> open_status, read_pkg, tx_meta;
  hls::stream<pkt128> notification;

  // TOE emulation: every openConnection succeeds ...
  for (int i = 0; i < a->useConn; i++) {
    pkt32 st;
    st.data(15, 0) = 100 + i;
    st.data(16, 16) = 1;
    open_status.write(st);
  }
  // ... and every tx request is granted in full, error = 0 (sendData, cpp:78-126).
  const long total_bytes = (long)a->batch_num * BATCH_SIZE * INPUT_SIZE_AXI_512 * 64;
  const long pkt_bytes = (long)a->pkgWordCount * 64;
  for (long i = 0; i < total_bytes / pkt_bytes + 2; i++) {
    pkt64 rsp;
    rsp.data(15, 0) = 100 + (int)(i % a->useConn);
    rsp.data(31, 16) = (int)pkt_bytes;
    rsp.data(61, 32) = 0xffff;
    rsp.data(63, 62) = 0;
    tx_status.write(rsp);
  }
  const axi_t* const* b = reinterpret_cast<const axi_t* const*>(a->banks);
  typedef char axi_is_16_bytes[sizeof(axi_t) == 16 ? 1 : -1];  // 4 packed fp32, as in HBM
  KRNL_TOP(H28(b), b[28], b[29], udp_rx, udp_tx, udp_rx_meta, udp_tx_meta, listen_port, port_status, open_conn,
           open_status, close_conn, notification, read_pkg, rx_meta, tcp_rx_data, tx_meta, tcp_tx_data, tx_status,
           a->useConn, a->pkgWordCount, 5001, 0x0A01D46E, a->batch_num);
  long n = 0;
  while (!tcp_tx_data.empty()) {
    pkt512 wd = tcp_tx_data.read();
    if (n + 16 <= a->cap) std::memcpy(a->out + n, wd.data.w, 64);  // raw little-endian fp32, as on the wire
    n += 16;
  }
  a->written = n;
  return NULL;
}

}  // namespace

extern "C" {

void ref_info(int* input_size, int* n_plram, int* fpga_batch) {
  *input_size = INPUT_SIZE;
  *n_plram = N_PLRAM;
  *fpga_batch = BATCH_SIZE;
}

// banks: 30 pointers (HBM0..27, DDR0, DDR1) to bank images laid out as host.cpp
// does (table rows back to back at ADDR_AXI_*).  out receives the tx byte stream
// as floats; *written = floats the kernel actually emitted.
int ref_run_top(const float* const* banks, int batch_num, int useConn, int pkgWordCount, float* out, long cap,
                long* written) {
  TopArgs a;
  a.banks = banks; a.batch_num = batch_num; a.useConn = useConn; a.pkgWordCount = pkgWordCount;
  a.out = out; a.cap = cap; a.written = 0;
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, 1ul << 30);  // the kernel keeps its on-chip tables on the stack
  pthread_t th;
  if (pthread_create(&th, &attr, run_top_thread, &a) != 0) return -1;
  pthread_join(th, NULL);
  *written = a.written;
  return 0;
}

// Stream-level run of the reference's gather_embeddings(): stream s (order
// HBM0..27, DDR0, DDR1, PLRAM0..) carries words[s] axi words per item, float f of
// item i on stream s tagged  i*65536 + s*256 + f  (exact in fp32).  out:
// [n_items][INPUT_SIZE].  n_items must be a multiple of the FPGA batch (32).
int ref_run_gather_tagged(int n_items, const int* words, float* out) {
  if (n_items % BATCH_SIZE) return -1;
  hls::stream<axi_t> H[28], D[2], P[N_PLRAM];
  hls::stream<network_t> net;
  for (int s = 0; s < 30 + N_PLRAM; s++) {
    hls::stream<axi_t>& st = s < 28 ? H[s] : (s < 30 ? D[s - 28] : P[s - 30]);
    for (int i = 0; i < n_items; i++)
      for (int k = 0; k < words[s]; k++) {
        float f[4];
        for (int l = 0; l < 4; l++) f[l] = (float)(i * 65536 + s * 256 + k * 4 + l);
        axi_t wd;
        std::memcpy(wd.w, f, 16);
        st.write(wd);
      }
  }
  gather_embeddings(H28(H), D[0], D[1], PALL(P), net, n_items / BATCH_SIZE);
  long n = 0;
  while (!net.empty()) {
    network_t wd = net.read();
    std::memcpy(out + n, wd.w, 64);
    n += 16;
  }
  return n == (long)n_items * INPUT_SIZE ? 0 : -2;
}

}  // extern "C"
