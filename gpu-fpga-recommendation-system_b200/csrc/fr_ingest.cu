// fr_ingest.cu -- B2-compatible streaming ingest.
//
// The reference's GPU server listens on PORT + i for i < THREAD_NUM, accepts ONE connection per
// port and, per batch, read()s exactly BLOCK_SIZE = BATCH_SIZE * INPUT_FEATURE_LEN * 4 bytes of raw
// little-endian fp32 -- no header, no framing -- before the H2D copy and the four GEMMs; batch
// numbers come off a mutex-guarded counter shared by all connections, up to TOTAL_BATCH_NUM
// (cuda_server.c:360-461,541-556; constant.h:33-41).  Its sender programs
// (multiple_connections_network_client_sender.c:55-100) and the FPGA's sendData()
// (embedding_47_krnl.cpp:45-147) produce exactly that byte stream.  This front-end accepts the same
// stream on the same ports and feeds fr_mlp_only (FR_INGEST_CONCAT), or, for senders that ship
// indices instead of gathered vectors, fr_infer (FR_INGEST_INDICES: BATCH * n_tables int32).
// Differences by design: two staging buffers per connection, so the receive of block k+1 overlaps
// the device work of block k, and a buffer is only reused after an event says its batch is done
// (the reference reuses its single pinned buffer without waiting); errors
// are returned, nothing calls exit().
#include <arpa/inet.h>
#include <errno.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <thread>

#include "fr_common.h"

namespace {
struct Conn {
  int listen_fd = -1, fd = -1;
  fr_stream stream = nullptr;
  char* in[2] = {nullptr, nullptr};      // pinned staging blocks
  float* out[2] = {nullptr, nullptr};    // pinned score blocks
  cudaEvent_t done[2] = {nullptr, nullptr};
  int64_t batches = 0, bytes = 0;
  std::vector<float> last_scores;
  int64_t last_batch_no = -1;
  fr_status status = FR_OK;
  std::string error;
  std::thread thread;
};
}  // namespace

struct fr_ingest {
  fr_engine* eng = nullptr;
  fr_ingest_config cfg;
  size_t block_bytes = 0;
  std::vector<int64_t> rows;        // rows of every table: bound of the FR_INGEST_INDICES validation
  std::vector<Conn> conns;
  std::mutex mu;
  int64_t global_batch_count = 0;   // cuda_server.c:23
  std::chrono::steady_clock::time_point t0, t1;
  bool joined = false;
};

namespace {

void conn_fail(Conn& c, fr_status st, const std::string& msg) {
  c.status = st;
  c.error = msg;
}

// scores of a finished batch go to the caller's sink and to the connection's "last scores"
void retire(fr_ingest* g, int ci, int slot, int64_t k, int64_t batch_no) {
  Conn& c = g->conns[ci];
  if (g->cfg.scores_out && k < g->cfg.max_batches_per_conn)
    memcpy(g->cfg.scores_out + ((size_t)ci * g->cfg.max_batches_per_conn + k) * g->cfg.batch, c.out[slot],
           (size_t)g->cfg.batch * sizeof(float));
  std::lock_guard<std::mutex> lk(g->mu);
  c.last_scores.assign(c.out[slot], c.out[slot] + g->cfg.batch);
  c.last_batch_no = batch_no;
}

void conn_main(fr_ingest* g, int ci) {
  Conn& c = g->conns[ci];
  cudaSetDevice(g->eng->device);
  sockaddr_in peer;
  socklen_t plen = sizeof(peer);
  c.fd = accept(c.listen_fd, reinterpret_cast<sockaddr*>(&peer), &plen);   // cuda_server.c:393
  if (c.fd < 0) return conn_fail(c, FR_ERR_STATE, std::string("accept: ") + strerror(errno));
  cudaStream_t cs = static_cast<cudaStream_t>(fr_stream_cuda(c.stream));
  int64_t pending_no[2] = {-1, -1}, pending_k[2] = {-1, -1};
  for (int64_t k = 0;; k++) {
    int64_t batch_no;
    {
      std::lock_guard<std::mutex> lk(g->mu);   // cuda_server.c:408-417
      if (g->cfg.total_batches > 0 && g->global_batch_count >= g->cfg.total_batches) break;
      batch_no = g->global_batch_count++;
    }
    const int slot = (int)(k & 1);
    if (pending_no[slot] >= 0) {   // the batch that used this staging block two blocks ago must be done
      if (cudaEventSynchronize(c.done[slot]) != cudaSuccess) return conn_fail(c, FR_ERR_CUDA, "cudaEventSynchronize failed");
      retire(g, ci, slot, pending_k[slot], pending_no[slot]);
      pending_no[slot] = -1;
    }
    size_t got = 0;
    while (got < g->block_bytes) {   // cuda_server.c:426-450: exactly BLOCK_SIZE bytes per batch
      const ssize_t r = read(c.fd, c.in[slot] + got, g->block_bytes - got);
      if (r < 0 && errno == EINTR) continue;
      if (r < 0) return conn_fail(c, FR_ERR_STATE, std::string("read: ") + strerror(errno));
      if (r == 0) break;             // sender closed
      got += (size_t)r;
    }
    if (got < g->block_bytes) {
      {
        std::lock_guard<std::mutex> lk(g->mu);   // hand the unused batch number back
        if (batch_no == g->global_batch_count - 1) g->global_batch_count--;
      }
      if (got != 0) conn_fail(c, FR_ERR_STATE, "sender closed inside a block (" + std::to_string(got) + " of " +
                                                   std::to_string(g->block_bytes) + " bytes)");
      break;
    }
    if (g->cfg.payload == FR_INGEST_INDICES) {
      // the block came off a socket: an index outside its table would be an out-of-bounds device read
      const int32_t* ix = reinterpret_cast<const int32_t*>(c.in[slot]);
      const size_t T = g->rows.size();
      size_t bad = (size_t)-1;
      for (size_t i = 0, n = (size_t)g->cfg.batch * T; i < n; i++)
        if ((uint64_t)(int64_t)ix[i] >= (uint64_t)g->rows[i % T]) {
          bad = i;
          break;
        }
      if (bad != (size_t)-1) {
        conn_fail(c, FR_ERR_INVALID, "block " + std::to_string(batch_no) + ": index " + std::to_string(ix[bad]) + " of table " +
                                         std::to_string(bad % T) + " (item " + std::to_string(bad / T) + ") is outside its " +
                                         std::to_string(g->rows[bad % T]) + " rows");
        break;   // the blocks already in flight are still retired below
      }
    }
    fr_status st = g->cfg.payload == FR_INGEST_CONCAT
                       ? fr_mlp_only(g->eng, reinterpret_cast<const float*>(c.in[slot]), g->cfg.batch, c.out[slot], c.stream)
                       : fr_infer(g->eng, reinterpret_cast<const int32_t*>(c.in[slot]), g->cfg.batch, c.out[slot], c.stream);
    if (st != FR_OK) return conn_fail(c, st, fr_last_error(g->eng));
    if (cudaEventRecord(c.done[slot], cs) != cudaSuccess) return conn_fail(c, FR_ERR_CUDA, "cudaEventRecord failed");
    pending_no[slot] = batch_no;
    pending_k[slot] = k;
    c.batches++;
    c.bytes += (int64_t)g->block_bytes;
  }
  // drain in batch order
  int order[2] = {0, 1};
  if (pending_k[0] > pending_k[1]) { order[0] = 1; order[1] = 0; }
  for (int i = 0; i < 2; i++) {
    const int slot = order[i];
    if (pending_no[slot] < 0) continue;
    if (cudaEventSynchronize(c.done[slot]) != cudaSuccess) return conn_fail(c, FR_ERR_CUDA, "cudaEventSynchronize failed");
    retire(g, ci, slot, pending_k[slot], pending_no[slot]);
  }
}

}  // namespace

extern "C" fr_status fr_ingest_start(fr_engine* e, const fr_ingest_config* cfg, fr_ingest** out) {
  if (!e || !cfg || !out) return fr_fail(e, FR_ERR_INVALID, "fr_ingest_start: null argument");
  *out = nullptr;
  if (cfg->n_conn <= 0 || cfg->n_conn > 64 || cfg->batch <= 0 || cfg->batch > e->max_batch || cfg->base_port <= 0 ||
      cfg->base_port + cfg->n_conn > 65535 || (cfg->payload != FR_INGEST_CONCAT && cfg->payload != FR_INGEST_INDICES) ||
      cfg->total_batches < 0 || (cfg->scores_out && cfg->max_batches_per_conn <= 0))
    return fr_fail(e, FR_ERR_INVALID, "fr_ingest_start: bad configuration (n_conn %d, batch %d, base_port %d, payload %d)",
                   cfg->n_conn, cfg->batch, cfg->base_port, cfg->payload);
  if (e->world > 1) return fr_fail(e, FR_ERR_UNSUPPORTED, "fr_ingest drives fr_infer / fr_mlp_only; not for a table-sharded engine");
  if (cfg->payload == FR_INGEST_INDICES && e->index_format != FR_IDX_I32)
    return fr_fail(e, FR_ERR_STATE, "FR_INGEST_INDICES blocks are int32 rows (the reference's index stream): set FR_OPT_INDEX_FORMAT back to FR_IDX_I32");
  FR_CUDA(e, cudaSetDevice(e->device));
  fr_ingest* g = new fr_ingest();
  g->eng = e;
  g->cfg = *cfg;
  g->block_bytes = (size_t)cfg->batch * (cfg->payload == FR_INGEST_CONCAT ? (size_t)e->D * sizeof(float)
                                                                          : e->tables.size() * sizeof(int32_t));
  for (const FrTable& t : e->tables) g->rows.push_back(t.rows);
  g->conns.resize(cfg->n_conn);
  auto fail = [&](fr_status st, const std::string& msg) {
    fr_ingest_destroy(g);
    return fr_fail(e, st, "fr_ingest_start: %s", msg.c_str());
  };
  for (int i = 0; i < cfg->n_conn; i++) {
    Conn& c = g->conns[i];
    fr_status st = fr_stream_create(e, &c.stream);
    if (st != FR_OK) {
      fr_ingest_destroy(g);
      return st;
    }
    for (int b = 0; b < 2; b++) {
      if (cudaHostAlloc(&c.in[b], g->block_bytes, cudaHostAllocDefault) != cudaSuccess ||
          cudaHostAlloc(&c.out[b], (size_t)cfg->batch * sizeof(float), cudaHostAllocDefault) != cudaSuccess ||
          cudaEventCreateWithFlags(&c.done[b], cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        return fail(FR_ERR_OOM, "pinned staging blocks");
      }
    }
    c.listen_fd = socket(AF_INET, SOCK_STREAM, 0);
    if (c.listen_fd < 0) return fail(FR_ERR_STATE, std::string("socket: ") + strerror(errno));
    int opt = 1;
    setsockopt(c.listen_fd, SOL_SOCKET, SO_REUSEADDR, &opt, sizeof(opt));   // cuda_server.c:372
    sockaddr_in addr;
    memset(&addr, 0, sizeof(addr));
    addr.sin_family = AF_INET;
    addr.sin_addr.s_addr = cfg->listen_any ? htonl(INADDR_ANY) : htonl(INADDR_LOOPBACK);   // :379 binds INADDR_ANY
    addr.sin_port = htons((uint16_t)(cfg->base_port + i));                                    // :380, :541
    if (bind(c.listen_fd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) < 0)
      return fail(FR_ERR_STATE, "bind port " + std::to_string(cfg->base_port + i) + ": " + strerror(errno));
    if (listen(c.listen_fd, 3) < 0) return fail(FR_ERR_STATE, std::string("listen: ") + strerror(errno));
  }
  // every port is bound and listening before the call returns, so senders may connect right away
  g->t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < cfg->n_conn; i++) g->conns[i].thread = std::thread(conn_main, g, i);
  *out = g;
  return FR_OK;
}

extern "C" fr_status fr_ingest_wait(fr_ingest* g, fr_ingest_stats* stats) {
  if (!g) return fr_fail(nullptr, FR_ERR_INVALID, "fr_ingest_wait: null handle");
  for (Conn& c : g->conns)
    if (c.thread.joinable()) c.thread.join();
  if (!g->joined) {
    g->t1 = std::chrono::steady_clock::now();
    g->joined = true;
  }
  fr_status st = FR_OK;
  std::string msg;
  int64_t batches = 0, bytes = 0;
  int connected = 0;
  for (size_t i = 0; i < g->conns.size(); i++) {
    const Conn& c = g->conns[i];
    batches += c.batches;
    bytes += c.bytes;
    connected += c.fd >= 0;
    if (c.status != FR_OK && st == FR_OK) {
      st = c.status;
      msg = "connection " + std::to_string(i) + ": " + c.error;
    }
  }
  if (stats) {
    stats->batches = batches;
    stats->bytes = bytes;
    stats->connections = connected;
    stats->seconds = std::chrono::duration<double>(g->t1 - g->t0).count();
  }
  if (st != FR_OK) return fr_fail(g->eng, st, "fr_ingest: %s", msg.c_str());
  return FR_OK;
}

extern "C" fr_status fr_ingest_last_scores(fr_ingest* g, int conn, float* scores, int64_t* batch_no) {
  if (!g || !scores || conn < 0 || conn >= (int)g->conns.size()) return fr_fail(g ? g->eng : nullptr, FR_ERR_INVALID, "fr_ingest_last_scores: bad argument");
  std::lock_guard<std::mutex> lk(g->mu);
  const Conn& c = g->conns[conn];
  if (c.last_batch_no < 0) return fr_fail(g->eng, FR_ERR_STATE, "connection %d has not finished a batch", conn);
  memcpy(scores, c.last_scores.data(), c.last_scores.size() * sizeof(float));
  if (batch_no) *batch_no = c.last_batch_no;
  return FR_OK;
}

extern "C" void fr_ingest_destroy(fr_ingest* g) {
  if (!g) return;
  for (Conn& c : g->conns) {   // unblock accept() / read() of threads still running
    if (c.listen_fd >= 0) shutdown(c.listen_fd, SHUT_RDWR);
    if (c.fd >= 0) shutdown(c.fd, SHUT_RDWR);
  }
  for (Conn& c : g->conns)
    if (c.thread.joinable()) c.thread.join();
  cudaSetDevice(g->eng->device);
  for (Conn& c : g->conns) {
    if (c.fd >= 0) close(c.fd);
    if (c.listen_fd >= 0) close(c.listen_fd);
    if (c.stream) {
      fr_sync(g->eng, c.stream);
      fr_stream_destroy(g->eng, c.stream);
    }
    for (int b = 0; b < 2; b++) {
      if (c.in[b]) cudaFreeHost(c.in[b]);
      if (c.out[b]) cudaFreeHost(c.out[b]);
      if (c.done[b]) cudaEventDestroy(c.done[b]);
    }
  }
  delete g;
}
