// fr_batcher.cu -- request-driven batch former in front of fr_infer.
//
// The reference hands out FIXED batches: THREAD_NUM workers pull the next batch number off a
// mutex-guarded global counter and block in read() until BATCH_SIZE x INPUT_SIZE floats have
// arrived (cuda_server.c:23-25,406-461), so a batch's latency is unbounded at low load.  This
// front-end forms batches from requests instead:
//
//   producers (any thread)  fr_batcher_submit(): copy n index rows into the OPEN batch's pinned
//                           staging buffer; a full batch is closed and queued at once
//   timer thread            closes the open batch when its oldest request has waited max_delay_us
//   worker threads (one per fr_stream)  take closed batches, fr_infer on pinned buffers, fr_sync,
//                           scatter the scores to every request's output, complete its ticket
//
// Host code only (C++ threads over the public C ABI); the device work is fr_infer's.
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <thread>

#include "fr_common.h"

namespace {
using Clock = std::chrono::steady_clock;

struct Request {
  uint64_t ticket;
  int offset, n;        // rows [offset, offset + n) of the batch
  float* scores_out;
  Clock::time_point t_submit;
};

struct Batch {
  int32_t* idx = nullptr;   // pinned [max_batch][T]
  float* scores = nullptr;  // pinned [max_batch]
  int count = 0;
  bool by_deadline = false;
  Clock::time_point t_first;
  std::vector<Request> reqs;
};
}  // namespace

struct fr_batcher {
  fr_engine* eng = nullptr;
  int max_batch = 0, T = 0;
  std::chrono::microseconds max_delay{0};

  std::mutex mu;
  std::condition_variable cv_work;    // closed batch available / shutdown
  std::condition_variable cv_free;    // a staging buffer returned to the pool
  std::condition_variable cv_timer;   // open batch appeared / changed / shutdown
  std::condition_variable cv_done;    // a ticket completed
  std::vector<Batch*> pool;           // free staging buffers
  Batch* open = nullptr;              // batch being filled
  std::deque<Batch*> closed;          // waiting for a worker
  std::vector<Batch*> all;
  bool stop = false;
  fr_status worker_error = FR_OK;
  std::string worker_error_msg;

  // One ticket per submit.  A request larger than the room left in the open batch is split over several batches,
  // which run concurrently on different workers and may finish in any order: the ticket completes when its LAST
  // OUTSTANDING part retires (not when its last-submitted part does).
  struct Pending {
    int outstanding = 0;   // parts handed to batches and not yet scored
    bool sealed = false;   // submit has handed out every part
  };
  uint64_t next_ticket = 1;
  uint64_t completed_below = 1;       // every ticket < this is complete
  std::vector<uint64_t> done_out_of_order;
  std::map<uint64_t, Pending> pending;

  std::vector<std::thread> workers;
  std::vector<fr_stream> streams;
  std::thread timer;

  // statistics
  int64_t n_batches = 0, n_items = 0, n_deadline = 0, n_requests = 0;
  std::vector<float> latency_us;      // per request, submit -> scores written (ring of the last 65536)
  size_t lat_pos = 0;
};

namespace {

void close_open_locked(fr_batcher* b, bool by_deadline) {
  if (!b->open || b->open->count == 0) return;
  b->open->by_deadline = by_deadline;
  b->closed.push_back(b->open);
  b->open = nullptr;
  b->cv_work.notify_one();
}

void complete_ticket_locked(fr_batcher* b, uint64_t t) {
  if (t == b->completed_below) {
    b->completed_below++;
    // absorb tickets that finished earlier out of order
    bool again = true;
    while (again) {
      again = false;
      for (size_t i = 0; i < b->done_out_of_order.size(); i++)
        if (b->done_out_of_order[i] == b->completed_below) {
          b->completed_below++;
          b->done_out_of_order.erase(b->done_out_of_order.begin() + i);
          again = true;
          break;
        }
    }
  } else {
    b->done_out_of_order.push_back(t);
  }
}

// a part of ticket t has been scored (or, with part_done = false, submit has just sealed the ticket)
void retire_part_locked(fr_batcher* b, uint64_t t, bool part_done) {
  auto it = b->pending.find(t);
  if (it == b->pending.end()) return;
  if (part_done) it->second.outstanding--;
  if (it->second.sealed && it->second.outstanding == 0) {
    b->pending.erase(it);
    complete_ticket_locked(b, t);
  }
}

bool ticket_done_locked(const fr_batcher* b, uint64_t t) {
  if (t < b->completed_below) return true;
  for (uint64_t d : b->done_out_of_order)
    if (d == t) return true;
  return false;
}

void worker_main(fr_batcher* b, int w) {
  cudaSetDevice(b->eng->device);
  for (;;) {
    Batch* bt = nullptr;
    {
      std::unique_lock<std::mutex> lk(b->mu);
      b->cv_work.wait(lk, [&] { return b->stop || !b->closed.empty(); });
      if (b->closed.empty()) return;   // stop requested and nothing left to run
      bt = b->closed.front();
      b->closed.pop_front();
    }
    // full batches replay a CUDA graph per (staging buffer, worker); a deadline-closed batch has a one-off size, which
    // would only fill the graph cache with entries never used again
    fr_status st = fr_infer_opts(b->eng, bt->idx, bt->count, bt->scores, b->streams[w], bt->count != b->max_batch);
    if (st == FR_OK) st = fr_sync(b->eng, b->streams[w]);
    const Clock::time_point now = Clock::now();
    if (st == FR_OK)
      for (const Request& r : bt->reqs) memcpy(r.scores_out, bt->scores + r.offset, (size_t)r.n * sizeof(float));
    {
      std::lock_guard<std::mutex> lk(b->mu);
      if (st != FR_OK && b->worker_error == FR_OK) {
        b->worker_error = st;
        b->worker_error_msg = fr_last_error(b->eng);
      }
      b->n_batches++;
      b->n_items += bt->count;
      if (bt->by_deadline) b->n_deadline++;
      for (const Request& r : bt->reqs) {
        const float us = std::chrono::duration<float, std::micro>(now - r.t_submit).count();
        if (b->latency_us.size() < 65536) b->latency_us.push_back(us);
        else b->latency_us[b->lat_pos++ % 65536] = us;
        retire_part_locked(b, r.ticket, true);
      }
      bt->count = 0;
      bt->reqs.clear();
      b->pool.push_back(bt);
    }
    b->cv_free.notify_all();
    b->cv_done.notify_all();
  }
}

void timer_main(fr_batcher* b) {
  std::unique_lock<std::mutex> lk(b->mu);
  while (!b->stop) {
    if (!b->open || b->open->count == 0) {
      b->cv_timer.wait(lk);
      continue;
    }
    const Clock::time_point deadline = b->open->t_first + b->max_delay;
    const Batch* watched = b->open;
    if (b->cv_timer.wait_until(lk, deadline) == std::cv_status::timeout && b->open == watched && b->open->count > 0 &&
        Clock::now() >= deadline)
      close_open_locked(b, true);
  }
}

}  // namespace

extern "C" fr_status fr_batcher_create(fr_engine* e, const fr_batcher_config* cfg, fr_batcher** out) {
  if (!e || !cfg || !out) return fr_fail(e, FR_ERR_INVALID, "fr_batcher_create: null argument");
  *out = nullptr;
  if (cfg->max_batch <= 0 || cfg->max_batch > e->max_batch)
    return fr_fail(e, FR_ERR_INVALID, "fr_batcher_create: max_batch %d outside (0, engine max_batch %d]", cfg->max_batch,
                   e->max_batch);
  if (cfg->n_workers <= 0 || cfg->n_workers > 64 || cfg->max_delay_us < 0)
    return fr_fail(e, FR_ERR_INVALID, "fr_batcher_create: n_workers %d / max_delay_us %d", cfg->n_workers, cfg->max_delay_us);
  if (e->world > 1) return fr_fail(e, FR_ERR_UNSUPPORTED, "fr_batcher drives fr_infer; a table-sharded engine needs lock-step ranks");
  FR_CUDA(e, cudaSetDevice(e->device));
  fr_batcher* b = new fr_batcher();
  b->eng = e;
  b->max_batch = cfg->max_batch;
  b->T = e->ipr_full;   // int32 words per index row in the engine's index format (FR_OPT_INDEX_FORMAT)
  b->max_delay = std::chrono::microseconds(cfg->max_delay_us);
  const int n_buf = 2 * cfg->n_workers + 1;   // one being filled, one queued and one in flight per worker
  for (int i = 0; i < n_buf; i++) {
    Batch* bt = new Batch();
    b->all.push_back(bt);
    if (cudaHostAlloc(&bt->idx, (size_t)b->max_batch * b->T * sizeof(int32_t), cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc(&bt->scores, (size_t)b->max_batch * sizeof(float), cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      fr_batcher_destroy(b);
      return fr_fail(e, FR_ERR_OOM, "fr_batcher_create: pinned staging buffers (%d x %d rows)", n_buf, b->max_batch);
    }
    b->pool.push_back(bt);
  }
  for (int w = 0; w < cfg->n_workers; w++) {
    fr_stream s = nullptr;
    fr_status st = fr_stream_create(e, &s);
    if (st != FR_OK) {
      fr_batcher_destroy(b);
      return st;
    }
    b->streams.push_back(s);
  }
  for (int w = 0; w < cfg->n_workers; w++) b->workers.emplace_back(worker_main, b, w);
  b->timer = std::thread(timer_main, b);
  *out = b;
  return FR_OK;
}

extern "C" fr_status fr_batcher_submit(fr_batcher* b, const int32_t* idx, int n, float* scores_out, uint64_t* ticket) {
  if (!b) return fr_fail(nullptr, FR_ERR_INVALID, "fr_batcher_submit: null batcher");
  if (!idx || !scores_out || !ticket || n <= 0) return fr_fail(b->eng, FR_ERR_INVALID, "fr_batcher_submit: bad argument");
  int done = 0;
  std::unique_lock<std::mutex> lk(b->mu);
  if (b->worker_error != FR_OK) return fr_fail(b->eng, b->worker_error, "batcher worker failed: %s", b->worker_error_msg.c_str());
  const uint64_t my_ticket = b->next_ticket++;
  b->pending[my_ticket] = fr_batcher::Pending();
  // seal the ticket on every way out: a ticket whose submit failed half-way still completes once the parts already
  // handed out have been scored, so nobody waits on it for ever
  auto seal = [&] {
    b->pending[my_ticket].sealed = true;
    retire_part_locked(b, my_ticket, false);
    b->cv_done.notify_all();
  };
  while (done < n) {
    if (b->stop) {
      seal();
      return fr_fail(b->eng, FR_ERR_STATE, "batcher is shutting down");
    }
    if (!b->open) {
      b->cv_free.wait(lk, [&] { return b->stop || !b->pool.empty() || b->open; });
      if (b->stop) {
        seal();
        return fr_fail(b->eng, FR_ERR_STATE, "batcher is shutting down");
      }
      if (b->open) continue;   // another producer opened one while this thread waited
      b->open = b->pool.back();
      b->pool.pop_back();
      b->open->count = 0;
      b->open->t_first = Clock::now();
      b->cv_timer.notify_one();
    }
    Batch* bt = b->open;
    const int take = std::min(n - done, b->max_batch - bt->count);
    memcpy(bt->idx + (size_t)bt->count * b->T, idx + (size_t)done * b->T, (size_t)take * b->T * sizeof(int32_t));
    bt->reqs.push_back({my_ticket, bt->count, take, scores_out + done, Clock::now()});
    b->pending[my_ticket].outstanding++;
    bt->count += take;
    done += take;
    b->n_requests++;
    if (bt->count == b->max_batch) close_open_locked(b, false);
  }
  seal();
  *ticket = my_ticket;
  return FR_OK;
}

extern "C" fr_status fr_batcher_flush(fr_batcher* b) {
  if (!b) return fr_fail(nullptr, FR_ERR_INVALID, "fr_batcher_flush: null batcher");
  std::lock_guard<std::mutex> lk(b->mu);
  close_open_locked(b, false);
  return FR_OK;
}

extern "C" fr_status fr_batcher_wait(fr_batcher* b, uint64_t ticket) {
  if (!b) return fr_fail(nullptr, FR_ERR_INVALID, "fr_batcher_wait: null batcher");
  std::unique_lock<std::mutex> lk(b->mu);
  if (ticket == 0 || ticket >= b->next_ticket) return fr_fail(b->eng, FR_ERR_INVALID, "fr_batcher_wait: unknown ticket");
  b->cv_done.wait(lk, [&] { return ticket_done_locked(b, ticket) || b->worker_error != FR_OK; });
  if (b->worker_error != FR_OK) return fr_fail(b->eng, b->worker_error, "batcher worker failed: %s", b->worker_error_msg.c_str());
  return FR_OK;
}

extern "C" fr_status fr_batcher_get_stats(fr_batcher* b, fr_batcher_stats* out) {
  if (!b || !out) return fr_fail(b ? b->eng : nullptr, FR_ERR_INVALID, "fr_batcher_get_stats: null argument");
  std::vector<float> lat;
  {
    std::lock_guard<std::mutex> lk(b->mu);
    out->batches = b->n_batches;
    out->items = b->n_items;
    out->requests = b->n_requests;
    out->closed_by_deadline = b->n_deadline;
    lat = b->latency_us;
  }
  out->latency_p50_us = out->latency_p99_us = 0.f;
  if (!lat.empty()) {
    std::sort(lat.begin(), lat.end());
    out->latency_p50_us = lat[lat.size() / 2];
    out->latency_p99_us = lat[std::min(lat.size() - 1, (size_t)(lat.size() * 0.99))];
  }
  return FR_OK;
}

extern "C" void fr_batcher_destroy(fr_batcher* b) {
  if (!b) return;
  {
    std::lock_guard<std::mutex> lk(b->mu);
    close_open_locked(b, false);   // drain: what was submitted is still scored
    b->stop = true;
  }
  b->cv_work.notify_all();
  b->cv_timer.notify_all();
  b->cv_free.notify_all();
  for (std::thread& t : b->workers)
    if (t.joinable()) t.join();
  if (b->timer.joinable()) b->timer.join();
  b->cv_done.notify_all();
  for (fr_stream s : b->streams) fr_stream_destroy(b->eng, s);
  cudaSetDevice(b->eng->device);
  for (Batch* bt : b->all) {
    if (bt->idx) cudaFreeHost(bt->idx);
    if (bt->scores) cudaFreeHost(bt->scores);
    delete bt;
  }
  delete b;
}
