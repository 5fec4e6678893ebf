// fleetrec_host.cpp -- C++ host program over the C ABI (include/fleetrec.h).
//
// Plays both reference host roles on one B200:
//   FPGA/host/embedding_47_krnl/host.cpp  : build the table images, hand them to the device,
//                                           start the lookup for `batch_num` batches
//   GPU/.../cuda_server.c main()+thread_consume() : THREAD_NUM workers, one stream each, pulling
//                                           batch numbers off a mutex-guarded global counter
//                                           (cuda_server.c:23-25,406-417,547-556), printing the
//                                           first outputs at the end (cuda_server.c:499-502)
// The TCP hop between the two (sendData / read()) does not exist: concat vectors stay in HBM.
//
// usage: fleetrec_host [model=small] [batch=2048] [total_batches=2048] [threads=4]
//                      [fill=reference|hash] [mode=linear|sigmoid] [prec=tf32|fp32] [row_cap=0]
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <mutex>
#include <string>
#include <vector>

#include "fleetrec.h"

namespace {

const int kIdxRandom[32] = {3,  99, 38, 72, 29, 57, 1,  72, 36, 76, 35, 50, 37, 57, 13, 66,
                            26, 70, 41, 93, 48, 82, 44, 78, 25, 52, 3,  92, 36, 56, 46, 88};  // embedding_47_krnl.cpp:903

struct Shared {
  fr_engine* eng;
  const fr_model_desc* desc;
  int batch, total_batches;
  bool reference_idx;
  std::mutex mu;
  int next_batch = 0;  // global_batch_count
  std::vector<int64_t> rows;
};

struct WorkerInfo {  // CUDA_thread_info
  Shared* sh;
  int id;
  std::vector<float> last_scores;
  int batches_done = 0;
  int status = 0;
};

void* thread_consume(void* vp) {
  WorkerInfo* w = static_cast<WorkerInfo*>(vp);
  Shared* sh = w->sh;
  const int T = sh->desc->n_tables, B = sh->batch;
  fr_stream st = nullptr;
  if (fr_stream_create(sh->eng, &st) != FR_OK) {
    fprintf(stderr, "worker %d: %s\n", w->id, fr_last_error(sh->eng));
    w->status = -1;
    return nullptr;
  }
  std::vector<int32_t> idx((size_t)B * T);
  w->last_scores.assign(B, 0.f);
  uint64_t lcg = 0x9E3779B97F4A7C15ull * (uint64_t)(w->id + 1);
  while (true) {
    int my_batch;
    {
      std::lock_guard<std::mutex> g(sh->mu);
      if (sh->next_batch >= sh->total_batches) break;
      my_batch = sh->next_batch++;
    }
    (void)my_batch;
    for (int b = 0; b < B; b++)
      for (int t = 0; t < T; t++) {
        if (sh->reference_idx) {
          idx[(size_t)b * T + t] = kIdxRandom[b % 32];
        } else {
          lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
          idx[(size_t)b * T + t] = (int32_t)((lcg >> 33) % (uint64_t)sh->rows[t]);
        }
      }
    if (fr_infer(sh->eng, idx.data(), B, w->last_scores.data(), st) != FR_OK || fr_sync(sh->eng, st) != FR_OK) {
      fprintf(stderr, "worker %d: %s\n", w->id, fr_last_error(sh->eng));
      w->status = -1;
      break;
    }
    w->batches_done++;
  }
  fr_stream_destroy(sh->eng, st);
  return nullptr;
}

}  // namespace

int main(int argc, char** argv) {
  const std::string model = argc > 1 ? argv[1] : "small";
  const int batch = argc > 2 ? atoi(argv[2]) : 2048;
  const int total = argc > 3 ? atoi(argv[3]) : 2048;
  const int threads = argc > 4 ? atoi(argv[4]) : 4;
  const std::string fill = argc > 5 ? argv[5] : "reference";
  const std::string mode = argc > 6 ? argv[6] : "linear";
  const std::string prec = argc > 7 ? argv[7] : "tf32";
  const int64_t row_cap = argc > 8 ? atoll(argv[8]) : 0;

  fr_model_desc desc;
  if (fr_model_builtin(model.c_str(), &desc) != FR_OK) {
    fprintf(stderr, "%s\n", fr_last_error(nullptr));
    return 1;
  }
  desc.mlp_mode = mode == "linear" ? FR_MLP_LINEAR : FR_MLP_BIAS_RELU_SIGMOID;
  desc.precision = prec == "fp32" ? FR_PREC_FP32 : FR_PREC_TF32;
  desc.max_batch = batch;
  const int dev = 0;
  fr_engine* eng = nullptr;
  if (fr_create(&desc, 1, &dev, &eng) != FR_OK) {
    fprintf(stderr, "fr_create: %s\n", fr_last_error(nullptr));
    return 1;
  }
  Shared sh;
  sh.eng = eng;
  sh.desc = &desc;
  sh.batch = batch;
  sh.total_batches = total;
  sh.reference_idx = (fill == "reference");
  for (int t = 0; t < desc.n_tables; t++) {
    int64_t rows = desc.tables[t].rows;
    if (row_cap > 0 && rows > row_cap) {
      rows = row_cap;
      if (fr_set_table_rows(eng, t, rows) != FR_OK) { fprintf(stderr, "%s\n", fr_last_error(eng)); return 1; }
    }
    sh.rows.push_back(rows);
    const fr_status s = sh.reference_idx ? fr_fill_table_reference(eng, t, 0) : fr_fill_table_hash(eng, t, 0x5EED);
    if (s != FR_OK) { fprintf(stderr, "table %d: %s\n", t, fr_last_error(eng)); return 1; }
  }
  // init_array(w, n, 1.0f) for every layer (cuda_server.c:152-160)
  int dims[5] = {desc.concat_floats, desc.hidden[0], desc.hidden[1], desc.hidden[2], desc.hidden[3]};
  for (int k = 0; k < 4; k++) {
    std::vector<float> w((size_t)dims[k] * dims[k + 1], mode == "linear" ? 1.0f : 1.0f / (float)dims[k]);
    if (fr_load_mlp(eng, k, w.data(), nullptr) != FR_OK) { fprintf(stderr, "%s\n", fr_last_error(eng)); return 1; }
  }
  printf("model %s: %d tables, %d floats/item, %.3f GB in HBM; batch %d x %d batches on %d workers\n", model.c_str(),
         desc.n_tables, desc.concat_floats, fr_table_bytes(eng) / 1e9, batch, total, threads);

  std::vector<WorkerInfo> info(threads);
  std::vector<pthread_t> th(threads);
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < threads; i++) {
    info[i].sh = &sh;
    info[i].id = i;
    pthread_create(&th[i], nullptr, thread_consume, &info[i]);
  }
  int rc = 0, done = 0;
  for (int i = 0; i < threads; i++) {
    pthread_join(th[i], nullptr);
    rc |= info[i].status;
    done += info[i].batches_done;
  }
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (int i = 0; i < 5 && i < batch; i++) printf("r_layer_out[%d] = %f\n", i, info[0].last_scores[i]);
  printf("%d batches, %.3f s, %.0f inferences/s (host wall clock, indices generated on the host), kernels launched %lld\n",
         done, sec, (double)done * batch / sec, (long long)fr_launch_count(eng));
  fr_destroy(eng);
  return rc ? 1 : 0;
}
