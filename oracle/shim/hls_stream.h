// hls_stream.h -- stand-in for Vitis HLS's C-simulation FIFO.
// TEST INFRASTRUCTURE ONLY (see ap_int.h).  Unbounded queue; reading an empty
// stream is a harness bug and aborts loudly instead of returning garbage.
#ifndef FR_SHIM_HLS_STREAM_H
#define FR_SHIM_HLS_STREAM_H
#include <cstdio>
#include <cstdlib>
#include <deque>
namespace hls {
template <typename T>
class stream {
 public:
  stream() {}
  explicit stream(const char*) {}
  void write(const T& v) { q_.push_back(v); }
  T read() {
    if (q_.empty()) { std::fprintf(stderr, "hls::stream shim: read on empty stream\n"); std::abort(); }
    T v = q_.front();
    q_.pop_front();
    return v;
  }
  void read(T& v) { v = read(); }
  bool read_nb(T& v) { if (q_.empty()) return false; v = read(); return true; }
  bool empty() const { return q_.empty(); }
  bool full() const { return false; }
  size_t size() const { return q_.size(); }
 private:
  stream(const stream&);
  stream& operator=(const stream&);
  std::deque<T> q_;
};
}  // namespace hls
#endif
