// ap_int.h -- minimal stand-in for Xilinx's arbitrary-precision integer header.
//
// TEST INFRASTRUCTURE ONLY.  Written for this repo so that the reference's
// Vitis-HLS lookup kernels (FPGA/kernel/user_krnl/embedding_{47,98,377}_krnl)
// can be compiled with plain g++ into oracle/_ref/ and executed as the parity
// reference.  It implements just the subset those sources use: fixed-width
// unsigned bit vectors, bit-range proxies (x.range(hi,lo) / x(hi,lo)), single
// bit access, and value semantics through one implicit conversion to
// unsigned long long (arithmetic and comparisons then use the built-in
// operators, which is exact for every width <= 64 the kernels do arithmetic on).
#ifndef FR_SHIM_AP_INT_H
#define FR_SHIM_AP_INT_H

#include <stdint.h>
#include <cstring>
#include <iostream>

template <int W> struct ap_uint;

namespace fr_shim {
inline bool get_bit(const uint64_t* w, int i) { return (w[i >> 6] >> (i & 63)) & 1u; }
inline void set_bit(uint64_t* w, int i, bool v) {
  const uint64_t m = 1ull << (i & 63);
  if (v) w[i >> 6] |= m; else w[i >> 6] &= ~m;
}
}  // namespace fr_shim

template <int W>
struct ap_range_ref {
  ap_uint<W>* p;
  int hi, lo;
  ap_range_ref(ap_uint<W>* p_, int hi_, int lo_) : p(p_), hi(hi_), lo(lo_) {}
  int width() const { return hi - lo + 1; }
  bool bit(int i) const;  // bit i of the range (0 = lo)
  ap_range_ref& operator=(unsigned long long v) {
    for (int i = 0; i < width(); i++) set(i, i < 64 ? ((v >> i) & 1ull) : 0);
    return *this;
  }
  template <int W2> ap_range_ref& operator=(const ap_uint<W2>& v);
  template <int W2> ap_range_ref& operator=(const ap_range_ref<W2>& v) {
    // copy through a temporary so that overlapping self-assignment is safe
    bool tmp[4096];
    const int n = width() < v.width() ? width() : v.width();
    for (int i = 0; i < n; i++) tmp[i] = v.bit(i);
    for (int i = 0; i < width(); i++) set(i, i < n ? tmp[i] : false);
    return *this;
  }
  ap_range_ref& operator=(const ap_range_ref& v) { return this->template operator=<W>(v); }
  operator unsigned long long() const {
    unsigned long long r = 0;
    for (int i = 0; i < width() && i < 64; i++) r |= (unsigned long long)bit(i) << i;
    return r;
  }
  void set(int i, bool v);
};

template <int W>
struct ap_bit_ref {
  ap_uint<W>* p;
  int i;
  ap_bit_ref(ap_uint<W>* p_, int i_) : p(p_), i(i_) {}
  ap_bit_ref& operator=(unsigned long long v);
  ap_bit_ref& operator=(const ap_bit_ref& o) { return *this = (unsigned long long)(bool)o; }
  operator bool() const;
};

template <int W>
struct ap_uint {
  typedef char width_check[(W >= 1 && W <= 4096) ? 1 : -1];
  enum { NW = (W + 63) / 64 };
  uint64_t w[NW];

  ap_uint() { std::memset(w, 0, sizeof(w)); }
  void from_signed(long long v) {
    std::memset(w, 0, sizeof(w));
    w[0] = (uint64_t)v;
    if (v < 0) for (int i = 1; i < NW; i++) w[i] = ~0ull;
    trim();
  }
  void from_unsigned(unsigned long long v) {
    std::memset(w, 0, sizeof(w));
    w[0] = v;
    trim();
  }
  ap_uint(bool v) { from_unsigned(v); }
  ap_uint(char v) { from_signed(v); }
  ap_uint(signed char v) { from_signed(v); }
  ap_uint(unsigned char v) { from_unsigned(v); }
  ap_uint(short v) { from_signed(v); }
  ap_uint(unsigned short v) { from_unsigned(v); }
  ap_uint(int v) { from_signed(v); }
  ap_uint(unsigned v) { from_unsigned(v); }
  ap_uint(long v) { from_signed(v); }
  ap_uint(unsigned long v) { from_unsigned(v); }
  ap_uint(long long v) { from_signed(v); }
  ap_uint(unsigned long long v) { from_unsigned(v); }
  template <int W2> ap_uint(const ap_uint<W2>& o) {
    std::memset(w, 0, sizeof(w));
    const int n = NW < ap_uint<W2>::NW ? NW : ap_uint<W2>::NW;
    for (int i = 0; i < n; i++) w[i] = o.w[i];
    trim();
  }
  template <int W2> ap_uint(const ap_range_ref<W2>& r) {
    std::memset(w, 0, sizeof(w));
    for (int i = 0; i < r.width() && i < W; i++) fr_shim::set_bit(w, i, r.bit(i));
  }
  template <int W2> ap_uint(const ap_bit_ref<W2>& b) {
    std::memset(w, 0, sizeof(w));
    w[0] = (bool)b;
  }
  void trim() {
    if (W % 64) w[NW - 1] &= (~0ull) >> (64 - W % 64);
  }
  operator unsigned long long() const { return w[0]; }

  ap_range_ref<W> range(int hi, int lo) { return ap_range_ref<W>(this, hi, lo); }
  ap_range_ref<W> operator()(int hi, int lo) { return ap_range_ref<W>(this, hi, lo); }
  ap_range_ref<W> range(int hi, int lo) const { return ap_range_ref<W>(const_cast<ap_uint*>(this), hi, lo); }
  ap_range_ref<W> operator()(int hi, int lo) const { return ap_range_ref<W>(const_cast<ap_uint*>(this), hi, lo); }
  ap_bit_ref<W> operator[](int i) { return ap_bit_ref<W>(this, i); }
  bool operator[](int i) const { return fr_shim::get_bit(w, i); }
  ap_bit_ref<W> operator()(int i) { return ap_bit_ref<W>(this, i); }
  bool operator()(int i) const { return fr_shim::get_bit(w, i); }

  ap_uint& operator++() { *this = ap_uint((unsigned long long)*this + 1); return *this; }
  ap_uint operator++(int) { ap_uint t = *this; ++*this; return t; }
  ap_uint& operator--() { *this = ap_uint((unsigned long long)*this - 1); return *this; }
  ap_uint operator--(int) { ap_uint t = *this; --*this; return t; }
  template <typename T> ap_uint& operator+=(T v) { *this = ap_uint((unsigned long long)*this + (unsigned long long)v); return *this; }
  template <typename T> ap_uint& operator-=(T v) { *this = ap_uint((unsigned long long)*this - (unsigned long long)v); return *this; }
};

template <int W> bool ap_range_ref<W>::bit(int i) const { return fr_shim::get_bit(p->w, lo + i); }
template <int W> void ap_range_ref<W>::set(int i, bool v) { fr_shim::set_bit(p->w, lo + i, v); }
template <int W> template <int W2>
ap_range_ref<W>& ap_range_ref<W>::operator=(const ap_uint<W2>& v) {
  for (int i = 0; i < width(); i++) set(i, i < W2 ? fr_shim::get_bit(v.w, i) : false);
  return *this;
}
template <int W> ap_bit_ref<W>& ap_bit_ref<W>::operator=(unsigned long long v) {
  fr_shim::set_bit(p->w, i, v & 1ull);
  return *this;
}
template <int W> ap_bit_ref<W>::operator bool() const { return fr_shim::get_bit(p->w, i); }

template <int W> std::ostream& operator<<(std::ostream& os, const ap_uint<W>& v) {
  return os << (unsigned long long)v;
}


#endif
