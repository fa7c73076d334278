// divsufsort.hpp -- C++ mirror of the reference crate `divsufsort`
// (crates/divsufsort/src/lib.rs:20-29; twin crates/cdivsufsort/src/lib.rs:9-30).
#pragma once
#include "sacabase.hpp"

namespace divsufsort {

// sort_in_place(text, sa): panics (throws) when lengths differ (divsufsort.rs:4-8) or the
// text is too large for i32 indices (divsufsort.rs:9-13); asserts the C return code is 0
// (cdivsufsort lib.rs:22).
inline void sort_in_place(const uint8_t *text, size_t text_len, int32_t *sa, size_t sa_len, int device = -1,
                          gsa_build_stats *stats = nullptr) {
  if (text_len != sa_len) throw std::logic_error("text and suffix array should have same len");
  if (text_len >= 0x7fffffffull) throw std::logic_error("text too large, should not exceed 2147483646 bytes");
  static const uint8_t dummy_t = 0;
  static int32_t dummy_sa = 0;
  const uint8_t *t = text_len ? text : &dummy_t;
  int32_t *s = sa_len ? sa : &dummy_sa;
  const int32_t rc = (device < 0 && !stats) ? gsa_divsufsort(t, s, (int32_t)text_len)
                                            : gsa_divsufsort_ex(t, s, (int32_t)text_len, device < 0 ? 0 : device, stats);
  if (rc != 0) throw std::runtime_error(std::string("divsufsort returned ") + std::to_string(rc) + ": " + gsa_last_error());
}

// sort(text) -> SuffixArray (lib.rs:25-29)
inline sacabase::SuffixArray sort(const uint8_t *text, size_t text_len, int device = -1, gsa_build_stats *stats = nullptr) {
  std::vector<int32_t> sa(text_len, 0);
  sort_in_place(text, text_len, sa.data(), sa.size(), device, stats);
  return sacabase::SuffixArray(text, text_len, std::move(sa), device < 0 ? 0 : device);
}
inline sacabase::SuffixArray sort(const std::string &text, int device = -1) {
  return sort(reinterpret_cast<const uint8_t *>(text.data()), text.size(), device);
}

}  // namespace divsufsort
