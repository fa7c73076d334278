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

// ---- libdivsufsort's adjacent API (c-sources/divsufsort.c:372-405, utils.c:111-156) and the LCP array ----

// divbwt: -> (transformed string, primary index)
inline std::pair<std::vector<uint8_t>, int32_t> bwt(const uint8_t *text, size_t n) {
  if (n >= 0x7fffffffull) throw std::logic_error("text too large, should not exceed 2147483646 bytes");
  std::vector<uint8_t> u(n);
  static const uint8_t dummy = 0;
  static uint8_t dummy_u = 0;
  const int32_t rc = gsa_divbwt(n ? text : &dummy, n ? u.data() : &dummy_u, nullptr, (int32_t)n);
  if (rc < 0) throw std::runtime_error(std::string("divbwt returned ") + std::to_string(rc) + ": " + gsa_last_error());
  return {std::move(u), rc};
}

// inverse_bw_transform: (transformed string, primary index) -> text
inline std::vector<uint8_t> inverse_bwt(const uint8_t *u, size_t n, int32_t primary_index) {
  std::vector<uint8_t> t(n);
  static const uint8_t dummy = 0;
  static uint8_t dummy_t = 0;
  const int32_t rc = gsa_inverse_bw_transform(n ? u : &dummy, n ? t.data() : &dummy_t, nullptr, (int32_t)n, primary_index);
  if (rc != 0) throw std::runtime_error(std::string("inverse_bw_transform returned ") + std::to_string(rc) + ": " + gsa_last_error());
  return t;
}

// LCP[0] = 0, LCP[j] = lcp(suffix sa[j-1], suffix sa[j])
inline std::vector<int32_t> lcp(const uint8_t *text, size_t n, const int32_t *sa, int device = 0) {
  std::vector<int32_t> out(n, 0);
  if (n == 0) return out;
  const int32_t rc = gsa_lcp(text, sa, out.data(), (int32_t)n, device);
  if (rc != 0) throw std::runtime_error(std::string("gsa_lcp returned ") + std::to_string(rc) + ": " + gsa_last_error());
  return out;
}

}  // namespace divsufsort
