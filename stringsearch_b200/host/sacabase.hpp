// sacabase.hpp -- C++ mirror of the reference crate `sacabase`
// (crates/sacabase/src/lib.rs) on top of the C ABI in include/gsa.h.
//
// Same names, argument meaning and failure behaviour as the Rust API; where Rust
// panics this throws std::logic_error / std::out_of_range with the reference's message.
// Searches run on the GPU (a device-resident copy of text + sa is created lazily and
// kept for the lifetime of the SuffixArray).  There is no CPU search path.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gsa.h"

namespace sacabase {

using Bytes = std::pair<const uint8_t *, size_t>;  // &[u8]

inline void gsa_check(int32_t rc, const char *where) {
  if (rc == GSA_OK) return;
  std::string msg = std::string(where) + ": rc=" + std::to_string(rc) + " " + gsa_last_error();
  if (rc == GSA_ENOMEM) throw std::bad_alloc();
  throw std::runtime_error(msg);
}

// lib.rs:4-21
struct LongestCommonSubstring {
  const uint8_t *text = nullptr;
  size_t text_len = 0;
  size_t start = 0;
  size_t len = 0;
  Bytes as_bytes() const { return {text + start, len}; }
  std::string debug() const { return "T[" + std::to_string(start) + ".." + std::to_string(start + len) + "]"; }
};

// lib.rs:26-35.  Plain host helper exposed by the reference API (not a search path).
inline size_t common_prefix_len(Bytes a, Bytes b) {
  const size_t n = std::min(a.second, b.second);
  for (size_t i = 0; i < n; ++i)
    if (a.first[i] != b.first[i]) return i;
  return n;
}

// lib.rs:102-123
struct NotSorted : std::runtime_error {
  size_t i, j;
  NotSorted(size_t i_, size_t j_)
      : std::runtime_error("invariant doesn't hold: suf(SA(" + std::to_string(i_) + ")) < suf(SA(" + std::to_string(j_) + "))"),
        i(i_), j(j_) {}
};

// A flat batch of needles: needle q is bytes[off[q] .. off[q+1]).
struct NeedleBatch {
  std::vector<uint8_t> bytes;
  std::vector<uint64_t> off{0};
  void push(const uint8_t *p, size_t n) {
    bytes.insert(bytes.end(), p, p + n);
    off.push_back(bytes.size());
  }
  void push(const std::string &s) { push(reinterpret_cast<const uint8_t *>(s.data()), s.size()); }
  uint64_t size() const { return off.size() - 1; }
};

// trait StringIndex (lib.rs:160-163)
struct StringIndex {
  virtual ~StringIndex() = default;
  virtual LongestCommonSubstring longest_substring_match(const uint8_t *needle, size_t needle_len) const = 0;
  LongestCommonSubstring longest_substring_match(const std::string &needle) const {
    return longest_substring_match(reinterpret_cast<const uint8_t *>(needle.data()), needle.size());
  }
};

// SuffixArray<'a, i32> (lib.rs:152-197): owns `sa`, borrows `text`.
class SuffixArray : public StringIndex {
 public:
  // SuffixArray::new (lib.rs:170-172)
  SuffixArray(const uint8_t *text, size_t text_len, std::vector<int32_t> sa, int device = 0)
      : text_(text), text_len_(text_len), sa_(std::move(sa)), device_(device) {}
  SuffixArray(SuffixArray &&) = default;
  SuffixArray &operator=(SuffixArray &&) = default;

  // into_parts (lib.rs:175-177)
  std::pair<Bytes, std::vector<int32_t>> into_parts() && { return {{text_, text_len_}, std::move(sa_)}; }
  Bytes text() const { return {text_, text_len_}; }  // lib.rs:185-187
  const std::vector<int32_t> &sa() const { return sa_; }

  // verify (lib.rs:127-149,180-182): throws NotSorted.  O(n) on the GPU.
  void verify() const {
    if (text_len_ == 0) throw std::logic_error("attempt to subtract with overflow");  // lib.rs:143
    int64_t bad = -1;
    const int32_t rc = gsa_index_verify(handle(), &bad);
    if (rc == 1) throw NotSorted((size_t)bad, (size_t)bad + 1);
    gsa_check(rc, "gsa_index_verify");
  }

  using StringIndex::longest_substring_match;
  // lib.rs:39-99,190-197
  LongestCommonSubstring longest_substring_match(const uint8_t *needle, size_t needle_len) const override {
    NeedleBatch b;
    b.push(needle, needle_len);
    return longest_substring_match_batch(b)[0];
  }

  std::vector<LongestCommonSubstring> longest_substring_match_batch(const NeedleBatch &b) const {
    if (sa_.empty()) throw std::out_of_range("index out of bounds: the len is 0 but the index is 0");  // lib.rs:89-91
    std::vector<uint64_t> st(b.size());
    std::vector<uint32_t> ln(b.size());
    gsa_check(gsa_lsm_batch(handle(), b.bytes.data(), b.off.data(), b.size(), st.data(), ln.data()), "gsa_lsm_batch");
    std::vector<LongestCommonSubstring> out(b.size());
    for (size_t q = 0; q < out.size(); ++q) out[q] = {text_, text_len_, (size_t)st[q], (size_t)ln[q]};
    return out;
  }

  // libdivsufsort sa_search semantics (c-sources/utils.c:258-325): occurrences are
  // sa()[left .. left+count).
  struct Range { int32_t left, count; };
  std::vector<Range> search_all_batch(const NeedleBatch &b) const {
    std::vector<int32_t> left(b.size()), count(b.size());
    gsa_check(gsa_search_all_batch(handle(), b.bytes.data(), b.off.data(), b.size(), left.data(), count.data()), "gsa_search_all_batch");
    std::vector<Range> out(b.size());
    for (size_t q = 0; q < out.size(); ++q) out[q] = {left[q], count[q]};
    return out;
  }
  std::vector<int32_t> search_all(const std::string &pattern) const {
    NeedleBatch b;
    b.push(pattern);
    const Range r = search_all_batch(b)[0];
    if (r.count <= 0) return {};
    return std::vector<int32_t>(sa_.begin() + r.left, sa_.begin() + r.left + r.count);
  }
  bool contains(const std::string &pattern) const {
    NeedleBatch b;
    b.push(pattern);
    return search_all_batch(b)[0].count > 0;
  }

 private:
  struct Deleter { void operator()(gsa_index *p) const { gsa_index_destroy(p); } };
  gsa_index *handle() const {
    if (!dev_) {
      gsa_index *h = nullptr;
      gsa_check(gsa_index_from_parts(text_, (int64_t)text_len_, sa_.data(), (int64_t)sa_.size(), device_, &h), "gsa_index_from_parts");
      dev_.reset(h);
    }
    return dev_.get();
  }
  const uint8_t *text_;
  size_t text_len_;
  std::vector<int32_t> sa_;
  int device_;
  mutable std::unique_ptr<gsa_index, Deleter> dev_;
};

// free functions of the crate (lib.rs:39-99, 127-149)
inline LongestCommonSubstring longest_substring_match(const uint8_t *text, size_t text_len, const std::vector<int32_t> &sa,
                                                      const uint8_t *needle, size_t needle_len) {
  return SuffixArray(text, text_len, sa).longest_substring_match(needle, needle_len);
}
inline void verify(const uint8_t *text, size_t text_len, const std::vector<int32_t> &sa) {
  SuffixArray(text, text_len, sa).verify();
}

}  // namespace sacabase
