// sacapart.hpp -- C++ mirror of the reference crate `sacapart`
// (crates/sacapart/src/lib.rs:26-98): a partitioned suffix array whose shards live on one
// or more GPUs.  The builder closure of the reference (`f: Fn(&[u8]) -> SuffixArray`) is
// fixed to the GPU divsufsort; `devices` says where shard i goes (devices[i % ndev]).
#pragma once
#include "sacabase.hpp"

namespace sacapart {

class PartitionedSuffixArray : public sacabase::StringIndex {
 public:
  // PartitionedSuffixArray::new(text, num_partitions, divsufsort::sort)  (lib.rs:39-58)
  PartitionedSuffixArray(const uint8_t *text, size_t text_len, size_t num_partitions, std::vector<int32_t> devices = {})
      : text_(text), text_len_(text_len) {
    if (num_partitions == 0) throw std::logic_error("attempt to divide by zero");  // lib.rs:43
    gsa_part *h = nullptr;
    sacabase::gsa_check(gsa_part_create(text, text_len, num_partitions, devices.empty() ? nullptr : devices.data(),
                                        (int32_t)devices.size(), &h),
                        "gsa_part_create");
    h_.reset(h);
  }

  size_t num_partitions() const { return (size_t)gsa_part_num_partitions(h_.get()); }  // lib.rs:60-62
  size_t partition_size() const { return (size_t)gsa_part_partition_size(h_.get()); }

  using sacabase::StringIndex::longest_substring_match;
  // lib.rs:69-97
  sacabase::LongestCommonSubstring longest_substring_match(const uint8_t *needle, size_t needle_len) const override {
    sacabase::NeedleBatch b;
    b.push(needle, needle_len);
    return longest_substring_match_batch(b)[0];
  }

  std::vector<sacabase::LongestCommonSubstring> longest_substring_match_batch(const sacabase::NeedleBatch &b) const {
    std::vector<uint64_t> st(b.size());
    std::vector<uint32_t> ln(b.size());
    const int32_t rc = gsa_part_lsm_batch(h_.get(), b.bytes.data(), b.off.data(), b.size(), st.data(), ln.data());
    if (rc == GSA_EPANIC)  // lib.rs:94-96
      throw std::logic_error("partitioned suffix arrays should always find at least one longest common substring");
    sacabase::gsa_check(rc, "gsa_part_lsm_batch");
    std::vector<sacabase::LongestCommonSubstring> out(b.size());
    for (size_t q = 0; q < out.size(); ++q) out[q] = {text_, text_len_, (size_t)st[q], (size_t)ln[q]};
    return out;
  }

 private:
  struct Deleter { void operator()(gsa_part *p) const { gsa_part_destroy(p); } };
  const uint8_t *text_;
  size_t text_len_;
  std::unique_ptr<gsa_part, Deleter> h_;
};

}  // namespace sacapart
