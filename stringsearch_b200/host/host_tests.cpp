// host_tests.cpp -- the reference's own unit tests, restated against the C++ mirror
// (run on a GPU box: tests/test_gpu_host_cpp.py builds and runs it).
//   crates/sacapart/src/lib.rs:105-165   worse_test, equivalent_test
//   crates/divsufsort/src/lib.rs:84-91   shruggy + sort-then-verify helper
#include <cassert>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include "divsufsort.hpp"
#include "sacapart.hpp"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static std::string bytes_of(const sacabase::LongestCommonSubstring &m) {
  auto b = m.as_bytes();
  return std::string(reinterpret_cast<const char *>(b.first), b.second);
}

int main() {
  {  // worse_test
    const std::string input = "totor";
    auto sa_full = divsufsort::sort(input);
    sacapart::PartitionedSuffixArray sa_part(reinterpret_cast<const uint8_t *>(input.data()), input.size(), 2);
    std::string needle = "tor";
    CHECK(bytes_of(sa_full.longest_substring_match(needle)) == needle);
    CHECK(bytes_of(sa_part.longest_substring_match(needle)) == needle.substr(0, 2));
    needle = "otor";
    CHECK(bytes_of(sa_full.longest_substring_match(needle)) == needle);
    CHECK(bytes_of(sa_part.longest_substring_match(needle)) == needle);
  }
  {  // equivalent_test
    const std::string input = "This is a rather long text. We can probably find matches that span two partitions. Oh yes.";
    auto sa_full = divsufsort::sort(input);
    for (size_t partitions : {1, 2, 3}) {
      for (const char *nd : {"rather long", "text. We can", "We can probably find matches that span"}) {
        sacapart::PartitionedSuffixArray sa_part(reinterpret_cast<const uint8_t *>(input.data()), input.size(), partitions);
        auto fm = sa_full.longest_substring_match(std::string(nd));
        auto pm = sa_part.longest_substring_match(std::string(nd));
        CHECK(bytes_of(fm) == bytes_of(pm));
        CHECK(fm.start == pm.start);
        CHECK(fm.len == pm.len);
      }
    }
  }
  {  // shruggy + verify; search_all / contains
    const std::string s = "\xc2\xaf\\_(\xe3\x83\x84)_/\xc2\xaf";
    auto sa = divsufsort::sort(s);
    sa.verify();
    const std::vector<int32_t> expect = {4, 8, 10, 2, 3, 9, 6, 7, 12, 1, 11, 0, 5};
    CHECK(sa.sa() == expect);
    const std::string banana = "banana";  // the SuffixArray borrows its text (&'a [u8])
    auto b = divsufsort::sort(banana);
    CHECK((b.search_all("ana") == std::vector<int32_t>{3, 1}));
    CHECK(b.contains("nan") && !b.contains("nab"));
  }
  {  // BWT round trip and LCP (divsufsort.c:372-405, utils.c:111-156; lcp.cu)
    const std::string banana = "banana";
    const auto *bt = reinterpret_cast<const uint8_t *>(banana.data());
    auto tr = divsufsort::bwt(bt, banana.size());
    CHECK(std::string(tr.first.begin(), tr.first.end()) == "annbaa" && tr.second == 4);
    auto back = divsufsort::inverse_bwt(tr.first.data(), tr.first.size(), tr.second);
    CHECK(std::string(back.begin(), back.end()) == banana);
    auto b = divsufsort::sort(banana);
    CHECK((divsufsort::lcp(bt, banana.size(), b.sa().data()) == std::vector<int32_t>{0, 1, 3, 0, 0, 2}));
    std::string big;
    for (int i = 0; i < 50000; ++i) big += "abracadabra"[(i * 7 + i / 13) % 11];
    const auto *gt = reinterpret_cast<const uint8_t *>(big.data());
    auto tr2 = divsufsort::bwt(gt, big.size());
    auto back2 = divsufsort::inverse_bwt(tr2.first.data(), tr2.first.size(), tr2.second);
    CHECK(std::string(back2.begin(), back2.end()) == big);
  }
  {  // panics
    bool threw = false;
    try { sacapart::PartitionedSuffixArray p(nullptr, 0, 0); } catch (const std::logic_error &) { threw = true; }
    CHECK(threw);
    threw = false;
    try {
      int32_t sa[2];
      divsufsort::sort_in_place(reinterpret_cast<const uint8_t *>("abc"), 3, sa, 2);
    } catch (const std::logic_error &) { threw = true; }
    CHECK(threw);
  }
  std::puts("host_tests: all passed");
  return 0;
}
