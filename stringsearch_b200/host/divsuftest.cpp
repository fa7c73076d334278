// divsuftest -- successor of the reference harness crates/divsuftest/src/main.rs:
//
//     divsuftest bench|run|verify INPUT [LENGTH] [--device D] [--partitions P] [--cpu-lib PATH]
//
//   run     time one GPU divsufsort::sort of INPUT[..LENGTH]            (main.rs:115-121)
//   bench   table of time / throughput per implementation                (main.rs:123-190):
//           "gpu-divsufsort (device)", "gpu-divsufsort (host->host)", "gpu-sacapart (P)",
//           and, when --cpu-lib names a libdivsufsort shared object exporting
//           divsufsort(T, SA, n), that CPU implementation timed on the same bytes and
//           compared byte for byte with the GPU result
//   verify  build on the GPU and run the O(n) GPU sufcheck (replaces `crosscheck`, whose
//           line-by-line trace diff is tied to sharing divsufsort's internal steps)
// LENGTH accepts k / m suffixes like the reference (main.rs:192-208).
#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "divsufsort.hpp"
#include "sacapart.hpp"

static size_t parse_size(std::string s) {  // main.rs:192-208
  size_t factor = 1;
  for (auto &c : s) c = (char)tolower(c);
  if (!s.empty() && s.back() == 'k') { factor = 1024; s.pop_back(); }
  else if (!s.empty() && s.back() == 'm') { factor = 1024 * 1024; s.pop_back(); }
  return (size_t)std::stoull(s) * factor;
}

static std::string human(double bytes_per_s) {
  const char *u[] = {"B/s", "KiB/s", "MiB/s", "GiB/s", "TiB/s"};
  int i = 0;
  while (bytes_per_s >= 1024.0 && i < 4) { bytes_per_s /= 1024.0; ++i; }
  char buf[64];
  snprintf(buf, sizeof buf, "%.2f %s", bytes_per_s, u[i]);
  return buf;
}

[[noreturn]] static void usage() {
  std::puts("Usage: divsuftest bench|run|verify INPUT [LENGTH] [--device D] [--partitions P] [--cpu-lib PATH]");
  std::exit(1);
}

struct Row { std::string name; double secs; };

int main(int argc, char **argv) {
  std::vector<std::string> free_args;
  int device = 0;
  size_t partitions = 8;
  std::string cpu_lib;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
    else if (a == "--partitions" && i + 1 < argc) partitions = (size_t)std::atoll(argv[++i]);
    else if (a == "--cpu-lib" && i + 1 < argc) cpu_lib = argv[++i];
    else free_args.push_back(a);
  }
  if (free_args.size() < 2) usage();
  const std::string cmd = free_args[0];
  if (cmd != "bench" && cmd != "run" && cmd != "verify") {
    std::puts("Command should be one of bench, run or verify");
    return 1;
  }
  std::ifstream f(free_args[1], std::ios::binary);
  if (!f) { std::fprintf(stderr, "cannot read %s\n", free_args[1].c_str()); return 1; }
  std::vector<uint8_t> full((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  size_t len = full.size();
  if (free_args.size() > 2) len = std::min(len, parse_size(free_args[2]));
  const uint8_t *input = full.data();
  std::printf("Input is size %zuB\n", len);
  if (gsa_device_count() < 1) { std::fprintf(stderr, "no CUDA device: divsuftest has no CPU path\n"); return 2; }

  using clk = std::chrono::steady_clock;
  auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  try {
    if (cmd == "run") {
      auto t0 = clk::now();
      auto sa = divsufsort::sort(input, len, device);
      std::printf("Done in %.6fs\n", secs(t0, clk::now()));
      return 0;
    }
    if (cmd == "verify") {
      gsa_build_stats st;
      auto sa = divsufsort::sort(input, len, device, &st);
      std::printf("built in %.3f ms on the device (%u rounds, sigma %u)\n", st.ms_total, st.rounds, st.sigma);
      if (len > 0) sa.verify();
      std::puts("suffix array verified (GPU sufcheck: permutation + adjacent-suffix order)");
      return 0;
    }
    // bench
    std::vector<Row> rows;
    std::vector<int32_t> gpu_sa(len);
    {
      divsufsort::sort_in_place(input, len, gpu_sa.data(), len, device);  // warm-up (context, module load)
      gsa_build_stats st;
      auto t0 = clk::now();
      divsufsort::sort_in_place(input, len, gpu_sa.data(), len, device, &st);
      rows.push_back({"gpu-divsufsort (host->host)", secs(t0, clk::now())});
      rows.push_back({"gpu-divsufsort (device only)", st.ms_total / 1e3});
      std::printf("rounds:");
      for (uint32_t r = 0; r < st.rounds && r < GSA_MAX_ROUNDS; ++r)
        std::printf(" [h=%llu L=%llu p=%u %.2fms]", (unsigned long long)st.round[r].depth,
                    (unsigned long long)st.round[r].live, st.round[r].passes, st.round[r].ms_total);
      std::puts("");
    }
    if (partitions > 0 && len > 0) {
      auto t0 = clk::now();
      sacapart::PartitionedSuffixArray psa(input, len, partitions, {device});
      rows.push_back({"gpu-sacapart (" + std::to_string(psa.num_partitions()) + " partitions)", secs(t0, clk::now())});
    }
    if (!cpu_lib.empty()) {
      void *h = dlopen(cpu_lib.c_str(), RTLD_NOW);
      if (!h) { std::fprintf(stderr, "dlopen %s: %s\n", cpu_lib.c_str(), dlerror()); return 1; }
      using fn_t = int32_t (*)(const uint8_t *, int32_t *, int32_t);
      fn_t fn = (fn_t)dlsym(h, "divsufsort");
      if (!fn) { std::fprintf(stderr, "%s does not export divsufsort\n", cpu_lib.c_str()); return 1; }
      std::vector<int32_t> cpu_sa(len);
      auto t0 = clk::now();
      const int32_t rc = fn(input, cpu_sa.data(), (int32_t)len);
      rows.push_back({"c-divsufsort (1 thread)", secs(t0, clk::now())});
      if (rc != 0) { std::fprintf(stderr, "cpu divsufsort rc=%d\n", rc); return 1; }
      std::printf("GPU vs CPU suffix arrays: %s\n", cpu_sa == gpu_sa ? "IDENTICAL" : "DIFFERENT");
      if (cpu_sa != gpu_sa) return 3;
    }
    std::printf("%-36s %14s %16s\n", "Algorithm", "Time", "Average speed");
    for (const Row &r : rows) std::printf("%-36s %12.6fs %16s\n", r.name.c_str(), r.secs, human(len / r.secs).c_str());
  } catch (const std::exception &e) {
    std::fprintf(stderr, "panic: %s\n", e.what());
    return 101;
  }
  return 0;
}
