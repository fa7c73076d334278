// api.cu -- the extern "C" boundary declared in include/gsa.h.
//
// Host-side logic only: argument checks that mirror the reference's entry points,
// device selection, host<->device staging, handle lifetime, and the sacapart fan-out.
// All compute is in sa_build.cu / search.cu / verify.cu.  There is no CPU fallback: with
// no usable GPU every compute entry point returns GSA_ECUDA.
#include <algorithm>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "builder.h"

namespace gsa {

static thread_local std::string g_last_error;

void set_error(const char *what, const char *file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s (%s:%d)", what ? what : "unknown error", file, line);
  g_last_error = buf;
}

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
    cudaError_t e = cudaSetDevice(dev);
    ok = (e == cudaSuccess);
    if (!ok) { set_error(cudaGetErrorString(e), __FILE__, __LINE__); cudaGetLastError(); }
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct Stream {
  cudaStream_t s = nullptr;
  int create() {
    GSA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    return GSA_OK;
  }
  ~Stream() { if (s) cudaStreamDestroy(s); }
};

template <typename T>
struct DevBuf {
  T *p = nullptr;
  int alloc(size_t count) {
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T) + 64);
    if (e != cudaSuccess) {
      p = nullptr;
      set_error(cudaGetErrorString(e), __FILE__, __LINE__);
      cudaGetLastError();
      return GSA_ENOMEM;
    }
    return GSA_OK;
  }
  T *release() { T *r = p; p = nullptr; return r; }
  ~DevBuf() { if (p) cudaFree(p); }
};

// Per-device cache of the scratch block used by the host-pointer entry points (text + SA +
// sort workspace).  cudaMalloc/cudaFree of tens of GB costs more than the copies they
// bracket; the block is kept between calls and reused when it is large enough and free.
// A concurrent call on the same device (sacapart-style threaded builders) simply allocates
// its own block.  gsa_release_cached_memory() returns everything to the driver.
// ---------------------------------------------------------------------------------------------
// Large copies from / to PAGEABLE host memory (a Rust Vec, a numpy array): the driver stages them
// through its own small pinned buffer at 11 GB/s up and 19 GB/s down (measured here; pinned memory
// gets 52-55 GB/s).  These two helpers bounce through two 32 MiB pinned buffers of our own and let a
// few host threads do the memcpy of one chunk while the next chunk is on the bus.  Pinned (or
// registered / managed) buffers and small copies take the plain cudaMemcpyAsync path.
// ---------------------------------------------------------------------------------------------
constexpr size_t kStageChunk = 32u << 20;
constexpr size_t kStageMin = 128u << 20;   // below this a plain copy is as good
constexpr int kStageThreads = 4;

struct StageCache {
  std::mutex mu;
  char *buf[2] = {nullptr, nullptr};
  bool busy = false;
} g_stage;

struct StageLease {
  char *buf[2] = {nullptr, nullptr};
  bool held = false;
  bool acquire() {
    std::lock_guard<std::mutex> lk(g_stage.mu);
    if (g_stage.busy) return false;  // another host thread is copying: that one uses the plain path
    for (int i = 0; i < 2; ++i) {
      if (!g_stage.buf[i] && cudaHostAlloc(reinterpret_cast<void **>(&g_stage.buf[i]), kStageChunk, cudaHostAllocDefault) != cudaSuccess) {
        g_stage.buf[i] = nullptr;
        cudaGetLastError();
        return false;
      }
      buf[i] = g_stage.buf[i];
    }
    g_stage.busy = held = true;
    return true;
  }
  ~StageLease() {
    if (held) { std::lock_guard<std::mutex> lk(g_stage.mu); g_stage.busy = false; }
  }
};

static bool is_pageable(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}

static void parallel_memcpy(char *dst, const char *src, size_t bytes) {
  const size_t per = align_up(div_up(bytes, kStageThreads), 4096);
  std::thread th[kStageThreads - 1];
  int started = 0;
  size_t done_to = std::min(bytes, per);  // [0, done_to) is copied by this thread
  for (int t = 1; t < kStageThreads; ++t) {
    const size_t lo = std::min(bytes, per * (size_t)t), hi = std::min(bytes, lo + per);
    if (hi <= lo) break;
    try {
      th[started] = std::thread([=] { memcpy(dst + lo, src + lo, hi - lo); });
      ++started;
    } catch (...) {  // no thread to be had: this one copies the rest as well (nothing may escape the C ABI)
      memcpy(dst + lo, src + lo, bytes - lo);
      break;
    }
  }
  memcpy(dst, src, done_to);
  for (int t = 0; t < started; ++t) th[t].join();
}

static int copy_h2d(void *d, const void *h, size_t bytes, cudaStream_t st) {
  StageLease lease;
  if (bytes < kStageMin || !is_pageable(h) || !lease.acquire()) {
    GSA_TRY(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st));
    return GSA_OK;
  }
  cudaEvent_t ev[2];
  GSA_TRY(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
  GSA_TRY(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
  struct EvFree { cudaEvent_t *e; ~EvFree() { cudaEventDestroy(e[0]); cudaEventDestroy(e[1]); } } evg{ev};
  size_t k = 0;
  for (size_t off = 0; off < bytes; off += kStageChunk, ++k) {
    const size_t len = std::min(kStageChunk, bytes - off);
    const int i = (int)(k & 1);
    if (k >= 2) GSA_TRY(cudaEventSynchronize(ev[i]));  // the DMA that last read this staging buffer is done
    parallel_memcpy(lease.buf[i], static_cast<const char *>(h) + off, len);
    GSA_TRY(cudaMemcpyAsync(static_cast<char *>(d) + off, lease.buf[i], len, cudaMemcpyHostToDevice, st));
    GSA_TRY(cudaEventRecord(ev[i], st));
  }
  GSA_TRY(cudaStreamSynchronize(st));  // the staging buffers go back to the cache
  return GSA_OK;
}

// Synchronous with respect to `st`: returns when the bytes are in `h`.
static int copy_d2h(void *h, const void *d, size_t bytes, cudaStream_t st) {
  StageLease lease;
  if (bytes < kStageMin || !is_pageable(h) || !lease.acquire()) {
    GSA_TRY(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st));
    return GSA_OK;
  }
  cudaEvent_t ev[2];
  GSA_TRY(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
  GSA_TRY(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
  struct EvFree { cudaEvent_t *e; ~EvFree() { cudaEventDestroy(e[0]); cudaEventDestroy(e[1]); } } evg{ev};
  const size_t chunks = div_up(bytes, kStageChunk);
  for (size_t k = 0; k <= chunks; ++k) {
    if (k < chunks) {  // put chunk k on the bus ...
      const size_t off = k * kStageChunk, len = std::min(kStageChunk, bytes - off);
      GSA_TRY(cudaMemcpyAsync(lease.buf[k & 1], static_cast<const char *>(d) + off, len, cudaMemcpyDeviceToHost, st));
      GSA_TRY(cudaEventRecord(ev[k & 1], st));
    }
    if (k >= 1) {      // ... and move chunk k - 1 to its place meanwhile
      const size_t off = (k - 1) * kStageChunk, len = std::min(kStageChunk, bytes - off);
      GSA_TRY(cudaEventSynchronize(ev[(k - 1) & 1]));
      parallel_memcpy(static_cast<char *>(h) + off, lease.buf[(k - 1) & 1], len);
    }
  }
  return GSA_OK;
}

struct ScratchCache {
  static constexpr int kMaxDev = 64;
  std::mutex mu;
  char *ptr[kMaxDev] = {};
  size_t bytes[kMaxDev] = {};
  bool busy[kMaxDev] = {};
} g_scratch;

struct Scratch {
  char *p = nullptr;
  size_t bytes = 0;
  int dev = -1;
  bool cached = false;
  int acquire(int device, size_t need) {
    dev = device;
    if (device >= 0 && device < ScratchCache::kMaxDev) {
      std::lock_guard<std::mutex> lk(g_scratch.mu);
      if (!g_scratch.busy[device]) {
        if (g_scratch.bytes[device] < need) {
          if (g_scratch.ptr[device]) cudaFree(g_scratch.ptr[device]);
          g_scratch.ptr[device] = nullptr;
          g_scratch.bytes[device] = 0;
          char *np = nullptr;
          if (cudaMalloc(&np, need) != cudaSuccess) {
            set_error("cudaMalloc failed (scratch)", __FILE__, __LINE__);
            cudaGetLastError();
            return GSA_ENOMEM;
          }
          g_scratch.ptr[device] = np;
          g_scratch.bytes[device] = need;
        }
        g_scratch.busy[device] = true;
        p = g_scratch.ptr[device];
        bytes = g_scratch.bytes[device];
        cached = true;
        return GSA_OK;
      }
    }
    if (cudaMalloc(&p, need) != cudaSuccess) {
      p = nullptr;
      set_error("cudaMalloc failed (scratch)", __FILE__, __LINE__);
      cudaGetLastError();
      return GSA_ENOMEM;
    }
    bytes = need;
    return GSA_OK;
  }
  ~Scratch() {
    if (!p) return;
    if (cached) {
      std::lock_guard<std::mutex> lk(g_scratch.mu);
      g_scratch.busy[dev] = false;
    } else {
      cudaFree(p);
    }
  }
};

int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return 0; }
  return d;
}

}  // namespace
}  // namespace gsa

using namespace gsa;

struct gsa_index {
  int device = 0;
  u64 n = 0;           // suffix-array length == logical text length
  u64 text_avail = 0;  // bytes resident behind d_text (n + halo)
  u8 *d_text = nullptr;
  i32 *d_sa = nullptr;
  // shard bookkeeping (halo top-up needs the caller's text; see gsa_part_create)
  const u8 *host_full = nullptr;
  u64 n_full = 0;
  u64 offset = 0;
  // prefix-bucket table of the searches (search.cu), built by the first query on this index
  std::mutex accel_mu;
  bool accel_ready = false;
  AccelView accel{};
  u32 *d_accel = nullptr;
};

struct gsa_part {
  const u8 *text = nullptr;
  u64 n = 0;
  u64 partition_size = 0;
  std::vector<gsa_index *> shards;
  std::vector<int> devices;
  std::mutex query_mu;  // gsa_part_lsm_batch may re-allocate a shard's text (halo top-up): one query at a time
};

extern "C" {

const char *gsa_last_error(void) { return g_last_error.c_str(); }
const char *gsa_version(void) { return "gsa 0.1 (sm_100a, prefix doubling + onesweep LSD radix)"; }

int32_t gsa_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
  return c;
}

int32_t gsa_current_device(void) { return current_device(); }

void gsa_release_cached_memory(void) {
  std::lock_guard<std::mutex> lk(g_scratch.mu);
  int prev = -1;
  cudaGetDevice(&prev);
  for (int d = 0; d < ScratchCache::kMaxDev; ++d) {
    if (g_scratch.ptr[d] && !g_scratch.busy[d]) {
      cudaSetDevice(d);
      cudaFree(g_scratch.ptr[d]);
      g_scratch.ptr[d] = nullptr;
      g_scratch.bytes[d] = 0;
    }
  }
  if (prev >= 0) cudaSetDevice(prev);
  cudaGetLastError();
}

void *gsa_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void gsa_host_free(void *p) { if (p) cudaFreeHost(p); }

size_t gsa_build_workspace_bytes(int32_t n) { return n <= 0 ? 0 : build_workspace_bytes((u32)n); }

int32_t gsa_build_device(const uint8_t *d_T, int32_t *d_SA, int32_t n, void *workspace, size_t workspace_bytes,
                         void *stream, gsa_build_stats *stats) {
  if (d_T == nullptr || d_SA == nullptr || n < 0) return GSA_EINVAL;
  return build_sa_device(d_T, d_SA, (u32)n, workspace, workspace_bytes, static_cast<cudaStream_t>(stream), stats);
}

// divsufsort(T, SA, n): argument checks and the n <= 2 shortcuts are the reference's own
// (c-sources/divsufsort.c:346-349; crates/divsufsort/src/divsufsort.rs:18-29).
int32_t gsa_divsufsort_ex(const uint8_t *T, int32_t *SA, int32_t n, int32_t device, gsa_build_stats *stats) {
  if (stats) memset(stats, 0, sizeof(*stats));
  if (T == nullptr || SA == nullptr || n < 0) return GSA_EINVAL;
  if (n == 0) return GSA_OK;
  if (n == 1) { SA[0] = 0; return GSA_OK; }
  if (n == 2) {
    const int m = T[0] < T[1];
    SA[m ^ 1] = 0;
    SA[m] = 1;
    return GSA_OK;
  }
  DeviceGuard dg(device);
  if (!dg.ok) return GSA_ECUDA;
  Stream st;
  GSA_TRY_RC(st.create());
  // one scratch block: [text | SA | sort workspace]
  const size_t text_bytes = align_up((size_t)n + 64, 256), sa_bytes = align_up((size_t)n * sizeof(i32), 256);
  const size_t ws_bytes = build_workspace_bytes((u32)n);
  Scratch sc;
  GSA_TRY_RC(sc.acquire(device, text_bytes + sa_bytes + ws_bytes));
  u8 *d_T = reinterpret_cast<u8 *>(sc.p);
  i32 *d_SA = reinterpret_cast<i32 *>(sc.p + text_bytes);
  void *d_ws = sc.p + text_bytes + sa_bytes;
  cudaEvent_t e0, e1, e2, e3;
  GSA_TRY(cudaEventCreate(&e0)); GSA_TRY(cudaEventCreate(&e1));
  GSA_TRY(cudaEventCreate(&e2)); GSA_TRY(cudaEventCreate(&e3));
  struct EvFree { cudaEvent_t a, b, c, d; ~EvFree() { cudaEventDestroy(a); cudaEventDestroy(b); cudaEventDestroy(c); cudaEventDestroy(d); } } evg{e0, e1, e2, e3};
  GSA_TRY(cudaEventRecord(e0, st.s));
  GSA_TRY_RC(copy_h2d(d_T, T, (size_t)n, st.s));
  GSA_TRY(cudaEventRecord(e1, st.s));
  gsa_build_stats local;
  gsa_build_stats *sp = stats ? stats : &local;
  GSA_TRY_RC(build_sa_device(d_T, d_SA, (u32)n, d_ws, ws_bytes, st.s, sp));
  GSA_TRY(cudaEventRecord(e2, st.s));
  GSA_TRY_RC(copy_d2h(SA, d_SA, (size_t)n * sizeof(i32), st.s));
  GSA_TRY(cudaEventRecord(e3, st.s));
  GSA_TRY(cudaStreamSynchronize(st.s));
  cudaEventElapsedTime(&sp->ms_h2d, e0, e1);
  cudaEventElapsedTime(&sp->ms_d2h, e2, e3);
  return GSA_OK;
}

int32_t gsa_divsufsort(const uint8_t *T, int32_t *SA, int32_t n) {
  return gsa_divsufsort_ex(T, SA, n, current_device(), nullptr);
}

// divbwt(T, U, A, n) (c-sources/divsufsort.c:372-405): `A` is only scratch in the reference and is
// ignored here (the SA is built in device memory).  Host pointers.
int32_t gsa_divbwt(const uint8_t *T, uint8_t *U, int32_t *A, int32_t n) {
  (void)A;
  if (T == nullptr || U == nullptr || n < 0) return GSA_EINVAL;
  if (n <= 1) { if (n == 1) U[0] = T[0]; return n; }
  const int device = current_device();
  DeviceGuard dg(device);
  if (!dg.ok) return GSA_ECUDA;
  Stream st;
  GSA_TRY_RC(st.create());
  const size_t text_bytes = align_up((size_t)n + 64, 256), sa_bytes = align_up((size_t)n * sizeof(i32), 256);
  const size_t ws_bytes = std::max(build_workspace_bytes((u32)n), text_bytes);
  Scratch sc;
  GSA_TRY_RC(sc.acquire(device, text_bytes + sa_bytes + ws_bytes));
  u8 *d_T = reinterpret_cast<u8 *>(sc.p);
  i32 *d_SA = reinterpret_cast<i32 *>(sc.p + text_bytes);
  char *d_ws = sc.p + text_bytes + sa_bytes;
  GSA_TRY_RC(copy_h2d(d_T, T, (size_t)n, st.s));
  GSA_TRY_RC(build_sa_device(d_T, d_SA, (u32)n, d_ws, ws_bytes, st.s, nullptr));
  u8 *d_U = reinterpret_cast<u8 *>(d_ws);  // the sort workspace is free again
  i32 pidx = 0;
  GSA_TRY_RC(bwt_device(d_T, d_SA, (u32)n, d_U, &pidx, st.s));
  GSA_TRY_RC(copy_d2h(U, d_U, (size_t)n, st.s));
  GSA_TRY(cudaStreamSynchronize(st.s));
  return pidx;
}

int32_t gsa_bwt_device(const uint8_t *d_T, const int32_t *d_SA, int32_t n, uint8_t *d_U, int32_t *primary_index,
                       void *stream) {
  if (n < 0 || (n > 0 && (!d_T || !d_SA || !d_U))) return GSA_EINVAL;
  return bwt_device(d_T, d_SA, (u32)n, d_U, primary_index, static_cast<cudaStream_t>(stream));
}

// inverse_bw_transform(T, U, A, n, idx) (c-sources/utils.c:111-156).  Host pointers; `A` is ignored.
int32_t gsa_inverse_bw_transform(const uint8_t *T, uint8_t *U, int32_t *A, int32_t n, int32_t idx) {
  (void)A;
  if (T == nullptr || U == nullptr || n < 0 || idx < 0 || n < idx || (0 < n && idx == 0)) return GSA_EINVAL;  // utils.c:120-123
  if (n <= 1) { if (n == 1) U[0] = T[0]; return GSA_OK; }
  const int device = current_device();
  DeviceGuard dg(device);
  if (!dg.ok) return GSA_ECUDA;
  Stream st;
  GSA_TRY_RC(st.create());
  const size_t text_bytes = align_up((size_t)n + 64, 256);
  const size_t ws_bytes = inverse_bwt_workspace_bytes((u32)n);
  Scratch sc;
  GSA_TRY_RC(sc.acquire(device, 2 * text_bytes + ws_bytes));
  u8 *d_T = reinterpret_cast<u8 *>(sc.p);
  u8 *d_U = reinterpret_cast<u8 *>(sc.p + text_bytes);
  GSA_TRY_RC(copy_h2d(d_T, T, (size_t)n, st.s));
  GSA_TRY_RC(inverse_bwt_device(d_T, d_U, (u32)n, (u32)idx, sc.p + 2 * text_bytes, ws_bytes, st.s));
  GSA_TRY_RC(copy_d2h(U, d_U, (size_t)n, st.s));
  GSA_TRY(cudaStreamSynchronize(st.s));
  return GSA_OK;
}

size_t gsa_inverse_bwt_workspace_bytes(int32_t n) { return inverse_bwt_workspace_bytes(n < 0 ? 0u : (u32)n); }

int32_t gsa_inverse_bwt_device(const uint8_t *d_T, uint8_t *d_U, int32_t n, int32_t idx, void *workspace,
                               size_t workspace_bytes, void *stream) {
  if (n < 0 || idx < 0 || n < idx || (0 < n && (idx == 0 || !d_T || !d_U))) return GSA_EINVAL;
  return inverse_bwt_device(d_T, d_U, (u32)n, (u32)idx, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t gsa_lcp_workspace_bytes(int32_t n) { return lcp_workspace_bytes(n < 0 ? 0u : (u32)n); }

int32_t gsa_lcp_device(const uint8_t *d_T, const int32_t *d_SA, int32_t *d_LCP, int32_t n, void *workspace,
                       size_t workspace_bytes, void *stream) {
  if (n < 0 || (n > 0 && (!d_T || !d_SA || !d_LCP))) return GSA_EINVAL;
  return lcp_device(d_T, d_SA, (u32)n, d_LCP, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

// Host-pointer forms.  build != 0: the suffix array is built here (SA may then be NULL = not wanted).
static int32_t lcp_host(const uint8_t *T, int32_t *SA, int32_t *LCP, int32_t n, int32_t device, int build) {
  if (n < 0 || (n > 0 && (T == nullptr || (SA == nullptr && !build)))) return GSA_EINVAL;
  if (n == 0) return GSA_OK;
  DeviceGuard dg(device);
  if (!dg.ok) return GSA_ECUDA;
  Stream st;
  GSA_TRY_RC(st.create());
  const size_t text_bytes = align_up((size_t)n + 64, 256), sa_bytes = align_up((size_t)n * sizeof(i32), 256);
  const size_t ws_bytes = std::max(build ? build_workspace_bytes((u32)n) : (size_t)0, sa_bytes + lcp_workspace_bytes((u32)n));
  Scratch sc;
  GSA_TRY_RC(sc.acquire(device, text_bytes + sa_bytes + ws_bytes));
  u8 *d_T = reinterpret_cast<u8 *>(sc.p);
  i32 *d_SA = reinterpret_cast<i32 *>(sc.p + text_bytes);
  char *d_ws = sc.p + text_bytes + sa_bytes;
  GSA_TRY_RC(copy_h2d(d_T, T, (size_t)n, st.s));
  if (build) {
    GSA_TRY_RC(build_sa_device(d_T, d_SA, (u32)n, d_ws, ws_bytes, st.s, nullptr));
    if (SA) GSA_TRY_RC(copy_d2h(SA, d_SA, (size_t)n * sizeof(i32), st.s));
  } else {
    GSA_TRY_RC(copy_h2d(d_SA, SA, (size_t)n * sizeof(i32), st.s));
  }
  if (LCP) {
    i32 *d_LCP = reinterpret_cast<i32 *>(d_ws);  // the sort workspace is free again
    GSA_TRY_RC(lcp_device(d_T, d_SA, (u32)n, d_LCP, d_ws + sa_bytes, ws_bytes - sa_bytes, st.s));
    GSA_TRY_RC(copy_d2h(LCP, d_LCP, (size_t)n * sizeof(i32), st.s));
  }
  GSA_TRY(cudaStreamSynchronize(st.s));
  return GSA_OK;
}

int32_t gsa_lcp(const uint8_t *T, const int32_t *SA, int32_t *LCP, int32_t n, int32_t device) {
  if (n > 0 && LCP == nullptr) return GSA_EINVAL;
  return lcp_host(T, const_cast<int32_t *>(SA), LCP, n, device, 0);
}

int32_t gsa_divsufsort_lcp(const uint8_t *T, int32_t *SA, int32_t *LCP, int32_t n, int32_t device) {
  return lcp_host(T, SA, LCP, n, device, 1);
}

int32_t gsa_sufcheck_device(const uint8_t *d_T, const int32_t *d_SA, int32_t n, void *stream, int64_t *bad_index) {
  if (n < 0 || (n > 0 && (d_T == nullptr || d_SA == nullptr))) return GSA_EINVAL;
  return sufcheck_device(d_T, d_SA, (u32)n, static_cast<cudaStream_t>(stream), bad_index);
}

int32_t gsa_sufcheck(const uint8_t *T, const int32_t *SA, int32_t n, int32_t device, int64_t *bad_index) {
  if (bad_index) *bad_index = -1;
  if (n < 0 || (n > 0 && (T == nullptr || SA == nullptr))) return GSA_EINVAL;
  if (n == 0) return 0;
  DeviceGuard dg(device);
  if (!dg.ok) return GSA_ECUDA;
  Stream st;
  GSA_TRY_RC(st.create());
  DevBuf<u8> d_T;
  DevBuf<i32> d_SA;
  GSA_TRY_RC(d_T.alloc((size_t)n));
  GSA_TRY_RC(d_SA.alloc((size_t)n));
  GSA_TRY(cudaMemcpyAsync(d_T.p, T, (size_t)n, cudaMemcpyHostToDevice, st.s));
  GSA_TRY(cudaMemcpyAsync(d_SA.p, SA, (size_t)n * 4, cudaMemcpyHostToDevice, st.s));
  return sufcheck_device(d_T.p, d_SA.p, (u32)n, st.s, bad_index);
}

// ------------------------------------------------------------------------------------------
// Resident index
// ------------------------------------------------------------------------------------------
static int index_create_impl(const u8 *T_full, u64 n_full, u64 offset, u64 len, u64 halo, int device, bool is_shard,
                             const i32 *host_sa, gsa_index **out, gsa_build_stats *stats) {
  if (out == nullptr) return GSA_EINVAL;
  *out = nullptr;
  if (stats) memset(stats, 0, sizeof(*stats));
  if ((T_full == nullptr && n_full > 0) || offset > n_full || len > n_full - offset) return GSA_EINVAL;
  if (len >= 0x7fffffffull) {  // crates/divsufsort/src/divsufsort.rs:9-13: text.len() < i32::MAX
    set_error("text too large, should not exceed 2147483646 bytes", __FILE__, __LINE__);
    return GSA_EINVAL;
  }
  DeviceGuard dg(device);
  if (!dg.ok) return GSA_ECUDA;
  halo = std::min<u64>(halo, n_full - offset - len);
  gsa_index *ix = new (std::nothrow) gsa_index();
  if (!ix) return GSA_ENOMEM;
  ix->device = device;
  ix->n = len;
  ix->text_avail = len + halo;
  if (is_shard) { ix->host_full = T_full; ix->n_full = n_full; ix->offset = offset; }
  DevBuf<u8> d_T;
  DevBuf<i32> d_SA;
  int rc = d_T.alloc((size_t)(len + halo) + 64);
  if (rc == GSA_OK) rc = d_SA.alloc((size_t)len);
  Stream st;
  if (rc == GSA_OK) rc = st.create();
  auto body = [&]() -> int {
    if (len + halo > 0) GSA_TRY(cudaMemcpyAsync(d_T.p, T_full + offset, (size_t)(len + halo), cudaMemcpyHostToDevice, st.s));
    if (host_sa) {
      if (len > 0) GSA_TRY(cudaMemcpyAsync(d_SA.p, host_sa, (size_t)len * 4, cudaMemcpyHostToDevice, st.s));
      // a caller-supplied SA: every entry must be a text position, or the search kernels would read
      // outside the text (the reference panics on its slice bounds check instead)
      i64 bad = -1;
      GSA_TRY_RC(sa_range_check_device(d_SA.p, (u32)len, st.s, &bad));
      if (bad >= 0) {
        char msg[128];
        snprintf(msg, sizeof msg, "suffix array entry %lld is out of range for a text of %llu bytes", (long long)bad, (unsigned long long)len);
        set_error(msg, __FILE__, __LINE__);
        return GSA_EPANIC;
      }
      return GSA_OK;
    }
    // the sort workspace comes from the per-device scratch cache (a cudaMalloc / cudaFree pair of ~66 bytes per
    // text byte costs more than a shard's upload); a concurrent build on the same device gets a block of its own
    const size_t ws_bytes = build_workspace_bytes((u32)len);
    Scratch sc;
    GSA_TRY_RC(sc.acquire(device, ws_bytes));
    return build_sa_device(d_T.p, d_SA.p, (u32)len, sc.p, sc.bytes, st.s, stats);
  };
  if (rc == GSA_OK) rc = body();
  if (rc != GSA_OK) { delete ix; return rc; }
  ix->d_text = d_T.release();
  ix->d_sa = d_SA.release();
  *out = ix;
  return GSA_OK;
}

int32_t gsa_index_create(const uint8_t *T, int64_t n, int32_t device, gsa_index **out, gsa_build_stats *stats) {
  if (n < 0) return GSA_EINVAL;
  return index_create_impl(T, (u64)n, 0, (u64)n, 0, device, false, nullptr, out, stats);
}

int32_t gsa_index_from_parts(const uint8_t *T, int64_t n, const int32_t *SA, int64_t sa_len, int32_t device,
                             gsa_index **out) {
  if (n < 0 || (n > 0 && SA == nullptr)) return GSA_EINVAL;
  if (sa_len != n) {
    set_error("text and suffix array should have same len", __FILE__, __LINE__);
    return GSA_EINVAL;
  }
  static const i32 dummy = 0;
  return index_create_impl(T, (u64)n, 0, (u64)n, 0, device, false, n > 0 ? SA : &dummy, out, nullptr);
}

int32_t gsa_index_create_shard(const uint8_t *T_full, uint64_t n_full, uint64_t offset, uint64_t len, uint64_t halo,
                               int32_t device, gsa_index **out, gsa_build_stats *stats) {
  return index_create_impl(T_full, n_full, offset, len, halo, device, true, nullptr, out, stats);
}

int64_t gsa_index_len(const gsa_index *ix) { return ix ? (int64_t)ix->n : -1; }
int32_t gsa_index_device(const gsa_index *ix) { return ix ? ix->device : -1; }
const uint8_t *gsa_index_device_text(const gsa_index *ix) { return ix ? ix->d_text : nullptr; }
const int32_t *gsa_index_device_sa(const gsa_index *ix) { return ix ? ix->d_sa : nullptr; }

int32_t gsa_index_sa(const gsa_index *ix, int32_t *out_sa) {
  if (!ix || (ix->n > 0 && !out_sa)) return GSA_EINVAL;
  if (ix->n == 0) return GSA_OK;
  DeviceGuard dg(ix->device);
  if (!dg.ok) return GSA_ECUDA;
  GSA_TRY(cudaMemcpy(out_sa, ix->d_sa, (size_t)ix->n * 4, cudaMemcpyDeviceToHost));
  return GSA_OK;
}

int32_t gsa_index_verify(const gsa_index *ix, int64_t *bad_index) {
  if (!ix) return GSA_EINVAL;
  if (ix->n == 0) return GSA_EPANIC;  // sacabase lib.rs:143: `input.len() - 1` underflows
  DeviceGuard dg(ix->device);
  if (!dg.ok) return GSA_ECUDA;
  Stream st;
  GSA_TRY_RC(st.create());
  return sufcheck_device(ix->d_text, ix->d_sa, (u32)ix->n, st.s, bad_index);
}

void gsa_index_destroy(gsa_index *ix) {
  if (!ix) return;
  DeviceGuard dg(ix->device);
  if (ix->d_text) cudaFree(ix->d_text);
  if (ix->d_sa) cudaFree(ix->d_sa);
  if (ix->d_accel) cudaFree(ix->d_accel);
  delete ix;
}

// Grow the halo behind a shard so that the may_extend rule can read `want` bytes past its end.
static int ensure_halo(gsa_index *ix, u64 want) {
  if (!ix->host_full) return GSA_OK;
  const u64 room = ix->n_full - ix->offset - ix->n;
  want = std::min(want, room);
  if (ix->text_avail - ix->n >= want) return GSA_OK;
  DevBuf<u8> nt;
  GSA_TRY_RC(nt.alloc((size_t)(ix->n + want) + 64));
  GSA_TRY(cudaMemcpy(nt.p, ix->d_text, (size_t)ix->n, cudaMemcpyDeviceToDevice));
  GSA_TRY(cudaMemcpy(nt.p + ix->n, ix->host_full + ix->offset + ix->n, (size_t)want, cudaMemcpyHostToDevice));
  cudaFree(ix->d_text);
  ix->d_text = nt.release();
  ix->text_avail = ix->n + want;
  return GSA_OK;
}

// ------------------------------------------------------------------------------------------
// Batched search, host pointers
// ------------------------------------------------------------------------------------------
// The table is built on first use (index creation stays as cheap as the build itself); the caller has
// selected the index's device.  The search kernels work without it (k == 0) if the allocation fails.
static int ensure_accel(const gsa_index *cix) {
  gsa_index *ix = const_cast<gsa_index *>(cix);
  std::lock_guard<std::mutex> lk(ix->accel_mu);
  if (ix->accel_ready) return GSA_OK;
  DeviceGuard dg(ix->device);
  if (!dg.ok) return GSA_ECUDA;
  Stream st;
  GSA_TRY_RC(st.create());
  const int rc = accel_build_device(ix->d_text, ix->d_sa, ix->n, 0, &ix->accel, &ix->d_accel, st.s);
  if (rc == GSA_ENOMEM) {
    // no memory for the table: the searches work without it (they start at [0, n])
    memset(&ix->accel, 0, sizeof(ix->accel));
    ix->d_accel = nullptr;
  } else if (rc != GSA_OK) {
    return rc;
  }
  ix->accel_ready = true;
  return GSA_OK;
}

static TextView view_of(const gsa_index *ix) { return TextView{ix->d_text, ix->d_sa, ix->n, ix->text_avail, ix->accel}; }

struct PatternsOnDevice {
  DevBuf<u8> pats;
  DevBuf<u64> off;
  int upload(const u8 *h_pats, const u64 *h_off, u64 Q, cudaStream_t st) {
    const u64 bytes = h_off[Q];
    GSA_TRY_RC(pats.alloc((size_t)bytes + 64));
    GSA_TRY_RC(off.alloc((size_t)Q + 1));
    // the kernels read whole aligned words: the bytes behind the last pattern are defined (masked out, never compared)
    GSA_TRY(cudaMemsetAsync(pats.p + bytes, 0, 64, st));
    if (bytes) GSA_TRY(cudaMemcpyAsync(pats.p, h_pats, (size_t)bytes, cudaMemcpyHostToDevice, st));
    GSA_TRY(cudaMemcpyAsync(off.p, h_off, (size_t)(Q + 1) * sizeof(u64), cudaMemcpyHostToDevice, st));
    return GSA_OK;
  }
};

static int check_patterns(const u8 *pats, const u64 *pat_off, u64 Q) {
  if (Q == 0) return GSA_OK;
  if (!pat_off) return GSA_EINVAL;
  if (pat_off[Q] > 0 && !pats) return GSA_EINVAL;
  return GSA_OK;
}

}  // extern "C" (a template in between)

// Host-pointer batch in chunks of kQueryChunk patterns on two streams: the upload of chunk c + 1 and the
// download of chunk c - 1 overlap the kernel of chunk c (the three use different engines), the longest
// pattern of a chunk is found by the CPU while the previous chunk is in flight, and the device buffers come
// from the per-device scratch cache instead of four cudaMalloc / cudaFree pairs per call.
// launch(d_pats, d_off_of_chunk, Qc, max_len_of_chunk, d_out_a_of_chunk, d_out_b_of_chunk, stream).
constexpr u64 kQueryChunk = 1u << 20;

template <typename Launch>
static int chunked_query(const gsa_index *ix, const u8 *pats, const u64 *pat_off, u64 Q, size_t a_bytes, void *out_a,
                         void *out_b, Launch launch) {
  Stream st[2];
  GSA_TRY_RC(st[0].create());
  GSA_TRY_RC(st[1].create());
  const u64 bytes = pat_off[Q];
  const size_t pat_sz = align_up((size_t)bytes + 64, 256), off_sz = align_up((size_t)(Q + 1) * 8, 256);
  const size_t a_sz = align_up((size_t)Q * a_bytes, 256), b_sz = align_up((size_t)Q * 4, 256);
  Scratch sc;
  GSA_TRY_RC(sc.acquire(ix->device, pat_sz + off_sz + a_sz + b_sz));
  u8 *d_pats = reinterpret_cast<u8 *>(sc.p);
  u64 *d_off = reinterpret_cast<u64 *>(sc.p + pat_sz);
  char *d_a = sc.p + pat_sz + off_sz;
  char *d_b = d_a + a_sz;
  // the kernels read whole aligned words: the bytes behind the last pattern are defined (masked out, never compared)
  GSA_TRY(cudaMemsetAsync(d_pats + bytes, 0, 64, st[0].s));
  GSA_TRY(cudaStreamSynchronize(st[0].s));
  u64 k = 0;
  for (u64 c0 = 0; c0 < Q; c0 += kQueryChunk, ++k) {
    const u64 c1 = std::min(Q, c0 + kQueryChunk), qc = c1 - c0;
    cudaStream_t s = st[k & 1].s;
    const u64 b0 = pat_off[c0], b1 = pat_off[c1];
    if (b1 > b0) GSA_TRY(cudaMemcpyAsync(d_pats + b0, pats + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, s));
    GSA_TRY(cudaMemcpyAsync(d_off + c0, pat_off + c0, (size_t)(qc + 1) * 8, cudaMemcpyHostToDevice, s));
    u64 m = 0;
    for (u64 q = c0; q < c1; ++q) m = std::max(m, pat_off[q + 1] - pat_off[q]);
    GSA_TRY_RC(launch(d_pats, d_off + c0, qc, (u32)std::min<u64>(m, 0xffffffffull), d_a + c0 * a_bytes, d_b + c0 * 4, s));
    GSA_TRY(cudaMemcpyAsync(static_cast<char *>(out_a) + c0 * a_bytes, d_a + c0 * a_bytes, (size_t)qc * a_bytes, cudaMemcpyDeviceToHost, s));
    GSA_TRY(cudaMemcpyAsync(static_cast<char *>(out_b) + c0 * 4, d_b + c0 * 4, (size_t)qc * 4, cudaMemcpyDeviceToHost, s));
  }
  GSA_TRY(cudaStreamSynchronize(st[0].s));
  GSA_TRY(cudaStreamSynchronize(st[1].s));
  return GSA_OK;
}

extern "C" {

int32_t gsa_lsm_batch(const gsa_index *ix, const uint8_t *pats, const uint64_t *pat_off, uint64_t Q,
                      uint64_t *out_start, uint32_t *out_len) {
  if (!ix || (Q > 0 && (!out_start || !out_len))) return GSA_EINVAL;
  GSA_TRY_RC(check_patterns(pats, pat_off, Q));
  if (ix->n == 0) return GSA_EPANIC;
  if (Q == 0) return GSA_OK;
  DeviceGuard dg(ix->device);
  if (!dg.ok) return GSA_ECUDA;
  GSA_TRY_RC(ensure_accel(ix));
  const TextView tv = view_of(ix);
  return chunked_query(ix, pats, pat_off, Q, 8, out_start, out_len,
                       [&](const u8 *dp, const u64 *doff, u64 qc, u32 ml, char *da, char *db, cudaStream_t s) {
                         return lsm_device(tv, dp, doff, qc, ml, 0, 0, reinterpret_cast<u64 *>(da), reinterpret_cast<u32 *>(db), s);
                       });
}

int32_t gsa_search_all_batch(const gsa_index *ix, const uint8_t *pats, const uint64_t *pat_off, uint64_t Q,
                             int32_t *out_left, int32_t *out_count) {
  if (!ix || (Q > 0 && (!out_left || !out_count))) return GSA_EINVAL;
  GSA_TRY_RC(check_patterns(pats, pat_off, Q));
  if (Q == 0) return GSA_OK;
  if (ix->n == 0) {  // utils.c:269,272: idx = -1, count 0 on an empty text / SA
    for (u64 q = 0; q < Q; ++q) { out_left[q] = -1; out_count[q] = 0; }
    return GSA_OK;
  }
  DeviceGuard dg(ix->device);
  if (!dg.ok) return GSA_ECUDA;
  GSA_TRY_RC(ensure_accel(ix));
  const TextView tv = view_of(ix);
  return chunked_query(ix, pats, pat_off, Q, 4, out_left, out_count,
                       [&](const u8 *dp, const u64 *doff, u64 qc, u32 ml, char *da, char *db, cudaStream_t s) {
                         return search_all_device(tv, dp, doff, qc, ml, reinterpret_cast<i32 *>(da), reinterpret_cast<i32 *>(db), s);
                       });
}

int32_t gsa_contains_batch(const gsa_index *ix, const uint8_t *pats, const uint64_t *pat_off, uint64_t Q,
                           uint8_t *out) try {
  if (Q > 0 && !out) return GSA_EINVAL;
  std::vector<i32> left((size_t)Q), count((size_t)Q);
  GSA_TRY_RC(gsa_search_all_batch(ix, pats, pat_off, Q, left.data(), count.data()));
  for (u64 q = 0; q < Q; ++q) out[q] = count[q] > 0;
  return GSA_OK;
} catch (const std::bad_alloc &) {  // nothing may escape the C ABI
  gsa::set_error("out of host memory", __FILE__, __LINE__);
  return GSA_ENOMEM;
} catch (...) {
  gsa::set_error("unexpected C++ exception", __FILE__, __LINE__);
  return GSA_ECUDA;
}

int32_t gsa_lsm_device(const gsa_index *ix, const uint8_t *d_pats, const uint64_t *d_pat_off, uint64_t Q,
                       uint32_t max_pat_len, uint64_t offset, int32_t accumulate, uint64_t *d_io_start,
                       uint32_t *d_io_len, void *stream) {
  if (!ix || (Q > 0 && (!d_pat_off || !d_io_start || !d_io_len))) return GSA_EINVAL;
  if (Q > 0 && ix->n > 0) GSA_TRY_RC(ensure_accel(ix));
  return lsm_device(view_of(ix), d_pats, d_pat_off, Q, max_pat_len, offset, accumulate, d_io_start, d_io_len,
                    static_cast<cudaStream_t>(stream));
}

int32_t gsa_search_all_device(const gsa_index *ix, const uint8_t *d_pats, const uint64_t *d_pat_off, uint64_t Q,
                              uint32_t max_pat_len, int32_t *d_out_left, int32_t *d_out_count, void *stream) {
  if (!ix || (Q > 0 && (!d_pat_off || !d_out_left || !d_out_count))) return GSA_EINVAL;
  if (ix->n == 0) return GSA_EINVAL;
  if (Q > 0) GSA_TRY_RC(ensure_accel(ix));
  return search_all_device(view_of(ix), d_pats, d_pat_off, Q, max_pat_len, d_out_left, d_out_count,
                           static_cast<cudaStream_t>(stream));
}

int32_t gsa_lsm_reduce_device(uint64_t *d_start, uint32_t *d_len, uint64_t Q, uint32_t nsets, void *stream) {
  if (Q > 0 && (!d_start || !d_len)) return GSA_EINVAL;
  return lsm_reduce_device(d_start, d_len, Q, nsets, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------
// Partitioned suffix array (sacapart)
// ------------------------------------------------------------------------------------------
static const u64 kDefaultHalo = 4096;

int32_t gsa_part_create(const uint8_t *T, uint64_t n, uint64_t num_partitions, const int32_t *devices, int32_t ndev,
                        gsa_part **out) try {
  if (!out) return GSA_EINVAL;
  *out = nullptr;
  if (T == nullptr && n > 0) return GSA_EINVAL;
  if (num_partitions == 0) {  // sacapart lib.rs:43 divides by zero
    set_error("attempt to divide by zero (num_partitions == 0)", __FILE__, __LINE__);
    return GSA_EPANIC;
  }
  struct PartFree { void operator()(gsa_part *q) const { gsa_part_destroy(q); } };
  std::unique_ptr<gsa_part, PartFree> owner(new (std::nothrow) gsa_part());  // freed with its shards on every error path
  gsa_part *p = owner.get();
  if (!p) return GSA_ENOMEM;
  p->text = T;
  p->n = n;
  p->partition_size = n / num_partitions + 1;                        // lib.rs:43
  const u64 nparts = (n + p->partition_size - 1) / p->partition_size;  // par_chunks, lib.rs:45-49
  if (devices && ndev > 0) p->devices.assign(devices, devices + ndev);
  else p->devices.push_back(current_device());
  p->shards.assign((size_t)nparts, nullptr);
  const size_t nd = p->devices.size();
  std::vector<int> rcs(nd, GSA_OK);
  std::vector<std::string> errs(nd);
  // One host thread per device (rayon's par_chunks in the reference); shards that share a
  // device are built one after the other so each can use the whole GPU and its workspace.
  auto worker = [&](size_t k) {
    for (u64 i = k; i < nparts; i += nd) {
      const u64 off = i * p->partition_size;
      const u64 len = std::min(p->partition_size, n - off);
      int rc = gsa_index_create_shard(T, n, off, len, kDefaultHalo, p->devices[k], &p->shards[(size_t)i], nullptr);
      if (rc != GSA_OK) { rcs[k] = rc; errs[k] = g_last_error; return; }
    }
  };
  if (nd == 1) {
    worker(0);
  } else {
    // joins whatever was started, also when a later spawn throws (a joinable std::thread that is
    // destroyed calls std::terminate before any handler of this function runs)
    struct Joiner {
      std::vector<std::thread> th;
      ~Joiner() { for (auto &t : th) if (t.joinable()) t.join(); }
    } j;
    j.th.reserve(nd);
    for (size_t k = 0; k < nd; ++k) {
      try {
        j.th.emplace_back(worker, k);
      } catch (...) {  // no thread to be had: this one builds the device's shards itself
        worker(k);
      }
    }
  }
  for (size_t k = 0; k < nd; ++k)
    if (rcs[k] != GSA_OK) {
      g_last_error = errs[k];
      return rcs[k];
    }
  *out = owner.release();
  return GSA_OK;
} catch (const std::bad_alloc &) {  // nothing may escape the C ABI
  gsa::set_error("out of host memory", __FILE__, __LINE__);
  return GSA_ENOMEM;
} catch (...) {
  gsa::set_error("unexpected C++ exception", __FILE__, __LINE__);
  return GSA_ECUDA;
}

uint64_t gsa_part_num_partitions(const gsa_part *p) { return p ? p->shards.size() : 0; }
uint64_t gsa_part_partition_size(const gsa_part *p) { return p ? p->partition_size : 0; }
const gsa_index *gsa_part_shard(const gsa_part *p, uint64_t i) {
  return (p && i < p->shards.size()) ? p->shards[(size_t)i] : nullptr;
}

void gsa_part_destroy(gsa_part *p) {
  if (!p) return;
  for (gsa_index *ix : p->shards) gsa_index_destroy(ix);
  delete p;
}

int32_t gsa_part_lsm_batch(gsa_part *p, const uint8_t *pats, const uint64_t *pat_off, uint64_t Q, uint64_t *out_start,
                           uint32_t *out_len) try {
  if (!p || (Q > 0 && (!out_start || !out_len))) return GSA_EINVAL;
  GSA_TRY_RC(check_patterns(pats, pat_off, Q));
  if (p->shards.empty()) {  // lib.rs:94-96: expect() on None
    set_error("partitioned suffix arrays should always find at least one longest common substring", __FILE__, __LINE__);
    return GSA_EPANIC;
  }
  if (Q == 0) return GSA_OK;
  u64 max_len = 0;
  for (u64 q = 0; q < Q; ++q) max_len = std::max(max_len, pat_off[q + 1] - pat_off[q]);
  std::lock_guard<std::mutex> one_query(p->query_mu);  // ensure_halo() below may replace a shard's text buffer

  const size_t nd = p->devices.size();
  struct PerDev {
    Stream st;
    PatternsOnDevice pd;
    DevBuf<u64> start;
    DevBuf<u32> len;
    bool used = false;
  };
  std::vector<PerDev> dev(nd);
  // fan out: every device answers for its shards, in ascending shard order (strict-greater
  // replacement then keeps the earliest partition, lib.rs:86-92)
  for (size_t k = 0; k < nd; ++k) {
    if (k >= p->shards.size()) break;
    DeviceGuard dg(p->devices[k]);
    if (!dg.ok) return GSA_ECUDA;
    PerDev &d = dev[k];
    d.used = true;
    GSA_TRY_RC(d.st.create());
    GSA_TRY_RC(d.pd.upload(pats, pat_off, Q, d.st.s));
    // device 0 receives every device's result set for the final merge
    GSA_TRY_RC(d.start.alloc((size_t)Q * (k == 0 ? nd : 1)));
    GSA_TRY_RC(d.len.alloc((size_t)Q * (k == 0 ? nd : 1)));
    bool first = true;
    for (size_t i = k; i < p->shards.size(); i += nd) {
      gsa_index *ix = p->shards[i];
      GSA_TRY_RC(ensure_halo(ix, max_len));
      GSA_TRY_RC(ensure_accel(ix));
      GSA_TRY_RC(lsm_device(view_of(ix), d.pd.pats.p, d.pd.off.p, Q, (u32)std::min<u64>(max_len, 0xffffffffull), ix->offset, first ? 0 : 1, d.start.p, d.len.p,
                            d.st.s));
      first = false;
    }
  }
  u32 nsets = 0;
  for (size_t k = 0; k < nd; ++k) {
    if (!dev[k].used) continue;
    DeviceGuard dg(p->devices[k]);
    GSA_TRY(cudaStreamSynchronize(dev[k].st.s));
    ++nsets;
  }
  // gather to device 0 and merge there
  DeviceGuard dg0(p->devices[0]);
  for (size_t k = 1; k < nd; ++k) {
    if (!dev[k].used) continue;
    GSA_TRY(cudaMemcpyPeerAsync(dev[0].start.p + (size_t)k * Q, p->devices[0], dev[k].start.p, p->devices[k], (size_t)Q * 8, dev[0].st.s));
    GSA_TRY(cudaMemcpyPeerAsync(dev[0].len.p + (size_t)k * Q, p->devices[0], dev[k].len.p, p->devices[k], (size_t)Q * 4, dev[0].st.s));
  }
  GSA_TRY_RC(lsm_reduce_device(dev[0].start.p, dev[0].len.p, Q, nsets, dev[0].st.s));
  GSA_TRY(cudaMemcpyAsync(out_start, dev[0].start.p, (size_t)Q * 8, cudaMemcpyDeviceToHost, dev[0].st.s));
  GSA_TRY(cudaMemcpyAsync(out_len, dev[0].len.p, (size_t)Q * 4, cudaMemcpyDeviceToHost, dev[0].st.s));
  GSA_TRY(cudaStreamSynchronize(dev[0].st.s));
  return GSA_OK;
} catch (const std::bad_alloc &) {  // nothing may escape the C ABI
  gsa::set_error("out of host memory", __FILE__, __LINE__);
  return GSA_ENOMEM;
} catch (...) {
  gsa::set_error("unexpected C++ exception", __FILE__, __LINE__);
  return GSA_ECUDA;
}

}  // extern "C"
