// search.cu -- batched suffix-array search kernels: one warp per pattern, warp-cooperative
// byte comparison against text resident in HBM.
//
//   k_lsm         sacabase::longest_substring_match (reference:
//                 crates/sacabase/src/lib.rs:39-99) with sacapart's per-shard step
//                 (crates/sacapart/src/lib.rs:71-92) folded in: offset, may_extend,
//                 strict-greater replacement.
//   k_search_all  libdivsufsort sa_search (crates/cdivsufsort/c-sources/utils.c:244-325):
//                 (left, count) of the SA range whose suffixes start with the pattern.
//   k_lsm_reduce  best-of merge of fanned-out partition results (sacapart lib.rs:86-92).
//
// Per query ~log2(n) dependent steps, each one 4-byte SA read and one <= m-byte text read
// (two 32-byte sectors): latency bound, so the kernel runs as many warps as fit.
#include "builder.h"

namespace gsa {

// Compares pattern P[0,m) with the text bytes [s, tend), starting at byte `from` (the first
// `from` bytes are known to be equal).  Warp-cooperative; every lane returns the same values.
//   cpl = length of the common prefix (<= min(m, tend - s))
//   gt  = pattern > suffix in Rust slice order / sa_search's r < 0
//         (first differing byte larger, or the suffix is a proper prefix of the pattern)
//   lt  = pattern < suffix at a differing byte (sa_search's r > 0)
struct CmpResult {
  u32 cpl;
  bool gt;
  bool lt;
};

__device__ __forceinline__ CmpResult warp_compare(const u8 *__restrict__ text, u64 s, u64 tend,
                                                  const u8 *__restrict__ pat, u32 m, u32 pat_lane0, u32 from) {
  const u32 lane = lane_id();
  const u64 rem = tend - s;
  const u32 lim = (rem < (u64)m) ? (u32)rem : m;
  CmpResult r;
  for (u32 off = from & ~31u; off < lim; off += 32) {
    const u32 i = off + lane;
    u32 pb = 0, tb = 0;
    if (i < lim) {
      pb = (off == 0) ? pat_lane0 : (u32)__ldg(pat + i);
      tb = (u32)__ldg(text + s + i);
    }
    const u32 neq = __ballot_sync(0xffffffffu, pb != tb);
    if (neq) {
      const int first = __ffs(neq) - 1;
      r.cpl = off + (u32)first;
      const u32 p1 = __shfl_sync(0xffffffffu, pb, first);
      const u32 t1 = __shfl_sync(0xffffffffu, tb, first);
      r.gt = p1 > t1;
      r.lt = p1 < t1;
      return r;
    }
  }
  r.cpl = lim;
  r.gt = m > lim;  // text ran out first: the suffix is a proper prefix of the pattern
  r.lt = false;
  return r;
}

struct LsmArgs {
  const u8 *text;
  const i32 *sa;
  u64 n;           // suffix-array length == logical text length
  u64 text_avail;  // readable text bytes (>= n): shard + halo
  const u8 *pats;
  const u64 *pat_off;
  u64 Q;
  u64 offset;      // added to start (partition offset)
  int accumulate;  // keep the previous (start,len) unless strictly longer
  u64 *io_start;
  u32 *io_len;
};

__global__ void __launch_bounds__(256) k_lsm(const LsmArgs a) {
  const u64 q = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= a.Q) return;
  const u32 lane = lane_id();
  const u64 p0 = a.pat_off[q];
  const u32 m = (u32)(a.pat_off[q + 1] - p0);
  const u8 *pat = a.pats + p0;
  const u32 pl0 = (lane < m) ? (u32)__ldg(pat + lane) : 0u;

  // sacabase lib.rs:75-98 on the window sa[lo .. lo+w)
  u64 lo = 0, w = a.n;
  while (w > 2) {
    const u64 mid = w >> 1;
    const u64 s = (u64)(u32)__ldg(a.sa + lo + mid);
    const CmpResult c = warp_compare(a.text, s, a.n, pat, m, pl0, 0);
    if (c.gt) { lo += mid; w -= mid; } else { w = mid + 1; }
  }
  u64 start = (u64)(u32)__ldg(a.sa + lo);
  u32 len = warp_compare(a.text, start, a.n, pat, m, pl0, 0).cpl;
  if (w == 2) {
    const u64 s1 = (u64)(u32)__ldg(a.sa + lo + 1);
    const u32 y = warp_compare(a.text, s1, a.n, pat, m, pl0, 0).cpl;
    if (!(len > y)) { start = s1; len = y; }  // `x > y` keeps the first, ties go to the second
  }
  // sacapart lib.rs:77-84: a match that touches the end of the shard may continue behind it
  if (start + len == a.n && a.text_avail > a.n) len = warp_compare(a.text, start, a.text_avail, pat, m, pl0, len).cpl;
  if (lane == 0) {
    if (!a.accumulate || len > a.io_len[q]) {  // lib.rs:86-92, strict
      a.io_start[q] = start + a.offset;
      a.io_len[q] = len;
    }
  }
}

struct SearchAllArgs {
  const u8 *text;
  const i32 *sa;
  u64 n;
  const u8 *pats;
  const u64 *pat_off;
  u64 Q;
  i32 *left;
  i32 *count;
};

__global__ void __launch_bounds__(256) k_search_all(const SearchAllArgs a) {
  const u64 q = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= a.Q) return;
  const u32 lane = lane_id();
  const u64 p0 = a.pat_off[q];
  const u32 m = (u32)(a.pat_off[q + 1] - p0);
  if (m == 0) {  // utils.c:273
    if (lane == 0) { a.left[q] = 0; a.count[q] = (i32)a.n; }
    return;
  }
  const u8 *pat = a.pats + p0;
  const u32 pl0 = (lane < m) ? (u32)__ldg(pat + lane) : 0u;
  // lower bound: suffixes with r < 0   (suffix < pattern)
  u64 lo = 0, hi = a.n;
  while (lo < hi) {
    const u64 mid = (lo + hi) >> 1;
    const CmpResult c = warp_compare(a.text, (u64)(u32)__ldg(a.sa + mid), a.n, pat, m, pl0, 0);
    if (c.gt) lo = mid + 1; else hi = mid;
  }
  const u64 left = lo;
  // upper bound: suffixes with r <= 0  (suffix < pattern, or pattern is a prefix of it)
  hi = a.n;
  while (lo < hi) {
    const u64 mid = (lo + hi) >> 1;
    const CmpResult c = warp_compare(a.text, (u64)(u32)__ldg(a.sa + mid), a.n, pat, m, pl0, 0);
    if (!c.lt) lo = mid + 1; else hi = mid;
  }
  if (lane == 0) {
    a.left[q] = (i32)left;  // first match, or the insertion point on a miss (utils.c:323)
    a.count[q] = (i32)(lo - left);
  }
}

__global__ void __launch_bounds__(256) k_lsm_reduce(u64 *__restrict__ start, u32 *__restrict__ len, u64 Q, u32 nsets) {
  const u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  u64 bs = start[q];
  u32 bl = len[q];
  for (u32 s = 1; s < nsets; ++s) {
    const u64 cs = start[(u64)s * Q + q];
    const u32 cl = len[(u64)s * Q + q];
    if (cl > bl || (cl == bl && cs < bs)) { bs = cs; bl = cl; }
  }
  start[q] = bs;
  len[q] = bl;
}

int lsm_device(const TextView &tv, const u8 *d_pats, const u64 *d_pat_off, u64 Q, u64 offset, int accumulate,
               u64 *d_io_start, u32 *d_io_len, cudaStream_t st) {
  if (Q == 0) return GSA_OK;
  if (tv.n == 0) return GSA_EPANIC;  // sacabase lib.rs:89-91 indexes sa[0]
  LsmArgs a{tv.text, tv.sa, tv.n, tv.text_avail, d_pats, d_pat_off, Q, offset, accumulate, d_io_start, d_io_len};
  const u64 blocks = div_up(Q * 32, 256);
  k_lsm<<<(unsigned)blocks, 256, 0, st>>>(a);
  GSA_TRY(cudaGetLastError());
  return GSA_OK;
}

int search_all_device(const TextView &tv, const u8 *d_pats, const u64 *d_pat_off, u64 Q, i32 *d_left, i32 *d_count,
                      cudaStream_t st) {
  if (Q == 0) return GSA_OK;
  SearchAllArgs a{tv.text, tv.sa, tv.n, d_pats, d_pat_off, Q, d_left, d_count};
  const u64 blocks = div_up(Q * 32, 256);
  k_search_all<<<(unsigned)blocks, 256, 0, st>>>(a);
  GSA_TRY(cudaGetLastError());
  return GSA_OK;
}

int lsm_reduce_device(u64 *d_start, u32 *d_len, u64 Q, u32 nsets, cudaStream_t st) {
  if (Q == 0 || nsets <= 1) return GSA_OK;
  k_lsm_reduce<<<(unsigned)div_up(Q, 256), 256, 0, st>>>(d_start, d_len, Q, nsets);
  GSA_TRY(cudaGetLastError());
  return GSA_OK;
}

}  // namespace gsa
