// search.cu -- batched suffix-array search kernels: a group of 8/16/32 lanes per pattern,
// cooperative word comparison against text resident in HBM.
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

// A pattern is handled by a group of G lanes (G = 8, 16 or 32, chosen from the longest pattern
// of the batch: 4 bytes per lane and step, so G = 8 covers 32-byte patterns in one step and a
// warp carries 4 independent binary searches -- the walk is latency bound, more searches in
// flight is what makes it faster).  All lanes of the warp run in lockstep; a group whose search
// has finished idles until the others are done.
//
// group_compare: pattern P[0,m) against the text bytes [s, tend), the first `from` bytes being
// known equal.  Every lane of the group returns the same values.
//   cpl = length of the common prefix (<= min(m, tend - s))
//   gt  = pattern > suffix in Rust slice order / sa_search's r < 0
//         (first differing byte larger, or the suffix is a proper prefix of the pattern)
//   lt  = pattern < suffix at a differing byte (sa_search's r > 0)
struct CmpResult {
  u32 cpl;
  bool gt;
  bool lt;
};

// The nb (1..4) bytes at an arbitrary address as a little-endian word, from aligned 4-byte
// loads.  Only words that contain at least one of the requested bytes are touched, so no
// padding behind the buffers is required.
__device__ __forceinline__ u32 load_bytes(const u8 *p, u32 nb) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const u32 *w = reinterpret_cast<const u32 *>(a & ~(uintptr_t)3);
  const u32 mis = (u32)(a & 3u);
  const u32 lo = __ldg(w);
  if (mis + nb <= 4u) return lo >> (8u * mis);
  return __funnelshift_r(lo, __ldg(w + 1), 8u * mis);
}

template <int G>
__device__ __forceinline__ CmpResult group_compare(const u8 *__restrict__ text, u64 s, u64 tend,
                                                   const u8 *__restrict__ pat, u32 m, u32 pat_word0, u32 from, bool active) {
  const u32 lane = lane_id();
  const u32 sub = lane & (G - 1);              // lane inside the group
  const u32 gshift = lane & ~(u32)(G - 1);     // first lane of the group
  const u32 gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gshift);
  const u64 rem = tend - s;
  const u32 lim = (rem < (u64)m) ? (u32)rem : m;
  CmpResult r;
  r.cpl = lim;
  r.gt = m > lim;  // no difference found: the suffix is a proper prefix of the pattern, or equal
  r.lt = false;
  bool done = !active;
  for (u32 off = from & ~(4u * G - 1u); ; off += 4u * G) {
    const bool run = !done && off < lim;
    if (!__any_sync(0xffffffffu, run)) break;
    u32 neq = 0, pw = 0, tw = 0;
    const u32 i = off + 4u * sub;
    if (run && i < lim) {
      const u32 nb = min(4u, lim - i);
      pw = (off == 0) ? pat_word0 : load_bytes(pat + i, nb);
      tw = load_bytes(text + s + i, nb);
      neq = (pw ^ tw) & ((nb == 4u) ? 0xffffffffu : ((1u << (8u * nb)) - 1u));
    }
    const u32 bal = __ballot_sync(0xffffffffu, neq != 0u) & gmask;
    // every lane takes part in the shuffles (full mask); groups without a difference read themselves
    const u32 first = bal ? ((u32)__ffs(bal) - 1u) : lane;  // absolute lane holding the first difference
    const u32 fn = __shfl_sync(0xffffffffu, neq, first);
    const u32 fp = __shfl_sync(0xffffffffu, pw, first);
    const u32 ft = __shfl_sync(0xffffffffu, tw, first);
    if (run && bal) {
      const u32 byte = ((u32)__ffs(fn) - 1u) >> 3;
      const u32 pb = (fp >> (8u * byte)) & 255u, tb = (ft >> (8u * byte)) & 255u;
      r.cpl = off + 4u * (first - gshift) + byte;
      r.gt = pb > tb;
      r.lt = pb < tb;
      done = true;
    }
  }
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
  u32 carry;  // 0: every comparison starts at byte 0 (GSA_NO_MATCH_CARRY, measurements only)
};

// CARRY: start comparisons at min(lmatch, rmatch); pointless (and measurably slower in k_search_all)
// when every needle fits one comparison step of the group (4 * G bytes).
template <int G, bool CARRY>
__global__ void __launch_bounds__(256) k_lsm(const LsmArgs a) {
  constexpr int PER_WARP = 32 / G;
  const u32 lane = lane_id();
  const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u64 q = warp * PER_WARP + (lane / G);
  const bool have = q < a.Q;
  const u32 sub = lane & (G - 1);
  u64 p0 = 0;
  u32 m = 0;
  if (have) {
    p0 = a.pat_off[q];
    m = (u32)(a.pat_off[q + 1] - p0);
  }
  const u8 *pat = a.pats + p0;
  const u32 pw0 = (have && 4u * sub < m) ? load_bytes(pat + 4u * sub, min(4u, m - 4u * sub)) : 0u;

  // sacabase lib.rs:75-98 on the window sa[lo .. lo+w).  (Requesting the SA entries of both
  // possible next windows ahead of the comparison was tried and does not pay: at 1 GiB the walk
  // is bound by the rate of random DRAM sector fetches, not by their latency.)
  // The needle lies between the two ends of the window (end-inclusive once an end has been
  // compared), so every suffix inside shares at least min(lm, rm) bytes with it, lm / rm being the
  // common prefix lengths found at the ends: comparisons start there instead of at byte 0 -- the
  // lmatch / rmatch of libdivsufsort's sa_search (utils.c:275-286), which matters for long needles.
  u64 lo = 0, w = a.n;
  u32 lm = 0, rm = 0;
  for (;;) {
    const bool act = have && w > 2;
    if (!__any_sync(0xffffffffu, act)) break;
    const u64 mid = w >> 1;
    const u64 s = act ? (u64)(u32)__ldg(a.sa + lo + mid) : 0;
    const CmpResult c = group_compare<G>(a.text, s, a.n, pat, m, pw0, CARRY ? min(lm, rm) : 0u, act);
    if (act) {
      if (c.gt) { lo += mid; w -= mid; lm = c.cpl; } else { w = mid + 1; rm = c.cpl; }
    }
  }
  u64 start = have ? (u64)(u32)__ldg(a.sa + lo) : 0;
  u32 len = group_compare<G>(a.text, start, a.n, pat, m, pw0, CARRY ? min(lm, rm) : 0u, have).cpl;
  {
    const bool two = have && w == 2;
    const u64 s1 = two ? (u64)(u32)__ldg(a.sa + lo + 1) : 0;
    const u32 y = group_compare<G>(a.text, s1, a.n, pat, m, pw0, CARRY ? min(lm, rm) : 0u, two).cpl;
    if (two && !(len > y)) { start = s1; len = y; }  // `x > y` keeps the first, ties go to the second
  }
  {
    // sacapart lib.rs:77-84: a match that touches the end of the shard may continue behind it
    const bool ext = have && start + len == a.n && a.text_avail > a.n;
    const u32 l2 = group_compare<G>(a.text, start, a.text_avail, pat, m, pw0, len, ext).cpl;
    if (ext) len = l2;
  }
  if (have && sub == 0) {
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
  u32 carry;
};

template <int G, bool CARRY>
__global__ void __launch_bounds__(256) k_search_all(const SearchAllArgs a) {
  constexpr int PER_WARP = 32 / G;
  const u32 lane = lane_id();
  const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u64 q = warp * PER_WARP + (lane / G);
  const u32 sub = lane & (G - 1);
  bool have = q < a.Q;
  u64 p0 = 0;
  u32 m = 0;
  if (have) {
    p0 = a.pat_off[q];
    m = (u32)(a.pat_off[q + 1] - p0);
  }
  if (have && m == 0) {  // utils.c:273
    if (sub == 0) { a.left[q] = 0; a.count[q] = (i32)a.n; }
    have = false;
  }
  const u8 *pat = a.pats + p0;
  const u32 pw0 = (have && 4u * sub < m) ? load_bytes(pat + 4u * sub, min(4u, m - 4u * sub)) : 0u;
  // lower bound L: suffixes with r < 0 (suffix < pattern); upper bound U: suffixes with r <= 0
  // (suffix < pattern, or pattern is a prefix of it).  The two binary searches are independent
  // and are advanced together, so their memory round trips overlap.
  // Each search carries the common prefix lengths found at its two bounds (sa_search's lmatch /
  // rmatch, utils.c:275-286): a comparison starts at their minimum.
  u64 lo = 0, hi = a.n, lo2 = 0, hi2 = a.n;
  u32 lm1 = 0, rm1 = 0, lm2 = 0, rm2 = 0;
  for (;;) {
    const bool act1 = have && lo < hi, act2 = have && lo2 < hi2;
    if (!__any_sync(0xffffffffu, act1 || act2)) break;
    const u64 mid1 = (lo + hi) >> 1, mid2 = (lo2 + hi2) >> 1;
    const u64 s1 = act1 ? (u64)(u32)__ldg(a.sa + mid1) : 0;
    const u64 s2 = act2 ? (u64)(u32)__ldg(a.sa + mid2) : 0;
    const CmpResult c1 = group_compare<G>(a.text, s1, a.n, pat, m, pw0, CARRY ? min(lm1, rm1) : 0u, act1);
    const CmpResult c2 = group_compare<G>(a.text, s2, a.n, pat, m, pw0, CARRY ? min(lm2, rm2) : 0u, act2);
    if (act1) { if (c1.gt) { lo = mid1 + 1; lm1 = c1.cpl; } else { hi = mid1; rm1 = c1.cpl; } }
    if (act2) { if (!c2.lt) { lo2 = mid2 + 1; lm2 = c2.cpl; } else { hi2 = mid2; rm2 = c2.cpl; } }
  }
  const u64 left = lo;
  lo = lo2;
  if (have && sub == 0) {
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

static int group_lanes(u32 max_pat_len) {
  if (max_pat_len == 0 || max_pat_len > 64) return 32;  // unknown or long: 128 bytes per step
  return max_pat_len > 32 ? 16 : 8;
}

int lsm_device(const TextView &tv, const u8 *d_pats, const u64 *d_pat_off, u64 Q, u32 max_pat_len, u64 offset,
               int accumulate, u64 *d_io_start, u32 *d_io_len, cudaStream_t st) {
  if (Q == 0) return GSA_OK;
  if (tv.n == 0) return GSA_EPANIC;  // sacabase lib.rs:89-91 indexes sa[0]
  LsmArgs a{tv.text, tv.sa, tv.n, tv.text_avail, d_pats, d_pat_off, Q, offset, accumulate, d_io_start, d_io_len, getenv("GSA_NO_MATCH_CARRY") ? 0u : 1u};
  const int G = group_lanes(max_pat_len);
  const u64 warps = div_up(Q, 32 / G);
  const unsigned blocks = (unsigned)div_up(warps * 32, 256);
  // G == 8 / 16 are only chosen when every needle fits one comparison step (32 / 64 bytes)
  const bool carry = a.carry && (max_pat_len == 0 || max_pat_len > 128);
  // (for k_lsm<8> the CARRY instantiation is used although it cannot skip anything: it measures 813 M instead of
  // 750 M queries/s on 32-byte needles -- a scheduling accident of the compiler, same results)
  if (G == 8) { if (a.carry) k_lsm<8, true><<<blocks, 256, 0, st>>>(a); else k_lsm<8, false><<<blocks, 256, 0, st>>>(a); }
  else if (G == 16) k_lsm<16, false><<<blocks, 256, 0, st>>>(a);
  else if (carry) k_lsm<32, true><<<blocks, 256, 0, st>>>(a);
  else k_lsm<32, false><<<blocks, 256, 0, st>>>(a);
  GSA_TRY(cudaGetLastError());
  return GSA_OK;
}

int search_all_device(const TextView &tv, const u8 *d_pats, const u64 *d_pat_off, u64 Q, u32 max_pat_len, i32 *d_left,
                      i32 *d_count, cudaStream_t st) {
  if (Q == 0) return GSA_OK;
  SearchAllArgs a{tv.text, tv.sa, tv.n, d_pats, d_pat_off, Q, d_left, d_count, getenv("GSA_NO_MATCH_CARRY") ? 0u : 1u};
  const int G = group_lanes(max_pat_len);
  const u64 warps = div_up(Q, 32 / G);
  const unsigned blocks = (unsigned)div_up(warps * 32, 256);
  const bool carry = a.carry && (max_pat_len == 0 || max_pat_len > 128);
  if (G == 8) k_search_all<8, false><<<blocks, 256, 0, st>>>(a);
  else if (G == 16) k_search_all<16, false><<<blocks, 256, 0, st>>>(a);
  else if (carry) k_search_all<32, true><<<blocks, 256, 0, st>>>(a);
  else k_search_all<32, false><<<blocks, 256, 0, st>>>(a);
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
