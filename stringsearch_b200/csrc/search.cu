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

// ------------------------------------------------------------------------------------------------
// Where a search starts: the prefix-bucket table (AccelView, builder.h).
//
// sacabase's narrowing loop (lib.rs:75-98) keeps a window [lo, lo + w) that always contains both
// clamp(ip - 1) and clamp(ip), ip = number of suffixes below the needle (it moves `lo` to an index
// whose suffix is below the needle and the last index to one that is not), and for n >= 2 it stops at
// w == 2 (w -> ceil(w / 2) or floor(w / 2) + 1, both >= 2 while w >= 3).  The final window is therefore
// [a, a + 1] with a = clamp(ip - 1, 0, n - 2) WHATEVER path led there, and the result is
// `cpl(a) > cpl(a + 1) ? a : a + 1`.  sa_search's (left, count) are bounds of the same kind.  So any
// way of finding ip gives the reference's answers, and the ~log2(n) top levels of the walk can be
// replaced by one table look-up on the needle's first k symbols:
//   lower bound in [T[key], T[key + 1]];   short needle (m < k): lower in the bucket of needle.000..,
//   upper in the bucket of needle.111..;   a needle byte that does not occur in the text at position
//   j < k: both bounds in the bucket of prefix . code(byte) . 000.. (code = number of smaller bytes that occur).
// ------------------------------------------------------------------------------------------------
struct Bounds {
  u64 lo_lo, lo_hi;  // the lower bound (first suffix >= pattern) lies in [lo_lo, lo_hi]
  u64 up_lo, up_hi;  // the upper bound (first suffix > pattern that does not start with it) in [up_lo, up_hi]
};

__device__ __forceinline__ Bounds bucket_bounds(const AccelView &ac, const u8 *__restrict__ pat, u32 m, u64 n) {
  Bounds r;
  if (ac.k == 0u) { r.lo_lo = r.up_lo = 0; r.lo_hi = r.up_hi = n; return r; }
  u32 key = 0, w = 0;
  for (u32 j = 0; j < ac.k; ++j) {
    if (j >= m) {  // short needle: every suffix that starts with it has a key in [needle.000.., needle.111..]
      const u32 sh = ac.b * (ac.k - j);
      const u32 klo = key << sh, khi = ((key + 1u) << sh) - 1u;
      r.lo_lo = __ldg(ac.T + klo); r.lo_hi = __ldg(ac.T + klo + 1u);
      r.up_lo = __ldg(ac.T + khi); r.up_hi = __ldg(ac.T + khi + 1u);
      return r;
    }
    if ((j & 3u) == 0u) w = load_bytes(pat + j, min(4u, m - j));
    const u32 v = (w >> (8u * (j & 3u))) & 255u;
    const u32 c = ac.code[v];
    if (!((ac.present[v >> 5] >> (v & 31u)) & 1u)) {
      // No suffix starts with the pattern; every suffix is decided against it at symbol j at the latest.  c is
      // the code of the smallest byte above v that occurs, so the insertion point lies in the bucket of
      // prefix.c.000.. (not always at its start: a suffix that ends inside the prefix is zero padded into the
      // same key when the prefix continues with code-0 symbols, and sorts below the pattern).  v above every
      // byte that occurs: c = 2^b carries into the prefix, and the point is exactly the start of that bucket.
      const u32 kk = ((key << ac.b) + c) << (ac.b * (ac.k - j - 1u));
      r.lo_lo = r.up_lo = __ldg(ac.T + kk);
      r.lo_hi = r.up_hi = (c >> ac.b) ? r.lo_lo : __ldg(ac.T + kk + 1u);
      return r;
    }
    key = (key << ac.b) | c;
  }
  r.lo_lo = r.up_lo = __ldg(ac.T + key);
  r.lo_hi = r.up_hi = __ldg(ac.T + key + 1u);
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
  AccelView ac;
};

// CARRY: start comparisons at min(lmatch, rmatch); pointless when every needle fits one comparison
// step of the group (4 * G bytes).
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

  // ip = number of suffixes below the needle, by binary search inside the needle's bucket.  lm / rm are the
  // common prefix lengths found at ip - 1 and ip once those have been compared (lk / rk): the needle lies
  // between the two, so every suffix in between shares min(lm, rm) bytes with it and comparisons may start
  // there -- the lmatch / rmatch of libdivsufsort's sa_search (utils.c:275-286), which matters for long needles.
  u64 lo = 0, hi = 0;
  if (have) {
    const Bounds b = bucket_bounds(a.ac, pat, m, a.n);
    lo = b.lo_lo;
    hi = b.lo_hi;
  }
  u32 lm = 0, rm = 0;
  bool lk = false, rk = false;
  for (;;) {
    const bool act = have && lo < hi;
    if (!__any_sync(0xffffffffu, act)) break;
    const u64 mid = (lo + hi) >> 1;
    const u64 s = act ? (u64)(u32)__ldg(a.sa + mid) : 0;
    const CmpResult c = group_compare<G>(a.text, s, a.n, pat, m, pw0, (CARRY && lk && rk) ? min(lm, rm) : 0u, act);
    if (act) {
      if (c.gt) { lo = mid + 1; lm = c.cpl; lk = true; } else { hi = mid; rm = c.cpl; rk = true; }
    }
  }
  // final window [a0, a0 + 1], a0 = clamp(ip - 1, 0, n - 2)  (n == 1: the single entry)
  const u64 ip = lo;
  const bool two = have && a.n >= 2;
  u64 a0 = (ip == 0) ? 0 : ip - 1;
  if (two && a0 > a.n - 2) a0 = a.n - 2;
  const bool x_known = lk && a0 + 1 == ip, y_known = two && rk && a0 + 1 == ip;
  const bool x_known2 = rk && a0 == ip;       // ip == 0: the first entry is the one compared as `ip`
  const bool y_known2 = two && lk && a0 + 2 == ip;  // ip == n: the last entry is the one compared as `ip - 1`
  u64 start = have ? (u64)(u32)__ldg(a.sa + a0) : 0;
  u32 len;
  {
    const bool need = have && !(x_known || x_known2);
    const u32 x = group_compare<G>(a.text, start, a.n, pat, m, pw0, 0u, need).cpl;
    len = x_known ? lm : (x_known2 ? rm : x);
  }
  {
    const u64 s1 = two ? (u64)(u32)__ldg(a.sa + a0 + 1) : 0;
    const bool need = two && !(y_known || y_known2);
    const u32 yc = group_compare<G>(a.text, s1, a.n, pat, m, pw0, 0u, need).cpl;
    const u32 y = y_known ? rm : (y_known2 ? lm : yc);
    if (two && !(len > y)) { start = s1; len = y; }  // lib.rs:80-88: `x > y` keeps the first, ties go to the second
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
  AccelView ac;
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
  // left = lower bound: first suffix with r >= 0 in sa_search's comparison (utils.c:244-255), by binary
  // search inside the pattern's bucket; lm / rm as in k_lsm.
  u64 lo = 0, hi = 0, up_lo = 0, up_hi = 0;
  if (have) {
    const Bounds b = bucket_bounds(a.ac, pat, m, a.n);
    lo = b.lo_lo; hi = b.lo_hi; up_lo = b.up_lo; up_hi = b.up_hi;
  }
  u32 lm = 0, rm = 0;
  bool lk = false, rk = false;
  for (;;) {
    const bool act = have && lo < hi;
    if (!__any_sync(0xffffffffu, act)) break;
    const u64 mid = (lo + hi) >> 1;
    const u64 s = act ? (u64)(u32)__ldg(a.sa + mid) : 0;
    const CmpResult c = group_compare<G>(a.text, s, a.n, pat, m, pw0, (CARRY && lk && rk) ? min(lm, rm) : 0u, act);
    if (act) {
      if (c.gt) { lo = mid + 1; lm = c.cpl; lk = true; } else { hi = mid; rm = c.cpl; rk = true; }
    }
  }
  const u64 left = lo;
  // Does the suffix at `left` start with the pattern?  If not there is no occurrence (count 0, and `left` is
  // the insertion point, utils.c:323); only a hit needs the upper bound.
  u32 at_left = rm;
  {
    const bool need = have && !rk && left < a.n;
    const u64 s = need ? (u64)(u32)__ldg(a.sa + left) : 0;
    const u32 c = group_compare<G>(a.text, s, a.n, pat, m, pw0, 0u, need).cpl;
    if (need) at_left = c;
  }
  const bool hit = have && left < a.n && at_left == m;
  // upper bound = first suffix behind `left` that does not start with the pattern: gallop (occurrence
  // counts are small for most patterns), then bisect the last interval
  u64 ulo = left + 1, uhi = (up_hi > left + 1) ? up_hi : left + 1;
  if (hit && up_lo > ulo) ulo = up_lo;  // short pattern: everything below the bucket of pattern.111.. starts with it
  u64 step = 1;
  bool gallop = true;
  for (;;) {
    const bool act = hit && ulo < uhi;
    if (!__any_sync(0xffffffffu, act)) break;
    u64 probe = gallop ? ulo + step - 1 : (ulo + uhi) >> 1;
    if (probe >= uhi) probe = uhi - 1;
    const u64 s = act ? (u64)(u32)__ldg(a.sa + probe) : 0;
    // every suffix in [left, uhi) is >= the pattern and the one at `left` starts with it: comparisons may start at 0 only
    const CmpResult c = group_compare<G>(a.text, s, a.n, pat, m, pw0, 0u, act);
    if (act) {
      if (!c.lt) { ulo = probe + 1; step <<= 1; } else { uhi = probe; gallop = false; }
    }
  }
  if (have && sub == 0) {
    a.left[q] = (i32)left;  // first match, or the insertion point on a miss (utils.c:323)
    a.count[q] = hit ? (i32)(ulo - left) : 0;
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
  LsmArgs a{tv.text, tv.sa, tv.n, tv.text_avail, d_pats, d_pat_off, Q, offset, accumulate, d_io_start, d_io_len, tv.ac};
  const int G = group_lanes(max_pat_len);
  const u64 warps = div_up(Q, 32 / G);
  const unsigned blocks = (unsigned)div_up(warps * 32, 256);
  // G == 8 / 16 are only chosen when every needle fits one comparison step (32 / 64 bytes)
  const bool carry = !getenv("GSA_NO_MATCH_CARRY") && (max_pat_len == 0 || max_pat_len > 128);
  if (G == 8) k_lsm<8, false><<<blocks, 256, 0, st>>>(a);
  else if (G == 16) k_lsm<16, false><<<blocks, 256, 0, st>>>(a);
  else if (carry) k_lsm<32, true><<<blocks, 256, 0, st>>>(a);
  else k_lsm<32, false><<<blocks, 256, 0, st>>>(a);
  GSA_TRY(cudaGetLastError());
  return GSA_OK;
}

int search_all_device(const TextView &tv, const u8 *d_pats, const u64 *d_pat_off, u64 Q, u32 max_pat_len, i32 *d_left,
                      i32 *d_count, cudaStream_t st) {
  if (Q == 0) return GSA_OK;
  SearchAllArgs a{tv.text, tv.sa, tv.n, d_pats, d_pat_off, Q, d_left, d_count, tv.ac};
  const int G = group_lanes(max_pat_len);
  const u64 warps = div_up(Q, 32 / G);
  const unsigned blocks = (unsigned)div_up(warps * 32, 256);
  const bool carry = !getenv("GSA_NO_MATCH_CARRY") && (max_pat_len == 0 || max_pat_len > 128);
  if (G == 8) k_search_all<8, false><<<blocks, 256, 0, st>>>(a);
  else if (G == 16) k_search_all<16, false><<<blocks, 256, 0, st>>>(a);
  else if (carry) k_search_all<32, true><<<blocks, 256, 0, st>>>(a);
  else k_search_all<32, false><<<blocks, 256, 0, st>>>(a);
  GSA_TRY(cudaGetLastError());
  return GSA_OK;
}

// ------------------------------------------------------------------------------------------------
// Construction of the prefix-bucket table from (text, SA).
//   k_accel_presence   which byte values occur (-> dense codes)
//   k_accel_sample     key of every STEP-th suffix-array entry
//   k_accel_mark       F[key] = first suffix-array index with that key: only sample intervals whose two ends
//                      differ are walked entry by entry (keys are non-decreasing along the SA)
//   k_accel_scan_*     T[c] = min over c' >= c of F[c']  (suffix minimum; T[2^(k b)] = n)
// Cost: n / STEP + (#distinct keys) * STEP random text reads and a scan of 2^(k b) words -- a few ms per GiB.
// ------------------------------------------------------------------------------------------------
constexpr u32 ACCEL_STEP = 8;

__global__ void __launch_bounds__(256) k_accel_presence(const u8 *__restrict__ T, u64 n, u32 *__restrict__ present) {
  __shared__ u32 sp[8];
  if (threadIdx.x < 8) sp[threadIdx.x] = 0;
  __syncthreads();
  u32 mine[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const u32 v = T[i];
    mine[v >> 5] |= 1u << (v & 31u);
  }
#pragma unroll
  for (int w = 0; w < 8; ++w) if (mine[w]) atomicOr(&sp[w], mine[w]);
  __syncthreads();
  if (threadIdx.x < 8 && sp[threadIdx.x]) atomicOr(&present[threadIdx.x], sp[threadIdx.x]);
}

__device__ __forceinline__ u32 accel_key(const AccelView &ac, const u8 *__restrict__ text, u64 s, u64 n) {
  u32 key = 0, w = 0;
  const u64 rem = n - s;
  for (u32 j = 0; j < ac.k; ++j) {
    u32 c = 0;
    if ((u64)j < rem) {
      if ((j & 3u) == 0u) w = load_bytes(text + s + j, (u32)min((u64)4, rem - j));
      c = ac.code[(w >> (8u * (j & 3u))) & 255u];
    }
    key = (key << ac.b) | c;
  }
  return key;
}

__global__ void __launch_bounds__(256) k_accel_sample(const AccelView ac, const u8 *__restrict__ text, const i32 *__restrict__ sa,
                                                      u64 n, u64 nsamples, u32 *__restrict__ skey) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsamples) return;
  const u64 j = min(t * ACCEL_STEP, n - 1);  // the last sample is the last entry
  skey[t] = accel_key(ac, text, (u64)(u32)__ldg(sa + j), n);
}

__global__ void __launch_bounds__(256) k_accel_mark(const AccelView ac, const u8 *__restrict__ text, const i32 *__restrict__ sa,
                                                    u64 n, u64 nsamples, const u32 *__restrict__ skey, u32 *__restrict__ F) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsamples) return;
  const u32 k0 = skey[t];
  if (t == 0) F[k0] = 0;  // entry 0 starts the first key
  if (t + 1 >= nsamples) return;
  const u32 k1 = skey[t + 1];
  if (k0 == k1) return;  // no boundary inside (keys are monotone)
  const u64 j0 = t * ACCEL_STEP, j1 = min((t + 1) * ACCEL_STEP, n - 1);
  u32 prev = k0;
  for (u64 j = j0 + 1; j <= j1; ++j) {
    const u32 kj = (j == j1) ? k1 : accel_key(ac, text, (u64)(u32)__ldg(sa + j), n);
    if (kj != prev) { F[kj] = (u32)j; prev = kj; }
  }
}

// suffix-minimum scan in three steps over tiles of 1024 * 8 words
constexpr u32 ASCAN_THREADS = 1024, ASCAN_IPT = 8, ASCAN_TILE = ASCAN_THREADS * ASCAN_IPT;

__global__ void __launch_bounds__(ASCAN_THREADS) k_accel_tile_min(const u32 *__restrict__ F, u64 len, u32 *__restrict__ tmin) {
  __shared__ u32 sw[32];
  const u64 base = (u64)blockIdx.x * ASCAN_TILE;
  u32 m = 0xffffffffu;
  for (u32 i = threadIdx.x; i < ASCAN_TILE; i += ASCAN_THREADS)
    if (base + i < len) m = min(m, F[base + i]);
  for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = sw[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) tmin[blockIdx.x] = m;
  }
}

__global__ void k_accel_tile_scan(u32 *__restrict__ tmin, u32 tiles, u32 n) {  // one thread: <= 2049 tiles
  u32 carry = n;
  for (u32 t = tiles; t > 0; --t) {
    const u32 v = tmin[t - 1];
    tmin[t - 1] = carry;  // minimum over all later tiles
    carry = min(carry, v);
  }
}

__global__ void __launch_bounds__(ASCAN_THREADS) k_accel_apply(u32 *__restrict__ F, u64 len, const u32 *__restrict__ tmin) {
  __shared__ u32 sw[32];
  const u64 base = (u64)blockIdx.x * ASCAN_TILE + (u64)threadIdx.x * ASCAN_IPT;
  u32 v[ASCAN_IPT];
  u32 m = 0xffffffffu;
#pragma unroll
  for (int i = ASCAN_IPT - 1; i >= 0; --i) {
    v[i] = (base + i < len) ? F[base + i] : 0xffffffffu;
    m = min(m, v[i]);
  }
  // exclusive suffix-min over the threads of the block (thread t needs the min of threads > t)
  const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  u32 inc = m;
  for (int o = 1; o < 32; o <<= 1) {
    const u32 y = __shfl_down_sync(0xffffffffu, inc, o);
    if (lane + o < 32) inc = min(inc, y);
  }
  if (lane == 0) sw[warp] = inc;
  u32 ex = __shfl_down_sync(0xffffffffu, inc, 1);
  if (lane == 31) ex = 0xffffffffu;
  __syncthreads();
  for (u32 w = warp + 1; w < 32; ++w) ex = min(ex, sw[w]);
  u32 carry = min(ex, tmin[blockIdx.x]);
#pragma unroll
  for (int i = ASCAN_IPT - 1; i >= 0; --i) {
    carry = min(carry, v[i]);
    if (base + i < len) F[base + i] = carry;
  }
}

int accel_build_device(const u8 *d_T, const i32 *d_SA, u64 n, u32 bits, AccelView *out, u32 **d_table, cudaStream_t st) {
  memset(out, 0, sizeof(*out));
  *d_table = nullptr;
  if (n < 4096 || getenv("GSA_NO_ACCEL")) return GSA_OK;  // k == 0: searches start at [0, n]
  if (const char *e = getenv("GSA_ACCEL_BITS")) bits = (u32)atoi(e);
  if (bits == 0) bits = 26;
  // buckets much finer than the text cannot pay: about 16 suffixes per bucket at least
  bits = std::min<u32>(std::min<u32>(bits, 26), std::max<u32>(8, bits_for(n) > 4 ? bits_for(n) - 4 : 8));
#define ACCEL_MALLOC(ptr, bytes)                                                          \
  do {                                                                                    \
    if (cudaMalloc(&(ptr), (bytes)) != cudaSuccess) {                                     \
      cudaGetLastError();                                                                 \
      set_error("cudaMalloc failed (prefix-bucket table)", __FILE__, __LINE__);           \
      return GSA_ENOMEM;                                                                  \
    }                                                                                     \
  } while (0)
  u32 *d_present = nullptr;
  ACCEL_MALLOC(d_present, 8 * sizeof(u32));
  struct Free { void *p; ~Free() { if (p) cudaFree(p); } } g0{d_present};
  GSA_TRY(cudaMemsetAsync(d_present, 0, 8 * sizeof(u32), st));
  k_accel_presence<<<148 * 8, 256, 0, st>>>(d_T, n, d_present);
  GSA_TRY(cudaGetLastError());
  GSA_TRY(cudaMemcpyAsync(out->present, d_present, 8 * sizeof(u32), cudaMemcpyDeviceToHost, st));
  GSA_TRY(cudaStreamSynchronize(st));
  u32 sigma = 0;
  for (u32 v = 0; v < 256; ++v) {
    out->code[v] = (u8)std::min<u32>(sigma, 255);  // number of smaller bytes that occur (255 only if all 256 occur: then v == 255 occurs too)
    if ((out->present[v >> 5] >> (v & 31)) & 1u) ++sigma;
  }
  // a byte above the largest one that occurs has code sigma, which must be representable next to the real codes
  u32 b = bits_for(sigma > 1 ? sigma - 1 : 1);
  const u32 k = bits / b;
  if (k == 0) return GSA_OK;
  out->b = b;
  out->k = k;
  const u64 len = ((u64)1 << (k * b)) + 1;  // T[0 .. 2^(k b)]
  u32 *F = nullptr;
  ACCEL_MALLOC(F, len * sizeof(u32));
  Free g1{F};
  const u64 nsamples = div_up(n, ACCEL_STEP) + 1;
  u32 *skey = nullptr;
  ACCEL_MALLOC(skey, nsamples * sizeof(u32));
  Free g2{skey};
  const u32 tiles = (u32)div_up(len, ASCAN_TILE);
  u32 *tmin = nullptr;
  ACCEL_MALLOC(tmin, (size_t)tiles * sizeof(u32));
  Free g3{tmin};
  GSA_TRY(cudaMemsetAsync(F, 0xff, len * sizeof(u32), st));
  AccelView dev = *out;  // T is not needed by the key computation
  k_accel_sample<<<(u32)div_up(nsamples, 256), 256, 0, st>>>(dev, d_T, d_SA, n, nsamples, skey);
  GSA_TRY(cudaGetLastError());
  k_accel_mark<<<(u32)div_up(nsamples, 256), 256, 0, st>>>(dev, d_T, d_SA, n, nsamples, skey, F);
  GSA_TRY(cudaGetLastError());
  k_accel_tile_min<<<tiles, ASCAN_THREADS, 0, st>>>(F, len, tmin);
  GSA_TRY(cudaGetLastError());
  k_accel_tile_scan<<<1, 1, 0, st>>>(tmin, tiles, (u32)n);
  GSA_TRY(cudaGetLastError());
  k_accel_apply<<<tiles, ASCAN_THREADS, 0, st>>>(F, len, tmin);
  GSA_TRY(cudaGetLastError());
  GSA_TRY(cudaStreamSynchronize(st));
  out->T = F;
  *d_table = F;
  g1.p = nullptr;  // ownership passes to the caller
  return GSA_OK;
}

int lsm_reduce_device(u64 *d_start, u32 *d_len, u64 Q, u32 nsets, cudaStream_t st) {
  if (Q == 0 || nsets <= 1) return GSA_OK;
  k_lsm_reduce<<<(unsigned)div_up(Q, 256), 256, 0, st>>>(d_start, d_len, Q, nsets);
  GSA_TRY(cudaGetLastError());
  return GSA_OK;
}

}  // namespace gsa
