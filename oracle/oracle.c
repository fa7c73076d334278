/*
 * oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h).
 *
 * Plain-C restatement of the reference's behaviour on the hot path, written from
 * the reference's semantics (cited file:line, relative to /root/reference), not
 * copied from it.  Pinned by tests/test_oracle_*.py against the golden vectors in
 * tests/golden and against the reference's own C library (oracle/_ref).
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* SA construction.  Contract of divsufsort() (divsufsort.c:331-370):         */
/*   -1 on NULL / negative n, n==0 -> 0, n==1 -> SA[0]=0, n==2 by comparing   */
/*   the two bytes (divsufsort.c:346-349); otherwise the sorted suffix order   */
/*   with a proper prefix sorting before the longer suffix.                   */
/* Algorithm here: Manber-Myers prefix doubling with two stable counting-sort */
/* passes per round (second key, then first key).  O(n log maxLCP).           */
/* ------------------------------------------------------------------------- */
int32_t oracle_sa_build(const uint8_t *T, int32_t *SA, int32_t n) {
  if (T == NULL || SA == NULL || n < 0) return -1;
  if (n == 0) return 0;
  if (n == 1) { SA[0] = 0; return 0; }
  if (n == 2) {
    int m = T[0] < T[1];
    SA[m ^ 1] = 0; SA[m] = 1;
    return 0;
  }
  size_t N = (size_t)n;
  int32_t *rk = (int32_t *)malloc(N * sizeof(int32_t));
  int32_t *tmp = (int32_t *)malloc(N * sizeof(int32_t));
  int32_t *sa2 = (int32_t *)malloc(N * sizeof(int32_t));
  int32_t *cnt = (int32_t *)malloc(((N > 256 ? N : 256) + 2) * sizeof(int32_t));
  if (!rk || !tmp || !sa2 || !cnt) { free(rk); free(tmp); free(sa2); free(cnt); return -2; }

  /* depth 1: counting sort on the first byte, dense ranks 1..m (0 = past the end) */
  memset(cnt, 0, 258 * sizeof(int32_t));
  for (size_t i = 0; i < N; ++i) cnt[T[i] + 1]++;
  for (int c = 1; c <= 256; ++c) cnt[c] += cnt[c - 1];
  for (size_t i = 0; i < N; ++i) SA[cnt[T[i]]++] = (int32_t)i;
  int32_t m = 1;
  rk[SA[0]] = 1;
  for (size_t j = 1; j < N; ++j) {
    if (T[SA[j]] != T[SA[j - 1]]) ++m;
    rk[SA[j]] = m;
  }

  for (int64_t h = 1; m < n; h <<= 1) {
    /* stable order by second key rank[i+h] (0 when i+h >= n): the suffixes with no
     * second key first, then i = SA[j]-h in SA order. */
    size_t p = 0;
    for (int64_t i = (n - h > 0 ? n - h : 0); i < n; ++i) sa2[p++] = (int32_t)i;
    for (size_t j = 0; j < N; ++j)
      if (SA[j] >= h) sa2[p++] = (int32_t)(SA[j] - h);
    /* stable counting sort by first key */
    memset(cnt, 0, ((size_t)m + 2) * sizeof(int32_t));
    for (size_t i = 0; i < N; ++i) cnt[rk[i] + 1]++;
    for (int32_t c = 1; c <= m + 1; ++c) cnt[c] += cnt[c - 1];
    for (size_t j = 0; j < N; ++j) { int32_t i = sa2[j]; SA[cnt[rk[i]]++] = i; }
    /* re-rank on the (first, second) pair */
    int32_t mm = 1;
    tmp[SA[0]] = 1;
    for (size_t j = 1; j < N; ++j) {
      int32_t a = SA[j - 1], b = SA[j];
      int32_t a2 = (a + h < n) ? rk[a + h] : 0;
      int32_t b2 = (b + h < n) ? rk[b + h] : 0;
      if (rk[a] != rk[b] || a2 != b2) ++mm;
      tmp[b] = mm;
    }
    int32_t *sw = rk; rk = tmp; tmp = sw;
    m = mm;
  }
  free(rk); free(tmp); free(sa2); free(cnt);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* sacabase::common_prefix_len  (crates/sacabase/src/lib.rs:26-35)            */
/* ------------------------------------------------------------------------- */
size_t oracle_common_prefix_len(const uint8_t *a, size_t alen, const uint8_t *b, size_t blen) {
  size_t n = alen < blen ? alen : blen;
  for (size_t i = 0; i < n; ++i)
    if (a[i] != b[i]) return i;
  return n;
}

/* Rust slice `a > b` on &[u8]: bytewise, a proper prefix is the smaller one. */
static int slice_gt(const uint8_t *a, size_t alen, const uint8_t *b, size_t blen) {
  size_t n = alen < blen ? alen : blen;
  int c = n ? memcmp(a, b, n) : 0;
  if (c != 0) return c > 0;
  return alen > blen;
}
static int slice_lt(const uint8_t *a, size_t alen, const uint8_t *b, size_t blen) {
  return slice_gt(b, blen, a, alen);
}

/* ------------------------------------------------------------------------- */
/* sacabase::longest_substring_match  (crates/sacabase/src/lib.rs:39-99)      */
/*   loop on the shrinking window `sa`:                                       */
/*     len 1 -> (sa[0], cpl)                                         :77-79   */
/*     len 2 -> x,y = cpl of both; x > y ? first : second            :80-88   */
/*     else  -> mid = len/2; needle > suff(mid) ? [mid..] : [..=mid] :89-96   */
/* ------------------------------------------------------------------------- */
int32_t oracle_longest_substring_match(const uint8_t *T, size_t n, const int32_t *SA, size_t sa_len,
                                       const uint8_t *needle, size_t m, uint64_t *start, uint64_t *len) {
  if (sa_len == 0) return -1; /* Rust: index out of bounds panic at sa[0] */
  const int32_t *sa = SA;
  size_t w = sa_len;
  for (;;) {
    if (w == 1) {
      size_t s = (size_t)sa[0];
      *start = s;
      *len = oracle_common_prefix_len(T + s, n - s, needle, m);
      return 0;
    } else if (w == 2) {
      size_t s0 = (size_t)sa[0], s1 = (size_t)sa[1];
      size_t x = oracle_common_prefix_len(T + s0, n - s0, needle, m);
      size_t y = oracle_common_prefix_len(T + s1, n - s1, needle, m);
      if (x > y) { *start = s0; *len = x; } else { *start = s1; *len = y; }
      return 0;
    } else {
      size_t mid = w / 2;
      size_t s = (size_t)sa[mid];
      if (slice_gt(needle, m, T + s, n - s)) { sa += mid; w -= mid; } else { w = mid + 1; }
    }
  }
}

/* ------------------------------------------------------------------------- */
/* sacabase::verify  (crates/sacabase/src/lib.rs:127-149)                     */
/* ------------------------------------------------------------------------- */
int32_t oracle_verify(const uint8_t *T, size_t n, const int32_t *SA, uint64_t *bad_i) {
  if (n == 0) return -1; /* `input.len() - 1` underflows (lib.rs:143) */
  for (size_t i = 0; i + 1 < n; ++i) {
    size_t a = (size_t)SA[i], b = (size_t)SA[i + 1];
    if (!slice_lt(T + a, n - a, T + b, n - b)) {
      if (bad_i) *bad_i = i;
      return 1;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* libdivsufsort sa_search  (utils.c:244-325)                                 */
/* _compare (:244-255): r = sign of first differing byte (text - pattern); if  */
/* the suffix ends before the pattern does r = -1; if the pattern is consumed  */
/* r = 0.  The lmatch/rmatch carry of the C code is a pure speed-up and is not */
/* reproduced; the half-interval walk is.                                      */
/* ------------------------------------------------------------------------- */
static int sa_compare(const uint8_t *T, int32_t Tsize, const uint8_t *P, int32_t Psize, int32_t suf) {
  int32_t i = suf, j = 0;
  while (i < Tsize && j < Psize) {
    int r = (int)T[i] - (int)P[j];
    if (r != 0) return r;
    ++i; ++j;
  }
  return (j != Psize) ? -1 : 0;
}

int32_t oracle_sa_search(const uint8_t *T, int32_t Tsize, const uint8_t *P, int32_t Psize,
                         const int32_t *SA, int32_t SAsize, int32_t *idx) {
  if (idx) *idx = -1;
  if (!T || !P || !SA || Tsize < 0 || Psize < 0 || SAsize < 0) return -1; /* :270-271 */
  if (Tsize == 0 || SAsize == 0) return 0;                                /* :272 */
  if (Psize == 0) { if (idx) *idx = 0; return SAsize; }                   /* :273 */
  int32_t i = 0, j = 0, k = 0, size, half;
  for (size = SAsize, half = size >> 1; 0 < size; size = half, half >>= 1) {
    int r = sa_compare(T, Tsize, P, Psize, SA[i + half]);
    if (r < 0) {
      i += half + 1;
      half -= (size & 1) ^ 1;
    } else if (r == 0) {
      int32_t lsize = half, rsize = size - half - 1;
      j = i; k = i + half + 1;
      for (half = lsize >> 1; 0 < lsize; lsize = half, half >>= 1) { /* left edge :290-301 */
        r = sa_compare(T, Tsize, P, Psize, SA[j + half]);
        if (r < 0) { j += half + 1; half -= (lsize & 1) ^ 1; }
      }
      for (half = rsize >> 1; 0 < rsize; rsize = half, half >>= 1) { /* right edge :304-316 */
        r = sa_compare(T, Tsize, P, Psize, SA[k + half]);
        if (r <= 0) { k += half + 1; half -= (rsize & 1) ^ 1; }
      }
      break;
    }
  }
  if (idx) *idx = (0 < (k - j)) ? j : i; /* :323 */
  return k - j;
}

/* ------------------------------------------------------------------------- */
/* libdivsufsort sufcheck, non-verbose  (utils.c:160-241)                     */
/* ------------------------------------------------------------------------- */
int32_t oracle_lcp_kasai(const uint8_t *T, const int32_t *SA, int32_t n, int32_t *LCP) {
  if (n < 0 || (n > 0 && (!T || !SA || !LCP))) return -1;
  if (n == 0) return 0;
  int32_t *isa = (int32_t *)malloc((size_t)n * sizeof(int32_t));
  if (!isa) return -1;
  for (int32_t j = 0; j < n; ++j) isa[SA[j]] = j;
  int32_t l = 0;
  for (int32_t i = 0; i < n; ++i) {
    const int32_t j = isa[i];
    if (j == 0) { LCP[0] = 0; l = 0; continue; }
    const int32_t p = SA[j - 1];
    while (i + l < n && p + l < n && T[i + l] == T[p + l]) ++l;
    LCP[j] = l;
    if (l > 0) --l;
  }
  free(isa);
  return 0;
}

int32_t oracle_sufcheck(const uint8_t *T, const int32_t *SA, int32_t n) {
  if (!T || !SA || n < 0) return -1;
  if (n == 0) return 0;
  for (int32_t i = 0; i < n; ++i)
    if (SA[i] < 0 || n <= SA[i]) return -2;
  for (int32_t i = 1; i < n; ++i)
    if (T[SA[i - 1]] > T[SA[i]]) return -3;
  int32_t C[256];
  memset(C, 0, sizeof C);
  for (int32_t i = 0; i < n; ++i) ++C[T[i]];
  for (int32_t i = 0, p = 0; i < 256; ++i) { int32_t t = C[i]; C[i] = p; p += t; }
  int32_t q = C[T[n - 1]];
  C[T[n - 1]] += 1;
  for (int32_t i = 0; i < n; ++i) {
    int32_t p = SA[i], t, c;
    if (0 < p) { c = T[--p]; t = C[c]; } else { c = T[p = n - 1]; t = q; }
    if (t < 0 || p != SA[t]) return -4;
    if (t != q) {
      ++C[c];
      if (n <= C[c] || T[SA[C[c]]] != c) C[c] = -1;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* sacapart  (crates/sacapart/src/lib.rs)                                     */
/* ------------------------------------------------------------------------- */
int32_t oracle_part_plan(uint64_t n, uint64_t num_partitions, uint64_t *partition_size,
                         uint64_t *actual_partitions) {
  if (num_partitions == 0) return -1;                 /* lib.rs:43 divides by zero */
  uint64_t ps = n / num_partitions + 1;               /* lib.rs:43 */
  if (partition_size) *partition_size = ps;
  if (actual_partitions) *actual_partitions = (n + ps - 1) / ps; /* par_chunks: lib.rs:45-49,60-62 */
  return 0;
}

int32_t oracle_part_lsm(const uint8_t *T, uint64_t n, uint64_t partition_size, uint64_t nparts,
                        const int32_t *const *SAs, const uint8_t *needle, size_t m, uint64_t *start,
                        uint64_t *len) {
  int have = 0;
  uint64_t best_start = 0, best_len = 0;
  for (uint64_t i = 0; i < nparts; ++i) {                       /* lib.rs:71 */
    uint64_t offset = i * partition_size;                       /* :73 */
    uint64_t clen = (offset + partition_size <= n) ? partition_size : n - offset;
    uint64_t s, l;
    if (oracle_longest_substring_match(T + offset, clen, SAs[i], clen, needle, m, &s, &l) != 0) return -1;
    int may_extend = (s + l == clen);                           /* :77 */
    s += offset;                                                /* :80 */
    if (may_extend) l = oracle_common_prefix_len(T + s, n - s, needle, m); /* :82-84 */
    if (!have || l > best_len) { have = 1; best_start = s; best_len = l; } /* :86-92 */
  }
  if (!have) return -1;                                         /* :94-96 expect() */
  *start = best_start; *len = best_len;
  return 0;
}

/* ------------------------------------------------------------------------- */
/* Batched forms: the "rayon CPU" baseline of BASELINE.md section 3.            */
/* ------------------------------------------------------------------------- */
int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int32_t oracle_lsm_batch(const uint8_t *T, size_t n, const int32_t *SA, size_t sa_len,
                         const uint8_t *pats, const uint64_t *pat_off, uint64_t Q, uint64_t *out_start,
                         uint32_t *out_len, int threads) {
  if (sa_len == 0) return -1;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#endif
  (void)threads;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(threads)
  for (int64_t q = 0; q < (int64_t)Q; ++q) {
    uint64_t s = 0, l = 0;
    oracle_longest_substring_match(T, n, SA, sa_len, pats + pat_off[q], pat_off[q + 1] - pat_off[q], &s, &l);
    out_start[q] = s;
    out_len[q] = (uint32_t)l;
  }
  return 0;
}

int32_t oracle_search_all_batch(const uint8_t *T, int32_t n, const int32_t *SA, const uint8_t *pats,
                                const uint64_t *pat_off, uint64_t Q, int32_t *out_left,
                                int32_t *out_count, int threads) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#endif
  (void)threads;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(threads)
  for (int64_t q = 0; q < (int64_t)Q; ++q) {
    int32_t idx = -1;
    out_count[q] = oracle_sa_search(T, n, pats + pat_off[q], (int32_t)(pat_off[q + 1] - pat_off[q]), SA, n, &idx);
    out_left[q] = idx;
  }
  return 0;
}

int32_t oracle_part_lsm_batch(const uint8_t *T, uint64_t n, uint64_t partition_size, uint64_t nparts,
                              const int32_t *const *SAs, const uint8_t *pats, const uint64_t *pat_off,
                              uint64_t Q, uint64_t *out_start, uint32_t *out_len, int threads) {
  if (nparts == 0) return -1;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#endif
  (void)threads;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(threads)
  for (int64_t q = 0; q < (int64_t)Q; ++q) {
    uint64_t s = 0, l = 0;
    oracle_part_lsm(T, n, partition_size, nparts, SAs, pats + pat_off[q], pat_off[q + 1] - pat_off[q], &s, &l);
    out_start[q] = s;
    out_len[q] = (uint32_t)l;
  }
  return 0;
}
