/*
 * oracle.h -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  Nothing under stringsearch_b200/ links,
 * imports or calls it; the product path is CUDA-only and fails loudly without
 * its extension.
 *
 * What is restated here (file:line relative to /root/reference):
 *   oracle_common_prefix_len        crates/sacabase/src/lib.rs:26-35
 *   oracle_longest_substring_match  crates/sacabase/src/lib.rs:39-99
 *   oracle_verify                   crates/sacabase/src/lib.rs:127-149
 *   oracle_part_plan / _part_lsm    crates/sacapart/src/lib.rs:39-58, 69-97
 *   oracle_sa_search                crates/cdivsufsort/c-sources/utils.c:244-325
 *   oracle_sufcheck                 crates/cdivsufsort/c-sources/utils.c:160-241
 *   oracle_sa_build                 the *result* of divsufsort::sort
 *                                   (crates/divsufsort/src/lib.rs:20-29,
 *                                   c-sources/divsufsort.c:331-370).  The suffix
 *                                   array of a text is a unique permutation, so the
 *                                   oracle does not restate induced sorting; it
 *                                   uses textbook Manber-Myers doubling and is
 *                                   PINNED against (a) every golden vector in
 *                                   tests/golden and (b) the reference's own C
 *                                   libdivsufsort compiled into oracle/_ref.
 *
 * Parity status: PINNED for SA build, longest_substring_match, partitioned
 * search (reference's own tests: crates/sacapart/src/lib.rs:105-165,
 * crates/divsufsort/src/lib.rs:33-86) and sa_search (vs oracle/_ref sa_search).
 * `contains`/`search_all` exist only in the un-vendored third-party crate
 * suffix_array 0.4.0 (Cargo.lock:460-468): for those names parity is pinned to
 * the in-tree C sa_search semantics, not to that crate ("parity unpinned" w.r.t.
 * suffix_array 0.4.0 itself).
 */
#ifndef GSA_ORACLE_H
#define GSA_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* SA construction: same contract and return codes as C divsufsort()
 * (divsufsort.c:346-349): 0 ok, -1 bad args, -2 allocation failure. */
int32_t oracle_sa_build(const uint8_t *T, int32_t *SA, int32_t n);

size_t oracle_common_prefix_len(const uint8_t *a, size_t alen, const uint8_t *b, size_t blen);

/* sacabase::longest_substring_match.  Returns 0, or -1 where the Rust code would
 * panic (empty suffix array). */
int32_t oracle_longest_substring_match(const uint8_t *T, size_t n, const int32_t *SA, size_t sa_len,
                                       const uint8_t *needle, size_t m, uint64_t *start, uint64_t *len);

/* sacabase::verify: 0 = sorted; 1 = NotSorted (bad_i receives i); -1 = the Rust
 * code would underflow (empty input). */
int32_t oracle_verify(const uint8_t *T, size_t n, const int32_t *SA, uint64_t *bad_i);

/* libdivsufsort sa_search: returns count, *idx = left (see utils.c:269-273,323-324). */
int32_t oracle_sa_search(const uint8_t *T, int32_t Tsize, const uint8_t *P, int32_t Psize,
                         const int32_t *SA, int32_t SAsize, int32_t *idx);

/* libdivsufsort sufcheck (non-verbose): 0 ok, -1..-4 as utils.c:160-241. */
int32_t oracle_sufcheck(const uint8_t *T, const int32_t *SA, int32_t n);

/* LCP array by Kasai et al. (CPM 2001): LCP[0] = 0, LCP[j] = lcp(suffix SA[j-1], suffix SA[j]).
 * No counterpart in the reference (SURVEY.md 8(f) rank 3 extension); pinned in tests/ against a
 * byte-by-byte comparison of adjacent suffixes.  Returns 0, -1 on bad arguments / allocation. */
int32_t oracle_lcp_kasai(const uint8_t *T, const int32_t *SA, int32_t n, int32_t *LCP);

/* sacapart chunking (lib.rs:43-51,60-62): partition_size = n / P + 1, number of
 * chunks actually produced by par_chunks.  Returns -1 if P == 0 (division by zero
 * in the reference). */
int32_t oracle_part_plan(uint64_t n, uint64_t num_partitions, uint64_t *partition_size,
                         uint64_t *actual_partitions);

/* sacapart query (lib.rs:69-97).  SAs[i] is the SA of chunk i.  Returns 0, or -1
 * where the reference panics (zero partitions). */
int32_t oracle_part_lsm(const uint8_t *T, uint64_t n, uint64_t partition_size, uint64_t nparts,
                        const int32_t *const *SAs, const uint8_t *needle, size_t m, uint64_t *start,
                        uint64_t *len);

/* Batched forms (OpenMP over patterns) used as the "rayon CPU" search baseline.
 * pat_off has Q+1 entries. threads <= 0 -> OpenMP default. */
int32_t oracle_lsm_batch(const uint8_t *T, size_t n, const int32_t *SA, size_t sa_len,
                         const uint8_t *pats, const uint64_t *pat_off, uint64_t Q, uint64_t *out_start,
                         uint32_t *out_len, int threads);
int32_t oracle_search_all_batch(const uint8_t *T, int32_t n, const int32_t *SA, const uint8_t *pats,
                                const uint64_t *pat_off, uint64_t Q, int32_t *out_left,
                                int32_t *out_count, int threads);
int32_t oracle_part_lsm_batch(const uint8_t *T, uint64_t n, uint64_t partition_size, uint64_t nparts,
                              const int32_t *const *SAs, const uint8_t *pats, const uint64_t *pat_off,
                              uint64_t Q, uint64_t *out_start, uint32_t *out_len, int threads);
int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
